// lk_tile5_kernel -- the default exact-mode kernel for window 5 -- run on the CPU from its own source file:
// optical-flow-fpga_b200/csrc/lk_tile5.cu is #included below and compiled by g++ on top of cuda_on_host.h (CUDA
// threads = OS threads).  TEST INFRASTRUCTURE; tests/test_kernel_host_emulation.py builds and drives it.
//   g++ -O1 -ffp-contract=off -std=c++17 -shared -fPIC -pthread -DOF_HOST_EMULATION \
//       -I optical-flow-fpga_b200/csrc -I /usr/local/cuda/include tests/host_emul/emul_lk_tile5.cpp -o ...so
#include "cuda_on_host.h"

#include "lk_tile5.cu"

using namespace ofb;

template <int SRC>
static void run(const TileArgs& a, int batch) {
    const int rows = (SRC == SRC_WARPED) ? a.row_hi - a.row_lo : a.H;
    dim3 grid((a.W + T5_TX - 1) / T5_TX, (rows + T5_TY - 1) / T5_TY, batch);
    cuda_on_host::launch(grid, T5_THREADS, [&a]() { lk_tile5_kernel<SRC>(a); });
}

extern "C" {

// single-scale LK on [batch][H][W] frames (SRC_FRAMES)
int emul_lk_tile5_frames(const float* prev, const float* curr, float* u, float* v, int batch, int H, int W) {
    TileArgs a;
    std::memset(&a, 0, sizeof(a));
    a.in0 = prev;
    a.in1 = curr;
    a.out_u = u;
    a.out_v = v;
    a.H = H;
    a.W = W;
    run<SRC_FRAMES>(a, batch);
    return 0;
}

// one refinement iteration on (prev, warped) (SRC_WARPED): flow[sel ^ 1] = flow[sel] + d on rows [row_lo, row_hi),
// per-block sums of |du|, |dv| over rows [own_lo, own_hi) -> partial[pair][block][2]; pairs with done[pair] skipped
int emul_lk_tile5_warped(const float* prev, const float* warped, float* flow_u0, float* flow_v0, float* flow_u1,
                         float* flow_v1, const int* sel, int sel_xor, const int* done, double* partial, int batch, int H,
                         int W, int row_lo, int row_hi, int own_lo, int own_hi) {
    TileArgs a;
    std::memset(&a, 0, sizeof(a));
    a.in0 = prev;
    a.in1 = warped;
    a.flow_u[0] = flow_u0;
    a.flow_v[0] = flow_v0;
    a.flow_u[1] = flow_u1;
    a.flow_v[1] = flow_v1;
    a.sel = sel;
    a.sel_xor = sel_xor;
    a.done = done;
    a.partial = partial;
    a.H = H;
    a.W = W;
    a.row_lo = row_lo;
    a.row_hi = row_hi;
    a.own_lo = own_lo;
    a.own_hi = own_hi;
    run<SRC_WARPED>(a, batch);
    return 0;
}
}
