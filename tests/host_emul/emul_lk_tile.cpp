// lk_tile_kernel<SRC, WIN> (every window, frames / gradients / in-tile warp) and iter_finalize_kernel run on the CPU
// from their own source (optical-flow-fpga_b200/csrc/lk_tile.cu) on top of cuda_on_host.h.  TEST INFRASTRUCTURE;
// tests/test_kernel_host_emulation.py builds and drives it.
#include "cuda_on_host.h"

#include "lk_tile.cu"

using namespace ofb;

template <int SRC, int WIN>
static void run_one(const TileArgs& a, int batch) {
    constexpr int HW = WIN / 2, GW = TX + 2 * HW, GH = TY + 2 * HW, FW = GW + 2, FH = GH + 2;  // as launch_one
    std::vector<float> smem(3 * GH * GW + 2 * FH * FW);
    cuda_on_host::dynamic_smem() = smem.data();
    const int rows = (SRC == SRC_WARP || SRC == SRC_WARPED) ? a.row_hi - a.row_lo : a.H;
    dim3 grid((a.W + TX - 1) / TX, (rows + TY - 1) / TY, batch);
    cuda_on_host::launch(grid, TILE_THREADS, [&a]() { lk_tile_kernel<SRC, WIN>(a); });
    cuda_on_host::dynamic_smem() = nullptr;
}

template <int SRC>
static int run_src(int window, const TileArgs& a, int batch) {
    switch (window) {
        case 1: run_one<SRC, 1>(a, batch); return 0;
        case 3: run_one<SRC, 3>(a, batch); return 0;
        case 5: run_one<SRC, 5>(a, batch); return 0;
        case 7: run_one<SRC, 7>(a, batch); return 0;
        case 9: run_one<SRC, 9>(a, batch); return 0;
        case 11: run_one<SRC, 11>(a, batch); return 0;
        default: return 1;
    }
}

extern "C" {

// src: 0 frames (in0 = prev, in1 = curr), 2 gradients (in0..2 = Ix, Iy, It); -> out_u, out_v
int emul_lk_tile(int src, int window, const float* in0, const float* in1, const float* in2, float* u, float* v, int batch,
                 int H, int W) {
    TileArgs a;
    std::memset(&a, 0, sizeof(a));
    a.in0 = in0;
    a.in1 = in1;
    a.in2 = in2;
    a.out_u = u;
    a.out_v = v;
    a.H = H;
    a.W = W;
    if (src == SRC_FRAMES) return run_src<SRC_FRAMES>(window, a, batch);
    if (src == SRC_GRADS) return run_src<SRC_GRADS>(window, a, batch);
    return 1;
}

// src: 1 = the kernel gathers curr through the flow itself, 3 = in1 is the warped plane
int emul_lk_tile_refine(int src, int window, const float* prev, const float* in1, float* flow_u0, float* flow_v0,
                        float* flow_u1, float* flow_v1, const int* sel, int sel_xor, const int* done, double* partial,
                        int batch, int H, int W, int row_lo, int row_hi, int own_lo, int own_hi) {
    TileArgs a;
    std::memset(&a, 0, sizeof(a));
    a.in0 = prev;
    a.in1 = in1;
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
    if (src == SRC_WARP) return run_src<SRC_WARP>(window, a, batch);
    if (src == SRC_WARPED) return run_src<SRC_WARPED>(window, a, batch);
    return 1;
}

// the iteration's convergence step: means of the per-block sums, early-exit test, ping-pong flip
int emul_iter_finalize(const double* partial, int blocks_per_pair, int H, int W, int* sel, int* done, int* iters_executed,
                       int iters_pair_stride, float* residuals, long resid_pair_stride, int iteration, int batch) {
    IterFinalizeArgs f;
    std::memset(&f, 0, sizeof(f));
    f.partial = partial;
    f.blocks_per_pair = blocks_per_pair;
    f.H = H;
    f.W = W;
    f.sel = sel;
    f.done = done;
    f.iters_executed = iters_executed;
    f.iters_pair_stride = iters_pair_stride;
    f.residuals = residuals;
    f.resid_pair_stride = (size_t)resid_pair_stride;
    f.iteration = iteration;
    cuda_on_host::launch(dim3(batch, 1, 1), 256, [&f]() { iter_finalize_kernel(f); });
    return 0;
}
}
