// lk_exact_march_kernel -- the default exact-mode kernel for window 5 -- run on the CPU from its own source file
// through its own launcher (band planning, dynamic shared memory as on the device): lk_exact_march.cu is #included
// below and compiled by g++ on top of cuda_on_host.h (CUDA threads = OS threads, __syncwarp a barrier, packed pairs
// two IEEE operations).  TEST INFRASTRUCTURE; tests/test_kernel_host_emulation.py builds and drives it.
#include "cuda_on_host.h"
#include "peer_on_host.h"  // the fused iteration tail (peer_device.cuh)

#include "lk_exact_march.cu"

using namespace ofb;

#ifndef EMUL_EXACT_MARCH_NO_TILE_GEOMETRY
// the partial-sum slot count both exact kernels share (lk_tile.cu's definition; that file is not part of this unit)
namespace ofb {
int lk_tile_blocks_per_pair(int rows, int W) { return ((W + 63) / 64) * ((rows + 15) / 16); }
}  // namespace ofb
#endif

extern "C" {

// single-scale LK on [batch][H][W] frames (SRC_FRAMES)
int emul_lk_exact_march_frames(const float* prev, const float* curr, float* u, float* v, int batch, int H, int W) {
    TileArgs a;
    std::memset(&a, 0, sizeof(a));
    a.in0 = prev;
    a.in1 = curr;
    a.out_u = u;
    a.out_v = v;
    a.H = H;
    a.W = W;
    return (int)launch_lk_exact_march(SRC_FRAMES, a, batch, nullptr);
}

// one refinement iteration on (prev, warped) (SRC_WARPED): flow[sel ^ 1] = flow[sel] + d on rows [row_lo, row_hi),
// per-unit sums of |du|, |dv| over rows [own_lo, own_hi) -> partial[pair][slot][2]; pairs with done[pair] skipped
int emul_lk_exact_march_warped(const float* prev, const float* warped, float* flow_u0, float* flow_v0, float* flow_u1,
                               float* flow_v1, const int* sel, int sel_xor, const int* done, double* partial, int batch,
                               int H, int W, int row_lo, int row_hi, int own_lo, int own_hi) {
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
    return (int)launch_lk_exact_march(SRC_WARPED, a, batch, nullptr);
}
}
