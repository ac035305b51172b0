// warp_rows_kernel<float> / <double> -- warp_image of the split refinement iteration -- run on the CPU from its
// own source (optical-flow-fpga_b200/csrc/warp_rows.cuh) on top of cuda_on_host.h.  TEST INFRASTRUCTURE;
// tests/test_kernel_host_emulation.py builds and drives it.
#include "cuda_on_host.h"

#include "warp_rows.cuh"

using namespace ofb;

extern "C" int emul_warp_rows(const float* curr, const float* flow_u, const float* flow_v, float* warped, int batch, int H,
                              int W, int row_lo, int row_hi, int exact) {
    WarpRowsArgs w;
    std::memset(&w, 0, sizeof(w));
    w.curr = curr;
    w.flow_u[0] = flow_u;
    w.flow_v[0] = flow_v;
    w.flow_u[1] = flow_u;
    w.flow_v[1] = flow_v;
    w.warped = warped;
    w.H = H;
    w.W = W;
    w.row_lo = row_lo;
    w.row_hi = row_hi;
    dim3 grid((W + 256 * WR_PER_THREAD - 1) / (256 * WR_PER_THREAD), row_hi - row_lo, batch);  // as launch_warp_rows
    if (exact)
        cuda_on_host::launch(grid, 256, [&w]() { warp_rows_kernel<double>(w); });
    else
        cuda_on_host::launch(grid, 256, [&w]() { warp_rows_kernel<float>(w); });
    return 0;
}
