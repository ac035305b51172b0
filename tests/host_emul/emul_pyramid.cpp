// pyramid.cu (gradients, tile pyramid kernel, warp_kernel, both upsample kernels, select_copy) run on the CPU from
// its own source INCLUDING its launchers (OF_LAUNCH); emul_pyramid_march.cpp adds pyramid_march.cu, which
// launch_pyramid_down dispatches to.  TEST INFRASTRUCTURE; tests/test_kernel_host_emulation.py drives the launchers.
#include "cuda_on_host.h"

#include "pyramid.cu"

using namespace ofb;

extern "C" {
int emul_gradients(const float* prev, const float* curr, float* ix, float* iy, float* it, int batch, int H, int W) {
    return (int)launch_gradients(prev, curr, ix, iy, it, batch, H, W, nullptr, nullptr);
}
int emul_pyramid_down(const float* src, float* dst, int batch, int H, int W, int oh, int ow, const double* weights,
                      int radius, int row_lo, int row_hi, int fast) {
    return (int)launch_pyramid_down(src, dst, batch, H, W, oh, ow, weights, radius, row_lo, row_hi, nullptr, nullptr, fast != 0);
}
int emul_warp(const float* img, const float* fu, const float* fv, float* out, int batch, int H, int W) {
    return (int)launch_warp(img, fu, fv, out, batch, H, W, nullptr, nullptr);
}
int emul_upsample_flow(const float* cu, const float* cv, float* fu, float* fv, int batch, int ch, int cw, int th, int tw,
                       int row_lo, int row_hi) {
    return (int)launch_upsample_flow(cu, cv, nullptr, nullptr, nullptr, 0, fu, fv, batch, ch, cw, th, tw, row_lo, row_hi,
                                     nullptr, nullptr);
}
}
