// apply_motion (python/generate_test_frames_natural.py:67-73): the fixture generators' sub-pixel
// shift, scipy.ndimage.shift(frame, (dy, dx), order=1, mode="constant", cval) on uint8 frames, for a
// batch of frames with one (dx, dy) each -- the step that produces the second frame of a synthetic
// pair, so benchmark / test batches can be built without leaving the device.
//
// SciPy's arithmetic (ni_interpolation.c, NI_ZoomShift), mirrored bit for bit: coordinate =
// index - shift in float64; outside [0, n-1] on either axis -> cval; else the four taps blended in
// float64 in row-major order, each as (value * wy) * wx, summed from 0.0; uint8 result =
// truncate(clamp(t + 0.5, 0, 255)).
#include <cuda_runtime.h>

#include "of_common.cuh"
#include "of_kernels.h"

namespace ofb {

namespace {

__global__ void __launch_bounds__(256) apply_motion_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           const double* __restrict__ dx, const double* __restrict__ dy,
                                                           int H, int W, double cval) {
    const int pair = blockIdx.z;
    const int y = blockIdx.y;
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    const uint8_t* __restrict__ img = src + (size_t)pair * H * W;
    const double Y = dsub((double)y, dy[pair]), X = dsub((double)x, dx[pair]);
    double t = cval;
    if (Y >= 0.0 && Y <= (double)(H - 1) && X >= 0.0 && X <= (double)(W - 1)) {
        const double fy0 = floor(Y), fx0 = floor(X);
        const double fy = dsub(Y, fy0), fx = dsub(X, fx0);
        const int y0 = (int)fy0, x0 = (int)fx0;
        const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);  // weight exactly 0 where it would leave the frame
        const double wy0 = dsub(1.0, fy), wx0 = dsub(1.0, fx);
        const uint8_t* r0 = img + (size_t)y0 * W;
        const uint8_t* r1 = img + (size_t)y1 * W;
        t = 0.0;
        t = dadd(t, dmul(dmul((double)r0[x0], wy0), wx0));
        t = dadd(t, dmul(dmul((double)r0[x1], wy0), fx));
        t = dadd(t, dmul(dmul((double)r1[x0], fy), wx0));
        t = dadd(t, dmul(dmul((double)r1[x1], fy), fx));
    }
    double v = t > 0.0 ? dadd(t, 0.5) : 0.0;
    v = v > 255.0 ? 255.0 : v;
    dst[(size_t)pair * H * W + (size_t)y * W + x] = (uint8_t)(int)v;
}

// apply_motion_opencv (python/generate_test_suite.py:165-204) = cv2.warpAffine(INTER_LINEAR,
// BORDER_CONSTANT): OpenCV's fixed-point bilinear warp (imgwarp.cpp), bit for bit.  The inverse map is
// sampled with 10 fractional bits, rounded to 1/32 pixel, blended with 15-bit integer weights:
//   X = (rint((M1 y + M2) 1024) + 16 + rint(M0 x 1024)) >> 5;  sx = X >> 5, ax = X & 31   (Y alike)
//   out = (32 * sum(tap * wy * wx) + 2^14) >> 15,  taps outside the frame = cval
// minv holds warpAffine's inverted matrices, computed on the host in OpenCV's operation order.
constexpr int WA_MAX_BATCH = 48;
struct WarpAffineArgs {
    const uint8_t* src;
    uint8_t* dst;
    int H, W, cval;
    double minv[WA_MAX_BATCH][6];
};

__global__ void __launch_bounds__(256) warp_affine_kernel(const WarpAffineArgs a) {
    const int pair = blockIdx.z, y = blockIdx.y;
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= a.W) return;
    const int H = a.H, W = a.W;
    const double* M = a.minv[pair];
    const uint8_t* __restrict__ img = a.src + (size_t)pair * H * W;
    // saturate_cast<int>(double) = round half to even; products and sums rounded separately (no FMA)
    const long long adelta = __double2ll_rn(dmul(dmul(M[0], (double)x), 1024.0));
    const long long bdelta = __double2ll_rn(dmul(dmul(M[3], (double)x), 1024.0));
    const long long X0 = __double2ll_rn(dmul(dadd(dmul(M[1], (double)y), M[2]), 1024.0)) + 16;
    const long long Y0 = __double2ll_rn(dmul(dadd(dmul(M[4], (double)y), M[5]), 1024.0)) + 16;
    const long long X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    const long long sxl = X >> 5, syl = Y >> 5;
    const int ax = (int)(X & 31), ay = (int)(Y & 31);
    int s = a.cval << 10;  // all four taps outside: cval * 1024
    if (sxl >= -1 && sxl < W && syl >= -1 && syl < H) {
        const int sx = (int)sxl, sy = (int)syl;
        const bool x0in = sx >= 0, x1in = sx + 1 < W, y0in = sy >= 0, y1in = sy + 1 < H;
        const uint8_t* r0 = img + (size_t)(y0in ? sy : 0) * W;
        const uint8_t* r1 = img + (size_t)(y1in ? sy + 1 : 0) * W;
        const int t00 = (y0in && x0in) ? r0[sx] : a.cval, t01 = (y0in && x1in) ? r0[sx + 1] : a.cval;
        const int t10 = (y1in && x0in) ? r1[sx] : a.cval, t11 = (y1in && x1in) ? r1[sx + 1] : a.cval;
        s = t00 * ((32 - ay) * (32 - ax)) + t01 * ((32 - ay) * ax) + t10 * (ay * (32 - ax)) + t11 * (ay * ax);
    }
    a.dst[(size_t)pair * H * W + (size_t)y * W + x] = (uint8_t)((s * 32 + (1 << 14)) >> 15);
}

}  // namespace

cudaError_t launch_warp_affine(const uint8_t* src, uint8_t* dst, const double* minv, int batch, int H, int W, int cval,
                               int* launches, cudaStream_t stream) {
    if (batch < 1 || H < 1 || H > 65535 || W < 1 || cval < 0 || cval > 255) return cudaErrorInvalidValue;
    for (int b0 = 0; b0 < batch; b0 += WA_MAX_BATCH) {
        const int nb = batch - b0 < WA_MAX_BATCH ? batch - b0 : WA_MAX_BATCH;
        WarpAffineArgs a;
        a.src = src + (size_t)b0 * H * W;
        a.dst = dst + (size_t)b0 * H * W;
        a.H = H;
        a.W = W;
        a.cval = cval;
        for (int b = 0; b < nb; ++b)
            for (int k = 0; k < 6; ++k) a.minv[b][k] = minv[(size_t)(b0 + b) * 6 + k];
        if (launches) *launches += 1;
        dim3 grid((W + 255) / 256, H, nb);
        OF_LAUNCH(warp_affine_kernel, grid, 256, 0, stream, a);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_apply_motion(const uint8_t* src, uint8_t* dst, const double* dx, const double* dy, int batch, int H,
                                int W, double cval, int* launches, cudaStream_t stream) {
    if (batch < 1 || batch > 65535 || H < 1 || H > 65535 || W < 1) return cudaErrorInvalidValue;
    if (launches) *launches += 1;
    dim3 grid((W + 255) / 256, H, batch);
    OF_LAUNCH(apply_motion_kernel, grid, 256, 0, stream, src, dst, dx, dy, H, W, cval);
    return cudaGetLastError();
}

}  // namespace ofb
