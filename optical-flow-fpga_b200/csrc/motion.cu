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

}  // namespace

cudaError_t launch_apply_motion(const uint8_t* src, uint8_t* dst, const double* dx, const double* dy, int batch, int H,
                                int W, double cval, int* launches, cudaStream_t stream) {
    if (batch < 1 || batch > 65535 || H < 1 || H > 65535 || W < 1) return cudaErrorInvalidValue;
    if (launches) *launches += 1;
    dim3 grid((W + 255) / 256, H, batch);
    apply_motion_kernel<<<grid, 256, 0, stream>>>(src, dst, dx, dy, H, W, cval);
    return cudaGetLastError();
}

}  // namespace ofb
