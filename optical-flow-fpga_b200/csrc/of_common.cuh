// Shared device helpers for the B200 Lucas-Kanade kernels.
//
// Every arithmetic step that must round exactly like the NumPy/SciPy reference is
// written with the *_rn intrinsics, which the compiler never contracts into FMAs
// (the library is also built with -fmad=false).  Explicit fmaf() is used only in
// the fast kernels, where it is stated why it cannot change the result.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define OF_DET_EPS 1e-4f          // lucas_kanade_core.py:131
#define OF_CONVERGENCE_EPS 0.01f  // lucas_kanade_pyramidal.py:221

// the kernel's dynamic shared memory as an array `name` of `type` (tests/host_emul/ points it at a host buffer)
#ifndef OF_DYNAMIC_SMEM
#define OF_DYNAMIC_SMEM(type, name) extern __shared__ type name[]
#define OF_DYNAMIC_SMEM_ALIGNED(align, type, name) extern __shared__ __align__(align) type name[]
#endif

namespace ofb {

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// 2x2 Cramer solve in the reference's operation order (lucas_kanade_core.py:122-133):
// every product and difference rounds to float32 on its own, IEEE division.
__device__ __forceinline__ void cramer_solve(float sxx, float syy, float sxy, float sxt, float syt,
                                             float& u, float& v) {
    const float b0 = -sxt;
    const float b1 = -syt;
    const float det = fsub(fmul(sxx, syy), fmul(sxy, sxy));
    if (fabsf(det) > OF_DET_EPS) {
        u = fdiv(fsub(fmul(syy, b0), fmul(sxy, b1)), det);
        v = fdiv(fsub(fmul(sxx, b1), fmul(sxy, b0)), det);
    } else {
        u = 0.0f;
        v = 0.0f;
    }
}

// Branch-free variant for the fused kernels: the division always runs on a safe denominator
// (no special-case slow path when det == 0) and the result is selected afterwards.
// `inside` = the pixel is not on the window_size//2 border, which stays 0.
__device__ __forceinline__ void cramer_solve_select(float sxx, float syy, float sxy, float sxt, float syt,
                                                    bool inside, float& u, float& v) {
    const float b0 = -sxt;
    const float b1 = -syt;
    const float det = fsub(fmul(sxx, syy), fmul(sxy, sxy));
    const bool ok = inside && (fabsf(det) > OF_DET_EPS);
    const float den = ok ? det : 1.0f;
    const float nu = fsub(fmul(syy, b0), fmul(sxy, b1));
    const float nv = fsub(fmul(sxx, b1), fmul(sxy, b0));
    const float qu = fdiv(nu, den);
    const float qv = fdiv(nv, den);
    u = ok ? qu : 0.0f;
    v = ok ? qv : 0.0f;
}

// scipy.ndimage.map_coordinates(order=1, mode="constant", cval=0) for one sample.
// float64 coordinates; inside iff 0 <= y <= H-1 and 0 <= x <= W-1; the four taps are
// blended in float64 in row-major order, each as (value*wy)*wx, summed from 0.0.
__device__ __forceinline__ float bilinear_f64(const float* __restrict__ img, int H, int W, double y,
                                              double x) {
    if (!(y >= 0.0 && y <= (double)(H - 1) && x >= 0.0 && x <= (double)(W - 1))) return 0.0f;
    const double fy0 = floor(y), fx0 = floor(x);
    const double fy = dsub(y, fy0), fx = dsub(x, fx0);
    const int y0 = (int)fy0, x0 = (int)fx0;
    // the tap past the last row/column is mirrored by SciPy; its weight is exactly 0 there
    const int y1 = (y0 + 1 > H - 1) ? (H >= 2 ? H - 2 : 0) : y0 + 1;
    const int x1 = (x0 + 1 > W - 1) ? (W >= 2 ? W - 2 : 0) : x0 + 1;
    const double wy0 = dsub(1.0, fy), wx0 = dsub(1.0, fx);
    const float* r0 = img + (size_t)y0 * W;
    const float* r1 = img + (size_t)y1 * W;
    double t = 0.0;
    t = dadd(t, dmul(dmul((double)__ldg(r0 + x0), wy0), wx0));
    t = dadd(t, dmul(dmul((double)__ldg(r0 + x1), wy0), fx));
    t = dadd(t, dmul(dmul((double)__ldg(r1 + x0), fy), wx0));
    t = dadd(t, dmul(dmul((double)__ldg(r1 + x1), fy), fx));
    return (float)t;
}

// np.linspace(0, n_src - 1, n_dst)[i] in float64: i * step, last element forced to the end.
__device__ __forceinline__ double linspace_coord(int i, int n_dst, int n_src, double step) {
    return (n_dst > 1 && i == n_dst - 1) ? (double)(n_src - 1) : dmul((double)i, step);
}

}  // namespace ofb
