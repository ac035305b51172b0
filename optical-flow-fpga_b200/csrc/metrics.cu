// Flow-field error metrics on the device: compute_all_metrics (python/flow_metrics.py:166-201) over
// the verifier's rectangular test region (python/optical_flow_verifier.py:96-138), for a batch of
// flow fields that never leave HBM.
//
// Per pixel the arithmetic is the reference's float32 expression, in its operand order:
//   eu = u - u_true, ev = v - v_true; |eu|, |ev|; sq = eu*eu + ev*ev; sqrt(sq);
//   angular error = rad2deg(arccos(clip((u*ut + v*vt + 1) / (sqrt(u*u + v*v + 1) * sqrt(ut*ut + vt*vt + 1)))))
// What differs is the reduction: the reference takes np.mean in float32 (pairwise), here the sums
// are float64, reduced in a fixed order (deterministic).  Metrics therefore agree with the
// reference to float32 rounding of a mean (a few 1e-7 relative), not bit for bit; arccos is CUDA's
// acosf (<= 1 ulp from NumPy's).
#include <cuda_runtime.h>

#include "of_common.cuh"
#include "of_kernels.h"

namespace ofb {

namespace {

constexpr int MET_THREADS = 256;
constexpr int MET_SUMS = 5;  // sum|eu|, sum|ev|, sum sq, sum sqrt(sq), sum angular error

struct MetricsArgs {
    const float* u;
    const float* v;
    const float* u_true;  // [batch]
    const float* v_true;
    int H, W, y0, y1, x0, x1;
    int blocks_per_pair;
    double* partial;  // [batch][blocks_per_pair][MET_SUMS + 1]; the last entry is max(u*u + v*v)
};

__global__ void __launch_bounds__(MET_THREADS) metrics_partial_kernel(const MetricsArgs a) {
    const int pair = blockIdx.y;
    const float ut = a.u_true[pair], vt = a.v_true[pair];
    const float* __restrict__ u = a.u + (size_t)pair * a.H * a.W;
    const float* __restrict__ v = a.v + (size_t)pair * a.H * a.W;
    const int rw = a.x1 - a.x0;
    const long long n = (long long)(a.y1 - a.y0) * rw;
    const float norm_true = sqrtf(fadd(fadd(fmul(ut, ut), fmul(vt, vt)), 1.0f));
    const float rad2deg = 57.29577951308232f;  // float32(180 / pi), NumPy's rad2deg factor
    double s[MET_SUMS] = {0.0, 0.0, 0.0, 0.0, 0.0};
    float mx = 0.0f;
    for (long long i = (long long)blockIdx.x * MET_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * MET_THREADS) {
        const int y = a.y0 + (int)(i / rw), x = a.x0 + (int)(i % rw);
        const size_t o = (size_t)y * a.W + x;
        const float pu = __ldg(u + o), pv = __ldg(v + o);
        const float eu = fsub(pu, ut), ev = fsub(pv, vt);
        const float sq = fadd(fmul(eu, eu), fmul(ev, ev));
        const float m2 = fadd(fmul(pu, pu), fmul(pv, pv));
        const float norm_pred = sqrtf(fadd(m2, 1.0f));
        float dot = fdiv(fadd(fadd(fmul(pu, ut), fmul(pv, vt)), 1.0f), fmul(norm_pred, norm_true));
        dot = fminf(fmaxf(dot, -1.0f), 1.0f);
        s[0] += (double)fabsf(eu);
        s[1] += (double)fabsf(ev);
        s[2] += (double)sq;
        s[3] += (double)sqrtf(sq);
        s[4] += (double)fmul(acosf(dot), rad2deg);
        mx = fmaxf(mx, m2);
    }
    __shared__ double red[MET_SUMS][MET_THREADS / 32];
    __shared__ float redm[MET_THREADS / 32];
#pragma unroll
    for (int k = 0; k < MET_SUMS; ++k)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s[k] += __shfl_down_sync(0xffffffffu, s[k], off);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < MET_SUMS; ++k) red[k][threadIdx.x >> 5] = s[k];
        redm[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double* out = a.partial + ((size_t)pair * a.blocks_per_pair + blockIdx.x) * (MET_SUMS + 1);
        for (int k = 0; k < MET_SUMS; ++k) {
            double t = 0.0;
            for (int w = 0; w < MET_THREADS / 32; ++w) t += red[k][w];
            out[k] = t;
        }
        float m = 0.0f;
        for (int w = 0; w < MET_THREADS / 32; ++w) m = fmaxf(m, redm[w]);
        out[MET_SUMS] = (double)m;
    }
}

// out[pair] = {mae_u, mae_v, rmse, epe, aae}
__global__ void metrics_final_kernel(const MetricsArgs a, double* __restrict__ out, int batch) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= batch) return;
    const double n = (double)(a.y1 - a.y0) * (double)(a.x1 - a.x0);
    double s[MET_SUMS] = {0.0, 0.0, 0.0, 0.0, 0.0};
    double mx = 0.0;
    const double* p = a.partial + (size_t)pair * a.blocks_per_pair * (MET_SUMS + 1);
    for (int b = 0; b < a.blocks_per_pair; ++b) {
        for (int k = 0; k < MET_SUMS; ++k) s[k] += p[b * (MET_SUMS + 1) + k];
        mx = fmax(mx, p[b * (MET_SUMS + 1) + MET_SUMS]);
    }
    const float ut = a.u_true[pair], vt = a.v_true[pair];
    // flow_metrics.py:143-147: no motion in truth and prediction -> angular error 0
    const bool still = sqrt((double)ut * ut + (double)vt * vt) < 1e-6 && sqrtf((float)mx) < 1e-6f;
    double* o = out + (size_t)pair * 5;
    o[0] = s[0] / n;
    o[1] = s[1] / n;
    o[2] = sqrt(s[2] / n);
    o[3] = s[3] / n;
    o[4] = still ? 0.0 : s[4] / n;
}

}  // namespace

int metrics_blocks_per_pair(int rows, int cols) {
    long long n = (long long)rows * cols;
    long long b = (n + MET_THREADS * 8 - 1) / (MET_THREADS * 8);
    if (b < 1) b = 1;
    if (b > 592) b = 592;  // 148 SMs x 4
    return (int)b;
}

cudaError_t launch_flow_metrics(const float* u, const float* v, const float* u_true, const float* v_true, int batch, int H,
                                int W, int y0, int y1, int x0, int x1, double* partial, double* out, int* launches,
                                cudaStream_t stream) {
    if (batch < 1 || batch > 65535 || y0 < 0 || y1 > H || x0 < 0 || x1 > W || y0 >= y1 || x0 >= x1) return cudaErrorInvalidValue;
    MetricsArgs a;
    a.u = u;
    a.v = v;
    a.u_true = u_true;
    a.v_true = v_true;
    a.H = H;
    a.W = W;
    a.y0 = y0;
    a.y1 = y1;
    a.x0 = x0;
    a.x1 = x1;
    a.blocks_per_pair = metrics_blocks_per_pair(y1 - y0, x1 - x0);
    a.partial = partial;
    if (launches) *launches += 2;
    dim3 grid(a.blocks_per_pair, batch);
    OF_LAUNCH(metrics_partial_kernel, grid, MET_THREADS, 0, stream, a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    OF_LAUNCH(metrics_final_kernel, (batch + 127) / 128, 128, 0, stream, a, out, batch);
    return cudaGetLastError();
}

}  // namespace ofb
