// K2 -- one level of build_gaussian_pyramid (lucas_kanade_pyramidal.py:44-59) as a marching
// kernel: gaussian_filter(sigma = 2: 17 taps, reflect) followed by bilinear resampling on the
// np.linspace grid, one pass over the fine image, nothing but the coarse level written to HBM.
//
// Arithmetic is the tile kernel's (pyramid.cu) and therefore SciPy's, bit for bit: per axis a
// float64 accumulation in the symmetric-kernel order (centre tap, then (x[c-k] + x[c+k]) * w[k]
// for k = 8..1, products and sums rounded separately), float32 store after each axis, rows first;
// bilinear taps blended in float64 in map_coordinates' order.
//
// Why a second kernel: the tile kernel spends ~320 instructions per fine pixel, most of them
// float32 -> float64 conversions (16 lanes / clk / SM) that the compiler re-materialises for every
// output, and 2x halo overhead.  Here a CTA owns a strip of 256 fine columns and marches down a
// band of rows, 8 rows per step:
//   axis 0   one thread = one column; the 24 source rows a step touches live in registers as
//            float64 (16 are carried over from the previous step), so every source value is loaded
//            and converted exactly once; next step's 8 rows are prefetched during the arithmetic;
//   axis 1   the 8 x 256 float32 results go through shared memory; one thread = one run of 8
//            consecutive columns of one row (24 LDS, 24 conversions for 8 outputs);
//   resample a 16-row ring of smoothed rows in shared memory; every coarse row whose two source
//            rows are complete is emitted.
// ~64 float64 operations per fine pixel: the kernel is bound by the FP64 pipe (64 lanes / clk /
// SM), not by HBM (5 B per fine pixel).
#include <cuda_runtime.h>

#include <type_traits>

#include "of_common.cuh"
#include "of_kernels.h"

namespace ofb {

namespace {

constexpr int PM_THREADS = 256;
#ifndef OF_PM_MIN_CTAS
#define OF_PM_MIN_CTAS 2
#endif
constexpr int PM_MIN_CTAS = OF_PM_MIN_CTAS;       // resident CTAs per SM the register budget is set for
constexpr int PM_R = 8;                           // radius of the sigma = 2 kernel
constexpr int PM_CH = 8;                          // fine rows per step
constexpr int PM_OUTC = PM_THREADS - 2 * PM_R;    // 240 smoothed columns per strip
constexpr int PM_SEG = 8;                         // axis-1 outputs per thread and step
constexpr int PM_NSEG = PM_OUTC / PM_SEG;         // 30
constexpr int PM_RING = 32;                       // smoothed rows kept in shared memory (four steps)
constexpr int PM_TPITCH = PM_THREADS + 2;         // 258 doubles: quarter-warps on 8 rows hit 8 distinct 16-byte slots
constexpr int PM_SPITCH = PM_OUTC + 4;            // 244: same for the 128-bit stores of the axis-1 results
constexpr int PM_TMP_BYTES = 2 * PM_CH * PM_TPITCH * 8;
constexpr int PM_SMEM_BYTES = PM_TMP_BYTES + PM_RING * PM_SPITCH * 4;
constexpr int PM_MAX_SPAN = PM_OUTC - 8;          // fine columns a strip's taps may span (strip start is 8-aligned)

struct PyrMarchArgs {
    const float* src;
    float* dst;
    int H, W, oh, ow;
    int row_lo, row_hi;  // coarse rows to produce
    int band_rows;       // coarse rows per CTA
    int strip_cols;      // coarse columns per CTA
    double step_y, step_x;
    double w[2 * PM_R + 1];
};

__device__ __forceinline__ int pm_reflect(int i, int n) {
    // scipy 'reflect' (d c b a | a b c d | d c b a) for -n <= i < 2n: one fold.  The launcher only
    // takes frames for which no index of a CTA leaves that range (H >= 16, W >= 248).
    if (i < 0) i = -1 - i;
    return i >= n ? 2 * n - 1 - i : i;
}

// centre tap, then the symmetric pairs from the outside in (scipy's correlate1d for symmetric kernels)
// FMA = false: products and sums rounded separately, like the C code SciPy compiles to on x86-64 (exact
// mode, and the stand-alone of_pyramid_down entry points).  FMA = true (fast mode of the pyramidal
// drivers): each pair term is one fused multiply-add -- 17 instead of 25 float64 operations.  The float64
// sums then differ in their last bits, which survives the float32 store only when the sum sits within
// ~2^-52 of a float32 rounding boundary: about 5 values in 10^9 move by one float32 ulp.
template <bool FMA>
__device__ __forceinline__ double pm_gauss(const double* x, const double* w) {
    if (FMA) {
        // fast mode: two independent accumulators (even / odd pairs) halve the dependent chain of float64 FMAs the
        // kernel waits on (its top stall is the fixed-latency dependency); one more addition, and the float64 sum
        // is associated differently -- like the FMA itself, visible after the float32 store in ~1e-9 of the values
        double a0 = dmul(x[PM_R], w[PM_R]), a1 = 0.0;
#pragma unroll
        for (int ii = -PM_R; ii < 0; ii += 2) {
            a0 = fma(dadd(x[PM_R + ii], x[PM_R - ii]), w[ii + PM_R], a0);
            a1 = fma(dadd(x[PM_R + ii + 1], x[PM_R - ii - 1]), w[ii + 1 + PM_R], a1);
        }
        return dadd(a0, a1);
    }
    double acc = dmul(x[PM_R], w[PM_R]);
#pragma unroll
    for (int ii = -PM_R; ii < 0; ++ii) acc = dadd(acc, dmul(dadd(x[PM_R + ii], x[PM_R - ii]), w[ii + PM_R]));
    return acc;
}
// float32 flavour (PM_F32, an experiment behind OF_B200_PYRAMID_FAST=f32): the same symmetric order with float32 fused
// multiply-adds -- no conversions, a quarter of the pipe time.  A smoothed value then carries a few float32
// roundings (about 2 ulp, 3e-5 on a 0..255 image) instead of one.
template <bool FMA>
__device__ __forceinline__ float pm_gauss(const float* x, const float* w) {
    float acc = fmul(x[PM_R], w[PM_R]);
#pragma unroll
    for (int ii = -PM_R; ii < 0; ++ii) acc = fmaf(fadd(x[PM_R + ii], x[PM_R - ii]), w[ii + PM_R], acc);
    return acc;
}

// flavours of the kernel: SciPy's bits; float64 with fused multiply-adds; float32
enum { PM_EXACT = 0, PM_F64_FMA = 1, PM_F32 = 2 };
template <typename A> struct PmPair;
template <> struct PmPair<double> { typedef double2 type; };
template <> struct PmPair<float> { typedef float2 type; };

template <int FLAVOUR>
__global__ void __launch_bounds__(PM_THREADS, PM_MIN_CTAS) pyramid_march_kernel(const PyrMarchArgs a) {
    constexpr bool FMA = FLAVOUR != PM_EXACT;
    typedef typename std::conditional<FLAVOUR == PM_F32, float, double>::type A;  // accumulator / register type
    typedef typename PmPair<A>::type A2;
    OF_DYNAMIC_SMEM_ALIGNED(16, unsigned char, pm_smem);
    // axis-0 results of two consecutive steps, already rounded to float32 but kept in the accumulator type
    // so that the axis-1 pass needs no conversions; then the ring of fully smoothed rows (float32)
    A* tmp = reinterpret_cast<A*>(pm_smem);                                         // [2][PM_CH][PM_TPITCH]
    float* smo = reinterpret_cast<float*>(pm_smem + PM_TMP_BYTES);                  // [PM_RING][PM_SPITCH]
    A wt[2 * PM_R + 1];
#pragma unroll
    for (int k = 0; k < 2 * PM_R + 1; ++k) wt[k] = (A)a.w[k];
    if (FLAVOUR == PM_F32) {
        // rounding 17 weights to float32 leaves their sum ~1e-8 off 1 -- a bias of that relative size in every
        // smoothed value; the centre tap absorbs it (weights that sum to 1 within half an ulp of the centre tap)
        double others = 0.0;
#pragma unroll
        for (int k = 0; k < 2 * PM_R + 1; ++k)
            if (k != PM_R) others += (double)wt[k];
        wt[PM_R] = (A)(1.0 - others);
    }

    const int H = a.H, W = a.W, tid = threadIdx.x;
    const float* __restrict__ src = a.src + (size_t)blockIdx.z * H * W;
    float* __restrict__ dst = a.dst + (size_t)blockIdx.z * a.oh * a.ow;

    const int j0 = blockIdx.x * a.strip_cols;
    const int jw = min(a.strip_cols, a.ow - j0);
    const int i_lo = a.row_lo + blockIdx.y * a.band_rows;
    const int i_hi = min(i_lo + a.band_rows, a.row_hi);
    if (jw <= 0 || i_lo >= i_hi) return;

    // first fine column a tap of this strip reads, aligned down so that warps load whole sectors
    const int cx0 = ((int)floor(linspace_coord(j0, a.ow, W, a.step_x))) & ~7;
    const int y_first = (int)floor(linspace_coord(i_lo, a.oh, H, a.step_y));
    const int y_last = min((int)floor(linspace_coord(i_hi - 1, a.oh, H, a.step_y)) + 1, H - 1);
    const int n_steps = (y_last - y_first + PM_CH) / PM_CH;

    // axis 0: this thread's source column (reflected at the frame edge)
    const float* __restrict__ colp = src + pm_reflect(cx0 - PM_R + tid, W);

    A x[PM_CH + 2 * PM_R];  // source rows ybase - 8 .. ybase + 15 of this column
#pragma unroll
    for (int k = 0; k < 2 * PM_R; ++k) x[k] = (A)__ldg(colp + (size_t)pm_reflect(y_first - PM_R + k, H) * W);
    float nxt[PM_CH];
#pragma unroll
    for (int k = 0; k < PM_CH; ++k) nxt[k] = __ldg(colp + (size_t)pm_reflect(y_first + PM_R + k, H) * W);

    // axis 1: this thread's row of the step and run of columns
    const int hk = tid & (PM_CH - 1), hseg = tid >> 3;

    // resampling: this thread's coarse column (threads 0..127 / 128..255 take alternate coarse rows)
    const int jj = tid & 127;
    const int j = j0 + min(jj, jw - 1);
    const double xx = linspace_coord(j, a.ow, W, a.step_x);
    const double fx0 = floor(xx);
    const double fx = dsub(xx, fx0), wx0 = dsub(1.0, fx);
    const int x0 = (int)fx0 - cx0;
    // the tap beyond the last row / column has weight exactly 0 (SciPy mirrors its index)
    const int x1 = min((int)fx0 + 1, W - 1) - cx0;

    // Software pipeline, one barrier per iteration: axis 0 of step `it`, axis 1 of step it - 1 and
    // the resampling of step it - 2 run between the same two barriers, so a warp in the FP64-heavy
    // part overlaps with warps in the conversion- and LDS-heavy parts.  tmp is double-buffered; the
    // ring holds four steps (the one being written, the two being read, one spare).
    int next_i = i_lo;
    for (int it = 0; it < n_steps + 2; ++it) {
        if (it < n_steps) {
            const int ybase = y_first + it * PM_CH;
#pragma unroll
            for (int k = 0; k < PM_CH; ++k) x[2 * PM_R + k] = (A)nxt[k];
            // Unconditional: inside a branch the loads' destinations are copied into the loop-carried registers
            // at the end of the block, and that copy waits for the loads (27 % of the kernel's stall samples,
            // profiles/r02_pyrmarch_f32_experiment_*).  The last step re-reads the band's first rows instead.
            const int nrow = (it + 1 < n_steps) ? ybase + PM_CH + PM_R : y_first;
#pragma unroll
            for (int k = 0; k < PM_CH; ++k) nxt[k] = __ldg(colp + (size_t)pm_reflect(nrow + k, H) * W);
            A* tw = tmp + (it & 1) * (PM_CH * PM_TPITCH) + tid;
#pragma unroll
            for (int k = 0; k < PM_CH; ++k) tw[k * PM_TPITCH] = (A)(float)pm_gauss<FMA>(x + k, wt);
#pragma unroll
            for (int k = 0; k < 2 * PM_R; ++k) x[k] = x[k + PM_CH];
        }
        if (it >= 1 && it - 1 < n_steps && hseg < PM_NSEG) {
            const int hs = it - 1;
            const A2* row2 =
                reinterpret_cast<const A2*>(tmp + (hs & 1) * (PM_CH * PM_TPITCH) + hk * PM_TPITCH + hseg * PM_SEG);
            A t[PM_SEG + 2 * PM_R];
#pragma unroll
            for (int q = 0; q < (PM_SEG + 2 * PM_R) / 2; ++q) {
                const A2 f = row2[q];
                t[2 * q + 0] = f.x;
                t[2 * q + 1] = f.y;
            }
            float o[PM_SEG];
#pragma unroll
            for (int k = 0; k < PM_SEG; ++k) o[k] = (float)pm_gauss<FMA>(t + k, wt);
            float4* out4 = reinterpret_cast<float4*>(smo + ((hs * PM_CH + hk) & (PM_RING - 1)) * PM_SPITCH + hseg * PM_SEG);
            out4[0] = make_float4(o[0], o[1], o[2], o[3]);
            out4[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
        if (it >= 2) {
            // Emit every coarse row whose two source rows were complete after step it - 2.  The test
            // is done in integers (same in every thread): row i reads fine rows <= m + 1 with
            // m = floor(i (H-1) / (oh-1)); the float64 coordinate's floor is m or, when the product
            // rounds below an exact integer, m - 1 -- both inside the ring.  The last step emits the
            // rest of the band (y_last was computed from the float64 coordinate itself).
            const int bs = it - 2;
            const int done_row = min(y_first + bs * PM_CH + PM_CH - 1, y_last);
            int i_end = i_hi;
            if (bs + 1 < n_steps) {
                const int lim = (int)(((unsigned)done_row * (unsigned)(a.oh - 1) - 1u) / (unsigned)(H - 1)) + 1;
                i_end = min(i_hi, max(next_i, lim));
            }
            // two coarse rows per pass: threads 0..127 / 128..255 take the columns (strip_cols <= 128)
            if (jj < jw) {
                for (int i = next_i + (tid >> 7); i < i_end; i += 2) {
                    const double y = linspace_coord(i, a.oh, H, a.step_y);
                    const double fy0 = floor(y);
                    const double fy = dsub(y, fy0), wy0 = dsub(1.0, fy);
                    const int y0 = (int)fy0;
                    const int y1 = min(y0 + 1, H - 1);
                    const float* s0 = smo + ((y0 - y_first) & (PM_RING - 1)) * PM_SPITCH;
                    const float* s1 = smo + ((y1 - y_first) & (PM_RING - 1)) * PM_SPITCH;
                    if (FLAVOUR == PM_F32) {
                        // float32 lerps with the float64 grid fractions rounded once
                        const float ffx = (float)fx, ffy = (float)fy;
                        const float top = fmaf(ffx, fsub(s0[x1], s0[x0]), s0[x0]);
                        const float bot = fmaf(ffx, fsub(s1[x1], s1[x0]), s1[x0]);
                        dst[(size_t)i * a.ow + j] = fmaf(ffy, fsub(bot, top), top);
                    } else {
                        double t = 0.0;
                        t = dadd(t, dmul(dmul((double)s0[x0], wy0), wx0));
                        t = dadd(t, dmul(dmul((double)s0[x1], wy0), fx));
                        t = dadd(t, dmul(dmul((double)s1[x0], fy), wx0));
                        t = dadd(t, dmul(dmul((double)s1[x1], fy), fx));
                        dst[(size_t)i * a.ow + j] = (float)t;
                    }
                }
            }
            next_i = i_end;
        }
        __syncthreads();
    }
}

}  // namespace

bool pyramid_march_supported(int H, int W, int oh, int ow, int radius) {
    if (radius != PM_R || oh < 2 || ow < 2 || H < 16 || W < PM_THREADS - PM_R) return false;
    const double sy = (double)(H - 1) / (double)(oh - 1), sx = (double)(W - 1) / (double)(ow - 1);
    // a coarse row must not need smoothed rows older than the ring holds; strips must fit 232 columns
    return sy >= 1.0 && sy <= 6.0 && sx >= 1.0 && sx <= 6.0;
}

cudaError_t launch_pyramid_march(const float* src, float* dst, int batch, int H, int W, int oh, int ow,
                                 const double* weights, int row_lo, int row_hi, int flavour, int* launches,
                                 cudaStream_t stream) {
    if (flavour < PM_EXACT || flavour > PM_F32) return cudaErrorInvalidValue;
    if (batch > 65535 || row_lo < 0 || row_hi > oh || row_lo >= row_hi) return cudaErrorInvalidValue;
    PyrMarchArgs a;
    a.src = src;
    a.dst = dst;
    a.H = H;
    a.W = W;
    a.oh = oh;
    a.ow = ow;
    a.row_lo = row_lo;
    a.row_hi = row_hi;
    a.step_y = (double)(H - 1) / (double)(oh - 1);  // np.linspace step
    a.step_x = (double)(W - 1) / (double)(ow - 1);
    for (int i = 0; i < 2 * PM_R + 1; ++i) a.w[i] = weights[i];
    // coarse columns per strip: floor(j1 * step) + 1 - floor(j0 * step) <= PM_MAX_SPAN - 1
    int sc = (int)((PM_MAX_SPAN - 3) / a.step_x) + 1;
    if (sc < 1) sc = 1;
    if (sc > 128) sc = 128;  // the resampling pass maps 128 threads to a strip's columns
    a.strip_cols = sc;
    const int n_strips = (ow + sc - 1) / sc;
    // bands: CTAs run in waves of 148 SMs x 2 resident CTAs and every band recomputes 16 fine rows
    // (two steps) of filter warm-up; minimise  waves x (steps per band + 2)
    const int rows = row_hi - row_lo;
    const long long per_band = (long long)batch * n_strips, slots = 148LL * PM_MIN_CTAS;
    const int max_bands = (rows + 7) / 8;
    long long best_cost = -1;
    a.band_rows = rows;
    for (int nb = 1; nb <= max_bands && nb <= 512; ++nb) {
        const int br = (rows + nb - 1) / nb;
        const int bands = (rows + br - 1) / br;
        const long long waves = (per_band * bands + slots - 1) / slots;
        const long long steps = (long long)(br * a.step_y + PM_CH) / PM_CH + 2;
        const long long cost = waves * steps;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            a.band_rows = br;
        }
    }
    const int n_bands = (rows + a.band_rows - 1) / a.band_rows;
    static SmemOptIn opt_in[3];
    {
        cudaError_t e = flavour == PM_F32       ? opt_in[2].ensure(pyramid_march_kernel<PM_F32>, PM_SMEM_BYTES)
                        : flavour == PM_F64_FMA ? opt_in[1].ensure(pyramid_march_kernel<PM_F64_FMA>, PM_SMEM_BYTES)
                                                : opt_in[0].ensure(pyramid_march_kernel<PM_EXACT>, PM_SMEM_BYTES);
        if (e != cudaSuccess) return e;
    }
    if (launches) *launches += 1;
    dim3 grid(n_strips, n_bands, batch);
    if (flavour == PM_F32)
        OF_LAUNCH(pyramid_march_kernel<PM_F32>, grid, PM_THREADS, PM_SMEM_BYTES, stream, a);
    else if (flavour == PM_F64_FMA)
        OF_LAUNCH(pyramid_march_kernel<PM_F64_FMA>, grid, PM_THREADS, PM_SMEM_BYTES, stream, a);
    else
        OF_LAUNCH(pyramid_march_kernel<PM_EXACT>, grid, PM_THREADS, PM_SMEM_BYTES, stream, a);
    return cudaGetLastError();
}

}  // namespace ofb
