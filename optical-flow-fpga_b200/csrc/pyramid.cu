// K2 / K4 and the stand-alone helpers of the pyramidal path.
//
//   pyramid_down   one level of build_gaussian_pyramid   (lucas_kanade_pyramidal.py:44-59)
//   warp           warp_image                            (lucas_kanade_pyramidal.py:66-97)
//   upsample_flow  upsample_flow                         (lucas_kanade_pyramidal.py:100-138)
//   gradients      compute_gradients                     (lucas_kanade_core.py:15-45)
//
// SciPy semantics that are mirrored (pinned by tests/golden):
//   gaussian_filter: per axis float64 accumulate in SciPy's symmetric-kernel order
//   (centre tap, then (x[c-k] + x[c+k]) * w[k] for k = radius..1), float32 store after
//   each axis, rows (axis 0) first, 'reflect' boundary.  map_coordinates(order=1,
//   mode="constant"): float64 coordinates and blend, outside -> exactly 0.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include "of_common.cuh"
#include "of_kernels.h"

namespace ofb {

// ---------------------------------------------------------------------------------------
// compute_gradients
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gradients_kernel(const float* __restrict__ prev, const float* __restrict__ curr,
                                                         float* __restrict__ ix, float* __restrict__ iy,
                                                         float* __restrict__ it, int H, int W) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)H * W;
    const float* p = prev + blockIdx.z * plane;
    const float* c = curr + blockIdx.z * plane;
    const float kx[3][3] = {{-0.125f, 0.0f, 0.125f}, {-0.25f, 0.0f, 0.25f}, {-0.125f, 0.0f, 0.125f}};
    const float ky[3][3] = {{-0.125f, -0.25f, -0.125f}, {0.0f, 0.0f, 0.0f}, {0.125f, 0.25f, 0.125f}};
    float ax = 0.0f, ay = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int yy = clampi(y + 1 - j, 0, H - 1), xx = clampi(x + 1 - k, 0, W - 1);
            const size_t o = (size_t)yy * W + xx;
            const float avg = fmul(fadd(__ldg(p + o), __ldg(c + o)), 0.5f);
            ax = fadd(ax, fmul(avg, kx[j][k]));
            ay = fadd(ay, fmul(avg, ky[j][k]));
        }
    const size_t o = (size_t)y * W + x;
    ix[blockIdx.z * plane + o] = ax;
    iy[blockIdx.z * plane + o] = ay;
    it[blockIdx.z * plane + o] = fsub(__ldg(p + o), __ldg(c + o));
}

cudaError_t launch_gradients(const float* prev, const float* curr, float* ix, float* iy, float* it, int batch, int H,
                             int W, int* launches, cudaStream_t stream) {
    if (launches) *launches += 1;
    dim3 grid((W + 63) / 64, (H + 3) / 4, batch);
    OF_LAUNCH(gradients_kernel, grid, 256, 0, stream, prev, curr, ix, iy, it, H, W);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// fused Gaussian blur + bilinear decimation
// ---------------------------------------------------------------------------------------
constexpr int PYR_MAX_RADIUS = 16;
constexpr int PYR_RMAX = 40;  // smoothed rows a CTA may need
constexpr int PYR_CMAX = 72;  // smoothed columns a CTA may need
constexpr int PYR_SEG = 6;    // outputs per sliding-window run (R = 36 and C = 66 are multiples)

struct PyrArgs {
    const float* src;
    float* dst;
    int H, W, oh, ow;
    int tile_h, tile_w;  // coarse pixels per CTA
    int row_lo, row_hi;  // coarse rows to produce
    int radius;
    double step_y, step_x;
    double w[2 * PYR_MAX_RADIUS + 1];
};

__device__ __forceinline__ int reflect_index(int i, int n) {
    // scipy 'reflect': (d c b a | a b c d | d c b a).  One fold covers -n <= i < 2n (every tile
    // of a frame taller / wider than the filter radius); the modulo form handles any overhang.
    if (i >= -n && i < 2 * n) {
        if (i < 0) i = -1 - i;
        return i >= n ? 2 * n - 1 - i : i;
    }
    const int period = 2 * n;
    int m = i % period;
    if (m < 0) m += period;
    return m >= n ? period - 1 - m : m;
}

// One Gaussian output in SciPy's symmetric-kernel order: centre tap, then for k = R .. 1 the
// pair (x[c-k] + x[c+k]) * w[R-k], everything float64, rounded to float32 by the caller.
template <int R>
__device__ __forceinline__ double gauss_tap_sum(const double* x, const double* w) {
    double acc = dmul(x[R], w[R]);
#pragma unroll
    for (int ii = -R; ii < 0; ++ii) acc = dadd(acc, dmul(dadd(x[R + ii], x[R - ii]), w[ii + R]));
    return acc;
}

// RADIUS > 0: sliding-window version (every input is read from shared memory and converted
// to float64 once per run of SEG outputs).  RADIUS == 0: any radius <= PYR_MAX_RADIUS, one
// output per thread and pass (only sigma != 2 gets here; the reference always uses sigma = 2).
template <int RADIUS>
__global__ void __launch_bounds__(256) pyramid_down_kernel(PyrArgs a) {
    OF_DYNAMIC_SMEM(float, smem);
    constexpr int SEG = PYR_SEG;
    const int H = a.H, W = a.W, r = RADIUS > 0 ? RADIUS : a.radius;
    const float* src = a.src + (size_t)blockIdx.z * H * W;
    float* dst = a.dst + (size_t)blockIdx.z * a.oh * a.ow;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    const int i0 = a.row_lo + blockIdx.y * a.tile_h, j0 = blockIdx.x * a.tile_w;
    const int i1 = min(i0 + a.tile_h, a.row_hi) - 1, j1 = min(j0 + a.tile_w, a.ow) - 1;
    // fine rows / columns whose smoothed value the bilinear taps of this tile can touch
    const int fy_lo = (int)floor(linspace_coord(i0, a.oh, H, a.step_y));
    const int fy_hi = min((int)floor(linspace_coord(i1, a.oh, H, a.step_y)) + 1, H - 1);
    const int fx_lo = (int)floor(linspace_coord(j0, a.ow, W, a.step_x));
    const int fx_hi = min((int)floor(linspace_coord(j1, a.ow, W, a.step_x)) + 1, W - 1);
    const int R = fy_hi - fy_lo + 1, C = fx_hi - fx_lo + 1;
    const int IW = C + 2 * r, IH = R + 2 * r;
    // odd pitches: the axis-1 pass and the resampling walk rows with lanes on different rows /
    // every other column.  Both intermediate planes are padded by one segment so that the
    // sliding windows can read past the tile end without a clamp (those outputs are discarded).
    const int IWP = (IW + SEG) | 1;
    const int CP = C | 1;

    float* img = smem;                    // [IH + SEG][IW]  source with reflected halo
    float* tmp = img + (IH + SEG) * IW;   // [R][IWP]        after the axis-0 pass (float32)
    float* smo = tmp + R * IWP;           // [R][CP]         after the axis-1 pass (float32)

    for (int y = wid; y < IH; y += 8) {
        const float* row = src + (size_t)reflect_index(fy_lo - r + y, H) * W;
        for (int x = lane; x < IW; x += 32) img[y * IW + x] = __ldg(row + reflect_index(fx_lo - r + x, W));
    }
    __syncthreads();

    if (RADIUS > 0) {
        constexpr int RR = RADIUS > 0 ? RADIUS : 1;
        // axis 0: item = (row segment, column); consecutive lanes take consecutive columns
        const int nseg_v = (R + SEG - 1) / SEG;
        for (int item = tid; item < nseg_v * IW; item += 256) {
            const int seg = item / IW, cc = item - seg * IW;
            const int r0 = seg * SEG;
            const float* col = img + r0 * IW + cc;
            double x[SEG + 2 * RR];
#pragma unroll
            for (int k = 0; k < SEG + 2 * RR; ++k) x[k] = (double)col[k * IW];
            float* out = tmp + r0 * IWP + cc;
#pragma unroll
            for (int k = 0; k < SEG; ++k)
                if (r0 + k < R) out[k * IWP] = (float)gauss_tap_sum<RR>(x + k, a.w);
        }
        __syncthreads();
        // axis 1: item = (column segment, row); consecutive lanes take consecutive rows
        const int nseg_h = (C + SEG - 1) / SEG;
        for (int item = tid; item < nseg_h * R; item += 256) {
            const int seg = item / R, rr = item - seg * R;
            const int c0 = seg * SEG;
            const float* row = tmp + rr * IWP + c0;
            double x[SEG + 2 * RR];
#pragma unroll
            for (int k = 0; k < SEG + 2 * RR; ++k) x[k] = (double)row[k];
            float* out = smo + rr * CP + c0;
#pragma unroll
            for (int k = 0; k < SEG; ++k)
                if (c0 + k < C) out[k] = (float)gauss_tap_sum<RR>(x + k, a.w);
        }
    } else {
        for (int i = tid; i < R * IW; i += 256) {
            const int rr = i / IW, cc = i % IW;
            const float* col = img + (rr + r) * IW + cc;
            double acc = dmul((double)col[0], a.w[r]);
            for (int ii = -r; ii < 0; ++ii)
                acc = dadd(acc, dmul(dadd((double)col[ii * IW], (double)col[-ii * IW]), a.w[ii + r]));
            tmp[rr * IWP + cc] = (float)acc;
        }
        __syncthreads();
        for (int i = tid; i < R * C; i += 256) {
            const int rr = i / C, cc = i % C;
            const float* row = tmp + rr * IWP + cc + r;
            double acc = dmul((double)row[0], a.w[r]);
            for (int ii = -r; ii < 0; ++ii)
                acc = dadd(acc, dmul(dadd((double)row[ii], (double)row[-ii]), a.w[ii + r]));
            smo[rr * CP + cc] = (float)acc;
        }
    }
    __syncthreads();
    const int th = i1 - i0 + 1, tw = j1 - j0 + 1;
    for (int o = tid; o < th * tw; o += 256) {
        const int i = i0 + o / tw, j = j0 + o % tw;
        const double y = linspace_coord(i, a.oh, H, a.step_y);
        const double x = linspace_coord(j, a.ow, W, a.step_x);
        const double fy0 = floor(y), fx0 = floor(x);
        const double fy = dsub(y, fy0), fx = dsub(x, fx0);
        const int y0 = (int)fy0, x0 = (int)fx0;
        // the tap beyond the last row/column has weight exactly 0 (SciPy mirrors its index)
        const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
        const double wy0 = dsub(1.0, fy), wx0 = dsub(1.0, fx);
        const float* s0 = smo + (y0 - fy_lo) * CP - fx_lo;
        const float* s1 = smo + (y1 - fy_lo) * CP - fx_lo;
        double t = 0.0;
        t = dadd(t, dmul(dmul((double)s0[x0], wy0), wx0));
        t = dadd(t, dmul(dmul((double)s0[x1], wy0), fx));
        t = dadd(t, dmul(dmul((double)s1[x0], fy), wx0));
        t = dadd(t, dmul(dmul((double)s1[x1], fy), fx));
        dst[(size_t)i * a.ow + j] = (float)t;
    }
}

static int pyr_tile_extent(int limit, double step) {
    // largest n with ceil((n - 1) * step) + 2 <= limit, capped
    int n = (int)((limit - 3) / step) + 1;
    if (n < 1) n = 1;
    return n;
}

cudaError_t launch_pyramid_down(const float* src, float* dst, int batch, int H, int W, int oh, int ow,
                                const double* weights, int radius, int row_lo, int row_hi, int* launches,
                                cudaStream_t stream, bool fast) {
    if (radius < 0 || radius > PYR_MAX_RADIUS || oh < 1 || ow < 1 || batch > 65535) return cudaErrorInvalidValue;
    if (row_lo < 0 || row_hi > oh || row_lo >= row_hi) return cudaErrorInvalidValue;
    // sigma = 2 (the reference's pyramid) on a ~2x decimation: the marching kernel; anything else
    // (and OF_B200_PYRAMID=tile, for A/B measurements) takes the generic tile kernel below.  Both
    // produce the same bits.
    static const bool force_tile = [] {
        const char* e = getenv("OF_B200_PYRAMID");
        return e != nullptr && strcmp(e, "tile") == 0;
    }();
    static const int fast_flavour = [] {
        // f32: the float32 flavour of the marching kernel (experiment: 2 x faster, but the per-pixel deviation of
        // the fast pyramidal flow from the reference grows -- see DESIGN.md K2); default: float64 + FMA
        const char* e = getenv("OF_B200_PYRAMID_FAST");
        return (e != nullptr && strcmp(e, "f32") == 0) ? 2 : 1;
    }();
    if (!force_tile && pyramid_march_supported(H, W, oh, ow, radius))
        return launch_pyramid_march(src, dst, batch, H, W, oh, ow, weights, row_lo, row_hi, fast ? fast_flavour : 0, launches,
                                    stream);
    PyrArgs a;
    a.src = src;
    a.dst = dst;
    a.H = H;
    a.W = W;
    a.oh = oh;
    a.ow = ow;
    a.radius = radius;
    a.row_lo = row_lo;
    a.row_hi = row_hi;
    a.step_y = oh > 1 ? (double)(H - 1) / (double)(oh - 1) : 0.0;  // np.linspace step
    a.step_x = ow > 1 ? (double)(W - 1) / (double)(ow - 1) : 0.0;
    for (int i = 0; i < 2 * radius + 1; ++i) a.w[i] = weights[i];
    a.tile_h = a.step_y > 0 ? pyr_tile_extent(PYR_RMAX, a.step_y) : 16;
    a.tile_w = a.step_x > 0 ? pyr_tile_extent(PYR_CMAX, a.step_x) : 32;
    if (a.tile_h > 16) a.tile_h = 16;
    if (a.tile_w > 32) a.tile_w = 32;
    const int IHmax = PYR_RMAX + 2 * radius, IWmax = PYR_CMAX + 2 * radius;
    const size_t smem =
        (size_t)((IHmax + PYR_SEG) * IWmax + PYR_RMAX * (IWmax + PYR_SEG + 1) + PYR_RMAX * (PYR_CMAX + 1)) * sizeof(float);
    static SmemOptIn opt_in[2];
    const int which = radius == 8 ? 1 : 0;
    {
        // the generic variant is opted in for its largest radius once, whatever radius comes first
        const int IHcap = PYR_RMAX + 2 * PYR_MAX_RADIUS, IWcap = PYR_CMAX + 2 * PYR_MAX_RADIUS;
        const size_t cap = (size_t)((IHcap + PYR_SEG) * IWcap + PYR_RMAX * (IWcap + PYR_SEG + 1) + PYR_RMAX * (PYR_CMAX + 1)) * sizeof(float);
        cudaError_t e = which ? opt_in[1].ensure(pyramid_down_kernel<8>, smem) : opt_in[0].ensure(pyramid_down_kernel<0>, cap);
        if (e != cudaSuccess) return e;
    }
    if (launches) *launches += 1;
    dim3 grid((ow + a.tile_w - 1) / a.tile_w, (row_hi - row_lo + a.tile_h - 1) / a.tile_h, batch);
    if (which)
        OF_LAUNCH(pyramid_down_kernel<8>, grid, 256, smem, stream, a);  // sigma = 2, the reference's pyramid
    else
        OF_LAUNCH(pyramid_down_kernel<0>, grid, 256, smem, stream, a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// warp_image / upsample_flow / select-copy
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) warp_kernel(const float* __restrict__ img, const float* __restrict__ fu,
                                                    const float* __restrict__ fv, float* __restrict__ out, int H,
                                                    int W) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)H * W, o = blockIdx.z * plane + (size_t)y * W + x;
    const double yw = dadd((double)y, (double)__ldg(fv + o));
    const double xw = dadd((double)x, (double)__ldg(fu + o));
    out[o] = bilinear_f64(img + blockIdx.z * plane, H, W, yw, xw);
}

cudaError_t launch_warp(const float* img, const float* fu, const float* fv, float* out, int batch, int H, int W,
                        int* launches, cudaStream_t stream) {
    if (launches) *launches += 1;
    dim3 grid((W + 63) / 64, (H + 3) / 4, batch);
    OF_LAUNCH(warp_kernel, grid, 256, 0, stream, img, fu, fv, out, H, W);
    return cudaGetLastError();
}

struct UpArgs {
    const float* cu[2];
    const float* cv[2];
    const int* sel;
    int sel_xor;
    float* fu;
    float* fv;
    int ch, cw, th, tw;
    int row_lo, row_hi;  // target rows to produce
    double step_y, step_x;
    float scale_y, scale_x;
};

// One thread = one target column and UP_ROWS consecutive target rows: the column's coordinate,
// taps and weights are computed once, u and v share every row's coordinate and weights, and the
// converted coarse taps are kept while consecutive target rows fall between the same two coarse
// rows (every other row at 2x), so each coarse value is loaded and widened once per thread.  The
// linspace grid never leaves [0, n - 1], so every sample is inside the coarse field.
constexpr int UP_ROWS = 16;

__global__ void __launch_bounds__(256) upsample_flow_kernel(UpArgs a) {
    // the rows' coordinates are the same for every column: 16 threads compute them once per CTA
    __shared__ int s_y0[UP_ROWS], s_y1[UP_ROWS];
    __shared__ double s_fy[UP_ROWS], s_wy0[UP_ROWS];
    const int y_begin = a.row_lo + blockIdx.y * UP_ROWS, y_end = min(y_begin + UP_ROWS, a.row_hi);
    if (threadIdx.x < UP_ROWS) {
        const int y = min(y_begin + (int)threadIdx.x, a.th - 1);
        const double yc = linspace_coord(y, a.th, a.ch, a.step_y);
        const double fy0 = floor(yc);
        const double fy = dsub(yc, fy0);
        const int y0 = (int)fy0;
        s_y0[threadIdx.x] = y0;
        s_y1[threadIdx.x] = min(y0 + 1, a.ch - 1);
        s_fy[threadIdx.x] = fy;
        s_wy0[threadIdx.x] = dsub(1.0, fy);
    }
    __syncthreads();
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= a.tw) return;
    const int pair = blockIdx.z;
    const int cur = (a.sel ? a.sel[pair] : 0) ^ a.sel_xor;
    const size_t cplane = (size_t)a.ch * a.cw;
    const float* __restrict__ cu = (cur ? a.cu[1] : a.cu[0]) + pair * cplane;
    const float* __restrict__ cv = (cur ? a.cv[1] : a.cv[0]) + pair * cplane;
    const size_t o0 = (size_t)pair * a.th * a.tw + (size_t)y_begin * a.tw + x;
    float* __restrict__ fu = a.fu + o0;
    float* __restrict__ fv = a.fv + o0;

    const double xc = linspace_coord(x, a.tw, a.cw, a.step_x);
    const double fx0 = floor(xc);
    const double fx = dsub(xc, fx0), wx0 = dsub(1.0, fx);
    const int x0 = (int)fx0;
    const int x1 = min(x0 + 1, a.cw - 1);  // weight 0 when it would leave the field
    int cy0 = -1, cy1 = -1;  // coarse rows held in (u0*, v0*) and (u1*, v1*)
    double u00 = 0.0, u01 = 0.0, u10 = 0.0, u11 = 0.0, v00 = 0.0, v01 = 0.0, v10 = 0.0, v11 = 0.0;
    const int n = y_end - y_begin;
    for (int k = 0; k < n; ++k) {
        const int y0 = s_y0[k], y1 = s_y1[k];
        const double fy = s_fy[k], wy0 = s_wy0[k];
        if (y0 != cy0) {  // uniform over the CTA
            if (y0 == cy1) {
                u00 = u10; u01 = u11; v00 = v10; v01 = v11;
            } else {
                const unsigned r0 = (unsigned)(y0 * a.cw);
                u00 = (double)__ldg(cu + r0 + x0); u01 = (double)__ldg(cu + r0 + x1);
                v00 = (double)__ldg(cv + r0 + x0); v01 = (double)__ldg(cv + r0 + x1);
            }
            cy0 = y0;
        }
        if (y1 != cy1) {
            const unsigned r1 = (unsigned)(y1 * a.cw);
            u10 = (double)__ldg(cu + r1 + x0); u11 = (double)__ldg(cu + r1 + x1);
            v10 = (double)__ldg(cv + r1 + x0); v11 = (double)__ldg(cv + r1 + x1);
            cy1 = y1;
        }
        // map_coordinates order: taps row-major, each (value * wy) * wx, summed from 0.0
        double tu = 0.0, tv = 0.0;
        tu = dadd(tu, dmul(dmul(u00, wy0), wx0));
        tu = dadd(tu, dmul(dmul(u01, wy0), fx));
        tu = dadd(tu, dmul(dmul(u10, fy), wx0));
        tu = dadd(tu, dmul(dmul(u11, fy), fx));
        tv = dadd(tv, dmul(dmul(v00, wy0), wx0));
        tv = dadd(tv, dmul(dmul(v01, wy0), fx));
        tv = dadd(tv, dmul(dmul(v10, fy), wx0));
        tv = dadd(tv, dmul(dmul(v11, fy), fx));
        const unsigned o = (unsigned)(k * a.tw);
        __stcs(fu + o, fmul((float)tu, a.scale_x));  // flow scales with the resolution, float32 multiply
        __stcs(fv + o, fmul((float)tv, a.scale_y));
    }
}

// Upsampling proper (target about twice the coarse size -- every call of the pyramidal path): the
// coarse tile a CTA needs (<= 10 rows x 132 columns of u and v) is loaded once, coalesced, widened
// to float64 on the way into shared memory, and every target pixel takes its 8 taps from there:
// no global-load latency and no conversions inside the row loop.
constexpr int UPS_CR = 10, UPS_CC = 132;

__global__ void __launch_bounds__(256) upsample_flow_tile_kernel(UpArgs a) {
    __shared__ int s_y0[UP_ROWS], s_y1[UP_ROWS];
    __shared__ double s_fy[UP_ROWS], s_wy0[UP_ROWS];
    __shared__ double2 s_uv[UPS_CR * UPS_CC];  // (u, v) of every staged coarse pixel
    const int y_begin = a.row_lo + blockIdx.y * UP_ROWS, y_end = min(y_begin + UP_ROWS, a.row_hi);
    const int xb = blockIdx.x * 256;
    const int pair = blockIdx.z;
    const int cur = (a.sel ? a.sel[pair] : 0) ^ a.sel_xor;
    const size_t cplane = (size_t)a.ch * a.cw;
    const float* __restrict__ cu = (cur ? a.cu[1] : a.cu[0]) + pair * cplane;
    const float* __restrict__ cv = (cur ? a.cv[1] : a.cv[0]) + pair * cplane;
    const int yf = (int)floor(linspace_coord(y_begin, a.th, a.ch, a.step_y));  // first coarse row / column
    const int xf = (int)floor(linspace_coord(xb, a.tw, a.cw, a.step_x));
    if (threadIdx.x < UP_ROWS) {
        const int y = min(y_begin + (int)threadIdx.x, a.th - 1);
        const double yc = linspace_coord(y, a.th, a.ch, a.step_y);
        const double fy0 = floor(yc);
        const double fy = dsub(yc, fy0);
        const int y0 = (int)fy0;
        s_y0[threadIdx.x] = (y0 - yf) * UPS_CC;
        s_y1[threadIdx.x] = (min(y0 + 1, a.ch - 1) - yf) * UPS_CC;
        s_fy[threadIdx.x] = fy;
        s_wy0[threadIdx.x] = dsub(1.0, fy);
    }
    for (int i = threadIdx.x; i < UPS_CR * UPS_CC; i += 256) {
        const int r = i / UPS_CC, c = i - r * UPS_CC;
        const unsigned o = (unsigned)(min(yf + r, a.ch - 1) * a.cw + min(xf + c, a.cw - 1));
        s_uv[i] = make_double2((double)__ldg(cu + o), (double)__ldg(cv + o));
    }
    __syncthreads();
    const int x = xb + threadIdx.x;
    if (x >= a.tw) return;
    const size_t o0 = (size_t)pair * a.th * a.tw + (size_t)y_begin * a.tw + x;
    float* __restrict__ fu = a.fu + o0;
    float* __restrict__ fv = a.fv + o0;
    const double xc = linspace_coord(x, a.tw, a.cw, a.step_x);
    const double fx0 = floor(xc);
    const double fx = dsub(xc, fx0), wx0 = dsub(1.0, fx);
    const int x0 = (int)fx0 - xf;
    const int x1 = min((int)fx0 + 1, a.cw - 1) - xf;  // weight 0 when it would leave the field
    const int n = y_end - y_begin;
    // the taps stay in registers while consecutive target rows fall between the same coarse rows
    int c0 = -1, c1 = -1;
    double2 t00 = make_double2(0.0, 0.0), t01 = t00, t10 = t00, t11 = t00;
    for (int k = 0; k < n; ++k) {
        const int r0 = s_y0[k], r1 = s_y1[k];
        const double fy = s_fy[k], wy0 = s_wy0[k];
        if (r0 != c0) {  // uniform over the CTA
            if (r0 == c1) {
                t00 = t10;
                t01 = t11;
            } else {
                t00 = s_uv[r0 + x0];
                t01 = s_uv[r0 + x1];
            }
            c0 = r0;
        }
        if (r1 != c1) {
            t10 = s_uv[r1 + x0];
            t11 = s_uv[r1 + x1];
            c1 = r1;
        }
        // map_coordinates order: taps row-major, each (value * wy) * wx, summed from 0.0
        double tu = 0.0, tv = 0.0;
        tu = dadd(tu, dmul(dmul(t00.x, wy0), wx0));
        tu = dadd(tu, dmul(dmul(t01.x, wy0), fx));
        tu = dadd(tu, dmul(dmul(t10.x, fy), wx0));
        tu = dadd(tu, dmul(dmul(t11.x, fy), fx));
        tv = dadd(tv, dmul(dmul(t00.y, wy0), wx0));
        tv = dadd(tv, dmul(dmul(t01.y, wy0), fx));
        tv = dadd(tv, dmul(dmul(t10.y, fy), wx0));
        tv = dadd(tv, dmul(dmul(t11.y, fy), fx));
        const unsigned o = (unsigned)(k * a.tw);
        __stcs(fu + o, fmul((float)tu, a.scale_x));  // flow scales with the resolution, float32 multiply
        __stcs(fv + o, fmul((float)tv, a.scale_y));
    }
}

cudaError_t launch_upsample_flow(const float* cu0, const float* cv0, const float* cu1, const float* cv1,
                                 const int* sel, int sel_xor, float* fu, float* fv, int batch, int ch, int cw, int th,
                                 int tw, int row_lo, int row_hi, int* launches, cudaStream_t stream) {
    if (row_lo < 0 || row_hi > th || row_lo >= row_hi) return cudaErrorInvalidValue;
    UpArgs a;
    a.cu[0] = cu0;
    a.cv[0] = cv0;
    a.cu[1] = cu1 ? cu1 : cu0;
    a.cv[1] = cv1 ? cv1 : cv0;
    a.sel = sel;
    a.sel_xor = sel_xor;
    a.fu = fu;
    a.fv = fv;
    a.ch = ch;
    a.cw = cw;
    a.th = th;
    a.tw = tw;
    a.row_lo = row_lo;
    a.row_hi = row_hi;
    a.step_y = th > 1 ? (double)(ch - 1) / (double)(th - 1) : 0.0;
    a.step_x = tw > 1 ? (double)(cw - 1) / (double)(tw - 1) : 0.0;
    a.scale_y = (float)((double)th / (double)ch);
    a.scale_x = (float)((double)tw / (double)cw);
    if (launches) *launches += 1;
    dim3 grid((tw + 255) / 256, (row_hi - row_lo + UP_ROWS - 1) / UP_ROWS, batch);
    // the staged tile covers 16 target rows x 256 target columns when the grid steps are <= ~0.5
    if (15.0 * a.step_y + 1.0 < UPS_CR - 1 && 255.0 * a.step_x + 1.0 < UPS_CC - 1)
        OF_LAUNCH(upsample_flow_tile_kernel, grid, 256, 0, stream, a);
    else
        OF_LAUNCH(upsample_flow_kernel, grid, 256, 0, stream, a);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) select_copy_kernel(const float* __restrict__ u0, const float* __restrict__ v0,
                                                           const float* __restrict__ u1, const float* __restrict__ v1,
                                                           const int* __restrict__ sel, int sel_xor,
                                                           float* __restrict__ out_u, float* __restrict__ out_v,
                                                           size_t n) {
    const int pair = blockIdx.y;
    const int cur = (sel ? sel[pair] : 0) ^ sel_xor;
    const float* su = (cur ? u1 : u0) + pair * n;
    const float* sv = (cur ? v1 : v0) + pair * n;
    if (su == out_u + pair * n) return;  // already in place
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        out_u[pair * n + i] = su[i];
        out_v[pair * n + i] = sv[i];
    }
}

cudaError_t launch_select_copy(const float* u0, const float* v0, const float* u1, const float* v1, const int* sel,
                               int sel_xor, float* out_u, float* out_v, int batch, size_t n, int* launches,
                               cudaStream_t stream) {
    if (launches) *launches += 1;
    unsigned bx = (unsigned)((n + 256 * 8 - 1) / (256 * 8));
    if (bx < 1) bx = 1;
    if (bx > 4096) bx = 4096;
    dim3 grid(bx, batch);
    OF_LAUNCH(select_copy_kernel, grid, 256, 0, stream, u0, v0, u1, v1, sel, sel_xor, out_u, out_v, n);
    return cudaGetLastError();
}

}  // namespace ofb
