// warp_image on the device for the split refinement iteration (python/lucas_kanade_pyramidal.py:66-97):
// the sample / blend helpers and warp_rows_kernel<F>.  Included by lk_march.cu (which launches the kernel) and,
// compiled by g++ on top of tests/host_emul/cuda_on_host.h, by the CPU test that runs this source against the
// reference's warp_image.
#pragma once
#include "of_common.cuh"

namespace ofb {

// One bilinear sample of warp_image, split in two so that a lane can put the loads of all
// its samples in flight before blending any of them (the blend needs ~40 dependent cycles,
// a gather from L2 several hundred).  Branch-free: indices are clamped into the frame and the
// result is zeroed afterwards when the sample lies outside.
template <typename F>
struct WarpTapT {
    float v00, v01, v10, v11;
    F fy, fx;
    bool inside;
};
using WarpTap = WarpTapT<float>;

template <typename F>
__device__ __forceinline__ float warp_blend(const WarpTapT<F>& t) {
    // float64 blend in SciPy's order: taps row-major, each (value * wy) * wx, summed from 0.0
    const double wy1 = (double)t.fy, wx1 = (double)t.fx;
    const double wy0 = dsub(1.0, wy1), wx0 = dsub(1.0, wx1);
    double acc = 0.0;
    acc = dadd(acc, dmul(dmul((double)t.v00, wy0), wx0));
    acc = dadd(acc, dmul(dmul((double)t.v01, wy0), wx1));
    acc = dadd(acc, dmul(dmul((double)t.v10, wy1), wx0));
    acc = dadd(acc, dmul(dmul((double)t.v11, wy1), wx1));
    return t.inside ? (float)acc : 0.0f;
}

// warp_image for the split refinement iteration: warped[y][x] = bilinear(curr, y + v, x + u) for
// rows [row_lo, row_hi) of every pair that has not converged; 4 pixels per thread (128-bit flow
// loads and stores, 16 gathers in flight).  Same arithmetic as the fused kernel's gather.
struct WarpRowsArgs {
    const float* curr;
    const float* flow_u[2];
    const float* flow_v[2];
    const int* sel;
    int sel_xor;
    const int* done;
    float* warped;
    int H, W, row_lo, row_hi;
};

constexpr int WR_PER_THREAD = 4;  // samples per thread, 256 columns apart: every gather of a
                                  // warp touches 32 adjacent pixels (coalesced), 16 loads in flight

// The same sample as warp_gather with fewer issue slots: floor() and the float -> int conversion
// (FRND + F2I, quarter-rate XU pipe) become one round-down add of 1.5 * 2^23, whose low mantissa
// bits are floor(v) for |v| < 2^22; the four taps are addressed by 32-bit element offsets from one
// base.  |v| >= 2^22 (and NaN) is "outside", as it is for the reference's float64 coordinates.
//
// The fraction v - floor(v) is exact in float32 for v >= 0; for v < 0 it can need one or two bits
// more than float32 has (-0.3 + 1), so the float32 fraction (F = float, fast mode) may be one
// rounding away from the reference's float64 fraction.  F = double takes the difference in
// float64, where it is always exact; that is the reference's fraction whenever the reference's
// own float64 coordinate y + v is exact (warp_rows_kernel<double> checks that, see there).
template <typename F>
__device__ __forceinline__ F warp_fraction(float v, float floor_v) {
    return (F)v - (F)floor_v;
}

// Where one sample's four taps lie (32-bit element offsets from the frame's base), its fractions and whether it
// lies inside the frame: the address half of a sample, shared by every kernel that warps (the loads differ:
// registers here, cp.async into shared memory in the marching kernel's epilogue).
template <typename F>
struct WarpAddrT {
    unsigned o00, o01, o10, o11;
    F fy, fx;
    bool inside;
};

template <typename F>
__device__ __forceinline__ void warp_address_magic(int H, int W, int yc, int xc, float v, float u, WarpAddrT<F>& t) {
    const float magic = 12582912.0f;  // 0x4B400000
    const float tv = __fadd_rd(v, magic), tu = __fadd_rd(u, magic);
    t.fy = warp_fraction<F>(v, tv - magic);
    t.fx = warp_fraction<F>(u, tu - magic);
    const int y0 = yc + (__float_as_int(tv) - 0x4B400000);
    const int x0 = xc + (__float_as_int(tu) - 0x4B400000);
    const bool sane = (fabsf(v) < 4194304.0f) & (fabsf(u) < 4194304.0f);
    // 0 <= y0 + fy <= H - 1 and 0 <= x0 + fx <= W - 1   (bitwise ops: no short-circuit branches)
    const bool in_y = ((unsigned)y0 < (unsigned)(H - 1)) | ((y0 == H - 1) & (t.fy == (F)0));
    const bool in_x = ((unsigned)x0 < (unsigned)(W - 1)) | ((x0 == W - 1) & (t.fx == (F)0));
    t.inside = in_y & in_x & sane;
    const int ys = min(max(y0, 0), H - 1), xs = min(max(x0, 0), W - 1);
    // The tap past the last row / column has weight exactly 0 (SciPy mirrors its index there);
    // any finite in-frame value gives the same sum, so it simply re-reads the last one.
    t.o00 = (unsigned)(ys * W + xs);
    t.o01 = t.o00 + ((xs < W - 1) ? 1u : 0u);
    const unsigned dy = (ys < H - 1) ? (unsigned)W : 0u;
    t.o10 = t.o00 + dy;
    t.o11 = t.o01 + dy;
}

// N samples of one lane on row y (columns xc[k], flow (lu[k], lv[k])).  Integer / fraction split of every sample
// first; if the 2x2 taps of ALL samples of the warp lie strictly inside the frame (the common case away from the
// border and for moderate flow), the addresses are formed without clamps, edge rules and the final select -- about
// a seventh fewer instructions.  Every lane of the warp must call it (__all_sync).
template <typename F, int N>
__device__ __forceinline__ void warp_address_n(int H, int W, int y, const int (&xc)[N], const float (&lu)[N],
                                               const float (&lv)[N], WarpAddrT<F> (&t)[N]) {
    const float magic = 12582912.0f;  // 1.5 * 2^23 (see warp_address_magic)
    int sy[N], sx[N];
    bool interior = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        sy[k] = y + (__float_as_int(__fadd_rd(lv[k], magic)) - 0x4B400000);
        sx[k] = xc[k] + (__float_as_int(__fadd_rd(lu[k], magic)) - 0x4B400000);
        interior &= ((unsigned)sy[k] < (unsigned)(H - 1)) & ((unsigned)sx[k] < (unsigned)(W - 1)) &
                    (fabsf(lv[k]) < 4194304.0f) & (fabsf(lu[k]) < 4194304.0f);
    }
    if (__all_sync(0xffffffffu, interior)) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            t[k].fy = warp_fraction<F>(lv[k], __fadd_rd(lv[k], magic) - magic);
            t[k].fx = warp_fraction<F>(lu[k], __fadd_rd(lu[k], magic) - magic);
            t[k].inside = true;
            t[k].o00 = (unsigned)(sy[k] * W + sx[k]);
            t[k].o01 = t[k].o00 + 1u;
            t[k].o10 = t[k].o00 + (unsigned)W;
            t[k].o11 = t[k].o00 + (unsigned)W + 1u;
        }
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) warp_address_magic<F>(H, W, y, xc[k], lv[k], lu[k], t[k]);
    }
}

// the loads of those samples into registers, all in flight together; one widening multiply-add per address
// (IMAD.WIDE.U32) instead of a 64-bit add + shift pair
template <typename F>
__device__ __forceinline__ void warp_load_taps(const float* __restrict__ img, const WarpAddrT<F>& a, WarpTapT<F>& t) {
    const char* base = reinterpret_cast<const char*>(img);
    t.fy = a.fy;
    t.fx = a.fx;
    t.inside = a.inside;
    t.v00 = __ldg(reinterpret_cast<const float*>(base + (size_t)a.o00 * 4u));
    t.v01 = __ldg(reinterpret_cast<const float*>(base + (size_t)a.o01 * 4u));
    t.v10 = __ldg(reinterpret_cast<const float*>(base + (size_t)a.o10 * 4u));
    t.v11 = __ldg(reinterpret_cast<const float*>(base + (size_t)a.o11 * 4u));
}

template <typename F>
__device__ __forceinline__ void warp_gather_magic(const float* __restrict__ img, int H, int W, int yc, int xc, float v,
                                                  float u, WarpTapT<F>& t) {
    WarpAddrT<F> a;
    warp_address_magic<F>(H, W, yc, xc, v, u, a);
    warp_load_taps<F>(img, a, t);
}

template <typename F, int N>
__device__ __forceinline__ void warp_gather_n(const float* __restrict__ img, int H, int W, int y, const int (&xc)[N],
                                              const float (&lu)[N], const float (&lv)[N], WarpTapT<F> (&t)[N]) {
    WarpAddrT<F> a[N];
    warp_address_n<F, N>(H, W, y, xc, lu, lv, a);
#pragma unroll
    for (int k = 0; k < N; ++k) warp_load_taps<F>(img, a[k], t[k]);
}

template <typename F>
__global__ void __launch_bounds__(256) warp_rows_kernel(WarpRowsArgs a) {
    const int pair = blockIdx.z;
    if (a.done != nullptr && a.done[pair]) return;
    const int y = a.row_lo + blockIdx.y;
    const int x0 = blockIdx.x * (256 * WR_PER_THREAD) + threadIdx.x;
    const int cur = (a.sel ? a.sel[pair] : 0) ^ a.sel_xor;
    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    const float* __restrict__ img = a.curr + pair * plane;
    // per-pair bases once, then 32-bit element offsets (H * W < 2^31) and one widening multiply-add per address:
    // 64-bit index arithmetic per load was a twelfth of the kernel's instructions
    const char* fub = reinterpret_cast<const char*>((cur ? a.flow_u[1] : a.flow_u[0]) + pair * plane);
    const char* fvb = reinterpret_cast<const char*>((cur ? a.flow_v[1] : a.flow_v[0]) + pair * plane);
    char* outb = reinterpret_cast<char*>(a.warped + pair * plane);
    const unsigned r0 = (unsigned)y * (unsigned)W;
    float* __restrict__ out = reinterpret_cast<float*>(outb + (size_t)r0 * 4u);
    float lu[WR_PER_THREAD], lv[WR_PER_THREAD];
#pragma unroll
    for (int k = 0; k < WR_PER_THREAD; ++k) {
        const unsigned xs = (unsigned)min(x0 + 256 * k, W - 1);  // keep the loads in range; the store is predicated
        lu[k] = __ldg(reinterpret_cast<const float*>(fub + (size_t)(r0 + xs) * 4u));
        lv[k] = __ldg(reinterpret_cast<const float*>(fvb + (size_t)(r0 + xs) * 4u));
    }
    if (sizeof(F) == 8) {
        // Exact flavour.  The reference's coordinate is the float64 sum y + v (lucas_kanade_pyramidal.py:88-92),
        // which is itself ROUNDED when v has bits below the sum's last place: coordinates below 2^16 keep
        // bits down to 2^-37, so the sum is exact -- and the integer / fraction split below is the reference's --
        // iff v == 0 or |v| >= 2^-14.  A warp holding any other flow value (tiny non-zero flow, NaN) takes
        // the rounded sum through bilinear_f64, the sample routine of warp_kernel / lk_tile_kernel<SRC_WARP>.
        bool exact_sum = (H <= 65536) & (W <= 65536);
#pragma unroll
        for (int k = 0; k < WR_PER_THREAD; ++k)
            exact_sum &= ((lv[k] == 0.0f) | (fabsf(lv[k]) >= 6.103515625e-05f)) &
                         ((lu[k] == 0.0f) | (fabsf(lu[k]) >= 6.103515625e-05f));
        if (!__all_sync(0xffffffffu, exact_sum)) {
#pragma unroll
            for (int k = 0; k < WR_PER_THREAD; ++k) {
                const int x = x0 + 256 * k;
                if (x < W) __stcs(out + x, bilinear_f64(img, H, W, dadd((double)y, (double)lv[k]), dadd((double)x, (double)lu[k])));
            }
            return;
        }
    }
    WarpTapT<F> t[WR_PER_THREAD];
    int xc[WR_PER_THREAD];
#pragma unroll
    for (int k = 0; k < WR_PER_THREAD; ++k) xc[k] = min(x0 + 256 * k, W - 1);
    warp_gather_n<F, WR_PER_THREAD>(img, H, W, y, xc, lu, lv, t);
#pragma unroll
    for (int k = 0; k < WR_PER_THREAD; ++k) {
        const int x = x0 + 256 * k;
        if (x < W) __stcs(out + x, warp_blend(t[k]));
    }
}

}  // namespace ofb
