// K1 / K3 (exact), window 5, marching form: the reference's operation order (bit-identical on any float32
// input, like lk_tile5.cu / lk_tile.cu) in a kernel that walks down the image instead of tiling it.
//
//   SRC_FRAMES  lucas_kanade_single_scale  (python/lucas_kanade_core.py:15-45, 48-70, 73-135)
//   SRC_WARPED  one refinement iteration of lucas_kanade_pyramidal on (prev, warped curr): flow_out = flow_in + d
//               and per-unit sums of |du|, |dv| (python/lucas_kanade_pyramidal.py:203-214)
//
// Why another kernel.  The tile kernels are issue-bound, and 40 % of what lk_tile5_kernel issues (100 of 265
// lane-instructions per pixel, profiles/r01c) is not the arithmetic np.sum's order forces but tile overhead:
// every 16 x 64 tile re-stages a 22 x 70 frame tile and re-derives a 20 x 68 product tile (1.5 / 1.33 pixels of
// work per output), and every scalar float32 operation takes a full issue slot.  Here
//   * a unit owns two adjacent 58-column strips and marches down a band of rows: every
//     frame value is loaded, averaged and scaled once, every gradient and product formed once (the only
//     redundancy is the 3-column halo of a strip, 6 / 64);
//   * lane L holds columns 2L, 2L + 1 of BOTH strips, packed: one 64-bit word = the same column of strip A and
//     strip B.  Every Sobel tap, product and window-sum addition is a Blackwell packed-pair instruction
//     (FADD2 / FMUL2: one issue slot, two independent IEEE operations) on naturally aligned operands;
//   * the rows in flight live in shared-memory rings (8 scaled-frame rows, 8 product rows), so a lane reads its
//     neighbours' columns directly -- no shuffles;
//   * the rings (34.8 KB) cap an SM at six units, so the latency has to be hidden inside the unit: a unit is a CTA of
//     TWO warps sharing one set of rings, a step handles FOUR rows (two per warp) between two barriers; a warp's two
//     output rows are two independent sets of accumulation chains and share every shared-memory load of the six
//     product rows they read (3.75 instead of 6.25 128-bit loads per output); frame and flow loads are issued at the
//     top of a step and consumed at its end.
// Per output the operations and their order are exactly lk_tile5_kernel's (scaled frame planes E = avg * 0.125,
// D = avg * 0.25 so that a Sobel tap is one addition; zero taps contribute value * 0.0; the 25 products summed
// like np.sum: 8 running lanes, tree, tail, + 0.0; Cramer without FMA), so the bits are the reference's.
#include <cuda_runtime.h>

#include "of_common.cuh"
#include "f32x2.cuh"
#include "of_kernels.h"
#include "peer_device.cuh"

namespace ofb {
namespace {

constexpr int XM_WARPS = 2;                 // a unit (= CTA) is two warps that share one set of rings
constexpr int XM_COLS = 64;                 // loaded columns per strip (2 per lane)
constexpr int XM_OUT = XM_COLS - 6;         // 58 outputs per strip: Sobel 1 + window 2 columns of halo on both sides
constexpr int XM_PAD = 2;                   // ring rows start 2 words in: word XM_PAD + j = local column j
constexpr int XM_PITCH = XM_COLS + 4;       // words per ring row (one readable halo word on both sides, 16-byte rows)
constexpr int XM_FRING = 8, XM_PRING = 8;   // ring depths (powers of two): scaled-frame rows, product rows
constexpr int XM_STEP = 2 * XM_WARPS;       // rows per step: two per warp
constexpr int XM_EXTRA = 12;                // rows' worth of steps a band spends before its first output row (3 steps x 4 rows)
constexpr int XM_UNIT_WORDS = (3 * XM_FRING + 5 * XM_PRING) * XM_PITCH;
constexpr size_t XM_SMEM_BYTES = (size_t)XM_UNIT_WORDS * sizeof(f32x2);  // 34 816
constexpr int XM_UNITS_PER_SM = 6;          // 209 KB of shared memory, 12 warps, <= 170 registers per thread

struct XmArgs {
    TileArgs t;
    int n_pairs_of_strips, n_bands, band_rows;
    long long n_units;
    int slots_per_pair;  // lk_tile_blocks_per_pair(rows, W): the partial-sum slots the iteration's tail reads
};

// np.sum's order as a streaming accumulator on packed pairs (see lk_tile5.cu)
struct Np25x2 {
    f32x2 lane[8];
    f32x2 res;
};
__device__ __forceinline__ void np25_add(Np25x2& s, int t, f32x2 p) {
    if (t < 8) {
        s.lane[t] = p;
    } else if (t < 24) {
        s.lane[t & 7] = add2(s.lane[t & 7], p);
    } else {
        s.res = add2(add2(add2(s.lane[0], s.lane[1]), add2(s.lane[2], s.lane[3])),
                     add2(add2(s.lane[4], s.lane[5]), add2(s.lane[6], s.lane[7])));
        s.res = add2(s.res, p);
    }
}

template <int SRC>
__global__ void __launch_bounds__(XM_WARPS * 32, XM_UNITS_PER_SM) lk_exact_march_kernel(const XmArgs xa) {
    constexpr bool FLOW = (SRC == SRC_WARPED);
    OF_DYNAMIC_SMEM_ALIGNED(16, unsigned char, xm_smem);
    const TileArgs& a = xa.t;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long unit = blockIdx.x;  // one unit per CTA: everything below up to the row loop is CTA-uniform
    const int sp = (int)(unit % xa.n_pairs_of_strips);
    const long long rest = unit / xa.n_pairs_of_strips;
    const int band = (int)(rest % xa.n_bands);
    const int pair = (int)(rest / xa.n_bands);
    if (FLOW && a.done != nullptr && a.done[pair]) return;  // level already converged

    f32x2* ring = reinterpret_cast<f32x2*>(xm_smem);
    // The scaled taps are formed in stage A and go through shared memory: ptxas fuses a packed mul.rn.f32x2 into a
    // packed addition that consumes it (it honours .rn only for scalars; it also sees through fma(x, k, -0.0)), and an
    // unrounded avg * k differs from the reference's rounded tap product when it is subnormal.  A store in between
    // is the one thing it cannot fuse across.
    f32x2* sE = ring;                               // avg * 0.125  [XM_FRING][XM_PITCH], word = (strip A, strip B)
    f32x2* sD = sE + XM_FRING * XM_PITCH;           // avg * 0.25
    f32x2* sT = sD + XM_FRING * XM_PITCH;           // It = p - c
    f32x2* prod = sT + XM_FRING * XM_PITCH;         // xx, yy, xy, xt, yt  [5][XM_PRING][XM_PITCH]

    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    const int row_lo = FLOW ? a.row_lo : 0, row_hi = FLOW ? a.row_hi : H;
    const int y0 = row_lo + band * xa.band_rows;
    const int y1 = min(y0 + xa.band_rows, row_hi);
    // image column of local column j: strip A (2 sp) starts at 58 * 2 sp - 3, strip B 58 columns further right
    const int xA = XM_OUT * 2 * sp - 3 + 2 * lane, xB = xA + XM_OUT;
    const float* __restrict__ gp = a.in0 + pair * plane;
    const float* __restrict__ gc = a.in1 + pair * plane;
    // replicated (= 'symm' for a 3 x 3 kernel) border: clamped coordinates
    const int cA0 = clampi(xA, 0, W - 1), cA1 = clampi(xA + 1, 0, W - 1);
    const int cB0 = clampi(xB, 0, W - 1), cB1 = clampi(xB + 1, 0, W - 1);

    const float* fin_u = nullptr;
    const float* fin_v = nullptr;
    float* out_u;
    float* out_v;
    if (FLOW) {
        const int cur = (a.sel ? a.sel[pair] : 0) ^ a.sel_xor;
        fin_u = (cur ? a.flow_u[1] : a.flow_u[0]) + pair * plane;
        fin_v = (cur ? a.flow_v[1] : a.flow_v[0]) + pair * plane;
        out_u = (cur ? a.flow_u[0] : a.flow_u[1]) + pair * plane;
        out_v = (cur ? a.flow_v[0] : a.flow_v[1]) + pair * plane;
    } else {
        out_u = a.out_u + pair * plane;
        out_v = a.out_v + pair * plane;
    }
    // which of this lane's four outputs per row exist: local columns 3 .. 60 of a strip, inside the frame
    // (strip B of the last pair of strips may lie beyond it)
    bool emit[2][2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const bool local_ok = (2 * lane + k >= 3) && (2 * lane + k < 3 + XM_OUT);
        emit[0][k] = local_ok && (xA + k < W);
        emit[1][k] = local_ok && (xB + k < W);
    }
    const f32x2 zero2 = pk(0.0f, 0.0f), half2 = pk(0.5f, 0.5f), k125 = pk(0.125f, 0.125f), k25 = pk(0.25f, 0.25f);
    const int wl = XM_PAD + 2 * lane;  // this lane's first word in a ring row
    double acc_u = 0.0, acc_v = 0.0;

    // Four rows per step, two per warp.  Step `it` (frame rows fr .. fr + 3, fr = fr0 + 4 it); warp w takes rows 2 w, 2 w + 1
    // of every stage:
    //   1. the frame values of its rows fr + 2 w, fr + 2 w + 1 are requested;
    //   2. stage B forms the products of its gradient rows g + 2 w, g + 2 w + 1 (g = fr - 5) from frame rows staged by
    //      earlier steps                                                                               __syncthreads()
    //   3. stage C finishes its output rows o + 2 w, o + 2 w + 1 (o = fr - 7) from six product rows (the newest written
    //      in 2. by either warp): the two rows share every shared-memory load and run as independent chains;
    //   4. stage A turns the values of 1. into E, D, T rows (over ring slots 2. has finished with)  __syncthreads()
    // Frame row fr0 + j lives in ring slot j & 7, the product row of gradient row y0 - 2 + p in slot p & 7.
    const int fr0 = y0 - 5;
    const int n_steps = (y1 - y0 + XM_STEP - 1) / XM_STEP + 3;  // outputs start at step 3
    // output pointers of this lane's first column of either strip at this warp's first row, advanced per emitting step
    const size_t o_first = (size_t)(y0 + 2 * warp) * W;
    float* pu[2] = {out_u + o_first + xA, out_u + o_first + xB};
    float* pv[2] = {out_v + o_first + xA, out_v + o_first + xB};
    const float* qu[2] = {FLOW ? fin_u + o_first + xA : nullptr, FLOW ? fin_u + o_first + xB : nullptr};
    const float* qv[2] = {FLOW ? fin_v + o_first + xA : nullptr, FLOW ? fin_v + o_first + xB : nullptr};
    for (int it = 0; it < n_steps; ++it) {
        const int jw = XM_STEP * it + 2 * warp;  // ring index of this warp's first frame row of the step
        // ---- 1. frame values of this warp's two rows (consumed in 4.: a whole step's arithmetic covers the latency)
        float fv[2][8];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            // 32-bit element offsets (H * W < 2^31, launcher) and one widening multiply-add per address
            const unsigned ro = (unsigned)clampi(fr0 + jw + rr, 0, H - 1) * (unsigned)W;
            const char* bp = reinterpret_cast<const char*>(gp);
            const char* bc = reinterpret_cast<const char*>(gc);
            fv[rr][0] = __ldg(reinterpret_cast<const float*>(bp + (size_t)(ro + (unsigned)cA0) * 4u));
            fv[rr][1] = __ldg(reinterpret_cast<const float*>(bp + (size_t)(ro + (unsigned)cA1) * 4u));
            fv[rr][2] = __ldg(reinterpret_cast<const float*>(bp + (size_t)(ro + (unsigned)cB0) * 4u));
            fv[rr][3] = __ldg(reinterpret_cast<const float*>(bp + (size_t)(ro + (unsigned)cB1) * 4u));
            fv[rr][4] = __ldg(reinterpret_cast<const float*>(bc + (size_t)(ro + (unsigned)cA0) * 4u));
            fv[rr][5] = __ldg(reinterpret_cast<const float*>(bc + (size_t)(ro + (unsigned)cA1) * 4u));
            fv[rr][6] = __ldg(reinterpret_cast<const float*>(bc + (size_t)(ro + (unsigned)cB0) * 4u));
            fv[rr][7] = __ldg(reinterpret_cast<const float*>(bc + (size_t)(ro + (unsigned)cB1) * 4u));
        }
        // the flow this warp's two output rows are added to (refinement iteration): requested now, used at the end of 3.
        const int o = y0 + XM_STEP * (it - 3) + 2 * warp;  // this warp's first output row of the step
        const bool emit_rows = (it >= 3 && o < y1);
        const bool second = (o + 1 < y1);  // the band may end on this warp's first row
        float fin[2][2][2][2];             // [row][half][column][u, v]
        if (FLOW && emit_rows) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                for (int half = 0; half < 2; ++half)
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const bool live = emit[half][k] && (rr == 0 || second);
                        const size_t po = (size_t)rr * W + k;
                        fin[rr][half][k][0] = live ? __ldg(qu[half] + po) : 0.0f;
                        fin[rr][half][k][1] = live ? __ldg(qv[half] + po) : 0.0f;
                    }
        }
        // ---- 2. stage B: this warp's gradient rows (ring indices jw - 5, jw - 4) from frame rows jw - 6 .. jw - 3: Sobel
        // in kernel order (j, k) -- tap (j, k) reads frame offset (1 - j, 1 - k) from the centre -- and the five products
        if (it >= 2) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                // 128-bit loads of aligned word pairs (conflict-free at the lanes' 16-byte stride): words wl - 2 .. wl + 3
                // of the top / middle / bottom rows, of which wl - 1 .. wl + 2 are the taps of this lane's two columns
                const int jg = jw - 5 + rr;  // ring index of the gradient row's own frame row
                const int s_top = ((jg - 1) & (XM_FRING - 1)) * XM_PITCH + wl - 2;
                const int s_mid = (jg & (XM_FRING - 1)) * XM_PITCH + wl - 2;
                const int s_bot = ((jg + 1) & (XM_FRING - 1)) * XM_PITCH + wl - 2;
                auto load4 = [](const f32x2* row, f32x2* dst) {  // words 1 .. 4 of the six at `row`
                    const ulonglong2* r2 = reinterpret_cast<const ulonglong2*>(row);
                    const ulonglong2 a0 = r2[0], a1 = r2[1], a2 = r2[2];
                    dst[0] = a0.y;
                    dst[1] = a1.x;
                    dst[2] = a1.y;
                    dst[3] = a2.x;
                };
                f32x2 E0[4], E2[4], D1[4];
                load4(sE + s_top, E0);
                load4(sE + s_bot, E2);
                load4(sD + s_mid, D1);
                const ulonglong2 d0 = *reinterpret_cast<const ulonglong2*>(sD + s_top + 2);
                const ulonglong2 d2 = *reinterpret_cast<const ulonglong2*>(sD + s_bot + 2);
                const ulonglong2 t1 = *reinterpret_cast<const ulonglong2*>(sT + s_mid + 2);
                const f32x2 D0m[2] = {d0.x, d0.y}, D2m[2] = {d2.x, d2.y}, Tm[2] = {t1.x, t1.y};
                f32x2 pr[5][2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {  // this lane's two columns: taps at words c (left), c + 1, c + 2 (right)
                    // zero taps: value * 0.0f (keeps the reference's signed zeros / NaN propagation)
                    const f32x2 Z0 = mul2(D0m[c], zero2), Z2 = mul2(D2m[c], zero2);
                    const f32x2 Z1lo = mul2(D1[c], zero2), Z1m = mul2(D1[c + 1], zero2), Z1hi = mul2(D1[c + 2], zero2);
                    f32x2 ax = zero2, ay = zero2;
                    // j = 0 (frame row g + 1): kx = -.125, 0, .125   ky = -.125, -.25, -.125
                    ax = sub2(ax, E2[c + 2]);  ay = sub2(ay, E2[c + 2]);
                    ax = add2(ax, Z2);         ay = sub2(ay, D2m[c]);
                    ax = add2(ax, E2[c]);      ay = sub2(ay, E2[c]);
                    // j = 1 (frame row g): kx = -.25, 0, .25          ky = 0, 0, 0
                    ax = sub2(ax, D1[c + 2]);  ay = add2(ay, Z1hi);
                    ax = add2(ax, Z1m);        ay = add2(ay, Z1m);
                    ax = add2(ax, D1[c]);      ay = add2(ay, Z1lo);
                    // j = 2 (frame row g - 1): kx = -.125, 0, .125   ky = .125, .25, .125
                    ax = sub2(ax, E0[c + 2]);  ay = add2(ay, E0[c + 2]);
                    ax = add2(ax, Z0);         ay = add2(ay, D0m[c]);
                    ax = add2(ax, E0[c]);      ay = add2(ay, E0[c]);
                    pr[0][c] = mul2(ax, ax);
                    pr[1][c] = mul2(ay, ay);
                    pr[2][c] = mul2(ax, ay);
                    pr[3][c] = mul2(ax, Tm[c]);
                    pr[4][c] = mul2(ay, Tm[c]);
                }
                // gradient row y0 - 2 + p with p = jw - 8 + rr (jg = jw - 5 + rr is frame row y0 - 5 + jg)
                f32x2* pw = prod + ((jw - 8 + rr) & (XM_PRING - 1)) * XM_PITCH + wl;
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    *reinterpret_cast<ulonglong2*>(pw + q * (XM_PRING * XM_PITCH)) = make_ulonglong2(pr[q][0], pr[q][1]);
            }
        }
        __syncthreads();
        // ---- 3. stage C: this warp's output rows o, o + 1 from product rows o - 2 .. o + 3 (indices jw - 12 .. jw - 7)
        if (emit_rows) {
            f32x2 sum[2][5][2];                // [row][quantity][column]
            int rs[6];                         // ring offsets of product rows o - 2 + i
#pragma unroll
            for (int i = 0; i < 6; ++i) rs[i] = ((jw - 12 + i) & (XM_PRING - 1)) * XM_PITCH + wl - 2;  // window of column j: words j - 2 .. j + 2
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const f32x2* P = prod + q * (XM_PRING * XM_PITCH);
                Np25x2 s0[2], s1[2];  // windows of row o / row o + 1, this lane's two columns
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const ulonglong2* row = reinterpret_cast<const ulonglong2*>(P + rs[i]);
                    const ulonglong2 q0 = row[0], q1 = row[1], q2 = row[2];
                    const f32x2 v[6] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y};
#pragma unroll
                    for (int w = 0; w < 2; ++w)
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            if (i < 5) np25_add(s0[w], 5 * i + k, v[w + k]);
                            if (i >= 1) np25_add(s1[w], 5 * (i - 1) + k, v[w + k]);
                        }
                }
#pragma unroll
                for (int w = 0; w < 2; ++w) {  // np.add.reduce starts from +0.0
                    sum[0][q][w] = add2(zero2, s0[w].res);
                    sum[1][q][w] = add2(zero2, s1[w].res);
                }
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                if (rr == 1 && !second) break;
                const int y = o + rr;
                const bool row_inside = (y >= 2 && y < H - 2);
                const bool row_owned = FLOW && (y >= a.own_lo && y < a.own_hi);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    // Cramer's determinants for both strips at once (lucas_kanade_core.py:122-133; b0 = -sxt, b1 = -syt and
                    // negation commutes with rounding): every product rounds on its own, then the difference.  Written as
                    // fma(c * d, -1, a * b): ptxas would fuse a packed multiply into a packed subtraction.
                    const f32x2 m1 = pk(-1.0f, -1.0f);
                    const f32x2 sxx = sum[rr][0][k], syy = sum[rr][1][k], sxy = sum[rr][2][k], sxt = sum[rr][3][k], syt = sum[rr][4][k];
                    const f32x2 det2 = fma2(mul2(sxy, sxy), m1, mul2(sxx, syy));
                    const f32x2 nu2 = fma2(mul2(syy, sxt), m1, mul2(sxy, syt));
                    const f32x2 nv2 = fma2(mul2(sxx, syt), m1, mul2(sxy, sxt));
                    float det[2], nu[2], nv[2];
                    unpk(det2, det[0], det[1]);
                    unpk(nu2, nu[0], nu[1]);
                    unpk(nv2, nv[0], nv[1]);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (!emit[half][k]) continue;
                        const int x = (half ? xB : xA) + k;
                        // branch-free: the division runs on a safe denominator, the border / singular case selects 0
                        const bool ok = row_inside && x >= 2 && x < W - 2 && (fabsf(det[half]) > OF_DET_EPS);
                        const float den = ok ? det[half] : 1.0f;
                        const float qu_ = fdiv(nu[half], den), qv_ = fdiv(nv[half], den);
                        const float u = ok ? qu_ : 0.0f, v = ok ? qv_ : 0.0f;
                        const size_t po = (size_t)rr * W + k;
                        if (FLOW) {
                            pu[half][po] = fadd(fin[rr][half][k][0], u);  // flow += d
                            pv[half][po] = fadd(fin[rr][half][k][1], v);
                            if (row_owned) {
                                acc_u += (double)fabsf(u);
                                acc_v += (double)fabsf(v);
                            }
                        } else {
                            pu[half][po] = u;
                            pv[half][po] = v;
                        }
                    }
                }
            }
        }
        if (it >= 3) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                pu[half] += XM_STEP * (size_t)W;
                pv[half] += XM_STEP * (size_t)W;
                if (FLOW) {
                    qu[half] += XM_STEP * (size_t)W;
                    qv[half] += XM_STEP * (size_t)W;
                }
            }
        }
        // ---- 4. stage A: this warp's frame rows -> E, D, T (ring slots jw & 7, (jw + 1) & 7)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const f32x2 p0 = pk(fv[rr][0], fv[rr][2]), p1 = pk(fv[rr][1], fv[rr][3]);
            const f32x2 c0 = pk(fv[rr][4], fv[rr][6]), c1 = pk(fv[rr][5], fv[rr][7]);
            const f32x2 avg0 = mul2(add2(p0, c0), half2), avg1 = mul2(add2(p1, c1), half2);  // (p + c) / 2.0
            const int w = ((jw + rr) & (XM_FRING - 1)) * XM_PITCH + wl;
            *reinterpret_cast<ulonglong2*>(sE + w) = make_ulonglong2(mul2(avg0, k125), mul2(avg1, k125));
            *reinterpret_cast<ulonglong2*>(sD + w) = make_ulonglong2(mul2(avg0, k25), mul2(avg1, k25));
            *reinterpret_cast<ulonglong2*>(sT + w) = make_ulonglong2(sub2(p0, c0), sub2(p1, c1));
        }
        __syncthreads();
    }

    if (FLOW && a.partial != nullptr) {
        // fixed shuffle tree over each warp, then warp 0 + warp 1; the unit's sums go to its slot, and the slots no unit
        // owns (the tail of the iteration reads lk_tile_blocks_per_pair(rows, W) of them) are cleared by the units in turn
        __shared__ double red[2][XM_WARPS];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            acc_u += __shfl_down_sync(0xffffffffu, acc_u, off);
            acc_v += __shfl_down_sync(0xffffffffu, acc_v, off);
        }
        if (lane == 0) {
            red[0][warp] = acc_u;
            red[1][warp] = acc_v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int units_per_pair = xa.n_bands * xa.n_pairs_of_strips;
            const int me = band * xa.n_pairs_of_strips + sp;
            double* part = a.partial + (size_t)pair * xa.slots_per_pair * 2;
            part[2 * me + 0] = red[0][0] + red[0][1];
            part[2 * me + 1] = red[1][0] + red[1][1];
            for (int s_ = me + units_per_pair; s_ < xa.slots_per_pair; s_ += units_per_pair) {
                part[2 * s_ + 0] = 0.0;
                part[2 * s_ + 1] = 0.0;
            }
        }
        if (a.tail.counter != nullptr && warp == 0) {
            // fused tail of the iteration (peer_device.cuh): the pair's last unit reduces the units' partials in a
            // fixed order and applies the reference's convergence test -- no iter_finalize launch
            unsigned ticket = 0;
            if (lane == 0) {
                __threadfence();
                ticket = atomicAdd(a.tail.counter + pair, 1u);
            }
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            const int units_per_pair = xa.n_bands * xa.n_pairs_of_strips;
            if (ticket == (unsigned)units_per_pair - 1u)
                warp_iteration_tail(a.tail, a.partial + (size_t)pair * xa.slots_per_pair * 2, units_per_pair, pair, lane);
        }
    }
}

}  // namespace

// window 5 on frames (single scale, or prev + the warped plane) for any frame size
cudaError_t launch_lk_exact_march(int src, const TileArgs& a, int batch, cudaStream_t stream) {
    if (batch < 1 || batch > 65535 || (size_t)a.H * a.W >= ((size_t)1 << 31)) return cudaErrorInvalidValue;
    const int row_lo = (src == SRC_WARPED) ? a.row_lo : 0, row_hi = (src == SRC_WARPED) ? a.row_hi : a.H;
    const int rows = row_hi - row_lo;
    if (rows <= 0) return cudaErrorInvalidValue;
    XmArgs x;
    x.t = a;
    const int n_strips = (a.W + XM_OUT - 1) / XM_OUT;
    x.n_pairs_of_strips = (n_strips + 1) / 2;
    x.slots_per_pair = lk_tile_blocks_per_pair(rows, a.W);
    // Bands: units (one per CTA of two warps) run in waves of 148 SMs x 6 resident units and every band spends XM_EXTRA rows on
    // warm-up; pick the band count that minimises  waves x (rows per band + XM_EXTRA).  Bands of at least 16 rows keep
    // the units of a pair within the partial-sum slots (16 x 64 tiles) the iteration's tail reads.
    const long long slots = 148LL * XM_UNITS_PER_SM;
    const long long per_band = (long long)batch * x.n_pairs_of_strips;
    const int max_bands = rows >= 32 ? rows / 16 : 1;
    long long best_cost = -1;
    int best_rows = rows;
    for (int nb = 1; nb <= max_bands && nb <= 1024; ++nb) {
        int br = (rows + nb - 1) / nb;
        if (br < 16) br = 16;
        const int bands = (rows + br - 1) / br;
        const long long waves = (per_band * bands + slots - 1) / slots;
        const long long cost = waves * (br + XM_EXTRA);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best_rows = br;
        }
    }
    x.band_rows = best_rows;
    x.n_bands = (rows + best_rows - 1) / best_rows;
    x.n_units = (long long)batch * x.n_bands * x.n_pairs_of_strips;
    if ((long long)x.n_bands * x.n_pairs_of_strips > x.slots_per_pair) return cudaErrorInvalidValue;  // cannot happen (see above)
    if (x.n_units > 0x7fffffffLL) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)x.n_units;  // one unit per CTA
    static SmemOptIn opt_in[2];
    switch (src) {
        case SRC_FRAMES: {
            const cudaError_t e = opt_in[0].ensure(lk_exact_march_kernel<SRC_FRAMES>, XM_SMEM_BYTES);
            if (e != cudaSuccess) return e;
            OF_LAUNCH(lk_exact_march_kernel<SRC_FRAMES>, grid, XM_WARPS * 32, XM_SMEM_BYTES, stream, x);
            break;
        }
        case SRC_WARPED: {
            const cudaError_t e = opt_in[1].ensure(lk_exact_march_kernel<SRC_WARPED>, XM_SMEM_BYTES);
            if (e != cudaSuccess) return e;
            OF_LAUNCH(lk_exact_march_kernel<SRC_WARPED>, grid, XM_WARPS * 32, XM_SMEM_BYTES, stream, x);
            break;
        }
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace ofb
