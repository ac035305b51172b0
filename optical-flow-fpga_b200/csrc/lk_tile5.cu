// K1 / K3 (exact), window 5: the reference-order tile kernel, third version (lk_tile.cu holds the first, which
// stays the path for the other windows and for SRC_GRADS / SRC_WARP).
//
// Same arithmetic, bit for bit (python/lucas_kanade_core.py:15-45, 73-135):
//   avg = (p + c) / 2, It = p - c;  Sobel = true convolution with the float32 accumulator fed tap by tap
//   in kernel order, zero taps included;  the 25 products of a window summed like np.sum (8 running
//   lanes, tree, tail, + 0.0);  Cramer without FMA.
// The first kernel was issue-bound at 406 lane-instructions per pixel: every thread re-formed the products of
// its neighbourhood and multiplied every Sobel tap by its coefficient.  The second version did each piece of
// work once (scaled frame planes E = avg * 0.125, D = avg * 0.25 so that a Sobel tap is one FADD / FSUB -- the
// coefficients are +-0.125, +-0.25, 0 and x * (-k) == -(x * k), a + (-b) == a - b hold exactly; product planes
// in shared memory; additions-only window sums): 265 lane-instructions per pixel, still issue-bound, and 125
// of them the additions np.sum's order forces.
// This version halves the floating-point instruction count of the Sobel, product and window-sum stages with
// Blackwell's packed float32 pairs (FADD2 / FMUL2: one issue slot, two independent IEEE operations): every
// shared-memory plane is stored INTERLEAVED BY HALVES -- element x of a tile row shares a 64-bit word with
// element x + 32 -- so that any tap of any stencil is one aligned 64-bit operand holding the same tap of two
// outputs 32 columns apart, and one thread computes eight outputs (four adjacent columns and their twins 32
// columns to the right) with the instruction stream of four.  The operations and their order per output are
// unchanged, so are the bits.
//
//   SRC_FRAMES  lucas_kanade_single_scale  (in0 = prev, in1 = curr)
//   SRC_WARPED  one refinement iteration of lucas_kanade_pyramidal on (prev, warped curr): flow_out = flow_in + d
//               and per-block sums of |du|, |dv| (python/lucas_kanade_pyramidal.py:203-214)
#include <cuda_runtime.h>

#include "of_common.cuh"
#include "f32x2.cuh"
#include "of_kernels.h"

namespace ofb {
namespace {

constexpr int T5_TX = 64, T5_TY = 16, T5_THREADS = 128;
constexpr int T5_HALF = T5_TX / 2;                   // column x is packed with column x + 32
constexpr int T5_GW = T5_TX + 4, T5_GH = T5_TY + 4;  // gradient / product tile: window halo of 2 (68 x 20)
constexpr int T5_FW = T5_GW + 2, T5_FH = T5_GH + 2;  // frame tile: + Sobel halo of 1 (70 x 22)
constexpr int T5_GP = T5_GW - T5_HALF;               // 36 packed product columns (columns 32..35 sit in both halves)
constexpr int T5_FP = T5_FW - T5_HALF;               // 38 packed frame columns
// packed words per row of every plane: 19 x 16 bytes -- an odd number of 16-byte units, so the eight threads of a
// 128-bit shared-memory access phase, which read eight different rows, hit eight different bank groups
constexpr int T5_PITCH = 38;
constexpr int T5_F2 = T5_FH * T5_PITCH;              // packed words per frame plane
constexpr int T5_G2 = T5_GH * T5_PITCH;              // packed words per product plane
constexpr int T5_F = T5_FH * T5_FW;                  // frame-tile elements (stage A's loop)
constexpr size_t T5_SMEM_BYTES = (size_t)(3 * T5_F2 + 5 * T5_G2) * sizeof(f32x2);  // 50 464
static_assert(T5_THREADS == (T5_TX / 8) * T5_TY, "one thread = 2 x 4 outputs");
static_assert((T5_PITCH * 8) % 16 == 0 && ((T5_PITCH * 8 / 16) & 1) == 1, "rows 16-byte aligned, odd pitch in 16-byte units");

// np.sum's order as a streaming accumulator, window 5, on packed pairs: taps arrive in index order t = 0 .. 24
// (t is a compile-time constant after unrolling): lanes t & 7 for t < 24, the fixed tree, the tail, + 0.0.
struct Np25 {
    f32x2 lane[8];
    f32x2 res;
};
__device__ __forceinline__ void np25_add(Np25& s, int t, f32x2 p) {
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
__global__ void __launch_bounds__(T5_THREADS, 4) lk_tile5_kernel(TileArgs a) {
    constexpr bool FLOW = (SRC == SRC_WARPED);
    OF_DYNAMIC_SMEM_ALIGNED(16, unsigned char, t5_smem);
    f32x2* sE = reinterpret_cast<f32x2*>(t5_smem);  // avg * 0.125   [T5_FH][T5_PITCH], word pc = (column pc, column pc + 32)
    f32x2* sD = sE + T5_F2;                         // avg * 0.25
    f32x2* sT = sD + T5_F2;                         // It = p - c
    f32x2* prod = sT + T5_F2;                       // xx, yy, xy, xt, yt planes [T5_GH][T5_PITCH]

    const int pair = blockIdx.z;
    if (FLOW && a.done != nullptr && a.done[pair]) return;  // level already converged

    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    const int ox = blockIdx.x * T5_TX, oy = (FLOW ? a.row_lo : 0) + blockIdx.y * T5_TY;
    const int y_end = FLOW ? a.row_hi : H;
    const int tid = threadIdx.x;

    // ---- stage A: frames with the replicated (= 'symm' for a 3 x 3 kernel) border ----------------
    // element i = tid + 128 n of the 22 x 70 frame tile: (row, column) advance by (1, 58) per step (128 =
    // 70 + 58), no division in the loop; fully unrolled so that all of a thread's loads are in flight.
    // Column fx goes to half 0 of word fx (fx < 38) and to half 1 of word fx - 32 (fx >= 32).
    {
        // one widening multiply-add per address (IMAD.WIDE.U32) from the pair's base
        const char* prev = reinterpret_cast<const char*>(a.in0 + pair * plane);
        const char* curr = reinterpret_cast<const char*>(a.in1 + pair * plane);
        int fr = tid / T5_FW, fx = tid - fr * T5_FW;
        constexpr int STEPS = (T5_F + T5_THREADS - 1) / T5_THREADS;
        float pv[STEPS], cv[STEPS];
#pragma unroll
        for (int n = 0; n < STEPS; ++n) {
            if ((n + 1) * T5_THREADS <= T5_F || tid + n * T5_THREADS < T5_F) {
                const int y = clampi(oy - 3 + fr, 0, H - 1);
                const int x = clampi(ox - 3 + fx, 0, W - 1);
                const unsigned o = (unsigned)y * (unsigned)W + (unsigned)x;  // H * W < 2^31 (launcher)
                pv[n] = __ldg(reinterpret_cast<const float*>(prev + (size_t)o * 4u));
                cv[n] = __ldg(reinterpret_cast<const float*>(curr + (size_t)o * 4u));
            } else {
                pv[n] = 0.0f;
                cv[n] = 0.0f;
            }
            fr += 1;
            fx += T5_THREADS - T5_FW;
            if (fx >= T5_FW) {
                fx -= T5_FW;
                fr += 1;
            }
        }
        fr = tid / T5_FW;
        fx = tid - fr * T5_FW;
        float* fE = reinterpret_cast<float*>(sE);
        float* fD = reinterpret_cast<float*>(sD);
        float* fT = reinterpret_cast<float*>(sT);
#pragma unroll
        for (int n = 0; n < STEPS; ++n) {
            if ((n + 1) * T5_THREADS <= T5_F || tid + n * T5_THREADS < T5_F) {
                const float avg = fmul(fadd(pv[n], cv[n]), 0.5f);  // (p + c) / 2.0, exact either way
                const float e = fmul(avg, 0.125f), d = fmul(avg, 0.25f), t = fsub(pv[n], cv[n]);
                const int w0 = 2 * (fr * T5_PITCH + fx);  // float index of half 0 of word fx
                if (fx < T5_FP) {
                    fE[w0] = e;
                    fD[w0] = d;
                    fT[w0] = t;
                }
                if (fx >= T5_HALF) {
                    fE[w0 - 2 * T5_HALF + 1] = e;
                    fD[w0 - 2 * T5_HALF + 1] = d;
                    fT[w0 - 2 * T5_HALF + 1] = t;
                }
            }
            fr += 1;
            fx += T5_THREADS - T5_FW;
            if (fx >= T5_FW) {
                fx -= T5_FW;
                fr += 1;
            }
        }
    }
    __syncthreads();

    // ---- stage B: Sobel (kernel order (j, k), tap (j, k) reads frame offset (2 - j, 2 - k)) and the five
    // products; one item = gradient pixel (r, pc) and its twin (r, pc + 32), every operand one packed word ----
    {
        const f32x2 zero2 = pk(0.0f, 0.0f);
        int r = tid / T5_GP, pc = tid - r * T5_GP;
        for (int item = tid; item < T5_GH * T5_GP; item += T5_THREADS) {
            const f32x2* e0 = sE + r * T5_PITCH + pc;        // frame-tile rows r .. r + 2, words pc .. pc + 2
            const f32x2* e2 = e0 + 2 * T5_PITCH;
            const f32x2* d0 = sD + r * T5_PITCH + pc;
            const f32x2* d1 = d0 + T5_PITCH;
            const f32x2* d2 = d1 + T5_PITCH;
            const f32x2 E0lo = e0[0], E0hi = e0[2], E2lo = e2[0], E2hi = e2[2];
            const f32x2 D0m = d0[1], D2m = d2[1], D1lo = d1[0], D1m = d1[1], D1hi = d1[2];
            // zero taps: value * 0.0f (keeps the reference's signed zeros / NaN propagation)
            const f32x2 Z0 = mul2(D0m, zero2), Z2 = mul2(D2m, zero2);
            const f32x2 Z1lo = mul2(D1lo, zero2), Z1m = mul2(D1m, zero2), Z1hi = mul2(D1hi, zero2);
            f32x2 ax = zero2, ay = zero2;
            // j = 0 (frame row 2): kx = -.125, 0, .125   ky = -.125, -.25, -.125
            ax = sub2(ax, E2hi);  ay = sub2(ay, E2hi);
            ax = add2(ax, Z2);    ay = sub2(ay, D2m);
            ax = add2(ax, E2lo);  ay = sub2(ay, E2lo);
            // j = 1 (frame row 1): kx = -.25, 0, .25      ky = 0, 0, 0
            ax = sub2(ax, D1hi);  ay = add2(ay, Z1hi);
            ax = add2(ax, Z1m);   ay = add2(ay, Z1m);
            ax = add2(ax, D1lo);  ay = add2(ay, Z1lo);
            // j = 2 (frame row 0): kx = -.125, 0, .125   ky = .125, .25, .125
            ax = sub2(ax, E0hi);  ay = add2(ay, E0hi);
            ax = add2(ax, Z0);    ay = add2(ay, D0m);
            ax = add2(ax, E0lo);  ay = add2(ay, E0lo);
            const f32x2 gt = sT[(r + 1) * T5_PITCH + pc + 1];
            const int g = r * T5_PITCH + pc;
            prod[0 * T5_G2 + g] = mul2(ax, ax);
            prod[1 * T5_G2 + g] = mul2(ay, ay);
            prod[2 * T5_G2 + g] = mul2(ax, ay);
            prod[3 * T5_G2 + g] = mul2(ax, gt);
            prod[4 * T5_G2 + g] = mul2(ay, gt);
            // next item: 128 = 3 * 36 + 20
            r += 3;
            pc += T5_THREADS - 3 * T5_GP;
            if (pc >= T5_GP) {
                pc -= T5_GP;
                r += 1;
            }
        }
    }
    __syncthreads();

    // ---- stage C: window sums in NumPy's order + Cramer; one thread = outputs (r, c .. c + 3) and their twins
    // (r, c + 32 .. c + 35): four packed windows.  Consecutive threads take consecutive rows (see T5_PITCH). ----
    const int r = tid & (T5_TY - 1), c = 4 * (tid >> 4);
    const int y = oy + r, x0 = ox + c;
    double acc_u = 0.0, acc_v = 0.0;
    if (y < y_end && x0 < W) {
        const f32x2 zero2 = pk(0.0f, 0.0f);
        f32x2 sum[5][4];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const f32x2* P = prod + q * T5_G2 + r * T5_PITCH + c;
            Np25 s[4];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const ulonglong2* row = reinterpret_cast<const ulonglong2*>(P + i * T5_PITCH);
                const ulonglong2 q0 = row[0], q1 = row[1], q2 = row[2], q3 = row[3];
                const f32x2 v[8] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y};
#pragma unroll
                for (int w = 0; w < 4; ++w)
#pragma unroll
                    for (int k = 0; k < 5; ++k) np25_add(s[w], 5 * i + k, v[w + k]);
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) sum[q][w] = add2(zero2, s[w].res);  // np.add.reduce starts from +0.0
        }
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
        const bool row_inside = (y >= 2 && y < H - 2);
        const bool row_owned = FLOW && (y >= a.own_lo && y < a.own_hi);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int x = x0 + T5_HALF * half + w;
                if (x < W) {
                    float sq[5];
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        float lo, hi;
                        unpk(sum[q][w], lo, hi);
                        sq[q] = half ? hi : lo;
                    }
                    float u, v;  // branch-free: the division runs on a safe denominator, the border / singular case selects 0
                    cramer_solve_select(sq[0], sq[1], sq[2], sq[3], sq[4], row_inside && x >= 2 && x < W - 2, u, v);
                    const size_t go = (size_t)y * W + x;
                    if (FLOW) {
                        out_u[go] = fadd(__ldg(fin_u + go), u);  // flow += d
                        out_v[go] = fadd(__ldg(fin_v + go), v);
                        if (row_owned) {
                            acc_u += (double)fabsf(u);
                            acc_v += (double)fabsf(v);
                        }
                    } else {
                        out_u[go] = u;
                        out_v[go] = v;
                    }
                }
            }
        }
    }

    if (FLOW && a.partial != nullptr) {
        // deterministic block reduction (fixed shuffle tree, then warps in order)
        __shared__ double red[2][T5_THREADS / 32];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            acc_u += __shfl_down_sync(0xffffffffu, acc_u, off);
            acc_v += __shfl_down_sync(0xffffffffu, acc_v, off);
        }
        if ((tid & 31) == 0) {
            red[0][tid >> 5] = acc_u;
            red[1][tid >> 5] = acc_v;
        }
        __syncthreads();
        if (tid == 0) {
            double su = 0.0, sv = 0.0;
            for (int w = 0; w < T5_THREADS / 32; ++w) {
                su += red[0][w];
                sv += red[1][w];
            }
            const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
            const size_t nblk = (size_t)gridDim.x * gridDim.y;
            a.partial[(pair * nblk + blk) * 2 + 0] = su;
            a.partial[(pair * nblk + blk) * 2 + 1] = sv;
        }
    }
}

}  // namespace

// same tile geometry as lk_tile.cu (16 x 64), so lk_tile_blocks_per_pair sizes the partial sums for both
cudaError_t launch_lk_tile5(int src, const TileArgs& a, int batch, cudaStream_t stream) {
    if (batch < 1 || batch > 65535 || (size_t)a.H * a.W >= ((size_t)1 << 31)) return cudaErrorInvalidValue;
    const int rows = (src == SRC_WARPED) ? a.row_hi - a.row_lo : a.H;
    if (rows <= 0) return cudaErrorInvalidValue;
    dim3 grid((a.W + T5_TX - 1) / T5_TX, (rows + T5_TY - 1) / T5_TY, batch);
    static SmemOptIn opt_in[2];
    switch (src) {
        case SRC_FRAMES: {
            const cudaError_t e = opt_in[0].ensure(lk_tile5_kernel<SRC_FRAMES>, T5_SMEM_BYTES);
            if (e != cudaSuccess) return e;
            OF_LAUNCH(lk_tile5_kernel<SRC_FRAMES>, grid, T5_THREADS, T5_SMEM_BYTES, stream, a);
            break;
        }
        case SRC_WARPED: {
            const cudaError_t e = opt_in[1].ensure(lk_tile5_kernel<SRC_WARPED>, T5_SMEM_BYTES);
            if (e != cudaSuccess) return e;
            OF_LAUNCH(lk_tile5_kernel<SRC_WARPED>, grid, T5_THREADS, T5_SMEM_BYTES, stream, a);
            break;
        }
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace ofb
