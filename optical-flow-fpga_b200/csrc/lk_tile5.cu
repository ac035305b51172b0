// K1 / K3 (exact), window 5: second version of the reference-order tile kernel (lk_tile.cu holds the
// first, which stays the path for the other windows and for SRC_GRADS / SRC_WARP).
//
// Same arithmetic, bit for bit (python/lucas_kanade_core.py:15-45, 73-135):
//   avg = (p + c) / 2, It = p - c;  Sobel = true convolution with the float32 accumulator fed tap by tap
//   in kernel order, zero taps included;  the 25 products of a window summed like np.sum (8 running
//   lanes, tree, tail, + 0.0);  Cramer without FMA.
// What changed is where the work is done.  The first kernel was issue-bound at 406 lane-instructions per
// pixel (profiles/r01c_tile_frames_ncu_full_summary.json): every thread re-formed the five products of
// every gradient pixel of its 5 x 6 neighbourhood (75 FMUL per output pixel) and multiplied every Sobel
// tap by its coefficient (24 FMUL per pixel).  Here
//   * stage A stores avg * 0.125 and avg * 0.25 (E, D): the Sobel coefficients are +-0.125, +-0.25 and 0,
//     and x * (-k) == -(x * k), a + (-b) == a - b hold exactly in IEEE arithmetic, so a tap is one FADD /
//     FSUB of E or D; a zero tap adds D * 0.0 (same sign / NaN as avg * 0.0), computed once per value;
//   * stage B forms the five products once per gradient pixel and stores five product planes;
//   * stage C is additions only: one thread = four horizontally adjacent outputs, every product row is two
//     128-bit shared-memory loads feeding the four windows' accumulators in np.sum's tap order.
// About 270 lane-instructions per pixel.
//
//   SRC_FRAMES  lucas_kanade_single_scale  (in0 = prev, in1 = curr)
//   SRC_WARPED  one refinement iteration of lucas_kanade_pyramidal on (prev, warped curr): flow_out = flow_in + d
//               and per-block sums of |du|, |dv| (python/lucas_kanade_pyramidal.py:203-214)
#include <cuda_runtime.h>

#include "of_common.cuh"
#include "of_kernels.h"

namespace ofb {
namespace {

constexpr int T5_TX = 64, T5_TY = 16, T5_THREADS = 256;
constexpr int T5_GW = T5_TX + 4, T5_GH = T5_TY + 4;  // gradient / product tile: window halo of 2
constexpr int T5_FW = T5_GW + 2, T5_FH = T5_GH + 2;  // frame tile: + Sobel halo of 1
constexpr int T5_G = T5_GH * T5_GW;                  // 1360 floats per product plane (16-byte multiple)
constexpr int T5_F = T5_FH * T5_FW;                  // 1540 floats per frame plane (16-byte multiple)

// np.sum's order as a streaming accumulator, window 5: taps arrive in index order t = 0 .. 24 (t is a
// compile-time constant after unrolling): lanes t & 7 for t < 24, the fixed tree, the tail, + 0.0.
struct Np25 {
    float lane[8];
    float res;
};
__device__ __forceinline__ void np25_add(Np25& s, int t, float p) {
    if (t < 8) {
        s.lane[t] = p;
    } else if (t < 24) {
        s.lane[t & 7] = fadd(s.lane[t & 7], p);
    } else {
        s.res = fadd(fadd(fadd(s.lane[0], s.lane[1]), fadd(s.lane[2], s.lane[3])),
                     fadd(fadd(s.lane[4], s.lane[5]), fadd(s.lane[6], s.lane[7])));
        s.res = fadd(s.res, p);
    }
}
__device__ __forceinline__ float np25_finish(const Np25& s) { return fadd(0.0f, s.res); }

template <int SRC>
__global__ void __launch_bounds__(T5_THREADS, 3) lk_tile5_kernel(TileArgs a) {
    constexpr bool FLOW = (SRC == SRC_WARPED);
    __shared__ __align__(16) float smem[5 * T5_G + 3 * T5_F];
    float* prod = smem;             // xx, yy, xy, xt, yt planes [T5_GH][T5_GW]
    float* sE = smem + 5 * T5_G;    // avg * 0.125   [T5_FH][T5_FW]
    float* sD = sE + T5_F;          // avg * 0.25
    float* sT = sD + T5_F;          // It = p - c

    const int pair = blockIdx.z;
    if (FLOW && a.done != nullptr && a.done[pair]) return;  // level already converged

    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    const int ox = blockIdx.x * T5_TX, oy = (FLOW ? a.row_lo : 0) + blockIdx.y * T5_TY;
    const int y_end = FLOW ? a.row_hi : H;
    const int tid = threadIdx.x;

    // ---- stage A: frames with the replicated (= 'symm' for a 3 x 3 kernel) border ----------------
    // element i = tid + 256 n of the 22 x 70 frame tile: (row, column) advance by (3, 46) per step (256 =
    // 3 * 70 + 46), no division in the loop; fully unrolled so that all of a thread's loads are in flight
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
            fr += 3;
            fx += T5_THREADS - 3 * T5_FW;
            if (fx >= T5_FW) {
                fx -= T5_FW;
                fr += 1;
            }
        }
#pragma unroll
        for (int n = 0; n < STEPS; ++n) {
            const int i = tid + n * T5_THREADS;
            if ((n + 1) * T5_THREADS <= T5_F || i < T5_F) {
                const float avg = fmul(fadd(pv[n], cv[n]), 0.5f);  // (p + c) / 2.0, exact either way
                sE[i] = fmul(avg, 0.125f);
                sD[i] = fmul(avg, 0.25f);
                sT[i] = fsub(pv[n], cv[n]);
            }
        }
    }
    __syncthreads();

    // ---- stage B: Sobel (kernel order (j, k), tap (j, k) reads frame offset (2 - j, 2 - k)) and the five
    // products, two horizontally adjacent gradient pixels per item ------------------------------------
    for (int item = tid; item < T5_GH * (T5_GW / 2); item += T5_THREADS) {
        const int r = item / (T5_GW / 2), c = 2 * (item - r * (T5_GW / 2));
        // frame-tile rows r .. r + 2, columns c .. c + 3 (c even, even pitch: 8-byte aligned pairs)
        float E0[4], E2[4], D0[4], D1[4], D2[4];
        {
            const float2* e0 = reinterpret_cast<const float2*>(sE + r * T5_FW + c);
            const float2* e2 = reinterpret_cast<const float2*>(sE + (r + 2) * T5_FW + c);
            const float2* d0 = reinterpret_cast<const float2*>(sD + r * T5_FW + c);
            const float2* d1 = reinterpret_cast<const float2*>(sD + (r + 1) * T5_FW + c);
            const float2* d2 = reinterpret_cast<const float2*>(sD + (r + 2) * T5_FW + c);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 a0 = e0[h], a2 = e2[h], b0 = d0[h], b1 = d1[h], b2 = d2[h];
                E0[2 * h] = a0.x; E0[2 * h + 1] = a0.y;
                E2[2 * h] = a2.x; E2[2 * h + 1] = a2.y;
                D0[2 * h] = b0.x; D0[2 * h + 1] = b0.y;
                D1[2 * h] = b1.x; D1[2 * h + 1] = b1.y;
                D2[2 * h] = b2.x; D2[2 * h + 1] = b2.y;
            }
        }
        // zero taps: value * 0.0f (keeps the reference's signed zeros / NaN propagation)
        const float Z0[2] = {fmul(D0[1], 0.0f), fmul(D0[2], 0.0f)};  // row 0, columns 1, 2
        const float Z2[2] = {fmul(D2[1], 0.0f), fmul(D2[2], 0.0f)};  // row 2, columns 1, 2
        const float Z1[4] = {fmul(D1[0], 0.0f), fmul(D1[1], 0.0f), fmul(D1[2], 0.0f), fmul(D1[3], 0.0f)};
        float gxv[2], gyv[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            float ax = 0.0f, ay = 0.0f;
            // j = 0 (frame row 2): kx = -.125, 0, .125   ky = -.125, -.25, -.125
            ax = fsub(ax, E2[s + 2]);  ay = fsub(ay, E2[s + 2]);
            ax = fadd(ax, Z2[s]);      ay = fsub(ay, D2[s + 1]);
            ax = fadd(ax, E2[s]);      ay = fsub(ay, E2[s]);
            // j = 1 (frame row 1): kx = -.25, 0, .25      ky = 0, 0, 0
            ax = fsub(ax, D1[s + 2]);  ay = fadd(ay, Z1[s + 2]);
            ax = fadd(ax, Z1[s + 1]);  ay = fadd(ay, Z1[s + 1]);
            ax = fadd(ax, D1[s]);      ay = fadd(ay, Z1[s]);
            // j = 2 (frame row 0): kx = -.125, 0, .125   ky = .125, .25, .125
            ax = fsub(ax, E0[s + 2]);  ay = fadd(ay, E0[s + 2]);
            ax = fadd(ax, Z0[s]);      ay = fadd(ay, D0[s + 1]);
            ax = fadd(ax, E0[s]);      ay = fadd(ay, E0[s]);
            gxv[s] = ax;
            gyv[s] = ay;
        }
        const float t0 = sT[(r + 1) * T5_FW + c + 1], t1 = sT[(r + 1) * T5_FW + c + 2];
        const int g = r * T5_GW + c;
        *reinterpret_cast<float2*>(prod + 0 * T5_G + g) = make_float2(fmul(gxv[0], gxv[0]), fmul(gxv[1], gxv[1]));
        *reinterpret_cast<float2*>(prod + 1 * T5_G + g) = make_float2(fmul(gyv[0], gyv[0]), fmul(gyv[1], gyv[1]));
        *reinterpret_cast<float2*>(prod + 2 * T5_G + g) = make_float2(fmul(gxv[0], gyv[0]), fmul(gxv[1], gyv[1]));
        *reinterpret_cast<float2*>(prod + 3 * T5_G + g) = make_float2(fmul(gxv[0], t0), fmul(gxv[1], t1));
        *reinterpret_cast<float2*>(prod + 4 * T5_G + g) = make_float2(fmul(gyv[0], t0), fmul(gyv[1], t1));
    }
    __syncthreads();

    // ---- stage C: window sums in NumPy's order + Cramer; one thread = outputs (r, c .. c + 3) -----------
    const int r = tid >> 4, c = 4 * (tid & 15);
    const int y = oy + r, x0 = ox + c;
    double acc_u = 0.0, acc_v = 0.0;
    if (y < y_end && x0 < W) {
        float sum[5][4];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const float* P = prod + q * T5_G + r * T5_GW + c;
            Np25 s[4];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const float4 lo = *reinterpret_cast<const float4*>(P + i * T5_GW);
                const float4 hi = *reinterpret_cast<const float4*>(P + i * T5_GW + 4);
                const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                for (int w = 0; w < 4; ++w)
#pragma unroll
                    for (int k = 0; k < 5; ++k) np25_add(s[w], 5 * i + k, v[w + k]);
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) sum[q][w] = np25_finish(s[w]);
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
        for (int w = 0; w < 4; ++w) {
            const int x = x0 + w;
            if (x < W) {
                float u = 0.0f, v = 0.0f;
                if (row_inside && x >= 2 && x < W - 2) cramer_solve(sum[0][w], sum[1][w], sum[2][w], sum[3][w], sum[4][w], u, v);
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
    switch (src) {
        case SRC_FRAMES: OF_LAUNCH(lk_tile5_kernel<SRC_FRAMES>, grid, T5_THREADS, 0, stream, a); break;
        case SRC_WARPED: OF_LAUNCH(lk_tile5_kernel<SRC_WARPED>, grid, T5_THREADS, 0, stream, a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace ofb
