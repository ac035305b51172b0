// K1 / K3 (exact): tile kernel that reproduces the reference's operation order bit for bit
// on ANY float32 input and any odd window up to 11.
//
//   SRC_FRAMES  lucas_kanade_single_scale      (python/lucas_kanade_core.py:48-70)
//   SRC_GRADS   lucas_kanade_from_gradients    (python/lucas_kanade_core.py:73-135)
//   SRC_WARP    one refinement iteration of lucas_kanade_pyramidal
//               (python/lucas_kanade_pyramidal.py:203-214): warp_image (float64 bilinear,
//               outside -> 0) -> single-scale LK against the previous frame -> flow += d,
//               plus per-block partial sums of |du|, |dv| for the convergence test.
//   SRC_WARPED  the same iteration with the warped frame already in HBM (written by
//               warp_rows_kernel<double>, lk_march.cu): stage A is SRC_FRAMES' on (prev, warped).
//
// Order of operations that is mirrored (pinned by tests/golden):
//   * avg = (p + c) / 2;  Sobel = true 2-D convolution, symmetric border, float32
//     accumulator fed tap by tap in kernel order (j, k), zero taps included;
//   * the w*w products of a window are summed like np.sum on a contiguous vector:
//     8 running lanes, fixed tree, tail, and the result added to the +0.0 identity;
//   * Cramer solve without FMA.
//
// One CTA = 16 x 64 output pixels.  Stage A stages the two frames (or prev + warped curr)
// with halo in shared memory, stage B turns them into Ix/Iy/It tiles in shared memory,
// stage C does the window sums + solve; Ix/Iy/It never go to HBM.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

#include "of_common.cuh"
#include "of_kernels.h"

namespace ofb {

constexpr int TX = 64;
constexpr int TY = 16;
constexpr int TILE_THREADS = 256;

template <int WIN>
__device__ __forceinline__ float numpy_order_sum(const float* __restrict__ a, const float* __restrict__ b,
                                                 int pitch) {
    // sum_t a[t] * b[t] over the WIN x WIN window at (a, b), t = WIN * i + j
    constexpr int N = WIN * WIN;
    if (N < 8) {
        float res = -0.0f;
#pragma unroll
        for (int i = 0; i < WIN; ++i)
#pragma unroll
            for (int j = 0; j < WIN; ++j) res = fadd(res, fmul(a[i * pitch + j], b[i * pitch + j]));
        return fadd(0.0f, res);
    }
    constexpr int FULL = N - (N % 8);
    float lane[8];
    float res = 0.0f;
#pragma unroll
    for (int t = 0; t < N; ++t) {
        const int i = t / WIN, j = t % WIN;
        const float p = fmul(a[i * pitch + j], b[i * pitch + j]);
        if (t < 8) {
            lane[t] = p;
        } else if (t < FULL) {
            lane[t & 7] = fadd(lane[t & 7], p);
        } else {
            if (t == FULL)
                res = fadd(fadd(fadd(lane[0], lane[1]), fadd(lane[2], lane[3])),
                           fadd(fadd(lane[4], lane[5]), fadd(lane[6], lane[7])));
            res = fadd(res, p);
        }
    }
    if (FULL == N)
        res = fadd(fadd(fadd(lane[0], lane[1]), fadd(lane[2], lane[3])),
                   fadd(fadd(lane[4], lane[5]), fadd(lane[6], lane[7])));
    return fadd(0.0f, res);
}

// np.sum's order as a streaming accumulator: taps arrive in index order t = 0 .. N-1 (N = WIN * WIN >= 8),
// t is a compile-time constant after unrolling.  Lets one thread feed the same per-pixel product
// into the windows of two adjacent output pixels.
template <int WIN>
struct NpAcc {
    float lane[8];
    float res;
};
__device__ __forceinline__ float np_tree(const float* lane) {
    return fadd(fadd(fadd(lane[0], lane[1]), fadd(lane[2], lane[3])), fadd(fadd(lane[4], lane[5]), fadd(lane[6], lane[7])));
}
template <int WIN>
__device__ __forceinline__ void np_add(NpAcc<WIN>& s, int t, float p) {
    constexpr int N = WIN * WIN, FULL = N - (N % 8);
    if (t < 8) {
        s.lane[t] = p;
    } else if (t < FULL) {
        s.lane[t & 7] = fadd(s.lane[t & 7], p);
    } else {
        if (t == FULL) s.res = np_tree(s.lane);
        s.res = fadd(s.res, p);
    }
}
template <int WIN>
__device__ __forceinline__ float np_finish(const NpAcc<WIN>& s) {
    constexpr int N = WIN * WIN, FULL = N - (N % 8);
    return fadd(0.0f, FULL == N ? np_tree(s.lane) : s.res);
}

template <int SRC, int WIN>
__global__ void __launch_bounds__(TILE_THREADS, WIN <= 5 ? (SRC == SRC_WARPED ? 3 : 2) : 1) lk_tile_kernel(TileArgs a) {
    OF_DYNAMIC_SMEM(float, smem);
    constexpr bool FLOW = (SRC == SRC_WARP || SRC == SRC_WARPED);  // refinement iteration: flow_out = flow_in + d
    constexpr bool GATHER = (SRC == SRC_WARP);                     // stage A gathers the current frame itself
    constexpr int HW = WIN / 2;
    constexpr int GW = TX + 2 * HW, GH = TY + 2 * HW;  // gradient tile
    constexpr int FW = GW + 2, FH = GH + 2;            // frame tile (Sobel halo)
    float* gx = smem;
    float* gy = gx + GH * GW;
    float* gt = gy + GH * GW;
    float* fp = gt + GH * GW;  // prev, then the frame average
    float* fc = fp + FH * FW;  // curr (or warped curr), then It

    const int pair = blockIdx.z;
    if (FLOW && a.done != nullptr && a.done[pair]) return;  // level already converged

    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    const int ox = blockIdx.x * TX, oy = (FLOW ? a.row_lo : 0) + blockIdx.y * TY;
    const int y_end = FLOW ? a.row_hi : H;
    const int tid = threadIdx.x;

    const float* fin_u = nullptr;
    const float* fin_v = nullptr;
    float* fout_u = nullptr;
    float* fout_v = nullptr;
    if (FLOW) {
        const int cur = (a.sel ? a.sel[pair] : 0) ^ a.sel_xor;
        if (SRC == SRC_WARPED) {
            // selects instead of a run-time index into the parameter block (which costs a local copy)
            fin_u = (cur ? a.flow_u[1] : a.flow_u[0]) + pair * plane;
            fin_v = (cur ? a.flow_v[1] : a.flow_v[0]) + pair * plane;
            fout_u = (cur ? a.flow_u[0] : a.flow_u[1]) + pair * plane;
            fout_v = (cur ? a.flow_v[0] : a.flow_v[1]) + pair * plane;
        } else {
            fin_u = a.flow_u[cur] + pair * plane;
            fin_v = a.flow_v[cur] + pair * plane;
            fout_u = a.flow_u[cur ^ 1] + pair * plane;
            fout_v = a.flow_v[cur ^ 1] + pair * plane;
        }
    }

    if (SRC == SRC_GRADS) {
        const float* ix = a.in0 + pair * plane;
        const float* iy = a.in1 + pair * plane;
        const float* it = a.in2 + pair * plane;
        for (int i = tid; i < GH * GW; i += TILE_THREADS) {
            const int y = oy - HW + i / GW, x = ox - HW + i % GW;
            const bool in = (y >= 0 && y < H && x >= 0 && x < W);
            const size_t o = (size_t)(in ? y : 0) * W + (in ? x : 0);
            gx[i] = in ? __ldg(ix + o) : 0.0f;
            gy[i] = in ? __ldg(iy + o) : 0.0f;
            gt[i] = in ? __ldg(it + o) : 0.0f;
        }
    } else {
        const float* prev = a.in0 + pair * plane;
        const float* curr = a.in1 + pair * plane;
        // stage A: frames with the replicated (= 'symm' for a 3x3 kernel) border
        for (int i = tid; i < FH * FW; i += TILE_THREADS) {
            const int y = clampi(oy - HW - 1 + i / FW, 0, H - 1);
            const int x = clampi(ox - HW - 1 + i % FW, 0, W - 1);
            const size_t o = (size_t)y * W + x;
            const float p = __ldg(prev + o);
            float c;
            if (GATHER) {
                // warp_image: coordinates are int64 + float32 -> float64
                const double yw = dadd((double)y, (double)__ldg(fin_v + o));
                const double xw = dadd((double)x, (double)__ldg(fin_u + o));
                c = bilinear_f64(curr, H, W, yw, xw);
            } else {
                c = __ldg(curr + o);
            }
            fp[i] = fmul(fadd(p, c), 0.5f);  // (p + c) / 2.0, exact either way
            fc[i] = fsub(p, c);              // It
        }
        __syncthreads();
        // stage B: Sobel by true convolution, accumulator fed in kernel order (j, k)
        for (int i = tid; i < GH * GW; i += TILE_THREADS) {
            const int r = i / GW + 1, c = i % GW + 1;  // centre in the frame tile
            const float kx[3][3] = {{-0.125f, 0.0f, 0.125f}, {-0.25f, 0.0f, 0.25f}, {-0.125f, 0.0f, 0.125f}};
            const float ky[3][3] = {{-0.125f, -0.25f, -0.125f}, {0.0f, 0.0f, 0.0f}, {0.125f, 0.25f, 0.125f}};
            float ax = 0.0f, ay = 0.0f;
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float v = fp[(r + 1 - j) * FW + (c + 1 - k)];
                    ax = fadd(ax, fmul(v, kx[j][k]));
                    ay = fadd(ay, fmul(v, ky[j][k]));
                }
            gx[i] = ax;
            gy[i] = ay;
            gt[i] = fc[r * FW + c];
        }
    }
    __syncthreads();

    // stage C: window sums in NumPy's order + Cramer.  One thread = two horizontally adjacent output
    // pixels: every gradient pixel of the 5 x 6 neighbourhood is loaded once and its five products are
    // formed once, then fed into both windows' accumulators (same tap order per window as np.sum).
    double acc_u = 0.0, acc_v = 0.0;
    auto emit = [&](int y, int x, float sxx, float syy, float sxy, float sxt, float syt) {
        if (y >= y_end || x >= W) return;
        float u = 0.0f, v = 0.0f;
        if (y >= HW && y < H - HW && x >= HW && x < W - HW) cramer_solve(sxx, syy, sxy, sxt, syt, u, v);
        const size_t go = (size_t)y * W + x;
        if (FLOW) {
            fout_u[go] = fadd(__ldg(fin_u + go), u);  // flow += d
            fout_v[go] = fadd(__ldg(fin_v + go), v);
            if (y >= a.own_lo && y < a.own_hi) {
                acc_u += (double)fabsf(u);
                acc_v += (double)fabsf(v);
            }
        } else {
            a.out_u[pair * plane + go] = u;
            a.out_v[pair * plane + go] = v;
        }
    };
    if (WIN * WIN >= 8) {
        for (int item = tid; item < TY * (TX / 2); item += TILE_THREADS) {
            const int r = item / (TX / 2), c = 2 * (item % (TX / 2));
            const int y = oy + r, x = ox + c;
            if (y >= y_end || x >= W) continue;
            NpAcc<WIN> s0[5], s1[5];
#pragma unroll
            for (int i = 0; i < WIN; ++i) {
                const float* rx = gx + (r + i) * GW + c;
                const float* ry = gy + (r + i) * GW + c;
                const float* rt = gt + (r + i) * GW + c;
                // WIN + 1 values per plane as 64-bit loads (c is even, the pitch GW is even): consecutive
                // lanes read consecutive 8-byte words, no bank conflicts
                float ax[WIN + 1], ay[WIN + 1], at[WIN + 1];
#pragma unroll
                for (int k = 0; k <= WIN; k += 2) {
                    const float2 fx2 = *reinterpret_cast<const float2*>(rx + k);
                    const float2 fy2 = *reinterpret_cast<const float2*>(ry + k);
                    const float2 ft2 = *reinterpret_cast<const float2*>(rt + k);
                    ax[k] = fx2.x; ax[k + 1] = fx2.y;
                    ay[k] = fy2.x; ay[k + 1] = fy2.y;
                    at[k] = ft2.x; at[k + 1] = ft2.y;
                }
#pragma unroll
                for (int k = 0; k <= WIN; ++k) {
                    const float vx = ax[k], vy = ay[k], vt = at[k];
                    const float p[5] = {fmul(vx, vx), fmul(vy, vy), fmul(vx, vy), fmul(vx, vt), fmul(vy, vt)};
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        if (k < WIN) np_add<WIN>(s0[q], WIN * i + k, p[q]);
                        if (k >= 1) np_add<WIN>(s1[q], WIN * i + k - 1, p[q]);
                    }
                }
            }
            emit(y, x, np_finish<WIN>(s0[0]), np_finish<WIN>(s0[1]), np_finish<WIN>(s0[2]), np_finish<WIN>(s0[3]), np_finish<WIN>(s0[4]));
            emit(y, x + 1, np_finish<WIN>(s1[0]), np_finish<WIN>(s1[1]), np_finish<WIN>(s1[2]), np_finish<WIN>(s1[3]), np_finish<WIN>(s1[4]));
        }
    } else {
        for (int o = tid; o < TY * TX; o += TILE_THREADS) {
            const int r = o / TX, c = o % TX;
            const float* wx = gx + r * GW + c;
            const float* wy = gy + r * GW + c;
            const float* wt = gt + r * GW + c;
            emit(oy + r, ox + c, numpy_order_sum<WIN>(wx, wx, GW), numpy_order_sum<WIN>(wy, wy, GW),
                 numpy_order_sum<WIN>(wx, wy, GW), numpy_order_sum<WIN>(wx, wt, GW), numpy_order_sum<WIN>(wy, wt, GW));
        }
    }

    if (FLOW && a.partial != nullptr) {
        // deterministic block reduction (fixed shuffle tree, then warps in order)
        __shared__ double red[2][TILE_THREADS / 32];
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
            for (int w = 0; w < TILE_THREADS / 32; ++w) {
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

// One CTA per pair: mean|du|, mean|dv| over the whole level, the reference's early exit,
// and the ping-pong flip.
__global__ void __launch_bounds__(256) iter_finalize_kernel(IterFinalizeArgs a) {
    const int pair = blockIdx.x;
    if (a.done[pair]) return;
    __shared__ double red[2][8];
    const double* part = a.partial + (size_t)pair * a.blocks_per_pair * 2;
    double su = 0.0, sv = 0.0;
    for (int i = threadIdx.x; i < a.blocks_per_pair; i += 256) {
        su += part[2 * i];
        sv += part[2 * i + 1];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        su += __shfl_down_sync(0xffffffffu, su, off);
        sv += __shfl_down_sync(0xffffffffu, sv, off);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = su;
        red[1][threadIdx.x >> 5] = sv;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        su = 0.0;
        sv = 0.0;
        for (int w = 0; w < 8; ++w) {
            su += red[0][w];
            sv += red[1][w];
        }
        const double n = (double)a.H * (double)a.W;
        const float mu = (float)(su / n), mv = (float)(sv / n);
        a.sel[pair] ^= 1;
        if (a.iters_executed) a.iters_executed[(size_t)pair * a.iters_pair_stride] += 1;
        if (a.residuals) {
            a.residuals[(size_t)pair * a.resid_pair_stride + 2 * a.iteration + 0] = mu;
            a.residuals[(size_t)pair * a.resid_pair_stride + 2 * a.iteration + 1] = mv;
        }
        if (mu < OF_CONVERGENCE_EPS && mv < OF_CONVERGENCE_EPS) a.done[pair] = 1;
    }
}

bool lk_tile_window_supported(int window) { return window >= 1 && window <= 11 && (window & 1); }

int lk_tile_blocks_per_pair(int rows, int W) { return ((W + TX - 1) / TX) * ((rows + TY - 1) / TY); }

// per pair: fixed-order float64 sum of the per-block partials -> sums[pair][2]
__global__ void __launch_bounds__(256) sum_partials_kernel(const double* __restrict__ partial, int blocks_per_pair,
                                                            double* __restrict__ sums) {
    const int pair = blockIdx.x;
    __shared__ double red[2][8];
    const double* part = partial + (size_t)pair * blocks_per_pair * 2;
    double su = 0.0, sv = 0.0;
    for (int i = threadIdx.x; i < blocks_per_pair; i += 256) {
        su += part[2 * i];
        sv += part[2 * i + 1];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        su += __shfl_down_sync(0xffffffffu, su, off);
        sv += __shfl_down_sync(0xffffffffu, sv, off);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = su;
        red[1][threadIdx.x >> 5] = sv;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        su = 0.0;
        sv = 0.0;
        for (int w = 0; w < 8; ++w) {
            su += red[0][w];
            sv += red[1][w];
        }
        sums[2 * pair] = su;
        sums[2 * pair + 1] = sv;
    }
}

__global__ void convergence_update_kernel(const double* __restrict__ sums, int batch, double n_pixels, int* sel,
                                          int* done, int* iters_executed, float* residuals, int max_iters,
                                          int iteration) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= batch || done[pair]) return;
    const float mu = (float)(sums[2 * pair] / n_pixels), mv = (float)(sums[2 * pair + 1] / n_pixels);
    sel[pair] ^= 1;
    if (iters_executed) iters_executed[pair] += 1;
    if (residuals) {
        residuals[((size_t)pair * max_iters + iteration) * 2 + 0] = mu;
        residuals[((size_t)pair * max_iters + iteration) * 2 + 1] = mv;
    }
    if (mu < OF_CONVERGENCE_EPS && mv < OF_CONVERGENCE_EPS) done[pair] = 1;
}

cudaError_t launch_convergence_update(const double* sums, int batch, double n_pixels, int* sel, int* done,
                                      int* iters_executed, float* residuals, int max_iters, int iteration,
                                      int* launches, cudaStream_t stream) {
    if (launches) *launches += 1;
    OF_LAUNCH(convergence_update_kernel, (batch + 127) / 128, 128, 0, stream, sums, batch, n_pixels, sel, done, iters_executed,
                                                                       residuals, max_iters, iteration);
    return cudaGetLastError();
}

cudaError_t launch_sum_partials(const double* partial, int blocks_per_pair, double* sums, int batch, int* launches,
                                cudaStream_t stream) {
    if (launches) *launches += 1;
    OF_LAUNCH(sum_partials_kernel, batch, 256, 0, stream, partial, blocks_per_pair, sums);
    return cudaGetLastError();
}

template <int SRC, int WIN>
static cudaError_t launch_one(const TileArgs& a, int batch, cudaStream_t stream) {
    constexpr int HW = WIN / 2;
    constexpr int GW = TX + 2 * HW, GH = TY + 2 * HW;
    constexpr int FW = GW + 2, FH = GH + 2;
    const size_t smem = (size_t)(3 * GH * GW + 2 * FH * FW) * sizeof(float);
    const int rows = (SRC == SRC_WARP || SRC == SRC_WARPED) ? a.row_hi - a.row_lo : a.H;
    if (rows <= 0) return cudaErrorInvalidValue;
    dim3 grid((a.W + TX - 1) / TX, (rows + TY - 1) / TY, batch);
    OF_LAUNCH((lk_tile_kernel<SRC, WIN>), grid, TILE_THREADS, smem, stream, a);
    return cudaGetLastError();
}

template <int SRC>
static cudaError_t launch_src(int window, const TileArgs& a, int batch, cudaStream_t stream) {
    switch (window) {
        case 1: return launch_one<SRC, 1>(a, batch, stream);
        case 3: return launch_one<SRC, 3>(a, batch, stream);
        case 5: return launch_one<SRC, 5>(a, batch, stream);
        case 7: return launch_one<SRC, 7>(a, batch, stream);
        case 9: return launch_one<SRC, 9>(a, batch, stream);
        case 11: return launch_one<SRC, 11>(a, batch, stream);
        default: return cudaErrorInvalidValue;
    }
}

#ifndef OF_TILE_V2_DEFAULT
#define OF_TILE_V2_DEFAULT 1  // lk_tile5_kernel passed the GPU suite (profiles/r01c_pytest_gpu_exact_v2.log)
#endif
// window 5 on frames (single scale, or prev + warped plane): the second version of the kernel (lk_tile5.cu)
// unless OF_B200_TILE=v1.  Same bits; every other window / source takes the kernel above.
static bool tile_v2() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("OF_B200_TILE");
        v = e ? (strcmp(e, "v2") == 0 ? 1 : 0) : OF_TILE_V2_DEFAULT;
    }
    return v == 1;
}

// window 5 on frames: the marching form (lk_exact_march.cu: packed pairs, warp-private shared-memory rings, two rows per
// step) unless OF_B200_EXACT=tile selects the tile kernel (A/B runs).  Same bits.
static bool exact_march() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("OF_B200_EXACT");
        v = (e && strcmp(e, "tile") == 0) ? 0 : 1;
    }
    return v == 1;
}

static bool takes_exact_march(int src, int window, const TileArgs& a) {
    return window == 5 && (src == SRC_FRAMES || src == SRC_WARPED) && (size_t)a.H * a.W < ((size_t)1 << 31) && exact_march();
}

bool lk_tile_fuses_tail(int src, int window, const TileArgs& a) { return src == SRC_WARPED && takes_exact_march(src, window, a); }

cudaError_t launch_lk_tile(int src, int window, const TileArgs& a, int batch, int* launches, cudaStream_t stream) {
    if (batch > 65535) return cudaErrorInvalidValue;
    if (a.tail.counter != nullptr && !lk_tile_fuses_tail(src, window, a)) return cudaErrorInvalidValue;  // nobody would run it
    if (launches) *launches += 1;
    if (takes_exact_march(src, window, a)) return launch_lk_exact_march(src, a, batch, stream);
    if (window == 5 && (src == SRC_FRAMES || src == SRC_WARPED) && (size_t)a.H * a.W < ((size_t)1 << 31) && tile_v2())
        return launch_lk_tile5(src, a, batch, stream);
    switch (src) {
        case SRC_FRAMES: return launch_src<SRC_FRAMES>(window, a, batch, stream);
        case SRC_WARP: return launch_src<SRC_WARP>(window, a, batch, stream);
        case SRC_WARPED: return launch_src<SRC_WARPED>(window, a, batch, stream);
        case SRC_GRADS: return launch_src<SRC_GRADS>(window, a, batch, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_iter_finalize(const IterFinalizeArgs& a, int batch, int* launches, cudaStream_t stream) {
    if (launches) *launches += 1;
    OF_LAUNCH(iter_finalize_kernel, batch, 256, 0, stream, a);
    return cudaGetLastError();
}

}  // namespace ofb
