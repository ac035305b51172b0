// C ABI of libof_b200.so (see include/of_b200.h).  Argument checking, device memory for the
// host-buffer entry points, and the coarse-to-fine driver of the pyramidal path live here;
// the arithmetic lives in the kernel files.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/of_b200.h"
#include "of_kernels.h"

using namespace ofb;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define OF_CUDA(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            cudaGetLastError();                                                                        \
            return fail(_e == cudaErrorMemoryAllocation ? OF_ERR_OUT_OF_MEMORY : OF_ERR_CUDA,          \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                           \
        }                                                                                              \
    } while (0)

#define OF_TRY(expr)               \
    do {                           \
        int _s = (expr);           \
        if (_s != OF_OK) return _s; \
    } while (0)

int need_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(OF_ERR_NO_DEVICE, "no CUDA device is available; this backend has no CPU fallback");
    }
    return OF_OK;
}

struct Counter {
    int n = 0;
    ~Counter() { g_launches += n; }
};

// grow-only device buffers reused by the host-buffer entry points
struct Arena {
    std::vector<void*> ptr;
    std::vector<size_t> cap;
    int get(size_t idx, size_t bytes, void** out) {
        if (ptr.size() <= idx) {
            ptr.resize(idx + 1, nullptr);
            cap.resize(idx + 1, 0);
        }
        if (cap[idx] < bytes) {
            if (ptr[idx]) cudaFree(ptr[idx]);
            ptr[idx] = nullptr;
            cap[idx] = 0;
            cudaError_t e = cudaMalloc(&ptr[idx], bytes ? bytes : 1);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(OF_ERR_OUT_OF_MEMORY, std::string("cudaMalloc: ") + cudaGetErrorString(e));
            }
            cap[idx] = bytes;
        }
        *out = ptr[idx];
        return OF_OK;
    }
    void release() {
        for (void* p : ptr)
            if (p) cudaFree(p);
        ptr.clear();
        cap.clear();
    }
};

// State of the host-buffer entry points, one per device ordinal: buffers and streams belong to the
// device that was current when they were created, so a process (or several threads) driving more
// than one GPU gets one set per GPU.  The mutex serialises the host calls of one device only.
constexpr int HOST_STREAMS = 3;
struct HostPath {
    std::mutex mutex;
    Arena arena;
    cudaStream_t streams[HOST_STREAMS] = {nullptr, nullptr, nullptr};
    int device = -1;
    // small page-locked staging buffer: a device->host copy into pageable memory blocks the calling thread
    // until it has completed, which would serialise the pipelined passes
    void* staging = nullptr;
    size_t staging_cap = 0;
    int get_staging(size_t bytes, void** out) {
        if (staging_cap < bytes) {
            if (staging) cudaFreeHost(staging);
            staging = nullptr;
            staging_cap = 0;
            cudaError_t e = cudaHostAlloc(&staging, bytes ? bytes : 1, cudaHostAllocDefault);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(OF_ERR_OUT_OF_MEMORY, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
            }
            staging_cap = bytes;
        }
        *out = staging;
        return OF_OK;
    }
    // every stream drained: called before a host entry point returns, on success and on failure, so no
    // copy into the caller's or the arena's buffers is in flight once the mutex is released
    cudaError_t drain() {
        cudaError_t first = cudaSuccess;
        for (int i = 0; i < HOST_STREAMS; ++i)
            if (streams[i]) {
                const cudaError_t e = cudaStreamSynchronize(streams[i]);
                if (e != cudaSuccess && first == cudaSuccess) first = e;
            }
        return first;
    }
};
constexpr int MAX_DEVICES = 64;
std::mutex g_host_paths_mutex;
HostPath* g_host_paths[MAX_DEVICES] = {nullptr};

int host_path(HostPath** out) {
    int dev = 0;
    OF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return fail(OF_ERR_UNSUPPORTED, "device ordinal out of range");
    std::lock_guard<std::mutex> lock(g_host_paths_mutex);
    if (!g_host_paths[dev]) {
        g_host_paths[dev] = new HostPath();
        g_host_paths[dev]->device = dev;
    }
    HostPath* hp = g_host_paths[dev];
    for (int i = 0; i < HOST_STREAMS; ++i)
        if (!hp->streams[i]) OF_CUDA(cudaStreamCreateWithFlags(&hp->streams[i], cudaStreamNonBlocking));
    *out = hp;
    return OF_OK;
}

// Drains the device's host-path streams when a host entry point returns -- on the failure paths too, so
// no copy into the caller's or the arena's buffers is still in flight once the mutex is released.
struct HostDrain {
    HostPath* hp;
    ~HostDrain() {
        if (hp->drain() != cudaSuccess) cudaGetLastError();
    }
};

// the current device's host-path state, locked for the rest of the calling scope
#define OF_HOST_PATH(hp)           \
    HostPath* hp = nullptr;        \
    OF_TRY(host_path(&hp));        \
    std::lock_guard<std::mutex> lock(hp->mutex); \
    HostDrain drain_guard{hp}

int check_frame(const void* a, const void* b, int H, int W) {
    if (!a || !b) return fail(OF_ERR_INVALID_ARGUMENT, "null buffer");
    if (H < 1 || W < 1) return fail(OF_ERR_INVALID_ARGUMENT, "height and width must be >= 1");
    if ((long long)H * W > (1LL << 31) - 1) return fail(OF_ERR_INVALID_ARGUMENT, "frame too large");
    return OF_OK;
}

int check_window(int window) {
    if (window < 1 || (window & 1) == 0)
        return fail(OF_ERR_INVALID_ARGUMENT, "window_size must be odd and >= 1");
    if (!lk_tile_window_supported(window))
        return fail(OF_ERR_UNSUPPORTED, "window_size > 11 is not supported (NumPy's summation order changes at 128 taps)");
    return OF_OK;
}

// ---------------------------------------------------------------------------------------
// single scale
// ---------------------------------------------------------------------------------------
int single_scale_dev(const float* prev, const float* curr, float* u, float* v, int batch, int H, int W, int window,
                     int mode, cudaStream_t stream, Counter& cnt) {
    if (mode != OF_MODE_EXACT && mode != OF_MODE_FAST) return fail(OF_ERR_INVALID_ARGUMENT, "unknown mode");
    // the marching kernel moves 128-bit words: all four planes must be 16-byte aligned
    const bool aligned = ((reinterpret_cast<uintptr_t>(prev) | reinterpret_cast<uintptr_t>(curr) |
                           reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (mode == OF_MODE_FAST && aligned && lk_march_supported(H, W, window)) {
        OF_CUDA(launch_lk_march(prev, curr, u, v, batch, H, W, window, 0, &cnt.n, stream));
        return OF_OK;
    }
    TileArgs a;
    memset(&a, 0, sizeof(a));
    a.in0 = prev;
    a.in1 = curr;
    a.out_u = u;
    a.out_v = v;
    a.H = H;
    a.W = W;
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
        const size_t off = (size_t)b0 * H * W;
        TileArgs c = a;
        c.in0 = prev + off;
        c.in1 = curr + off;
        c.out_u = u + off;
        c.out_v = v + off;
        OF_CUDA(launch_lk_tile(SRC_FRAMES, window, c, nb, &cnt.n, stream));
    }
    return OF_OK;
}

// ---------------------------------------------------------------------------------------
// pyramidal driver
// ---------------------------------------------------------------------------------------
struct PyrPlan {
    int L = 0;
    std::vector<int> h, w;                      // k = 0 finest ... L-1 coarsest
    std::vector<size_t> prev_off, curr_off;      // k >= 1
    std::vector<size_t> au_off, av_off, bu_off, bv_off;  // flow ping-pong (A of k = 0 is the caller's u, v)
    std::vector<size_t> warped_off, warped2_off;         // warped current frame (split refinement), ping-pong
    size_t partial_off = 0, sel_off = 0, done_off = 0, cnt_off = 0, total = 0;
    int max_blocks = 0;
};

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
bool refine_chain_warp();  // below: OF_B200_REFINE_WARP=chain needs a second warped plane per level

int make_plan(int batch, int H, int W, int levels, PyrPlan& p) {
    if (levels < 1 || levels > 16) return fail(OF_ERR_INVALID_ARGUMENT, "num_levels must be in 1..16");
    p.L = levels;
    p.h.assign(levels, 0);
    p.w.assign(levels, 0);
    p.h[0] = H;
    p.w[0] = W;
    for (int k = 1; k < levels; ++k) {
        p.h[k] = p.h[k - 1] / 2;  // int(h * 0.5)
        p.w[k] = p.w[k - 1] / 2;
        if (p.h[k] < 1 || p.w[k] < 1) return fail(OF_ERR_INVALID_ARGUMENT, "frame too small for this many pyramid levels");
    }
    size_t off = 0;
    p.prev_off.assign(levels, 0);
    p.curr_off.assign(levels, 0);
    p.au_off.assign(levels, 0);
    p.av_off.assign(levels, 0);
    p.bu_off.assign(levels, 0);
    p.bv_off.assign(levels, 0);
    p.warped_off.assign(levels, 0);
    p.warped2_off.assign(levels, 0);
    p.max_blocks = 0;
    for (int k = 0; k < levels; ++k) {
        const size_t bytes = align_up((size_t)batch * p.h[k] * p.w[k] * sizeof(float));
        if (k >= 1) {
            p.prev_off[k] = off; off += bytes;
            p.curr_off[k] = off; off += bytes;
            p.au_off[k] = off; off += bytes;
            p.av_off[k] = off; off += bytes;
        }
        p.bu_off[k] = off; off += bytes;
        p.bv_off[k] = off; off += bytes;
        p.warped_off[k] = off; off += bytes;
        p.warped2_off[k] = p.warped_off[k];
        if (refine_chain_warp()) { p.warped2_off[k] = off; off += bytes; }
        const int nb = lk_tile_blocks_per_pair(p.h[k], p.w[k]);
        if (nb > p.max_blocks) p.max_blocks = nb;
    }
    p.partial_off = off; off += align_up((size_t)batch * p.max_blocks * 2 * sizeof(double));
    p.sel_off = off; off += align_up((size_t)levels * batch * sizeof(int));
    p.done_off = off; off += align_up((size_t)levels * batch * sizeof(int));
    p.cnt_off = off; off += align_up((size_t)batch * sizeof(unsigned));  // ticket counters of the fused iteration tail
    p.total = off;
    return OF_OK;
}

// fast-mode refinement: "split" = warp kernel + K1 marching kernel (default), "fused" = one
// kernel that gathers the warped frame itself.  Both are kept; OF_B200_REFINE picks.
bool refine_split() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("OF_B200_REFINE");
        v = (e && std::string(e) == "fused") ? 0 : 1;
    }
    return v == 1;
}

// OF_B200_REFINE=ws: the split form's marching kernel with producer warps that gather the warped rows into its ring
// stages (lk_march_kernel<..., WS>): no warped plane, one launch per iteration.  Window 5; other windows stay split.
bool refine_ws() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("OF_B200_REFINE");
        v = (e && std::string(e) == "ws") ? 1 : 0;
    }
    return v == 1;
}
cudaError_t launch_refine_split_form(const RefineArgs& ra, float* warped, int batch, int* launches, cudaStream_t stream) {
    if (refine_ws() && ra.window == 5 && !ra.warped_ready && ra.warped_next == nullptr)
        return launch_lk_refine_ws(ra, batch, launches, stream);
    return launch_lk_refine_split(ra, warped, batch, launches, stream);
}

#ifndef OF_EXACT_REFINE_SPLIT_DEFAULT
#define OF_EXACT_REFINE_SPLIT_DEFAULT 1  // split passed the GPU suite (profiles/r01c_pytest_gpu_exact_v2.log)
#endif
// split refinement, OF_B200_REFINE_WARP=chain: the marching kernel warps the next iteration's input itself (its
// epilogue gathers through the flow it has in registers), so only a level's first iteration launches warp_rows.
// Measured and NOT the default: the marching warps are latency-bound at 8 per SM, so every instruction added to them
// costs its full latency -- 4 x 4K finest level: 308 us per iteration against 165 + 124 us for the two launches
// (profiles/r02_refine_chain_*).  Kept behind the switch with its bit-equality test.
bool refine_chain_warp() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("OF_B200_REFINE_WARP");
        v = (e && std::string(e) == "chain") ? 1 : 0;
    }
    return v == 1;
}

// exact-mode refinement: "split" = warp kernel with float64 fractions + the tile kernel on (prev, warped)
// ; "fused" = the tile kernel gathers its halo tile itself (the first implementation: 1.5 gathers per
// pixel at 16 warps per SM, latency-bound).  Same bits; OF_B200_EXACT_REFINE=split|fused picks.
bool exact_refine_split() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("OF_B200_EXACT_REFINE");
        v = e ? (std::string(e) == "split" ? 1 : 0) : OF_EXACT_REFINE_SPLIT_DEFAULT;
    }
    return v == 1;
}

int pyramidal_dev(const float* prev, const float* curr, float* u, float* v, int batch, int H, int W, int levels,
                  int window, int iterations, int mode, const double* gw, int radius, void* workspace, size_t ws_bytes,
                  int* iters_dev, float* resid_dev, cudaStream_t stream, Counter& cnt) {
    if (mode != OF_MODE_EXACT && mode != OF_MODE_FAST) return fail(OF_ERR_INVALID_ARGUMENT, "unknown mode");
    if (iterations < 0) return fail(OF_ERR_INVALID_ARGUMENT, "num_iterations must be >= 0");
    if (batch > 65535) return fail(OF_ERR_INVALID_ARGUMENT, "at most 65535 pairs per call on the device entry point");
    if (levels > 1 && (!gw || radius < 0 || radius > OF_MAX_GAUSS_RADIUS))
        return fail(OF_ERR_INVALID_ARGUMENT, "gaussian weights missing or radius out of range");
    PyrPlan p;
    OF_TRY(make_plan(batch, H, W, levels, p));
    if (!workspace || ws_bytes < p.total) return fail(OF_ERR_INVALID_ARGUMENT, "workspace too small");
    char* ws = static_cast<char*>(workspace);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    int* sel = reinterpret_cast<int*>(ws + p.sel_off);
    int* done = reinterpret_cast<int*>(ws + p.done_off);
    double* partial = reinterpret_cast<double*>(ws + p.partial_off);

    OF_CUDA(cudaMemsetAsync(ws + p.sel_off, 0, p.total - p.sel_off, stream));
    if (iters_dev) OF_CUDA(cudaMemsetAsync(iters_dev, 0, (size_t)batch * levels * sizeof(int), stream));
    if (resid_dev && iterations > 0)
        OF_CUDA(cudaMemsetAsync(resid_dev, 0, (size_t)batch * levels * iterations * 2 * sizeof(float), stream));

    // Gaussian pyramids of both frames, fine -> coarse
    std::vector<const float*> lp(levels), lc(levels);
    lp[0] = prev;
    lc[0] = curr;
    for (int k = 1; k < levels; ++k) {
        OF_CUDA(launch_pyramid_down(lp[k - 1], F(p.prev_off[k]), batch, p.h[k - 1], p.w[k - 1], p.h[k], p.w[k], gw,
                                    radius, 0, p.h[k], &cnt.n, stream, mode == OF_MODE_FAST));
        OF_CUDA(launch_pyramid_down(lc[k - 1], F(p.curr_off[k]), batch, p.h[k - 1], p.w[k - 1], p.h[k], p.w[k], gw,
                                    radius, 0, p.h[k], &cnt.n, stream, mode == OF_MODE_FAST));
        lp[k] = F(p.prev_off[k]);
        lc[k] = F(p.curr_off[k]);
    }
    auto flowA_u = [&](int k) { return k == 0 ? u : F(p.au_off[k]); };
    auto flowA_v = [&](int k) { return k == 0 ? v : F(p.av_off[k]); };
    auto flow_u = [&](int k, int idx) { return idx ? F(p.bu_off[k]) : flowA_u(k); };
    auto flow_v = [&](int k, int idx) { return idx ? F(p.bv_off[k]) : flowA_v(k); };
    // Each iteration writes the other ping-pong buffer.  Starting every level in buffer
    // (iterations & 1) makes a pair that runs all its iterations end in buffer 0, which at the
    // finest level is the caller's (u, v): no final copy unless a pair stopped early.
    const int start = iterations & 1;

    // zero flow at the coarsest level
    const int kc = levels - 1;
    const size_t coarse_bytes = (size_t)batch * p.h[kc] * p.w[kc] * sizeof(float);
    OF_CUDA(cudaMemsetAsync(flow_u(kc, start), 0, coarse_bytes, stream));
    OF_CUDA(cudaMemsetAsync(flow_v(kc, start), 0, coarse_bytes, stream));

    for (int k = kc; k >= 0; --k) {
        const int ref_level = kc - k;  // the reference counts levels from the coarsest
        int* sel_k = sel + (size_t)k * batch;
        int* done_k = done + (size_t)k * batch;
        if (k < kc) {
            OF_CUDA(launch_upsample_flow(flowA_u(k + 1), flowA_v(k + 1), F(p.bu_off[k + 1]), F(p.bv_off[k + 1]),
                                         sel + (size_t)(k + 1) * batch, start, flow_u(k, start), flow_v(k, start), batch,
                                         p.h[k + 1], p.w[k + 1], p.h[k], p.w[k], 0, p.h[k], &cnt.n, stream));
        }
        // fast mode: the register-marching kernel where the level allows TMA (width % 4 == 0,
        // window 5 or 7); otherwise, and always in exact mode, the reference-order kernels
        RefineArgs ra;
        memset(&ra, 0, sizeof(ra));
        ra.prev = lp[k];
        ra.curr = lc[k];
        ra.flow_u[0] = flowA_u(k);
        ra.flow_v[0] = flowA_v(k);
        ra.flow_u[1] = F(p.bu_off[k]);
        ra.flow_v[1] = F(p.bv_off[k]);
        ra.sel = sel_k;
        ra.sel_xor = start;
        ra.done = done_k;
        ra.partial = partial;
        ra.H = p.h[k];
        ra.W = p.w[k];
        ra.window = window;
        ra.row_lo = 0;
        ra.row_hi = p.h[k];
        ra.own_lo = 0;
        ra.own_hi = p.h[k];
        const bool fast_level = (mode == OF_MODE_FAST) && lk_refine_supported(ra, window);
        // split refinement: the marching kernel's last warp per pair also does the convergence step
        const bool fused_tail = fast_level && refine_split();
        // opt-in: its epilogue also warps the current frame through the flow it has just produced, so only the level's
        // first iteration launches warp_rows; the two warped planes alternate (bands run independently)
        const bool chain_warp = fused_tail && refine_chain_warp();
        for (int it = 0; it < iterations; ++it) {
            if (fast_level) {
                float* warped_it = F((chain_warp && (it & 1)) ? p.warped2_off[k] : p.warped_off[k]);
                ra.warped_ready = chain_warp && it > 0;
                ra.warped_next = (chain_warp && it + 1 < iterations) ? F((it & 1) ? p.warped_off[k] : p.warped2_off[k]) : nullptr;
                if (fused_tail) {
                    ra.tail.counter = reinterpret_cast<unsigned*>(ws + p.cnt_off);
                    ra.tail.peers = 0;
                    ra.tail.n_pixels = (double)p.h[k] * (double)p.w[k];
                    ra.tail.sel = sel_k;
                    ra.tail.done = done_k;
                    ra.tail.iters_executed = iters_dev ? iters_dev + ref_level : nullptr;
                    ra.tail.iters_pair_stride = levels;
                    ra.tail.residuals = resid_dev ? resid_dev + (size_t)ref_level * iterations * 2 : nullptr;
                    ra.tail.resid_pair_stride = (size_t)levels * iterations * 2;
                    ra.tail.iteration = it;
                    OF_CUDA(launch_refine_split_form(ra, warped_it, batch, &cnt.n, stream));
                    continue;
                }
                if (refine_split())
                    OF_CUDA(launch_refine_split_form(ra, warped_it, batch, &cnt.n, stream));
                else
                    OF_CUDA(launch_lk_refine(ra, batch, &cnt.n, stream));
            } else {
                TileArgs a;
                memset(&a, 0, sizeof(a));
                a.in0 = lp[k];
                a.in1 = lc[k];
                a.flow_u[0] = flowA_u(k);
                a.flow_v[0] = flowA_v(k);
                a.flow_u[1] = F(p.bu_off[k]);
                a.flow_v[1] = F(p.bv_off[k]);
                a.sel = sel_k;
                a.sel_xor = start;
                a.done = done_k;
                a.partial = partial;
                a.H = p.h[k];
                a.W = p.w[k];
                a.row_lo = 0;
                a.row_hi = p.h[k];
                a.own_lo = 0;
                a.own_hi = p.h[k];
                if (exact_refine_split()) {
                    // the whole level warped once (coalesced gathers, 4 samples per thread), then the tile
                    // kernel reads it like a frame: its replicated border is the warp at the clamped pixel
                    OF_CUDA(launch_warp_rows(ra, F(p.warped_off[k]), 0, p.h[k], true, batch, &cnt.n, stream));
                    a.in1 = F(p.warped_off[k]);
                    // the marching exact kernel finishes the iteration itself (its pair's last unit reduces the
                    // partial sums and applies the convergence test): no iter_finalize launch
                    const bool tile_tail = lk_tile_fuses_tail(SRC_WARPED, window, a);
                    if (tile_tail) {
                        a.tail.counter = reinterpret_cast<unsigned*>(ws + p.cnt_off);
                        a.tail.peers = 0;
                        a.tail.n_pixels = (double)p.h[k] * (double)p.w[k];
                        a.tail.sel = sel_k;
                        a.tail.done = done_k;
                        a.tail.iters_executed = iters_dev ? iters_dev + ref_level : nullptr;
                        a.tail.iters_pair_stride = levels;
                        a.tail.residuals = resid_dev ? resid_dev + (size_t)ref_level * iterations * 2 : nullptr;
                        a.tail.resid_pair_stride = (size_t)levels * iterations * 2;
                        a.tail.iteration = it;
                    }
                    OF_CUDA(launch_lk_tile(SRC_WARPED, window, a, batch, &cnt.n, stream));
                    if (tile_tail) continue;
                } else {
                    OF_CUDA(launch_lk_tile(SRC_WARP, window, a, batch, &cnt.n, stream));
                }
            }
            IterFinalizeArgs f;
            f.partial = partial;
            f.blocks_per_pair = fast_level ? lk_refine_units_per_pair(batch, p.h[k], p.w[k])
                                           : lk_tile_blocks_per_pair(p.h[k], p.w[k]);
            f.H = p.h[k];
            f.W = p.w[k];
            f.sel = sel_k;
            f.done = done_k;
            f.iters_executed = iters_dev ? iters_dev + ref_level : nullptr;
            f.iters_pair_stride = levels;
            f.residuals = resid_dev ? resid_dev + (size_t)ref_level * iterations * 2 : nullptr;
            f.resid_pair_stride = (size_t)levels * iterations * 2;
            f.iteration = it;
            OF_CUDA(launch_iter_finalize(f, batch, &cnt.n, stream));
        }
    }
    // the finest level's current buffer -> caller's (u, v) (no-op for pairs already there)
    OF_CUDA(launch_select_copy(u, v, F(p.bu_off[0]), F(p.bv_off[0]), sel, start, u, v, batch, (size_t)H * W, &cnt.n,
                               stream));
    return OF_OK;
}

}  // namespace

// =======================================================================================
// exported C ABI
// =======================================================================================
extern "C" {

int of_version(void) { return 100; }

const char* of_last_error(void) { return g_err.c_str(); }

int of_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int of_set_device(int ordinal) {
    OF_TRY(need_device());
    OF_CUDA(cudaSetDevice(ordinal));
    return OF_OK;
}

long long of_kernel_launches(void) { return g_launches.load(); }

int of_host_alloc_pinned(void** ptr, size_t bytes) {
    if (!ptr) return fail(OF_ERR_INVALID_ARGUMENT, "null pointer");
    OF_TRY(need_device());
    // portable: pinned for every CUDA context of the process (a process that drives several GPUs)
    OF_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return OF_OK;
}

int of_release_host_buffers(void) {
    int dev = 0;
    OF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return OF_OK;
    HostPath* hp = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_host_paths_mutex);
        hp = g_host_paths[dev];
    }
    if (!hp) return OF_OK;
    std::lock_guard<std::mutex> lock(hp->mutex);
    if (hp->drain() != cudaSuccess) cudaGetLastError();
    hp->arena.release();
    if (hp->staging) cudaFreeHost(hp->staging);
    hp->staging = nullptr;
    hp->staging_cap = 0;
    return OF_OK;
}

int of_host_free_pinned(void* ptr) {
    if (ptr) OF_CUDA(cudaFreeHost(ptr));
    return OF_OK;
}

int of_lk_single_scale_f32_dev(const float* prev, const float* curr, float* u, float* v, int batch, int height,
                               int width, int window, int mode, void* stream) {
    OF_TRY(check_frame(prev, curr, height, width));
    OF_TRY(check_frame(u, v, height, width));
    OF_TRY(check_window(window));
    if (batch < 0) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be >= 0");
    if (batch == 0) return OF_OK;
    OF_TRY(need_device());
    Counter cnt;
    return single_scale_dev(prev, curr, u, v, batch, height, width, window, mode, static_cast<cudaStream_t>(stream), cnt);
}

int of_lk_single_scale_f32(const float* prev, const float* curr, float* u, float* v, int batch, int height, int width,
                           int window, int mode) {
    OF_TRY(check_frame(prev, curr, height, width));
    OF_TRY(check_frame(u, v, height, width));
    OF_TRY(check_window(window));
    if (batch < 0) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be >= 0");
    if (batch == 0) return OF_OK;
    OF_TRY(need_device());
    OF_HOST_PATH(hp);
    Counter cnt;
    const size_t plane = (size_t)height * width;
    // chunks of <= 64 MiB per array, three in flight: H2D of chunk i+1 and D2H of chunk i-1
    // overlap the kernel of chunk i (needs pinned host memory to actually overlap)
    static const size_t chunk_mib = [] {
        const char* e = getenv("OF_B200_CHUNK_MB");  // experiments: pipeline granularity of the host path
        const long v = e ? atol(e) : 0;
        return (size_t)(v >= 1 && v <= 1024 ? v : 64);
    }();
    size_t per_chunk = (chunk_mib << 20) / (plane * sizeof(float));
    if (per_chunk < 1) per_chunk = 1;
    if (per_chunk > (size_t)batch) per_chunk = batch;
    const size_t chunk_bytes = per_chunk * plane * sizeof(float);
    const int n_chunks = (int)((batch + per_chunk - 1) / per_chunk);
    const int slots = n_chunks < 3 ? n_chunks : 3;
    float* d[3][4];
    for (int s = 0; s < slots; ++s)
        for (int j = 0; j < 4; ++j) OF_TRY(hp->arena.get(s * 4 + j, chunk_bytes, reinterpret_cast<void**>(&d[s][j])));
    for (int c = 0; c < n_chunks; ++c) {
        const int s = c % slots;
        cudaStream_t st = hp->streams[s];
        const size_t b0 = (size_t)c * per_chunk;
        const int nb = (int)((size_t)batch - b0 < per_chunk ? (size_t)batch - b0 : per_chunk);
        const size_t bytes = (size_t)nb * plane * sizeof(float);
        OF_CUDA(cudaMemcpyAsync(d[s][0], prev + b0 * plane, bytes, cudaMemcpyHostToDevice, st));
        OF_CUDA(cudaMemcpyAsync(d[s][1], curr + b0 * plane, bytes, cudaMemcpyHostToDevice, st));
        OF_TRY(single_scale_dev(d[s][0], d[s][1], d[s][2], d[s][3], nb, height, width, window, mode, st, cnt));
        OF_CUDA(cudaMemcpyAsync(u + b0 * plane, d[s][2], bytes, cudaMemcpyDeviceToHost, st));
        OF_CUDA(cudaMemcpyAsync(v + b0 * plane, d[s][3], bytes, cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < slots; ++s) OF_CUDA(cudaStreamSynchronize(hp->streams[s]));
    return OF_OK;
}

int of_gradients_f32(const float* prev, const float* curr, float* ix, float* iy, float* it, int height, int width) {
    OF_TRY(check_frame(prev, curr, height, width));
    OF_TRY(check_frame(ix, iy, height, width));
    if (!it) return fail(OF_ERR_INVALID_ARGUMENT, "null buffer");
    OF_TRY(need_device());
    OF_HOST_PATH(hp);
    Counter cnt;
    const size_t bytes = (size_t)height * width * sizeof(float);
    float* d[5];
    for (int j = 0; j < 5; ++j) OF_TRY(hp->arena.get(j, bytes, reinterpret_cast<void**>(&d[j])));
    cudaStream_t st = hp->streams[0];
    OF_CUDA(cudaMemcpyAsync(d[0], prev, bytes, cudaMemcpyHostToDevice, st));
    OF_CUDA(cudaMemcpyAsync(d[1], curr, bytes, cudaMemcpyHostToDevice, st));
    OF_CUDA(launch_gradients(d[0], d[1], d[2], d[3], d[4], 1, height, width, &cnt.n, st));
    OF_CUDA(cudaMemcpyAsync(ix, d[2], bytes, cudaMemcpyDeviceToHost, st));
    OF_CUDA(cudaMemcpyAsync(iy, d[3], bytes, cudaMemcpyDeviceToHost, st));
    OF_CUDA(cudaMemcpyAsync(it, d[4], bytes, cudaMemcpyDeviceToHost, st));
    OF_CUDA(cudaStreamSynchronize(st));
    return OF_OK;
}

int of_lk_from_gradients_f32(const float* ix, const float* iy, const float* it, float* u, float* v, int height,
                             int width, int window) {
    OF_TRY(check_frame(ix, iy, height, width));
    OF_TRY(check_frame(u, v, height, width));
    if (!it) return fail(OF_ERR_INVALID_ARGUMENT, "null buffer");
    OF_TRY(check_window(window));
    OF_TRY(need_device());
    OF_HOST_PATH(hp);
    Counter cnt;
    const size_t bytes = (size_t)height * width * sizeof(float);
    float* d[5];
    for (int j = 0; j < 5; ++j) OF_TRY(hp->arena.get(j, bytes, reinterpret_cast<void**>(&d[j])));
    cudaStream_t st = hp->streams[0];
    OF_CUDA(cudaMemcpyAsync(d[0], ix, bytes, cudaMemcpyHostToDevice, st));
    OF_CUDA(cudaMemcpyAsync(d[1], iy, bytes, cudaMemcpyHostToDevice, st));
    OF_CUDA(cudaMemcpyAsync(d[2], it, bytes, cudaMemcpyHostToDevice, st));
    TileArgs a;
    memset(&a, 0, sizeof(a));
    a.in0 = d[0];
    a.in1 = d[1];
    a.in2 = d[2];
    a.out_u = d[3];
    a.out_v = d[4];
    a.H = height;
    a.W = width;
    OF_CUDA(launch_lk_tile(SRC_GRADS, window, a, 1, &cnt.n, st));
    OF_CUDA(cudaMemcpyAsync(u, d[3], bytes, cudaMemcpyDeviceToHost, st));
    OF_CUDA(cudaMemcpyAsync(v, d[4], bytes, cudaMemcpyDeviceToHost, st));
    OF_CUDA(cudaStreamSynchronize(st));
    return OF_OK;
}

int of_pyramid_down_f32(const float* src, float* dst, int height, int width, int out_height, int out_width,
                        const double* weights, int radius) {
    OF_TRY(check_frame(src, dst, height, width));
    if (out_height < 1 || out_width < 1) return fail(OF_ERR_INVALID_ARGUMENT, "output size must be >= 1");
    if (!weights || radius < 0 || radius > OF_MAX_GAUSS_RADIUS)
        return fail(OF_ERR_INVALID_ARGUMENT, "gaussian weights missing or radius out of range");
    OF_TRY(need_device());
    OF_HOST_PATH(hp);
    Counter cnt;
    const size_t ib = (size_t)height * width * sizeof(float), ob = (size_t)out_height * out_width * sizeof(float);
    float *ds, *dd;
    OF_TRY(hp->arena.get(0, ib, reinterpret_cast<void**>(&ds)));
    OF_TRY(hp->arena.get(1, ob, reinterpret_cast<void**>(&dd)));
    cudaStream_t st = hp->streams[0];
    OF_CUDA(cudaMemcpyAsync(ds, src, ib, cudaMemcpyHostToDevice, st));
    OF_CUDA(launch_pyramid_down(ds, dd, 1, height, width, out_height, out_width, weights, radius, 0, out_height, &cnt.n,
                                st));
    OF_CUDA(cudaMemcpyAsync(dst, dd, ob, cudaMemcpyDeviceToHost, st));
    OF_CUDA(cudaStreamSynchronize(st));
    return OF_OK;
}

int of_warp_f32(const float* image, const float* flow_u, const float* flow_v, float* out, int height, int width) {
    OF_TRY(check_frame(image, out, height, width));
    OF_TRY(check_frame(flow_u, flow_v, height, width));
    OF_TRY(need_device());
    OF_HOST_PATH(hp);
    Counter cnt;
    const size_t bytes = (size_t)height * width * sizeof(float);
    float* d[4];
    for (int j = 0; j < 4; ++j) OF_TRY(hp->arena.get(j, bytes, reinterpret_cast<void**>(&d[j])));
    cudaStream_t st = hp->streams[0];
    OF_CUDA(cudaMemcpyAsync(d[0], image, bytes, cudaMemcpyHostToDevice, st));
    OF_CUDA(cudaMemcpyAsync(d[1], flow_u, bytes, cudaMemcpyHostToDevice, st));
    OF_CUDA(cudaMemcpyAsync(d[2], flow_v, bytes, cudaMemcpyHostToDevice, st));
    OF_CUDA(launch_warp(d[0], d[1], d[2], d[3], 1, height, width, &cnt.n, st));
    OF_CUDA(cudaMemcpyAsync(out, d[3], bytes, cudaMemcpyDeviceToHost, st));
    OF_CUDA(cudaStreamSynchronize(st));
    return OF_OK;
}

int of_upsample_flow_f32(const float* coarse_u, const float* coarse_v, float* u, float* v, int coarse_height,
                         int coarse_width, int target_height, int target_width) {
    OF_TRY(check_frame(coarse_u, coarse_v, coarse_height, coarse_width));
    OF_TRY(check_frame(u, v, target_height, target_width));
    OF_TRY(need_device());
    OF_HOST_PATH(hp);
    Counter cnt;
    const size_t cb = (size_t)coarse_height * coarse_width * sizeof(float);
    const size_t tb = (size_t)target_height * target_width * sizeof(float);
    float* d[4];
    OF_TRY(hp->arena.get(0, cb, reinterpret_cast<void**>(&d[0])));
    OF_TRY(hp->arena.get(1, cb, reinterpret_cast<void**>(&d[1])));
    OF_TRY(hp->arena.get(2, tb, reinterpret_cast<void**>(&d[2])));
    OF_TRY(hp->arena.get(3, tb, reinterpret_cast<void**>(&d[3])));
    cudaStream_t st = hp->streams[0];
    OF_CUDA(cudaMemcpyAsync(d[0], coarse_u, cb, cudaMemcpyHostToDevice, st));
    OF_CUDA(cudaMemcpyAsync(d[1], coarse_v, cb, cudaMemcpyHostToDevice, st));
    OF_CUDA(launch_upsample_flow(d[0], d[1], nullptr, nullptr, nullptr, 0, d[2], d[3], 1, coarse_height, coarse_width,
                                 target_height, target_width, 0, target_height, &cnt.n, st));
    OF_CUDA(cudaMemcpyAsync(u, d[2], tb, cudaMemcpyDeviceToHost, st));
    OF_CUDA(cudaMemcpyAsync(v, d[3], tb, cudaMemcpyDeviceToHost, st));
    OF_CUDA(cudaStreamSynchronize(st));
    return OF_OK;
}

size_t of_lk_pyramidal_workspace_bytes(int batch, int height, int width, int levels, int iterations) {
    (void)iterations;
    PyrPlan p;
    if (batch < 1 || height < 1 || width < 1 || make_plan(batch, height, width, levels, p) != OF_OK) return 0;
    return p.total;
}

int of_lk_pyramidal_f32_dev(const float* prev, const float* curr, float* u, float* v, int batch, int height, int width,
                            int levels, int window, int iterations, int mode, const double* gauss_weights,
                            int gauss_radius, void* workspace, size_t workspace_bytes, int* iters_executed_dev,
                            float* residuals_dev, void* stream) {
    OF_TRY(check_frame(prev, curr, height, width));
    OF_TRY(check_frame(u, v, height, width));
    OF_TRY(check_window(window));
    if (batch < 0) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be >= 0");
    if (batch == 0) return OF_OK;
    OF_TRY(need_device());
    Counter cnt;
    return pyramidal_dev(prev, curr, u, v, batch, height, width, levels, window, iterations, mode, gauss_weights,
                         gauss_radius, workspace, workspace_bytes, iters_executed_dev, residuals_dev,
                         static_cast<cudaStream_t>(stream), cnt);
}

int of_lk_pyramidal_f32(const float* prev, const float* curr, float* u, float* v, int batch, int height, int width,
                        int levels, int window, int iterations, int mode, const double* gauss_weights,
                        int gauss_radius, int* iters_executed, float* residuals) {
    OF_TRY(check_frame(prev, curr, height, width));
    OF_TRY(check_frame(u, v, height, width));
    OF_TRY(check_window(window));
    if (batch < 0) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be >= 0");
    if (iterations < 0) return fail(OF_ERR_INVALID_ARGUMENT, "num_iterations must be >= 0");
    if (batch == 0) return OF_OK;
    OF_TRY(need_device());
    OF_HOST_PATH(hp);
    Counter cnt;
    const size_t plane = (size_t)height * width;
    // Passes of a few pairs each, three in flight on three streams with their own frames, workspace and trace
    // buffers: H2D of pass i+1 and D2H of pass i-1 overlap the kernels of pass i (pinned host memory needed
    // for the copies to actually overlap).  A pass holds at most 128 MiB per frame stack, and a batch is cut
    // into at least six passes when it has that many pairs, so that the pipeline has something to overlap.
    size_t per_pass = ((size_t)128 << 20) / (plane * sizeof(float));
    if (per_pass < 1) per_pass = 1;
    const size_t sixth = ((size_t)batch + 5) / 6;
    if (per_pass > sixth) per_pass = sixth;
    if (per_pass > 65535) per_pass = 65535;
    const int n_pass = (int)(((size_t)batch + per_pass - 1) / per_pass);
    const int slots = n_pass < HOST_STREAMS ? n_pass : HOST_STREAMS;
    PyrPlan plan;
    OF_TRY(make_plan((int)per_pass, height, width, levels, plan));
    const size_t fb = per_pass * plane * sizeof(float);
    const size_t ib = per_pass * levels * sizeof(int);
    const size_t rb = per_pass * levels * (size_t)(iterations > 0 ? iterations : 1) * 2 * sizeof(float);
    float* d[HOST_STREAMS][4];
    void* ws[HOST_STREAMS];
    int* d_iters[HOST_STREAMS] = {nullptr, nullptr, nullptr};
    float* d_res[HOST_STREAMS] = {nullptr, nullptr, nullptr};
    for (int s = 0; s < slots; ++s) {
        const size_t base = 32 + (size_t)s * 8;
        for (int j = 0; j < 4; ++j) OF_TRY(hp->arena.get(base + j, fb, reinterpret_cast<void**>(&d[s][j])));
        OF_TRY(hp->arena.get(base + 4, plan.total, &ws[s]));
        if (iters_executed) OF_TRY(hp->arena.get(base + 5, ib, reinterpret_cast<void**>(&d_iters[s])));
        if (residuals) OF_TRY(hp->arena.get(base + 6, rb, reinterpret_cast<void**>(&d_res[s])));
    }
    // the per-pair trace (a few bytes) goes through page-locked staging memory and reaches the caller's
    // arrays after the last pass: a copy into pageable memory would stall the pipeline at every pass
    const size_t it_total = iters_executed ? (size_t)batch * levels * sizeof(int) : 0;
    const size_t rs_total = (residuals && iterations > 0) ? (size_t)batch * levels * iterations * 2 * sizeof(float) : 0;
    char* stage_mem = nullptr;
    if (it_total + rs_total) OF_TRY(hp->get_staging(it_total + rs_total, reinterpret_cast<void**>(&stage_mem)));
    int* st_iters = reinterpret_cast<int*>(stage_mem);
    float* st_res = reinterpret_cast<float*>(stage_mem + it_total);
    for (int c = 0; c < n_pass; ++c) {
        const int s = c % slots;
        cudaStream_t st = hp->streams[s];
        const size_t b0 = (size_t)c * per_pass;
        const int nb = (int)((size_t)batch - b0 < per_pass ? (size_t)batch - b0 : per_pass);
        const size_t bytes = (size_t)nb * plane * sizeof(float);
        OF_CUDA(cudaMemcpyAsync(d[s][0], prev + b0 * plane, bytes, cudaMemcpyHostToDevice, st));
        OF_CUDA(cudaMemcpyAsync(d[s][1], curr + b0 * plane, bytes, cudaMemcpyHostToDevice, st));
        PyrPlan pp;
        OF_TRY(make_plan(nb, height, width, levels, pp));
        OF_TRY(pyramidal_dev(d[s][0], d[s][1], d[s][2], d[s][3], nb, height, width, levels, window, iterations, mode,
                             gauss_weights, gauss_radius, ws[s], pp.total, d_iters[s], d_res[s], st, cnt));
        OF_CUDA(cudaMemcpyAsync(u + b0 * plane, d[s][2], bytes, cudaMemcpyDeviceToHost, st));
        OF_CUDA(cudaMemcpyAsync(v + b0 * plane, d[s][3], bytes, cudaMemcpyDeviceToHost, st));
        if (it_total)
            OF_CUDA(cudaMemcpyAsync(st_iters + b0 * levels, d_iters[s], (size_t)nb * levels * sizeof(int),
                                    cudaMemcpyDeviceToHost, st));
        if (rs_total)
            OF_CUDA(cudaMemcpyAsync(st_res + b0 * levels * iterations * 2, d_res[s],
                                    (size_t)nb * levels * iterations * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < slots; ++s) OF_CUDA(cudaStreamSynchronize(hp->streams[s]));
    if (it_total) memcpy(iters_executed, st_iters, it_total);
    if (rs_total) memcpy(residuals, st_res, rs_total);
    return OF_OK;
}

int of_pyramid_down_f32_dev(const float* src, float* dst, int batch, int height, int width, int out_height,
                            int out_width, const double* weights, int radius, int row_lo, int row_hi, int mode,
                            void* stream) {
    OF_TRY(check_frame(src, dst, height, width));
    if (out_height < 1 || out_width < 1) return fail(OF_ERR_INVALID_ARGUMENT, "output size must be >= 1");
    if (!weights || radius < 0 || radius > OF_MAX_GAUSS_RADIUS)
        return fail(OF_ERR_INVALID_ARGUMENT, "gaussian weights missing or radius out of range");
    if (batch < 1 || batch > 65535) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be in 1..65535");
    OF_TRY(need_device());
    Counter cnt;
    if (row_lo < 0 || row_hi > out_height || row_lo >= row_hi) return fail(OF_ERR_INVALID_ARGUMENT, "bad output row range");
    if (mode != OF_MODE_EXACT && mode != OF_MODE_FAST) return fail(OF_ERR_INVALID_ARGUMENT, "unknown mode");
    OF_CUDA(launch_pyramid_down(src, dst, batch, height, width, out_height, out_width, weights, radius, row_lo, row_hi,
                                &cnt.n, static_cast<cudaStream_t>(stream), mode == OF_MODE_FAST));
    return OF_OK;
}

int of_upsample_flow_f32_dev(const float* coarse_u, const float* coarse_v, float* u, float* v, int batch,
                             int coarse_height, int coarse_width, int target_height, int target_width, int row_lo,
                             int row_hi, void* stream) {
    OF_TRY(check_frame(coarse_u, coarse_v, coarse_height, coarse_width));
    OF_TRY(check_frame(u, v, target_height, target_width));
    if (batch < 1 || batch > 65535) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be in 1..65535");
    if (row_lo < 0 || row_hi > target_height || row_lo >= row_hi)
        return fail(OF_ERR_INVALID_ARGUMENT, "bad target row range");
    OF_TRY(need_device());
    Counter cnt;
    OF_CUDA(launch_upsample_flow(coarse_u, coarse_v, nullptr, nullptr, nullptr, 0, u, v, batch, coarse_height,
                                 coarse_width, target_height, target_width, row_lo, row_hi, &cnt.n,
                                 static_cast<cudaStream_t>(stream)));
    return OF_OK;
}

static size_t refine_partial_bytes(int batch, int height, int width) {
    return align_up((size_t)batch * lk_tile_blocks_per_pair(height, width) * 2 * sizeof(double));
}

size_t of_lk_refine_workspace_bytes(int batch, int height, int width) {
    if (batch < 1 || height < 1 || width < 1) return 0;
    // per-block partial sums + one plane for the warped current frame (split refinement)
    return refine_partial_bytes(batch, height, width) + align_up((size_t)batch * height * width * sizeof(float));
}

static int refine_dev_impl(const float* prev, const float* curr, float* flow0_u, float* flow0_v, float* flow1_u,
                           float* flow1_v, const int* sel, const int* done, int batch, int height, int width, int window,
                           int mode, int row_lo, int row_hi, int own_lo, int own_hi, double* sums, void* workspace,
                           size_t workspace_bytes, void* stream);

int of_lk_refine_f32_dev(const float* prev, const float* curr, const float* flow_in_u, const float* flow_in_v,
                         float* flow_out_u, float* flow_out_v, int batch, int height, int width, int window, int mode,
                         int row_lo, int row_hi, int own_lo, int own_hi, double* sums, void* workspace,
                         size_t workspace_bytes, void* stream) {
    return refine_dev_impl(prev, curr, const_cast<float*>(flow_in_u), const_cast<float*>(flow_in_v), flow_out_u,
                           flow_out_v, nullptr, nullptr, batch, height, width, window, mode, row_lo, row_hi, own_lo,
                           own_hi, sums, workspace, workspace_bytes, stream);
}

int of_lk_refine_pingpong_f32_dev(const float* prev, const float* curr, float* flow0_u, float* flow0_v, float* flow1_u,
                                  float* flow1_v, const int* sel, const int* done, int batch, int height, int width,
                                  int window, int mode, int row_lo, int row_hi, int own_lo, int own_hi, double* sums,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    if (!sel || !done) return fail(OF_ERR_INVALID_ARGUMENT, "sel / done missing");
    return refine_dev_impl(prev, curr, flow0_u, flow0_v, flow1_u, flow1_v, sel, done, batch, height, width, window, mode,
                           row_lo, row_hi, own_lo, own_hi, sums, workspace, workspace_bytes, stream);
}

int of_lk_convergence_update_dev(const double* sums, int batch, double n_pixels, int* sel, int* done,
                                 int* iters_executed, float* residuals, int max_iterations, int iteration, void* stream) {
    if (!sums || !sel || !done || batch < 1 || n_pixels <= 0) return fail(OF_ERR_INVALID_ARGUMENT, "bad argument");
    OF_TRY(need_device());
    Counter cnt;
    OF_CUDA(launch_convergence_update(sums, batch, n_pixels, sel, done, iters_executed, residuals, max_iterations,
                                      iteration, &cnt.n, static_cast<cudaStream_t>(stream)));
    return OF_OK;
}

static int refine_dev_impl(const float* prev, const float* curr, float* flow_in_u, float* flow_in_v, float* flow_out_u,
                           float* flow_out_v, const int* sel, const int* done, int batch, int height, int width, int window,
                           int mode, int row_lo, int row_hi, int own_lo, int own_hi, double* sums, void* workspace,
                           size_t workspace_bytes, void* stream) {
    OF_TRY(check_frame(prev, curr, height, width));
    OF_TRY(check_frame(flow_in_u, flow_in_v, height, width));
    OF_TRY(check_frame(flow_out_u, flow_out_v, height, width));
    OF_TRY(check_window(window));
    if (mode != OF_MODE_EXACT && mode != OF_MODE_FAST) return fail(OF_ERR_INVALID_ARGUMENT, "unknown mode");
    if (batch < 1 || batch > 65535) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be in 1..65535");
    if (row_lo < 0 || row_hi > height || row_lo >= row_hi || (row_lo & 1))
        return fail(OF_ERR_INVALID_ARGUMENT, "bad row range (row_lo must be even)");
    if (flow_in_u == flow_out_u || flow_in_v == flow_out_v)
        return fail(OF_ERR_INVALID_ARGUMENT, "flow_in and flow_out must be different buffers");
    if (!sums || !workspace || workspace_bytes < of_lk_refine_workspace_bytes(batch, height, width))
        return fail(OF_ERR_INVALID_ARGUMENT, "sums / workspace missing or too small");
    OF_TRY(need_device());
    Counter cnt;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* partial = static_cast<double*>(workspace);
    RefineArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.prev = prev;
    ra.curr = curr;
    ra.flow_u[0] = flow_in_u;
    ra.flow_v[0] = flow_in_v;
    ra.flow_u[1] = flow_out_u;
    ra.flow_v[1] = flow_out_v;
    ra.sel = sel;
    ra.done = done;
    ra.partial = partial;
    ra.H = height;
    ra.W = width;
    ra.window = window;
    ra.row_lo = row_lo;
    ra.row_hi = row_hi;
    ra.own_lo = own_lo;
    ra.own_hi = own_hi;
    int blocks;
    if (mode == OF_MODE_FAST && lk_refine_supported(ra, window)) {
        float* warped = reinterpret_cast<float*>(static_cast<char*>(workspace) + refine_partial_bytes(batch, height, width));
        if (refine_split())
            OF_CUDA(launch_refine_split_form(ra, warped, batch, &cnt.n, st));
        else
            OF_CUDA(launch_lk_refine(ra, batch, &cnt.n, st));
        blocks = lk_refine_units_per_pair(batch, row_hi - row_lo, width);
    } else {
        TileArgs a;
        memset(&a, 0, sizeof(a));
        a.in0 = prev;
        a.in1 = curr;
        a.flow_u[0] = ra.flow_u[0];
        a.flow_v[0] = ra.flow_v[0];
        a.flow_u[1] = flow_out_u;
        a.flow_v[1] = flow_out_v;
        a.sel = sel;
        a.done = done;
        a.partial = partial;
        a.H = height;
        a.W = width;
        a.row_lo = row_lo;
        a.row_hi = row_hi;
        a.own_lo = own_lo;
        a.own_hi = own_hi;
        if (exact_refine_split()) {
            float* warped = reinterpret_cast<float*>(static_cast<char*>(workspace) + refine_partial_bytes(batch, height, width));
            const int halo = window / 2 + 1;
            OF_CUDA(launch_warp_rows(ra, warped, row_lo - halo < 0 ? 0 : row_lo - halo,
                                     row_hi + halo > height ? height : row_hi + halo, true, batch, &cnt.n, st));
            a.in1 = warped;
            OF_CUDA(launch_lk_tile(SRC_WARPED, window, a, batch, &cnt.n, st));
        } else {
            OF_CUDA(launch_lk_tile(SRC_WARP, window, a, batch, &cnt.n, st));
        }
        blocks = lk_tile_blocks_per_pair(row_hi - row_lo, width);
    }
    OF_CUDA(launch_sum_partials(partial, blocks, sums, batch, &cnt.n, st));
    return OF_OK;
}

int of_lk_single_scale_fx_dev(const uint8_t* prev, const uint8_t* curr, int16_t* u, int16_t* v, int batch, int height,
                              int width, int flags, void* stream) {
    OF_TRY(check_frame(prev, curr, height, width));
    OF_TRY(check_frame(u, v, height, width));
    if (batch < 0) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be >= 0");
    if (batch == 0) return OF_OK;
    OF_TRY(need_device());
    Counter cnt;
    const size_t plane = (size_t)height * width;
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
        // the marching kernel (TMA uint8 boxes: width % 16 == 0, 16-byte aligned planes); else the tile kernel
        static const bool force_tile = [] {
            const char* e = getenv("OF_B200_FIXED");
            return e != nullptr && strcmp(e, "tile") == 0;  // A/B runs
        }();
        const int quirk = (flags & OF_FX_MIRROR_AVG_QUIRK) ? 1 : 0;
        if (!force_tile && lk_march_fx_supported(prev + b0 * plane, curr + b0 * plane, u + b0 * plane, v + b0 * plane, height, width))
            OF_CUDA(launch_lk_march_fx(prev + b0 * plane, curr + b0 * plane, u + b0 * plane, v + b0 * plane, nb, height,
                                       width, quirk, &cnt.n, static_cast<cudaStream_t>(stream)));
        else
            OF_CUDA(launch_lk_fixed(prev + b0 * plane, curr + b0 * plane, u + b0 * plane, v + b0 * plane, nb, height, width,
                                    quirk, &cnt.n, static_cast<cudaStream_t>(stream)));
    }
    return OF_OK;
}

int of_lk_single_scale_fx(const uint8_t* prev, const uint8_t* curr, int16_t* u, int16_t* v, int batch, int height,
                          int width, int flags) {
    OF_TRY(check_frame(prev, curr, height, width));
    OF_TRY(check_frame(u, v, height, width));
    if (batch < 0) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be >= 0");
    if (batch == 0) return OF_OK;
    OF_TRY(need_device());
    OF_HOST_PATH(hp);
    const size_t plane = (size_t)height * width;
    // chunks of <= 64 MiB of flow per array, three in flight on three streams: H2D of chunk i+1 and D2H of chunk
    // i-1 overlap the kernel of chunk i (pinned host memory needed for the copies to actually overlap)
    size_t per_chunk = ((size_t)64 << 20) / (plane * sizeof(int16_t));
    if (per_chunk < 1) per_chunk = 1;
    if (per_chunk > (size_t)batch) per_chunk = batch;
    if (per_chunk > 65535) per_chunk = 65535;
    const int n_chunks = (int)(((size_t)batch + per_chunk - 1) / per_chunk);
    const int slots = n_chunks < HOST_STREAMS ? n_chunks : HOST_STREAMS;
    uint8_t* d8[HOST_STREAMS][2];
    int16_t* d16[HOST_STREAMS][2];
    for (int s = 0; s < slots; ++s)
        for (int j = 0; j < 2; ++j) {
            OF_TRY(hp->arena.get(64 + (size_t)s * 4 + j, per_chunk * plane, reinterpret_cast<void**>(&d8[s][j])));
            OF_TRY(hp->arena.get(64 + (size_t)s * 4 + 2 + j, per_chunk * plane * sizeof(int16_t), reinterpret_cast<void**>(&d16[s][j])));
        }
    for (int c = 0; c < n_chunks; ++c) {
        const int s = c % slots;
        cudaStream_t st = hp->streams[s];
        const size_t b0 = (size_t)c * per_chunk;
        const int nb = (int)((size_t)batch - b0 < per_chunk ? (size_t)batch - b0 : per_chunk);
        const size_t n = (size_t)nb * plane;
        OF_CUDA(cudaMemcpyAsync(d8[s][0], prev + b0 * plane, n, cudaMemcpyHostToDevice, st));
        OF_CUDA(cudaMemcpyAsync(d8[s][1], curr + b0 * plane, n, cudaMemcpyHostToDevice, st));
        OF_TRY(of_lk_single_scale_fx_dev(d8[s][0], d8[s][1], d16[s][0], d16[s][1], nb, height, width, flags, st));
        OF_CUDA(cudaMemcpyAsync(u + b0 * plane, d16[s][0], n * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
        OF_CUDA(cudaMemcpyAsync(v + b0 * plane, d16[s][1], n * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < slots; ++s) OF_CUDA(cudaStreamSynchronize(hp->streams[s]));
    return OF_OK;
}

// ---- uint8 ingest (the reference's on-disk frame formats) -------------------------------------
int of_lk_single_scale_u8_dev(const uint8_t* prev, const uint8_t* curr, float* u, float* v, int batch, int height,
                              int width, int window, void* stream) {
    OF_TRY(check_frame(prev, curr, height, width));
    OF_TRY(check_frame(u, v, height, width));
    OF_TRY(check_window(window));
    if (batch < 0) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be >= 0");
    if (batch == 0) return OF_OK;
    OF_TRY(need_device());
    if (!lk_march_u8_supported(prev, curr, u, v, height, width, window))
        return fail(OF_ERR_UNSUPPORTED,
                    "the device-buffer uint8 entry point needs window 5 or 7, width % 16 == 0 and 16-byte aligned planes "
                    "(of_lk_single_scale_u8 takes any frame)");
    Counter cnt;
    OF_CUDA(launch_lk_march_u8(prev, curr, u, v, batch, height, width, window, &cnt.n, static_cast<cudaStream_t>(stream)));
    return OF_OK;
}

int of_lk_single_scale_u8(const uint8_t* prev, const uint8_t* curr, float* u, float* v, int batch, int height, int width,
                          int window, int mode) {
    OF_TRY(check_frame(prev, curr, height, width));
    OF_TRY(check_frame(u, v, height, width));
    OF_TRY(check_window(window));
    if (mode != OF_MODE_EXACT && mode != OF_MODE_FAST) return fail(OF_ERR_INVALID_ARGUMENT, "unknown mode");
    if (batch < 0) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be >= 0");
    if (batch == 0) return OF_OK;
    OF_TRY(need_device());
    OF_HOST_PATH(hp);
    Counter cnt;
    const size_t plane = (size_t)height * width;
    size_t per_chunk = ((size_t)64 << 20) / (plane * sizeof(float));
    if (per_chunk < 1) per_chunk = 1;
    if (per_chunk > (size_t)batch) per_chunk = batch;
    const int n_chunks = (int)((batch + per_chunk - 1) / per_chunk);
    const int slots = n_chunks < 3 ? n_chunks : 3;
    // per slot: uint8 prev / curr, float u / v, and (fallback only) float prev / curr
    uint8_t* d8[3][2];
    float* df[3][4];
    for (int s = 0; s < slots; ++s) {
        for (int j = 0; j < 2; ++j) OF_TRY(hp->arena.get(12 + s * 2 + j, per_chunk * plane, reinterpret_cast<void**>(&d8[s][j])));
        for (int j = 0; j < 2; ++j)
            OF_TRY(hp->arena.get(s * 4 + 2 + j, per_chunk * plane * sizeof(float), reinterpret_cast<void**>(&df[s][2 + j])));
    }
    for (int c = 0; c < n_chunks; ++c) {
        const int s = c % slots;
        cudaStream_t st = hp->streams[s];
        const size_t b0 = (size_t)c * per_chunk;
        const int nb = (int)((size_t)batch - b0 < per_chunk ? (size_t)batch - b0 : per_chunk);
        OF_CUDA(cudaMemcpyAsync(d8[s][0], prev + b0 * plane, (size_t)nb * plane, cudaMemcpyHostToDevice, st));
        OF_CUDA(cudaMemcpyAsync(d8[s][1], curr + b0 * plane, (size_t)nb * plane, cudaMemcpyHostToDevice, st));
        if (mode == OF_MODE_FAST && lk_march_u8_supported(d8[s][0], d8[s][1], df[s][2], df[s][3], height, width, window)) {
            OF_CUDA(launch_lk_march_u8(d8[s][0], d8[s][1], df[s][2], df[s][3], nb, height, width, window, &cnt.n, st));
        } else {
            for (int j = 0; j < 2; ++j) {
                OF_TRY(hp->arena.get(s * 4 + j, per_chunk * plane * sizeof(float), reinterpret_cast<void**>(&df[s][j])));
                OF_CUDA(launch_u8_to_f32(d8[s][j], df[s][j], (size_t)nb * plane, &cnt.n, st));
            }
            OF_TRY(single_scale_dev(df[s][0], df[s][1], df[s][2], df[s][3], nb, height, width, window, mode, st, cnt));
        }
        OF_CUDA(cudaMemcpyAsync(u + b0 * plane, df[s][2], (size_t)nb * plane * sizeof(float), cudaMemcpyDeviceToHost, st));
        OF_CUDA(cudaMemcpyAsync(v + b0 * plane, df[s][3], (size_t)nb * plane * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < slots; ++s) OF_CUDA(cudaStreamSynchronize(hp->streams[s]));
    return OF_OK;
}

/* frame_00.bin / frame_00.mem of generate_test_suite.py:259-271: raw bytes, or one two-digit hex
 * byte per line (the RTL testbenches' $readmemh format).  Chosen by the file extension. */
int of_load_frame_u8(const char* path, uint8_t* out, int height, int width) {
    if (!path || !out || height < 1 || width < 1) return fail(OF_ERR_INVALID_ARGUMENT, "bad argument");
    const size_t n = (size_t)height * width;
    const std::string p(path);
    const bool is_mem = p.size() >= 4 && p.compare(p.size() - 4, 4, ".mem") == 0;
    FILE* f = fopen(path, is_mem ? "r" : "rb");
    if (!f) return fail(OF_ERR_INVALID_ARGUMENT, std::string("cannot open ") + path);
    size_t got = 0;
    if (!is_mem) {
        got = fread(out, 1, n, f);
        unsigned char extra;
        if (got == n && fread(&extra, 1, 1, f) == 1) got = n + 1;  // longer than the frame
    } else {
        char line[256];
        while (fgets(line, sizeof(line), f)) {
            char* q = line;
            while (*q == ' ' || *q == '\t') ++q;
            if (*q == '\0' || *q == '\n' || *q == '\r' || (q[0] == '/' && q[1] == '/')) continue;  // blank / comment
            char* end = nullptr;
            const unsigned long val = strtoul(q, &end, 16);
            if (end == q || val > 255) {
                fclose(f);
                return fail(OF_ERR_INVALID_ARGUMENT, std::string("not a hex byte in ") + path + ": " + line);
            }
            if (got < n) out[got] = (uint8_t)val;
            ++got;
        }
    }
    fclose(f);
    if (got != n)
        return fail(OF_ERR_INVALID_ARGUMENT, std::string(path) + ": holds " + std::to_string(got) + " pixels, expected " +
                                                 std::to_string(n));
    return OF_OK;
}

// ---- flow-field text export ---------------------------------------------------------------------
static int export_flow_common(const char* path, const char* title, int height, int width, int x_min, int x_max, int y_min,
                              int y_max, const float* u, const float* v, const int16_t* u_fx, const int16_t* v_fx) {
    if (!path || height < 1 || width < 1) return fail(OF_ERR_INVALID_ARGUMENT, "bad argument");
    FILE* f = fopen(path, "w");
    if (!f) return fail(OF_ERR_INVALID_ARGUMENT, std::string("cannot open ") + path + " for writing");
    std::vector<char> buf(1 << 20);
    setvbuf(f, buf.data(), _IOFBF, buf.size());
    fprintf(f, "%s\n# Format: x y u v\n# Image size: %dx%d\n", title, width, height);
    if (x_min >= 0) fprintf(f, "# Test region: x[%d:%d], y[%d:%d]\n", x_min, x_max, y_min, y_max);
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            const size_t o = (size_t)y * width + x;
            const double uu = u ? (double)u[o] : (double)u_fx[o] / 128.0;  // S8.7 -> pixels
            const double vv = v ? (double)v[o] : (double)v_fx[o] / 128.0;
            fprintf(f, "%d %d %.6f %.6f\n", x, y, uu, vv);
        }
    const bool bad = ferror(f) != 0;
    if (fclose(f) != 0 || bad) return fail(OF_ERR_INVALID_ARGUMENT, std::string("write error on ") + path);
    return OF_OK;
}

/* export_flow_field_txt (python/lucas_kanade_reference.py:78-103): "# ..." header lines, then one
 * "x y u v" line per pixel, row-major, six decimals -- the file scripts/visualize_flow.py reads. */
int of_export_flow_txt(const char* path, const float* u, const float* v, int height, int width, int x_min, int x_max,
                       int y_min, int y_max) {
    if (!u || !v) return fail(OF_ERR_INVALID_ARGUMENT, "null buffer");
    return export_flow_common(path, "# Optical flow field data (Python reference)", height, width, x_min, x_max, y_min, y_max,
                              u, v, nullptr, nullptr);
}

/* The same file from the fixed-point mode's S8.7 flow, with the header the RTL testbench writes
 * (tb/tb_optical_flow_top.sv:340-358), so a simulator dump can be diffed against it line by line. */
int of_export_flow_fx_txt(const char* path, const int16_t* u, const int16_t* v, int height, int width, int x_min, int x_max,
                          int y_min, int y_max) {
    if (!u || !v) return fail(OF_ERR_INVALID_ARGUMENT, "null buffer");
    return export_flow_common(path, "# Optical flow field data", height, width, x_min, x_max, y_min, y_max, nullptr, nullptr,
                              u, v);
}

// ---- apply_motion (fixture generators) ---------------------------------------------------------
int of_apply_motion_u8_dev(const uint8_t* frames, uint8_t* out, int batch, int height, int width, const double* dx,
                           const double* dy, double cval, void* stream) {
    OF_TRY(check_frame(frames, out, height, width));
    if (!dx || !dy) return fail(OF_ERR_INVALID_ARGUMENT, "null buffer");
    if (frames == out) return fail(OF_ERR_INVALID_ARGUMENT, "in-place shifting is not supported");
    if (batch < 1 || batch > 65535 || height > 65535) return fail(OF_ERR_INVALID_ARGUMENT, "batch and height must be in 1..65535");
    OF_TRY(need_device());
    Counter cnt;
    OF_CUDA(launch_apply_motion(frames, out, dx, dy, batch, height, width, cval, &cnt.n, static_cast<cudaStream_t>(stream)));
    return OF_OK;
}

int of_apply_motion_u8(const uint8_t* frames, uint8_t* out, int batch, int height, int width, const double* dx,
                       const double* dy, double cval) {
    OF_TRY(check_frame(frames, out, height, width));
    if (!dx || !dy) return fail(OF_ERR_INVALID_ARGUMENT, "null buffer");
    if (batch < 1 || batch > 65535 || height > 65535) return fail(OF_ERR_INVALID_ARGUMENT, "batch and height must be in 1..65535");
    OF_TRY(need_device());
    OF_HOST_PATH(hp);
    const size_t n = (size_t)batch * height * width;
    uint8_t *ds, *dd;
    double* dsh;
    OF_TRY(hp->arena.get(0, n, reinterpret_cast<void**>(&ds)));
    OF_TRY(hp->arena.get(1, n, reinterpret_cast<void**>(&dd)));
    OF_TRY(hp->arena.get(2, (size_t)batch * 2 * sizeof(double), reinterpret_cast<void**>(&dsh)));
    cudaStream_t st = hp->streams[0];
    OF_CUDA(cudaMemcpyAsync(ds, frames, n, cudaMemcpyHostToDevice, st));
    OF_CUDA(cudaMemcpyAsync(dsh, dx, (size_t)batch * sizeof(double), cudaMemcpyHostToDevice, st));
    OF_CUDA(cudaMemcpyAsync(dsh + batch, dy, (size_t)batch * sizeof(double), cudaMemcpyHostToDevice, st));
    OF_TRY(of_apply_motion_u8_dev(ds, dd, batch, height, width, dsh, dsh + batch, cval, st));
    OF_CUDA(cudaMemcpyAsync(out, dd, n, cudaMemcpyDeviceToHost, st));
    OF_CUDA(cudaStreamSynchronize(st));
    return OF_OK;
}

// warpAffine's inversion of the forward 2x3 matrix (imgwarp.cpp, !WARP_INVERSE_MAP), same operation
// order; volatile keeps the host compiler from contracting a*b+c into an FMA
static void invert_affine_cv(const double* m, double* out) {
    volatile double D = m[0] * m[4];
    volatile double t = m[1] * m[3];
    D = D - t;
    D = D != 0 ? 1.0 / D : 0.0;
    volatile double a11 = m[4] * D, a22 = m[0] * D;
    volatile double m1 = m[1] * (-D), m3 = m[3] * (-D);
    volatile double p0 = -a11 * m[2], p1 = m1 * m[5];
    volatile double q0 = -m3 * m[2], q1 = a22 * m[5];
    out[0] = a11;
    out[1] = m1;
    out[2] = p0 - p1;
    out[3] = m3;
    out[4] = a22;
    out[5] = q0 - q1;
}

int of_warp_affine_u8_dev(const uint8_t* frames, uint8_t* out, int batch, int height, int width, const double* matrices,
                          int cval, void* stream) {
    OF_TRY(check_frame(frames, out, height, width));
    if (!matrices) return fail(OF_ERR_INVALID_ARGUMENT, "null buffer");
    if (frames == out) return fail(OF_ERR_INVALID_ARGUMENT, "in-place warping is not supported");
    if (batch < 1 || batch > 65535 || height > 65535) return fail(OF_ERR_INVALID_ARGUMENT, "batch and height must be in 1..65535");
    if (cval < 0 || cval > 255) return fail(OF_ERR_INVALID_ARGUMENT, "border value must be in 0..255");
    OF_TRY(need_device());
    std::vector<double> minv((size_t)batch * 6);
    for (int b = 0; b < batch; ++b) invert_affine_cv(matrices + (size_t)b * 6, minv.data() + (size_t)b * 6);
    Counter cnt;
    OF_CUDA(launch_warp_affine(frames, out, minv.data(), batch, height, width, cval, &cnt.n, static_cast<cudaStream_t>(stream)));
    return OF_OK;
}

int of_warp_affine_u8(const uint8_t* frames, uint8_t* out, int batch, int height, int width, const double* matrices,
                      int cval) {
    OF_TRY(check_frame(frames, out, height, width));
    if (!matrices) return fail(OF_ERR_INVALID_ARGUMENT, "null buffer");
    if (batch < 1 || batch > 65535 || height > 65535) return fail(OF_ERR_INVALID_ARGUMENT, "batch and height must be in 1..65535");
    OF_TRY(need_device());
    OF_HOST_PATH(hp);
    const size_t n = (size_t)batch * height * width;
    uint8_t *ds, *dd;
    OF_TRY(hp->arena.get(0, n, reinterpret_cast<void**>(&ds)));
    OF_TRY(hp->arena.get(1, n, reinterpret_cast<void**>(&dd)));
    cudaStream_t st = hp->streams[0];
    OF_CUDA(cudaMemcpyAsync(ds, frames, n, cudaMemcpyHostToDevice, st));
    OF_TRY(of_warp_affine_u8_dev(ds, dd, batch, height, width, matrices, cval, st));
    OF_CUDA(cudaMemcpyAsync(out, dd, n, cudaMemcpyDeviceToHost, st));
    OF_CUDA(cudaStreamSynchronize(st));
    return OF_OK;
}

size_t of_flow_metrics_workspace_bytes(int batch, int height, int width) {
    if (batch < 1 || height < 1 || width < 1) return 0;
    return align_up((size_t)batch * metrics_blocks_per_pair(height, width) * 6 * sizeof(double));
}

int of_flow_metrics_f32_dev(const float* u, const float* v, const float* u_true, const float* v_true, int batch, int height,
                            int width, int y0, int y1, int x0, int x1, double* metrics, void* workspace,
                            size_t workspace_bytes, void* stream) {
    OF_TRY(check_frame(u, v, height, width));
    if (!u_true || !v_true || !metrics) return fail(OF_ERR_INVALID_ARGUMENT, "null buffer");
    if (batch < 1 || batch > 65535) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be in 1..65535");
    if (y0 < 0 || y1 > height || x0 < 0 || x1 > width || y0 >= y1 || x0 >= x1)
        return fail(OF_ERR_INVALID_ARGUMENT, "empty or out-of-frame test region");
    if (!workspace || workspace_bytes < of_flow_metrics_workspace_bytes(batch, y1 - y0, x1 - x0))
        return fail(OF_ERR_INVALID_ARGUMENT, "workspace missing or too small");
    OF_TRY(need_device());
    Counter cnt;
    OF_CUDA(launch_flow_metrics(u, v, u_true, v_true, batch, height, width, y0, y1, x0, x1, static_cast<double*>(workspace),
                                metrics, &cnt.n, static_cast<cudaStream_t>(stream)));
    return OF_OK;
}

int of_flow_metrics_f32(const float* u, const float* v, const float* u_true, const float* v_true, int batch, int height,
                        int width, int y0, int y1, int x0, int x1, double* metrics) {
    OF_TRY(check_frame(u, v, height, width));
    if (!u_true || !v_true || !metrics) return fail(OF_ERR_INVALID_ARGUMENT, "null buffer");
    if (batch < 1 || batch > 65535) return fail(OF_ERR_INVALID_ARGUMENT, "batch must be in 1..65535");
    OF_TRY(need_device());
    OF_HOST_PATH(hp);
    const size_t n = (size_t)batch * height * width * sizeof(float);
    const size_t ws = of_flow_metrics_workspace_bytes(batch, height, width);
    float *du, *dv, *dt;
    double *dm, *dw;
    OF_TRY(hp->arena.get(0, n, reinterpret_cast<void**>(&du)));
    OF_TRY(hp->arena.get(1, n, reinterpret_cast<void**>(&dv)));
    OF_TRY(hp->arena.get(2, (size_t)batch * 2 * sizeof(float), reinterpret_cast<void**>(&dt)));
    OF_TRY(hp->arena.get(3, (size_t)batch * 5 * sizeof(double), reinterpret_cast<void**>(&dm)));
    OF_TRY(hp->arena.get(4, ws, reinterpret_cast<void**>(&dw)));
    cudaStream_t st = hp->streams[0];
    OF_CUDA(cudaMemcpyAsync(du, u, n, cudaMemcpyHostToDevice, st));
    OF_CUDA(cudaMemcpyAsync(dv, v, n, cudaMemcpyHostToDevice, st));
    OF_CUDA(cudaMemcpyAsync(dt, u_true, (size_t)batch * sizeof(float), cudaMemcpyHostToDevice, st));
    OF_CUDA(cudaMemcpyAsync(dt + batch, v_true, (size_t)batch * sizeof(float), cudaMemcpyHostToDevice, st));
    OF_TRY(of_flow_metrics_f32_dev(du, dv, dt, dt + batch, batch, height, width, y0, y1, x0, x1, dm, dw, ws, st));
    OF_CUDA(cudaMemcpyAsync(metrics, dm, (size_t)batch * 5 * sizeof(double), cudaMemcpyDeviceToHost, st));
    OF_CUDA(cudaStreamSynchronize(st));
    return OF_OK;
}

}  // extern "C"

#include "of_rowband.inl"
