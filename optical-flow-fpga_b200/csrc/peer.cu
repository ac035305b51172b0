// Collectives of the row-band (multi-GPU) pyramidal path over NVLink / NVSwitch peer memory.
//
// Every rank owns one "arena" (a single cudaMalloc) with the same layout; peers map each other's
// arenas (CUDA IPC between processes, plain pointers between threads of one process), so a rank
// addresses any peer's copy of a plane as peer_base + offset.  Three device-side primitives, all
// stream-ordered, none of them touching the host or NCCL:
//
//   push_rows        all-gather by stores: the rows a rank owns go straight into every peer's
//                    arena (and its own "gathered" plane), read once, written `world` times;
//   peer_sync        release-store of a sequence number into every peer's flag slot, then an
//                    acquire-spin until every rank's number has arrived (the rows pushed before are
//                    then visible): the barrier that ends an all-gather;
//   allreduce_update the reference's global early-exit test (lucas_kanade_pyramidal.py:213-223)
//                    for row bands: reduce this rank's per-block |du|, |dv| partials, publish the
//                    two sums in every peer's exchange slot, wait for all ranks, add them in rank
//                    order (so every rank takes the same decision), flip the ping-pong selector.
//
// Sequence numbers only grow and a slot is reused after PEER_SLOTS collectives, while no rank can
// run more than one collective ahead of the slowest one, so flags never need a reset.  Every spin
// has a time-out (a peer that died must not hang the GPU): it sets an error word that makes
// all later waits of the run fall through; the host reads it after the run (of_rowband_trace /
// of_rowband_status return OF_ERR_PEER_TIMEOUT) and the next run after that read starts with a clean word.
#include <cuda_runtime.h>
#include <string.h>

#include "of_common.cuh"
#include "of_kernels.h"
#include "peer_device.cuh"

namespace ofb {

namespace {

struct PushArgs {
    const float* src[2];  // ping-pong candidates (same plane offsets), current = sel ? src[1] : src[0]
    const int* sel;       // nullable
    int sel_xor;
    size_t dst_off;       // byte offset of the gathered plane in every arena
    char* peer[PEER_MAX_WORLD];
    int world;
    size_t first;         // first element (row_a * W) and number of elements (rows * W) this rank owns
    size_t count;
    // what each rank receives: elements [lo[r], hi[r]) of the plane, clipped to the owned range by the launcher
    // (the whole owned range for an all-gather, the rows a neighbour's next stage reads for a halo push, empty
    // for a rank that already holds the data)
    size_t lo[PEER_MAX_WORLD], hi[PEER_MAX_WORLD];
};

__global__ void __launch_bounds__(256) push_rows_kernel(const PushArgs a) {
    const int cur = (a.sel ? a.sel[0] : 0) ^ a.sel_xor;
    const float* __restrict__ src = (cur ? a.src[1] : a.src[0]) + a.first;
    const size_t stride = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | (a.dst_off + a.first * 4)) & 15) == 0;
    const size_t n4 = vec ? a.count / 4 : 0;
    for (size_t i = tid; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
        const size_t e = a.first + 4 * i;  // first of the four elements
#pragma unroll
        for (int r = 0; r < PEER_MAX_WORLD; ++r) {
            if (r >= a.world || e + 4 <= a.lo[r] || e >= a.hi[r]) continue;
            float* d = reinterpret_cast<float*>(a.peer[r] + a.dst_off) + e;
            if (e >= a.lo[r] && e + 4 <= a.hi[r]) {
                *reinterpret_cast<float4*>(d) = v;
            } else {  // a range that starts or ends inside this word
                const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (e + j >= a.lo[r] && e + j < a.hi[r]) d[j] = x[j];
            }
        }
    }
    for (size_t i = n4 * 4 + tid; i < a.count; i += stride) {
        const float v = __ldg(src + i);
        const size_t e = a.first + i;
#pragma unroll
        for (int r = 0; r < PEER_MAX_WORLD; ++r)
            if (r < a.world && e >= a.lo[r] && e < a.hi[r]) reinterpret_cast<float*>(a.peer[r] + a.dst_off)[e] = v;
    }
}

__global__ void __launch_bounds__(32) peer_sync_kernel(const PeerSync a) {
    const int r = threadIdx.x;
    const unsigned long long seq = *a.run_id * a.ops_per_run + a.op;
    const int slot = (int)(seq % PEER_SLOTS);
    if (r < a.world) {
        __threadfence_system();  // the rows pushed by the kernels before this one
        st_release_sys(reinterpret_cast<unsigned long long*>(a.peer[r] + a.flag_off) + slot * PEER_MAX_WORLD + a.rank, seq);
        wait_flag(reinterpret_cast<const unsigned long long*>(a.peer[a.rank] + a.flag_off) + slot * PEER_MAX_WORLD + r, seq,
                  a.err, a.timeout_ns);
    }
}

struct AllreduceArgs {
    IterTail t;             // sel / done / trace of the level (pair 0), t.sync = the collective
    const double* partial;  // this rank's per-block sums of |du|, |dv| (nullptr: the rank owns no rows)
    int blocks;
};

// stand-alone form of the iteration tail (exact mode's tile kernel, ranks without rows); the fast
// path's marching kernel runs the same code in its last warp (lk_march.cu)
__global__ void __launch_bounds__(256) peer_allreduce_update_kernel(const AllreduceArgs a) {
    __shared__ double red[2][8];
    if (a.t.done[0]) return;  // the level has converged: the ranks agree on that, nothing is exchanged
    double su = 0.0, sv = 0.0;
    if (a.partial != nullptr) {
        for (int i = threadIdx.x; i < a.blocks; i += 256) {
            su += a.partial[2 * i];
            sv += a.partial[2 * i + 1];
        }
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
    if (threadIdx.x < 32) {
        su = 0.0;
        sv = 0.0;
        for (int w = 0; w < 8; ++w) {
            su += red[0][w];
            sv += red[1][w];
        }
        double tu, tv;
        bool ok;
        warp_peer_allreduce(a.t.sync, su, sv, tu, tv, ok);
        if (threadIdx.x == 0) apply_convergence(a.t, 0, tu, tv, ok);
    }
}

}  // namespace

// first kernel of a run: next run number; clear_err: the host has read the error word of an earlier run
// (of_rowband_trace / of_rowband_status), so this run may wait for its peers again
__global__ void peer_bump_run_kernel(unsigned long long* run_id, int* err, int clear_err) {
    *run_id += 1;
    if (clear_err) *err = 0;
}

void fill_peer_sync(PeerSync& s, const PeerView& pv, unsigned long long op) {
    for (int r = 0; r < PEER_MAX_WORLD; ++r) s.peer[r] = r < pv.world ? pv.peer[r] : nullptr;
    s.world = pv.world;
    s.rank = pv.rank;
    s.flag_off = pv.flag_off;
    s.xchg_off = pv.xchg_off;
    s.run_id = pv.run_id;
    s.ops_per_run = pv.ops_per_run;
    s.op = op;
    s.err = pv.err;
    s.timeout_ns = pv.timeout_ns;
}

cudaError_t launch_peer_push_ranges(const PeerView& pv, const float* src0, const float* src1, const int* sel, int sel_xor,
                                    size_t dst_off, size_t first, size_t count, const size_t* lo, const size_t* hi,
                                    int* launches, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    PushArgs a;
    a.src[0] = src0;
    a.src[1] = src1 ? src1 : src0;
    a.sel = sel;
    a.sel_xor = sel_xor;
    a.dst_off = dst_off;
    a.world = pv.world;
    a.first = first;
    a.count = count;
    size_t lo_min = first + count, hi_max = first;
    for (int r = 0; r < PEER_MAX_WORLD; ++r) {
        a.peer[r] = r < pv.world ? pv.peer[r] : nullptr;
        size_t l = r < pv.world ? lo[r] : 0, h = r < pv.world ? hi[r] : 0;
        if (l < first) l = first;
        if (h > first + count) h = first + count;
        if (h <= l) l = h = 0;
        a.lo[r] = l;
        a.hi[r] = h;
        if (h > l) {
            if (l < lo_min) lo_min = l;
            if (h > hi_max) hi_max = h;
        }
    }
    if (hi_max <= lo_min) return cudaSuccess;  // nobody needs anything of this rank's rows
    // only the span somebody receives is read
    a.first = lo_min;
    a.count = hi_max - lo_min;
    size_t blocks = (a.count / 4 + 256 * 4 - 1) / (256 * 4);
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (launches) *launches += 1;
    OF_LAUNCH(push_rows_kernel, (unsigned)blocks, 256, 0, stream, a);
    return cudaGetLastError();
}

cudaError_t launch_peer_push_rows(const PeerView& pv, const float* src0, const float* src1, const int* sel, int sel_xor,
                                  size_t dst_off, size_t first, size_t count, bool skip_self, int* launches,
                                  cudaStream_t stream) {
    size_t lo[PEER_MAX_WORLD], hi[PEER_MAX_WORLD];
    for (int r = 0; r < PEER_MAX_WORLD; ++r) {
        const bool to = r < pv.world && !(skip_self && r == pv.rank);
        lo[r] = to ? first : 0;
        hi[r] = to ? first + count : 0;
    }
    return launch_peer_push_ranges(pv, src0, src1, sel, sel_xor, dst_off, first, count, lo, hi, launches, stream);
}

cudaError_t launch_peer_begin_run(const PeerView& pv, bool clear_error, int* launches, cudaStream_t stream) {
    if (launches) *launches += 1;
    OF_LAUNCH(peer_bump_run_kernel, 1, 1, 0, stream, const_cast<unsigned long long*>(pv.run_id), pv.err, clear_error ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_peer_sync(const PeerView& pv, unsigned long long op, int* launches, cudaStream_t stream) {
    PeerSync s;
    fill_peer_sync(s, pv, op);
    if (launches) *launches += 1;
    OF_LAUNCH(peer_sync_kernel, 1, 32, 0, stream, s);
    return cudaGetLastError();
}

cudaError_t launch_peer_allreduce_update(const PeerView& pv, unsigned long long op, const double* partial, int blocks,
                                         double n_pixels, int* sel, int* done, int* iters_executed, float* residuals,
                                         int iteration, int* launches, cudaStream_t stream) {
    AllreduceArgs a;
    memset(&a, 0, sizeof(a));
    fill_peer_sync(a.t.sync, pv, op);
    a.t.peers = 1;
    a.t.n_pixels = n_pixels;
    a.t.sel = sel;
    a.t.done = done;
    a.t.iters_executed = iters_executed;
    a.t.iters_pair_stride = 0;
    a.t.residuals = residuals;
    a.t.resid_pair_stride = 0;
    a.t.iteration = iteration;
    a.partial = partial;
    a.blocks = blocks;
    if (launches) *launches += 1;
    OF_LAUNCH(peer_allreduce_update_kernel, 1, 256, 0, stream, a);
    return cudaGetLastError();
}

}  // namespace ofb
