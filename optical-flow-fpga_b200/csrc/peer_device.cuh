// Device-side pieces shared by peer.cu (stand-alone collective kernels) and lk_march.cu (the same
// collectives fused into the tail of the refinement kernel): system-scope flag stores / loads with a
// time-out, and the "tail" of one refinement iteration -- reduce the per-unit |du|, |dv| partials,
// optionally all-reduce the two sums over the ranks through peer memory, apply the reference's
// convergence test (lucas_kanade_pyramidal.py:213-223) and flip the ping-pong selector.
#pragma once
#include <cuda_runtime.h>

#include "of_common.cuh"
#include "of_kernels.h"

namespace ofb {

#ifndef OF_HOST_EMULATION  // host stand-ins: tests/host_emul/
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

// wait until *flag >= seq; false on time-out or when an earlier wait already failed
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long seq, int* err,
                                          unsigned long long timeout_ns) {
    if (*reinterpret_cast<volatile int*>(err) != 0) return false;
    const unsigned long long t0 = global_timer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(flag) < seq) {
        if ((++spins & 63u) == 0) {
            if (global_timer_ns() - t0 > timeout_ns) {
                atomicExch(err, 1);
                return false;
            }
            __nanosleep(200);
        }
    }
    return true;
}

// One warp (all 32 lanes must call): lane r < world publishes (su, sv) in rank r's exchange slot and
// flag, then collects rank r's contribution from the local arena; returns the rank-ordered totals in
// every lane.  ok = false if a peer did not answer in time.
__device__ __forceinline__ void warp_peer_allreduce(const PeerSync& s, double su, double sv, double& tu, double& tv, bool& ok) {
    const int lane = threadIdx.x & 31;
    const unsigned long long seq = *s.run_id * s.ops_per_run + s.op;
    const int slot = (int)(seq % PEER_SLOTS);
    double cu = 0.0, cv = 0.0;
    int got = 1;
    if (lane < s.world) {
        double* x = reinterpret_cast<double*>(s.peer[lane] + s.xchg_off) + ((size_t)slot * PEER_MAX_WORLD + s.rank) * 2;
        x[0] = su;
        x[1] = sv;
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(s.peer[lane] + s.flag_off) + slot * PEER_MAX_WORLD + s.rank, seq);
        got = wait_flag(reinterpret_cast<const unsigned long long*>(s.peer[s.rank] + s.flag_off) + slot * PEER_MAX_WORLD + lane,
                        seq, s.err, s.timeout_ns)
                  ? 1
                  : 0;
        const double* y = reinterpret_cast<const double*>(s.peer[s.rank] + s.xchg_off) + ((size_t)slot * PEER_MAX_WORLD + lane) * 2;
        cu = got ? *reinterpret_cast<const volatile double*>(y) : 0.0;
        cv = got ? *reinterpret_cast<const volatile double*>(y + 1) : 0.0;
    }
    tu = 0.0;
    tv = 0.0;
    int all = 1;
    for (int q = 0; q < s.world; ++q) {  // rank order: the same sum on every rank
        tu += __shfl_sync(0xffffffffu, cu, q);
        tv += __shfl_sync(0xffffffffu, cv, q);
        all &= __shfl_sync(0xffffffffu, got, q);
    }
    ok = all != 0;
}

// The reference's early-exit test on the level's totals, the trace and the ping-pong flip (one thread).
__device__ __forceinline__ void apply_convergence(const IterTail& t, int pair, double tu, double tv, bool ok) {
    if (!ok) {
        t.done[pair] = 1;  // a peer never answered: stop iterating (the error word is set)
        return;
    }
    const float mu = (float)(tu / t.n_pixels), mv = (float)(tv / t.n_pixels);
    t.sel[pair] ^= 1;
    if (t.iters_executed) t.iters_executed[(size_t)pair * t.iters_pair_stride] += 1;
    if (t.residuals) {
        t.residuals[(size_t)pair * t.resid_pair_stride + 2 * t.iteration + 0] = mu;
        t.residuals[(size_t)pair * t.resid_pair_stride + 2 * t.iteration + 1] = mv;
    }
    if (mu < OF_CONVERGENCE_EPS && mv < OF_CONVERGENCE_EPS) t.done[pair] = 1;
}

// Fused tail of a refinement iteration.  Every unit of a pair stores its partial sums, fences and takes a ticket from
// the pair's counter; the unit that draws the last ticket (every other unit's partial is then visible) calls this with
// ONE full warp: the n_units partials are reduced in a fixed order -- whichever unit it is --, all-reduced over the
// ranks if the level is split, and the reference's convergence test is applied.  Saves a launch per iteration.
__device__ __forceinline__ void warp_iteration_tail(const IterTail& t, const double* part, int n_units, int pair, int lane) {
    __threadfence();
    double su = 0.0, sv = 0.0;
    for (int i = lane; i < n_units; i += 32) {
        su += __ldcg(part + 2 * i);
        sv += __ldcg(part + 2 * i + 1);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        su += __shfl_down_sync(0xffffffffu, su, off);
        sv += __shfl_down_sync(0xffffffffu, sv, off);
    }
    su = __shfl_sync(0xffffffffu, su, 0);
    sv = __shfl_sync(0xffffffffu, sv, 0);
    double tu = su, tv = sv;
    bool ok = true;
    if (t.peers) warp_peer_allreduce(t.sync, su, sv, tu, tv, ok);
    if (lane == 0) {
        t.counter[pair] = 0;  // ready for the next launch
        apply_convergence(t, pair, tu, tv, ok);
    }
}

}  // namespace ofb
