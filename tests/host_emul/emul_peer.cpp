// peer.cu compiled for the CPU (the row-band collectives; only linked, several ranks at once are not emulated).
#include "cuda_on_host.h"
namespace ofb {
static inline void st_release_sys(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline unsigned long long ld_acquire_sys(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline unsigned long long global_timer_ns() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (unsigned long long)t.tv_sec * 1000000000ull + (unsigned long long)t.tv_nsec;
}
}  // namespace ofb
#include "peer.cu"
