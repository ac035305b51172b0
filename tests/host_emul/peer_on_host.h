// Host stand-ins for what csrc/peer_device.cuh takes from the device: system-scope flag accesses and the global
// nanosecond timer.  TEST INFRASTRUCTURE (tests/host_emul/): included before any kernel source that includes
// peer_device.cuh.
#pragma once
#include <time.h>

namespace ofb {

static inline void st_release_sys(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline unsigned long long ld_acquire_sys(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline unsigned long long global_timer_ns() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (unsigned long long)t.tv_sec * 1000000000ull + (unsigned long long)t.tv_nsec;
}

}  // namespace ofb
