// Just enough of the CUDA execution model to run a kernel's SOURCE on the CPU: every CUDA thread of a block is
// an OS thread, __syncthreads() is a pthread barrier, __shared__ arrays are function-local statics, warp shuffles
// exchange through a per-warp scratch line.  Blocks run one after the other.  TEST INFRASTRUCTURE ONLY: it lets
// `pytest -m "not gpu"` check a kernel's indexing and operation order against the oracle without a GPU.
//
// Arithmetic intrinsics (__fadd_rn ...) are single IEEE operations here too (volatile keeps g++ from fusing or
// reassociating; build with -ffp-contract=off), so a kernel that is bit-exact on the host is bit-exact on the
// device as long as it only uses the operations listed below.
#pragma once
#include <cuda_runtime.h>  // vector types (float2, float4, dim3, uint3), host_defines.h
#include <pthread.h>
#include <sched.h>

#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>

// host_defines.h makes the CUDA qualifiers vanish without nvcc; __shared__ must become storage that all threads
// of the block see
#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)

namespace cuda_on_host {
constexpr int MAX_THREADS = 1024;
struct Block {
    pthread_barrier_t all;
    pthread_barrier_t warp[MAX_THREADS / 32];
    double scratch[MAX_THREADS / 32][32];
};
// Per OS thread: the launching thread sets both, every worker thread of the launch copies them, so several host
// threads (the ranks of a row-band run) can have kernels in flight at the same time.
inline Block*& block() {
    static thread_local Block* b = nullptr;
    return b;
}
// dynamic shared memory of the kernel being emulated (set by the harness before launch)
inline void*& dynamic_smem() {
    static thread_local void* p = nullptr;
    return p;
}
}  // namespace cuda_on_host
#define OF_DYNAMIC_SMEM(type, name) type* name = static_cast<type*>(cuda_on_host::dynamic_smem())
#define OF_DYNAMIC_SMEM_ALIGNED(align, type, name) type* name = static_cast<type*>(cuda_on_host::dynamic_smem())
// launchers written with OF_LAUNCH (of_common.cuh) run unchanged: grid / block / dynamic shared memory as given
#define OF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cuda_on_host::launch_dynamic(grid, block, smem, [&]() { kernel(__VA_ARGS__); })
#ifndef CUDA_ON_HOST_FAKE_RUNTIME  // the whole-library build links tests/host_emul/fake_cudart.cpp instead
#define cudaGetLastError() cudaSuccess
#endif

static thread_local uint3 threadIdx;
static thread_local uint3 blockIdx;
static thread_local dim3 blockDim;
static thread_local dim3 gridDim;

static inline void __syncthreads() { pthread_barrier_wait(&cuda_on_host::block()->all); }
namespace cuda_on_host {
// one warp-wide exchange: every lane publishes `v`, then reads lane `src` (its own value when src is out of range)
template <typename T>
static inline T warp_exchange(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffles move at most 8 bytes");
    Block* b = block();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    std::memcpy(&b->scratch[w][lane], &v, sizeof(T));
    pthread_barrier_wait(&b->warp[w]);
    T r = v;
    if (src >= 0 && src < 32) std::memcpy(&r, &b->scratch[w][src], sizeof(T));
    pthread_barrier_wait(&b->warp[w]);
    return r;
}
}  // namespace cuda_on_host
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src) { return cuda_on_host::warp_exchange(v, src & 31); }
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int off) { return cuda_on_host::warp_exchange(v, (int)(threadIdx.x & 31) + off); }
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int off) { return cuda_on_host::warp_exchange(v, (int)(threadIdx.x & 31) - off); }
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&cuda_on_host::block()->warp[threadIdx.x >> 5]); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fadd_rd(float a, float b) {  // round toward -infinity (build with -frounding-math)
    const int mode = fegetround();
    fesetround(FE_DOWNWARD);
    volatile float x = a, y = b;
    volatile float r = x + y;
    fesetround(mode);
    return r;
}
static inline int __float_as_int(float f) {
    int i;
    std::memcpy(&i, &f, 4);
    return i;
}
template <typename T>
static inline void __stcs(T* p, T v) { *p = v; }
template <typename T>
static inline T __ldcg(const T* p) { return *p; }
static inline float __int_as_float(int i) {
    float f;
    std::memcpy(&f, &i, 4);
    return f;
}
static inline float __uint_as_float(unsigned i) {
    float f;
    std::memcpy(&f, &i, 4);
    return f;
}
static inline unsigned __float_as_uint(float f) {
    unsigned i;
    std::memcpy(&i, &f, 4);
    return i;
}
static inline int __float2int_rd(float f) { return (int)std::floor(f); }
static inline int __double2int_rz(double d) { return (int)d; }
static inline float __fmaf_rd(float a, float b, float c) {  // exact product and sum in float64, one rounding down
    const int mode = fegetround();
    fesetround(FE_DOWNWARD);
    volatile double t = (double)a * (double)b + (double)c;  // callers keep this exact in float64 (see floor_div8)
    volatile float r = (float)t;
    fesetround(mode);
    return r;
}
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {  // PRMT, default mode
    const unsigned long long src = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(src >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline void __nanosleep(unsigned) { sched_yield(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicInc(unsigned* p, unsigned limit) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST), next;
    do next = old >= limit ? 0 : old + 1;
    while (!__atomic_compare_exchange_n(p, &old, next, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return old;
}
#define __grid_constant__
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline bool __all_sync(unsigned, bool pred) {
    cuda_on_host::Block* b = cuda_on_host::block();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    b->scratch[w][lane] = pred ? 1.0 : 0.0;
    pthread_barrier_wait(&b->warp[w]);
    bool all = true;
    for (int l = 0; l < 32; ++l) all = all && b->scratch[w][l] != 0.0;
    pthread_barrier_wait(&b->warp[w]);
    return all;
}
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }

namespace cuda_on_host {
// kernel(args) for every thread of every block of `grid`, `threads` threads per block (multiple of 32)
template <typename Kernel>
void launch(dim3 grid, int threads, Kernel kernel) {
    Block blk;
    block() = &blk;
    pthread_barrier_init(&blk.all, nullptr, threads);
    for (int w = 0; w < threads / 32; ++w) pthread_barrier_init(&blk.warp[w], nullptr, 32);
    struct Arg {
        int tid, threads;
        dim3 grid;
        Kernel* k;
        Block* blk;
        void* smem;
    };
    std::vector<Arg> args(threads);
    std::vector<pthread_t> tids(threads);
    auto body = [](void* p) -> void* {
        Arg* a = static_cast<Arg*>(p);
        block() = a->blk;
        dynamic_smem() = a->smem;
        blockDim = dim3(a->threads, 1, 1);
        gridDim = a->grid;
        threadIdx = uint3{(unsigned)a->tid, 0, 0};
        for (unsigned z = 0; z < a->grid.z; ++z)
            for (unsigned y = 0; y < a->grid.y; ++y)
                for (unsigned x = 0; x < a->grid.x; ++x) {
                    blockIdx = uint3{x, y, z};
                    (*a->k)();
                    pthread_barrier_wait(&block()->all);  // the next block reuses the static "shared memory"
                }
        return nullptr;
    };
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    for (int t = 0; t < threads; ++t) {
        args[t] = Arg{t, threads, grid, &kernel, &blk, dynamic_smem()};
        pthread_create(&tids[t], &attr, body, &args[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tids[t], nullptr);
    pthread_attr_destroy(&attr);
    pthread_barrier_destroy(&blk.all);
    for (int w = 0; w < threads / 32; ++w) pthread_barrier_destroy(&blk.warp[w]);
    block() = nullptr;
}
template <typename Kernel>
void launch_dynamic(dim3 grid, dim3 block, size_t smem_bytes, Kernel kernel) {
#ifdef OF_EMUL_EXACT_SMEM
    // exactly the bytes the launch asked for, so that AddressSanitizer (tests/host_emul/asan_check.sh) sees any
    // access past the kernel's dynamic shared memory
    void* buf = nullptr;
    if (posix_memalign(&buf, 128, smem_bytes ? smem_bytes : 128) != 0) abort();
    dynamic_smem() = buf;
    launch(grid, (int)block.x, kernel);
    dynamic_smem() = nullptr;
    free(buf);
#else
    std::vector<char> buf(smem_bytes + 256);
    dynamic_smem() = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(buf.data()) + 127) & ~(uintptr_t)127);
    launch(grid, (int)block.x, kernel);
    dynamic_smem() = nullptr;
#endif
}
}  // namespace cuda_on_host
