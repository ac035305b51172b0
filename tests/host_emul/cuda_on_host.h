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
inline Block*& block() {
    static Block* b = nullptr;
    return b;
}
// dynamic shared memory of the kernel being emulated (set by the harness before launch)
inline void*& dynamic_smem() {
    static void* p = nullptr;
    return p;
}
}  // namespace cuda_on_host
#define OF_DYNAMIC_SMEM(type, name) type* name = static_cast<type*>(cuda_on_host::dynamic_smem())
#define OF_DYNAMIC_SMEM_ALIGNED(align, type, name) type* name = static_cast<type*>(cuda_on_host::dynamic_smem())
// launchers written with OF_LAUNCH (of_common.cuh) run unchanged: grid / block / dynamic shared memory as given
#define OF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cuda_on_host::launch_dynamic(grid, block, smem, [&]() { kernel(__VA_ARGS__); })
#define cudaGetLastError() cudaSuccess

static thread_local uint3 threadIdx;
static thread_local uint3 blockIdx;
static thread_local dim3 blockDim;
static thread_local dim3 gridDim;

static inline void __syncthreads() { pthread_barrier_wait(&cuda_on_host::block()->all); }
static inline double __shfl_down_sync(unsigned, double v, int off) {
    cuda_on_host::Block* b = cuda_on_host::block();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    b->scratch[w][lane] = v;
    pthread_barrier_wait(&b->warp[w]);
    const double r = lane + off < 32 ? b->scratch[w][lane + off] : v;
    pthread_barrier_wait(&b->warp[w]);
    return r;
}
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
static inline void __stcs(float* p, float v) { *p = v; }
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
    };
    std::vector<Arg> args(threads);
    std::vector<pthread_t> tids(threads);
    auto body = [](void* p) -> void* {
        Arg* a = static_cast<Arg*>(p);
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
        args[t] = Arg{t, threads, grid, &kernel};
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
    std::vector<double> buf(smem_bytes / sizeof(double) + 2);  // 16-byte aligned is all a kernel may assume
    void* p = buf.data();
    if (reinterpret_cast<uintptr_t>(p) & 15) p = static_cast<char*>(p) + 8;
    dynamic_smem() = p;
    launch(grid, (int)block.x, kernel);
    dynamic_smem() = nullptr;
}
}  // namespace cuda_on_host
