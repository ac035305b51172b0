// The handful of CUDA runtime calls the library's host code makes, for the whole-library host emulation
// (tests/host_emul/build_emulated_library.py): "device" memory is host memory, streams are synchronous, there is one
// "device".  TEST INFRASTRUCTURE ONLY -- it exists so that the C ABI's own driver code (buffer management, level loop,
// kernel selection) can be exercised by `pytest -m "not gpu"`; nothing in the product links or loads it.
#include <cuda_runtime.h>

#include <sys/mman.h>

#include <cstdlib>
#include <cstring>

extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime error"; }
// Several ranks = several PROCESSES (forked after cuda_on_host_shared_init): "device" allocations then come from one
// shared mapping that every process sees at the same address, each process bump-allocating from its own slice, so the
// arena pointers the ranks exchange are valid everywhere -- the stand-in for peer-mapped device memory.
static char* g_shared = nullptr;
static size_t g_shared_bytes = 0, g_lo = 0, g_hi = 0, g_next = 0;
int cuda_on_host_shared_init(size_t bytes) {
    void* m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) return 1;
    g_shared = static_cast<char*>(m);
    g_shared_bytes = bytes;
    return 0;
}
int cuda_on_host_use_slice(int index, int count) {
    if (!g_shared || count < 1 || index < 0 || index >= count) return 1;
    const size_t slice = (g_shared_bytes / count) & ~(size_t)4095;
    g_lo = g_next = slice * index;
    g_hi = g_lo + slice;
    return 0;
}
cudaError_t cudaMalloc(void** p, size_t n) {
    *p = nullptr;
    if (g_hi > g_lo) {
        const size_t at = (g_next + 255) & ~(size_t)255;
        if (at + n > g_hi) return cudaErrorMemoryAllocation;
        *p = g_shared + at;
        g_next = at + (n ? n : 1);
        return cudaSuccess;
    }
    return posix_memalign(p, 256, n ? n : 1) == 0 ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void* p) {
    if (g_shared && static_cast<char*>(p) >= g_shared && static_cast<char*>(p) < g_shared + g_shared_bytes) return cudaSuccess;
    free(p);
    return cudaSuccess;
}
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(0x1); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
// peer memory needs several processes / devices: not emulated
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }
}
