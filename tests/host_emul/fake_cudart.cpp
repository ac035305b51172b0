// The handful of CUDA runtime calls the library's host code makes, for the whole-library host emulation
// (tests/host_emul/build_emulated_library.py): "device" memory is host memory, streams are synchronous, there is one
// "device".  TEST INFRASTRUCTURE ONLY -- it exists so that the C ABI's own driver code (buffer management, level loop,
// kernel selection) can be exercised by `pytest -m "not gpu"`; nothing in the product links or loads it.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime error"; }
cudaError_t cudaMalloc(void** p, size_t n) {
    *p = nullptr;
    return posix_memalign(p, 256, n ? n : 1) == 0 ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
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
