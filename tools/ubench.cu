// Micro-benchmarks used to size the fast LK kernel on B200: issue throughput of scalar vs
// packed (f32x2) FP32 adds/FMAs, warp shuffles and IEEE division.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
#include <cuda_runtime.h>
#include <cstdio>

#define ITERS 4096
#define CHAINS 8

__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int MODE>
__global__ void k(float* out, float seed) {
    float a[CHAINS];
    unsigned long long p[CHAINS];
    double dd[CHAINS];
    for (int i = 0; i < CHAINS; ++i) dd[i] = seed + threadIdx.x + i;
    for (int i = 0; i < CHAINS; ++i) {
        a[i] = seed + threadIdx.x + i;
        float2 t = make_float2(a[i], a[i] + 0.5f);
        p[i] = *reinterpret_cast<unsigned long long*>(&t);
    }
    float2 inc2 = make_float2(seed, seed * 0.5f);
    unsigned long long inc = *reinterpret_cast<unsigned long long*>(&inc2);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) a[i] = __fadd_rn(a[i], seed);
            if (MODE == 1) p[i] = add2(p[i], inc);
            if (MODE == 2) a[i] = __fmaf_rn(a[i], seed, seed);
            if (MODE == 3) p[i] = fma2(p[i], inc, inc);
            if (MODE == 4) a[i] = __shfl_down_sync(0xffffffffu, a[i], 1);
            if (MODE == 5) a[i] = __fdiv_rn(seed, a[i]);
            if (MODE == 6) a[i] = __fmul_rn(a[i], seed);
            if (MODE == 7) dd[i] = __dadd_rn(dd[i], (double)seed);
            if (MODE == 8) dd[i] = __dmul_rn(dd[i], (double)seed);
            if (MODE == 9) dd[i] = __fma_rn(dd[i], (double)seed, (double)seed);
            if (MODE == 10) dd[i] = (double)(float)dd[i] + 1.0;
        }
    }
    float s = 0;
    for (int i = 0; i < CHAINS; ++i) {
        float2 t = *reinterpret_cast<float2*>(&p[i]);
        s += a[i] + t.x + t.y + (float)dd[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int flops_per_op) {
    float* out;
    const int blocks = 148 * 8, threads = 256;
    cudaMalloc(&out, blocks * threads * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * ITERS * CHAINS;  // lane-instructions
    printf("%-10s %8.3f ms  %8.2f T lane-instr/s  %8.2f T elem-ops/s\n", name, ms, ops / ms / 1e9,
           ops * flops_per_op / ms / 1e9);
    cudaFree(out);
}

int main() {
    run<0>("FADD", 1);
    run<1>("FADD2", 2);
    run<2>("FFMA", 1);
    run<3>("FFMA2", 2);
    run<6>("FMUL", 1);
    run<4>("SHFL", 1);
    run<5>("FDIV.rn", 1);
    run<7>("DADD", 1);
    run<8>("DMUL", 1);
    run<9>("DFMA", 1);
    run<10>("F2F+DADD", 1);
    return 0;
}
