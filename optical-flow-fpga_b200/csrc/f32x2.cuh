// Packed float32 pairs (Blackwell FADD2 / FMUL2 / FFMA2: one issue slot, two values) shared by the marching
// kernels (lk_march.cu) and the exact tile kernel (lk_tile5.cu).  Under OF_HOST_EMULATION (tests/host_emul/: the
// kernels' sources compiled by g++ and run on the CPU) every packed operation is two IEEE float32 operations.
#pragma once
#ifndef OF_HOST_EMULATION
#include <cuda_runtime.h>

namespace ofb {

// ---- packed FP32 pairs (Blackwell FADD2 / FMUL2 / FFMA2): one issue slot, two columns -----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

}  // namespace ofb
#else
#include <cmath>
#include <cstring>

namespace ofb {

typedef unsigned long long f32x2;
static inline f32x2 pk(float lo, float hi) {
    unsigned a, b;
    std::memcpy(&a, &lo, 4);
    std::memcpy(&b, &hi, 4);
    return ((f32x2)b << 32) | a;
}
static inline void unpk(f32x2 v, float& lo, float& hi) {
    const unsigned a = (unsigned)v, b = (unsigned)(v >> 32);
    std::memcpy(&lo, &a, 4);
    std::memcpy(&hi, &b, 4);
}
#define OF_PAIR_OP(name, expr_lo, expr_hi)                   \
    static inline f32x2 name(f32x2 a, f32x2 b) {              \
        float al, ah, bl, bh;                                 \
        unpk(a, al, ah);                                      \
        unpk(b, bl, bh);                                      \
        volatile float rl = expr_lo, rh = expr_hi;            \
        return pk(rl, rh);                                    \
    }
OF_PAIR_OP(add2, al + bl, ah + bh)
OF_PAIR_OP(sub2, al - bl, ah - bh)
OF_PAIR_OP(mul2, al* bl, ah* bh)
static inline f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    float al, ah, bl, bh, cl, ch;
    unpk(a, al, ah);
    unpk(b, bl, bh);
    unpk(c, cl, ch);
    return pk(std::fmaf(al, bl, cl), std::fmaf(ah, bh, ch));
}

}  // namespace ofb
#endif
