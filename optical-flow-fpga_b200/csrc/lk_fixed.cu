// Fixed-point single-scale LK that mirrors the reference RTL's integer datapath bit for bit:
//   rtl/unopt/gradient_compute.sv:108-140   9-bit average (sign-extension quirk), Sobel >>> 3,
//                                           It = prev - curr
//   rtl/unopt/window_accumulator.sv:128-167 24-bit products, 25-term 32-bit sums
//   rtl/unopt/flow_solver.sv:45,83-148      64-bit products truncated to their low 32 bits,
//                                           |det| > 1000, (num <<< 7) / det truncating, low 16
//                                           bits, clamp to +-1024 (S8.7, +-8 px)
// on geometrically correct 3x3 / 5x5 windows (interior pixels only; the RTL has no border
// handling).  uint8 frames in, int16 S8.7 flow out: 6 B per pixel of traffic.
#include <cuda_runtime.h>
#include <stdint.h>

#include "of_kernels.h"

namespace ofb {

constexpr int FTX = 64, FTY = 16;

__device__ __forceinline__ int avg_rtl(int p, int c, bool quirk) {
    if (quirk) {
        // both operands are `logic signed [7:0]`, the sum is 9 bits wide, `>>` is logical
        const int sp = p >= 128 ? p - 256 : p;
        const int sc = c >= 128 ? c - 256 : c;
        return ((sc + sp) & 0x1FF) >> 1;
    }
    return (p + c) >> 1;
}

__global__ void __launch_bounds__(256) lk_fixed_kernel(const uint8_t* __restrict__ prev, const uint8_t* __restrict__ curr,
                                                        int16_t* __restrict__ u, int16_t* __restrict__ v, int H, int W,
                                                        int quirk) {
    constexpr int GW = FTX + 4, GH = FTY + 4;  // gradient tile
    constexpr int FW = GW + 2, FH = GH + 2;    // pixel tile
    __shared__ int16_t s_avg[FH * FW];
    __shared__ int16_t s_it[FH * FW];
    __shared__ int16_t s_gx[GH * GW], s_gy[GH * GW], s_gt[GH * GW];

    const size_t plane = (size_t)H * W;
    const uint8_t* p = prev + blockIdx.z * plane;
    const uint8_t* c = curr + blockIdx.z * plane;
    const int ox = blockIdx.x * FTX, oy = blockIdx.y * FTY;
    const int tid = threadIdx.x;

    for (int i = tid; i < FH * FW; i += 256) {
        const int y = oy - 3 + i / FW, x = ox - 3 + i % FW;
        int a = 0, t = 0;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const int pv = p[(size_t)y * W + x], cv = c[(size_t)y * W + x];
            a = avg_rtl(pv, cv, quirk != 0);
            t = pv - cv;
        }
        s_avg[i] = (int16_t)a;
        s_it[i] = (int16_t)t;
    }
    __syncthreads();
    for (int i = tid; i < GH * GW; i += 256) {
        const int r = i / GW + 1, cc = i % GW + 1;
        const int16_t* a = s_avg + r * FW + cc;
        const int left = -a[-FW - 1] - 2 * a[-1] - a[FW - 1];
        const int right = a[-FW + 1] + 2 * a[1] + a[FW + 1];
        const int top = -a[-FW - 1] - 2 * a[-FW] - a[-FW + 1];
        const int bottom = a[FW - 1] + 2 * a[FW] + a[FW + 1];
        s_gx[i] = (int16_t)((left + right) >> 3);  // arithmetic shift = floor
        s_gy[i] = (int16_t)((top + bottom) >> 3);
        s_gt[i] = s_it[r * FW + cc];
    }
    __syncthreads();
    for (int o = tid; o < FTY * FTX; o += 256) {
        const int r = o / FTX, cc = o % FTX;
        const int y = oy + r, x = ox + cc;
        if (y >= H || x >= W) continue;
        int16_t ou = 0, ov = 0;
        // gradients exist for 1..H-2 / 1..W-2, so 5x5 windows for 3..H-4 / 3..W-4
        if (y >= 3 && y < H - 3 && x >= 3 && x < W - 3) {
            int sxx = 0, syy = 0, sxy = 0, sxt = 0, syt = 0;
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int k = (r + i) * GW + cc + j;
                    const int gx = s_gx[k], gy = s_gy[k], gt = s_gt[k];
                    sxx += gx * gx;
                    syy += gy * gy;
                    sxy += gx * gy;
                    sxt += gx * gt;
                    syt += gy * gt;
                }
            // 32 x 32 -> low 32 bits, 32-bit wrapping subtraction
            const uint32_t det_u = (uint32_t)sxx * (uint32_t)syy - (uint32_t)sxy * (uint32_t)sxy;
            const uint32_t nu_u = (uint32_t)syy * (uint32_t)sxt - (uint32_t)sxy * (uint32_t)syt;
            const uint32_t nv_u = (uint32_t)sxx * (uint32_t)syt - (uint32_t)sxy * (uint32_t)sxt;
            const int det = (int)det_u, nu = (int)nu_u, nv = (int)nv_u;
            if (det > 1000 || det < -1000) {
                const long long qu = ((long long)nu * 128) / (long long)det;  // truncates toward zero
                const long long qv = ((long long)nv * 128) / (long long)det;
                int fu = (int16_t)(uint16_t)(qu & 0xFFFF);
                int fv = (int16_t)(uint16_t)(qv & 0xFFFF);
                fu = fu > 1024 ? 1024 : (fu < -1024 ? -1024 : fu);
                fv = fv > 1024 ? 1024 : (fv < -1024 ? -1024 : fv);
                ou = (int16_t)fu;
                ov = (int16_t)fv;
            }
        }
        u[blockIdx.z * plane + (size_t)y * W + x] = ou;
        v[blockIdx.z * plane + (size_t)y * W + x] = ov;
    }
}

cudaError_t launch_lk_fixed(const uint8_t* prev, const uint8_t* curr, int16_t* u, int16_t* v, int batch, int H, int W,
                            int mirror_avg_quirk, int* launches, cudaStream_t stream) {
    if (batch > 65535) return cudaErrorInvalidValue;
    if (launches) *launches += 1;
    dim3 grid((W + FTX - 1) / FTX, (H + FTY - 1) / FTY, batch);
    OF_LAUNCH(lk_fixed_kernel, grid, 256, 0, stream, prev, curr, u, v, H, W, mirror_avg_quirk);
    return cudaGetLastError();
}

}  // namespace ofb
