// Shared helpers for the sm_100a kernels of libnerfstyle_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/nerfstyle_b200.h"

#define NRF_EXPORT extern "C" __attribute__((visibility("default")))

extern thread_local int g_nrf_last_cuda_error;

static inline int nrf_check_launch() {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        g_nrf_last_cuda_error = (int)e;
        cudaGetLastError();
        return NRF_E_CUDA;
    }
    return NRF_OK;
}

static inline uint32_t ceil_div_u32(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

#define NRF_FULL_MASK 0xffffffffu

__device__ __forceinline__ float nrf_clamp(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(NRF_FULL_MASK, v, d);
    return v;
}

// inclusive scans across a warp
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float o = __shfl_up_sync(NRF_FULL_MASK, v, d);
        if (lane >= d) v += o;
    }
    return v;
}
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float o = __shfl_up_sync(NRF_FULL_MASK, v, d);
        if (lane >= d) v *= o;
    }
    return v;
}
__device__ __forceinline__ uint32_t warp_scan_add_u32(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(NRF_FULL_MASK, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// 10-bit-per-axis Morton code and its inverse (/root/reference/raymarching/src/raymarching.cu:56-81)
__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton3D_dev(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
__device__ __forceinline__ uint32_t morton3D_invert_dev(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}
