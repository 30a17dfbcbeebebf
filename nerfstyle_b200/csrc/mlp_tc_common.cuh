// Row-owner / issuer building blocks shared by the tcgen05 MLP kernels (mlp_tc.cu: one network per launch;
// field_tc.cu: the four networks of the field in one launch).  See mlp_tc.cu for the design notes.
#pragma once
#include "common.cuh"
#include "tc05.cuh"

namespace {


// chunk stride of 128-row activation tiles: 2048 B of rows + a pad that makes the warp-cooperative (coalesced) staging of
// the input rows bank-conflict free (chunk c of row r lands in bank 4 (r + c * pad / 16) mod 32: the pad spreads the
// in/8 chunks of one row over distinct banks; row-owner accesses -- 512 contiguous bytes per chunk -- never conflict)
__host__ __device__ constexpr uint32_t ch_for(int in_kt) { return in_kt == 1 ? 2112u : in_kt == 2 ? 2080u : 2064u; }
constexpr uint32_t CHW = 1024;   // chunk stride of 64-row weight tiles (W1, Wh)
constexpr uint32_t CHO = 256;    // chunk stride of the 16-row output weight tile
constexpr int TC_ROWS = 128;     // row-owner threads (warps 0-3)
constexpr int TC_THREADS = 160;  // + the issuer warp

__device__ __forceinline__ float tact_fwd(float z, int act) {
    switch (act) {
        case NRF_ACT_RELU: return fmaxf(z, 0.0f);
        case NRF_ACT_SIGMOID: return __frcp_rn(1.0f + __expf(-z));
        case NRF_ACT_EXP: return __expf(z);
        case NRF_ACT_TRUNC_EXP: return __expf(__half2float(__float2half_rn(z)));      // trunc_exp of the f16 network output
        default: return z;
    }
}
__device__ __forceinline__ float tact_bwd(float z, int act) {
    switch (act) {
        case NRF_ACT_RELU: return z > 0.0f ? 1.0f : 0.0f;
        case NRF_ACT_SIGMOID: { const float y = __frcp_rn(1.0f + __expf(-z)); return y * (1.0f - y); }
        case NRF_ACT_EXP: return __expf(z);
        case NRF_ACT_TRUNC_EXP: return __expf(fminf(fmaxf(__half2float(__float2half_rn(z)), -15.0f), 15.0f));   // tcnn_nerf.py:66-68
        default: return 1.0f;
    }
}
__device__ __forceinline__ uint32_t tpack(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// round two f32 to f16 and clamp negatives to zero in ONE instruction (F2FP.RELU.F16.F32.PACK_AB): relu(round(x)) ==
// round(relu(x)), so this is the hidden activation exactly as cvt + HMNMX2 produced it
__device__ __forceinline__ uint32_t tpack_relu(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// c + a.lo * b.lo, then + a.hi * b.hi, for packed f16 pairs a, b: the mixed-precision FMA of sm_100 (FHFMA: f16 x f16 + f32 ->
// f32, one rounding) -- the same value as converting both halves to f32 and using FFMA, without the conversions
__device__ __forceinline__ float fhfma2(uint32_t a, uint32_t b, float c) {
    float d;
    asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t.reg .f32 t;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
        "fma.rn.f32.f16 t, al, bl, %3;\n\tfma.rn.f32.f16 %0, ah, bh, t;\n\t}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
    return d;
}
__device__ __forceinline__ uint32_t h2bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 bits2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

// columns [8c, 8c+8) of row `row` of a row-major [B, n] matrix (f16 or f32) as 8 packed halfs; zero outside.
// On the vectorised f16 path the result is the raw load (no dependent instruction: usable as a prefetch).
__device__ __forceinline__ uint4 load_chunk(const void* __restrict__ base, int dt, size_t row, uint32_t c, uint32_t n, bool row_ok,
                                            bool vec_ok) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (!row_ok || 8 * c >= n) return r;
    if (dt == NRF_DTYPE_F16) {
        const __half* p = reinterpret_cast<const __half*>(base) + row * n + 8 * c;
        if (vec_ok) return __ldg(reinterpret_cast<const uint4*>(p));
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (8 * c + j < n) ? __half2float(p[j]) : 0.0f;
        return make_uint4(tpack(v[0], v[1]), tpack(v[2], v[3]), tpack(v[4], v[5]), tpack(v[6], v[7]));
    }
    const float* p = reinterpret_cast<const float*>(base) + row * n + 8 * c;
    float v[8];
    if (vec_ok) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (8 * c + j < n) ? p[j] : 0.0f;
    }
    return make_uint4(tpack(v[0], v[1]), tpack(v[2], v[3]), tpack(v[4], v[5]), tpack(v[6], v[7]));
}

// The input rows of a 128-row tile -> registers -> the chunked shared-memory tile.  coal: warp-cooperative mapping over
// the warp's 32 rows x CPR chunks (piece i * 32 + lane: consecutive lanes read consecutive 16-byte pieces of global
// memory -- 4 cache lines per request instead of 16 for the row-owner mapping; ncu showed the MLP kernels L1TEX-bound);
// otherwise every thread reads its own row (f32 inputs, odd widths).
// WIDE: 32-byte pieces (one sector, LDG.E.256) -- a warp request reads 1 KB of consecutive global memory.  Measured: the
// backward (4 CTAs/SM) gains from it, the forward (6 CTAs/SM) loses 5 %, so the forward keeps 16-byte pieces.
template <int CPR, bool WIDE>
__device__ __forceinline__ void load_x_tile(uint4 (&xr)[CPR], const void* __restrict__ x, int x_dt, size_t tile_row0, bool tile_ok, uint32_t B,
                                            uint32_t n_in, bool x_vec, bool coal, int warp, int lane, int tid) {
    if (coal && WIDE) {
        constexpr int PPR = CPR / 2;
#pragma unroll
        for (int i = 0; i < PPR; i++) {
            const int idx = i * 32 + lane, r = idx / PPR, c2 = idx - r * PPR;
            const size_t row = tile_row0 + warp * 32 + r;
            if (tile_ok && row < B) {
                const __half* p = reinterpret_cast<const __half*>(x) + row * n_in + 16 * c2;
                asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(xr[2 * i].x), "=r"(xr[2 * i].y), "=r"(xr[2 * i].z), "=r"(xr[2 * i].w),
                               "=r"(xr[2 * i + 1].x), "=r"(xr[2 * i + 1].y), "=r"(xr[2 * i + 1].z), "=r"(xr[2 * i + 1].w) : "l"(p));
            } else {
                xr[2 * i] = make_uint4(0u, 0u, 0u, 0u);
                xr[2 * i + 1] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    } else if (coal) {
#pragma unroll
        for (int i = 0; i < CPR; i++) {
            const int idx = i * 32 + lane, r = idx / CPR, c = idx - r * CPR;
            const size_t row = tile_row0 + warp * 32 + r;
            xr[i] = (tile_ok && row < B) ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(x) + row * n_in) + c)
                                         : make_uint4(0u, 0u, 0u, 0u);
        }
    } else {
        const size_t row = tile_row0 + tid;
#pragma unroll
        for (int c = 0; c < CPR; c++) xr[c] = load_chunk(x, x_dt, row, c, n_in, tile_ok && row < B, x_vec);
    }
}
template <int CPR, bool WIDE>
__device__ __forceinline__ void stage_x_tile(const uint4 (&xr)[CPR], uint8_t* sX, uint32_t CH, bool coal, int warp, int lane, int tid) {
    if (coal && WIDE) {
        constexpr int PPR = CPR / 2;
#pragma unroll
        for (int i = 0; i < PPR; i++) {
            const int idx = i * 32 + lane, r = idx / PPR, c2 = idx - r * PPR;
            *reinterpret_cast<uint4*>(sX + (2 * c2) * CH + (warp * 32 + r) * 16) = xr[2 * i];
            *reinterpret_cast<uint4*>(sX + (2 * c2 + 1) * CH + (warp * 32 + r) * 16) = xr[2 * i + 1];
        }
    } else if (coal) {
#pragma unroll
        for (int i = 0; i < CPR; i++) {
            const int idx = i * 32 + lane, r = idx / CPR, c = idx - r * CPR;
            *reinterpret_cast<uint4*>(sX + c * CH + (warp * 32 + r) * 16) = xr[i];
        }
    } else {
#pragma unroll
        for (int c = 0; c < CPR; c++) *reinterpret_cast<uint4*>(sX + c * CH + tid * 16) = xr[c];
    }
}

// 8 consecutive f32 values -> columns [8c, 8c+8) of row `row` of a row-major matrix with n columns and row stride ld
__device__ __forceinline__ void store_chunk(void* __restrict__ base, int dt, size_t row, uint32_t c, uint32_t n, uint32_t ld,
                                            const float (&v)[8], bool vec_ok) {
    if (8 * c >= n) return;
    if (dt == NRF_DTYPE_F16) {
        __half* p = reinterpret_cast<__half*>(base) + row * ld + 8 * c;
        if (vec_ok) { *reinterpret_cast<uint4*>(p) = make_uint4(tpack(v[0], v[1]), tpack(v[2], v[3]), tpack(v[4], v[5]), tpack(v[6], v[7])); return; }
#pragma unroll
        for (int j = 0; j < 8; j++) if (8 * c + j < n) p[j] = __float2half_rn(v[j]);
        return;
    }
    float* p = reinterpret_cast<float*>(base) + row * ld + 8 * c;
    if (vec_ok) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) if (8 * c + j < n) p[j] = v[j];
}
// same, but ADDED to what is there with fire-and-forget vector reductions (REDG.ADD.F16x8 / F32x4): lets two networks that
// share an input accumulate their input gradients into one buffer without a read-modify-write pass
__device__ __forceinline__ void red_chunk(void* __restrict__ base, int dt, size_t row, uint32_t c, uint32_t n, uint32_t ld,
                                          const float (&v)[8], bool vec_ok) {
    if (8 * c >= n) return;
    if (dt == NRF_DTYPE_F16) {
        __half* p = reinterpret_cast<__half*>(base) + row * ld + 8 * c;
        if (vec_ok) {
            asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(tpack(v[0], v[1])), "r"(tpack(v[2], v[3])),
                         "r"(tpack(v[4], v[5])), "r"(tpack(v[6], v[7])) : "memory");
            return;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) if (8 * c + j < n) atomicAdd(p + j, __float2half_rn(v[j]));
        return;
    }
    float* p = reinterpret_cast<float*>(base) + row * ld + 8 * c;
    if (vec_ok) {
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) if (8 * c + j < n) atomicAdd(p + j, v[j]);
}

__device__ __forceinline__ bool vec_ok_for(const void* base, uint32_t n, uint32_t ld) {
    if (!base) return false;
    return (n % 8 == 0) && (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
}

// this row's output gradient (<= 16 values), RAW (one register per element, or two uint4 on the vector path): no
// instruction depends on the loaded data until dy_convert(), so the load can stay in flight for a whole tile
__device__ __forceinline__ void dy_load_raw(const void* __restrict__ dy, int dy_dt, size_t row, uint32_t n_out, uint32_t ld, bool row_ok,
                                            bool vec_ok, uint32_t (&raw)[16]) {
#pragma unroll
    for (int j = 0; j < 16; j++) raw[j] = 0u;
    if (!row_ok) return;
    if (dy_dt == NRF_DTYPE_F16) {
        const __half* p = reinterpret_cast<const __half*>(dy) + row * ld;
        if (vec_ok) {                                  // n_out in {8, 16}
            if (n_out > 8 && ld % 16 == 0 && (reinterpret_cast<uintptr_t>(dy) & 31) == 0) {      // one 256-bit load (LDG.E.256)
                asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(raw[0]), "=r"(raw[1]), "=r"(raw[2]), "=r"(raw[3]), "=r"(raw[4]), "=r"(raw[5]), "=r"(raw[6]), "=r"(raw[7]) : "l"(p));
                return;
            }
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
            raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w;
            if (n_out > 8) { const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1); raw[4] = b.x; raw[5] = b.y; raw[6] = b.z; raw[7] = b.w; }
            return;
        }
#pragma unroll
        for (int j = 0; j < 16; j++) if ((uint32_t)j < n_out) raw[j] = (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p) + j);
        return;
    }
    const float* p = reinterpret_cast<const float*>(dy) + row * ld;
#pragma unroll
    for (int j = 0; j < 16; j++) if ((uint32_t)j < n_out) raw[j] = __float_as_uint(__ldg(p + j));
}
// raw -> loss_scale * dy as 8 packed f16x2
__device__ __forceinline__ void dy_convert(const uint32_t (&raw)[16], int dy_dt, bool vec_ok, float loss_scale, uint32_t (&out)[8]) {
    if (dy_dt == NRF_DTYPE_F16 && vec_ok) {
        const __half2 ls = __float2half2_rn(loss_scale);
#pragma unroll
        for (int q = 0; q < 8; q++) out[q] = h2bits(__hmul2(bits2h(raw[q]), ls));
        return;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
        float a, b;
        if (dy_dt == NRF_DTYPE_F16) {
            a = __half2float(__ushort_as_half((unsigned short)raw[2 * q]));
            b = __half2float(__ushort_as_half((unsigned short)raw[2 * q + 1]));
        } else { a = __uint_as_float(raw[2 * q]); b = __uint_as_float(raw[2 * q + 1]); }
        out[q] = tpack(a * loss_scale, b * loss_scale);
    }
}

// stage a row-major f16 [rows, cols] weight matrix into the chunked layout (chunk stride ch), 16 bytes per step
__device__ __forceinline__ void stage_weight(uint8_t* dst, uint32_t ch, const __half* __restrict__ src, int rows, int cols) {
    const int cpr = cols / 8;
    for (int i = threadIdx.x; i < rows * cpr; i += TC_THREADS) {
        const int r = i / cpr, c = i - r * cpr;
        *reinterpret_cast<uint4*>(dst + c * ch + r * 16) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * cols + 8 * c));
    }
}

// forward hidden-layer epilogue: 64 f32 accumulator columns of this thread's row -> f16 -> relu -> chunks of `dst_row`.
// relu is applied AFTER the rounding on packed halfs (rounding is monotonic and sign-preserving: same result, half the
// instructions: one cvt.rn.f16x2.f32 + one HMNMX2 per pair of columns).
__device__ __forceinline__ void hidden_fwd_epilogue(uint32_t tacc_lane, uint8_t* dst_row, bool relu, uint32_t CH) {
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t v[32];
        tc05::tmem_ld32(tacc_lane + 32 * half, v);
        tc05::tmem_ld_wait();
        if (relu) {            // warp-uniform
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t p[4];
#pragma unroll
                for (int q = 0; q < 4; q++) p[q] = tpack_relu(__uint_as_float(v[8 * c + 2 * q]), __uint_as_float(v[8 * c + 2 * q + 1]));
                *reinterpret_cast<uint4*>(dst_row + (4 * half + c) * CH) = make_uint4(p[0], p[1], p[2], p[3]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t p[4];
#pragma unroll
                for (int q = 0; q < 4; q++) p[q] = tpack(__uint_as_float(v[8 * c + 2 * q]), __uint_as_float(v[8 * c + 2 * q + 1]));
                *reinterpret_cast<uint4*>(dst_row + (4 * half + c) * CH) = make_uint4(p[0], p[1], p[2], p[3]);
            }
        }
    }
}

// backward hidden-layer epilogue: dH = round_f16(acc) masked by relu'(h), where h is this row of the recomputed
// activation tile in shared memory (h > 0 <=> pre-activation > 0): one cvt + one HSET2 + one LOP3 per pair of columns.
__device__ __forceinline__ void hidden_bwd_epilogue(uint32_t tacc_lane, const uint8_t* h_row, uint8_t* dst_row, bool relu, uint32_t CH) {
    const __half2 zero = __float2half2_rn(0.0f);
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t v[32];
        tc05::tmem_ld32(tacc_lane + 32 * half, v);
        tc05::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint4 hv = *reinterpret_cast<const uint4*>(h_row + (4 * half + c) * CH);
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
            uint32_t p[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                p[q] = h2bits(__floats2half2_rn(__uint_as_float(v[8 * c + 2 * q]), __uint_as_float(v[8 * c + 2 * q + 1])));
                if (relu) p[q] &= __hgt2_mask(bits2h(hw[q]), zero);
            }
            *reinterpret_cast<uint4*>(dst_row + (4 * half + c) * CH) = make_uint4(p[0], p[1], p[2], p[3]);
        }
    }
}

// row owner -> issuer: my operand writes (generic proxy) and my TMEM reads are done
__device__ __forceinline__ void publish(uint64_t* ready) {
    tc05::fence_async_smem();
    tc05::fence_before_sync();
    tc05::mbar_arrive(ready);
}
// block-wide variant used in the prologue / teardown
__device__ __forceinline__ void publish_and_sync() {
    tc05::fence_async_smem();
    tc05::fence_before_sync();
    __syncthreads();
}

// advance the start-address field of a shared-memory descriptor by `bytes` (no carry out of the 14-bit field: smem < 256 KB)
__device__ __forceinline__ uint64_t dadv(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

}  // namespace
