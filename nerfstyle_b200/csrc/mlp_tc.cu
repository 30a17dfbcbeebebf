// tcnn-style fully fused MLP on the 5th-generation tensor cores (tcgen05 + TMEM) -- forward and backward.
//
// Same contract as mlp.cu (call sites /root/reference/networks/style_nerf.py:44-98):
// y = act_out(W_n relu(... relu(W_1 x))), f16 operands, f32 accumulation, hidden activations rounded to f16.
//
// Design (one CTA = 128 threads = one 128-row tile at a time, persistent over tiles, several CTAs per SM):
//   * thread t owns row t of the tile == TMEM lane t.  Every operand tile lives in shared memory in the "chunked"
//     SWIZZLE_NONE layout of tc05.cuh, which is a legal K-major AND MN-major UMMA operand, so the backward forms
//     dH = dZ W, dX = dH W1 and dW = dH^T X from the very same tiles with no transposed copy;
//   * one elected thread issues tcgen05.mma (M=128 for activations, M=64 for weight gradients), the f32 accumulator
//     sits in TMEM; after the commit lands on an mbarrier every thread pulls its row with tcgen05.ld, applies
//     relu / act', rounds to f16 and writes the next layer's operand back to shared memory (generic -> async proxy fence);
//   * the weight gradients are accumulated IN TMEM across all tiles of the persistent CTA (f32), read out once at the
//     end and flushed with one float atomic per weight per CTA.  The dW MMAs of a tile run on the tensor pipe while
//     the threads are already busy with the next epilogue (separate mbarrier, waited on at the top of the next tile);
//   * the backward recomputes the hidden activations from x (nothing but x is saved by the forward); relu masks are
//     kept as 64-bit register masks per row.
// The tensor pipe is nearly idle by design (a 64-wide MLP is ~20 FLOP/B): what the tcgen05 path buys is that the
// accumulators, transposes and weight-gradient reductions leave the register file / LSU, so the kernel runs at the
// rate rows stream through HBM.
#include "common.cuh"
#include "tc05.cuh"

namespace {

constexpr uint32_t CH = 2048;    // chunk stride of 128-row activation tiles
constexpr uint32_t CHW = 1024;   // chunk stride of 64-row weight tiles (W1, Wh)
constexpr uint32_t CHO = 256;    // chunk stride of the 16-row output weight tile
constexpr int TC_THREADS = 128;

__device__ __forceinline__ float tact_fwd(float z, int act) {
    switch (act) {
        case NRF_ACT_RELU: return fmaxf(z, 0.0f);
        case NRF_ACT_SIGMOID: return 1.0f / (1.0f + __expf(-z));
        case NRF_ACT_EXP: return __expf(z);
        default: return z;
    }
}
__device__ __forceinline__ float tact_bwd(float z, int act) {
    switch (act) {
        case NRF_ACT_RELU: return z > 0.0f ? 1.0f : 0.0f;
        case NRF_ACT_SIGMOID: { const float y = 1.0f / (1.0f + __expf(-z)); return y * (1.0f - y); }
        case NRF_ACT_EXP: return __expf(z);
        default: return 1.0f;
    }
}
__device__ __forceinline__ uint32_t tpack(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// columns [8c, 8c+8) of row `row` of a row-major [B, n] matrix (f16 or f32) as 8 packed halfs; zero outside
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

// 8 consecutive f32 values -> columns [8c, 8c+8) of row `row` of a row-major [B, n] matrix (f16 or f32)
__device__ __forceinline__ void store_chunk(void* __restrict__ base, int dt, size_t row, uint32_t c, uint32_t n, const float (&v)[8],
                                            bool vec_ok) {
    if (8 * c >= n) return;
    if (dt == NRF_DTYPE_F16) {
        __half* p = reinterpret_cast<__half*>(base) + row * n + 8 * c;
        if (vec_ok) { *reinterpret_cast<uint4*>(p) = make_uint4(tpack(v[0], v[1]), tpack(v[2], v[3]), tpack(v[4], v[5]), tpack(v[6], v[7])); return; }
#pragma unroll
        for (int j = 0; j < 8; j++) if (8 * c + j < n) p[j] = __float2half_rn(v[j]);
        return;
    }
    float* p = reinterpret_cast<float*>(base) + row * n + 8 * c;
    if (vec_ok) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) if (8 * c + j < n) p[j] = v[j];
}

__device__ __forceinline__ bool vec_ok_for(const void* base, int dt, uint32_t n) {
    if (!base) return false;
    return (n % 8 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
}

// stage a row-major f16 [rows, cols] weight matrix into the chunked layout (chunk stride ch), 16 bytes per step
__device__ __forceinline__ void stage_weight(uint8_t* dst, uint32_t ch, const __half* __restrict__ src, int rows, int cols) {
    const int cpr = cols / 8;
    for (int i = threadIdx.x; i < rows * cpr; i += TC_THREADS) {
        const int r = i / cpr, c = i - r * cpr;
        *reinterpret_cast<uint4*>(dst + c * ch + r * 16) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * cols + 8 * c));
    }
}

// hidden-layer epilogue: 64 f32 accumulator columns of this thread's row -> act -> f16 chunks in `dst` (+ sign mask)
// MASKED = false: forward relu (or identity) and the mask of positive pre-activations is returned
// MASKED = true : values are zeroed where `mask` is clear (back-propagation through relu)
template <bool MASKED>
__device__ __forceinline__ unsigned long long hidden_epilogue(uint32_t tacc_lane, uint8_t* dst_row, bool relu, unsigned long long mask) {
    unsigned long long out_mask = 0ull;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t v[32];
        tc05::tmem_ld32(tacc_lane + 32 * half, v);
        tc05::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                float a = __uint_as_float(v[8 * c + j]);
                const int bit = 32 * half + 8 * c + j;
                if (MASKED) {
                    if (relu && !((mask >> bit) & 1ull)) a = 0.0f;
                } else {
                    if (a > 0.0f) out_mask |= (1ull << bit);
                    if (relu) a = fmaxf(a, 0.0f);
                }
                f[j] = a;
            }
            *reinterpret_cast<uint4*>(dst_row + (4 * half + c) * CH) =
                make_uint4(tpack(f[0], f[1]), tpack(f[2], f[3]), tpack(f[4], f[5]), tpack(f[6], f[7]));
        }
    }
    return out_mask;
}

// publish this thread's shared-memory / TMEM accesses to the MMA-issuing thread, then block barrier
__device__ __forceinline__ void publish_and_sync() {
    tc05::fence_async_smem();
    tc05::fence_before_sync();
    __syncthreads();
}

struct TcSmem { uint32_t W1, Wh, Wo, X, H1, H2, dZ, dH1, dH2, total; };

template <int IN_KT, int NH>
__host__ __device__ constexpr TcSmem fwd_smem() {
    TcSmem s{};
    uint32_t o = 0;
    s.W1 = o; o += IN_KT * 2 * CHW;
    s.Wh = o; o += (NH - 1) * 8 * CHW;
    s.Wo = o; o += 8 * CHO;
    s.X = o;  o += IN_KT * 2 * CH;
    s.H1 = o; o += 8 * CH;
    s.total = o;
    return s;
}
template <int IN_KT, int NH>
__host__ __device__ constexpr TcSmem bwd_smem() {
    TcSmem s{};
    uint32_t o = 0;
    s.W1 = o; o += IN_KT * 2 * CHW;
    s.Wh = o; o += (NH - 1) * 8 * CHW;
    s.Wo = o; o += 8 * CHO;
    s.X = o;  o += IN_KT * 2 * CH;
    s.H1 = o; o += 8 * CH;
    s.H2 = o; o += (NH - 1) * 8 * CH;
    s.dZ = o; o += 2 * CH;
    s.dH1 = o; o += 8 * CH;
    s.dH2 = o; o += (NH - 1) * 8 * CH;
    s.total = o;
    return s;
}

// ------------------------------------------------------------------------------------------------ forward
template <int IN_KT, int NH>
__global__ void __launch_bounds__(TC_THREADS, 5)
k_mlp_fwd_tc(const void* __restrict__ x, int x_dt, const __half* __restrict__ params, uint32_t B, uint32_t n_in, uint32_t n_out,
             int hidden_act, int out_act, void* __restrict__ y, int y_dt) {
    constexpr int IN_PAD = IN_KT * 16;
    constexpr TcSmem L = fwd_smem<IN_KT, NH>();
    constexpr uint32_t TCOLS = 64;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    stage_weight(smem + L.W1, CHW, params, 64, IN_PAD);
    if constexpr (NH == 2) stage_weight(smem + L.Wh, CHW, params + 64 * IN_PAD, 64, 64);
    stage_weight(smem + L.Wo, CHO, params + 64 * IN_PAD + (NH - 1) * 64 * 64, 16, 64);
    if (warp == 0) { tc05::tmem_alloc(&tmem_slot, TCOLS); tc05::tmem_relinquish(); }
    if (tid == 0) { tc05::mbar_init(&bar, 1); tc05::fence_mbar_init(); }
    publish_and_sync();
    tc05::fence_after_sync();
    const uint32_t tacc = tmem_slot;
    const uint32_t tacc_lane = tacc + ((uint32_t)(warp * 32) << 16);
    const uint32_t sW1 = tc05::smem_u32(smem + L.W1), sWh = tc05::smem_u32(smem + L.Wh), sWo = tc05::smem_u32(smem + L.Wo);
    const uint32_t sX = tc05::smem_u32(smem + L.X), sH = tc05::smem_u32(smem + L.H1);
    constexpr uint32_t ID_H = tc05::idesc_f16(128, 64, false, false);
    constexpr uint32_t ID_O = tc05::idesc_f16(128, 16, false, false);
    const bool relu = hidden_act == NRF_ACT_RELU;
    const bool x_vec = vec_ok_for(x, x_dt, n_in) && (x_dt == NRF_DTYPE_F16 || (reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const bool y_vec = vec_ok_for(y, y_dt, n_out);
    const uint32_t ntiles = (B + 127) / 128;
    uint32_t phase = 0;
    uint8_t* xrow = smem + L.X + tid * 16;
    uint8_t* hrow = smem + L.H1 + tid * 16;

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t row = (size_t)tile * 128 + tid;
        const bool row_ok = row < B;
#pragma unroll
        for (int c = 0; c < IN_PAD / 8; c++) *reinterpret_cast<uint4*>(xrow + c * CH) = load_chunk(x, x_dt, row, c, n_in, row_ok, x_vec);
        publish_and_sync();
        if (tid == 0) {
            tc05::fence_after_sync();
#pragma unroll
            for (int k = 0; k < IN_KT; k++)
                tc05::mma_f16(tacc, tc05::desc_kmajor(sX + k * 2 * CH, CH), tc05::desc_kmajor(sW1 + k * 2 * CHW, CHW), ID_H, k > 0);
            tc05::mma_commit(&bar);
        }
        tc05::mbar_wait(&bar, phase); phase ^= 1;
        tc05::fence_after_sync();
#pragma unroll
        for (int l = 0; l < NH; l++) {
            hidden_epilogue<false>(tacc_lane, hrow, relu, 0ull);
            publish_and_sync();
            if (tid == 0) {
                tc05::fence_after_sync();
                if (l < NH - 1) {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        tc05::mma_f16(tacc, tc05::desc_kmajor(sH + k * 2 * CH, CH), tc05::desc_kmajor(sWh + k * 2 * CHW, CHW), ID_H, k > 0);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        tc05::mma_f16(tacc, tc05::desc_kmajor(sH + k * 2 * CH, CH), tc05::desc_kmajor(sWo + k * 2 * CHO, CHO), ID_O, k > 0);
                }
                tc05::mma_commit(&bar);
            }
            tc05::mbar_wait(&bar, phase); phase ^= 1;
            tc05::fence_after_sync();
        }
        uint32_t z[16];
        tc05::tmem_ld16(tacc_lane, z);
        tc05::tmem_ld_wait();
        if (row_ok) {
#pragma unroll
            for (int c = 0; c < 2; c++) {
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; j++) f[j] = tact_fwd(__uint_as_float(z[8 * c + j]), out_act);
                store_chunk(y, y_dt, row, c, n_out, f, y_vec);
            }
        }
        // the next tile's publish_and_sync orders these TMEM reads before the next MMA overwrites the accumulator
    }
    publish_and_sync();
    if (warp == 0) tc05::tmem_dealloc(tacc, TCOLS);
}

// ------------------------------------------------------------------------------------------------ backward
template <int IN_KT, int NH>
__host__ __device__ constexpr uint32_t bwd_tmem_cols() {
    const uint32_t need = 64 + IN_KT * 16 + 16 + (NH - 1) * 64;
    return need <= 128 ? 128u : 256u;
}

template <int IN_KT, int NH>
__global__ void __launch_bounds__(TC_THREADS, (NH == 2 || IN_KT == 4) ? 2 : 4)
k_mlp_bwd_tc(const void* __restrict__ x, int x_dt, const __half* __restrict__ params, const void* __restrict__ dy, int dy_dt,
             uint32_t B, uint32_t n_in, uint32_t n_out, int hidden_act, int out_act, float loss_scale, void* __restrict__ dx,
             float* __restrict__ dparams) {
    constexpr int IN_PAD = IN_KT * 16;
    constexpr TcSmem L = bwd_smem<IN_KT, NH>();
    constexpr uint32_t TCOLS = bwd_tmem_cols<IN_KT, NH>();
    constexpr uint32_t T_W1 = 64, T_WO = 64 + IN_PAD, T_WH = 64 + IN_PAD + 16;     // TMEM columns of the dW accumulators
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2];      // [0]: epilogue-critical MMAs, [1]: weight-gradient MMAs of a tile
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    stage_weight(smem + L.W1, CHW, params, 64, IN_PAD);
    if constexpr (NH == 2) stage_weight(smem + L.Wh, CHW, params + 64 * IN_PAD, 64, 64);
    stage_weight(smem + L.Wo, CHO, params + 64 * IN_PAD + (NH - 1) * 64 * 64, 16, 64);
    if (warp == 0) { tc05::tmem_alloc(&tmem_slot, TCOLS); tc05::tmem_relinquish(); }
    if (tid == 0) { tc05::mbar_init(&bars[0], 1); tc05::mbar_init(&bars[1], 1); tc05::fence_mbar_init(); }
    publish_and_sync();
    tc05::fence_after_sync();
    const uint32_t tacc = tmem_slot;
    const uint32_t tacc_lane = tacc + ((uint32_t)(warp * 32) << 16);
    const uint32_t sW1 = tc05::smem_u32(smem + L.W1), sWh = tc05::smem_u32(smem + L.Wh), sWo = tc05::smem_u32(smem + L.Wo);
    const uint32_t sX = tc05::smem_u32(smem + L.X), sH1 = tc05::smem_u32(smem + L.H1), sH2 = tc05::smem_u32(smem + L.H2);
    const uint32_t sdZ = tc05::smem_u32(smem + L.dZ), sdH1 = tc05::smem_u32(smem + L.dH1), sdH2 = tc05::smem_u32(smem + L.dH2);
    const uint32_t sHlast = (NH == 2) ? sH2 : sH1, sdHlast = (NH == 2) ? sdH2 : sdH1;
    constexpr uint32_t ID_H = tc05::idesc_f16(128, 64, false, false);        // X W1^T, H1 Wh^T      (A K-major, B K-major)
    constexpr uint32_t ID_O = tc05::idesc_f16(128, 16, false, false);        // Hlast Wo^T
    constexpr uint32_t ID_DH = tc05::idesc_f16(128, 64, false, true);        // dZ Wo, dH2 Wh        (B MN-major)
    constexpr uint32_t ID_DX = tc05::idesc_f16(128, IN_PAD, false, true);    // dH1 W1
    constexpr uint32_t ID_GW1 = tc05::idesc_f16(64, IN_PAD, true, true);     // dH1^T X
    constexpr uint32_t ID_GWH = tc05::idesc_f16(64, 64, true, true);         // dH2^T H1
    constexpr uint32_t ID_GWO = tc05::idesc_f16(64, 16, true, true);         // Hlast^T dZ  (= dWo^T)
    const bool relu = hidden_act == NRF_ACT_RELU;
    const bool x_vec = vec_ok_for(x, x_dt, n_in) && (x_dt == NRF_DTYPE_F16 || (reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const bool dx_vec = vec_ok_for(dx, x_dt, n_in);
    const float inv_scale = 1.0f / loss_scale;
    const uint32_t ntiles = (B + 127) / 128;
    const bool want_dw = dparams != nullptr;
    uint32_t phase = 0, phase_w = 0;
    bool first = true;

    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t row = (size_t)tile * 128 + tid;
        const bool row_ok = row < B;
        // ---- global loads of this tile's rows first (they overlap the previous tile's weight-gradient MMAs)
        uint4 xr[IN_PAD / 8];
#pragma unroll
        for (int c = 0; c < IN_PAD / 8; c++) xr[c] = load_chunk(x, x_dt, row, c, n_in, row_ok, x_vec);
        float dyr[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            float v = 0.0f;
            if (row_ok && (uint32_t)j < n_out)
                v = (dy_dt == NRF_DTYPE_F16) ? __half2float(reinterpret_cast<const __half*>(dy)[row * n_out + j])
                                             : reinterpret_cast<const float*>(dy)[row * n_out + j];
            dyr[j] = v;
        }
        if (!first) { tc05::mbar_wait(&bars[1], phase_w); phase_w ^= 1; tc05::fence_after_sync(); }   // operand tiles are free again
#pragma unroll
        for (int c = 0; c < IN_PAD / 8; c++) *reinterpret_cast<uint4*>(smem + L.X + c * CH + tid * 16) = xr[c];
        publish_and_sync();
        // ---- recompute: H1 = act(X W1^T)
        if (tid == 0) {
            tc05::fence_after_sync();
#pragma unroll
            for (int k = 0; k < IN_KT; k++)
                tc05::mma_f16(tacc, tc05::desc_kmajor(sX + k * 2 * CH, CH), tc05::desc_kmajor(sW1 + k * 2 * CHW, CHW), ID_H, k > 0);
            tc05::mma_commit(&bars[0]);
        }
        tc05::mbar_wait(&bars[0], phase); phase ^= 1;
        tc05::fence_after_sync();
        const unsigned long long m1 = hidden_epilogue<false>(tacc_lane, smem + L.H1 + tid * 16, relu, 0ull);
        publish_and_sync();
        unsigned long long m2 = 0ull;
        if constexpr (NH == 2) {
            if (tid == 0) {
                tc05::fence_after_sync();
#pragma unroll
                for (int k = 0; k < 4; k++)
                    tc05::mma_f16(tacc, tc05::desc_kmajor(sH1 + k * 2 * CH, CH), tc05::desc_kmajor(sWh + k * 2 * CHW, CHW), ID_H, k > 0);
                tc05::mma_commit(&bars[0]);
            }
            tc05::mbar_wait(&bars[0], phase); phase ^= 1;
            tc05::fence_after_sync();
            m2 = hidden_epilogue<false>(tacc_lane, smem + L.H2 + tid * 16, relu, 0ull);
            publish_and_sync();
        }
        // ---- Z = Hlast Wo^T, dZ = loss_scale * dy * act'(Z)
        if (tid == 0) {
            tc05::fence_after_sync();
#pragma unroll
            for (int k = 0; k < 4; k++)
                tc05::mma_f16(tacc, tc05::desc_kmajor(sHlast + k * 2 * CH, CH), tc05::desc_kmajor(sWo + k * 2 * CHO, CHO), ID_O, k > 0);
            tc05::mma_commit(&bars[0]);
        }
        tc05::mbar_wait(&bars[0], phase); phase ^= 1;
        tc05::fence_after_sync();
        {
            uint32_t z[16];
            tc05::tmem_ld16(tacc_lane, z);
            tc05::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 2; c++) {
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; j++) f[j] = dyr[8 * c + j] * loss_scale * tact_bwd(__uint_as_float(z[8 * c + j]), out_act);
                *reinterpret_cast<uint4*>(smem + L.dZ + c * CH + tid * 16) =
                    make_uint4(tpack(f[0], f[1]), tpack(f[2], f[3]), tpack(f[4], f[5]), tpack(f[6], f[7]));
            }
        }
        publish_and_sync();
        // ---- dHlast = (dZ Wo) . relu'   |   dWo^T += Hlast^T dZ
        if (tid == 0) {
            tc05::fence_after_sync();
            tc05::mma_f16(tacc, tc05::desc_kmajor(sdZ, CH), tc05::desc_mnmajor(sWo, CHO), ID_DH, 0);
            tc05::mma_commit(&bars[0]);
            if (want_dw) {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    tc05::mma_f16(tacc + T_WO, tc05::desc_mnmajor(sHlast + k * 256, CH), tc05::desc_mnmajor(sdZ + k * 256, CH), ID_GWO,
                                  (!first || k > 0) ? 1u : 0u);
            }
        }
        tc05::mbar_wait(&bars[0], phase); phase ^= 1;
        tc05::fence_after_sync();
        hidden_epilogue<true>(tacc_lane, smem + ((NH == 2) ? L.dH2 : L.dH1) + tid * 16, relu, (NH == 2) ? m2 : m1);
        publish_and_sync();
        if constexpr (NH == 2) {
            // ---- dH1 = (dH2 Wh) . relu'   |   dWh += dH2^T H1
            if (tid == 0) {
                tc05::fence_after_sync();
#pragma unroll
                for (int k = 0; k < 4; k++)
                    tc05::mma_f16(tacc, tc05::desc_kmajor(sdH2 + k * 2 * CH, CH), tc05::desc_mnmajor(sWh + k * 256, CHW), ID_DH, k > 0);
                tc05::mma_commit(&bars[0]);
                if (want_dw) {
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        tc05::mma_f16(tacc + T_WH, tc05::desc_mnmajor(sdH2 + k * 256, CH), tc05::desc_mnmajor(sH1 + k * 256, CH), ID_GWH,
                                      (!first || k > 0) ? 1u : 0u);
                }
            }
            tc05::mbar_wait(&bars[0], phase); phase ^= 1;
            tc05::fence_after_sync();
            hidden_epilogue<true>(tacc_lane, smem + L.dH1 + tid * 16, relu, m1);
            publish_and_sync();
        }
        // ---- dX = dH1 W1   |   dW1 += dH1^T X
        if (tid == 0) {
            tc05::fence_after_sync();
#pragma unroll
            for (int k = 0; k < 4; k++)
                tc05::mma_f16(tacc, tc05::desc_kmajor(sdH1 + k * 2 * CH, CH), tc05::desc_mnmajor(sW1 + k * 256, CHW), ID_DX, k > 0);
            tc05::mma_commit(&bars[0]);
            if (want_dw) {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    tc05::mma_f16(tacc + T_W1, tc05::desc_mnmajor(sdH1 + k * 256, CH), tc05::desc_mnmajor(sX + k * 256, CH), ID_GW1,
                                  (!first || k > 0) ? 1u : 0u);
            }
            tc05::mma_commit(&bars[1]);
        }
        tc05::mbar_wait(&bars[0], phase); phase ^= 1;
        tc05::fence_after_sync();
        if (dx != nullptr) {
#pragma unroll
            for (int g = 0; g < IN_PAD / 16; g++) {
                uint32_t v[16];
                tc05::tmem_ld16(tacc_lane + 16 * g, v);
                tc05::tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) f[j] = __uint_as_float(v[8 * c + j]) * inv_scale;
                        store_chunk(dx, x_dt, row, 2 * g + c, n_in, f, dx_vec);
                    }
                }
            }
        }
        first = false;
        // the next tile's publish_and_sync orders these TMEM reads before its first MMA
    }
    // ---- flush the weight gradients accumulated in TMEM (M=64 layout: row m lives in lane (m % 16) + 32 * (m / 16))
    if (!first) { tc05::mbar_wait(&bars[1], phase_w); tc05::fence_after_sync(); }
    if (want_dw && !first) {
        // tcgen05.ld is warp-collective (.sync.aligned): every lane loads, only lanes 0-15 hold rows of an M=64 accumulator
        const int m = warp * 16 + (lane & 15);
        const bool owner = lane < 16;
        float* dW1 = dparams;
        float* dWh = dparams + 64 * IN_PAD;
        float* dWo = dWh + (NH - 1) * 64 * 64;
#pragma unroll 1
        for (int g = 0; g < IN_PAD / 16; g++) {        // dW1[m][:]
            uint32_t v[16];
            tc05::tmem_ld16(tacc_lane + T_W1 + 16 * g, v);
            tc05::tmem_ld_wait();
            if (owner) {
#pragma unroll
                for (int j = 0; j < 16; j++) atomicAdd(dW1 + m * IN_PAD + 16 * g + j, __uint_as_float(v[j]) * inv_scale);
            }
        }
        if constexpr (NH == 2) {
#pragma unroll 1
            for (int g = 0; g < 4; g++) {              // dWh[m][:]
                uint32_t v[16];
                tc05::tmem_ld16(tacc_lane + T_WH + 16 * g, v);
                tc05::tmem_ld_wait();
                if (owner) {
#pragma unroll
                    for (int j = 0; j < 16; j++) atomicAdd(dWh + m * 64 + 16 * g + j, __uint_as_float(v[j]) * inv_scale);
                }
            }
        }
        {                                              // dWo^T[m][o] -> dWo[o][m]
            uint32_t v[16];
            tc05::tmem_ld16(tacc_lane + T_WO, v);
            tc05::tmem_ld_wait();
            if (owner) {
#pragma unroll
                for (int j = 0; j < 16; j++) atomicAdd(dWo + j * 64 + m, __uint_as_float(v[j]) * inv_scale);
            }
        }
    }
    publish_and_sync();
    if (warp == 0) tc05::tmem_dealloc(tacc, TCOLS);
}

int tc_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

int g_fwd_ctas_per_sm = 5, g_bwd_ctas_per_sm = 4;

template <int IN_KT, int NH>
int launch_fwd(const void* x, int xdt, const void* params, uint32_t B, uint32_t n_in, uint32_t n_out, int hact, int oact, void* y, int ydt,
               cudaStream_t s) {
    constexpr TcSmem L = fwd_smem<IN_KT, NH>();
    auto kern = k_mlp_fwd_tc<IN_KT, NH>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    const uint32_t ntiles = ceil_div_u32(B, 128);
    const uint32_t per_sm = (uint32_t)max(1, min(min(g_fwd_ctas_per_sm, (int)(220 * 1024 / (L.total + 1024))), 8));
    const uint32_t grid = (uint32_t)min((uint64_t)ntiles, (uint64_t)tc_sm_count() * per_sm);
    kern<<<grid, TC_THREADS, L.total, s>>>(x, xdt, (const __half*)params, B, n_in, n_out, hact, oact, y, ydt);
    return nrf_check_launch();
}

template <int IN_KT, int NH>
int launch_bwd(const void* x, int xdt, const void* params, const void* dy, int dydt, uint32_t B, uint32_t n_in, uint32_t n_out, int hact,
               int oact, float ls, void* dx, float* dparams, cudaStream_t s) {
    constexpr TcSmem L = bwd_smem<IN_KT, NH>();
    constexpr uint32_t TCOLS = bwd_tmem_cols<IN_KT, NH>();
    auto kern = k_mlp_bwd_tc<IN_KT, NH>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    const uint32_t ntiles = ceil_div_u32(B, 128);
    const uint32_t per_sm = (uint32_t)max(1, min(min(g_bwd_ctas_per_sm, (int)(220 * 1024 / (L.total + 1024))), (int)(512 / TCOLS)));
    const uint32_t grid = (uint32_t)min((uint64_t)ntiles, (uint64_t)tc_sm_count() * per_sm);
    kern<<<grid, TC_THREADS, L.total, s>>>(x, xdt, (const __half*)params, dy, dydt, B, n_in, n_out, hact, oact, ls, dx, dparams);
    return nrf_check_launch();
}

}  // namespace

// entry points used by mlp.cu's dispatcher (same argument meaning as nrf_mlp_forward / nrf_mlp_backward)
int nrf_mlp_tc_forward(const void* x, int x_dtype, const void* params_f16, uint32_t B, uint32_t n_in, uint32_t n_out, uint32_t n_hidden,
                       int hidden_act, int out_act, void* y, int y_dtype, cudaStream_t s) {
    const int kt = (int)((n_in + 15) / 16);
#define TC_FWD(K, H) if (kt == K && (int)n_hidden == H) return launch_fwd<K, H>(x, x_dtype, params_f16, B, n_in, n_out, hidden_act, out_act, y, y_dtype, s)
    TC_FWD(1, 1); TC_FWD(2, 1); TC_FWD(3, 1); TC_FWD(4, 1);
    TC_FWD(1, 2); TC_FWD(2, 2); TC_FWD(3, 2); TC_FWD(4, 2);
#undef TC_FWD
    return NRF_E_UNSUPPORTED;
}

int nrf_mlp_tc_backward(const void* x, int x_dtype, const void* params_f16, const void* dy, int dy_dtype, uint32_t B, uint32_t n_in,
                        uint32_t n_out, uint32_t n_hidden, int hidden_act, int out_act, float loss_scale, void* dx, float* dparams,
                        cudaStream_t s) {
    const int kt = (int)((n_in + 15) / 16);
#define TC_BWD(K, H) if (kt == K && (int)n_hidden == H) return launch_bwd<K, H>(x, x_dtype, params_f16, dy, dy_dtype, B, n_in, n_out, hidden_act, out_act, loss_scale, dx, dparams, s)
    TC_BWD(1, 1); TC_BWD(2, 1); TC_BWD(3, 1); TC_BWD(4, 1);
    TC_BWD(1, 2); TC_BWD(2, 2); TC_BWD(3, 2); TC_BWD(4, 2);
#undef TC_BWD
    return NRF_E_UNSUPPORTED;
}

void nrf_mlp_tc_set_ctas(int fwd_per_sm, int bwd_per_sm) {
    if (fwd_per_sm > 0) g_fwd_ctas_per_sm = fwd_per_sm;
    if (bwd_per_sm > 0) g_bwd_ctas_per_sm = bwd_per_sm;
}
