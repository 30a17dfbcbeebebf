// The four networks of the field (networks/style_nerf.py:120-142, use_dir=False) in ONE tcgen05 launch.
//
//   sigma   = trunc_exp(density_net(enc_d))                      32 -> 64 -> 1
//   classes = class_net(enc_c)                                   32 -> 64 -> K
//   rgb     = color2_net(color1_net(enc_c))                      32 -> 64 -> 16 -> 64 -> 64 -> 3 (sigmoid)
//   rgbs    = cat(rgb, classes)
//
// Same arithmetic as four nrf_mlp_forward_ex calls (f16 operands, f32 accumulation, hidden activations and network
// outputs rounded to f16 at the same points), but a 128-point tile goes through FIVE MMA round trips instead of nine and
// the encodings are staged once:
//   S1  Hd = Xd W1d^T (N=64)  |  [Hk | Hc] = Xc [W1k ; W1c]^T (ONE N=128 product: the two heads that share enc_c)
//   S2  [zk | c1] = [Hk | Hc] blockdiag(Wok, Woc)^T (ONE K=128, N=32 chain)  -- the density head's single output is a
//       64-term dot product of the row owner's own registers (CUDA cores, in the S1 epilogue): no MMA, no Hd tile
//   S3  G1 = c1 W1_2^T        S4  G2 = G1 Wh_2^T        S5  z = G2 Wo_2^T
// Accumulators live in 256 TMEM columns (two CTAs per SM), the [Hk | Hc] tile is reused for G1 | G2.  Row owners /
// issuer / mbarrier protocol and the chunked operand layout are those of mlp_tc.cu (mlp_tc_common.cuh, tc05.cuh).
#include "mlp_tc_common.cuh"

namespace {

constexpr uint32_t FCH = 2080;        // chunk stride of the 128-row activation tiles (= ch_for(2))
constexpr uint32_t CHW128 = 2048;     // chunk stride of the 128-row stacked first-layer weights [W1k ; W1c]
constexpr uint32_t CHW32 = 512;       // chunk stride of the 32-row block-diagonal output weights

struct FieldSmem { uint32_t W1d, W1kc, Wkc, W2a, W2h, W2o, wod, X, H, C1, total; };
__host__ __device__ constexpr FieldSmem field_smem(int n_xbuf) {
    FieldSmem s{};
    uint32_t o = 0;
    s.W1d = o;  o += 4 * CHW;          // density W1   [64 x 32]
    s.W1kc = o; o += 4 * CHW128;       // class W1 over color1 W1   [128 x 32]
    s.Wkc = o;  o += 16 * CHW32;       // blockdiag(class Wo, color1 Wo)   [32 x 128]
    s.W2a = o;  o += 2 * CHW;          // color2 W1   [64 x 16]
    s.W2h = o;  o += 8 * CHW;          // color2 Wh   [64 x 64]
    s.W2o = o;  o += 8 * CHO;          // color2 Wo   [16 x 64]
    s.wod = o;  o += 64 * 4;           // density Wo row 0 (64 f16 in the first 128 bytes)
    s.X = o;    o += (uint32_t)n_xbuf * 8 * FCH;      // Xd (4 chunks) | Xc (4 chunks), per buffer
    s.H = o;    o += 16 * FCH;         // [Hk | Hc], later [G1 | G2]
    s.C1 = o;   o += 2 * FCH;          // color1 output, color2's input
    s.total = o;
    return s;
}

// rows [r0, r0 + rows) of a row-major f16 [rows, cols] matrix into a chunked tile whose chunks start at column chunk c0
__device__ __forceinline__ void stage_block(uint8_t* dst, uint32_t ch, const __half* __restrict__ src, int rows, int cols, int r0, int c0) {
    const int cpr = cols / 8;
    for (int i = threadIdx.x; i < rows * cpr; i += TC_THREADS) {
        const int r = i / cpr, c = i - r * cpr;
        *reinterpret_cast<uint4*>(dst + (c0 + c) * ch + (r0 + r) * 16) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * cols + 8 * c));
    }
}

__global__ void __launch_bounds__(TC_THREADS, 2)
k_field_fwd_tc(const __half* __restrict__ enc_d, const __half* __restrict__ enc_c, const __half* __restrict__ w_density,
               const __half* __restrict__ w_class, const __half* __restrict__ w_color1, const __half* __restrict__ w_color2, uint32_t B,
               uint32_t n_classes, float* __restrict__ sigmas, float* __restrict__ rgbs, uint32_t ld_rgbs, __half* __restrict__ c1_out,
               const int32_t* __restrict__ B_dev) {
    if (B_dev) B = min(B, (uint32_t)*B_dev);
    const uint32_t ntiles = (B + 127) / 128;
    if (blockIdx.x >= ntiles) return;
    constexpr FieldSmem L = field_smem(1);
    constexpr uint32_t TCOLS = 256;
    constexpr uint32_t T_HD = 0, T_HKC = 64, T_OUT = 192, T_G1 = 0, T_G2 = 64, T_Z = 128;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_ready, bar_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- weights -> chunked tiles (tcnn layout: W1 [64, in_pad], (Wh [64, 64],) Wo [16, 64], row-major, concatenated)
    stage_block(smem + L.W1d, CHW, w_density, 64, 32, 0, 0);
    stage_block(smem + L.W1kc, CHW128, w_class, 64, 32, 0, 0);
    stage_block(smem + L.W1kc, CHW128, w_color1, 64, 32, 64, 0);
    for (int i = tid; i < 16 * CHW32 / 16; i += TC_THREADS) reinterpret_cast<uint4*>(smem + L.Wkc)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    stage_block(smem + L.Wkc, CHW32, w_class + 64 * 32, 16, 64, 0, 0);           // rows 0-15,  K columns 0-63   (Hk)
    stage_block(smem + L.Wkc, CHW32, w_color1 + 64 * 32, 16, 64, 16, 8);         // rows 16-31, K columns 64-127 (Hc)
    stage_block(smem + L.W2a, CHW, w_color2, 64, 16, 0, 0);
    stage_block(smem + L.W2h, CHW, w_color2 + 64 * 16, 64, 64, 0, 0);
    stage_block(smem + L.W2o, CHO, w_color2 + 64 * 16 + 64 * 64, 16, 64, 0, 0);
    if (tid < 64) reinterpret_cast<__half*>(smem + L.wod)[tid] = __ldg(w_density + 64 * 32 + tid);
    if (warp == 0) { tc05::tmem_alloc(&tmem_slot, TCOLS); tc05::tmem_relinquish(); }
    if (tid == 0) { tc05::mbar_init(&bar_ready, TC_ROWS); tc05::mbar_init(&bar_done, 1); tc05::fence_mbar_init(); }
    publish_and_sync();
    tc05::fence_after_sync();
    const uint32_t tacc = tmem_slot;

    if (warp == 4) {
        // ================================================================ MMA issuer (one lane)
        if (lane == 0) {
            const uint64_t kXd = tc05::desc_kmajor(tc05::smem_u32(smem + L.X), FCH);
            const uint64_t kXc = tc05::desc_kmajor(tc05::smem_u32(smem + L.X + 4 * FCH), FCH);
            const uint64_t kH = tc05::desc_kmajor(tc05::smem_u32(smem + L.H), FCH);
            const uint64_t kG2 = tc05::desc_kmajor(tc05::smem_u32(smem + L.H + 8 * FCH), FCH);
            const uint64_t kC1 = tc05::desc_kmajor(tc05::smem_u32(smem + L.C1), FCH);
            const uint64_t kW1d = tc05::desc_kmajor(tc05::smem_u32(smem + L.W1d), CHW);
            const uint64_t kW1kc = tc05::desc_kmajor(tc05::smem_u32(smem + L.W1kc), CHW128);
            const uint64_t kWkc = tc05::desc_kmajor(tc05::smem_u32(smem + L.Wkc), CHW32);
            const uint64_t kW2a = tc05::desc_kmajor(tc05::smem_u32(smem + L.W2a), CHW);
            const uint64_t kW2h = tc05::desc_kmajor(tc05::smem_u32(smem + L.W2h), CHW);
            const uint64_t kW2o = tc05::desc_kmajor(tc05::smem_u32(smem + L.W2o), CHO);
            constexpr uint32_t ID64 = tc05::idesc_f16(128, 64, false, false), ID128 = tc05::idesc_f16(128, 128, false, false);
            constexpr uint32_t ID32 = tc05::idesc_f16(128, 32, false, false), ID16 = tc05::idesc_f16(128, 16, false, false);
            uint32_t ph = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                // S1: Hd | [Hk | Hc]
                tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
#pragma unroll
                for (int k = 0; k < 2; k++) tc05::mma_f16(tacc + T_HD, dadv(kXd, k * 2 * FCH), dadv(kW1d, k * 2 * CHW), ID64, k > 0);
#pragma unroll
                for (int k = 0; k < 2; k++) tc05::mma_f16(tacc + T_HKC, dadv(kXc, k * 2 * FCH), dadv(kW1kc, k * 2 * CHW128), ID128, k > 0);
                tc05::mma_commit(&bar_done);
                // S2: [zk | c1] = [Hk | Hc] blockdiag^T
                tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
#pragma unroll
                for (int k = 0; k < 8; k++) tc05::mma_f16(tacc + T_OUT, dadv(kH, k * 2 * FCH), dadv(kWkc, k * 2 * CHW32), ID32, k > 0);
                tc05::mma_commit(&bar_done);
                // S3: G1 = c1 W1_2^T
                tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
                tc05::mma_f16(tacc + T_G1, kC1, kW2a, ID64, 0);
                tc05::mma_commit(&bar_done);
                // S4: G2 = G1 Wh_2^T
                tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
#pragma unroll
                for (int k = 0; k < 4; k++) tc05::mma_f16(tacc + T_G2, dadv(kH, k * 2 * FCH), dadv(kW2h, k * 2 * CHW), ID64, k > 0);
                tc05::mma_commit(&bar_done);
                // S5: z = G2 Wo_2^T
                tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
#pragma unroll
                for (int k = 0; k < 4; k++) tc05::mma_f16(tacc + T_Z, dadv(kG2, k * 2 * FCH), dadv(kW2o, k * 2 * CHO), ID16, k > 0);
                tc05::mma_commit(&bar_done);
            }
        }
        __syncwarp();
    } else {
        // ================================================================ row owners
        const uint32_t tl = tacc + ((uint32_t)(warp * 32) << 16);
        uint32_t phase = 0;
        uint8_t* const hrow = smem + L.H + tid * 16;
        uint8_t* const c1row = smem + L.C1 + tid * 16;
        const uint4* const wod = reinterpret_cast<const uint4*>(smem + L.wod);      // 8 x 8 f16 weights of the density output
        const __half2 zero2 = __float2half2_rn(0.0f);
        uint4 xd[4], xc[4];
        load_x_tile<4, false>(xd, enc_d, NRF_DTYPE_F16, (size_t)blockIdx.x * 128, true, B, 32, true, true, warp, lane, tid);
        load_x_tile<4, false>(xc, enc_c, NRF_DTYPE_F16, (size_t)blockIdx.x * 128, true, B, 32, true, true, warp, lane, tid);
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const size_t row = (size_t)tile * 128 + tid;
            const bool row_ok = row < B;
            stage_x_tile<4, false>(xd, smem + L.X, FCH, true, warp, lane, tid);
            stage_x_tile<4, false>(xc, smem + L.X + 4 * FCH, FCH, true, warp, lane, tid);
            publish(&bar_ready);
            {
                const uint32_t nt = tile + gridDim.x;
                load_x_tile<4, false>(xd, enc_d, NRF_DTYPE_F16, (size_t)nt * 128, nt < ntiles, B, 32, true, true, warp, lane, tid);
                load_x_tile<4, false>(xc, enc_c, NRF_DTYPE_F16, (size_t)nt * 128, nt < ntiles, B, 32, true, true, warp, lane, tid);
            }
            // ---- epilogue 1: density output on the CUDA cores; [Hk | Hc] -> f16 -> relu -> shared memory
            tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
            {
                float zd = 0.0f;
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    uint32_t v[32];
                    tc05::tmem_ld32(tl + T_HD + 32 * half, v);
                    tc05::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const uint4 w8 = wod[4 * half + c];
                        const uint32_t wp[4] = {w8.x, w8.y, w8.z, w8.w};
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            // the hidden activation exactly as the MMA path would see it (rounded to f16, then relu), times the
                            // f16 weight, accumulated in f32: FHFMA, same value as cvt + FFMA
                            zd = fhfma2(tpack_relu(__uint_as_float(v[8 * c + 2 * q]), __uint_as_float(v[8 * c + 2 * q + 1])), wp[q], zd);
                        }
                    }
                }
                if (row_ok) sigmas[row] = tact_fwd(zd, NRF_ACT_TRUNC_EXP);
#pragma unroll
                for (int piece = 0; piece < 4; piece++) {
                    uint32_t v[32];
                    tc05::tmem_ld32(tl + T_HKC + 32 * piece, v);
                    tc05::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        uint32_t p[4];
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            p[q] = tpack_relu(__uint_as_float(v[8 * c + 2 * q]), __uint_as_float(v[8 * c + 2 * q + 1]));
                        *reinterpret_cast<uint4*>(hrow + (4 * piece + c) * FCH) = make_uint4(p[0], p[1], p[2], p[3]);
                    }
                }
            }
            publish(&bar_ready);
            // ---- epilogue 2: class logits -> rgbs[:, 3:3+K];  c1 -> f16 -> shared memory (color2's input) and global (saved)
            tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
            {
                uint32_t v[32];
                tc05::tmem_ld32(tl + T_OUT, v);
                tc05::tmem_ld_wait();
                if (row_ok) {
                    float* o = rgbs + row * ld_rgbs + 3;
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if ((uint32_t)j < n_classes) o[j] = __half2float(__float2half_rn(__uint_as_float(v[j])));
                }
                uint32_t p[8];
#pragma unroll
                for (int q = 0; q < 8; q++) p[q] = tpack(__uint_as_float(v[16 + 2 * q]), __uint_as_float(v[16 + 2 * q + 1]));
                *reinterpret_cast<uint4*>(c1row) = make_uint4(p[0], p[1], p[2], p[3]);
                *reinterpret_cast<uint4*>(c1row + FCH) = make_uint4(p[4], p[5], p[6], p[7]);
                if (row_ok && c1_out) {
                    uint4* dst = reinterpret_cast<uint4*>(c1_out + row * 16);
                    dst[0] = make_uint4(p[0], p[1], p[2], p[3]);
                    dst[1] = make_uint4(p[4], p[5], p[6], p[7]);
                }
            }
            publish(&bar_ready);
            // ---- epilogues 3, 4: color2's hidden layers (G1 -> chunks 0-7, G2 -> chunks 8-15 of the H tile)
            tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
            hidden_fwd_epilogue(tl + T_G1, hrow, true, FCH);
            publish(&bar_ready);
            tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
            hidden_fwd_epilogue(tl + T_G2, hrow + 8 * FCH, true, FCH);
            publish(&bar_ready);
            // ---- epilogue 5: rgb = sigmoid(z), f16-rounded like the network output
            tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
            {
                uint32_t z[8];
                tc05::tmem_ld8(tl + T_Z, z);
                tc05::tmem_ld_wait();
                if (row_ok) {
                    float* o = rgbs + row * ld_rgbs;
#pragma unroll
                    for (int j = 0; j < 3; j++) o[j] = __half2float(__float2half_rn(tact_fwd(__uint_as_float(z[j]), NRF_ACT_SIGMOID)));
                }
            }
            // the next tile's publish() orders these TMEM reads before the issuer overwrites the accumulators
        }
    }
    publish_and_sync();
    if (warp == 0) tc05::tmem_dealloc(tacc, TCOLS);
}

int field_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace

// One launch for the whole field head (see the file header).  enc_d / enc_c: f16 [M, 32] (32-byte aligned rows), the four
// weight vectors: f16, tcnn FullyFusedMLP layout of the respective network (density 32->64->1, class 32->64->K, color1
// 32->64->16, color2 16->64->64->3).  sigmas f32 [M]; rgbs f32 [M, ld_rgbs] (columns 0-2 rgb, 3..3+K class logits, both
// holding the f16 network outputs widened); c1_out f16 [M, 16] or NULL (color1's output, what the backward needs).
// M_dev (device int32 or NULL): rows actually present (device-driven inference loop).
NRF_EXPORT int nrf_field_forward(const void* enc_d, const void* enc_c, const void* w_density, const void* w_class, const void* w_color1,
                                 const void* w_color2, uint32_t M, uint32_t n_classes, float* sigmas, float* rgbs, uint32_t ld_rgbs,
                                 void* c1_out, const int32_t* M_dev, void* stream) {
    if (M == 0) return NRF_OK;
    if (!enc_d || !enc_c || !w_density || !w_class || !w_color1 || !w_color2 || !sigmas || !rgbs) return NRF_E_INVALID;
    if (n_classes > 16 || ld_rgbs < 3 + n_classes) return NRF_E_INVALID;
    if ((((uintptr_t)enc_d | (uintptr_t)enc_c) & 31) || (((uintptr_t)w_density | (uintptr_t)w_class | (uintptr_t)w_color1 | (uintptr_t)w_color2) & 15) ||
        (((uintptr_t)c1_out) & 15)) return NRF_E_UNSUPPORTED;
    constexpr FieldSmem L = field_smem(1);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_field_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total); attr_set = true; }
    const uint32_t ntiles = ceil_div_u32(M, 128);
    const uint32_t grid = (uint32_t)min((uint64_t)ntiles, (uint64_t)field_sm_count() * 2);
    k_field_fwd_tc<<<grid, TC_THREADS, L.total, (cudaStream_t)stream>>>((const __half*)enc_d, (const __half*)enc_c, (const __half*)w_density,
                                                                       (const __half*)w_class, (const __half*)w_color1, (const __half*)w_color2, M,
                                                                       n_classes, sigmas, rgbs, ld_rgbs, (__half*)c1_out, M_dev);
    return nrf_check_launch();
}
