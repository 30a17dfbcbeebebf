// Nearest-neighbour feature matching on the 5th-generation tensor cores (tcgen05 + TMEM + bulk async copies).
//
// Same contract as nnfm.cu (loss.py:32-36 cosine_dists, :199-214 mask + amin): min_j (1 - a_i . b_j) over the allowed
// columns, with the N1 x N2 matrix never materialised.  Structure of one CTA (192 threads, 1 CTA per SM):
//   warp 0 (one lane)  producer : cp.async.bulk global -> shared, one 16 KB A block + one 32 KB B block per K-stage,
//                                 completion by mbarrier transaction bytes (4-stage ring, 48 KB per stage)
//   warp 1 (one lane)  issuer   : 4 x tcgen05.mma (M=128, N=256, K=16) per stage into one of TWO 256-column TMEM
//                                 accumulators; tcgen05.commit frees the stage / publishes the accumulator
//   warps 2-5          epilogue : each thread owns one row of the 128 x 256 tile (its TMEM lane), applies the
//                                 (image class -> style cluster) mask and keeps a running arg-max of the similarity
//                                 across all N tiles of the CTA in registers; one 64-bit atomicMax per row at the end.
// The epilogue of tile j overlaps the MMAs of tile j+1 (double-buffered TMEM).
// Operands are pre-packed once per call into the SWIZZLE_NONE "chunked" tile layout of tc05.cuh
// ([row tile][k block of 64][k chunk of 8][row][8 halfs]) so that every stage operand is ONE contiguous block in global
// memory: a single bulk copy each, no tensor map, and the same K-major descriptors the MLP kernels use.
#include "common.cuh"
#include "tc05.cuh"

namespace {

constexpr int TBM = 128, TBN = 256, TBK = 64, STAGES = 4;
constexpr uint32_t A_STAGE = TBM * TBK * 2, B_STAGE = TBN * TBK * 2;       // 16 KB, 32 KB
constexpr uint32_t CH_A = TBM * 16, CH_B = TBN * 16;                         // chunk strides of the packed blocks
constexpr int TC_NN_THREADS = 192;

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc05::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc05::smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(tc05::smem_u32(bar)) : "memory");
}

// order-preserving map float -> uint32 and the packed (similarity, ~column) key shared with nnfm.cu
__device__ __forceinline__ uint32_t f2ord_tc(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long make_key_tc(float sim, uint32_t col) {
    return ((unsigned long long)f2ord_tc(sim) << 32) | (unsigned long long)(0xFFFFFFFFu - col);
}

// row-major [N, K] f16 -> packed tiles of R rows: dst[((tile * nkb + kb) * 8 + c) * R * 8 + r * 8 + e], zero padded
__global__ void k_nnfm_pack(const __half* __restrict__ src, uint32_t N, uint32_t K, uint32_t R, uint32_t n_tiles, uint32_t nkb,
                            __half* __restrict__ dst) {
    const uint64_t total = (uint64_t)n_tiles * R * nkb * 8;      // 16-byte pieces
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        // consecutive threads -> consecutive rows of one k chunk: coalesced 16-byte writes
        const uint32_t r = (uint32_t)(i % R);
        uint64_t q = i / R;
        const uint32_t c = (uint32_t)(q % 8); q /= 8;
        const uint32_t kb = (uint32_t)(q % nkb);
        const uint32_t tile = (uint32_t)(q / nkb);
        const uint32_t row = tile * R + r, k = kb * TBK + c * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row < N && k < K) v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)row * K + k));
        reinterpret_cast<uint4*>(dst)[i] = v;
    }
}

__global__ void __launch_bounds__(TC_NN_THREADS, 1)
k_nnfm_gemm_tc(const __half* __restrict__ Ap, const __half* __restrict__ Bp, uint32_t N1, uint32_t N2, uint32_t nkb,
               const int32_t* __restrict__ row_req, const int32_t* __restrict__ b_label, uint32_t n_tiles, uint32_t tiles_per_split,
               unsigned long long* __restrict__ best) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) int32_t s_label[2][TBN];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sA = smem;                          // [STAGES][A_STAGE]
    uint8_t* sB = smem + STAGES * A_STAGE;       // [STAGES][B_STAGE]
    const uint32_t m_tile = blockIdx.x;
    const uint32_t t_begin = blockIdx.y * tiles_per_split;
    const uint32_t t_end = min(n_tiles, t_begin + tiles_per_split);

    if (warp == 0) { tc05::tmem_alloc(&tmem_slot, 512); tc05::tmem_relinquish(); }
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { tc05::mbar_init(&bar_full[s], 1); tc05::mbar_init(&bar_empty[s], 1); }
        for (int a = 0; a < 2; a++) { tc05::mbar_init(&bar_tfull[a], 1); tc05::mbar_init(&bar_tempty[a], 128); }
        tc05::fence_mbar_init();
    }
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tbase = tmem_slot;

    if (warp == 0) {
        // ================================================================ producer
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t t = t_begin; t < t_end; t++) {
                for (uint32_t kb = 0; kb < nkb; kb++, it++) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    tc05::mbar_wait(&bar_empty[s], ph ^ 1);
                    mbar_expect_tx(&bar_full[s], A_STAGE + B_STAGE);
                    bulk_g2s(sA + s * A_STAGE, Ap + ((size_t)m_tile * nkb + kb) * (A_STAGE / 2), A_STAGE, &bar_full[s]);
                    bulk_g2s(sB + s * B_STAGE, Bp + ((size_t)t * nkb + kb) * (B_STAGE / 2), B_STAGE, &bar_full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        if (lane == 0) {
            constexpr uint32_t IDESC = tc05::idesc_f16(TBM, TBN, false, false);
            uint32_t it = 0, j = 0;
            for (uint32_t t = t_begin; t < t_end; t++, j++) {
                const uint32_t acc = j & 1;
                tc05::mbar_wait(&bar_tempty[acc], ((j >> 1) & 1) ^ 1);
                tc05::fence_after_sync();
                for (uint32_t kb = 0; kb < nkb; kb++, it++) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    tc05::mbar_wait(&bar_full[s], ph);
                    tc05::fence_after_sync();
                    const uint64_t da = tc05::desc_kmajor(tc05::smem_u32(sA + s * A_STAGE), CH_A);
                    const uint64_t db = tc05::desc_kmajor(tc05::smem_u32(sB + s * B_STAGE), CH_B);
#pragma unroll
                    for (int k = 0; k < TBK / 16; k++)
                        tc05::mma_f16(tbase + acc * TBN, da + (uint64_t)((k * 2 * CH_A) >> 4), db + (uint64_t)((k * 2 * CH_B) >> 4), IDESC,
                                      (kb > 0 || k > 0) ? 1u : 0u);
                    tc05::mma_commit(&bar_empty[s]);
                }
                tc05::mma_commit(&bar_tfull[acc]);
            }
        }
        __syncwarp();
    } else {
        // ================================================================ epilogue (warps 2-5: TMEM lanes 32 * (warp % 4))
        const uint32_t q = warp & 3;
        const uint32_t r_local = q * 32 + lane;
        const uint32_t row = m_tile * TBM + r_local;
        const int etid = tid - 64;
        const bool masked = b_label != nullptr;
        const int32_t req = (row < N1) ? (masked ? __ldg(row_req + row) : -1) : -2;
        float best_s = -3.0e38f;
        int32_t best_c = -1;
        uint32_t j = 0;
        for (uint32_t t = t_begin; t < t_end; t++, j++) {
            const uint32_t acc = j & 1, n0 = t * TBN;
            if (masked) {
                // stage this tile's column labels (double buffered; the named barrier below orders them)
                for (int c = etid; c < TBN; c += 128) s_label[acc][c] = (n0 + c < N2) ? __ldg(b_label + n0 + c) : -3;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            tc05::mbar_wait(&bar_tfull[acc], (j >> 1) & 1);
            tc05::fence_after_sync();
            const uint32_t ncols = min((uint32_t)TBN, N2 - n0);
#pragma unroll 1
            for (uint32_t g = 0; g < TBN / 32; g++) {
                if (g * 32 >= ncols) break;
                uint32_t v[32];
                tc05::tmem_ld32(tbase + ((q * 32) << 16) + acc * TBN + g * 32, v);
                tc05::tmem_ld_wait();
                if (req != -2) {
                    const bool full_chunk = (g + 1) * 32 <= ncols;
                    if (req < 0 && full_chunk) {
#pragma unroll
                        for (int e = 0; e < 32; e++) {
                            const float sv = __uint_as_float(v[e]);
                            if (sv > best_s) { best_s = sv; best_c = (int32_t)(n0 + g * 32 + e); }
                        }
                    } else {
#pragma unroll
                        for (int e4 = 0; e4 < 8; e4++) {
                            int4 lb = make_int4(req, req, req, req);
                            if (masked) lb = *reinterpret_cast<const int4*>(&s_label[acc][g * 32 + e4 * 4]);
                            const int32_t lbs[4] = {lb.x, lb.y, lb.z, lb.w};
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const uint32_t c = g * 32 + e4 * 4 + e;
                                const float sv = __uint_as_float(v[e4 * 4 + e]);
                                const bool ok = (c < ncols) && (req < 0 || lbs[e] == req);
                                if (ok && sv > best_s) { best_s = sv; best_c = (int32_t)(n0 + c); }
                            }
                        }
                    }
                }
            }
            tc05::fence_before_sync();
            tc05::mbar_arrive(&bar_tempty[acc]);
        }
        if (req != -2 && best_c >= 0) atomicMax(best + row, make_key_tc(best_s, (uint32_t)best_c));
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tbase, 512);
}

}  // namespace

uint64_t nrf_nnfm_tc_pack_bytes(uint32_t N1, uint32_t N2, uint32_t K) {
    const uint64_t kp = (uint64_t)ceil_div_u32(K, TBK) * TBK;
    return ((uint64_t)ceil_div_u32(N1, TBM) * TBM + (uint64_t)ceil_div_u32(N2, TBN) * TBN) * kp * 2 + 256;
}

// packed = 128-byte aligned scratch of nrf_nnfm_tc_pack_bytes(); row_req / best prepared by the caller (nnfm.cu)
int nrf_nnfm_tc_gemm(const __half* a, const __half* b, uint32_t N1, uint32_t N2, uint32_t K, const int32_t* row_req,
                     const int32_t* b_label, unsigned long long* best, void* packed, cudaStream_t s) {
    const uint32_t nkb = ceil_div_u32(K, TBK);
    const uint32_t m_tiles = ceil_div_u32(N1, TBM), n_tiles = ceil_div_u32(N2, TBN);
    __half* Ap = (__half*)packed;
    __half* Bp = Ap + (size_t)m_tiles * TBM * nkb * TBK;
    k_nnfm_pack<<<1184, 256, 0, s>>>(a, N1, K, TBM, m_tiles, nkb, Ap);
    k_nnfm_pack<<<1184, 256, 0, s>>>(b, N2, K, TBN, n_tiles, nkb, Bp);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // split the N range so that the grid fills whole waves of one CTA per SM (each CTA keeps >= 2 column tiles when possible)
    uint32_t best_splits = 1;
    double best_eff = 0.0;
    for (uint32_t sp = 1; sp <= n_tiles; sp++) {
        const uint32_t per = ceil_div_u32(n_tiles, sp);
        const uint32_t real = ceil_div_u32(n_tiles, per);
        if (real != sp) continue;
        if (per < 2 && sp > 1 && n_tiles >= 2) continue;
        const uint64_t ctas = (uint64_t)m_tiles * sp;
        const uint64_t waves = (ctas + sms - 1) / sms;
        // work per wave is proportional to `per`; total time ~ waves * per (+ a small per-CTA overhead)
        const double cost = (double)waves * ((double)per + 0.25);
        const double eff = 1.0 / cost;
        if (eff > best_eff) { best_eff = eff; best_splits = sp; }
    }
    const uint32_t per = ceil_div_u32(n_tiles, best_splits);
    const uint32_t splits = ceil_div_u32(n_tiles, per);
    const size_t smem = (size_t)STAGES * (A_STAGE + B_STAGE);
    cudaFuncSetAttribute(k_nnfm_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_nnfm_gemm_tc<<<dim3(m_tiles, splits), TC_NN_THREADS, smem, s>>>(Ap, Bp, N1, N2, nkb, row_req, b_label, n_tiles, per, best);
    return nrf_check_launch();
}
