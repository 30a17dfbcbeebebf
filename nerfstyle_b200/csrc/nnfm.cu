// Segment-wise nearest-neighbour feature matching (loss.py:32-36 cosine_dists, :199-214 mask + amin).
//
// The reference materialises the N1 x N2 cosine-distance matrix (11 844 x 15 876 at room size), overwrites
// the entries whose (image class, style cluster) pair is not matched with +inf in a python loop over the
// classes, and takes a row-wise amin.  Here it is ONE tensor-core kernel: a 128x128-tile GEMM over the
// pre-normalised f16 features with the class mask and a running row arg-max of the similarity fused into the
// epilogue (min distance = 1 - max similarity).  The matrix never exists in memory.
#include "common.cuh"

#define NN_BM 128
#define NN_BN 128
#define NN_BK 32
#define NN_LD (NN_BK + 8)
#define NN_THREADS 256

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void ldsm_x4(uint32_t r[4], const __half* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816_nn(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// order-preserving map float -> uint32
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
// key = (ordered similarity << 32) | (~column): max key = max similarity, ties -> smallest column
__device__ __forceinline__ unsigned long long make_key(float sim, uint32_t col) {
    return ((unsigned long long)f2ord(sim) << 32) | (unsigned long long)(0xFFFFFFFFu - col);
}

__global__ void k_nnfm_prepare(const int32_t* __restrict__ a_label, const int32_t* __restrict__ match, uint32_t n_class,
                               uint32_t N1, int32_t* __restrict__ row_req, unsigned long long* __restrict__ best) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N1) return;
    int32_t req = -1;
    if (a_label && match) {
        const int32_t c = a_label[i];
        if (c >= 0 && (uint32_t)c < n_class) req = match[c];
    }
    row_req[i] = req;
    best[i] = 0ull;
}

__global__ void __launch_bounds__(NN_THREADS)
k_nnfm_gemm(const __half* __restrict__ A, const __half* __restrict__ Bm, uint32_t N1, uint32_t N2, uint32_t K,
            const int32_t* __restrict__ row_req, const int32_t* __restrict__ b_label, uint32_t n_tiles_per_split,
            unsigned long long* __restrict__ best) {
    __shared__ __align__(16) __half As[2][NN_BM][NN_LD];
    __shared__ __align__(16) __half Bs[2][NN_BN][NN_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_m = warp >> 2, warp_n = warp & 3;       // 2 x 4 warps, each 64 x 32
    const int g = lane >> 2, t = lane & 3;
    const uint32_t m0 = blockIdx.x * NN_BM;
    const uint32_t n_tiles = (N2 + NN_BN - 1) / NN_BN;
    const uint32_t tile_begin = blockIdx.y * n_tiles_per_split;
    const uint32_t tile_end = min(n_tiles, tile_begin + n_tiles_per_split);
    const uint32_t nk = (K + NN_BK - 1) / NN_BK;

    // per-thread rows: m-tile i (0..3), half h (0..1) -> row m0 + warp_m*64 + i*16 + g + 8h
    int32_t req[8];
    unsigned long long bestk[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t r = m0 + warp_m * 64 + (i >> 1) * 16 + g + (i & 1) * 8;
        req[i] = (r < N1) ? row_req[r] : -2;
        bestk[i] = 0ull;
    }

    // global -> shared staging: 128 rows x 32 halfs = 512 chunks of 16 B per operand, 2 per thread
    auto stage = [&](int buf, uint32_t n0, uint32_t kb) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int chunk = tid + c * NN_THREADS;
            const int row = chunk >> 2, kc = (chunk & 3) * 8;
            const uint32_t k = kb * NN_BK + kc;
            const bool kin = k < K;
            const uint32_t ra = m0 + row, rb = n0 + row;
            cp_async16(&As[buf][row][kc], A + (size_t)min(ra, N1 - 1) * K + (kin ? k : 0), kin && ra < N1);
            cp_async16(&Bs[buf][row][kc], Bm + (size_t)min(rb, N2 - 1) * K + (kin ? k : 0), kin && rb < N2);
        }
        cp_async_commit();
    };

    for (uint32_t tile = tile_begin; tile < tile_end; tile++) {
        const uint32_t n0 = tile * NN_BN;
        float acc[4][4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0f;
        stage(0, n0, 0);
        for (uint32_t kb = 0; kb < nk; kb++) {
            const int buf = kb & 1;
            if (kb + 1 < nk) { stage(buf ^ 1, n0, kb + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
            __syncthreads();
#pragma unroll
            for (int ks = 0; ks < NN_BK / 16; ks++) {
                uint32_t af[4][4], bf[2][4];
#pragma unroll
                for (int i = 0; i < 4; i++)
                    ldsm_x4(af[i], &As[buf][warp_m * 64 + i * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][ks * 16 + ((lane >> 4) & 1) * 8]);
#pragma unroll
                for (int j = 0; j < 2; j++)
                    ldsm_x4(bf[j], &Bs[buf][warp_n * 32 + j * 16 + (lane & 7) + ((lane >> 4) & 1) * 8][ks * 16 + ((lane >> 3) & 1) * 8]);
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) mma16816_nn(acc[i][j], af[i], bf[j >> 1][(j & 1) * 2], bf[j >> 1][(j & 1) * 2 + 1]);
            }
            __syncthreads();
        }
        // fused epilogue: class mask + running arg-max of the similarity
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t c0 = n0 + warp_n * 32 + j * 8 + 2 * t;
            const int32_t l0 = (c0 < N2 && b_label) ? __ldg(b_label + c0) : -1;
            const int32_t l1 = (c0 + 1 < N2 && b_label) ? __ldg(b_label + c0 + 1) : -1;
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int ri = i * 2 + h;
                    if (req[ri] == -2) continue;
                    if (c0 < N2 && (req[ri] < 0 || l0 == req[ri])) {
                        const unsigned long long k0 = make_key(acc[i][j][2 * h], c0);
                        if (k0 > bestk[ri]) bestk[ri] = k0;
                    }
                    if (c0 + 1 < N2 && (req[ri] < 0 || l1 == req[ri])) {
                        const unsigned long long k1 = make_key(acc[i][j][2 * h + 1], c0 + 1);
                        if (k1 > bestk[ri]) bestk[ri] = k1;
                    }
                }
            }
        }
    }
    // quad reduction, then one atomic per (row, warp)
#pragma unroll
    for (int i = 0; i < 8; i++) {
        unsigned long long k = bestk[i];
        const unsigned long long o1 = __shfl_xor_sync(NRF_FULL_MASK, k, 1);
        if (o1 > k) k = o1;
        const unsigned long long o2 = __shfl_xor_sync(NRF_FULL_MASK, k, 2);
        if (o2 > k) k = o2;
        const uint32_t r = m0 + warp_m * 64 + (i >> 1) * 16 + g + (i & 1) * 8;
        if (t == 0 && r < N1 && k != 0ull) atomicMax(best + r, k);
    }
}

__global__ void k_nnfm_finish(const unsigned long long* __restrict__ best, uint32_t N1, float* __restrict__ min_dist,
                              int32_t* __restrict__ argmin) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N1) return;
    const unsigned long long k = best[i];
    if (k == 0ull) { min_dist[i] = __int_as_float(0x7f800000); if (argmin) argmin[i] = -1; return; }
    min_dist[i] = 1.0f - ord2f((uint32_t)(k >> 32));
    if (argmin) argmin[i] = (int32_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
}

// nnfm_tc.cu: the tcgen05 / TMEM implementation (default); the mma.sync kernel above is the selectable second one
uint64_t nrf_nnfm_tc_pack_bytes(uint32_t N1, uint32_t N2, uint32_t K);
int nrf_nnfm_tc_gemm(const __half* a, const __half* b, uint32_t N1, uint32_t N2, uint32_t K, const int32_t* row_req,
                     const int32_t* b_label, unsigned long long* best, void* packed, cudaStream_t s);
static int g_nnfm_mode = 0;     // 0 = tcgen05 (nnfm_tc.cu), 1 = mma.sync
NRF_EXPORT void nrf_nnfm_set_mode(int mode) { g_nnfm_mode = mode; }

static uint64_t nnfm_small_bytes(uint32_t N1) {
    return (((uint64_t)N1 * (sizeof(unsigned long long) + sizeof(int32_t)) + 64) + 255) & ~255ull;
}
NRF_EXPORT uint64_t nrf_nnfm_scratch_bytes(uint32_t N1, uint32_t N2, uint32_t K) {
    return nnfm_small_bytes(N1) + nrf_nnfm_tc_pack_bytes(N1, N2, K);
}

NRF_EXPORT int nrf_nnfm_forward(const void* a_f16, const void* b_f16, uint32_t N1, uint32_t N2, uint32_t K,
                                const int32_t* a_label, const int32_t* b_label, const int32_t* match, uint32_t n_class,
                                float* min_dist, int32_t* argmin, void* scratch, void* stream) {
    if (N1 == 0) return NRF_OK;
    if (!a_f16 || !b_f16 || !min_dist || !scratch) return NRF_E_INVALID;
    if (N2 == 0 || K == 0 || (K % 8) != 0) return NRF_E_UNSUPPORTED;
    if ((((uintptr_t)a_f16) & 15) || (((uintptr_t)b_f16) & 15) || (((uintptr_t)scratch) & 7)) return NRF_E_INVALID;
    if (match && (!a_label || !b_label)) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long* best = (unsigned long long*)scratch;
    int32_t* row_req = (int32_t*)(best + N1);
    k_nnfm_prepare<<<ceil_div_u32(N1, 256), 256, 0, s>>>(a_label, match, n_class, N1, row_req, best);
    if (g_nnfm_mode == 0) {
        void* packed = (void*)((((uintptr_t)scratch + nnfm_small_bytes(N1)) + 127) & ~(uintptr_t)127);
        const int rc = nrf_nnfm_tc_gemm((const __half*)a_f16, (const __half*)b_f16, N1, N2, K, row_req, match ? b_label : nullptr, best,
                                        packed, s);
        if (rc != NRF_OK) return rc;
        k_nnfm_finish<<<ceil_div_u32(N1, 256), 256, 0, s>>>(best, N1, min_dist, argmin);
        return nrf_check_launch();
    }
    const uint32_t m_tiles = ceil_div_u32(N1, NN_BM), n_tiles = ceil_div_u32(N2, NN_BN);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // split the N range so that the grid has >= 2 CTAs per SM when possible
    uint32_t splits = 1;
    while (m_tiles * splits < (uint32_t)(2 * sms) && splits < n_tiles) splits++;
    const uint32_t per = ceil_div_u32(n_tiles, splits);
    splits = ceil_div_u32(n_tiles, per);
    k_nnfm_gemm<<<dim3(m_tiles, splits), NN_THREADS, 0, s>>>((const __half*)a_f16, (const __half*)b_f16, N1, N2, K, row_req,
                                                           match ? b_label : nullptr, per, best);
    k_nnfm_finish<<<ceil_div_u32(N1, 256), 256, 0, s>>>(best, N1, min_dist, argmin);
    return nrf_check_launch();
}
