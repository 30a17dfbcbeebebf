// Stand-alone probe of the tcgen05 conventions used by nerfstyle_b200/csrc/tc05.cuh (run on a B200):
//   * SWIZZLE_NONE "chunked" tiles as K-major and as MN-major operands (descriptor LBO/SBO roles)
//   * M=128 and M=64 accumulator lane mapping in TMEM, N=16/32/64
//   * micro timings: MMA issue->commit->wait round trip, tcgen05.ld throughput
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tc_probe tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../nerfstyle_b200/csrc/tc05.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct ProbeArgs {
    uint32_t a_bytes, b_bytes;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    uint32_t idesc, nk, a_step, b_step, ncols;
};

__global__ void __launch_bounds__(128) k_probe(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, ProbeArgs p,
                                               float* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* sa = smem;
    uint8_t* sb = smem + ((p.a_bytes + 127u) & ~127u);
    for (uint32_t i = threadIdx.x; i < p.a_bytes / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (uint32_t i = threadIdx.x; i < p.b_bytes / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tc05::tmem_alloc(&tmem_slot, 128); tc05::tmem_relinquish(); }
    if (threadIdx.x == 0) { tc05::mbar_init(&bar, 1); tc05::fence_mbar_init(); }
    tc05::fence_async_smem();
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if (threadIdx.x == 0) {
        for (uint32_t k = 0; k < p.nk; k++) {
            const uint64_t ad = tc05::smem_desc(tc05::smem_u32(sa) + k * p.a_step, p.a_lbo, p.a_sbo);
            const uint64_t bd = tc05::smem_desc(tc05::smem_u32(sb) + k * p.b_step, p.b_lbo, p.b_sbo);
            tc05::mma_f16(tbase, ad, bd, p.idesc, k > 0 ? 1u : 0u);
        }
        tc05::mma_commit(&bar);
    }
    tc05::mbar_wait(&bar, 0);
    tc05::fence_after_sync();
    for (uint32_t c0 = 0; c0 < p.ncols; c0 += 8) {
        uint32_t v[8];
        tc05::tmem_ld8(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
        tc05::tmem_ld_wait();
        for (int j = 0; j < 8; j++) out[(size_t)(warp * 32 + lane) * p.ncols + c0 + j] = __uint_as_float(v[j]);
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tbase, 128);
}

// chunked image of a logical [R][C] matrix (see tc05.cuh)
static std::vector<uint8_t> chunked(const std::vector<float>& X, int R, int C, int CH) {
    std::vector<uint8_t> img((size_t)(C / 8) * CH, 0);
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) {
            __half h = __float2half(X[(size_t)r * C + c]);
            memcpy(&img[(size_t)(c / 8) * CH + r * 16 + (c % 8) * 2], &h, 2);
        }
    return img;
}
static std::vector<float> rnd(int n, unsigned seed) {
    std::vector<float> v(n);
    srand(seed);
    for (auto& x : v) x = __half2float(__float2half(((rand() % 2001) - 1000) / 1000.0f));
    return v;
}

static int run_case(const char* name, const std::vector<uint8_t>& ai, const std::vector<uint8_t>& bi, ProbeArgs p,
                    const std::vector<float>& expect /* [M][ncols] logical rows */, int M, bool m64_layout) {
    uint8_t *da, *db; float* dout;
    p.a_bytes = (uint32_t)ai.size(); p.b_bytes = (uint32_t)bi.size();
    CK(cudaMalloc(&da, ai.size())); CK(cudaMalloc(&db, bi.size())); CK(cudaMalloc(&dout, 128 * p.ncols * 4));
    CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0, 128 * p.ncols * 4));
    const size_t smem = ((ai.size() + 127) & ~127ull) + bi.size() + 128;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_probe<<<1, 128, smem>>>(da, db, p, dout);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * p.ncols);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < M; m++) {
        const int lane = m64_layout ? (m % 16) + 32 * (m / 16) : m;
        for (uint32_t c = 0; c < p.ncols; c++) {
            const double e = fabs(out[(size_t)lane * p.ncols + c] - expect[(size_t)m * p.ncols + c]);
            if (e > maxerr) maxerr = e;
            if (e > 2e-3) bad++;
        }
    }
    printf("%-44s max|err| = %.3g  bad = %d  -> %s\n", name, maxerr, bad, bad ? "FAIL" : "ok");
    if (bad) {
        printf("   lane0 got:");  for (int c = 0; c < 8; c++) printf(" %8.4f", out[c]);
        printf("\n   row0 want:"); for (int c = 0; c < 8; c++) printf(" %8.4f", expect[c]);
        printf("\n   lane1 got:");  for (int c = 0; c < 8; c++) printf(" %8.4f", out[p.ncols + c]);
        printf("\n   row1 want:"); for (int c = 0; c < 8; c++) printf(" %8.4f", expect[p.ncols + c]);
        printf("\n   lane16 got:"); for (int c = 0; c < 8; c++) printf(" %8.4f", out[16 * p.ncols + c]);
        printf("\n   lane32 got:"); for (int c = 0; c < 8; c++) printf(" %8.4f", out[32 * p.ncols + c]);
        printf("\n   row16 want:"); for (int c = 0; c < 8; c++) printf(" %8.4f", expect[16 * p.ncols + c]);
        printf("\n");
    }
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return bad;
}

// ---------------------------------------------------------------------------------------------- timing
__global__ void __launch_bounds__(128) k_time(long long* out, int iters, int mode) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (uint32_t i = threadIdx.x; i < 40960 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tc05::tmem_alloc(&tmem_slot, 128); tc05::tmem_relinquish(); }
    if (threadIdx.x == 0) { tc05::mbar_init(&bar, 1); tc05::fence_mbar_init(); }
    tc05::fence_async_smem();
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t sa = tc05::smem_u32(smem), sb = sa + 16384;
    const uint32_t idesc = tc05::idesc_f16(128, 64, false, false);
    uint32_t acc = 0;
    long long t0 = clock64();
    if (mode == 0) {          // round trip: 2 MMAs (K=32) -> commit -> everyone waits
        for (int it = 0; it < iters; it++) {
            if (threadIdx.x == 0) {
                tc05::mma_f16(tbase, tc05::desc_kmajor(sa, 2048), tc05::desc_kmajor(sb, 1024), idesc, 0);
                tc05::mma_f16(tbase, tc05::desc_kmajor(sa + 4096, 2048), tc05::desc_kmajor(sb + 2048, 1024), idesc, 1);
                tc05::mma_commit(&bar);
            }
            tc05::mbar_wait(&bar, it & 1);
            tc05::fence_after_sync();
            __syncthreads();      // parity waits alias after two phases: keep everyone within one phase
        }
    } else if (mode == 1) {   // tcgen05.ld of 64 columns per thread (128 lanes x 64 cols x 4 B = 32 KB per CTA per iteration)
        for (int it = 0; it < iters; it++) {
            uint32_t v[32], w[32];
            tc05::tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16), v);
            tc05::tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + 32, w);
            tc05::tmem_ld_wait();
            for (int j = 0; j < 32; j++) acc += v[j] ^ w[j];
        }
    } else {                  // round trip + 64-column read + smem write-back of a f16 tile + syncthreads (one MLP layer hop)
        for (int it = 0; it < iters; it++) {
            if (threadIdx.x == 0) {
                tc05::mma_f16(tbase, tc05::desc_kmajor(sa, 2048), tc05::desc_kmajor(sb, 1024), idesc, 0);
                tc05::mma_f16(tbase, tc05::desc_kmajor(sa + 4096, 2048), tc05::desc_kmajor(sb + 2048, 1024), idesc, 1);
                tc05::mma_commit(&bar);
            }
            tc05::mbar_wait(&bar, it & 1);
            tc05::fence_after_sync();
            uint32_t v[32], w[32];
            tc05::tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16), v);
            tc05::tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + 32, w);
            tc05::tmem_ld_wait();
            uint4* dst = reinterpret_cast<uint4*>(smem + 20480 + threadIdx.x * 16);
            for (int j = 0; j < 4; j++) dst[j * 128] = make_uint4(v[8 * j] & 0, v[8 * j + 1] & 0, w[8 * j] & 0, w[8 * j + 1] & 0);
            tc05::fence_async_smem();
            tc05::fence_before_sync();
            __syncthreads();
            tc05::fence_after_sync();
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = acc; }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tbase, 128);
}


// ---------------------------------------------------------------------------------------------- MMA throughput by shape / major
__global__ void __launch_bounds__(128) k_mma_rate(long long* out, int n_mma, uint32_t idesc, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo,
                                                  uint32_t b_sbo) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (uint32_t i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tc05::tmem_alloc(&tmem_slot, 128); tc05::tmem_relinquish(); }
    if (threadIdx.x == 0) { tc05::mbar_init(&bar, 1); tc05::fence_mbar_init(); }
    tc05::fence_async_smem();
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint64_t ad = tc05::smem_desc(tc05::smem_u32(smem), a_lbo, a_sbo), bd = tc05::smem_desc(tc05::smem_u32(smem) + 32768, b_lbo, b_sbo);
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_mma; i++) tc05::mma_f16(tbase, ad, bd, idesc, 1);
        tc05::mma_commit(&bar);
    }
    tc05::mbar_wait(&bar, 0);
    tc05::fence_after_sync();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tbase, 128);
}

static void mma_rate(const char* name, int M, int N, bool a_mn, bool b_mn, uint32_t cha, uint32_t chb) {
    long long* d; CK(cudaMalloc(&d, 8 * 1024));
    CK(cudaFuncSetAttribute(k_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const uint32_t idesc = tc05::idesc_f16(M, N, a_mn, b_mn);
    const uint32_t a_lbo = a_mn ? 128 : cha, a_sbo = a_mn ? cha : 128, b_lbo = b_mn ? 128 : chb, b_sbo = b_mn ? chb : 128;
    long long r[2];
    for (int pass = 0; pass < 2; pass++) {
        const int n = pass == 0 ? 16 : 272;
        k_mma_rate<<<1, 128, 65536>>>(d, n, idesc, a_lbo, a_sbo, b_lbo, b_sbo);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&r[pass], d, 8, cudaMemcpyDeviceToHost));
    }
    printf("mma rate: %-40s %6.1f cycles/MMA  (16 MMAs + commit + wait: %lld cycles)\n", name, (double)(r[1] - r[0]) / 256.0, r[0]);
    cudaFree(d);
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    int dev = 0; cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    printf("device: %s sm_%d%d, %d SMs\n", pr.name, pr.major, pr.minor, pr.multiProcessorCount);
    int fails = 0;
    // ---- case 1: D[128x64] = X[128x32] * W[64x32]^T, both K-major
    {
        auto X = rnd(128 * 32, 1), W = rnd(64 * 32, 2);
        std::vector<float> E(128 * 64, 0.f);
        for (int m = 0; m < 128; m++) for (int n = 0; n < 64; n++) { float s = 0; for (int k = 0; k < 32; k++) s += X[m * 32 + k] * W[n * 32 + k]; E[m * 64 + n] = s; }
        ProbeArgs p{}; p.a_lbo = 2048; p.a_sbo = 128; p.b_lbo = 1024; p.b_sbo = 128; p.idesc = tc05::idesc_f16(128, 64, false, false);
        p.nk = 2; p.a_step = 2 * 2048; p.b_step = 2 * 1024; p.ncols = 64;
        fails += run_case("1 K-major A, K-major B, M128 N64 K32", chunked(X, 128, 32, 2048), chunked(W, 64, 32, 1024), p, E, 128, false);
        // same with padded chunk stride (CH = 2048+32)
        p.a_lbo = 2080; p.a_step = 2 * 2080;
        fails += run_case("1b same, A chunk stride 2080", chunked(X, 128, 32, 2080), chunked(W, 64, 32, 1024), p, E, 128, false);
    }
    // ---- case 2: D[128x32] = H[128x64] * W1[64x32]  (B MN-major: N = cols of W1, K = rows of W1)
    {
        auto H = rnd(128 * 64, 3), W = rnd(64 * 32, 4);
        std::vector<float> E(128 * 32, 0.f);
        for (int m = 0; m < 128; m++) for (int n = 0; n < 32; n++) { float s = 0; for (int k = 0; k < 64; k++) s += H[m * 64 + k] * W[k * 32 + n]; E[m * 32 + n] = s; }
        ProbeArgs p{}; p.a_lbo = 2048; p.a_sbo = 128; p.b_lbo = 128; p.b_sbo = 1024; p.idesc = tc05::idesc_f16(128, 32, false, true);
        p.nk = 4; p.a_step = 2 * 2048; p.b_step = 256; p.ncols = 32;
        fails += run_case("2 K-major A, MN-major B, M128 N32 K64", chunked(H, 128, 64, 2048), chunked(W, 64, 32, 1024), p, E, 128, false);
    }
    // ---- case 3: D[64x32] = P[128x64]^T * Q[128x32]  (A and B MN-major, K = 128 rows, M = 64)
    {
        auto P = rnd(128 * 64, 5), Q = rnd(128 * 32, 6);
        std::vector<float> E(64 * 32, 0.f);
        for (int m = 0; m < 64; m++) for (int n = 0; n < 32; n++) { float s = 0; for (int k = 0; k < 128; k++) s += P[k * 64 + m] * Q[k * 32 + n]; E[m * 32 + n] = s; }
        ProbeArgs p{}; p.a_lbo = 128; p.a_sbo = 2048; p.b_lbo = 128; p.b_sbo = 2048; p.idesc = tc05::idesc_f16(64, 32, true, true);
        p.nk = 8; p.a_step = 256; p.b_step = 256; p.ncols = 32;
        fails += run_case("3 MN-major A and B, M64 N32 K128", chunked(P, 128, 64, 2048), chunked(Q, 128, 32, 2048), p, E, 64, true);
    }
    // ---- case 4: D[128x16] = H[128x64] * Wo[16x64]^T (K-major, N=16)
    {
        auto H = rnd(128 * 64, 7), W = rnd(16 * 64, 8);
        std::vector<float> E(128 * 16, 0.f);
        for (int m = 0; m < 128; m++) for (int n = 0; n < 16; n++) { float s = 0; for (int k = 0; k < 64; k++) s += H[m * 64 + k] * W[n * 64 + k]; E[m * 16 + n] = s; }
        ProbeArgs p{}; p.a_lbo = 2048; p.a_sbo = 128; p.b_lbo = 256; p.b_sbo = 128; p.idesc = tc05::idesc_f16(128, 16, false, false);
        p.nk = 4; p.a_step = 2 * 2048; p.b_step = 2 * 256; p.ncols = 16;
        fails += run_case("4 K-major A,B M128 N16 K64", chunked(H, 128, 64, 2048), chunked(W, 16, 64, 256), p, E, 128, false);
    }
    // ---- case 5: D[128x64] = dZ[128x16] * Wo[16x64]  (B MN-major, K = 16 rows of Wo)
    {
        auto Z = rnd(128 * 16, 9), W = rnd(16 * 64, 10);
        std::vector<float> E(128 * 64, 0.f);
        for (int m = 0; m < 128; m++) for (int n = 0; n < 64; n++) { float s = 0; for (int k = 0; k < 16; k++) s += Z[m * 16 + k] * W[k * 64 + n]; E[m * 64 + n] = s; }
        ProbeArgs p{}; p.a_lbo = 2048; p.a_sbo = 128; p.b_lbo = 128; p.b_sbo = 256; p.idesc = tc05::idesc_f16(128, 64, false, true);
        p.nk = 1; p.a_step = 0; p.b_step = 0; p.ncols = 64;
        fails += run_case("5 K-major A, MN-major B, M128 N64 K16", chunked(Z, 128, 16, 2048), chunked(W, 16, 64, 256), p, E, 128, false);
    }
    // ---- case 6: D[64x16] = H[128x64]^T * dZ[128x16]  (dWo^T; M64 N16 K128)
    {
        auto P = rnd(128 * 64, 11), Q = rnd(128 * 16, 12);
        std::vector<float> E(64 * 16, 0.f);
        for (int m = 0; m < 64; m++) for (int n = 0; n < 16; n++) { float s = 0; for (int k = 0; k < 128; k++) s += P[k * 64 + m] * Q[k * 16 + n]; E[m * 16 + n] = s; }
        ProbeArgs p{}; p.a_lbo = 128; p.a_sbo = 2048; p.b_lbo = 128; p.b_sbo = 2048; p.idesc = tc05::idesc_f16(64, 16, true, true);
        p.nk = 8; p.a_step = 256; p.b_step = 256; p.ncols = 16;
        fails += run_case("6 MN-major A and B, M64 N16 K128", chunked(P, 128, 64, 2048), chunked(Q, 128, 16, 2048), p, E, 64, true);
    }
    // ---- timings
    long long* dt; CK(cudaMalloc(&dt, 8 * 2 * 1024));
    CK(cudaFuncSetAttribute(k_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
    const char* names[3] = {"MMA round trip (2 MMA K32 N64 + commit + wait)", "tcgen05.ld 128 lanes x 64 cols", "layer hop (MMA + ld 64 cols + sts + bar)"};
    for (int mode = 0; mode < 3; mode++) {
        for (int grid : {1, 148, 148 * 4}) {
            const int iters = 2000;
            k_time<<<grid, 128, 40960>>>(dt, iters, mode);
            CK(cudaDeviceSynchronize());
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            k_time<<<grid, 128, 40960>>>(dt, iters, mode);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[2]; CK(cudaMemcpy(h, dt, 16, cudaMemcpyDeviceToHost));
            printf("time: %-50s grid %4d: %8.1f cycles/iter (block 0), kernel %.3f ms -> %.1f ns/iter\n", names[mode], grid, (double)h[0] / iters, ms, ms * 1e6 / iters);
        }
    }
    mma_rate("M128 N64 A:K  B:K", 128, 64, false, false, 2048, 1024);
    mma_rate("M128 N64 A:K  B:MN", 128, 64, false, true, 2048, 1024);
    mma_rate("M128 N32 A:K  B:MN (dX)", 128, 32, false, true, 2048, 1024);
    mma_rate("M128 N16 A:K  B:K  (Z)", 128, 16, false, false, 2048, 256);
    mma_rate("M64  N32 A:MN B:MN (dW1)", 64, 32, true, true, 2048, 2048);
    mma_rate("M64  N16 A:MN B:MN (dWo)", 64, 16, true, true, 2048, 2048);
    mma_rate("M64  N64 A:MN B:MN (dWh)", 64, 64, true, true, 2048, 2048);
    mma_rate("M64  N32 A:K  B:K", 64, 32, false, false, 2048, 1024);
    mma_rate("M64  N32 A:MN B:K", 64, 32, true, false, 2048, 1024);
    mma_rate("M64  N32 A:K  B:MN", 64, 32, false, true, 2048, 2048);
    mma_rate("M128 N32 A:MN B:MN", 128, 32, true, true, 2048, 2048);
    mma_rate("M128 N48 A:MN B:MN (dW1|dWo stacked)", 128, 48, true, true, 2048, 2048);
    mma_rate("M128 N64 A:MN B:MN", 128, 64, true, true, 2048, 2048);
    mma_rate("M128 N128 A:MN B:MN", 128, 128, true, true, 2048, 2048);
    mma_rate("M128 N128 A:K B:K", 128, 128, false, false, 2048, 2048);
    mma_rate("M128 N256 A:K B:K", 128, 256, false, false, 2048, 4096);
    printf(fails ? "PROBE FAILED (%d)\n" : "PROBE OK\n", fails);
    return fails ? 1 : 0;
}
