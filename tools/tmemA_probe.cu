// Probe (run on a B200): tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st), B in shared memory.
//   * layout check: thread t = row t = TMEM lane t stores its row as packed f16 pairs (column c holds k = 2c, 2c + 1)
//   * hop timing: accumulator -> f16 -> next MMA's A operand, through shared memory (STS + proxy fence) vs through TMEM
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tmemA_probe tools/tmemA_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../nerfstyle_b200/csrc/tc05.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                   "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
                   "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
                   "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int K = 64, N = 64;
constexpr uint32_t CHB = 1024;      // chunk stride of the 64-row B tile

// mode 0: A from TMEM; mode 1: A from shared memory (reference)
__global__ void __launch_bounds__(160) k_probe(const __half* __restrict__ A, const uint8_t* __restrict__ b_img, float* __restrict__ out, int mode,
                                               long long* __restrict__ cyc, int hops) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_ready, bar_done;
    __shared__ uint32_t tmem_slot;
    uint8_t* sb = smem;                      // B: 8 chunks x 1024
    uint8_t* sa = smem + 8 * CHB;            // A (smem mode): 8 chunks x 2048
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 8 * (int)CHB / 16; i += 160) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    if (warp == 0) { tc05::tmem_alloc(&tmem_slot, 256); tc05::tmem_relinquish(); }
    if (tid == 0) { tc05::mbar_init(&bar_ready, 128); tc05::mbar_init(&bar_done, 1); tc05::fence_mbar_init(); }
    tc05::fence_async_smem(); tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tacc = tmem_slot;
    const uint32_t T_D = 0, T_A = 128;       // accumulator columns, A-operand columns (32 columns = 64 halfs)
    constexpr uint32_t ID = tc05::idesc_f16(128, N, false, false);
    if (warp == 4) {
        if (lane == 0) {
            const uint64_t kB = tc05::desc_kmajor(tc05::smem_u32(sb), CHB), kA = tc05::desc_kmajor(tc05::smem_u32(sa), 2048);
            uint32_t ph = 0;
            for (int h = 0; h < hops; h++) {
                tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
                for (int k = 0; k < K / 16; k++) {
                    const uint64_t bd = kB + (uint64_t)((k * 2 * CHB) >> 4);
                    if (mode == 0) mma_f16_ts(tacc + T_D, tacc + T_A + 8 * k, bd, ID, k > 0);
                    else tc05::mma_f16(tacc + T_D, kA + (uint64_t)((k * 2 * 2048) >> 4), bd, ID, k > 0);
                }
                tc05::mma_commit(&bar_done);
            }
        }
        __syncwarp();
    } else {
        const uint32_t tl = tacc + ((uint32_t)(warp * 32) << 16);
        uint32_t v[32];
        const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + (size_t)tid * K);
        for (int c = 0; c < 32; c++) v[c] = arow[c];
        uint32_t phase = 0;
        const long long t0 = clock64();
        for (int h = 0; h < hops; h++) {
            if (mode == 0) {
                tmem_st32(tl + T_A, v);
                tmem_st_wait();
                tc05::fence_before_sync();
                tc05::mbar_arrive(&bar_ready);
            } else {
                for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(sa + c * 2048 + tid * 16) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                tc05::fence_async_smem(); tc05::fence_before_sync();
                tc05::mbar_arrive(&bar_ready);
            }
            tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
            if (h + 1 < hops) {
                // next hop's A = this hop's accumulator rounded to f16 and scaled down (keeps magnitudes bounded)
                for (int half = 0; half < 2; half++) {
                    uint32_t d[32];
                    tc05::tmem_ld32(tl + T_D + 32 * half, d);
                    tc05::tmem_ld_wait();
                    for (int q = 0; q < 16; q++) {
                        __half2 hh = __floats2half2_rn(__uint_as_float(d[2 * q]) * 0.125f, __uint_as_float(d[2 * q + 1]) * 0.125f);
                        v[16 * half + q] = *reinterpret_cast<uint32_t*>(&hh);
                    }
                }
            }
        }
        if (tid == 0) cyc[0] = clock64() - t0;
        for (int half = 0; half < 2; half++) {
            uint32_t d[32];
            tc05::tmem_ld32(tl + T_D + 32 * half, d);
            tc05::tmem_ld_wait();
            for (int j = 0; j < 32; j++) out[(size_t)tid * N + 32 * half + j] = __uint_as_float(d[j]);
        }
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tacc, 256);
}

int main() {
    std::vector<float> A(128 * K), B(N * K);
    srand(1);
    for (auto& x : A) x = __half2float(__float2half(((rand() % 2001) - 1000) / 1000.0f));
    for (auto& x : B) x = __half2float(__float2half(((rand() % 2001) - 1000) / 1000.0f));
    std::vector<__half> Ah(128 * K);
    for (size_t i = 0; i < A.size(); i++) Ah[i] = __float2half(A[i]);
    std::vector<uint8_t> bimg(8 * CHB, 0);
    for (int r = 0; r < N; r++) for (int c = 0; c < K; c++) { __half h = __float2half(B[r * K + c]); memcpy(&bimg[(c / 8) * CHB + r * 16 + (c % 8) * 2], &h, 2); }
    __half* dA; uint8_t* dB; float* dout; long long* dc;
    CK(cudaMalloc(&dA, Ah.size() * 2)); CK(cudaMalloc(&dB, bimg.size())); CK(cudaMalloc(&dout, 128 * N * 4)); CK(cudaMalloc(&dc, 8));
    CK(cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
    const size_t smem = 8 * CHB + 8 * 2048 + 128;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> want(128 * N, 0.0f);
    for (int m = 0; m < 128; m++) for (int n = 0; n < N; n++) { float s = 0; for (int k = 0; k < K; k++) s += A[m * K + k] * B[n * K + k]; want[m * N + n] = s; }
    for (int mode = 0; mode < 2; mode++) {
        CK(cudaMemset(dout, 0, 128 * N * 4));
        k_probe<<<1, 160, smem>>>(dA, dB, dout, mode, dc, 1);
        CK(cudaDeviceSynchronize());
        std::vector<float> out(128 * N);
        CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; int bad = 0;
        for (size_t i = 0; i < out.size(); i++) { double e = fabs(out[i] - want[i]); if (e > maxerr) maxerr = e; if (e > 2e-3) bad++; }
        printf("A from %-6s: max|err| = %.3g  bad = %d -> %s\n", mode == 0 ? "TMEM" : "smem", maxerr, bad, bad ? "FAIL" : "ok");
        if (bad) { printf("  got :"); for (int c = 0; c < 8; c++) printf(" %8.4f", out[c]); printf("\n  want:"); for (int c = 0; c < 8; c++) printf(" %8.4f", want[c]); printf("\n"); }
    }
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) { k_probe<<<1, 160, smem>>>(dA, dB, dout, mode, dc, 64); CK(cudaDeviceSynchronize()); }
        long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
        printf("A from %-6s: %lld cycles per hop (operand write + 4 MMAs K=64 N=64 + commit + wait + 64-column tcgen05.ld + cvt), 1 CTA\n", mode == 0 ? "TMEM" : "smem", c / 64);
    }
    return 0;
}
