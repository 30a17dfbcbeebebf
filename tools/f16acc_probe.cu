// Probe (run on a B200): tcgen05.mma kind::f16 with an F16 accumulator (instruction-descriptor c_format = 0).
// Where do the 16-bit results land in tensor memory, and how far are they from the f32-accumulated product?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/f16acc_probe tools/f16acc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../nerfstyle_b200/csrc/tc05.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int K = 32, N = 64;
constexpr uint32_t CHA = 2048, CHB = 1024;

__global__ void __launch_bounds__(160) k_probe(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, uint32_t* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    uint8_t* sb = smem;                         // B: K/8 chunks x 1024 (64 rows)
    uint8_t* sa = smem + (K / 8) * CHB;         // A: K/8 chunks x 2048 (128 rows)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (K / 8) * (int)CHB / 16; i += 160) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    for (int i = tid; i < (K / 8) * (int)CHA / 16; i += 160) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    if (warp == 0) { tc05::tmem_alloc(&tmem_slot, 128); tc05::tmem_relinquish(); }
    if (tid == 0) { tc05::mbar_init(&bar_done, 1); tc05::fence_mbar_init(); }
    tc05::fence_async_smem(); tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tacc = tmem_slot;
    if (warp < 4) {      // poison the accumulator columns so untouched halves are visible
        uint32_t z[32];
        for (int i = 0; i < 32; i++) z[i] = 0x7e007e00u;    // f16 NaN pairs
        const uint32_t tl = tacc + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                     "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                     ::"r"(tl), "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]), "r"(z[8]), "r"(z[9]),
                       "r"(z[10]), "r"(z[11]), "r"(z[12]), "r"(z[13]), "r"(z[14]), "r"(z[15]), "r"(z[16]), "r"(z[17]), "r"(z[18]), "r"(z[19]),
                       "r"(z[20]), "r"(z[21]), "r"(z[22]), "r"(z[23]), "r"(z[24]), "r"(z[25]), "r"(z[26]), "r"(z[27]), "r"(z[28]), "r"(z[29]),
                       "r"(z[30]), "r"(z[31]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                     "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                     ::"r"(tl + 32), "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]), "r"(z[8]), "r"(z[9]),
                       "r"(z[10]), "r"(z[11]), "r"(z[12]), "r"(z[13]), "r"(z[14]), "r"(z[15]), "r"(z[16]), "r"(z[17]), "r"(z[18]), "r"(z[19]),
                       "r"(z[20]), "r"(z[21]), "r"(z[22]), "r"(z[23]), "r"(z[24]), "r"(z[25]), "r"(z[26]), "r"(z[27]), "r"(z[28]), "r"(z[29]),
                       "r"(z[30]), "r"(z[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    // c_format (bits 4-5) = 0: F16 accumulator
    constexpr uint32_t ID = tc05::idesc_f16(128, N, false, false) & ~(3u << 4);
    if (warp == 4 && lane == 0) {
        const uint64_t kB = tc05::desc_kmajor(tc05::smem_u32(sb), CHB), kA = tc05::desc_kmajor(tc05::smem_u32(sa), CHA);
        for (int k = 0; k < K / 16; k++)
            tc05::mma_f16(tacc, kA + (uint64_t)((k * 2 * CHA) >> 4), kB + (uint64_t)((k * 2 * CHB) >> 4), ID, k > 0);
        tc05::mma_commit(&bar_done);
    }
    if (warp < 4) {
        tc05::mbar_wait(&bar_done, 0); tc05::fence_after_sync();
        const uint32_t tl = tacc + ((uint32_t)(warp * 32) << 16);
        for (int half = 0; half < 2; half++) {
            uint32_t d[32];
            tc05::tmem_ld32(tl + 32 * half, d);
            tc05::tmem_ld_wait();
            for (int j = 0; j < 32; j++) out[(size_t)tid * 64 + 32 * half + j] = d[j];
        }
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tacc, 128);
}

static float h2f(uint16_t h) { __half x; memcpy(&x, &h, 2); return __half2float(x); }

int main() {
    std::vector<float> A(128 * K), B(N * K);
    srand(1);
    for (auto& x : A) x = __half2float(__float2half(((rand() % 2001) - 1000) / 1000.0f));
    for (auto& x : B) x = __half2float(__float2half(((rand() % 2001) - 1000) / 1000.0f));
    std::vector<uint8_t> aimg((K / 8) * CHA, 0), bimg((K / 8) * CHB, 0);
    for (int r = 0; r < 128; r++) for (int c = 0; c < K; c++) { __half h = __float2half(A[r * K + c]); memcpy(&aimg[(c / 8) * CHA + r * 16 + (c % 8) * 2], &h, 2); }
    for (int r = 0; r < N; r++) for (int c = 0; c < K; c++) { __half h = __float2half(B[r * K + c]); memcpy(&bimg[(c / 8) * CHB + r * 16 + (c % 8) * 2], &h, 2); }
    uint8_t *dA, *dB; uint32_t* dout;
    CK(cudaMalloc(&dA, aimg.size())); CK(cudaMalloc(&dB, bimg.size())); CK(cudaMalloc(&dout, 128 * 64 * 4));
    CK(cudaMemcpy(dA, aimg.data(), aimg.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
    const size_t smem = (K / 8) * (CHA + CHB) + 128;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_probe<<<1, 160, smem>>>(dA, dB, dout);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> out(128 * 64);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<float> want(128 * N);
    for (int m = 0; m < 128; m++) for (int n = 0; n < N; n++) { float s = 0; for (int k = 0; k < K; k++) s += A[m * K + k] * B[n * K + k]; want[m * N + n] = s; }
    printf("row 0, raw columns 0..7:"); for (int c = 0; c < 8; c++) printf(" %08x", out[c]); printf("\n");
    printf("row 0, raw columns 32..35:"); for (int c = 32; c < 36; c++) printf(" %08x", out[c]); printf("\n");
    printf("want n = 0..7:"); for (int n = 0; n < 8; n++) printf(" %.4f", want[n]); printf("\n");
    // hypothesis P: column c = (n = 2c low half, n = 2c + 1 high half);  hypothesis U: column c low half = n = c
    double errP = 0, errU = 0, errS = 0; int nanP = 0;
    for (int m = 0; m < 128; m++) for (int n = 0; n < N; n++) {
        const uint32_t wp = out[m * 64 + n / 2];
        const float vp = h2f((uint16_t)((n & 1) ? (wp >> 16) : (wp & 0xffff)));
        const float vu = h2f((uint16_t)(out[m * 64 + n] & 0xffff));
        const float ws = out[m * 64 + (n % 32)];      // hypothesis S: column c = (n = c low, n = c + 32 high)
        (void)ws;
        const uint32_t w2 = out[m * 64 + (n % 32)];
        const float vs = h2f((uint16_t)((n >= 32) ? (w2 >> 16) : (w2 & 0xffff)));
        if (vp != vp) nanP++;
        errP = fmax(errP, fabs((vp == vp ? vp : 1e9) - want[m * N + n]));
        errU = fmax(errU, fabs((vu == vu ? vu : 1e9) - want[m * N + n]));
        errS = fmax(errS, fabs((vs == vs ? vs : 1e9) - want[m * N + n]));
    }
    printf("max|err| vs f32 product:  packed (2c, 2c+1): %.4g   unpacked (low half of column n): %.4g   split (c, c+32): %.4g\n", errP, errU, errS);
    // error of the f16 accumulation against the f32 product rounded once to f16 (packed hypothesis)
    double e1 = 0; int differ = 0;
    for (int m = 0; m < 128; m++) for (int n = 0; n < N; n++) {
        const uint32_t wp = out[m * 64 + n / 2];
        const float vp = h2f((uint16_t)((n & 1) ? (wp >> 16) : (wp & 0xffff)));
        const float r = __half2float(__float2half(want[m * N + n]));
        if (vp != r) differ++;
        e1 = fmax(e1, fabs(vp - r));
    }
    printf("packed hypothesis vs round_f16(f32 product): %d of %d elements differ, max |diff| %.4g\n", differ, 128 * N, e1);
    return 0;
}
