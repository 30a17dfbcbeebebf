/* A plain-C host of libnerfstyle_b200.so: no Python, no torch -- only the C ABI of include/nerfstyle_b200.h and the CUDA
 * runtime for device memory.  It runs three ops of the ray-marching set and checks them against closed forms computed
 * here (Morton round trip, packbits bit pattern, slab intersection of axis-aligned rays).
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/c_host.c -o examples/c_host \
 *       -L nerfstyle_b200 -lnerfstyle_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../nerfstyle_b200' -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "nerfstyle_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA: %s\n", cudaGetErrorString(e_)); return 2; } } while (0)
#define NRF(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s: %s (cuda %d)\n", #x, nrf_error_string(rc_), nrf_last_cuda_error()); return 3; } } while (0)

int main(void) {
    printf("nerfstyle_b200 ABI version %d\n", nrf_version());
    int sms = 0, maj = 0, min = 0, l2 = 0;
    NRF(nrf_device_info(&sms, &maj, &min, &l2));
    printf("device: sm_%d%d, %d SMs, L2 %d MB\n", maj, min, sms, l2 >> 20);
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    int bad = 0;

    /* 1. morton3D / morton3D_invert round trip (raymarching.cu:313-359) */
    enum { N = 4096 };
    int32_t h_c[N * 3], h_back[N * 3], h_idx[N];
    for (int i = 0; i < N; i++) { h_c[3 * i] = (i * 7) & 127; h_c[3 * i + 1] = (i * 13) & 127; h_c[3 * i + 2] = (i * 29) & 127; }
    int32_t *d_c, *d_idx, *d_back;
    CK(cudaMalloc((void**)&d_c, sizeof h_c)); CK(cudaMalloc((void**)&d_idx, sizeof h_idx)); CK(cudaMalloc((void**)&d_back, sizeof h_back));
    CK(cudaMemcpyAsync(d_c, h_c, sizeof h_c, cudaMemcpyHostToDevice, s));
    NRF(nrf_morton3D(d_c, N, d_idx, s));
    NRF(nrf_morton3D_invert(d_idx, N, d_back, s));
    CK(cudaMemcpyAsync(h_back, d_back, sizeof h_back, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h_idx, d_idx, sizeof h_idx, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    bad += memcmp(h_c, h_back, sizeof h_c) != 0;
    {   /* closed form for one point: interleave bits x -> bit 3k, y -> 3k+1, z -> 3k+2 */
        int x = h_c[3], y = h_c[4], z = h_c[5], m = 0;
        for (int k = 0; k < 10; k++) m |= ((x >> k) & 1) << (3 * k) | ((y >> k) & 1) << (3 * k + 1) | ((z >> k) & 1) << (3 * k + 2);
        bad += h_idx[1] != m;
    }
    printf("morton3D round trip + closed form: %s\n", bad ? "FAIL" : "ok");

    /* 2. packbits (raymarching.cu:367-388): bit i of byte n = grid[8n+i] > thresh */
    enum { NB = 1024 };
    float h_g[NB * 8];
    unsigned char h_bits[NB];
    for (int i = 0; i < NB * 8; i++) h_g[i] = (float)((i * 2654435761u) >> 24) / 255.0f;
    float* d_g; unsigned char* d_bits;
    CK(cudaMalloc((void**)&d_g, sizeof h_g)); CK(cudaMalloc((void**)&d_bits, NB));
    CK(cudaMemcpyAsync(d_g, h_g, sizeof h_g, cudaMemcpyHostToDevice, s));
    NRF(nrf_packbits(d_g, NB, 0.5f, d_bits, s));
    CK(cudaMemcpyAsync(h_bits, d_bits, NB, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    int bad2 = 0;
    for (int n = 0; n < NB; n++) {
        unsigned char e = 0;
        for (int i = 0; i < 8; i++) e |= (unsigned char)((h_g[8 * n + i] > 0.5f) << i);
        bad2 += e != h_bits[n];
    }
    printf("packbits: %s\n", bad2 ? "FAIL" : "ok");

    /* 3. near_far_from_aabb (raymarching.cu:191-244): axis-aligned rays from x = -3 through the box [-2,2]^3 */
    enum { NR = 256 };
    float h_o[NR * 3], h_d[NR * 3], h_near[NR], h_far[NR], h_aabb[6] = {-2, -2, -2, 2, 2, 2};
    for (int i = 0; i < NR; i++) {
        h_o[3 * i] = -3.0f; h_o[3 * i + 1] = -1.9f + 3.8f * i / NR; h_o[3 * i + 2] = 0.25f;
        h_d[3 * i] = 1.0f; h_d[3 * i + 1] = 0.0f; h_d[3 * i + 2] = 0.0f;
    }
    float *d_o, *d_d, *d_aabb, *d_near, *d_far;
    CK(cudaMalloc((void**)&d_o, sizeof h_o)); CK(cudaMalloc((void**)&d_d, sizeof h_d)); CK(cudaMalloc((void**)&d_aabb, sizeof h_aabb));
    CK(cudaMalloc((void**)&d_near, sizeof h_near)); CK(cudaMalloc((void**)&d_far, sizeof h_far));
    CK(cudaMemcpyAsync(d_o, h_o, sizeof h_o, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_d, h_d, sizeof h_d, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_aabb, h_aabb, sizeof h_aabb, cudaMemcpyHostToDevice, s));
    NRF(nrf_near_far_from_aabb(d_o, d_d, d_aabb, NR, 0.2f, d_near, d_far, s));
    CK(cudaMemcpyAsync(h_near, d_near, sizeof h_near, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h_far, d_far, sizeof h_far, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    int bad3 = 0;
    for (int i = 0; i < NR; i++) bad3 += (fabsf(h_near[i] - 1.0f) > 1e-6f) || (fabsf(h_far[i] - 5.0f) > 1e-6f);
    printf("near_far_from_aabb: %s\n", bad3 ? "FAIL" : "ok");

    /* error path: NULL pointers are rejected before any launch */
    int rc = nrf_packbits(NULL, 8, 0.5f, NULL, s);
    printf("argument validation: %s (%s)\n", rc == NRF_E_INVALID ? "ok" : "FAIL", nrf_error_string(rc));
    const int fails = bad + bad2 + bad3 + (rc != NRF_E_INVALID);
    printf(fails ? "C HOST FAILED\n" : "C HOST OK\n");
    return fails ? 1 : 0;
}
