// Occupancy-grid update as device-side passes (SURVEY.md 8f NEXT-2, second half): the work of Renderer.update_state
// (/root/reference/renderer.py:120-194) around the density query, without the meshgrid / cat / index_put / boolean-mask
// ops and without any host read-back (the reference reads the mean density and the occupied-cell count on the host).
//   nrf_occ_points_full    cell centres of every cascade in MORTON order, scaled to the cascade and jittered (:120-134,143-155)
//   nrf_occ_points_sparse  the later-phase sampling: random cells + random OCCUPIED cells (:157-181)
//   nrf_occ_flags          occupied-cell flags for nrf_compact_alive (the device-side torch.nonzero of :163)
//   nrf_occ_scatter_max    tmp_grid[indices] = sigmas for the sparse phase (duplicates: the largest wins)
//   nrf_occ_update         grid = max(grid * decay, tmp) where both are valid (:183-186) + partial sums of clamp(grid, 0)
//   nrf_occ_finish         mean density and the packbits threshold min(mean, density_thresh) on the device (:187-188)
//   nrf_packbits_dev       packbits with the threshold read from device memory (:189)
#include "common.cuh"

// xyzs = 2 * coords.float() / (H - 1) - 1   -- torch divides by a host scalar as a multiplication by its f32 reciprocal
__device__ __forceinline__ float cell_centre(uint32_t c, float inv_hm1) {
    return __fsub_rn(__fmul_rn(__fmul_rn(2.0f, (float)c), inv_hm1), 1.0f);
}
// cas_xyzs = xyzs * (bound - half_grid_size);  cas_xyzs += (rand * 2 - 1) * half_grid_size      (renderer.py:127-131)
__device__ __forceinline__ float jittered(float x, float scale, float u, float hgs) {
    return __fadd_rn(__fmul_rn(x, scale), __fmul_rn(__fsub_rn(__fmul_rn(u, 2.0f), 1.0f), hgs));
}

// pts [C, H^3, 3] in Morton order; noise [C, H^3, 3] uniform in [0,1) (or NULL: no jitter, u = 0.5).
// scale_c = bound_c - hgs_c, hgs_c = bound_c / H with bound_c = min(2^c, bound): passed per cascade (host doubles -> f32).
__global__ void __launch_bounds__(256)
k_occ_points_full(float* __restrict__ pts, const float* __restrict__ noise, uint32_t H3, uint32_t C, float inv_hm1,
                  const float* __restrict__ scale_hgs /* [C][2] */) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cas = blockIdx.y;
    if (i >= H3 || cas >= C) return;
    const float scale = __ldg(scale_hgs + 2 * cas), hgs = __ldg(scale_hgs + 2 * cas + 1);
    const size_t o = ((size_t)cas * H3 + i) * 3;
    const uint32_t cx = morton3D_invert_dev(i), cy = morton3D_invert_dev(i >> 1), cz = morton3D_invert_dev(i >> 2);
    const float u0 = noise ? __ldg(noise + o) : 0.5f, u1 = noise ? __ldg(noise + o + 1) : 0.5f, u2 = noise ? __ldg(noise + o + 2) : 0.5f;
    pts[o] = jittered(cell_centre(cx, inv_hm1), scale, u0, hgs);
    pts[o + 1] = jittered(cell_centre(cy, inv_hm1), scale, u1, hgs);
    pts[o + 2] = jittered(cell_centre(cz, inv_hm1), scale, u2, hgs);
}

NRF_EXPORT int nrf_occ_points_full(float* pts, const float* noise, uint32_t H, uint32_t C, const float* scale_hgs, void* stream) {
    if (!pts || !scale_hgs || H < 2 || H > 1024 || C == 0) return NRF_E_INVALID;
    const uint32_t H3 = H * H * H;
    const float inv_hm1 = 1.0f / (float)(H - 1);
    k_occ_points_full<<<dim3(ceil_div_u32(H3, 256), C), 256, 0, (cudaStream_t)stream>>>(pts, noise, H3, C, inv_hm1, scale_hgs);
    return nrf_check_launch();
}

// Sparse phase, per cascade: slots [0, N) take uniformly random cells (rnd_cells [C,N,3] i32 in [0,H)), slots [N, 2N) take
// random OCCUPIED cells: occ_list [C, H^3] i32 holds the Morton indices of the cells with grid > 0 compacted to the front,
// occ_count [C] i32 their number; pick [C,N] uniform in [0,1) selects occ_list[min(floor(pick * count), count - 1)].
// A cascade with no occupied cell fills its second half with index -1 (the scatter skips it; the reference's randint(0, 0)
// raises there).  Outputs indices [C, 2N] i32 and pts [C, 2N, 3].
__global__ void __launch_bounds__(256)
k_occ_points_sparse(float* __restrict__ pts, int32_t* __restrict__ indices, const float* __restrict__ noise,
                    const int32_t* __restrict__ rnd_cells, const float* __restrict__ pick, const int32_t* __restrict__ occ_list,
                    const int32_t* __restrict__ occ_count, uint32_t N, uint32_t H3, uint32_t C, float inv_hm1,
                    const float* __restrict__ scale_hgs) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cas = blockIdx.y;
    if (j >= 2 * N || cas >= C) return;
    const float scale = __ldg(scale_hgs + 2 * cas), hgs = __ldg(scale_hgs + 2 * cas + 1);
    uint32_t cx, cy, cz;
    int32_t idx;
    if (j < N) {
        const int32_t* rc = rnd_cells + ((size_t)cas * N + j) * 3;
        cx = (uint32_t)__ldg(rc); cy = (uint32_t)__ldg(rc + 1); cz = (uint32_t)__ldg(rc + 2);
        idx = (int32_t)morton3D_dev(cx, cy, cz);
    } else {
        const int32_t cnt = __ldg(occ_count + cas);
        if (cnt <= 0) {
            idx = -1; cx = cy = cz = 0;
        } else {
            const float u = __ldg(pick + (size_t)cas * N + (j - N));
            const int32_t k = min((int32_t)(u * (float)cnt), cnt - 1);
            idx = __ldg(occ_list + (size_t)cas * H3 + k);
            cx = morton3D_invert_dev((uint32_t)idx); cy = morton3D_invert_dev((uint32_t)idx >> 1); cz = morton3D_invert_dev((uint32_t)idx >> 2);
        }
    }
    const size_t o = ((size_t)cas * 2 * N + j) * 3;
    const float u0 = noise ? __ldg(noise + o) : 0.5f, u1 = noise ? __ldg(noise + o + 1) : 0.5f, u2 = noise ? __ldg(noise + o + 2) : 0.5f;
    indices[(size_t)cas * 2 * N + j] = idx;
    pts[o] = jittered(cell_centre(cx, inv_hm1), scale, u0, hgs);
    pts[o + 1] = jittered(cell_centre(cy, inv_hm1), scale, u1, hgs);
    pts[o + 2] = jittered(cell_centre(cz, inv_hm1), scale, u2, hgs);
}

NRF_EXPORT int nrf_occ_points_sparse(float* pts, int32_t* indices, const float* noise, const int32_t* rnd_cells, const float* pick,
                                     const int32_t* occ_list, const int32_t* occ_count, uint32_t N, uint32_t H, uint32_t C,
                                     const float* scale_hgs, void* stream) {
    if (N == 0) return NRF_OK;
    if (!pts || !indices || !rnd_cells || !pick || !occ_list || !occ_count || !scale_hgs || H < 2 || H > 1024 || C == 0) return NRF_E_INVALID;
    const float inv_hm1 = 1.0f / (float)(H - 1);
    k_occ_points_sparse<<<dim3(ceil_div_u32(2 * (uint64_t)N, 256), C), 256, 0, (cudaStream_t)stream>>>(
        pts, indices, noise, rnd_cells, pick, occ_list, occ_count, N, H * H * H, C, inv_hm1, scale_hgs);
    return nrf_check_launch();
}

// occ_list / occ_count of the sparse phase come from nrf_compact_alive (raymarching.cu) applied per cascade to these flags:
// flags[c][i] = i when grid[c][i] > 0, else -1  (torch.nonzero(self.density_grid[cas] > 0), renderer.py:163, without the sync)
__global__ void __launch_bounds__(256)
k_occ_flags(const float* __restrict__ grid, uint64_t n, uint32_t H3, int32_t* __restrict__ flags) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (__ldg(grid + i) > 0.0f) ? (int32_t)(i % H3) : -1;
}

NRF_EXPORT int nrf_occ_flags(const float* grid, uint32_t H, uint32_t C, int32_t* flags, void* stream) {
    if (!grid || !flags || H < 2 || H > 1024 || C == 0) return NRF_E_INVALID;
    const uint32_t H3 = H * H * H;
    const uint64_t n = (uint64_t)H3 * C;
    k_occ_flags<<<ceil_div_u32(n, 256), 256, 0, (cudaStream_t)stream>>>(grid, n, H3, flags);
    return nrf_check_launch();
}

// tmp[cas][indices[cas][j]] = max over duplicates of sigma * density_scale  (tmp pre-filled with -1; values are >= 0, so the
// int ordering of the float bits is the float ordering).  The reference's index_put keeps an arbitrary duplicate.
__global__ void __launch_bounds__(256)
k_occ_scatter_max(float* __restrict__ tmp, const int32_t* __restrict__ indices, const float* __restrict__ sigmas, float density_scale,
                  uint32_t n_per_cas, uint32_t H3, uint32_t C) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cas = blockIdx.y;
    if (j >= n_per_cas || cas >= C) return;
    const int32_t idx = __ldg(indices + (size_t)cas * n_per_cas + j);
    if (idx < 0 || (uint32_t)idx >= H3) return;
    const float v = __fmul_rn(__ldg(sigmas + (size_t)cas * n_per_cas + j), density_scale);
    if (!(v >= 0.0f)) return;                       // negative / NaN never becomes valid (tmp stays -1)
    atomicMax(reinterpret_cast<int*>(tmp + (size_t)cas * H3 + idx), __float_as_int(v));
}

NRF_EXPORT int nrf_occ_scatter_max(float* tmp, const int32_t* indices, const float* sigmas, float density_scale, uint32_t n_per_cas,
                                   uint32_t H, uint32_t C, void* stream) {
    if (n_per_cas == 0) return NRF_OK;
    if (!tmp || !indices || !sigmas || C == 0) return NRF_E_INVALID;
    k_occ_scatter_max<<<dim3(ceil_div_u32(n_per_cas, 256), C), 256, 0, (cudaStream_t)stream>>>(tmp, indices, sigmas, density_scale,
                                                                                               n_per_cas, H * H * H, C);
    return nrf_check_launch();
}

// grid[i] = max(grid[i] * decay, tmp[i]) where grid[i] >= 0 and tmp[i] >= 0 (renderer.py:183-186); tmp = values[i] * tmp_scale
// (full phase: values = the sigmas in Morton order, tmp_scale = density_scale; sparse phase: values = the scattered tmp grid,
// tmp_scale = 1).  partial[b] = sum over the block of max(grid, 0) in double (deterministic two-stage mean).
__global__ void __launch_bounds__(256)
k_occ_update(float* __restrict__ grid, const float* __restrict__ values, float tmp_scale, float decay, uint64_t n, double* __restrict__ partial) {
    __shared__ double wsum[8];
    double acc = 0.0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        float g = grid[i];
        const float t = __fmul_rn(__ldg(values + i), tmp_scale);
        if (g >= 0.0f && t >= 0.0f) { g = fmaxf(__fmul_rn(g, decay), t); grid[i] = g; }
        acc += (double)fmaxf(g, 0.0f);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(NRF_FULL_MASK, acc, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += wsum[w];
        partial[blockIdx.x] = s;
    }
}

// state[0] = mean density, state[1] = min(mean, density_thresh)
__global__ void k_occ_finish(const double* __restrict__ partial, uint32_t nblocks, uint64_t n, float density_thresh, float* __restrict__ state) {
    __shared__ double wsum[32];
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < nblocks; i += blockDim.x) acc += partial[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(NRF_FULL_MASK, acc, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (uint32_t w = 0; w < blockDim.x / 32; w++) s += wsum[w];
        const float mean = (float)(s / (double)n);
        state[0] = mean;
        state[1] = fminf(mean, density_thresh);
    }
}

#define OCC_UPDATE_BLOCKS 1184     // 148 SMs x 8 resident blocks

NRF_EXPORT uint64_t nrf_occ_scratch_bytes(void) { return OCC_UPDATE_BLOCKS * sizeof(double); }

NRF_EXPORT int nrf_occ_update(float* grid, const float* values, float tmp_scale, float decay, uint64_t n, float density_thresh,
                              float* state, void* scratch, void* stream) {
    if (n == 0) return NRF_OK;
    if (!grid || !values || !state || !scratch) return NRF_E_INVALID;
    const uint32_t nb = (uint32_t)min((uint64_t)OCC_UPDATE_BLOCKS, (n + 255) / 256);
    k_occ_update<<<nb, 256, 0, (cudaStream_t)stream>>>(grid, values, tmp_scale, decay, n, (double*)scratch);
    k_occ_finish<<<1, 256, 0, (cudaStream_t)stream>>>((const double*)scratch, nb, n, density_thresh, state);
    return nrf_check_launch();
}

// packbits (raymarching.cu:367-388) with the threshold in device memory: bit i of byte n = grid[8n + i] > *thresh
__global__ void __launch_bounds__(256)
k_packbits_dev(const float4* __restrict__ grid4, uint32_t N, const float* __restrict__ thresh, uint8_t* __restrict__ out) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float th = __ldg(thresh);
    const float4 a = __ldg(grid4 + 2 * (size_t)n), b = __ldg(grid4 + 2 * (size_t)n + 1);
    uint32_t bits = 0;
    bits |= (a.x > th) ? 1u : 0u; bits |= (a.y > th) ? 2u : 0u; bits |= (a.z > th) ? 4u : 0u; bits |= (a.w > th) ? 8u : 0u;
    bits |= (b.x > th) ? 16u : 0u; bits |= (b.y > th) ? 32u : 0u; bits |= (b.z > th) ? 64u : 0u; bits |= (b.w > th) ? 128u : 0u;
    out[n] = (uint8_t)bits;
}

NRF_EXPORT int nrf_packbits_dev(const float* grid, uint32_t N, const float* density_thresh_dev, uint8_t* bitfield, void* stream) {
    if (N == 0) return NRF_OK;
    if (!grid || !density_thresh_dev || !bitfield || (((uintptr_t)grid) & 15)) return NRF_E_INVALID;
    k_packbits_dev<<<ceil_div_u32(N, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(grid), N, density_thresh_dev, bitfield);
    return nrf_check_launch();
}
