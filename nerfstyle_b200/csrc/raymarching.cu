// Ray-marching operator set for sm_100a (C ABI in include/nerfstyle_b200.h).
//
// Behavioural contract: /root/reference/raymarching/src/raymarching.cu (cited per function).  The code is
// written from scratch: every floating-point operation that decides an integer (cell index, sample
// count) is spelled with explicit round-to-nearest intrinsics (__fmaf_rn/__fmul_rn/__fadd_rn) at exactly the
// contraction points nvcc chose for the reference kernels, so sample counts and
// occupancy indices are bit-exact and independent of compiler flags.
#include "common.cuh"
#include <float.h>

thread_local int g_nrf_last_cuda_error = 0;

NRF_EXPORT const char* nrf_error_string(int code) {
    switch (code) {
        case NRF_OK: return "ok";
        case NRF_E_INVALID: return "invalid argument";
        case NRF_E_UNSUPPORTED: return "unsupported configuration";
        case NRF_E_CUDA: return "CUDA error";
        default: return "unknown error";
    }
}
NRF_EXPORT int nrf_last_cuda_error(void) { return g_nrf_last_cuda_error; }
NRF_EXPORT int nrf_version(void) { return 5; }   // 5: field forward, staged march, peer-memory optimizer, step-major device loop; 4: fp32 parity-mode MLP (nrf_mlp_*_f32); 3: paired tables, ray generation, occupancy update, loss head, _ex forms
NRF_EXPORT int nrf_device_info(int* sm_count, int* cc_major, int* cc_minor, int* l2_bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    cudaDeviceProp p;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) { g_nrf_last_cuda_error = (int)e; return NRF_E_CUDA; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (l2_bytes) *l2_bytes = p.l2CacheSize;
    return NRF_OK;
}

// ------------------------------------------------------------------------------------------------
// small utilities: near/far, spherical coords, Morton codes, bit packing
// ------------------------------------------------------------------------------------------------

// raymarching.cu:191-244
__global__ void k_near_far_from_aabb(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                     const float* __restrict__ aabb, uint32_t N, float min_near,
                                     float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[3 * n], oy = rays_o[3 * n + 1], oz = rays_o[3 * n + 2];
    const float dx = rays_d[3 * n], dy = rays_d[3 * n + 1], dz = rays_d[3 * n + 2];
    const float rdx = __fdiv_rn(1.0f, dx), rdy = __fdiv_rn(1.0f, dy), rdz = __fdiv_rn(1.0f, dz);
    float near = __fmul_rn(__fsub_rn(aabb[0], ox), rdx), far = __fmul_rn(__fsub_rn(aabb[3], ox), rdx);
    if (near > far) { float t = near; near = far; far = t; }
    float near_y = __fmul_rn(__fsub_rn(aabb[1], oy), rdy), far_y = __fmul_rn(__fsub_rn(aabb[4], oy), rdy);
    if (near_y > far_y) { float t = near_y; near_y = far_y; far_y = t; }
    if (near > far_y || near_y > far) { nears[n] = FLT_MAX; fars[n] = FLT_MAX; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = __fmul_rn(__fsub_rn(aabb[2], oz), rdz), far_z = __fmul_rn(__fsub_rn(aabb[5], oz), rdz);
    if (near_z > far_z) { float t = near_z; near_z = far_z; far_z = t; }
    if (near > far_z || near_z > far) { nears[n] = FLT_MAX; fars[n] = FLT_MAX; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    nears[n] = near;
    fars[n] = far;
}

NRF_EXPORT int nrf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                                      float min_near, float* nears, float* fars, void* stream) {
    if (N == 0) return NRF_OK;
    if (!rays_o || !rays_d || !aabb || !nears || !fars) return NRF_E_INVALID;
    k_near_far_from_aabb<<<ceil_div_u32(N, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    return nrf_check_launch();
}

// raymarching.cu:262-297
__global__ void k_sph_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float radius,
                               uint32_t N, float* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float RPI = 0.3183098861837907f;
    const float ox = rays_o[3 * n], oy = rays_o[3 * n + 1], oz = rays_o[3 * n + 2];
    const float dx = rays_d[3 * n], dy = rays_d[3 * n + 1], dz = rays_d[3 * n + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float B = ox * dx + oy * dy + oz * dz;
    const float C = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-B + sqrtf(B * B - A * C)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    coords[2 * n] = 2 * theta * RPI - 1;
    coords[2 * n + 1] = phi * RPI;
}

NRF_EXPORT int nrf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                                void* stream) {
    if (N == 0) return NRF_OK;
    if (!rays_o || !rays_d || !coords) return NRF_E_INVALID;
    k_sph_from_ray<<<ceil_div_u32(N, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, radius, N, coords);
    return nrf_check_launch();
}

// expand_bits / morton3D_dev / morton3D_invert_dev (raymarching.cu:56-81) live in common.cuh

// raymarching.cu:313-325
__global__ void k_morton3D(const int32_t* __restrict__ coords, uint32_t N, int32_t* __restrict__ indices) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    indices[n] = (int32_t)morton3D_dev((uint32_t)coords[3 * n], (uint32_t)coords[3 * n + 1], (uint32_t)coords[3 * n + 2]);
}
// raymarching.cu:336-353
__global__ void k_morton3D_invert(const int32_t* __restrict__ indices, uint32_t N, int32_t* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int32_t ind = indices[n];
    coords[3 * n + 0] = (int32_t)morton3D_invert_dev((uint32_t)(ind >> 0));
    coords[3 * n + 1] = (int32_t)morton3D_invert_dev((uint32_t)(ind >> 1));
    coords[3 * n + 2] = (int32_t)morton3D_invert_dev((uint32_t)(ind >> 2));
}
NRF_EXPORT int nrf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream) {
    if (N == 0) return NRF_OK;
    if (!coords || !indices) return NRF_E_INVALID;
    k_morton3D<<<ceil_div_u32(N, 256), 256, 0, (cudaStream_t)stream>>>(coords, N, indices);
    return nrf_check_launch();
}
NRF_EXPORT int nrf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream) {
    if (N == 0) return NRF_OK;
    if (!coords || !indices) return NRF_E_INVALID;
    k_morton3D_invert<<<ceil_div_u32(N, 256), 256, 0, (cudaStream_t)stream>>>(indices, N, coords);
    return nrf_check_launch();
}

// raymarching.cu:367-388.  One thread packs 4 bytes (32 cells = 8 x 16-byte loads) so that global
// stores are 4 bytes wide; the tail (N % 4 bytes) falls back to one byte per thread.
__global__ void k_packbits4(const float4* __restrict__ grid4, uint32_t N4, float thresh, uint32_t* __restrict__ out) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N4) return;
    uint32_t word = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float4 v = __ldg(grid4 + (size_t)n * 8 + j);
        uint32_t nib = (v.x > thresh ? 1u : 0u) | (v.y > thresh ? 2u : 0u) | (v.z > thresh ? 4u : 0u) | (v.w > thresh ? 8u : 0u);
        word |= nib << (4 * j);
    }
    out[n] = word;
}
__global__ void k_packbits1(const float* __restrict__ grid, uint32_t n0, uint32_t N, float thresh, uint8_t* __restrict__ out) {
    const uint32_t n = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    uint8_t bits = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) bits |= (grid[(size_t)n * 8 + i] > thresh) ? (uint8_t)(1u << i) : 0;
    out[n] = bits;
}
NRF_EXPORT int nrf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, void* stream) {
    if (N == 0) return NRF_OK;
    if (!grid || !bitfield) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t N4 = 0;
    if ((((uintptr_t)grid) & 15) == 0 && (((uintptr_t)bitfield) & 3) == 0) N4 = N / 4;
    if (N4) k_packbits4<<<ceil_div_u32(N4, 256), 256, 0, s>>>((const float4*)grid, N4, density_thresh, (uint32_t*)bitfield);
    if (N4 * 4 < N) k_packbits1<<<ceil_div_u32(N - N4 * 4, 256), 256, 0, s>>>(grid, N4 * 4, N, density_thresh, bitfield);
    return nrf_check_launch();
}

// ------------------------------------------------------------------------------------------------
// The marching state machine shared by train and inference marching
// (raymarching.cu:436-500 / :1037-1119; geometry normative in SURVEY.md 8a.2, contraction 8a.3)
// ------------------------------------------------------------------------------------------------
struct MarchCtx {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, sx, sy, sz;
    float rH, H3, Hf, Hm1, bound, nbound, dt_gamma, dt_min, dt_max;
    int Cm1;
};

__device__ __forceinline__ void march_init(MarchCtx& c, const float* __restrict__ o, const float* __restrict__ d,
                                           float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    c.ox = o[0]; c.oy = o[1]; c.oz = o[2];
    c.dx = d[0]; c.dy = d[1]; c.dz = d[2];
    c.rdx = __fdiv_rn(1.0f, c.dx); c.rdy = __fdiv_rn(1.0f, c.dy); c.rdz = __fdiv_rn(1.0f, c.dz);
    c.sx = 0.5f * copysignf(1.0f, c.dx); c.sy = 0.5f * copysignf(1.0f, c.dy); c.sz = 0.5f * copysignf(1.0f, c.dz);
    c.Hf = (float)H;
    c.Hm1 = (float)(H - 1);
    c.rH = __fdiv_rn(1.0f, c.Hf);
    c.H3 = (float)(H * H * H);
    c.bound = bound; c.nbound = -bound;
    c.dt_gamma = dt_gamma;
    const float two_sqrt3 = 2.0f * 1.7320508075688772f;
    c.dt_min = __fdiv_rn(two_sqrt3, (float)max_steps);
    c.dt_max = __fdiv_rn(__fmul_rn(two_sqrt3, (float)(1 << (C - 1))), c.Hf);
    c.Cm1 = (int)C - 1;
}

// frexpf exponent clamped to [0, Cm1] (raymarching.cu:42-54): for finite v >= 0 the frexp exponent is
// (biased exponent - 126); zero / denormals clamp to 0 either way.
__device__ __forceinline__ int mip_exponent(float v, int Cm1) {
    const int e = (int)((__float_as_uint(v) >> 23) & 0xffu) - 126;
    return min(Cm1, max(0, e));
}

__device__ __forceinline__ float march_t0(const MarchCtx& c, float t, float noise) {
    return __fmaf_rn(noise, nrf_clamp(__fmul_rn(t, c.dt_gamma), c.dt_min, c.dt_max), t);
}

// One visit of the loop body at t.  Returns true when the cell is occupied (x,y,z,dt = the sample); else
// advances t past the empty cell exactly like the reference's do/while and returns false.
__device__ __forceinline__ bool march_visit(const MarchCtx& c, const uint8_t* __restrict__ grid, float& t,
                                            float& x, float& y, float& z, float& dt) {
    x = nrf_clamp(__fmaf_rn(c.dx, t, c.ox), c.nbound, c.bound);
    y = nrf_clamp(__fmaf_rn(c.dy, t, c.oy), c.nbound, c.bound);
    z = nrf_clamp(__fmaf_rn(c.dz, t, c.oz), c.nbound, c.bound);
    dt = nrf_clamp(__fmul_rn(t, c.dt_gamma), c.dt_min, c.dt_max);
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    const int level = max(mip_exponent(mx, c.Cm1), mip_exponent(__fmul_rn(dt, c.Hf) * 0.5f, c.Cm1));
    const float mip_bound = fminf(__uint_as_float((uint32_t)(127 + level) << 23), c.bound);
    const float mip_rbound = __fdiv_rn(1.0f, mip_bound);
    const int nx = (int)nrf_clamp(__fmul_rn(0.5f * __fmaf_rn(x, mip_rbound, 1.0f), c.Hf), 0.0f, c.Hm1);
    const int ny = (int)nrf_clamp(__fmul_rn(0.5f * __fmaf_rn(y, mip_rbound, 1.0f), c.Hf), 0.0f, c.Hm1);
    const int nz = (int)nrf_clamp(__fmul_rn(0.5f * __fmaf_rn(z, mip_rbound, 1.0f), c.Hf), 0.0f, c.Hm1);
    const uint32_t index = (uint32_t)__fmaf_rn(c.H3, (float)level, (float)morton3D_dev((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    const bool occ = (__ldg(grid + (index >> 3)) >> (index & 7u)) & 1u;
    if (occ) return true;
    const float tx = __fmul_rn(c.rdx, __fsub_rn(__fmul_rn(mip_bound, __fmaf_rn(__fmul_rn(c.rH, __fadd_rn(c.sx, __fadd_rn((float)nx, 0.5f))), 2.0f, -1.0f)), x));
    const float ty = __fmul_rn(c.rdy, __fsub_rn(__fmul_rn(mip_bound, __fmaf_rn(__fmul_rn(c.rH, __fadd_rn(c.sy, __fadd_rn((float)ny, 0.5f))), 2.0f, -1.0f)), y));
    const float tz = __fmul_rn(c.rdz, __fsub_rn(__fmul_rn(mip_bound, __fmaf_rn(__fmul_rn(c.rH, __fadd_rn(c.sz, __fadd_rn((float)nz, 0.5f))), 2.0f, -1.0f)), z));
    const float tt = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
    do {
        t = __fadd_rn(t, nrf_clamp(__fmul_rn(t, c.dt_gamma), c.dt_min, c.dt_max));
    } while (t < tt);
    return false;
}

// Occupancy test at t without advancing: x,y,z,dt = the sample; tt = exit time of the cell when it is empty.
__device__ __forceinline__ bool march_eval(const MarchCtx& c, const uint8_t* __restrict__ grid, float t,
                                           float& x, float& y, float& z, float& dt, float& tt) {
    x = nrf_clamp(__fmaf_rn(c.dx, t, c.ox), c.nbound, c.bound);
    y = nrf_clamp(__fmaf_rn(c.dy, t, c.oy), c.nbound, c.bound);
    z = nrf_clamp(__fmaf_rn(c.dz, t, c.oz), c.nbound, c.bound);
    dt = nrf_clamp(__fmul_rn(t, c.dt_gamma), c.dt_min, c.dt_max);
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    const int level = max(mip_exponent(mx, c.Cm1), mip_exponent(__fmul_rn(dt, c.Hf) * 0.5f, c.Cm1));
    const float mip_bound = fminf(__uint_as_float((uint32_t)(127 + level) << 23), c.bound);
    const float mip_rbound = __fdiv_rn(1.0f, mip_bound);
    const int nx = (int)nrf_clamp(__fmul_rn(0.5f * __fmaf_rn(x, mip_rbound, 1.0f), c.Hf), 0.0f, c.Hm1);
    const int ny = (int)nrf_clamp(__fmul_rn(0.5f * __fmaf_rn(y, mip_rbound, 1.0f), c.Hf), 0.0f, c.Hm1);
    const int nz = (int)nrf_clamp(__fmul_rn(0.5f * __fmaf_rn(z, mip_rbound, 1.0f), c.Hf), 0.0f, c.Hm1);
    const uint32_t index = (uint32_t)__fmaf_rn(c.H3, (float)level, (float)morton3D_dev((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    const bool occ = (__ldg(grid + (index >> 3)) >> (index & 7u)) & 1u;
    const float tx = __fmul_rn(c.rdx, __fsub_rn(__fmul_rn(mip_bound, __fmaf_rn(__fmul_rn(c.rH, __fadd_rn(c.sx, __fadd_rn((float)nx, 0.5f))), 2.0f, -1.0f)), x));
    const float ty = __fmul_rn(c.rdy, __fsub_rn(__fmul_rn(mip_bound, __fmaf_rn(__fmul_rn(c.rH, __fadd_rn(c.sy, __fadd_rn((float)ny, 0.5f))), 2.0f, -1.0f)), y));
    const float tz = __fmul_rn(c.rdz, __fsub_rn(__fmul_rn(mip_bound, __fmaf_rn(__fmul_rn(c.rH, __fadd_rn(c.sz, __fadd_rn((float)nz, 0.5f))), 2.0f, -1.0f)), z));
    tt = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
    return occ;
}

__device__ __forceinline__ float march_next_t(const MarchCtx& c, float t) {
    return __fadd_rn(t, nrf_clamp(__fmul_rn(t, c.dt_gamma), c.dt_min, c.dt_max));
}

// ------------------------------------------------------------------------------------------------
// Warp-per-ray walker.  The candidate times t_0, t_1 = t_0 + dt(t_0), ... do not depend on the occupancy
// (occupied: t += dt; empty: t += dt until t >= exit time), so a warp tests 32 consecutive candidates at
// once: every lane runs the (cheap, strictly sequential) float recurrence and keeps its own t_i, tests its
// cell with one byte load, and the data-dependent skipping ("after an empty cell jump to the first
// candidate past its exit time") is resolved with warp ballots / or-reductions by pointer doubling over the
// per-lane successor indices.  Visited-and-occupied lanes are compacted with a prefix popcount.  Bit-exact
// with the sequential walk (same float ops in the same order for every value that is kept).
// ------------------------------------------------------------------------------------------------
#define MW_WARPS 4

template <bool WRITE>
__global__ void __launch_bounds__(MW_WARPS * 32)
k_march_warp(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid,
             float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, uint32_t rows,
             const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ noises,
             int32_t* __restrict__ rays, float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
             float* __restrict__ t_stage) {
    // t_stage (count pass only, or NULL): [N, max_steps] -- the marching time of every sample the ray emits, in order.  A
    // sample is a pure function of its time (x = clamp(fma(d, t, o)), dt = clamp(t dt_gamma), delta = t_after - previous
    // t_after), so the second pass need not walk the occupancy grid again: k_march_emit streams these times back.
    __shared__ float s_t[MW_WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t n = blockIdx.x * MW_WARPS + wib;
    if (n >= N) return;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t limit = max_steps, point_index = 0;
    if (WRITE) {
        point_index = (uint32_t)rays[3 * (size_t)n + 1];
        limit = (uint32_t)rays[3 * (size_t)n + 2];
        if (limit == 0) return;
        if (point_index + limit >= M) {      // dropped ray (raymarching.cu:517): its slots stay zero
            for (uint32_t i = point_index + lane; i < min(point_index + limit, rows); i += 32) {
                xyzs[3 * (size_t)i] = 0; xyzs[3 * (size_t)i + 1] = 0; xyzs[3 * (size_t)i + 2] = 0;
                dirs[3 * (size_t)i] = 0; dirs[3 * (size_t)i + 1] = 0; dirs[3 * (size_t)i + 2] = 0;
                reinterpret_cast<float4*>(deltas)[i] = make_float4(0, 0, 0, 0);
            }
            return;
        }
    }
    MarchCtx c;
    march_init(c, rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H);
    const float far = fars[n];
    float cur_t = march_t0(c, nears[n], noises ? noises[n] : 0.0f);
    float last_t = cur_t;
    uint32_t count = 0;
    while (cur_t < far && count < limit) {
        // 32 consecutive candidates; lane i keeps t_i, tc ends as t_32
        float tc = cur_t, my_t = cur_t;
#pragma unroll
        for (int i = 0; i < 32; i++) { if (lane == i) my_t = tc; tc = march_next_t(c, tc); }
        const bool valid = my_t < far;
        float x, y, z, dt, tt;
        const bool occ = march_eval(c, grid, my_t, x, y, z, dt, tt);
        s_t[wib][lane] = my_t;
        __syncwarp();
        int nxt = lane + 1;
        if (!occ) {      // first j >= lane+1 whose t_j is not < tt (the do/while of raymarching.cu:497-499)
            int lo_ = lane + 1, hi_ = 32;
            while (lo_ < hi_) { const int mid = (lo_ + hi_) >> 1; if (!(s_t[wib][mid] < tt)) hi_ = mid; else lo_ = mid + 1; }
            nxt = lo_;
        }
        __syncwarp();
        // which lanes does the sequential walk visit?  pointer doubling from lane 0
        bool marked = (lane == 0) && valid;
        int jump = nxt;
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const uint32_t m = __reduce_or_sync(NRF_FULL_MASK, (marked && jump < 32) ? (1u << jump) : 0u);
            if (((m >> lane) & 1u) && valid) marked = true;
            const int j2 = __shfl_sync(NRF_FULL_MASK, jump, min(jump, 31));
            jump = (jump < 32) ? j2 : 32;
        }
        const uint32_t V = __ballot_sync(NRF_FULL_MASK, marked);
        uint32_t E = __ballot_sync(NRF_FULL_MASK, marked && occ);
        const uint32_t remaining = limit - count;
        bool finished = false;
        if ((uint32_t)__popc(E) >= remaining) {          // the step cap falls inside this chunk: keep the first `remaining`
            const uint32_t pos = __fns(E, 0, (int)remaining + 1);
            if (pos != 0xffffffffu) E &= (1u << pos) - 1u;
            finished = true;
        }
        if (WRITE) {
            const float t_after = __fadd_rn(my_t, dt);
            const uint32_t before = E & lt_mask;
            const int src = before ? (31 - __clz(before)) : lane;
            const float prev_after = __shfl_sync(NRF_FULL_MASK, t_after, src);
            if ((E >> lane) & 1u) {
                const size_t o = (size_t)point_index + count + __popc(before);
                xyzs[3 * o] = x; xyzs[3 * o + 1] = y; xyzs[3 * o + 2] = z;
                dirs[3 * o] = c.dx; dirs[3 * o + 1] = c.dy; dirs[3 * o + 2] = c.dz;
                reinterpret_cast<float4*>(deltas)[o] = make_float4(dt, __fsub_rn(t_after, before ? prev_after : last_t), 0.0f, 0.0f);
            }
            if (E) last_t = __shfl_sync(NRF_FULL_MASK, t_after, 31 - __clz(E));
        }
        if (!WRITE && t_stage && ((E >> lane) & 1u)) t_stage[(size_t)n * max_steps + count + __popc(E & lt_mask)] = my_t;
        count += __popc(E);
        if (finished) break;
        const int Lv = 31 - __clz(V);                       // V != 0: lane 0 is valid inside the loop
        const int nxtL = __shfl_sync(NRF_FULL_MASK, nxt, Lv);
        const bool occL = __shfl_sync(NRF_FULL_MASK, (int)occ, Lv) != 0;
        const float ttL = __shfl_sync(NRF_FULL_MASK, tt, Lv);
        if (nxtL < 32) break;                               // the successor exists in this chunk but is past `far`
        cur_t = tc;
        if (!occL) { while (cur_t < ttL) cur_t = march_next_t(c, cur_t); }
    }
    if (!WRITE && lane == 0) {
        rays[3 * (size_t)n + 0] = (int32_t)n;
        rays[3 * (size_t)n + 1] = 0;
        rays[3 * (size_t)n + 2] = (int32_t)count;
    }
}

// block-local exclusive scan of the per-ray counts (rays[:,2]) into rays[:,1]; block totals to block_sums
#define SCAN_BLOCK 1024
__global__ void __launch_bounds__(SCAN_BLOCK)
k_scan_counts_local(int32_t* __restrict__ rays, uint32_t N, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t warp_tot[SCAN_BLOCK / 32];
    const uint32_t n = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t v = n < N ? (uint32_t)rays[3 * (size_t)n + 2] : 0u;
    const uint32_t incl = warp_scan_add_u32(v, lane);
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = warp_tot[lane];
        const uint32_t wi = warp_scan_add_u32(w, lane);
        warp_tot[lane] = wi - w;
        if (lane == 31) block_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    if (n < N) rays[3 * (size_t)n + 1] = (int32_t)(warp_tot[warp] + incl - v);
}
__global__ void k_add_block_offsets_g(int32_t* __restrict__ rays, const uint32_t* __restrict__ block_offs, uint32_t N, uint32_t per_block) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    rays[3 * (size_t)n + 1] += (int32_t)block_offs[n / per_block];
}

// ------------------------------------------------------------------------------------------------
// march_rays_train: count -> scan -> write   (reference: one kernel, two passes + 2 atomics per ray,
// raymarching.cu:411-589; here offsets are the exclusive scan in ray order = one valid schedule of the
// reference's racing atomicAdd, and deterministic).
// ------------------------------------------------------------------------------------------------
#define MARCH_BLOCK 64
static int g_march_warp_per_ray = 1;   // 1: warp-per-ray walker (default), 0: thread-per-ray (also used for NDC)
NRF_EXPORT void nrf_march_set_mode(int warp_per_ray) { g_march_warp_per_ray = warp_per_ray ? 1 : 0; }

// rays[n] = (n, block-local exclusive offset, count); block_sums[blockIdx.x] = sum of counts
__global__ void __launch_bounds__(MARCH_BLOCK)
k_march_count(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid,
              float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
              const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ noises,
              int32_t* __restrict__ rays, uint32_t* __restrict__ block_sums) {
    const uint32_t n = blockIdx.x * MARCH_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t num_steps = 0;
    if (n < N) {
        MarchCtx c;
        march_init(c, rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H);
        const float far = fars[n];
        float t = march_t0(c, nears[n], noises ? noises[n] : 0.0f);
        float x, y, z, dt;
        while (t < far && num_steps < max_steps) {
            if (march_visit(c, grid, t, x, y, z, dt)) { num_steps++; t = __fadd_rn(t, dt); }
        }
    }
    __shared__ uint32_t warp_tot[MARCH_BLOCK / 32];
    const uint32_t incl = warp_scan_add_u32(num_steps, lane);
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < MARCH_BLOCK / 32; w++) if (w < warp) base += warp_tot[w];
    if (n < N) {
        rays[3 * (size_t)n + 0] = (int32_t)n;
        rays[3 * (size_t)n + 1] = (int32_t)(base + incl - num_steps);
        rays[3 * (size_t)n + 2] = (int32_t)num_steps;
    }
    if (threadIdx.x == MARCH_BLOCK - 1) block_sums[blockIdx.x] = base + incl;
}

// single block: exclusive scan of block_sums (in place) starting from counter[0]; updates the counter.
__global__ void __launch_bounds__(1024)
k_scan_block_sums(uint32_t* __restrict__ block_sums, uint32_t nb, int32_t* __restrict__ counter, uint32_t N) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = counter ? (uint32_t)counter[0] : 0u;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < nb; i0 += 1024) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t v = i < nb ? block_sums[i] : 0u;
        const uint32_t incl = warp_scan_add_u32(v, lane);
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = warp_tot[lane];
            const uint32_t wi = warp_scan_add_u32(w, lane);
            warp_tot[lane] = wi - w;   // exclusive over warps
        }
        __syncthreads();
        const uint32_t excl = carry + warp_tot[warp] + incl - v;
        if (i < nb) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && counter) {
        counter[0] = (int32_t)carry;
        counter[1] = counter[1] + (int32_t)N;
    }
}

__global__ void k_add_block_offsets(int32_t* __restrict__ rays, const uint32_t* __restrict__ block_offs, uint32_t N) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    rays[3 * (size_t)n + 1] += (int32_t)block_offs[n / MARCH_BLOCK];
}

NRF_EXPORT uint64_t nrf_march_scratch_bytes(uint32_t N) {
    return ((uint64_t)ceil_div_u32(N, MARCH_BLOCK) + 1024) * sizeof(uint32_t);
}

static int march_count_impl(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma, uint32_t max_steps,
                            uint32_t N, uint32_t C, uint32_t H, const float* nears, const float* fars, const float* noises, int32_t* rays,
                            int32_t* counter, void* scratch, float* t_stage, void* stream);

NRF_EXPORT int nrf_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                          float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                          const float* nears, const float* fars, const float* noises,
                                          int32_t* rays, int32_t* counter, void* scratch, void* stream) {
    return march_count_impl(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays, counter, scratch, nullptr,
                            stream);
}

// Count pass that also records the marching time of every emitted sample in t_stage [N, max_steps] f32 (warp-per-ray
// walker only); nrf_march_rays_train_emit then produces the samples without a second walk of the occupancy grid.
NRF_EXPORT int nrf_march_rays_train_count_staged(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                                 float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                                 const float* nears, const float* fars, const float* noises, int32_t* rays,
                                                 int32_t* counter, void* scratch, float* t_stage, void* stream) {
    if (!t_stage || !g_march_warp_per_ray) return NRF_E_UNSUPPORTED;
    return march_count_impl(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays, counter, scratch, t_stage,
                            stream);
}

static int march_count_impl(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma, uint32_t max_steps,
                            uint32_t N, uint32_t C, uint32_t H, const float* nears, const float* fars, const float* noises, int32_t* rays,
                            int32_t* counter, void* scratch, float* t_stage, void* stream) {
    if (N == 0) return NRF_OK;
    if (!rays_o || !rays_d || !grid || !nears || !fars || !rays || !scratch) return NRF_E_INVALID;
    if (C < 1 || C > 24 || H < 1 || H > 1024) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t* block_sums = (uint32_t*)scratch;
    if (g_march_warp_per_ray) {
        k_march_warp<false><<<ceil_div_u32(N, MW_WARPS), MW_WARPS * 32, 0, s>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H,
                                                                              0, 0, nears, fars, noises, rays, nullptr, nullptr, nullptr, t_stage);
        const uint32_t nb = ceil_div_u32(N, SCAN_BLOCK);
        k_scan_counts_local<<<nb, SCAN_BLOCK, 0, s>>>(rays, N, block_sums);
        k_scan_block_sums<<<1, 1024, 0, s>>>(block_sums, nb, counter, N);
        k_add_block_offsets_g<<<ceil_div_u32(N, 256), 256, 0, s>>>(rays, block_sums, N, SCAN_BLOCK);
        return nrf_check_launch();
    }
    const uint32_t nb = ceil_div_u32(N, MARCH_BLOCK);
    k_march_count<<<nb, MARCH_BLOCK, 0, s>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays, block_sums);
    k_scan_block_sums<<<1, 1024, 0, s>>>(block_sums, nb, counter, N);
    k_add_block_offsets<<<ceil_div_u32(N, 256), 256, 0, s>>>(rays, block_sums, N);
    return nrf_check_launch();
}

// second pass: thread per ray re-marches and writes its samples (raymarching.cu:519-588)
__global__ void __launch_bounds__(MARCH_BLOCK)
k_march_write(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ z_hats,
              const uint8_t* __restrict__ grid, float bound, float dt_gamma, uint32_t max_steps, bool is_ndc,
              uint32_t N, uint32_t C, uint32_t H, uint32_t M, uint32_t rows,
              const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ noises,
              const int32_t* __restrict__ rays, float* __restrict__ xyzs, float* __restrict__ dirs,
              float* __restrict__ deltas) {
    const uint32_t n = blockIdx.x * MARCH_BLOCK + threadIdx.x;
    if (n >= N) return;
    const uint32_t point_index = (uint32_t)rays[3 * (size_t)n + 1];
    const uint32_t num_steps = (uint32_t)rays[3 * (size_t)n + 2];
    if (num_steps == 0) return;
    if (point_index + num_steps >= M) {
        // dropped ray (raymarching.cu:517): its slots stay zero in the reference's zero-filled buffers
        for (uint32_t i = point_index; i < min(point_index + num_steps, rows); i++) {
            xyzs[3 * (size_t)i] = 0; xyzs[3 * (size_t)i + 1] = 0; xyzs[3 * (size_t)i + 2] = 0;
            dirs[3 * (size_t)i] = 0; dirs[3 * (size_t)i + 1] = 0; dirs[3 * (size_t)i + 2] = 0;
            reinterpret_cast<float4*>(deltas)[i] = make_float4(0, 0, 0, 0);
        }
        return;
    }
    MarchCtx c;
    march_init(c, rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H);
    const float far = fars[n];
    float t = march_t0(c, nears[n], noises ? noises[n] : 0.0f);
    float* px = xyzs + 3 * (size_t)point_index;
    float* pd = dirs + 3 * (size_t)point_index;
    float4* pl = reinterpret_cast<float4*>(deltas) + point_index;
    uint32_t step = 0;
    float last_t = t;
    float last_z = nrf_clamp(__fmaf_rn(c.dz, t, c.oz), c.nbound, c.bound);
    const float zh = is_ndc ? z_hats[n] : 1.0f;
    float x, y, z, dt;
    while (t < far && step < num_steps) {
        if (march_visit(c, grid, t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
            t = __fadd_rn(t, dt);
            float4 dl = make_float4(dt, __fsub_rn(t, last_t), 0.0f, 0.0f);
            last_t = t;
            if (is_ndc) {   // raymarching.cu:566-571 (train updates last_z = z)
                const float new_z = nrf_clamp(__fmaf_rn(c.dz, t, c.oz), c.nbound, c.bound);
                dl.z = (2 / (new_z - 1) - 2 / (z - 1)) / zh;
                dl.w = (2 / (new_z - 1) - 2 / (last_z - 1)) / zh;
                last_z = z;
            }
            *pl = dl;
            px += 3; pd += 3; pl += 1; step++;
        }
    }
}

__global__ void k_zero_rows(float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
                            uint32_t from, uint32_t to) {
    const uint32_t i = from + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= to) return;
    xyzs[3 * (size_t)i] = 0; xyzs[3 * (size_t)i + 1] = 0; xyzs[3 * (size_t)i + 2] = 0;
    dirs[3 * (size_t)i] = 0; dirs[3 * (size_t)i + 1] = 0; dirs[3 * (size_t)i + 2] = 0;
    reinterpret_cast<float4*>(deltas)[i] = make_float4(0, 0, 0, 0);
}

NRF_EXPORT int nrf_march_rays_train_write(const float* rays_o, const float* rays_d, const float* z_hats,
                                          const uint8_t* grid, float bound, float dt_gamma, uint32_t max_steps,
                                          int is_ndc, uint32_t N, uint32_t C, uint32_t H, uint32_t M, uint32_t rows,
                                          uint32_t zero_from, const float* nears, const float* fars,
                                          const float* noises, const int32_t* rays, float* xyzs, float* dirs,
                                          float* deltas, void* stream) {
    if (N == 0) return NRF_OK;
    if (!rays_o || !rays_d || !grid || !nears || !fars || !rays || !xyzs || !dirs || !deltas) return NRF_E_INVALID;
    if (is_ndc && !z_hats) return NRF_E_INVALID;
    if ((((uintptr_t)deltas) & 15) != 0) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (zero_from < rows) k_zero_rows<<<ceil_div_u32(rows - zero_from, 256), 256, 0, s>>>(xyzs, dirs, deltas, zero_from, rows);
    if (g_march_warp_per_ray && !is_ndc) {
        k_march_warp<true><<<ceil_div_u32(N, MW_WARPS), MW_WARPS * 32, 0, s>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M,
                                                                             rows, nears, fars, noises, const_cast<int32_t*>(rays), xyzs, dirs,
                                                                             deltas, nullptr);
        return nrf_check_launch();
    }
    k_march_write<<<ceil_div_u32(N, MARCH_BLOCK), MARCH_BLOCK, 0, s>>>(rays_o, rays_d, z_hats, grid, bound, dt_gamma, max_steps,
                                                                  is_ndc != 0, N, C, H, M, rows, nears, fars, noises, rays,
                                                                  xyzs, dirs, deltas);
    return nrf_check_launch();
}

// Second pass of the staged marching: warp per ray, lane j produces sample j, j + 32, ... from its recorded time.  The same
// float operations as the walker's write pass (k_march_warp<true>): x = clamp(fma(d, t, o)), dt = clamp(t dt_gamma),
// t_after = t + dt, delta = (dt, t_after - previous t_after | t0) -- bit-identical samples, no occupancy lookups.
__global__ void __launch_bounds__(MW_WARPS * 32)
k_march_emit(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
             uint32_t C, uint32_t H, uint32_t M, uint32_t rows, const float* __restrict__ nears, const float* __restrict__ noises,
             const int32_t* __restrict__ rays, const float* __restrict__ t_stage, float* __restrict__ xyzs, float* __restrict__ dirs,
             float* __restrict__ deltas) {
    const int lane = threadIdx.x & 31;
    const uint32_t n = blockIdx.x * MW_WARPS + (threadIdx.x >> 5);
    if (n >= N) return;
    const uint32_t point_index = (uint32_t)rays[3 * (size_t)n + 1], count = (uint32_t)rays[3 * (size_t)n + 2];
    if (count == 0) return;
    if (point_index + count >= M) {      // dropped ray (raymarching.cu:517): its slots stay zero
        for (uint32_t i = point_index + lane; i < min(point_index + count, rows); i += 32) {
            xyzs[3 * (size_t)i] = 0; xyzs[3 * (size_t)i + 1] = 0; xyzs[3 * (size_t)i + 2] = 0;
            if (dirs) { dirs[3 * (size_t)i] = 0; dirs[3 * (size_t)i + 1] = 0; dirs[3 * (size_t)i + 2] = 0; }
            reinterpret_cast<float4*>(deltas)[i] = make_float4(0, 0, 0, 0);
        }
        return;
    }
    MarchCtx c;
    march_init(c, rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, dt_gamma, max_steps, C, H);
    const float t0 = march_t0(c, nears[n], noises ? noises[n] : 0.0f);
    const float* ts = t_stage + (size_t)n * max_steps;
    for (uint32_t j = lane; j < count; j += 32) {
        const float t = __ldg(ts + j);
        const float x = nrf_clamp(__fmaf_rn(c.dx, t, c.ox), c.nbound, c.bound);
        const float y = nrf_clamp(__fmaf_rn(c.dy, t, c.oy), c.nbound, c.bound);
        const float z = nrf_clamp(__fmaf_rn(c.dz, t, c.oz), c.nbound, c.bound);
        const float dt = nrf_clamp(__fmul_rn(t, c.dt_gamma), c.dt_min, c.dt_max);
        const float t_after = __fadd_rn(t, dt);
        float prev = t0;
        if (j > 0) { const float tp = __ldg(ts + j - 1); prev = __fadd_rn(tp, nrf_clamp(__fmul_rn(tp, c.dt_gamma), c.dt_min, c.dt_max)); }
        const size_t o = (size_t)point_index + j;
        xyzs[3 * o] = x; xyzs[3 * o + 1] = y; xyzs[3 * o + 2] = z;
        if (dirs) { dirs[3 * o] = c.dx; dirs[3 * o + 1] = c.dy; dirs[3 * o + 2] = c.dz; }
        reinterpret_cast<float4*>(deltas)[o] = make_float4(dt, __fsub_rn(t_after, prev), 0.0f, 0.0f);
    }
}

// Same arguments as nrf_march_rays_train_write minus the occupancy grid / fars (not needed) plus the staged times.
NRF_EXPORT int nrf_march_rays_train_emit(const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                                         uint32_t C, uint32_t H, uint32_t M, uint32_t rows, uint32_t zero_from, const float* nears,
                                         const float* noises, const int32_t* rays, const float* t_stage, float* xyzs, float* dirs,
                                         float* deltas, void* stream) {
    if (N == 0) return NRF_OK;
    if (!rays_o || !rays_d || !nears || !rays || !t_stage || !xyzs || !dirs || !deltas) return NRF_E_INVALID;
    if ((((uintptr_t)deltas) & 15) != 0) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (zero_from < rows) k_zero_rows<<<ceil_div_u32(rows - zero_from, 256), 256, 0, s>>>(xyzs, dirs, deltas, zero_from, rows);
    k_march_emit<<<ceil_div_u32(N, MW_WARPS), MW_WARPS * 32, 0, s>>>(rays_o, rays_d, bound, dt_gamma, max_steps, N, C, H, M, rows, nears, noises,
                                                                   rays, t_stage, xyzs, dirs, deltas);
    return nrf_check_launch();
}

NRF_EXPORT int nrf_march_rays_train(const float* rays_o, const float* rays_d, const float* z_hats, const uint8_t* grid,
                                    float bound, float dt_gamma, uint32_t max_steps, int is_ndc, uint32_t N, uint32_t C,
                                    uint32_t H, uint32_t M, const float* nears, const float* fars, float* xyzs,
                                    float* dirs, float* deltas, int32_t* rays, int32_t* counter, const float* noises,
                                    void* scratch, void* stream) {
    int rc = nrf_march_rays_train_count(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, noises,
                                        rays, counter, scratch, stream);
    if (rc != NRF_OK) return rc;
    return nrf_march_rays_train_write(rays_o, rays_d, z_hats, grid, bound, dt_gamma, max_steps, is_ndc, N, C, H, M, M, M,
                                      nears, fars, noises, rays, xyzs, dirs, deltas, stream);
}

// ------------------------------------------------------------------------------------------------
// composite_rays_train forward / backward as per-ray segmented scans: one warp per ray, 32 consecutive
// samples per step (coalesced), transmittance by a multiplicative warp scan, early termination by ballot.
// Contract: raymarching.cu:807-879 (fwd), :905-986 (bwd); normative restatement SURVEY.md 8a.4.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float alpha_from(float sigma, float delta) {
    // 1 - __expf(-sigma*delta): the reference compiles to mul, mul by -log2(e), ex2.approx, sub
    float e;
    asm("ex2.approx.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(__fmul_rn(sigma, delta), -1.4426950408889634f)));
    return 1.0f - e;
}

#define COMP_WARPS 4

template <int CMAX>
__global__ void __launch_bounds__(COMP_WARPS * 32)
k_composite_train_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
                      const int32_t* __restrict__ rays, uint32_t M, uint32_t N, uint32_t C, float T_thresh, bool is_ndc,
                      float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[3 * (size_t)n], offset = (uint32_t)rays[3 * (size_t)n + 1], num_steps = (uint32_t)rays[3 * (size_t)n + 2];
    float acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; c++) acc[c] = 0.0f;
    float ws = 0.0f, d = 0.0f;
    if (num_steps != 0 && offset + num_steps < M) {
        float T_run = 1.0f, t_run = 0.0f;
        // software pipeline: the loads of block k + 1 are issued before the scans of block k (the blocks of a ray are a serial
        // chain -- T_run / t_run -- so without the prefetch every block pays a full memory round trip)
        float4 n_dl = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float n_sigma = 0.0f;
        float n_rgb[CMAX];
        auto fetch = [&](uint32_t base) {
            const uint32_t i = base + lane;
            n_dl = make_float4(0.0f, 0.0f, 0.0f, 0.0f); n_sigma = 0.0f;
#pragma unroll
            for (int c = 0; c < CMAX; c++) n_rgb[c] = 0.0f;
            if (i < num_steps) {
                n_dl = __ldg(reinterpret_cast<const float4*>(deltas) + offset + i);
                n_sigma = __ldg(sigmas + offset + i);
                const float* r = rgbs + (size_t)(offset + i) * C;
#pragma unroll
                for (int c = 0; c < CMAX; c++) if (c < (int)C) n_rgb[c] = __ldg(r + c);
            }
        };
        fetch(0);
        for (uint32_t base = 0; base < num_steps; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < num_steps;
            const float sigma = n_sigma, d0 = is_ndc ? n_dl.z : n_dl.x, d1 = is_ndc ? n_dl.w : n_dl.y;
            float rgb[CMAX];
#pragma unroll
            for (int c = 0; c < CMAX; c++) rgb[c] = n_rgb[c];
            if (base + 32 < num_steps) fetch(base + 32);
            const float alpha = valid ? alpha_from(sigma, d0) : 0.0f;
            const float P = warp_scan_mul(1.0f - alpha, lane);            // inclusive product
            float Pex = __shfl_up_sync(NRF_FULL_MASK, P, 1);
            if (lane == 0) Pex = 1.0f;
            const float T_before = T_run * Pex, T_after = T_run * P;
            const uint32_t term = __ballot_sync(NRF_FULL_MASK, valid && (T_after < T_thresh));
            const int last = term ? (__ffs(term) - 1) : 31;               // the terminating sample still counts (fwd)
            const float w = (valid && lane <= last) ? alpha * T_before : 0.0f;
            const float tsum = warp_scan_add(d1, lane);
            d = __fmaf_rn(w, t_run + tsum, d);
            ws += w;
            if (w != 0.0f) {
#pragma unroll
                for (int c = 0; c < CMAX; c++) if (c < (int)C) acc[c] = __fmaf_rn(w, rgb[c], acc[c]);
            }
            if (term) break;
            T_run = __shfl_sync(NRF_FULL_MASK, T_after, 31);
            t_run += __shfl_sync(NRF_FULL_MASK, tsum, 31);
        }
        ws = warp_sum(ws);
        d = warp_sum(d);
#pragma unroll
        for (int c = 0; c < CMAX; c++) acc[c] = warp_sum(acc[c]);
    }
    if (lane == 0) { weights_sum[index] = ws; depth[index] = d; }
#pragma unroll
    for (int c = 0; c < CMAX; c++) if (c < (int)C && lane == (c & 31)) image[(size_t)index * C + c] = acc[c];
}

// generic fallback (any C): one thread per ray, serial like the reference
__global__ void k_composite_train_fwd_serial(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                             const float* __restrict__ deltas, const int32_t* __restrict__ rays, uint32_t M,
                                             uint32_t N, uint32_t C, float T_thresh, bool is_ndc, float* __restrict__ weights_sum,
                                             float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[3 * (size_t)n], offset = (uint32_t)rays[3 * (size_t)n + 1], num_steps = (uint32_t)rays[3 * (size_t)n + 2];
    float* img = image + (size_t)index * C;
    for (uint32_t c = 0; c < C; c++) img[c] = 0.0f;
    if (num_steps == 0 || offset + num_steps >= M) { weights_sum[index] = 0; depth[index] = 0; return; }
    float T = 1.0f, ws = 0, t = 0, d = 0;
    for (uint32_t i = 0; i < num_steps; i++) {
        const float4 dl = reinterpret_cast<const float4*>(deltas)[offset + i];
        const float alpha = alpha_from(sigmas[offset + i], is_ndc ? dl.z : dl.x);
        const float w = alpha * T;
        const float* r = rgbs + (size_t)(offset + i) * C;
        for (uint32_t c = 0; c < C; c++) img[c] = __fmaf_rn(w, r[c], img[c]);
        t += is_ndc ? dl.w : dl.y;
        d = __fmaf_rn(w, t, d);
        ws += w;
        T *= 1.0f - alpha;
        if (T < T_thresh) break;
    }
    weights_sum[index] = ws; depth[index] = d;
}

NRF_EXPORT int nrf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                                const int32_t* rays, uint32_t M, uint32_t N, uint32_t C, float T_thresh,
                                                int is_ndc, float* weights_sum, float* depth, float* image, void* stream) {
    if (N == 0) return NRF_OK;
    if (!sigmas || !rgbs || !deltas || !rays || !weights_sum || !depth || !image) return NRF_E_INVALID;
    if ((((uintptr_t)deltas) & 15) != 0) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t nb = ceil_div_u32(N, COMP_WARPS);
    const bool ndc = is_ndc != 0;
#define LAUNCH_FWD(CM) k_composite_train_fwd<CM><<<nb, COMP_WARPS * 32, 0, s>>>(sigmas, rgbs, deltas, rays, M, N, C, T_thresh, ndc, weights_sum, depth, image)
    if (C <= 4) LAUNCH_FWD(4);
    else if (C <= 8) LAUNCH_FWD(8);
    else if (C <= 16) LAUNCH_FWD(16);
    else if (C <= 32) LAUNCH_FWD(32);
    else k_composite_train_fwd_serial<<<ceil_div_u32(N, 128), 128, 0, s>>>(sigmas, rgbs, deltas, rays, M, N, C, T_thresh, ndc, weights_sum, depth, image);
#undef LAUNCH_FWD
    return nrf_check_launch();
}

template <int CMAX>
__global__ void __launch_bounds__(COMP_WARPS * 32)
k_composite_train_bwd(const float* __restrict__ grad_ws, const float* __restrict__ grad_image, const float* __restrict__ sigmas,
                      const float* __restrict__ rgbs, const float* __restrict__ deltas, const int32_t* __restrict__ rays, bool is_ndc,
                      const float* __restrict__ weights_sum, const float* __restrict__ image, uint32_t M, uint32_t N, uint32_t C,
                      float T_thresh, float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs, bool write_zeros) {
    const uint32_t n = blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[3 * (size_t)n], offset = (uint32_t)rays[3 * (size_t)n + 1], num_steps = (uint32_t)rays[3 * (size_t)n + 2];
    // write_zeros: the caller did NOT zero-fill the gradients (the reference does, raymarching.py:339-340); every sample of this
    // ray that receives no gradient -- the terminating one, those behind it, a dropped ray's -- is written as zero here
    auto zero_rows = [&](uint32_t first, uint32_t last) {     // samples [first, last) of this ray, clipped to the buffer
        for (uint32_t i = first + lane; i < last && offset + i < M; i += 32) {
            grad_sigmas[offset + i] = 0.0f;
            float* gr = grad_rgbs + (size_t)(offset + i) * C;
            for (uint32_t c = 0; c < C; c++) gr[c] = 0.0f;
        }
    };
    if (num_steps == 0 || offset + num_steps >= M) {
        if (write_zeros && num_steps != 0) zero_rows(0, num_steps);
        return;
    }
    float g[CMAX];
    float G_total = 0.0f;     // sum_c g_c * image_c
#pragma unroll
    for (int c = 0; c < CMAX; c++) {
        g[c] = (c < (int)C) ? __ldg(grad_image + (size_t)index * C + c) : 0.0f;
        if (c < (int)C) G_total = __fmaf_rn(g[c], __ldg(image + (size_t)index * C + c), G_total);
    }
    const float ws_term = __ldg(grad_ws + index) * (1.0f - __ldg(weights_sum + index));
    float T_run = 1.0f, pre_run = 0.0f;
    // software pipeline, as in the forward: block k + 1 is in flight while block k is scanned
    float4 n_dl = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float n_sigma = 0.0f;
    float n_rgb[CMAX];
    auto fetch = [&](uint32_t base) {
        const uint32_t i = base + lane;
        n_dl = make_float4(0.0f, 0.0f, 0.0f, 0.0f); n_sigma = 0.0f;
#pragma unroll
        for (int c = 0; c < CMAX; c++) n_rgb[c] = 0.0f;
        if (i < num_steps) {
            n_dl = __ldg(reinterpret_cast<const float4*>(deltas) + offset + i);
            n_sigma = __ldg(sigmas + offset + i);
            const float* r = rgbs + (size_t)(offset + i) * C;
#pragma unroll
            for (int c = 0; c < CMAX; c++) if (c < (int)C) n_rgb[c] = __ldg(r + c);
        }
    };
    fetch(0);
    for (uint32_t base = 0; base < num_steps; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < num_steps;
        const float sigma = n_sigma, d0 = is_ndc ? n_dl.z : n_dl.x;
        float gdot = 0.0f;    // g . rgb_i
#pragma unroll
        for (int c = 0; c < CMAX; c++) gdot = __fmaf_rn(g[c], n_rgb[c], gdot);
        if (base + 32 < num_steps) fetch(base + 32);
        const float alpha = valid ? alpha_from(sigma, d0) : 0.0f;
        const float P = warp_scan_mul(1.0f - alpha, lane);
        float Pex = __shfl_up_sync(NRF_FULL_MASK, P, 1);
        if (lane == 0) Pex = 1.0f;
        const float T_before = T_run * Pex, T_after = T_run * P;
        const uint32_t term = __ballot_sync(NRF_FULL_MASK, valid && (T_after < T_thresh));
        const int last = term ? (__ffs(term) - 1) : 32;
        // the terminating sample contributes to rgbs_buf but gets no gradient (raymarching.cu:954-961)
        const float w = (valid && lane <= last) ? alpha * T_before : 0.0f;
        const float pre = pre_run + warp_scan_add(w * gdot, lane);        // sum_{j<=i} w_j (g . rgb_j)
        if (valid && lane < last) {
            float* gr = grad_rgbs + (size_t)(offset + i) * C;
#pragma unroll
            for (int c = 0; c < CMAX; c++) if (c < (int)C) gr[c] = g[c] * w;
            grad_sigmas[offset + i] = d0 * (__fmaf_rn(T_after, gdot, -(G_total - pre)) + ws_term);
        } else if (write_zeros && valid) {
            float* gr = grad_rgbs + (size_t)(offset + i) * C;
#pragma unroll
            for (int c = 0; c < CMAX; c++) if (c < (int)C) gr[c] = 0.0f;
            grad_sigmas[offset + i] = 0.0f;
        }
        if (term) {
            if (write_zeros) zero_rows(base + 32, num_steps);
            break;
        }
        T_run = __shfl_sync(NRF_FULL_MASK, T_after, 31);
        pre_run = __shfl_sync(NRF_FULL_MASK, pre, 31);
    }
}

__global__ void k_composite_train_bwd_serial(const float* __restrict__ grad_ws, const float* __restrict__ grad_image,
                                             const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                             const float* __restrict__ deltas, const int32_t* __restrict__ rays, bool is_ndc,
                                             const float* __restrict__ weights_sum, const float* __restrict__ image, uint32_t M,
                                             uint32_t N, uint32_t C, float T_thresh, float* __restrict__ grad_sigmas,
                                             float* __restrict__ grad_rgbs) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[3 * (size_t)n], offset = (uint32_t)rays[3 * (size_t)n + 1], num_steps = (uint32_t)rays[3 * (size_t)n + 2];
    if (num_steps == 0 || offset + num_steps >= M) return;
    const float* gi = grad_image + (size_t)index * C;
    const float* img = image + (size_t)index * C;
    float G_total = 0.0f;
    for (uint32_t c = 0; c < C; c++) G_total = __fmaf_rn(gi[c], img[c], G_total);
    const float ws_term = grad_ws[index] * (1.0f - weights_sum[index]);
    float T = 1.0f, pre = 0.0f;
    for (uint32_t i = 0; i < num_steps; i++) {
        const float4 dl = reinterpret_cast<const float4*>(deltas)[offset + i];
        const float d0 = is_ndc ? dl.z : dl.x;
        const float alpha = alpha_from(sigmas[offset + i], d0);
        const float w = alpha * T;
        const float* r = rgbs + (size_t)(offset + i) * C;
        float gdot = 0.0f;
        for (uint32_t c = 0; c < C; c++) gdot = __fmaf_rn(gi[c], r[c], gdot);
        pre = __fmaf_rn(w, gdot, pre);
        T *= 1.0f - alpha;
        if (T < T_thresh) break;
        float* gr = grad_rgbs + (size_t)(offset + i) * C;
        for (uint32_t c = 0; c < C; c++) gr[c] = gi[c] * w;
        grad_sigmas[offset + i] = d0 * (__fmaf_rn(T, gdot, -(G_total - pre)) + ws_term);
    }
}

NRF_EXPORT int nrf_composite_rays_train_backward_ex(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                                    const float* rgbs, const float* deltas, const int32_t* rays, int is_ndc,
                                                    const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                                    uint32_t C, float T_thresh, float* grad_sigmas, float* grad_rgbs,
                                                    int write_zeros, void* stream);

NRF_EXPORT int nrf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                                 const float* rgbs, const float* deltas, const int32_t* rays, int is_ndc,
                                                 const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                                 uint32_t C, float T_thresh, float* grad_sigmas, float* grad_rgbs,
                                                 void* stream) {
    return nrf_composite_rays_train_backward_ex(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, is_ndc, weights_sum, image,
                                                M, N, C, T_thresh, grad_sigmas, grad_rgbs, 0, stream);
}

// write_zeros != 0: grad_sigmas / grad_rgbs need not be zero-filled by the caller -- every sample slot that belongs to a ray
// is written (gradient or zero); slots that belong to NO ray (padding rows) remain the caller's business.  The warp-per-ray
// kernels only (C <= 32).
NRF_EXPORT int nrf_composite_rays_train_backward_ex(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                                    const float* rgbs, const float* deltas, const int32_t* rays, int is_ndc,
                                                    const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                                    uint32_t C, float T_thresh, float* grad_sigmas, float* grad_rgbs,
                                                    int write_zeros, void* stream) {
    if (N == 0) return NRF_OK;
    if (write_zeros && C > 32) return NRF_E_UNSUPPORTED;
    const bool wz = write_zeros != 0;
    if (!grad_weights_sum || !grad_image || !sigmas || !rgbs || !deltas || !rays || !weights_sum || !image || !grad_sigmas || !grad_rgbs)
        return NRF_E_INVALID;
    if ((((uintptr_t)deltas) & 15) != 0) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t nb = ceil_div_u32(N, COMP_WARPS);
    const bool ndc = is_ndc != 0;
#define LAUNCH_BWD(CM) k_composite_train_bwd<CM><<<nb, COMP_WARPS * 32, 0, s>>>(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, ndc, weights_sum, image, M, N, C, T_thresh, grad_sigmas, grad_rgbs, wz)
    if (C <= 4) LAUNCH_BWD(4);
    else if (C <= 8) LAUNCH_BWD(8);
    else if (C <= 16) LAUNCH_BWD(16);
    else if (C <= 32) LAUNCH_BWD(32);
    else k_composite_train_bwd_serial<<<ceil_div_u32(N, 128), 128, 0, s>>>(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, ndc, weights_sum, image, M, N, C, T_thresh, grad_sigmas, grad_rgbs);
#undef LAUNCH_BWD
    return nrf_check_launch();
}

// ------------------------------------------------------------------------------------------------
// inference: march_rays (raymarching.cu:1005-1120) / composite_rays (:1134-1231) / alive-ray compaction
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive, const float* __restrict__ rays_t,
             const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ z_hats, float bound,
             float dt_gamma, uint32_t max_steps, bool is_ndc, uint32_t C, uint32_t H, const uint8_t* __restrict__ grid,
             const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
             const float* __restrict__ noises, uint32_t Mpad, bool zero_fill, const int32_t* __restrict__ ctl) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (ctl) {      // device-driven loop (nrf_march_rays_dev): the launch is sized for the cap, the counts live on the device
        n_alive = (uint32_t)ctl[0]; n_step = (uint32_t)ctl[1];
        if (n >= n_alive) return;
    }
    if (n >= n_alive) {
        // padding rows [n_alive*n_step, Mpad): spread over the surplus threads of the grid
        if (zero_fill) {
            const uint32_t total_threads = gridDim.x * blockDim.x;
            for (uint32_t i = n_alive * n_step + (n - n_alive); i < Mpad; i += total_threads - n_alive) {
                xyzs[3 * (size_t)i] = 0; xyzs[3 * (size_t)i + 1] = 0; xyzs[3 * (size_t)i + 2] = 0;
                if (dirs) { dirs[3 * (size_t)i] = 0; dirs[3 * (size_t)i + 1] = 0; dirs[3 * (size_t)i + 2] = 0; }
                reinterpret_cast<float4*>(deltas)[i] = make_float4(0, 0, 0, 0);
            }
        }
        return;
    }
    const int32_t index = rays_alive[n];
    float* px = xyzs + 3 * (size_t)n * n_step;
    float* pd = dirs ? dirs + 3 * (size_t)n * n_step : nullptr;      // NULL: the field has no direction input
    float4* pl = reinterpret_cast<float4*>(deltas) + (size_t)n * n_step;
    uint32_t step = 0;
    if (index < 0) {
        // extension: a dead slot (-1) left in the list by a caller that compacts without reading the count back every
        // iteration (Renderer.render_test of the host mirror); its rows are zero so that composite_rays skips them
        for (; step < n_step; step++) {
            px[0] = 0; px[1] = 0; px[2] = 0; if (pd) { pd[0] = 0; pd[1] = 0; pd[2] = 0; }
            *pl = make_float4(0, 0, 0, 0);
            px += 3; if (pd) pd += 3; pl += 1;
        }
        return;
    }
    MarchCtx c;
    march_init(c, rays_o + 3 * (size_t)index, rays_d + 3 * (size_t)index, bound, dt_gamma, max_steps, C, H);
    float t = rays_t[(size_t)index * (is_ndc ? 2 : 1)];
    const float far = fars[index];
    t = march_t0(c, t, noises ? noises[n] : 0.0f);
    float last_t = t;
    float last_z = nrf_clamp(__fmaf_rn(c.dz, t, c.oz), c.nbound, c.bound);
    const float zh = is_ndc ? z_hats[index] : 1.0f;
    float x, y, z, dt;
    while (t < far && step < n_step) {
        if (march_visit(c, grid, t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            if (pd) { pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz; }
            t = __fadd_rn(t, dt);
            float4 dl = make_float4(dt, __fsub_rn(t, last_t), 0.0f, 0.0f);
            if (is_ndc) {   // raymarching.cu:1094-1099 (inference updates last_z = new_z)
                const float new_z = nrf_clamp(__fmaf_rn(c.dz, t, c.oz), c.nbound, c.bound);
                dl.z = (2 / (new_z - 1) - 2 / (z - 1)) / zh;
                dl.w = (2 / (new_z - 1) - 2 / (last_z - 1)) / zh;
                last_z = new_z;
            }
            last_t = t;
            *pl = dl;
            px += 3; if (pd) pd += 3; pl += 1; step++;
        }
    }
    if (zero_fill) {
        for (; step < n_step; step++) {
            px[0] = 0; px[1] = 0; px[2] = 0; if (pd) { pd[0] = 0; pd[1] = 0; pd[2] = 0; }
            *pl = make_float4(0, 0, 0, 0);
            px += 3; if (pd) pd += 3; pl += 1;
        }
    }
}

// Block-staged form of k_march_rays for n_step > 1 (non-NDC): a block's 128 consecutive alive slots own ONE contiguous span
// of the output (rows [128 b n_step, 128 (b + 1) n_step)), so every ray writes its samples into shared memory and the block
// then stores the span with full 16-byte coalesced writes -- the per-thread version writes 12-byte pieces n_step * 12 bytes
// apart.  Same marching state machine, same values.
#define MRB_THREADS 128
#define MRB_MAX_STEP 8
__global__ void __launch_bounds__(MRB_THREADS)
k_march_rays_blk(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive, const float* __restrict__ rays_t,
                 const float* __restrict__ rays_o, const float* __restrict__ rays_d, float bound, float dt_gamma, uint32_t max_steps,
                 uint32_t C, uint32_t H, const uint8_t* __restrict__ grid, const float* __restrict__ fars, float* __restrict__ xyzs,
                 float* __restrict__ dirs, float* __restrict__ deltas, const float* __restrict__ noises, const int32_t* __restrict__ ctl) {
    __shared__ __align__(16) float s_xyz[MRB_THREADS * MRB_MAX_STEP * 3];
    __shared__ __align__(16) float s_dir[MRB_THREADS * MRB_MAX_STEP * 3];
    __shared__ __align__(16) float4 s_del[MRB_THREADS * MRB_MAX_STEP];
    if (ctl) { n_alive = (uint32_t)ctl[0]; n_step = (uint32_t)ctl[1]; }
    const uint32_t n0 = blockIdx.x * MRB_THREADS;
    if (n0 >= n_alive) return;
    const uint32_t n = n0 + threadIdx.x;
    const uint32_t rays_here = min((uint32_t)MRB_THREADS, n_alive - n0);
    if (ctl && ctl[7] != 0) {
        // STEP-MAJOR rows (device-driven loop only, ctl[7] = 1): sample s of alive slot n lives in row s * n_alive + n, so
        // consecutive lanes write -- and composite_rays later reads -- consecutive rows: coalesced without staging
        if (n >= n_alive) return;
        const int32_t index = rays_alive[n];
        uint32_t step = 0;
        if (index >= 0) {
            MarchCtx c;
            march_init(c, rays_o + 3 * (size_t)index, rays_d + 3 * (size_t)index, bound, dt_gamma, max_steps, C, H);
            float t = rays_t[index];
            const float far = fars[index];
            t = march_t0(c, t, noises ? noises[n] : 0.0f);
            float last_t = t;
            float x, y, z, dt;
            while (t < far && step < n_step) {
                if (march_visit(c, grid, t, x, y, z, dt)) {
                    const size_t row = (size_t)step * n_alive + n;
                    xyzs[3 * row] = x; xyzs[3 * row + 1] = y; xyzs[3 * row + 2] = z;
                    if (dirs) { dirs[3 * row] = c.dx; dirs[3 * row + 1] = c.dy; dirs[3 * row + 2] = c.dz; }
                    t = __fadd_rn(t, dt);
                    reinterpret_cast<float4*>(deltas)[row] = make_float4(dt, __fsub_rn(t, last_t), 0.0f, 0.0f);
                    last_t = t;
                    step++;
                }
            }
        }
        for (; step < n_step; step++) {
            const size_t row = (size_t)step * n_alive + n;
            xyzs[3 * row] = 0; xyzs[3 * row + 1] = 0; xyzs[3 * row + 2] = 0;
            if (dirs) { dirs[3 * row] = 0; dirs[3 * row + 1] = 0; dirs[3 * row + 2] = 0; }
            reinterpret_cast<float4*>(deltas)[row] = make_float4(0, 0, 0, 0);
        }
        return;
    }
    if (n < n_alive) {
        const int32_t index = rays_alive[n];
        float* px = s_xyz + 3 * threadIdx.x * n_step;
        float* pd = s_dir + 3 * threadIdx.x * n_step;
        float4* pl = s_del + threadIdx.x * n_step;
        uint32_t step = 0;
        if (index >= 0) {
            MarchCtx c;
            march_init(c, rays_o + 3 * (size_t)index, rays_d + 3 * (size_t)index, bound, dt_gamma, max_steps, C, H);
            float t = rays_t[index];
            const float far = fars[index];
            t = march_t0(c, t, noises ? noises[n] : 0.0f);
            float last_t = t;
            float x, y, z, dt;
            while (t < far && step < n_step) {
                if (march_visit(c, grid, t, x, y, z, dt)) {
                    px[0] = x; px[1] = y; px[2] = z;
                    if (dirs) { pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz; }
                    t = __fadd_rn(t, dt);
                    *pl = make_float4(dt, __fsub_rn(t, last_t), 0.0f, 0.0f);
                    last_t = t;
                    px += 3; pd += 3; pl += 1; step++;
                }
            }
        }
        for (; step < n_step; step++) {       // unused slots (ray left the volume / dead slot) are zero: composite_rays stops at them
            px[0] = 0; px[1] = 0; px[2] = 0;
            if (dirs) { pd[0] = 0; pd[1] = 0; pd[2] = 0; }
            *pl = make_float4(0, 0, 0, 0);
            px += 3; pd += 3; pl += 1;
        }
    }
    __syncthreads();
    // coalesced copy-out of the block's span: rows [n0 * n_step, (n0 + rays_here) * n_step)
    const size_t row0 = (size_t)n0 * n_step;
    const uint32_t rows = rays_here * n_step;
    float4* gd = reinterpret_cast<float4*>(deltas) + row0;
    for (uint32_t i = threadIdx.x; i < rows; i += MRB_THREADS) gd[i] = s_del[i];
    // 3 floats per row: the span starts at float 3 * row0 (16-byte aligned because row0 is a multiple of 128 * n_step ... of 4)
    const uint32_t nf = rows * 3;
    float* gx = xyzs + 3 * row0;
    if ((((uintptr_t)gx) & 15) == 0) {
        for (uint32_t i = threadIdx.x; i < nf / 4; i += MRB_THREADS) reinterpret_cast<float4*>(gx)[i] = reinterpret_cast<const float4*>(s_xyz)[i];
        for (uint32_t i = (nf / 4) * 4 + threadIdx.x; i < nf; i += MRB_THREADS) gx[i] = s_xyz[i];
    } else {
        for (uint32_t i = threadIdx.x; i < nf; i += MRB_THREADS) gx[i] = s_xyz[i];
    }
    if (dirs) {
        float* gdr = dirs + 3 * row0;
        if ((((uintptr_t)gdr) & 15) == 0) {
            for (uint32_t i = threadIdx.x; i < nf / 4; i += MRB_THREADS) reinterpret_cast<float4*>(gdr)[i] = reinterpret_cast<const float4*>(s_dir)[i];
            for (uint32_t i = (nf / 4) * 4 + threadIdx.x; i < nf; i += MRB_THREADS) gdr[i] = s_dir[i];
        } else {
            for (uint32_t i = threadIdx.x; i < nf; i += MRB_THREADS) gdr[i] = s_dir[i];
        }
    }
}

NRF_EXPORT int nrf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                              const float* rays_o, const float* rays_d, const float* z_hats, float bound, float dt_gamma,
                              uint32_t max_steps, int is_ndc, uint32_t C, uint32_t H, const uint8_t* grid, const float* nears,
                              const float* fars, float* xyzs, float* dirs, float* deltas, const float* noises,
                              uint32_t Mpad, int zero_fill, void* stream) {
    (void)nears;
    if (!xyzs || !dirs || !deltas) return NRF_E_INVALID;
    if ((((uintptr_t)deltas) & 15) != 0) return NRF_E_INVALID;
    if (n_alive && (!rays_alive || !rays_t || !rays_o || !rays_d || !grid || !fars)) return NRF_E_INVALID;
    if (is_ndc && !z_hats) return NRF_E_INVALID;
    if ((uint64_t)n_alive * n_step > Mpad) return NRF_E_INVALID;
    const uint32_t pad = Mpad - n_alive * n_step;
    if (!is_ndc && n_step >= 2 && n_step <= MRB_MAX_STEP && n_alive > 0) {
        k_march_rays_blk<<<ceil_div_u32(n_alive, MRB_THREADS), MRB_THREADS, 0, (cudaStream_t)stream>>>(n_alive, n_step, rays_alive, rays_t, rays_o,
                                                                                                   rays_d, bound, dt_gamma, max_steps, C, H, grid,
                                                                                                   fars, xyzs, dirs, deltas, noises, nullptr);
        if (zero_fill && pad) k_zero_rows<<<ceil_div_u32(pad, 256), 256, 0, (cudaStream_t)stream>>>(xyzs, dirs, deltas, n_alive * n_step, Mpad);
        return nrf_check_launch();
    }
    const uint32_t threads = n_alive + (zero_fill ? min(pad, 4096u) : 0u);
    if (threads == 0) return NRF_OK;
    k_march_rays<<<ceil_div_u32(threads, 128), 128, 0, (cudaStream_t)stream>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d,
                                                                             z_hats, bound, dt_gamma, max_steps, is_ndc != 0, C, H,
                                                                             grid, fars, xyzs, dirs, deltas, noises, Mpad,
                                                                             zero_fill != 0, nullptr);
    return nrf_check_launch();
}

// Device-driven inference loop (SURVEY.md 8f NEXT-2): the same kernels, but n_alive / n_step come from a control block in
// device memory, ctl = int32[8] {n_alive, n_step, n_rows = n_alive * n_step, steps_done, N, max_steps, iterations, -}, which
// nrf_compact_alive_dev advances at the end of every iteration.  Every launch is sized for the cap (n_alive_cap rays,
// n_alive * n_step <= N rows), so an iteration is shape-static and a pair of them can be captured in a CUDA graph and
// replayed with no host involvement.
NRF_EXPORT int nrf_march_rays_dev(const int32_t* ctl, uint32_t n_alive_cap, const int32_t* rays_alive, const float* rays_t,
                                  const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                                  uint32_t H, const uint8_t* grid, const float* fars, float* xyzs, float* dirs, float* deltas,
                                  void* stream) {
    if (!ctl || !rays_alive || !rays_t || !rays_o || !rays_d || !grid || !fars || !xyzs || !deltas) return NRF_E_INVALID;      // dirs may be NULL
    if ((((uintptr_t)deltas) & 15) != 0) return NRF_E_INVALID;
    if (n_alive_cap == 0) return NRF_OK;
    // n_step (device-side, <= 8) is not known to the host: the block-staged kernel handles every n_step
    k_march_rays_blk<<<ceil_div_u32(n_alive_cap, MRB_THREADS), MRB_THREADS, 0, (cudaStream_t)stream>>>(n_alive_cap, 1, rays_alive, rays_t, rays_o, rays_d,
                                                                                                  bound, dt_gamma, max_steps, C, H, grid, fars, xyzs,
                                                                                                  dirs, deltas, nullptr, ctl);
    return nrf_check_launch();
}

__global__ void __launch_bounds__(128)
k_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* __restrict__ rays_alive, float* __restrict__ rays_t,
                 const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas, uint32_t C,
                 bool is_ndc, float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image,
                 const int32_t* __restrict__ ctl) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (ctl) { n_alive = (uint32_t)ctl[0]; n_step = (uint32_t)ctl[1]; }
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    if (index < 0) return;          // dead slot (see k_march_rays)
    // row of sample `step` of this slot: n * n_step + step (the reference's ray-major layout) or, in the device-driven loop
    // with ctl[7] = 1, step * n_alive + n (step-major: lanes read consecutive rows)
    const bool step_major = ctl && ctl[7] != 0;
    const size_t row0 = step_major ? (size_t)n : (size_t)n * n_step;
    const size_t row_inc = step_major ? (size_t)n_alive : 1;
    const float* s = sigmas + row0;
    const float* r = rgbs + row0 * C;
    const float4* dl4 = reinterpret_cast<const float4*>(deltas) + row0;
    float* rt = rays_t + (size_t)index * (is_ndc ? 2 : 1);
    float* img = image + (size_t)index * C;
    float t_rm = 0.0f, t_phy;
    if (is_ndc) { t_rm = rt[0]; t_phy = rt[1]; } else { t_phy = rt[0]; }
    float weight_sum = weights_sum[index], d = depth[index];
    uint32_t step = 0;
    if (C <= 16) {
        // the ray's image row stays in registers across its n_step samples (the reference does a global read-modify-write
        // per sample per channel, raymarching.cu:1196-1198); same fma order per channel -> same bits
        float acc[16];
#pragma unroll
        for (int c = 0; c < 16; c++) acc[c] = ((uint32_t)c < C) ? img[c] : 0.0f;
        while (step < n_step) {
            const float4 dl = __ldg(dl4 + step * row_inc);
            if (dl.x == 0.0f) break;
            const float alpha = alpha_from(__ldg(s + step * row_inc), is_ndc ? dl.z : dl.x);
            const float T = 1.0f - weight_sum;
            const float w = alpha * T;
            weight_sum += w;
            if (is_ndc) { t_rm += dl.y; t_phy += dl.w; } else { t_phy += dl.y; }
            d = __fmaf_rn(w, t_phy, d);
            const float* rr = r + (size_t)step * row_inc * C;
#pragma unroll
            for (int c = 0; c < 16; c++) if ((uint32_t)c < C) acc[c] = __fmaf_rn(w, __ldg(rr + c), acc[c]);
            if (T < T_thresh) break;
            step++;
        }
#pragma unroll
        for (int c = 0; c < 16; c++) if ((uint32_t)c < C) img[c] = acc[c];
    } else {
        while (step < n_step) {
            const float4 dl = __ldg(dl4 + step * row_inc);
            if (dl.x == 0.0f) break;
            const float alpha = alpha_from(__ldg(s + step * row_inc), is_ndc ? dl.z : dl.x);
            const float T = 1.0f - weight_sum;
            const float w = alpha * T;
            weight_sum += w;
            if (is_ndc) { t_rm += dl.y; t_phy += dl.w; } else { t_phy += dl.y; }
            d = __fmaf_rn(w, t_phy, d);
            const float* rr = r + (size_t)step * row_inc * C;
            for (uint32_t c = 0; c < C; c++) img[c] = __fmaf_rn(w, __ldg(rr + c), img[c]);
            if (T < T_thresh) break;
            step++;
        }
    }
    if (step < n_step) rays_alive[n] = -1;
    else { if (is_ndc) { rt[0] = t_rm; rt[1] = t_phy; } else rt[0] = t_phy; }
    weights_sum[index] = weight_sum;
    depth[index] = d;
}

NRF_EXPORT int nrf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                                  const float* sigmas, const float* rgbs, const float* deltas, uint32_t C, int is_ndc,
                                  float* weights_sum, float* depth, float* image, void* stream) {
    if (n_alive == 0) return NRF_OK;
    if (!rays_alive || !rays_t || !sigmas || !rgbs || !deltas || !weights_sum || !depth || !image) return NRF_E_INVALID;
    if ((((uintptr_t)deltas) & 15) != 0) return NRF_E_INVALID;
    k_composite_rays<<<ceil_div_u32(n_alive, 128), 128, 0, (cudaStream_t)stream>>>(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas,
                                                                                  rgbs, deltas, C, is_ndc != 0, weights_sum, depth, image,
                                                                                  nullptr);
    return nrf_check_launch();
}

NRF_EXPORT int nrf_composite_rays_dev(const int32_t* ctl, uint32_t n_alive_cap, float T_thresh, int32_t* rays_alive, float* rays_t,
                                      const float* sigmas, const float* rgbs, const float* deltas, uint32_t C, float* weights_sum,
                                      float* depth, float* image, void* stream) {
    if (!ctl || !rays_alive || !rays_t || !sigmas || !rgbs || !deltas || !weights_sum || !depth || !image) return NRF_E_INVALID;
    if ((((uintptr_t)deltas) & 15) != 0) return NRF_E_INVALID;
    if (n_alive_cap == 0) return NRF_OK;
    k_composite_rays<<<ceil_div_u32(n_alive_cap, 128), 128, 0, (cudaStream_t)stream>>>(n_alive_cap, 1, T_thresh, rays_alive, rays_t, sigmas, rgbs,
                                                                                     deltas, C, false, weights_sum, depth, image, ctl);
    return nrf_check_launch();
}

// stable compaction of the non-negative entries (replaces rays_alive[rays_alive >= 0], renderer.py:284)
#define COMPACT_BLOCK 256
__global__ void __launch_bounds__(COMPACT_BLOCK)
k_compact_count(const int32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ block_sums, const int32_t* __restrict__ ctl) {
    if (ctl) n = (uint32_t)ctl[0];
    const uint32_t i = blockIdx.x * COMPACT_BLOCK + threadIdx.x;
    const bool keep = i < n && in[i] >= 0;
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = (uint32_t)c;
}
__global__ void k_compact_finish(const uint32_t* __restrict__ block_sums_scanned_end, int32_t* __restrict__ n_out) {
    *n_out = (int32_t)*block_sums_scanned_end;
}
__global__ void __launch_bounds__(COMPACT_BLOCK)
k_compact_write(const int32_t* __restrict__ in, uint32_t n, const uint32_t* __restrict__ block_offs, int32_t* __restrict__ out,
                const int32_t* __restrict__ ctl) {
    __shared__ uint32_t warp_tot[COMPACT_BLOCK / 32];
    if (ctl) n = (uint32_t)ctl[0];
    const uint32_t i = blockIdx.x * COMPACT_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t v = i < n ? in[i] : -1;
    const bool keep = v >= 0;
    const uint32_t m = __ballot_sync(NRF_FULL_MASK, keep);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    uint32_t base = block_offs[blockIdx.x];
    for (int w = 0; w < warp; w++) base += warp_tot[w];
    if (keep) out[base + __popc(m & ((1u << lane) - 1u))] = v;
}

NRF_EXPORT int nrf_compact_alive(const int32_t* in, uint32_t n, int32_t* out, int32_t* n_out, void* scratch, void* stream) {
    if (!n_out) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return cudaMemsetAsync(n_out, 0, sizeof(int32_t), s) == cudaSuccess ? NRF_OK : NRF_E_CUDA;
    if (!in || !out || !scratch) return NRF_E_INVALID;
    const uint32_t nb = ceil_div_u32(n, COMPACT_BLOCK);
    uint32_t* block_sums = (uint32_t*)scratch;      // nb entries + 1 (grand total written by the scan via counter)
    // scratch sized by nrf_march_scratch_bytes(n): ceil(n/64)+1024 words >= nb + 2
    int32_t* total = (int32_t*)(block_sums + nb);
    cudaMemsetAsync(total, 0, 2 * sizeof(int32_t), s);
    cudaMemsetAsync(out, 0xFF, (size_t)n * sizeof(int32_t), s);      // entries past the new count read -1 (dead slots)
    k_compact_count<<<nb, COMPACT_BLOCK, 0, s>>>(in, n, block_sums, nullptr);
    k_scan_block_sums<<<1, 1024, 0, s>>>(block_sums, nb, total, 0);
    k_compact_write<<<nb, COMPACT_BLOCK, 0, s>>>(in, n, block_sums, out, nullptr);
    cudaMemcpyAsync(n_out, total, sizeof(int32_t), cudaMemcpyDeviceToDevice, s);
    return nrf_check_launch();
}

// end of one iteration of the device-driven loop: the new alive count, the steps taken so far, the next n_step
// (renderer.py:253: max(min(N // n_alive, 8), 1); ctl[4] is the row budget -- N for the reference's schedule) and the row count the
// next iteration's kernels will process
__global__ void k_ctl_update(int32_t* __restrict__ ctl, const int32_t* __restrict__ total) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int n_alive = total[0];
    const int steps = ctl[3] + ctl[1];
    if (steps >= ctl[5]) n_alive = 0;                    // max_steps reached (renderer.py:249)
    const int n_step = n_alive > 0 ? max(min(ctl[4] / n_alive, 8), 1) : 0;
    ctl[0] = n_alive; ctl[1] = n_step; ctl[2] = n_alive * n_step; ctl[3] = steps; ctl[6] += 1;
}

NRF_EXPORT int nrf_compact_alive_dev(int32_t* ctl, uint32_t n_cap, const int32_t* in, int32_t* out, void* scratch, void* stream) {
    if (!ctl || !in || !out || !scratch) return NRF_E_INVALID;
    if (n_cap == 0) return NRF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t nb = ceil_div_u32(n_cap, COMPACT_BLOCK);
    uint32_t* block_sums = (uint32_t*)scratch;
    int32_t* total = (int32_t*)(block_sums + nb);
    cudaMemsetAsync(total, 0, 2 * sizeof(int32_t), s);
    cudaMemsetAsync(out, 0xFF, (size_t)n_cap * sizeof(int32_t), s);
    k_compact_count<<<nb, COMPACT_BLOCK, 0, s>>>(in, n_cap, block_sums, ctl);
    k_scan_block_sums<<<1, 1024, 0, s>>>(block_sums, nb, total, 0);
    k_compact_write<<<nb, COMPACT_BLOCK, 0, s>>>(in, n_cap, block_sums, out, ctl);
    k_ctl_update<<<1, 32, 0, s>>>(ctl, total);
    return nrf_check_launch();
}
