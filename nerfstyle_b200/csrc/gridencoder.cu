// Multiresolution hash-grid encoder for sm_100a (C ABI in include/nerfstyle_b200.h).
//
// Behavioural contract: /root/reference/gridencoder/src/gridencoder.cu (cited per function).  Written from
// scratch: the forward is a vectorised 8-corner gather that writes point-major [B, L*C] directly (the
// reference writes [L,B,C] and pays a permute copy, grid.py:58).  The backward exists in two forms: a walk (one lane follows
// a chunk of consecutive samples at one level and keeps the current cell's sums in registers; cells it leaves are parked in a
// shared-memory ring and reduced by full warps -- k_grid_bwd_walkq, the default for f32 gradient tables) and the round-1
// thread-per-sample kernel with warp aggregation (lanes that fall into the same cell are reduced with shuffles before ONE
// vector atomic per corner -- k_grid_bwd_d3c2, the fallback and the f16-gradient-table reference mode).
// Index arithmetic (hash, strides, modulo) is bit-exact with the reference; the float interpolation uses
// explicit __fmaf_rn at the reference's contraction points so f32 outputs are bit-exact too.
#include "common.cuh"

#define GRID_MAX_LEVELS 32
#define GRID_BLOCK 256
#define TR_STRIDE 36      // floats per lane of the backward's transpose buffer: 32 products + 4 pad (conflict-free 16-byte stores)
#define TR_BYTES (GRID_BLOCK / 32 * 32 * TR_STRIDE * 4)

__host__ __device__ __forceinline__ uint32_t prime_of(int i) {
    // gridencoder.cu:42
    return i == 0 ? 1u : i == 1 ? 2654435761u : i == 2 ? 805459861u : i == 3 ? 3674653429u : i == 4 ? 2097192037u
         : i == 5 ? 1434869437u : 2165219737u;
}

// per-level constants, computed once per block into shared memory
struct LevelP {
    uint32_t offset;        // row offset of the level
    uint32_t size;          // hashmap_size (rows)
    uint32_t resolution;
    float scale;
    uint32_t stride[3];     // dense strides actually accumulated (0 = dimension not accumulated)
    uint32_t style_term;    // style*stride when the style stride fits (dense), else 0
    uint32_t hash_style;    // style * primes[D]
    uint32_t use_hash;
    uint32_t mul, shift;    // exact x / size for any uint32 x:  t=umulhi(x,mul); q=(t+((x-t)>>1))>>(shift-1)
    uint32_t pow2mask;      // size-1 when size is a power of two, else 0
};

// gridencoder.cu:55-80 stride logic, and :134-138 resolution/scale, for D dims
__device__ __forceinline__ void level_setup(LevelP& p, const int32_t* __restrict__ offsets, uint32_t level, uint32_t D,
                                            float S, uint32_t H, uint32_t gridtype, bool align_corners, uint32_t style) {
    p.offset = (uint32_t)offsets[level];
    p.size = (uint32_t)(offsets[level + 1] - offsets[level]);
    p.resolution = (uint32_t)floorf(exp2f((float)level * S) * (float)H);
    p.scale = (float)(p.resolution - (align_corners ? 0u : 1u));
    uint32_t stride = 1;
    p.stride[0] = p.stride[1] = p.stride[2] = 0;
    for (uint32_t d = 0; d < D && stride <= p.size; d++) {
        if (d < 3) p.stride[d] = stride;
        stride *= (p.resolution + 1);
    }
    p.style_term = 0;
    if (stride <= p.size) { p.style_term = style * stride; stride *= 512u; }
    p.use_hash = (gridtype == 0 && stride > p.size) ? 1u : 0u;
    p.hash_style = style * prime_of((int)D);
    p.pow2mask = ((p.size & (p.size - 1)) == 0) ? p.size - 1 : 0u;
    // Granlund-Montgomery round-up magic for exact unsigned division by p.size
    uint32_t l = 0;
    while ((1ull << l) < (uint64_t)p.size) l++;
    p.shift = l;
    p.mul = (uint32_t)((((1ull << 32) * ((1ull << l) - (uint64_t)p.size)) / (uint64_t)p.size) + 1ull);
}

__device__ __forceinline__ uint32_t mod_size(const LevelP& p, uint32_t x) {
    if (p.pow2mask) return x & p.pow2mask;
    if (p.size == 1) return 0;
    const uint32_t t = __umulhi(x, p.mul);
    const uint32_t q = (t + ((x - t) >> 1)) >> (p.shift - 1);
    return x - q * p.size;
}

// position within a level (gridencoder.cu:144-149): pos = fma(x, scale, 0 | 0.5), cell = min(floor, res-1)
__device__ __forceinline__ void locate1(float in, const LevelP& p, bool align_corners, uint32_t& cell, float& frac) {
    const float pos = __fmaf_rn(in, p.scale, align_corners ? 0.0f : 0.5f);
    cell = (uint32_t)fminf(floorf(pos), (float)(p.resolution - 1));
    frac = __fsub_rn(pos, (float)cell);
}

// ------------------------------------------------------------------------------------------------
// forward, fast path D=3 C=2 (the only instance the model uses)
// ------------------------------------------------------------------------------------------------
// optional input transform of the dual entry points: xform = {min[3], size[3], bound}; reproduces, operation for operation,
// BBox.normalize (common.py:288: (x - min) / size) followed by GridEncoder.forward's (x + bound) / (2 bound) (grid.py:174)
__device__ __forceinline__ float xform1(float v, const float* __restrict__ xf, int d) {
    const float t = __fdiv_rn(__fsub_rn(v, __ldg(xf + d)), __ldg(xf + 3 + d));
    const float b = __ldg(xf + 6);
    return __fmul_rn(__fadd_rn(t, b), __fdiv_rn(1.0f, __fmul_rn(2.0f, b)));     // torch divides by a scalar as a * (1 / s)
}

template <typename T> struct Vec2;
template <> struct Vec2<float> { typedef float2 type; };
template <> struct Vec2<__half> { typedef __half2 type; };

__device__ __forceinline__ void corner_rows_d3(const LevelP& p, uint32_t cx, uint32_t cy, uint32_t cz, uint32_t rows[8]) {
    if (p.use_hash) {
        const uint32_t hx0 = cx, hx1 = cx + 1;
        const uint32_t hy0 = cy * 2654435761u, hy1 = (cy + 1) * 2654435761u;
        const uint32_t hz0 = (cz * 805459861u) ^ p.hash_style, hz1 = ((cz + 1) * 805459861u) ^ p.hash_style;
#pragma unroll
        for (int k = 0; k < 8; k++)
            rows[k] = mod_size(p, ((k & 1) ? hx1 : hx0) ^ ((k & 2) ? hy1 : hy0) ^ ((k & 4) ? hz1 : hz0));
    } else {
        const uint32_t base = cx * p.stride[0] + cy * p.stride[1] + cz * p.stride[2] + p.style_term;
#pragma unroll
        for (int k = 0; k < 8; k++)
            rows[k] = mod_size(p, base + ((k & 1) ? p.stride[0] : 0u) + ((k & 2) ? p.stride[1] : 0u) + ((k & 4) ? p.stride[2] : 0u));
    }
}

__device__ __forceinline__ void corner_weights_d3(float fx, float fy, float fz, float w[8]) {
    // gridencoder.cu:157-170: w = 1; w *= (bit d ? f_d : 1-f_d) for d = 0,1,2 (in that order)
    const float gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy), gz = __fsub_rn(1.0f, fz);
#pragma unroll
    for (int k = 0; k < 8; k++)
        w[k] = __fmul_rn(__fmul_rn((k & 1) ? fx : gx, (k & 2) ? fy : gy), (k & 4) ? fz : gz);
}

__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ uint32_t h2u(float2) { return 0u; }

// NE = 2: two encoders with IDENTICAL geometry (offsets, scale, resolution) evaluated in one pass -- positions, hash rows
// and trilinear weights are computed once and reused for both tables (the model's x_density_embedder / x_color_embedder,
// networks/style_nerf.py:29-30, are such a pair).
// PAIR (NE = 2 only): the two tables are ONE interleaved buffer [row][encoder][2] (table0; 8 bytes per row in f16, 16 in
// f32), so a corner of both encoders is a single vector gather from a single sector instead of two.
template <typename T> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<__half> { typedef uint2 type; };
__device__ __forceinline__ float2 pair_part(const float4& q, int e) { return e == 0 ? make_float2(q.x, q.y) : make_float2(q.z, q.w); }
__device__ __forceinline__ __half2 pair_part(const uint2& q, int e) { uint32_t u = e == 0 ? q.x : q.y; return *reinterpret_cast<__half2*>(&u); }

template <typename T, int LPT, int NE, bool PAIR = false>
__global__ void __launch_bounds__(GRID_BLOCK)
k_grid_fwd_d3c2(const float* __restrict__ inputs, const T* __restrict__ table0, const T* __restrict__ table1,
                const int32_t* __restrict__ offsets, T* __restrict__ outputs0, T* __restrict__ outputs1, uint32_t B, uint32_t L,
                float S, uint32_t H, uint32_t gridtype, bool align_corners, uint32_t style, bool point_major,
                const float* __restrict__ xform, const int32_t* __restrict__ B_dev, const float4* __restrict__ row_deltas) {
    typedef typename Vec2<T>::type V2;
    __shared__ LevelP lp[LPT];
    if (B_dev) B = min(B, (uint32_t)*B_dev);      // device-driven inference loop: the launch is sized for the cap
    if (blockIdx.x * GRID_BLOCK >= B) return;     // (block-uniform) nothing to do: leave before the level set-up
    const uint32_t l0 = blockIdx.y * LPT;
    if (threadIdx.x < LPT && l0 + threadIdx.x < L) level_setup(lp[threadIdx.x], offsets, l0 + threadIdx.x, 3, S, H, gridtype, align_corners, style);
    __syncthreads();
    const uint32_t b = blockIdx.x * GRID_BLOCK + threadIdx.x;
    if (b >= B) return;
    float x = __ldg(inputs + 3 * (size_t)b), y = __ldg(inputs + 3 * (size_t)b + 1), z = __ldg(inputs + 3 * (size_t)b + 2);
    if (xform) { x = xform1(x, xform, 0); y = xform1(y, xform, 1); z = xform1(z, xform, 2); }
    bool oob = (x < 0 || x > 1) || (y < 0 || y > 1) || (z < 0 || z > 1);   // gridencoder.cu:107-114
    // inference loop: a sample slot with delta == 0 is padding (its ray left the volume before the slot was used);
    // composite_rays stops at it (raymarching.cu:1173), so its encoding is never read -- skip the 256 gathers
    if (row_deltas && __ldg(&row_deltas[b]).x == 0.0f) oob = true;
    V2 res[NE][LPT];
#pragma unroll
    for (int j = 0; j < LPT; j++) {
#pragma unroll
        for (int e = 0; e < NE; e++) {
            if constexpr (sizeof(T) == 4) res[e][j] = make_float2(0.0f, 0.0f); else res[e][j] = __floats2half2_rn(0.0f, 0.0f);
        }
        if (l0 + j >= L || oob) continue;
        const LevelP& p = lp[j];
        uint32_t cx, cy, cz; float fx, fy, fz;
        locate1(x, p, align_corners, cx, fx);
        locate1(y, p, align_corners, cy, fy);
        locate1(z, p, align_corners, cz, fz);
        uint32_t rows[8]; float w[8];
        corner_rows_d3(p, cx, cy, cz, rows);
        corner_weights_d3(fx, fy, fz, w);
        typename Vec4<T>::type q[PAIR ? 8 : 1];
        if constexpr (PAIR) {
            const typename Vec4<T>::type* tp = reinterpret_cast<const typename Vec4<T>::type*>(table0) + p.offset;
#pragma unroll
            for (int k = 0; k < 8; k++) q[k] = __ldg(tp + rows[k]);
        }
#pragma unroll
        for (int e = 0; e < NE; e++) {
            V2 v[8];
            if constexpr (PAIR) {
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = pair_part(q[k], e);
            } else {
                const V2* tl = reinterpret_cast<const V2*>(e == 0 ? table0 : table1) + p.offset;
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = __ldg(tl + rows[k]);
            }
            if constexpr (sizeof(T) == 4) {
                float r0 = 0.0f, r1 = 0.0f;   // gridencoder.cu:177 compiles to fma.rn
#pragma unroll
                for (int k = 0; k < 8; k++) { r0 = __fmaf_rn(w[k], v[k].x, r0); r1 = __fmaf_rn(w[k], v[k].y, r1); }
                res[e][j] = make_float2(r0, r1);
            } else {
                // scalar_t = at::Half: the product is rounded to half, then the half sum is rounded to half
                __half2 acc = __floats2half2_rn(0.0f, 0.0f);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float2 g = __half22float2(v[k]);
                    acc = __hadd2(acc, __floats2half2_rn(__fmul_rn(w[k], g.x), __fmul_rn(w[k], g.y)));
                }
                res[e][j] = acc;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < NE; e++) {
        T* outputs = e == 0 ? outputs0 : outputs1;
        if (point_major) {
            V2* o = reinterpret_cast<V2*>(outputs) + (size_t)b * L + l0;
            constexpr int PER16 = 16 / (int)sizeof(V2);
            if constexpr (sizeof(T) == 2 && LPT == 8) {
                // 8 levels of f16 pairs = one 32-byte sector: a single 256-bit store (STG.E.256, sm_100)
                if ((L % 8) == 0 && l0 + LPT <= L && ((uintptr_t)outputs & 31) == 0) {
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "r"(h2u(res[e][0])), "r"(h2u(res[e][1])),
                                 "r"(h2u(res[e][2])), "r"(h2u(res[e][3])), "r"(h2u(res[e][4])), "r"(h2u(res[e][5])), "r"(h2u(res[e][6])),
                                 "r"(h2u(res[e][7])) : "memory");
                    continue;
                }
            }
            if (LPT % PER16 == 0 && (L % PER16) == 0 && l0 + LPT <= L && ((uintptr_t)outputs & 15) == 0) {
#pragma unroll
                for (int j = 0; j < LPT; j += PER16) {
                    uint4 u;
                    if constexpr (sizeof(T) == 4) {
                        u = make_uint4(__float_as_uint(res[e][j].x), __float_as_uint(res[e][j].y), __float_as_uint(res[e][(j + 1) % LPT].x), __float_as_uint(res[e][(j + 1) % LPT].y));
                    } else {
                        u = make_uint4(h2u(res[e][j]), h2u(res[e][(j + 1) % LPT]), h2u(res[e][(j + 2) % LPT]), h2u(res[e][(j + 3) % LPT]));
                    }
                    *reinterpret_cast<uint4*>(o + j) = u;
                }
            } else {
#pragma unroll
                for (int j = 0; j < LPT; j++) if (l0 + j < L) o[j] = res[e][j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < LPT; j++) if (l0 + j < L) reinterpret_cast<V2*>(outputs)[(size_t)(l0 + j) * B + b] = res[e][j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward, generic (D in {2,3}, C in {1,2,4,8}, optional dy_dx): thread per (point, level), kernel_grid :83-235
// ------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ uint32_t grid_row(const LevelP& p, const uint32_t pg[D]) {
    uint32_t index;
    if (p.use_hash) {
        index = 0;
#pragma unroll
        for (int d = 0; d < D; d++) index ^= pg[d] * prime_of(d);
        index ^= p.hash_style;
    } else {
        index = p.style_term;
#pragma unroll
        for (int d = 0; d < D; d++) index += pg[d] * p.stride[d];
    }
    return mod_size(p, index);
}

__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __half* p) { return __half2float(__ldg(p)); }
__device__ __forceinline__ void st_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_from_float(__half* p, float v) { *p = __float2half_rn(v); }
// `acc += w*g` in the reference's scalar_t arithmetic
__device__ __forceinline__ float acc_step(float acc, float w, float g, float) { return __fmaf_rn(w, g, acc); }
__device__ __forceinline__ float acc_step(float acc, float w, float g, __half) {
    const float prod = __half2float(__float2half_rn(__fmul_rn(w, g)));
    return __half2float(__float2half_rn(__fadd_rn(acc, prod)));
}

template <typename T, int D, int C>
__global__ void __launch_bounds__(GRID_BLOCK)
k_grid_fwd_generic(const float* __restrict__ inputs, const T* __restrict__ table, const int32_t* __restrict__ offsets,
                   T* __restrict__ outputs, uint32_t B, uint32_t L, float S, uint32_t H, bool calc_grad_inputs,
                   T* __restrict__ dy_dx, uint32_t gridtype, bool align_corners, uint32_t style, bool point_major) {
    __shared__ LevelP p;
    const uint32_t level = blockIdx.y;
    if (threadIdx.x == 0) level_setup(p, offsets, level, D, S, H, gridtype, align_corners, style);
    __syncthreads();
    const uint32_t b = blockIdx.x * GRID_BLOCK + threadIdx.x;
    if (b >= B) return;
    T* out = point_major ? outputs + ((size_t)b * L + level) * C : outputs + ((size_t)level * B + b) * C;
    T* dd = dy_dx + (size_t)b * D * L * C + (size_t)level * D * C;
    float in[D];
    bool oob = false;
#pragma unroll
    for (int d = 0; d < D; d++) { in[d] = __ldg(inputs + (size_t)b * D + d); if (in[d] < 0 || in[d] > 1) oob = true; }
    if (oob) {
#pragma unroll
        for (int c = 0; c < C; c++) st_from_float(out + c, 0.0f);
        if (calc_grad_inputs) {
#pragma unroll
            for (int k = 0; k < D * C; k++) st_from_float(dd + k, 0.0f);
        }
        return;
    }
    uint32_t pg[D]; float pos[D];
#pragma unroll
    for (int d = 0; d < D; d++) locate1(in[d], p, align_corners, pg[d], pos[d]);
    const T* tl = table + (size_t)p.offset * C;
    float res[C];
#pragma unroll
    for (int c = 0; c < C; c++) res[c] = 0.0f;
#pragma unroll
    for (int idx = 0; idx < (1 << D); idx++) {
        float w = 1.0f; uint32_t pl[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            if ((idx & (1 << d)) == 0) { w = __fmul_rn(w, __fsub_rn(1.0f, pos[d])); pl[d] = pg[d]; }
            else { w = __fmul_rn(w, pos[d]); pl[d] = pg[d] + 1; }
        }
        const uint32_t row = grid_row<D>(p, pl);
#pragma unroll
        for (int c = 0; c < C; c++) res[c] = acc_step(res[c], w, ld_as_float(tl + (size_t)row * C + c), T());
    }
#pragma unroll
    for (int c = 0; c < C; c++) st_from_float(out + c, res[c]);
    if (calc_grad_inputs) {   // gridencoder.cu:191-234
#pragma unroll
        for (int gd = 0; gd < D; gd++) {
            float rg[C];
#pragma unroll
            for (int c = 0; c < C; c++) rg[c] = 0.0f;
#pragma unroll
            for (int idx = 0; idx < (1 << (D - 1)); idx++) {
                float w = p.scale; uint32_t pl[D];
#pragma unroll
                for (int nd = 0; nd < D - 1; nd++) {
                    const int d = (nd >= gd) ? (nd + 1) : nd;
                    if ((idx & (1 << nd)) == 0) { w = __fmul_rn(w, __fsub_rn(1.0f, pos[d])); pl[d] = pg[d]; }
                    else { w = __fmul_rn(w, pos[d]); pl[d] = pg[d] + 1; }
                }
                pl[gd] = pg[gd];
                const uint32_t rl = grid_row<D>(p, pl);
                pl[gd] = pg[gd] + 1;
                const uint32_t rr = grid_row<D>(p, pl);
#pragma unroll
                for (int c = 0; c < C; c++) {
                    float diff = __fsub_rn(ld_as_float(tl + (size_t)rr * C + c), ld_as_float(tl + (size_t)rl * C + c));
                    if (sizeof(T) == 2) diff = __half2float(__float2half_rn(diff));
                    rg[c] = acc_step(rg[c], w, diff, T());
                }
            }
#pragma unroll
            for (int c = 0; c < C; c++) st_from_float(dd + gd * C + c, rg[c]);
        }
    }
}

static int g_fwd_lpt = 16;   // levels per thread of the fast path (tunable: nrf_grid_set_tuning)
static int g_bwd_lpt = 16;
static int g_bwd_agg = 1;    // warp aggregation of the scatter on (1) / off (0)
static int g_bwd_tr_min = 9; // paired scatter: runs of at least this many lanes are reduced through the shared-memory transpose

NRF_EXPORT void nrf_grid_set_transpose_min(int min_run) { if (min_run > 0) g_bwd_tr_min = min_run; }

NRF_EXPORT void nrf_grid_set_tuning(int fwd_lpt, int bwd_lpt, int bwd_agg) {
    if (fwd_lpt > 0) g_fwd_lpt = fwd_lpt;
    if (bwd_lpt > 0) g_bwd_lpt = bwd_lpt;
    if (bwd_agg >= 0) g_bwd_agg = bwd_agg;
}

template <typename T>
static int launch_fwd(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs, uint32_t B,
                      uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, bool calc, void* dy_dx, uint32_t gridtype,
                      bool ac, uint32_t style, bool pm, cudaStream_t s) {
    const T* tab = (const T*)embeddings; T* out = (T*)outputs; T* dd = (T*)dy_dx;
    const uint32_t nbx = ceil_div_u32(B, GRID_BLOCK);
    if (D == 3 && C == 2 && !calc && (((uintptr_t)embeddings) & 7) == 0) {
        const int lpt = g_fwd_lpt;
#define FWD_FAST(LPT) k_grid_fwd_d3c2<T, LPT, 1><<<dim3(nbx, ceil_div_u32(L, LPT)), GRID_BLOCK, 0, s>>>(inputs, tab, nullptr, offsets, out, nullptr, B, L, S, H, gridtype, ac, style, pm, nullptr, nullptr, nullptr)
        if (lpt >= 16) FWD_FAST(16); else if (lpt >= 8) FWD_FAST(8); else if (lpt >= 4) FWD_FAST(4); else if (lpt >= 2) FWD_FAST(2); else FWD_FAST(1);
#undef FWD_FAST
        return nrf_check_launch();
    }
    const dim3 grid(nbx, L);
#define FWD_GEN(DD, CC) k_grid_fwd_generic<T, DD, CC><<<grid, GRID_BLOCK, 0, s>>>(inputs, tab, offsets, out, B, L, S, H, calc, dd, gridtype, ac, style, pm)
    if (D == 3) { switch (C) { case 1: FWD_GEN(3, 1); break; case 2: FWD_GEN(3, 2); break; case 4: FWD_GEN(3, 4); break; case 8: FWD_GEN(3, 8); break; default: return NRF_E_UNSUPPORTED; } }
    else if (D == 2) { switch (C) { case 1: FWD_GEN(2, 1); break; case 2: FWD_GEN(2, 2); break; case 4: FWD_GEN(2, 4); break; case 8: FWD_GEN(2, 8); break; default: return NRF_E_UNSUPPORTED; } }
    else return NRF_E_UNSUPPORTED;
#undef FWD_GEN
    return nrf_check_launch();
}

NRF_EXPORT int nrf_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs,
                                       uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                       int calc_grad_inputs, void* dy_dx, uint32_t gridtype, int align_corners, uint32_t style,
                                       int dtype, int point_major, void* stream) {
    if (B == 0) return NRF_OK;
    if (!inputs || !embeddings || !offsets || !outputs) return NRF_E_INVALID;
    if (calc_grad_inputs && !dy_dx) return NRF_E_INVALID;
    if (L == 0 || L > GRID_MAX_LEVELS) return NRF_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == NRF_DTYPE_F32)
        return launch_fwd<float>(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs != 0, dy_dx, gridtype, align_corners != 0, style, point_major != 0, s);
    if (dtype == NRF_DTYPE_F16)
        return launch_fwd<__half>(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs != 0, dy_dx, gridtype, align_corners != 0, style, point_major != 0, s);
    return NRF_E_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// backward, fast path D=3 C=2: warp-aggregated scatter (kernel_grid_backward :238-328)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_add2(float* base, uint32_t row, float a, float b) {
    atomicAdd(reinterpret_cast<float2*>(base) + row, make_float2(a, b));      // RED.ADD.F32x2 (sm_90+)
}
__device__ __forceinline__ void atomic_add2(__half* base, uint32_t row, float a, float b) {
    atomicAdd(reinterpret_cast<__half2*>(base) + row, __floats2half2_rn(a, b));
}

// scatter the 8 corner contributions (v0[k], v1[k]) of one cell: one 8-byte vector reduction per corner.
// (Measured and rejected: pairing the x / x+1 corners of an even cell on power-of-two hashed levels into one 16-byte
//  REDG.ADD.F32x4 -- the rows differ only in bit 0 -- is 10 % SLOWER on B200: a 16-byte reduction costs two 8-byte ones.)
template <typename TO>
__device__ __forceinline__ void scatter_cell(TO* gl, const uint32_t (&rows)[8], const float (&v0)[8], const float (&v1)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k++) atomic_add2(gl, rows[k], v0[k], v1[k]);
}

// NE = 2: the gradients of two encoders with identical geometry are scattered in one pass (cells, hash rows, weights
// and the warp-aggregation bookkeeping are computed once).
// PAIR (NE = 2, TO = float only): grad_table0 is ONE interleaved gradient buffer [row][encoder][2] and a corner of both
// encoders is a single 16-byte reduction.
template <typename T, typename TO, int LPT, int NE, bool PAIR = false>
__global__ void __launch_bounds__(GRID_BLOCK, NE == 1 ? 1 : (PAIR ? 3 : 4))
k_grid_bwd_d3c2(const T* __restrict__ grad0, const T* __restrict__ grad1, const float* __restrict__ inputs,
                const int32_t* __restrict__ offsets, TO* __restrict__ grad_table0, TO* __restrict__ grad_table1, uint32_t B, uint32_t L,
                float S, uint32_t H, uint32_t gridtype, bool align_corners, uint32_t style, bool point_major, int agg_max_groups,
                const float* __restrict__ xform, int tr_min) {
    typedef typename Vec2<T>::type V2;
    extern __shared__ __align__(16) float tbuf_all[];  // PAIR: per-warp transpose buffers (TR_STRIDE floats per lane)
    constexpr int WPL = (int)sizeof(V2) / 4;           // 32-bit words per level of one point's gradient
    constexpr int ROW = LPT * WPL + 1;                 // padded smem row (conflict-free column reads)
    __shared__ LevelP lp[LPT];
    __shared__ uint32_t sg[NE][GRID_BLOCK * ROW];      // this block's gradients, [point][level] (16.3 / 32.3 KB per encoder at LPT=16)
    const uint32_t l0 = blockIdx.y * LPT;
    if (threadIdx.x < LPT && l0 + threadIdx.x < L) level_setup(lp[threadIdx.x], offsets, l0 + threadIdx.x, 3, S, H, gridtype, align_corners, style);
    const uint32_t b0 = blockIdx.x * GRID_BLOCK;
    const uint32_t b = b0 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    // ---- stage the block's gradient rows through shared memory with fully coalesced 16-byte loads
    const uint32_t nlev = min((uint32_t)LPT, L - l0);
#pragma unroll
    for (int e = 0; e < NE; e++) {
        const T* grad = e == 0 ? grad0 : grad1;
        uint32_t* sge = sg[e];
        if (grad == nullptr) {                             // this table takes no gradient (frozen): all-zero rows, skipped below
            for (int c = threadIdx.x; c < GRID_BLOCK * ROW; c += GRID_BLOCK) sge[c] = 0u;
        } else if (point_major) {
            constexpr int WORDS = LPT * WPL;               // words per point in this level group
            const bool vec = (WORDS % 4 == 0) && ((L * WPL) % 4 == 0) && ((l0 * WPL) % 4 == 0) && (nlev == (uint32_t)LPT) &&
                             (((uintptr_t)grad & 15) == 0);
            if (vec) {
                constexpr int CPP = WORDS / 4;             // 16-byte chunks per point
                for (int c = threadIdx.x; c < GRID_BLOCK * CPP; c += GRID_BLOCK) {
                    const int pt = c / CPP, ch = c - pt * CPP;
                    if (b0 + pt < B) {
                        const uint4 v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(grad) + ((size_t)(b0 + pt) * L + l0) * WPL) + ch);
                        uint32_t* d = sge + pt * ROW + ch * 4;
                        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                    }
                }
            } else {
                for (int c = threadIdx.x; c < GRID_BLOCK * WORDS; c += GRID_BLOCK) {
                    const int pt = c / WORDS, wd = c - pt * WORDS;
                    if (b0 + pt < B && (uint32_t)(wd / WPL) < nlev)
                        sge[pt * ROW + wd] = __ldg(reinterpret_cast<const uint32_t*>(grad) + ((size_t)(b0 + pt) * L + l0) * WPL + wd);
                }
            }
        } else {
            for (uint32_t j = 0; j < nlev; j++) {
                if (b < B) {
#pragma unroll
                    for (int w = 0; w < WPL; w++)
                        sge[threadIdx.x * ROW + j * WPL + w] = __ldg(reinterpret_cast<const uint32_t*>(grad) + ((size_t)(l0 + j) * B + b) * WPL + w);
                }
            }
        }
    }
    float x = -1.0f, y = -1.0f, z = -1.0f;
    if (b < B) {
        x = __ldg(inputs + 3 * (size_t)b); y = __ldg(inputs + 3 * (size_t)b + 1); z = __ldg(inputs + 3 * (size_t)b + 2);
        if (xform) { x = xform1(x, xform, 0); y = xform1(y, xform, 1); z = xform1(z, xform, 2); }
    }
    const bool active = !((x < 0 || x > 1) || (y < 0 || y > 1) || (z < 0 || z > 1));   // oob points contribute nothing (:268-273)
    __syncthreads();
#pragma unroll 1
    for (uint32_t j = 0; j < nlev; j++) {
        const LevelP& p = lp[j];
        uint32_t cx = 0, cy = 0, cz = 0; float fx = 0, fy = 0, fz = 0;
        uint32_t rows[8]; float w[8];
        if (active) {
            locate1(x, p, align_corners, cx, fx);
            locate1(y, p, align_corners, cy, fy);
            locate1(z, p, align_corners, cz, fz);
            corner_rows_d3(p, cx, cy, cz, rows);
            corner_weights_d3(fx, fy, fz, w);
        }
        // lanes in the same cell share all 8 corner rows: reduce each run of equal cells with shuffles and issue ONE vector
        // atomic per corner per run.  The butterfly depth adapts to the longest run in the warp (coarse levels: 5
        // steps for 1-2 runs; fine levels: 0-2 steps).
        bool agg = false;
        int lo = lane, hi = lane, maxlen = 1;
        uint32_t cellmask = 1u << lane;
        if (agg_max_groups > 0) {
            const unsigned long long key = active ? (((unsigned long long)cz << 42) | ((unsigned long long)cy << 21) | (unsigned long long)cx)
                                                  : (0xFFFFFFFF00000000ull | (unsigned)lane);
            cellmask = __match_any_sync(NRF_FULL_MASK, key);
            lo = __ffs(cellmask) - 1; hi = 31 - __clz(cellmask);
            const uint32_t span = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
            agg = __all_sync(NRF_FULL_MASK, cellmask == span);
            if (agg) maxlen = (int)__reduce_max_sync(NRF_FULL_MASK, (unsigned)(hi - lo + 1));
            if (maxlen < agg_max_groups) agg = false;     // runs shorter than this are cheaper as plain per-lane reductions
        }
        if constexpr (PAIR) {
            float g[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (active) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    if constexpr (sizeof(T) == 4) {
                        g[2 * e] = __uint_as_float(sg[e][threadIdx.x * ROW + j * 2]);
                        g[2 * e + 1] = __uint_as_float(sg[e][threadIdx.x * ROW + j * 2 + 1]);
                    } else {
                        uint32_t u = sg[e][threadIdx.x * ROW + j];
                        const float2 gg = __half22float2(*reinterpret_cast<__half2*>(&u));
                        g[2 * e] = gg.x; g[2 * e + 1] = gg.y;
                    }
                }
            }
            const uint32_t nzmask = __ballot_sync(NRF_FULL_MASK, g[0] != 0.0f || g[1] != 0.0f || g[2] != 0.0f || g[3] != 0.0f);
            if (nzmask == 0u) continue;
            float4* gl = reinterpret_cast<float4*>(grad_table0) + p.offset;
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float wk = active ? w[k] : 0.0f;
                v[k] = make_float4(__fmul_rn(wk, g[0]), __fmul_rn(wk, g[1]), __fmul_rn(wk, g[2]), __fmul_rn(wk, g[3]));
            }
            if (agg && maxlen >= tr_min) {
                // Long runs (coarse levels): the shuffle butterfly costs 32 shuffles per step -- 160 for a 32-lane run -- and
                // ncu shows the LSU data pipe (shuffles + reductions) is what bounds this kernel.  Transpose through shared
                // memory instead: every lane stores its 32 products (8 conflict-free 16-byte stores), then the products are
                // summed down the lanes of each run from shared memory and each run issues its 8 corner reductions once.
                float* tb = tbuf_all + (threadIdx.x >> 5) * (32 * TR_STRIDE);
#pragma unroll
                for (int k = 0; k < 8; k++) *reinterpret_cast<float4*>(tb + lane * TR_STRIDE + 4 * k) = v[k];
                __syncwarp();
                const uint32_t endmask = __ballot_sync(NRF_FULL_MASK, lane == hi);
                const uint32_t flushmask = __ballot_sync(NRF_FULL_MASK, lane == hi && active && (nzmask & cellmask) != 0u);
                // lane t owns product t (corner t / 4, component t % 4): all 32 loads are issued back to back (conflict-free),
                // then summed run by run; a run's flush is ONE 4-byte reduction whose 32 lanes cover 8 rows x 16 bytes.
                // (Measured and rejected: 8 lanes with 16-byte loads + reductions, 25 % slower; regrouping the sums into 8 lanes
                //  for 16-byte reductions, 9 % slower; fetching the run's rows with one shuffle instead of eight, 5 % slower.)
                float q[32];
#pragma unroll
                for (int r = 0; r < 32; r++) q[r] = tb[r * TR_STRIDE + lane];
                float* glf = reinterpret_cast<float*>(gl);
                float acc = 0.0f;
#pragma unroll
                for (int r = 0; r < 32; r++) {
                    acc += q[r];
                    if ((endmask >> r) & 1u) {                       // warp-uniform
                        if ((flushmask >> r) & 1u) {
                            uint32_t rr = 0;
#pragma unroll
                            for (int k = 0; k < 8; k++) {
                                const uint32_t t = __shfl_sync(NRF_FULL_MASK, rows[k], r);
                                if ((lane >> 2) == k) rr = t;
                            }
                            atomicAdd(glf + (size_t)rr * 4 + (lane & 3), acc);
                        }
                        acc = 0.0f;
                    }
                }
                __syncwarp();
            } else if (agg) {
                for (int d = 1; d < maxlen; d <<= 1) {
                    const bool take = (lane + d <= hi);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const float o0 = __shfl_down_sync(NRF_FULL_MASK, v[k].x, d);
                        const float o1 = __shfl_down_sync(NRF_FULL_MASK, v[k].y, d);
                        const float o2 = __shfl_down_sync(NRF_FULL_MASK, v[k].z, d);
                        const float o3 = __shfl_down_sync(NRF_FULL_MASK, v[k].w, d);
                        if (take) { v[k].x += o0; v[k].y += o1; v[k].z += o2; v[k].w += o3; }
                    }
                }
                if (active && lane == lo && (nzmask & cellmask)) {
#pragma unroll
                    for (int k = 0; k < 8; k++) atomicAdd(gl + rows[k], v[k]);       // RED.ADD.F32x4
                }
            } else if (active && ((nzmask >> lane) & 1u)) {
#pragma unroll
                for (int k = 0; k < 8; k++) atomicAdd(gl + rows[k], v[k]);
            }
            continue;
        }
#pragma unroll
        for (int e = 0; e < NE; e++) {
            float g0 = 0.0f, g1 = 0.0f;
            if (active) {
                if constexpr (sizeof(T) == 4) {
                    g0 = __uint_as_float(sg[e][threadIdx.x * ROW + j * 2]);
                    g1 = __uint_as_float(sg[e][threadIdx.x * ROW + j * 2 + 1]);
                } else {
                    uint32_t u = sg[e][threadIdx.x * ROW + j];
                    const float2 gg = __half22float2(*reinterpret_cast<__half2*>(&u));
                    g0 = gg.x; g1 = gg.y;
                }
            }
            // exact zeros add nothing: samples behind an early-terminated ray arrive here with all-zero gradients
            // (composite_rays_train backward never writes them) and are skipped instead of costing 8 reductions each
            const uint32_t nzmask = __ballot_sync(NRF_FULL_MASK, g0 != 0.0f || g1 != 0.0f);
            if (nzmask == 0u) continue;
            TO* gl = (e == 0 ? grad_table0 : grad_table1) + (size_t)p.offset * 2;
            float v0[8], v1[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { v0[k] = active ? __fmul_rn(w[k], g0) : 0.0f; v1[k] = active ? __fmul_rn(w[k], g1) : 0.0f; }
            if (agg) {
                for (int d = 1; d < maxlen; d <<= 1) {
                    const bool take = (lane + d <= hi);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const float o0 = __shfl_down_sync(NRF_FULL_MASK, v0[k], d);
                        const float o1 = __shfl_down_sync(NRF_FULL_MASK, v1[k], d);
                        if (take) { v0[k] += o0; v1[k] += o1; }
                    }
                }
                if (active && lane == lo && (nzmask & cellmask)) scatter_cell<TO>(gl, rows, v0, v1);
            } else if (active && ((nzmask >> lane) & 1u)) {
                scatter_cell<TO>(gl, rows, v0, v1);
            }
        }
    }
}

// generic backward: thread per (point, level, channel pair) like the reference
template <typename T, typename TO, int D, int C>
__global__ void __launch_bounds__(GRID_BLOCK)
k_grid_bwd_generic(const T* __restrict__ grad, const float* __restrict__ inputs, const int32_t* __restrict__ offsets,
                   TO* __restrict__ grad_table, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align_corners,
                   uint32_t style, bool point_major) {
    __shared__ LevelP p;
    const uint32_t level = blockIdx.y;
    if (threadIdx.x == 0) level_setup(p, offsets, level, D, S, H, gridtype, align_corners, style);
    __syncthreads();
    const uint32_t b = blockIdx.x * GRID_BLOCK + threadIdx.x;
    if (b >= B) return;
    float in[D];
#pragma unroll
    for (int d = 0; d < D; d++) { in[d] = __ldg(inputs + (size_t)b * D + d); if (in[d] < 0 || in[d] > 1) return; }
    uint32_t pg[D]; float pos[D];
#pragma unroll
    for (int d = 0; d < D; d++) locate1(in[d], p, align_corners, pg[d], pos[d]);
    const T* gp = point_major ? grad + ((size_t)b * L + level) * C : grad + ((size_t)level * B + b) * C;
    float g[C];
#pragma unroll
    for (int c = 0; c < C; c++) g[c] = ld_as_float(gp + c);
    TO* gl = grad_table + (size_t)p.offset * C;
#pragma unroll
    for (int idx = 0; idx < (1 << D); idx++) {
        float w = 1.0f; uint32_t pl[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            if ((idx & (1 << d)) == 0) { w = __fmul_rn(w, __fsub_rn(1.0f, pos[d])); pl[d] = pg[d]; }
            else { w = __fmul_rn(w, pos[d]); pl[d] = pg[d] + 1; }
        }
        const uint32_t row = grid_row<D>(p, pl);
        if constexpr (C == 1) {
            if constexpr (sizeof(TO) == 4) atomicAdd(reinterpret_cast<float*>(gl) + row, __fmul_rn(w, g[0]));
            else atomicAdd(reinterpret_cast<__half*>(gl) + row, __float2half_rn(__fmul_rn(w, g[0])));
        } else {
#pragma unroll
            for (int c = 0; c < C; c += 2) atomic_add2(gl + (size_t)row * C + c, 0, __fmul_rn(w, g[c]), __fmul_rn(w, g[c + 1]));
        }
    }
}

// kernel_input_backward :331-357
template <typename T>
__global__ void k_grid_input_bwd(const T* __restrict__ grad, const T* __restrict__ dy_dx, T* __restrict__ grad_inputs,
                                 uint32_t B, uint32_t D, uint32_t C, uint32_t L, bool point_major) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const T* dd = dy_dx + (size_t)b * L * D * C;
    float result = 0.0f;
    for (uint32_t l = 0; l < L; l++)
        for (uint32_t c = 0; c < C; c++) {
            const T* gp = point_major ? grad + ((size_t)b * L + l) * C + c : grad + ((size_t)l * B + b) * C + c;
            if (sizeof(T) == 4) result = __fmaf_rn(ld_as_float(gp), ld_as_float(dd + l * D * C + d * C + c), result);
            else result = acc_step(result, ld_as_float(gp), ld_as_float(dd + l * D * C + d * C + c), T());
        }
    st_from_float(grad_inputs + t, result);
}

template <typename T, typename TO>
static bool launch_bwd_walk_single(const void* grad, const float* inputs, const int32_t* offsets, void* grad_table, uint32_t B, uint32_t L, float S,
                                   uint32_t H, uint32_t gridtype, bool ac, uint32_t style, cudaStream_t s);

template <typename T, typename TO>
static int launch_bwd(const void* grad, const float* inputs, const int32_t* offsets, void* grad_embeddings, uint32_t B,
                      uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, bool calc, const void* dy_dx, void* grad_inputs,
                      uint32_t gridtype, bool ac, uint32_t style, bool pm, cudaStream_t s) {
    const T* g = (const T*)grad; TO* ge = (TO*)grad_embeddings;
    const uint32_t nbx = ceil_div_u32(B, GRID_BLOCK);
    if (D == 3 && C == 2 && pm && g_bwd_lpt >= 16 && g_bwd_agg > 0 && (((uintptr_t)grad_embeddings) & 7) == 0 &&
        launch_bwd_walk_single<T, TO>(grad, inputs, offsets, grad_embeddings, B, L, S, H, gridtype, ac, style, s)) {
        // the walk form (below) took it
    } else if (D == 3 && C == 2 && (((uintptr_t)grad_embeddings) & 7) == 0) {
        const int lpt = g_bwd_lpt;
#define BWD_FAST(LPT) k_grid_bwd_d3c2<T, TO, LPT, 1><<<dim3(nbx, ceil_div_u32(L, LPT)), GRID_BLOCK, 0, s>>>(g, nullptr, inputs, offsets, ge, nullptr, B, L, S, H, gridtype, ac, style, pm, g_bwd_agg, nullptr, 1 << 30)
        if (lpt >= 16) BWD_FAST(16); else if (lpt >= 8) BWD_FAST(8); else if (lpt >= 4) BWD_FAST(4); else if (lpt >= 2) BWD_FAST(2); else BWD_FAST(1);
#undef BWD_FAST
    } else {
        const dim3 grid(nbx, L);
#define BWD_GEN(DD, CC) k_grid_bwd_generic<T, TO, DD, CC><<<grid, GRID_BLOCK, 0, s>>>(g, inputs, offsets, ge, B, L, S, H, gridtype, ac, style, pm)
        if (D == 3) { switch (C) { case 1: BWD_GEN(3, 1); break; case 2: BWD_GEN(3, 2); break; case 4: BWD_GEN(3, 4); break; case 8: BWD_GEN(3, 8); break; default: return NRF_E_UNSUPPORTED; } }
        else if (D == 2) { switch (C) { case 1: BWD_GEN(2, 1); break; case 2: BWD_GEN(2, 2); break; case 4: BWD_GEN(2, 4); break; case 8: BWD_GEN(2, 8); break; default: return NRF_E_UNSUPPORTED; } }
        else return NRF_E_UNSUPPORTED;
#undef BWD_GEN
    }
    if (calc) k_grid_input_bwd<T><<<ceil_div_u32((uint64_t)B * D, 256), 256, 0, s>>>(g, (const T*)dy_dx, (T*)grad_inputs, B, D, C, L, pm);
    return nrf_check_launch();
}

NRF_EXPORT int nrf_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings, const int32_t* offsets,
                                        void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                        uint32_t H, int calc_grad_inputs, const void* dy_dx, void* grad_inputs,
                                        uint32_t gridtype, int align_corners, uint32_t style, int dtype, int grad_table_dtype,
                                        int point_major, void* stream) {
    (void)embeddings;
    if (B == 0) return NRF_OK;
    if (!grad || !inputs || !offsets || !grad_embeddings) return NRF_E_INVALID;
    if (calc_grad_inputs && (!dy_dx || !grad_inputs)) return NRF_E_INVALID;
    if (L == 0 || L > GRID_MAX_LEVELS) return NRF_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const bool calc = calc_grad_inputs != 0, ac = align_corners != 0, pm = point_major != 0;
    if (dtype == NRF_DTYPE_F32 && grad_table_dtype == NRF_DTYPE_F32)
        return launch_bwd<float, float>(grad, inputs, offsets, grad_embeddings, B, D, C, L, S, H, calc, dy_dx, grad_inputs, gridtype, ac, style, pm, s);
    if (dtype == NRF_DTYPE_F16 && grad_table_dtype == NRF_DTYPE_F16)
        return launch_bwd<__half, __half>(grad, inputs, offsets, grad_embeddings, B, D, C, L, S, H, calc, dy_dx, grad_inputs, gridtype, ac, style, pm, s);
    if (dtype == NRF_DTYPE_F16 && grad_table_dtype == NRF_DTYPE_F32)
        return launch_bwd<__half, float>(grad, inputs, offsets, grad_embeddings, B, D, C, L, S, H, calc, dy_dx, grad_inputs, gridtype, ac, style, pm, s);
    return NRF_E_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// dual encoders: two tables with identical geometry in one pass (D=3, C=2, point-major, no input gradients)
// ------------------------------------------------------------------------------------------------
template <typename T>
static int launch_fwd_dual(const float* inputs, const void* e0, const void* e1, const int32_t* offsets, void* o0, void* o1, uint32_t B,
                           uint32_t L, float S, uint32_t H, uint32_t gridtype, bool ac, uint32_t style, const float* xform,
                           const int32_t* B_dev, const float* row_deltas, cudaStream_t s) {
    const uint32_t nbx = ceil_div_u32(B, GRID_BLOCK);
    // 8 levels per thread keeps the register footprint of the two result sets at the single-encoder kernel's (48 regs)
    k_grid_fwd_d3c2<T, 8, 2><<<dim3(nbx, ceil_div_u32(L, 8)), GRID_BLOCK, 0, s>>>(inputs, (const T*)e0, (const T*)e1, offsets, (T*)o0, (T*)o1,
                                                                                B, L, S, H, gridtype, ac, style, true, xform, B_dev, reinterpret_cast<const float4*>(row_deltas));
    return nrf_check_launch();
}

NRF_EXPORT int nrf_grid_encode_forward_dual_dev(const float* inputs, const void* embeddings0, const void* embeddings1,
                                                const int32_t* offsets, void* outputs0, void* outputs1, uint32_t B_cap, uint32_t L,
                                                float S, uint32_t H, uint32_t gridtype, int align_corners, uint32_t style, int dtype,
                                                const float* xform, const int32_t* B_dev, const float* row_deltas, void* stream);

NRF_EXPORT int nrf_grid_encode_forward_dual(const float* inputs, const void* embeddings0, const void* embeddings1,
                                            const int32_t* offsets, void* outputs0, void* outputs1, uint32_t B, uint32_t L, float S,
                                            uint32_t H, uint32_t gridtype, int align_corners, uint32_t style, int dtype,
                                            const float* xform, void* stream) {
    return nrf_grid_encode_forward_dual_dev(inputs, embeddings0, embeddings1, offsets, outputs0, outputs1, B, L, S, H, gridtype,
                                            align_corners, style, dtype, xform, nullptr, nullptr, stream);
}

// B_dev (device int32, or NULL): the number of rows actually present (<= B_cap, which sizes the launch).
// row_deltas (the [B,4] deltas of march_rays, or NULL): rows whose delta is 0 are padding slots and get zeros.
NRF_EXPORT int nrf_grid_encode_forward_dual_dev(const float* inputs, const void* embeddings0, const void* embeddings1,
                                                const int32_t* offsets, void* outputs0, void* outputs1, uint32_t B, uint32_t L,
                                                float S, uint32_t H, uint32_t gridtype, int align_corners, uint32_t style, int dtype,
                                                const float* xform, const int32_t* B_dev, const float* row_deltas, void* stream) {
    if (B == 0) return NRF_OK;
    if (!inputs || !embeddings0 || !embeddings1 || !offsets || !outputs0 || !outputs1) return NRF_E_INVALID;
    if (L == 0 || L > GRID_MAX_LEVELS) return NRF_E_UNSUPPORTED;
    if ((((uintptr_t)embeddings0) & 7) || (((uintptr_t)embeddings1) & 7)) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == NRF_DTYPE_F32) return launch_fwd_dual<float>(inputs, embeddings0, embeddings1, offsets, outputs0, outputs1, B, L, S, H, gridtype, align_corners != 0, style, xform, B_dev, row_deltas, s);
    if (dtype == NRF_DTYPE_F16) return launch_fwd_dual<__half>(inputs, embeddings0, embeddings1, offsets, outputs0, outputs1, B, L, S, H, gridtype, align_corners != 0, style, xform, B_dev, row_deltas, s);
    return NRF_E_UNSUPPORTED;
}

// two separate f32 gradient tables through the walk form: one single-table walk per live table (defined with the walk kernels below)
static bool launch_bwd_walk_dual(int dtype, const void* grad0, const void* grad1, const float* inputs, const int32_t* offsets, void* ge0, void* ge1,
                                 uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool ac, uint32_t style, const float* xform, cudaStream_t s);

NRF_EXPORT int nrf_grid_encode_backward_dual(const void* grad0, const void* grad1, const float* inputs, const int32_t* offsets,
                                             void* grad_embeddings0, void* grad_embeddings1, uint32_t B, uint32_t L, float S, uint32_t H,
                                             uint32_t gridtype, int align_corners, uint32_t style, int dtype, int grad_table_dtype,
                                             const float* xform, void* stream) {
    if (B == 0) return NRF_OK;
    // a table that takes no gradient (frozen) passes NULL for BOTH its output gradient and its table gradient
    if ((grad0 == nullptr) != (grad_embeddings0 == nullptr) || (grad1 == nullptr) != (grad_embeddings1 == nullptr)) return NRF_E_INVALID;
    if ((!grad0 && !grad1) || !inputs || !offsets) return NRF_E_INVALID;
    if (L == 0 || L > GRID_MAX_LEVELS) return NRF_E_UNSUPPORTED;
    if ((((uintptr_t)grad_embeddings0) & 7) || (((uintptr_t)grad_embeddings1) & 7)) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const bool ac = align_corners != 0;
    const uint32_t nbx = ceil_div_u32(B, GRID_BLOCK);
    if (grad_table_dtype == NRF_DTYPE_F32 && g_bwd_agg > 0 &&
        launch_bwd_walk_dual(dtype, grad0, grad1, inputs, offsets, grad_embeddings0, grad_embeddings1, B, L, S, H, gridtype, ac, style, xform, s))
        return nrf_check_launch();
    if (dtype == NRF_DTYPE_F16 && grad_table_dtype == NRF_DTYPE_F32)
        k_grid_bwd_d3c2<__half, float, 16, 2><<<dim3(nbx, ceil_div_u32(L, 16)), GRID_BLOCK, 0, s>>>(
            (const __half*)grad0, (const __half*)grad1, inputs, offsets, (float*)grad_embeddings0, (float*)grad_embeddings1, B, L, S, H,
            gridtype, ac, style, true, g_bwd_agg, xform, 1 << 30);
    else if (dtype == NRF_DTYPE_F16 && grad_table_dtype == NRF_DTYPE_F16)
        k_grid_bwd_d3c2<__half, __half, 16, 2><<<dim3(nbx, ceil_div_u32(L, 16)), GRID_BLOCK, 0, s>>>(
            (const __half*)grad0, (const __half*)grad1, inputs, offsets, (__half*)grad_embeddings0, (__half*)grad_embeddings1, B, L, S, H,
            gridtype, ac, style, true, g_bwd_agg, xform, 1 << 30);
    else if (dtype == NRF_DTYPE_F32 && grad_table_dtype == NRF_DTYPE_F32)      // 8 levels per block: the staged f32 gradients of two encoders fit 48 KB
        k_grid_bwd_d3c2<float, float, 8, 2><<<dim3(nbx, ceil_div_u32(L, 8)), GRID_BLOCK, 0, s>>>(
            (const float*)grad0, (const float*)grad1, inputs, offsets, (float*)grad_embeddings0, (float*)grad_embeddings1, B, L, S, H,
            gridtype, ac, style, true, g_bwd_agg, xform, 1 << 30);
    else return NRF_E_UNSUPPORTED;
    return nrf_check_launch();
}

// Interleaved ("paired") forms: both tables in ONE buffer [row][encoder][2] -- `table_pair` in the tables' dtype for the
// forward, `grad_pair` in f32 for the backward.  Same arithmetic as the dual forms; half the gathers / reductions.
NRF_EXPORT int nrf_grid_encode_forward_pair(const float* inputs, const void* table_pair, const int32_t* offsets, void* outputs0,
                                            void* outputs1, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                            int align_corners, uint32_t style, int dtype, const float* xform, const int32_t* B_dev,
                                            const float* row_deltas, void* stream) {
    if (B == 0) return NRF_OK;
    if (!inputs || !table_pair || !offsets || !outputs0 || !outputs1) return NRF_E_INVALID;
    if (L == 0 || L > GRID_MAX_LEVELS) return NRF_E_UNSUPPORTED;
    if (((uintptr_t)table_pair) & 15) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t nbx = ceil_div_u32(B, GRID_BLOCK);
    const dim3 grid(nbx, ceil_div_u32(L, 8));
    const float4* rd = reinterpret_cast<const float4*>(row_deltas);
    const bool ac = align_corners != 0;
    if (dtype == NRF_DTYPE_F16)
        k_grid_fwd_d3c2<__half, 8, 2, true><<<grid, GRID_BLOCK, 0, s>>>(inputs, (const __half*)table_pair, nullptr, offsets, (__half*)outputs0,
                                                                      (__half*)outputs1, B, L, S, H, gridtype, ac, style, true, xform, B_dev, rd);
    else if (dtype == NRF_DTYPE_F32)
        k_grid_fwd_d3c2<float, 8, 2, true><<<grid, GRID_BLOCK, 0, s>>>(inputs, (const float*)table_pair, nullptr, offsets, (float*)outputs0,
                                                                     (float*)outputs1, B, L, S, H, gridtype, ac, style, true, xform, B_dev, rd);
    else return NRF_E_UNSUPPORTED;
    return nrf_check_launch();
}

// ------------------------------------------------------------------------------------------------
// paired backward, "walk" form: the samples of a train step arrive ray by ray, and consecutive samples of a ray stay in
// the same cell for many steps on the coarse and middle levels (a cell of level 0 holds ~150 march steps, of level 8 ~8).
// Instead of one thread per sample and a shuffle / shared-memory reduction over the lanes of a warp, ONE LANE WALKS a chunk
// of CH consecutive samples at ONE level (a half-warp = 16 levels of one chunk) and keeps the 8 corners x 4 components of
// its current cell in registers; the cell is flushed with 8 16-byte reductions when the walk leaves it.  The aggregation
// costs nothing (32 fma per sample-level, no shuffles except the 3 that broadcast the position) and it spans the whole
// chunk instead of one warp's 32 samples.  Gradient rows are read as 64 contiguous bytes per half-warp (point-major
// [B, L] pairs), four steps ahead of their use.
// ------------------------------------------------------------------------------------------------
static int g_bwd_walk = 128;  // chunk length of the walk kernels (0: the thread-per-sample kernel above)
NRF_EXPORT void nrf_grid_set_bwd_walk(int chunk) { g_bwd_walk = chunk; }

__device__ __forceinline__ void walk_grads(const float2& a, const float2& b, float (&g)[4]) { g[0] = a.x; g[1] = a.y; g[2] = b.x; g[3] = b.y; }
__device__ __forceinline__ void walk_grads(const __half2& a, const __half2& b, float (&g)[4]) {
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    g[0] = fa.x; g[1] = fa.y; g[2] = fb.x; g[3] = fb.y;
}

#define WALK_WARPS 8       // warps per block

// CH consecutive samples per half-warp.  L must be 16 and the gradient rows 16-byte aligned (checked by the launcher).
template <typename T, int CH>
__global__ void __launch_bounds__(WALK_WARPS * 32, sizeof(T) == 2 ? 2 : 1)
k_grid_bwd_walk(const T* __restrict__ grad0, const T* __restrict__ grad1, const float* __restrict__ inputs,
                const int32_t* __restrict__ offsets, float4* __restrict__ grad_pair, uint32_t B, float S, uint32_t H,
                uint32_t gridtype, bool align_corners, uint32_t style, const float* __restrict__ xform) {
    typedef typename Vec2<T>::type V2;
    static_assert(CH % 16 == 0, "chunks are walked in groups of 16 samples");
    constexpr int WPL = (int)sizeof(V2) / 4;      // 32-bit words per level of one sample's gradient row
    constexpr int WR = 16 * WPL;                  // words per gradient row (16 levels)
    constexpr int CPR = WR / 4;                   // 16-byte pieces per row
    constexpr int STAGE = 2 * 16 * 2 * WR;        // words per stage of one warp: [encoder][step][half][level]
    extern __shared__ __align__(16) uint32_t wbuf_all[];
    __shared__ LevelP lps[16];
    if (threadIdx.x < 16) level_setup(lps[threadIdx.x], offsets, threadIdx.x, 3, S, H, gridtype, align_corners, style);
    __syncthreads();
    const int lane = threadIdx.x & 31, lvl = lane & 15, hf = lane >> 4, hb = lane & 16;
    const LevelP p = lps[lvl];
    uint32_t* wbuf = wbuf_all + (threadIdx.x >> 5) * (2 * STAGE);
    const size_t chunk = ((size_t)blockIdx.x * WALK_WARPS + (threadIdx.x >> 5)) * 2 + hf;
    const size_t b0 = chunk * CH;
    const uint32_t nvalid = b0 < B ? (uint32_t)min((size_t)CH, (size_t)B - b0) : 0u;     // samples of this half-warp's chunk

    // group g of the chunk = 16 samples = 16 contiguous gradient rows per encoder: copied global -> shared asynchronously
    // (16-byte pieces, zero-filled past the end of the chunk), one group ahead of the walk
    auto stage_grads = [&](int g, int buf) {
        uint32_t* dst = wbuf + buf * STAGE;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(e == 0 ? grad0 : grad1) + (b0 + (size_t)g * 16) * WR;
#pragma unroll
            for (int i = 0; i < CPR; i++) {
                const int c = i * 16 + lvl, s = c / CPR, part = c - s * CPR;
                const bool ok = (uint32_t)(g * 16 + s) < nvalid;
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst + ((e * 16 + s) * 2 + hf) * WR + part * 4);
                const uint32_t* ga = ok ? src + (size_t)s * WR + part * 4 : reinterpret_cast<const uint32_t*>(grad0);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(ga), "r"(ok ? 16 : 0) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // the position of sample (16 g + lvl) of the chunk, transformed once per 16 steps and broadcast with shuffles;
    // x = -1 marks a sample that contributes nothing (outside [0,1]^3, gridencoder.cu:268-273, or past the end)
    auto fetch_pos = [&](int g, float& x, float& y, float& z) {
        x = -1.0f; y = 0.0f; z = 0.0f;
        const uint32_t s = (uint32_t)g * 16u + (uint32_t)lvl;
        if (s < nvalid) {
            const float* ip = inputs + 3 * (b0 + s);
            float a = __ldg(ip), b = __ldg(ip + 1), c = __ldg(ip + 2);
            if (xform) { a = xform1(a, xform, 0); b = xform1(b, xform, 1); c = xform1(c, xform, 2); }
            if (!((a < 0 || a > 1) || (b < 0 || b > 1) || (c < 0 || c > 1))) { x = a; y = b; z = c; }
        }
    };

    float acc[8][4];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.0f; }
    uint32_t pcx = 0, pcy = 0, pcz = 0;
    bool have = false;
    // exact x % size without the power-of-two / size-1 branches of mod_size (the round-up magic is exact for every size >= 2)
    const uint32_t msh = p.shift > 0 ? p.shift - 1 : 0u;
    auto flush = [&]() {
        uint32_t idx[8];
        if (p.use_hash) {
            const uint32_t hy0 = pcy * 2654435761u, hy1 = hy0 + 2654435761u;
            const uint32_t hz0 = (pcz * 805459861u) ^ p.hash_style, hz1 = ((pcz + 1) * 805459861u) ^ p.hash_style;
#pragma unroll
            for (int k = 0; k < 8; k++) idx[k] = (pcx + (k & 1)) ^ ((k & 2) ? hy1 : hy0) ^ ((k & 4) ? hz1 : hz0);
        } else {
            const uint32_t base = pcx * p.stride[0] + pcy * p.stride[1] + pcz * p.stride[2] + p.style_term;
#pragma unroll
            for (int k = 0; k < 8; k++) idx[k] = base + ((k & 1) ? p.stride[0] : 0u) + ((k & 2) ? p.stride[1] : 0u) + ((k & 4) ? p.stride[2] : 0u);
        }
        float4* gl = grad_pair + p.offset;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t v = idx[k], t = __umulhi(v, p.mul);
            uint32_t row = v - (((t + ((v - t) >> 1)) >> msh) * p.size);
            if (p.size == 1u) row = 0u;
            atomicAdd(gl + row, make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]));   // RED.ADD.F32x4
        }
    };

    float nx, ny, nz;
    fetch_pos(0, nx, ny, nz);
    stage_grads(0, 0);
#pragma unroll 1
    for (int g = 0; g < CH / 16; g++) {
        const float px = nx, py = ny, pz = nz;
        if (g + 1 < CH / 16) {
            fetch_pos(g + 1, nx, ny, nz);
            stage_grads(g + 1, (g + 1) & 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        const uint32_t* rb = wbuf + (g & 1) * STAGE + hf * WR + lvl * WPL;
#pragma unroll 2
        for (int s = 0; s < 16; s++) {
            const V2 c0 = *reinterpret_cast<const V2*>(rb + (s * 2) * WR);
            const V2 c1 = *reinterpret_cast<const V2*>(rb + ((16 + s) * 2) * WR);
            const float x = __shfl_sync(NRF_FULL_MASK, px, hb + s);
            const float y = __shfl_sync(NRF_FULL_MASK, py, hb + s);
            const float z = __shfl_sync(NRF_FULL_MASK, pz, hb + s);
            float gq[4];
            walk_grads(c0, c1, gq);
            // exact zeros add nothing (samples behind an early-terminated ray): they neither open nor extend a cell
            const bool contrib = x >= 0.0f && (gq[0] != 0.0f || gq[1] != 0.0f || gq[2] != 0.0f || gq[3] != 0.0f);
            uint32_t cx, cy, cz; float fx, fy, fz;
            locate1(x, p, align_corners, cx, fx);
            locate1(y, p, align_corners, cy, fy);
            locate1(z, p, align_corners, cz, fz);
            if (contrib && (!have || cx != pcx || cy != pcy || cz != pcz)) {
                if (have) flush();
#pragma unroll
                for (int k = 0; k < 8; k++) { acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.0f; }
                pcx = cx; pcy = cy; pcz = cz; have = true;
            }
            if (contrib) {
                float w[8];
                corner_weights_d3(fx, fy, fz, w);
#pragma unroll
                for (int k = 0; k < 8; k++) {
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[k][c] = __fmaf_rn(w[k], gq[c], acc[k][c]);
                }
            }
        }
        __syncwarp();      // every lane is done with this stage before the copies of group g + 2 overwrite it
    }
    if (have) flush();
}

// Queue form of the walk: in the kernel above the flush block (8 hashes + 8 reductions) runs whenever ANY lane leaves
// its cell -- nearly every step, because the fine levels change cell at every sample -- with on average 4.6 of 16 lanes
// active.  Here a lane that leaves a cell only PARKS it (32 sums + cell + level = 144 bytes) in a per-warp ring in shared
// memory; whenever the ring holds four cells the warp drains them with all 32 lanes busy: lane t takes corner t % 8 of
// parked cell t / 8 -- one 16-byte shared load, one hash, one 16-byte reduction.
#define WALKQ_SLOTS 40     // >= 3 left over + 32 parked in one step
static int g_bwd_walk_queue = 1;
NRF_EXPORT void nrf_grid_set_bwd_walk_queue(int on) { g_bwd_walk_queue = on; }

// NE = 2: both tables, interleaved f32 gradient rows [row][table][2] (TO = float).  NE = 1: one table (grad0), gradient rows
// of two TO (float: one 8-byte reduction per corner; __half: the reference's __half2 atomics, gridencoder.cu:313-319).
__device__ __forceinline__ void walk_red(float* base, uint32_t row, float a, float b) { atomicAdd(reinterpret_cast<float2*>(base) + row, make_float2(a, b)); }
__device__ __forceinline__ void walk_red(__half* base, uint32_t row, float a, float b) { atomicAdd(reinterpret_cast<__half2*>(base) + row, __floats2half2_rn(a, b)); }

template <typename T, typename TO, int NE, int CH>
__global__ void __launch_bounds__(WALK_WARPS * 32, 2)
k_grid_bwd_walkq(const T* __restrict__ grad0, const T* __restrict__ grad1, const float* __restrict__ inputs,
                 const int32_t* __restrict__ offsets, TO* __restrict__ grad_table, uint32_t B, float S, uint32_t H,
                 uint32_t gridtype, bool align_corners, uint32_t style, const float* __restrict__ xform) {
    typedef typename Vec2<T>::type V2;
    static_assert(NE == 1 || sizeof(TO) == 4, "the interleaved pair buffer is f32");
    constexpr int NA = 2 * NE;                    // sums per corner
    constexpr int QW = 8 * NA + 4;                // words of a parked cell: the sums + (cx, cy, cz | level << 16, the level's offset and size)
    static_assert(CH % 16 == 0, "chunks are walked in groups of 16 samples");
    constexpr int WPL = (int)sizeof(V2) / 4;      // 32-bit words per level of one sample's gradient row
    constexpr int WR = 16 * WPL;                  // words per gradient row (16 levels)
    constexpr int CPR = WR / 4;                   // 16-byte pieces per row
    constexpr int STAGE = NE * 16 * 2 * WR;       // words per stage of one warp: [encoder][step][half][level]
    constexpr int PERWARP = 2 * STAGE + WALKQ_SLOTS * QW;
    extern __shared__ __align__(16) uint32_t wbuf_all[];
    __shared__ LevelP lps[16];
    if (threadIdx.x < 16) level_setup(lps[threadIdx.x], offsets, threadIdx.x, 3, S, H, gridtype, align_corners, style);
    __syncthreads();
    const int lane = threadIdx.x & 31, lvl = lane & 15, hf = lane >> 4, hb = lane & 16;
    // what the drain needs of this lane's level travels with every parked cell, so the drain has no dependent second load:
    // row offset (bits 0-23), log2(size) of a hashed power-of-two level (bits 24-28; 0: slow path), level (bits 29-31 + cz word)
    uint32_t lvl_word;
    {
        const LevelP& q = lps[lvl];
        const uint32_t lg = (q.use_hash && q.pow2mask && q.offset < (1u << 24)) ? (uint32_t)__popc(q.pow2mask) : 0u;
        lvl_word = (lg ? q.offset : 0u) | (lg << 24);
    }
    const float scale = lps[lvl].scale;
    const float rmax = (float)(lps[lvl].resolution - 1);
    const uint32_t hash_style = style * prime_of(3);
    uint32_t* wbuf = wbuf_all + (threadIdx.x >> 5) * PERWARP;
    float4* ring = reinterpret_cast<float4*>(wbuf + 2 * STAGE);
    const size_t chunk = ((size_t)blockIdx.x * WALK_WARPS + (threadIdx.x >> 5)) * 2 + hf;
    const size_t b0 = chunk * CH;
    const uint32_t nvalid = b0 < B ? (uint32_t)min((size_t)CH, (size_t)B - b0) : 0u;     // samples of this half-warp's chunk

    auto stage_grads = [&](int g, int buf) {
        uint32_t* dst = wbuf + buf * STAGE;
#pragma unroll
        for (int e = 0; e < NE; e++) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(e == 0 ? grad0 : grad1) + (b0 + (size_t)g * 16) * WR;
#pragma unroll
            for (int i = 0; i < CPR; i++) {
                const int c = i * 16 + lvl, s = c / CPR, part = c - s * CPR;
                const bool ok = (uint32_t)(g * 16 + s) < nvalid;
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst + ((e * 16 + s) * 2 + hf) * WR + part * 4);
                const uint32_t* ga = ok ? src + (size_t)s * WR + part * 4 : reinterpret_cast<const uint32_t*>(grad0);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(ga), "r"(ok ? 16 : 0) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto fetch_pos = [&](int g, float& x, float& y, float& z) {
        x = -1.0f; y = 0.0f; z = 0.0f;
        const uint32_t s = (uint32_t)g * 16u + (uint32_t)lvl;
        if (s < nvalid) {
            const float* ip = inputs + 3 * (b0 + s);
            float a = __ldg(ip), b = __ldg(ip + 1), c = __ldg(ip + 2);
            if (xform) { a = xform1(a, xform, 0); b = xform1(b, xform, 1); c = xform1(c, xform, 2); }
            if (!((a < 0 || a > 1) || (b < 0 || b > 1) || (c < 0 || c > 1))) { x = a; y = b; z = c; }
        }
    };

    float acc[8][NA];
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int c = 0; c < NA; c++) acc[k][c] = 0.0f;
    }
    uint32_t pcx = 0, pcy = 0, pcz = 0;
    bool have = false;
    int qhead = 0, qcount = 0;                     // warp-uniform ring state
    // drain parked cells, four per round (all of them when `all`)
    // (Measured and rejected: eight cells per round, two corners x / x + 1 per lane -- fewer shared-memory wavefronts per
    //  cell, but 1.06 ms instead of 0.945.)
    auto drain = [&](bool all) {
        while (qcount >= 4 || (all && qcount > 0)) {
            const int e = lane >> 3, k = lane & 7;
            if (e < qcount) {
                int slot = qhead + e; if (slot >= WALKQ_SLOTS) slot -= WALKQ_SLOTS;
                const float4* qs = ring + slot * (QW / 4);
                const uint4 cc = *reinterpret_cast<const uint4*>(qs + 2 * NA);   // {cx, cy, cz | level << 16, offset | log2(size) << 24}
                float v[NA];
                if constexpr (NE == 2) { const float4 t = qs[k]; v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
                else { const float2 t = reinterpret_cast<const float2*>(qs)[k]; v[0] = t.x; v[1] = t.y; }
                const uint32_t ccx = cc.x + (k & 1), ccy = cc.y + ((k >> 1) & 1), ccz = (cc.z & 0xffffu) + (k >> 2);
                const uint32_t lg = cc.w >> 24;
                uint32_t row, off = cc.w & 0xffffffu;
                if (lg) {                           // every level of the model's grids but the coarsest five
                    row = (ccx ^ (ccy * 2654435761u) ^ (ccz * 805459861u) ^ hash_style) & ((1u << lg) - 1u);
                } else {
                    const LevelP& q = lps[cc.z >> 16];
                    uint32_t idx;
                    if (q.use_hash) idx = ccx ^ (ccy * 2654435761u) ^ (ccz * 805459861u) ^ q.hash_style;
                    else idx = ccx * q.stride[0] + ccy * q.stride[1] + ccz * q.stride[2] + q.style_term;
                    row = mod_size(q, idx); off = q.offset;
                }
                if constexpr (NE == 2) atomicAdd(reinterpret_cast<float4*>(grad_table) + off + row, make_float4(v[0], v[1], v[2], v[3]));   // RED.ADD.F32x4
                else walk_red(grad_table + (size_t)off * 2, row, v[0], v[1]);
            }
            const int n = min(qcount, 4);
            qhead += n; if (qhead >= WALKQ_SLOTS) qhead -= WALKQ_SLOTS;
            qcount -= n;
            __syncwarp();
        }
    };
    auto park = [&](bool mine) {
        const uint32_t fm = __ballot_sync(NRF_FULL_MASK, mine);
        if (fm == 0u) return;
        if (mine) {
            int slot = qhead + qcount + __popc(fm & ((1u << lane) - 1u));
            if (slot >= WALKQ_SLOTS) slot -= WALKQ_SLOTS;
            float4* qs = ring + slot * (QW / 4);
            if constexpr (NE == 2) {
#pragma unroll
                for (int k = 0; k < 8; k++) qs[k] = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; k += 2) qs[k / 2] = make_float4(acc[k][0], acc[k][1], acc[k + 1][0], acc[k + 1][1]);
            }
            *reinterpret_cast<uint4*>(qs + 2 * NA) = make_uint4(pcx, pcy, pcz | ((uint32_t)lvl << 16), lvl_word);
#pragma unroll
            for (int k = 0; k < 8; k++) {
#pragma unroll
                for (int c = 0; c < NA; c++) acc[k][c] = 0.0f;
            }
        }
        qcount += __popc(fm);
        __syncwarp();
    };

    float nx, ny, nz;
    fetch_pos(0, nx, ny, nz);
    stage_grads(0, 0);
#pragma unroll 1
    for (int g = 0; g < CH / 16; g++) {
        const float px = nx, py = ny, pz = nz;
        if (g + 1 < CH / 16) {
            fetch_pos(g + 1, nx, ny, nz);
            stage_grads(g + 1, (g + 1) & 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        const uint32_t* rb = wbuf + (g & 1) * STAGE + hf * WR + lvl * WPL;
#pragma unroll 2
        for (int s = 0; s < 16; s++) {
            const V2 c0 = *reinterpret_cast<const V2*>(rb + (s * 2) * WR);
            const V2 c1 = NE == 2 ? *reinterpret_cast<const V2*>(rb + ((16 * (NE - 1) + s) * 2) * WR) : c0;
            const float x = __shfl_sync(NRF_FULL_MASK, px, hb + s);
            const float y = __shfl_sync(NRF_FULL_MASK, py, hb + s);
            const float z = __shfl_sync(NRF_FULL_MASK, pz, hb + s);
            float gq[4];
            walk_grads(c0, c1, gq);
            // exact zeros add nothing (samples behind an early-terminated ray): they neither open nor extend a cell
            const bool contrib = x >= 0.0f && (gq[0] != 0.0f || gq[1] != 0.0f || (NE == 2 && (gq[2] != 0.0f || gq[3] != 0.0f)));
            // locate1, with the level's constants in registers
            const float posx = __fmaf_rn(x, scale, align_corners ? 0.0f : 0.5f), posy = __fmaf_rn(y, scale, align_corners ? 0.0f : 0.5f),
                        posz = __fmaf_rn(z, scale, align_corners ? 0.0f : 0.5f);
            const uint32_t cx = (uint32_t)fminf(floorf(posx), rmax), cy = (uint32_t)fminf(floorf(posy), rmax), cz = (uint32_t)fminf(floorf(posz), rmax);
            const bool change = contrib && (!have || cx != pcx || cy != pcy || cz != pcz);
            // (Measured and rejected: letting the finest levels, which change cell at nearly every sample, reduce straight from
            //  their registers instead of parking -- 8 reductions issued with 2-8 active lanes cost more than the trip through
            //  shared memory: 1.05 ms instead of 0.945 for levels >= 13.)
            // (also measured: draining before parking, so a step never waits for its own shared-memory stores -- 0.934 ms, no gain)
            park(change && have);
            drain(false);
            if (change) { pcx = cx; pcy = cy; pcz = cz; have = true; }
            if (contrib) {
                float w[8];
                corner_weights_d3(__fsub_rn(posx, (float)cx), __fsub_rn(posy, (float)cy), __fsub_rn(posz, (float)cz), w);
#pragma unroll
                for (int k = 0; k < 8; k++) {
#pragma unroll
                    for (int c = 0; c < NA; c++) acc[k][c] = __fmaf_rn(w[k], gq[c], acc[k][c]);
                }
            }
        }
        __syncwarp();      // every lane is done with this stage before the copies of group g + 2 overwrite it
    }
    park(have);
    drain(true);
}

template <typename T, typename TO, int NE>
static void launch_walkq(uint32_t CH, const void* grad0, const void* grad1, const float* inputs, const int32_t* offsets, void* grad_table,
                         uint32_t B, float S, uint32_t H, uint32_t gridtype, bool ac, uint32_t style, const float* xform, cudaStream_t s) {
    const uint32_t blocks = ceil_div_u32(ceil_div_u32(ceil_div_u32(B, CH), 2), WALK_WARPS);
    constexpr int SMEMQ = WALK_WARPS * 4 * (2 * (NE * 16 * 2 * 16 * ((int)sizeof(typename Vec2<T>::type) / 4)) + WALKQ_SLOTS * (8 * 2 * NE + 4));
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_grid_bwd_walkq<T, TO, NE, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEMQ);
        cudaFuncSetAttribute(k_grid_bwd_walkq<T, TO, NE, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEMQ);
        cudaFuncSetAttribute(k_grid_bwd_walkq<T, TO, NE, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEMQ);
        cudaFuncSetAttribute(k_grid_bwd_walkq<T, TO, NE, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEMQ);
        attr_set = true;
    }
#define WALKQ(N) k_grid_bwd_walkq<T, TO, NE, N><<<blocks, WALK_WARPS * 32, SMEMQ, s>>>((const T*)grad0, (const T*)grad1, inputs, offsets, (TO*)grad_table, B, S, H, gridtype, ac, style, xform)
    if (CH == 256) WALKQ(256); else if (CH == 128) WALKQ(128); else if (CH == 64) WALKQ(64); else WALKQ(32);
#undef WALKQ
}

static uint32_t walk_chunk(int ch) { return ch >= 256 ? 256u : (ch >= 128 ? 128u : (ch >= 64 ? 64u : (ch >= 32 ? 32u : 16u))); }

// the walk kernels need: 16 levels (a half-warp per chunk), 16-byte aligned point-major gradient rows, cell coordinates < 2^16
static bool walk_applicable(uint32_t L, float S, uint32_t H, const void* g0, const void* g1) {
    return g_bwd_walk > 0 && L == 16 && ((((uintptr_t)g0) | ((uintptr_t)g1)) & 15) == 0 && floorf(exp2f(15.0f * S) * (float)H) < 65535.0f;
}

template <typename T>
static void launch_bwd_walk(int ch, const void* grad0, const void* grad1, const float* inputs, const int32_t* offsets, float* grad_pair,
                            uint32_t B, float S, uint32_t H, uint32_t gridtype, bool ac, uint32_t style, const float* xform,
                            cudaStream_t s) {
    const uint32_t CH = walk_chunk(ch);
    if constexpr (sizeof(T) == 2) {      // (f32 gradient rows -- parity mode: the ring would leave one block per SM)
        if (g_bwd_walk_queue && CH >= 32) {
            launch_walkq<T, float, 2>(CH, grad0, grad1, inputs, offsets, grad_pair, B, S, H, gridtype, ac, style, xform, s);
            return;
        }
    }
    const uint32_t blocks = ceil_div_u32(ceil_div_u32(ceil_div_u32(B, CH), 2), WALK_WARPS);
    constexpr int STAGES = WALK_WARPS * 2 * (2 * 16 * 2 * 16 * (int)sizeof(typename Vec2<T>::type));   // two stages per warp
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_grid_bwd_walk<T, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES);
        cudaFuncSetAttribute(k_grid_bwd_walk<T, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES);
        cudaFuncSetAttribute(k_grid_bwd_walk<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES);
        attr_set = true;
    }
#define WALK_ARGS (const T*)grad0, (const T*)grad1, inputs, offsets, reinterpret_cast<float4*>(grad_pair), B, S, H, gridtype, ac, style, xform
    if (CH >= 64) k_grid_bwd_walk<T, 64><<<ceil_div_u32(ceil_div_u32(ceil_div_u32(B, 64), 2), WALK_WARPS), WALK_WARPS * 32, STAGES, s>>>(WALK_ARGS);
    else if (CH == 32) k_grid_bwd_walk<T, 32><<<blocks, WALK_WARPS * 32, STAGES, s>>>(WALK_ARGS);
    else k_grid_bwd_walk<T, 16><<<blocks, WALK_WARPS * 32, STAGES, s>>>(WALK_ARGS);
#undef WALK_ARGS
}

// single table (the reference-facing GridEncoder, gridencoder/grid.py:71-97): same walk, one 8-byte (f32) / 4-byte (f16) reduction per corner
template <typename T, typename TO>
static bool launch_bwd_walk_single(const void* grad, const float* inputs, const int32_t* offsets, void* grad_table, uint32_t B, uint32_t L, float S,
                                   uint32_t H, uint32_t gridtype, bool ac, uint32_t style, cudaStream_t s) {
    // f16 gradient tables keep the thread-per-sample kernel: that mode exists to reproduce the reference's per-sample __half2 atomics
    if constexpr (sizeof(TO) == 4) {
        if (!g_bwd_walk_queue || !walk_applicable(L, S, H, grad, grad) || walk_chunk(g_bwd_walk) < 32) return false;
        launch_walkq<T, TO, 1>(walk_chunk(g_bwd_walk), grad, nullptr, inputs, offsets, grad_table, B, S, H, gridtype, ac, style, nullptr, s);
        return true;
    } else {
        return false;
    }
}

static bool launch_bwd_walk_dual(int dtype, const void* grad0, const void* grad1, const float* inputs, const int32_t* offsets, void* ge0, void* ge1,
                                 uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool ac, uint32_t style, const float* xform, cudaStream_t s) {
    const void* any = grad0 ? grad0 : grad1;
    if (!g_bwd_walk_queue || walk_chunk(g_bwd_walk) < 32 || !walk_applicable(L, S, H, grad0 ? grad0 : any, grad1 ? grad1 : any)) return false;
    if (dtype != NRF_DTYPE_F16 && dtype != NRF_DTYPE_F32) return false;
    const uint32_t CH = walk_chunk(g_bwd_walk);
    for (int e = 0; e < 2; e++) {
        const void* g = e == 0 ? grad0 : grad1;
        void* ge = e == 0 ? ge0 : ge1;
        if (!g) continue;                  // frozen table
        if (dtype == NRF_DTYPE_F16) launch_walkq<__half, float, 1>(CH, g, nullptr, inputs, offsets, ge, B, S, H, gridtype, ac, style, xform, s);
        else launch_walkq<float, float, 1>(CH, g, nullptr, inputs, offsets, ge, B, S, H, gridtype, ac, style, xform, s);
    }
    return true;
}

NRF_EXPORT int nrf_grid_encode_backward_pair(const void* grad0, const void* grad1, const float* inputs, const int32_t* offsets,
                                             float* grad_pair, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                             int align_corners, uint32_t style, int dtype, const float* xform, void* stream) {
    if (B == 0) return NRF_OK;
    if (!grad0 || !grad1 || !inputs || !offsets || !grad_pair) return NRF_E_INVALID;
    if (L == 0 || L > GRID_MAX_LEVELS) return NRF_E_UNSUPPORTED;
    if (((uintptr_t)grad_pair) & 15) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const bool ac = align_corners != 0;
    const uint32_t nbx = ceil_div_u32(B, GRID_BLOCK);
    static bool attr_set = false;
    if (!attr_set) {          // static staging (34 KB) + dynamic transpose buffers (36 KB) exceed the 48 KB default
        cudaFuncSetAttribute(k_grid_bwd_d3c2<__half, float, 16, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_BYTES);
        cudaFuncSetAttribute(k_grid_bwd_d3c2<float, float, 8, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TR_BYTES);
        attr_set = true;
    }
    const int tr_min = g_bwd_agg > 0 ? g_bwd_tr_min : (1 << 30);
    const bool walk_ok = walk_applicable(L, S, H, grad0, grad1);
    if (walk_ok && (dtype == NRF_DTYPE_F16 || dtype == NRF_DTYPE_F32)) {
        if (dtype == NRF_DTYPE_F16) launch_bwd_walk<__half>(g_bwd_walk, grad0, grad1, inputs, offsets, grad_pair, B, S, H, gridtype, ac, style, xform, s);
        else launch_bwd_walk<float>(g_bwd_walk, grad0, grad1, inputs, offsets, grad_pair, B, S, H, gridtype, ac, style, xform, s);
        return nrf_check_launch();
    }
    if (dtype == NRF_DTYPE_F16)
        k_grid_bwd_d3c2<__half, float, 16, 2, true><<<dim3(nbx, ceil_div_u32(L, 16)), GRID_BLOCK, TR_BYTES, s>>>(
            (const __half*)grad0, (const __half*)grad1, inputs, offsets, grad_pair, nullptr, B, L, S, H, gridtype, ac, style, true, g_bwd_agg, xform, tr_min);
    else if (dtype == NRF_DTYPE_F32)
        k_grid_bwd_d3c2<float, float, 8, 2, true><<<dim3(nbx, ceil_div_u32(L, 8)), GRID_BLOCK, TR_BYTES, s>>>(
            (const float*)grad0, (const float*)grad1, inputs, offsets, grad_pair, nullptr, B, L, S, H, gridtype, ac, style, true, g_bwd_agg, xform, tr_min);
    else return NRF_E_UNSUPPORTED;
    return nrf_check_launch();
}

// ------------------------------------------------------------------------------------------------
// grid_initialize (kernel_grid_initialize :497-531 and its host loop :534-548), D=3 C=2 f32
// ------------------------------------------------------------------------------------------------
__global__ void k_grid_initialize(const float* __restrict__ ref_grid, float* __restrict__ grid, const int32_t* __restrict__ ref_offsets,
                                  const int32_t* __restrict__ offsets, uint32_t level, float S, uint32_t H, uint32_t Ns) {
    __shared__ LevelP pr;
    const uint32_t tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    if (tid == 0) level_setup(pr, ref_offsets, level, 3, S, H, 0, true, 0);
    __syncthreads();
    uint32_t pg[3] = { blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y * blockDim.y + threadIdx.y, blockIdx.z * blockDim.z + threadIdx.z };
    const uint32_t resolution = pr.resolution;
    if (pg[0] > resolution || pg[1] > resolution || pg[2] > resolution) return;
    const float2 v = reinterpret_cast<const float2*>(ref_grid)[pr.offset + grid_row<3>(pr, pg)];
    for (uint32_t s = 0; s < Ns; s++) {
        LevelP p;
        level_setup(p, offsets, level, 3, S, H, 0, true, s);
        reinterpret_cast<float2*>(grid)[p.offset + grid_row<3>(p, pg)] = v;
    }
}

NRF_EXPORT int nrf_grid_initialize(const float* ref_embeddings, float* embeddings, const int32_t* ref_offsets,
                                   const int32_t* offsets, uint32_t L, float S, uint32_t H, uint32_t Ns, void* stream) {
    if (!ref_embeddings || !embeddings || !ref_offsets || !offsets) return NRF_E_INVALID;
    if (L == 0 || L > GRID_MAX_LEVELS) return NRF_E_UNSUPPORTED;
    for (uint32_t level = 0; level < L; level++) {
        const uint32_t resolution = (uint32_t)floorf(exp2f((float)level * S) * (float)H);   // host mirror of :539
        const uint32_t nb = ceil_div_u32(resolution + 1, 8);
        k_grid_initialize<<<dim3(nb, nb, nb), dim3(8, 8, 8), 0, (cudaStream_t)stream>>>(ref_embeddings, embeddings, ref_offsets, offsets, level, S, H, Ns);
    }
    return nrf_check_launch();
}
