// tcnn-style fully fused MLP (bias-free, 64 wide) for sm_100a -- forward and backward.
//
// Contract: y = act_out(W_n relu(... relu(W_1 x))), weights f16 in tcnn's
// FullyFusedMLP layout ([out,in] row-major per layer, in/out padded to 16), hidden activations rounded
// to f16 between layers, f32 accumulation on the tensor cores.  Call sites replaced:
// /root/reference/networks/style_nerf.py:44-98 (tcnn.Network x4), networks/tcnn_nerf.py:97-122.
//
// Design: each warp owns a 16-row tile and chains the layers entirely in registers (the D fragment of
// one mma is repacked as the A fragment of the next), weights live in shared memory for the life of
// the (persistent) block.  The backward kernel RECOMPUTES the hidden activations from x instead of
// reading them back from HBM (nothing but x and y is saved by the forward), back-propagates through
// the layers in registers, stores dx, and forms the weight gradients on the tensor cores from
// shared-memory tiles (ldmatrix.trans) with per-warp register accumulators that are flushed with
// one float atomicAdd per weight per block at the end.
#include "common.cuh"

#define MLP_WIDTH 64
#define MLP_WARPS 4
#define MLP_THREADS (MLP_WARPS * 32)
#define HLD 72          // padded row stride (halfs) of 64-wide tiles in shared memory: conflict-free for 32-bit
                        // fragment accesses (stride 36 words) and for ldmatrix (row shift 16 B)

__device__ __forceinline__ void mma16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
    return __half22float2(*reinterpret_cast<__half2*>(&v));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t r[4], const __half* smem_ptr) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ float act_fwd(float z, int act) {
    switch (act) {
        case NRF_ACT_RELU: return fmaxf(z, 0.0f);
        case NRF_ACT_SIGMOID: return 1.0f / (1.0f + __expf(-z));
        case NRF_ACT_EXP: return __expf(z);
        default: return z;
    }
}
// d act(z) / dz
__device__ __forceinline__ float act_bwd(float z, int act) {
    switch (act) {
        case NRF_ACT_RELU: return z > 0.0f ? 1.0f : 0.0f;
        case NRF_ACT_SIGMOID: { const float y = 1.0f / (1.0f + __expf(-z)); return y * (1.0f - y); }
        case NRF_ACT_EXP: return __expf(z);
        default: return 1.0f;
    }
}

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }

// two consecutive elements (row r, cols col, col+1) of a [B, n] row-major matrix as packed f16x2; zero outside
template <typename XT>
__device__ __forceinline__ uint32_t load_pair(const XT* __restrict__ x, uint32_t r, uint32_t col, uint32_t B, uint32_t n, bool vec_ok) {
    if (r >= B || col >= n) return 0u;
    const XT* p = x + (size_t)r * n + col;
    if (vec_ok) {
        if constexpr (sizeof(XT) == 2) return __ldg(reinterpret_cast<const uint32_t*>(p));
        else { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); return pack_h2(v.x, v.y); }
    }
    const float lo = to_f(p[0]);
    const float hi = (col + 1 < n) ? to_f(p[1]) : 0.0f;
    return pack_h2(lo, hi);
}
template <typename XT>
__device__ __forceinline__ float2 load_pair_f(const XT* __restrict__ x, uint32_t r, uint32_t col, uint32_t B, uint32_t n) {
    float2 v = make_float2(0.0f, 0.0f);
    if (r >= B || col >= n) return v;
    const XT* p = x + (size_t)r * n + col;
    v.x = to_f(p[0]);
    if (col + 1 < n) v.y = to_f(p[1]);
    return v;
}
template <typename YT>
__device__ __forceinline__ void store_pair(YT* __restrict__ y, uint32_t r, uint32_t col, uint32_t B, uint32_t n, float v0, float v1, bool vec_ok) {
    if (r >= B || col >= n) return;
    YT* p = y + (size_t)r * n + col;
    if (vec_ok) {
        if constexpr (sizeof(YT) == 2) *reinterpret_cast<uint32_t*>(p) = pack_h2(v0, v1);
        else *reinterpret_cast<float2*>(p) = make_float2(v0, v1);
        return;
    }
    if constexpr (sizeof(YT) == 2) { p[0] = __float2half_rn(v0); if (col + 1 < n) p[1] = __float2half_rn(v1); }
    else { p[0] = v0; if (col + 1 < n) p[1] = v1; }
}

// copy a [rows, cols] f16 row-major matrix from global into shared with row stride ld (cols <= ld)
__device__ __forceinline__ void stage_matrix(__half* dst, int ld, const __half* __restrict__ src, int rows, int cols) {
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) dst[(i / cols) * ld + (i % cols)] = src[i];
}
// transposed: dst[c][r] = src[r][c]
__device__ __forceinline__ void stage_matrix_T(__half* dst, int ld, const __half* __restrict__ src, int rows, int cols) {
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) dst[(i % cols) * ld + (i / cols)] = src[i];
}

// acc[NT][4] += A[KT k-tiles] * W^T, with W [n_rows >= 8*NT, ld] row-major f16 in shared (B fragment k-contiguous)
template <int KT, int NT>
__device__ __forceinline__ void warp_gemm(float (&acc)[NT][4], const uint32_t (&a)[KT][4], const __half* W, int ld, int g, int t) {
#pragma unroll
    for (int j = 0; j < NT; j++) {
        const __half* wr = W + (8 * j + g) * ld + 2 * t;
#pragma unroll
        for (int kk = 0; kk < KT; kk++) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr + 16 * kk);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wr + 16 * kk + 8);
            mma16816(acc[j], a[kk], b0, b1);
        }
    }
}

// relu (or identity) + round to f16 + repack 8 D n-tiles as 4 A k-tiles
template <int NT>
__device__ __forceinline__ void repack_act(uint32_t (&a)[NT / 2][4], const float (&acc)[NT][4], bool relu) {
#pragma unroll
    for (int kk = 0; kk < NT / 2; kk++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float* c = acc[2 * kk + h];
            float v0 = c[0], v1 = c[1], v2 = c[2], v3 = c[3];
            if (relu) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); v2 = fmaxf(v2, 0.0f); v3 = fmaxf(v3, 0.0f); }
            a[kk][2 * h + 0] = pack_h2(v0, v1);
            a[kk][2 * h + 1] = pack_h2(v2, v3);
        }
    }
}

// write A fragments (KT k-tiles) of a 16-row tile to shared, row-major with stride ld
template <int KT>
__device__ __forceinline__ void store_tile(__half* tile, int ld, const uint32_t (&a)[KT][4], int g, int t) {
#pragma unroll
    for (int kk = 0; kk < KT; kk++) {
        *reinterpret_cast<uint32_t*>(tile + g * ld + 16 * kk + 2 * t) = a[kk][0];
        *reinterpret_cast<uint32_t*>(tile + (g + 8) * ld + 16 * kk + 2 * t) = a[kk][1];
        *reinterpret_cast<uint32_t*>(tile + g * ld + 16 * kk + 8 + 2 * t) = a[kk][2];
        *reinterpret_cast<uint32_t*>(tile + (g + 8) * ld + 16 * kk + 8 + 2 * t) = a[kk][3];
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int IN_KT, int NH, int OUT_NT, typename XT, typename YT>
__global__ void __launch_bounds__(MLP_THREADS)
k_mlp_fwd(const XT* __restrict__ x, const __half* __restrict__ params, uint32_t B, uint32_t n_in, uint32_t n_out,
          int hidden_act, int out_act, YT* __restrict__ y) {
    constexpr int IN_PAD = IN_KT * 16, IN_LD = IN_PAD + 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sW1 = reinterpret_cast<__half*>(smem_raw);             // [64][IN_LD]
    __half* sWh = sW1 + MLP_WIDTH * IN_LD;                        // [NH-1][64][HLD]
    __half* sWo = sWh + (NH - 1) * MLP_WIDTH * HLD;               // [16][HLD]
    stage_matrix(sW1, IN_LD, params, MLP_WIDTH, IN_PAD);
    for (int l = 0; l < NH - 1; l++)
        stage_matrix(sWh + l * MLP_WIDTH * HLD, HLD, params + MLP_WIDTH * IN_PAD + l * MLP_WIDTH * MLP_WIDTH, MLP_WIDTH, MLP_WIDTH);
    stage_matrix(sWo, HLD, params + MLP_WIDTH * IN_PAD + (NH - 1) * MLP_WIDTH * MLP_WIDTH, 16, MLP_WIDTH);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const uint32_t ntiles = (B + 15) / 16;
    const bool x_vec = (n_in % 2 == 0) && (((uintptr_t)x) % (2 * sizeof(XT)) == 0);
    const bool y_vec = (n_out % 2 == 0) && (((uintptr_t)y) % (2 * sizeof(YT)) == 0);
    const bool relu = hidden_act == NRF_ACT_RELU;
    for (uint32_t tile = blockIdx.x * MLP_WARPS + warp; tile < ntiles; tile += gridDim.x * MLP_WARPS) {
        const uint32_t r0 = tile * 16 + g, r1 = r0 + 8;
        uint32_t xa[IN_KT][4];
#pragma unroll
        for (int kk = 0; kk < IN_KT; kk++) {
            xa[kk][0] = load_pair(x, r0, 16 * kk + 2 * t, B, n_in, x_vec);
            xa[kk][1] = load_pair(x, r1, 16 * kk + 2 * t, B, n_in, x_vec);
            xa[kk][2] = load_pair(x, r0, 16 * kk + 8 + 2 * t, B, n_in, x_vec);
            xa[kk][3] = load_pair(x, r1, 16 * kk + 8 + 2 * t, B, n_in, x_vec);
        }
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f; }
        warp_gemm<IN_KT, 8>(acc, xa, sW1, IN_LD, g, t);
        uint32_t ha[4][4];
        repack_act<8>(ha, acc, relu);
#pragma unroll
        for (int l = 0; l < NH - 1; l++) {
#pragma unroll
            for (int j = 0; j < 8; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f; }
            warp_gemm<4, 8>(acc, ha, sWh + l * MLP_WIDTH * HLD, HLD, g, t);
            repack_act<8>(ha, acc, relu);
        }
        float zo[OUT_NT][4];
#pragma unroll
        for (int j = 0; j < OUT_NT; j++) { zo[j][0] = zo[j][1] = zo[j][2] = zo[j][3] = 0.0f; }
        warp_gemm<4, OUT_NT>(zo, ha, sWo, HLD, g, t);
#pragma unroll
        for (int j = 0; j < OUT_NT; j++) {
            store_pair(y, r0, 8 * j + 2 * t, B, n_out, act_fwd(zo[j][0], out_act), act_fwd(zo[j][1], out_act), y_vec);
            store_pair(y, r1, 8 * j + 2 * t, B, n_out, act_fwd(zo[j][2], out_act), act_fwd(zo[j][3], out_act), y_vec);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward (recompute + dx + dW)
// ------------------------------------------------------------------------------------------------
struct BwdSmem {
    // offsets in halfs
    int W1, W1T, Wh, WhT, Wo, WoT, X, H1, H2, dH1, dH2, dZ, total;
};
template <int IN_KT, int NH>
__host__ __device__ constexpr BwdSmem bwd_layout() {
    constexpr int IN_PAD = IN_KT * 16, IN_LD = IN_PAD + 8, ZLD = 24, ROWS = MLP_WARPS * 16;
    BwdSmem s{};
    int o = 0;
    s.W1 = o;  o += MLP_WIDTH * IN_LD;               // [64][IN_LD]
    s.W1T = o; o += IN_PAD * HLD;                    // [IN_PAD][HLD]
    s.Wh = o;  o += (NH - 1) * MLP_WIDTH * HLD;      // [64][HLD]
    s.WhT = o; o += (NH - 1) * MLP_WIDTH * HLD;
    s.Wo = o;  o += 16 * HLD;                        // [16][HLD]
    s.WoT = o; o += MLP_WIDTH * ZLD;                 // [64][ZLD]
    s.X = o;   o += ROWS * IN_LD;
    s.H1 = o;  o += ROWS * HLD;
    s.H2 = o;  o += (NH - 1) * ROWS * HLD;
    s.dH1 = o; o += ROWS * HLD;
    s.dH2 = o; o += (NH - 1) * ROWS * HLD;
    s.dZ = o;  o += ROWS * ZLD;
    s.total = o;
    return s;
}

// acc[NT][4] += dOut^T[m-tile mt, all ROWS] * In[ROWS, n-tiles nt0 .. nt0+NT-1]; tiles are row-major in shared.
template <int NT>
__device__ __forceinline__ void wgrad_gemm(float (&acc)[NT][4], const __half* sOut, int ld_out, int mt, const __half* sIn, int ld_in,
                                           int nt0, int lane) {
    static_assert(NT % 2 == 0, "n-tiles are fetched in pairs");
    const int i = lane >> 3, r = lane & 7;
#pragma unroll
    for (int ks = 0; ks < MLP_WARPS; ks++) {
        uint32_t a[4];
        ldmatrix_x4_trans(a, sOut + (ks * 16 + r + ((i & 2) ? 8 : 0)) * ld_out + mt * 16 + ((i & 1) ? 8 : 0));
#pragma unroll
        for (int j = 0; j < NT; j += 2) {
            uint32_t b[4];
            ldmatrix_x4_trans(b, sIn + (ks * 16 + r + ((i & 1) ? 8 : 0)) * ld_in + 8 * (nt0 + j) + ((i & 2) ? 8 : 0));
            mma16816(acc[j], a, b[0], b[1]);
            mma16816(acc[j + 1], a, b[2], b[3]);
        }
    }
}

// flush a per-warp accumulator slice into the global f32 gradient (row-major [*, ld])
template <int NT>
__device__ __forceinline__ void flush_acc(float* __restrict__ dW, int ld, int m0, int n0, const float (&acc)[NT][4], float scale,
                                          int m_limit, int g, int t) {
#pragma unroll
    for (int j = 0; j < NT; j++) {
        const int n = n0 + 8 * j + 2 * t;
        if (m0 + g < m_limit) {
            atomicAdd(dW + (m0 + g) * ld + n, acc[j][0] * scale);
            atomicAdd(dW + (m0 + g) * ld + n + 1, acc[j][1] * scale);
        }
        if (m0 + g + 8 < m_limit) {
            atomicAdd(dW + (m0 + g + 8) * ld + n, acc[j][2] * scale);
            atomicAdd(dW + (m0 + g + 8) * ld + n + 1, acc[j][3] * scale);
        }
    }
}

template <int IN_KT, int NH, int OUT_NT, typename XT, typename DYT, typename DXT>
__global__ void __launch_bounds__(MLP_THREADS)
k_mlp_bwd(const XT* __restrict__ x, const __half* __restrict__ params, const DYT* __restrict__ dy, uint32_t B, uint32_t n_in,
          uint32_t n_out, int hidden_act, int out_act, float loss_scale, DXT* __restrict__ dx, float* __restrict__ dparams) {
    constexpr int IN_PAD = IN_KT * 16, IN_LD = IN_PAD + 8, ZLD = 24, ROWS = MLP_WARPS * 16;
    constexpr BwdSmem L = bwd_layout<IN_KT, NH>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sm = reinterpret_cast<__half*>(smem_raw);
    const __half* pW1 = params;
    const __half* pWh = params + MLP_WIDTH * IN_PAD;
    const __half* pWo = pWh + (NH - 1) * MLP_WIDTH * MLP_WIDTH;
    stage_matrix(sm + L.W1, IN_LD, pW1, MLP_WIDTH, IN_PAD);
    stage_matrix_T(sm + L.W1T, HLD, pW1, MLP_WIDTH, IN_PAD);
    if constexpr (NH == 2) { stage_matrix(sm + L.Wh, HLD, pWh, MLP_WIDTH, MLP_WIDTH); stage_matrix_T(sm + L.WhT, HLD, pWh, MLP_WIDTH, MLP_WIDTH); }
    stage_matrix(sm + L.Wo, HLD, pWo, 16, MLP_WIDTH);
    stage_matrix_T(sm + L.WoT, ZLD, pWo, 16, MLP_WIDTH);
    // zero the tile area once (padding columns are never read by ldmatrix, but keep it clean)
    for (int i = threadIdx.x; i < L.total - L.X; i += blockDim.x) sm[L.X + i] = __float2half(0.0f);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const uint32_t nchunks = (B + ROWS - 1) / ROWS;
    const bool x_vec = (n_in % 2 == 0) && (((uintptr_t)x) % (2 * sizeof(XT)) == 0);
    const bool dx_vec = dx && (n_in % 2 == 0) && (((uintptr_t)dx) % (2 * sizeof(DXT)) == 0);
    const bool relu = hidden_act == NRF_ACT_RELU;
    const float inv_scale = 1.0f / loss_scale;

    // per-warp weight-gradient accumulators (f32): warp w owns out-features [16w, 16w+16) of W1 / W2 and
    // in-features [16w, 16w+16) of W_out
    float gW1[IN_KT * 2][4], gWh[NH == 2 ? 8 : 2][4], gWo[2][4];
#pragma unroll
    for (int j = 0; j < IN_KT * 2; j++) gW1[j][0] = gW1[j][1] = gW1[j][2] = gW1[j][3] = 0.0f;
#pragma unroll
    for (int j = 0; j < (NH == 2 ? 8 : 2); j++) gWh[j][0] = gWh[j][1] = gWh[j][2] = gWh[j][3] = 0.0f;
#pragma unroll
    for (int j = 0; j < 2; j++) gWo[j][0] = gWo[j][1] = gWo[j][2] = gWo[j][3] = 0.0f;

    __half* tX = sm + L.X + warp * 16 * IN_LD;
    __half* tH1 = sm + L.H1 + warp * 16 * HLD;
    __half* tH2 = sm + L.H2 + warp * 16 * HLD;
    __half* tdH1 = sm + L.dH1 + warp * 16 * HLD;
    __half* tdH2 = sm + L.dH2 + warp * 16 * HLD;
    __half* tdZ = sm + L.dZ + warp * 16 * ZLD;

    for (uint32_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const uint32_t r0 = chunk * ROWS + warp * 16 + g, r1 = r0 + 8;
        // ---------------- phase A: per-warp recompute + backward through the layers
        uint32_t xa[IN_KT][4];
#pragma unroll
        for (int kk = 0; kk < IN_KT; kk++) {
            xa[kk][0] = load_pair(x, r0, 16 * kk + 2 * t, B, n_in, x_vec);
            xa[kk][1] = load_pair(x, r1, 16 * kk + 2 * t, B, n_in, x_vec);
            xa[kk][2] = load_pair(x, r0, 16 * kk + 8 + 2 * t, B, n_in, x_vec);
            xa[kk][3] = load_pair(x, r1, 16 * kk + 8 + 2 * t, B, n_in, x_vec);
        }
        store_tile<IN_KT>(tX, IN_LD, xa, g, t);
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f; }
        warp_gemm<IN_KT, 8>(acc, xa, sm + L.W1, IN_LD, g, t);
        uint32_t h1a[4][4], h2a[4][4];
        repack_act<8>(h1a, acc, relu);
        store_tile<4>(tH1, HLD, h1a, g, t);
        if constexpr (NH == 2) {
#pragma unroll
            for (int j = 0; j < 8; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f; }
            warp_gemm<4, 8>(acc, h1a, sm + L.Wh, HLD, g, t);
            repack_act<8>(h2a, acc, relu);
            store_tile<4>(tH2, HLD, h2a, g, t);
        }
        const uint32_t (&hlast)[4][4] = (NH == 2) ? h2a : h1a;
        float zo[OUT_NT][4];
#pragma unroll
        for (int j = 0; j < OUT_NT; j++) { zo[j][0] = zo[j][1] = zo[j][2] = zo[j][3] = 0.0f; }
        warp_gemm<4, OUT_NT>(zo, hlast, sm + L.Wo, HLD, g, t);
        // dZ = loss_scale * dy * act'(z) as the A fragment of a single 16-wide k-tile
        uint32_t dza[1][4] = {{0u, 0u, 0u, 0u}};
#pragma unroll
        for (int j = 0; j < OUT_NT; j++) {
            const float2 d0 = load_pair_f(dy, r0, 8 * j + 2 * t, B, n_out);
            const float2 d1 = load_pair_f(dy, r1, 8 * j + 2 * t, B, n_out);
            dza[0][2 * j + 0] = pack_h2(d0.x * loss_scale * act_bwd(zo[j][0], out_act), d0.y * loss_scale * act_bwd(zo[j][1], out_act));
            dza[0][2 * j + 1] = pack_h2(d1.x * loss_scale * act_bwd(zo[j][2], out_act), d1.y * loss_scale * act_bwd(zo[j][3], out_act));
        }
        store_tile<1>(tdZ, ZLD, dza, g, t);
        // dH_last = dZ * W_out, masked by relu'
#pragma unroll
        for (int j = 0; j < 8; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f; }
        warp_gemm<1, 8>(acc, dza, sm + L.WoT, ZLD, g, t);
        uint32_t dha[4][4];
        auto mask_pack = [&](uint32_t (&out)[4][4], const uint32_t (&h)[4][4]) {
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    const float* c = acc[2 * kk + hh];
                    float v0 = c[0], v1 = c[1], v2 = c[2], v3 = c[3];
                    if (relu) {
                        const float2 m0 = unpack_h2(h[kk][2 * hh + 0]), m1 = unpack_h2(h[kk][2 * hh + 1]);
                        if (!(m0.x > 0.0f)) v0 = 0.0f;
                        if (!(m0.y > 0.0f)) v1 = 0.0f;
                        if (!(m1.x > 0.0f)) v2 = 0.0f;
                        if (!(m1.y > 0.0f)) v3 = 0.0f;
                    }
                    out[kk][2 * hh + 0] = pack_h2(v0, v1);
                    out[kk][2 * hh + 1] = pack_h2(v2, v3);
                }
            }
        };
        mask_pack(dha, hlast);
        if constexpr (NH == 2) {
            store_tile<4>(tdH2, HLD, dha, g, t);
#pragma unroll
            for (int j = 0; j < 8; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f; }
            warp_gemm<4, 8>(acc, dha, sm + L.WhT, HLD, g, t);
            mask_pack(dha, h1a);
        }
        store_tile<4>(tdH1, HLD, dha, g, t);
        if (dx) {
            float ax[IN_KT * 2][4];
#pragma unroll
            for (int j = 0; j < IN_KT * 2; j++) { ax[j][0] = ax[j][1] = ax[j][2] = ax[j][3] = 0.0f; }
            warp_gemm<4, IN_KT * 2>(ax, dha, sm + L.W1T, HLD, g, t);
#pragma unroll
            for (int j = 0; j < IN_KT * 2; j++) {
                store_pair(dx, r0, 8 * j + 2 * t, B, n_in, ax[j][0] * inv_scale, ax[j][1] * inv_scale, dx_vec);
                store_pair(dx, r1, 8 * j + 2 * t, B, n_in, ax[j][2] * inv_scale, ax[j][3] * inv_scale, dx_vec);
            }
        }
        __syncthreads();
        // ---------------- phase B: weight gradients over the block's ROWS rows
        if (dparams) {
            wgrad_gemm<IN_KT * 2>(gW1, sm + L.dH1, HLD, warp, sm + L.X, IN_LD, 0, lane);
            if constexpr (NH == 2) {
                float (&gw)[8][4] = reinterpret_cast<float (&)[8][4]>(gWh);
                wgrad_gemm<8>(gw, sm + L.dH2, HLD, warp, sm + L.H1, HLD, 0, lane);
            }
            wgrad_gemm<2>(gWo, sm + L.dZ, ZLD, 0, (NH == 2) ? sm + L.H2 : sm + L.H1, HLD, 2 * warp, lane);
        }
        __syncthreads();
    }
    if (dparams) {
        float* dW1 = dparams;
        float* dWh = dparams + MLP_WIDTH * IN_PAD;
        float* dWo = dWh + (NH - 1) * MLP_WIDTH * MLP_WIDTH;
        flush_acc<IN_KT * 2>(dW1, IN_PAD, 16 * warp, 0, gW1, inv_scale, MLP_WIDTH, g, t);
        if constexpr (NH == 2) {
            float (&gw)[8][4] = reinterpret_cast<float (&)[8][4]>(gWh);
            flush_acc<8>(dWh, MLP_WIDTH, 16 * warp, 0, gw, inv_scale, MLP_WIDTH, g, t);
        }
        flush_acc<2>(dWo, MLP_WIDTH, 0, 16 * warp, gWo, inv_scale, 16, g, t);
    }
}

// ------------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------------
// mlp_tc.cu: the tcgen05 / TMEM implementation (default); the mma.sync kernels above stay as the selectable
// second implementation (nrf_mlp_set_mode(1)) that the parity tests run side by side.
int nrf_mlp_tc_forward(const void* x, int x_dtype, const void* params_f16, uint32_t B, uint32_t n_in, uint32_t n_out, uint32_t n_hidden,
                       int hidden_act, int out_act, void* y, int y_dtype, uint32_t ld_y, const int32_t* B_dev, cudaStream_t s);
int nrf_mlp_tc_backward(const void* x, int x_dtype, const void* params_f16, const void* dy, int dy_dtype, uint32_t ld_dy, uint32_t B,
                        uint32_t n_in, uint32_t n_out, uint32_t n_hidden, int hidden_act, int out_act, float loss_scale, void* dx,
                        int dx_accumulate, float* dparams, cudaStream_t s);
void nrf_mlp_tc_set_ctas(int fwd_per_sm, int bwd_per_sm);
void nrf_mlp_tc_set_prof(unsigned long long* buf16);
static int g_mlp_mode = 0;      // 0 = tcgen05 (mlp_tc.cu), 1 = mma.sync (this file)
NRF_EXPORT void nrf_mlp_set_mode(int mode) { g_mlp_mode = mode; }
NRF_EXPORT int nrf_mlp_get_mode(void) { return g_mlp_mode; }
NRF_EXPORT void nrf_mlp_set_profile(void* device_buf_16_u64) { nrf_mlp_tc_set_prof((unsigned long long*)device_buf_16_u64); }
NRF_EXPORT void nrf_mlp_set_tuning(int fwd_ctas_per_sm, int bwd_ctas_per_sm) { nrf_mlp_tc_set_ctas(fwd_ctas_per_sm, bwd_ctas_per_sm); }

static int g_sm_count = 0;
static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

template <int IN_KT, int NH, int OUT_NT, typename XT, typename YT>
static int launch_fwd_t(const void* x, const void* params, uint32_t B, uint32_t n_in, uint32_t n_out, int hact, int oact, void* y,
                        cudaStream_t s) {
    constexpr int IN_LD = IN_KT * 16 + 8;
    const size_t smem = sizeof(__half) * (MLP_WIDTH * IN_LD + (NH - 1) * MLP_WIDTH * HLD + 16 * HLD);
    auto kern = k_mlp_fwd<IN_KT, NH, OUT_NT, XT, YT>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint32_t ntiles = (B + 15) / 16;
    const uint32_t grid = (uint32_t)min((uint64_t)ceil_div_u32(ntiles, MLP_WARPS), (uint64_t)sm_count() * 8);
    kern<<<grid, MLP_THREADS, smem, s>>>((const XT*)x, (const __half*)params, B, n_in, n_out, hact, oact, (YT*)y);
    return nrf_check_launch();
}

template <int IN_KT, int NH, int OUT_NT>
static int launch_fwd_dt(const void* x, int xdt, const void* params, uint32_t B, uint32_t n_in, uint32_t n_out, int hact, int oact,
                         void* y, int ydt, cudaStream_t s) {
    if (xdt == NRF_DTYPE_F16 && ydt == NRF_DTYPE_F16) return launch_fwd_t<IN_KT, NH, OUT_NT, __half, __half>(x, params, B, n_in, n_out, hact, oact, y, s);
    if (xdt == NRF_DTYPE_F32 && ydt == NRF_DTYPE_F16) return launch_fwd_t<IN_KT, NH, OUT_NT, float, __half>(x, params, B, n_in, n_out, hact, oact, y, s);
    if (xdt == NRF_DTYPE_F16 && ydt == NRF_DTYPE_F32) return launch_fwd_t<IN_KT, NH, OUT_NT, __half, float>(x, params, B, n_in, n_out, hact, oact, y, s);
    if (xdt == NRF_DTYPE_F32 && ydt == NRF_DTYPE_F32) return launch_fwd_t<IN_KT, NH, OUT_NT, float, float>(x, params, B, n_in, n_out, hact, oact, y, s);
    return NRF_E_UNSUPPORTED;
}

NRF_EXPORT int nrf_mlp_forward_ex(const void* x, int x_dtype, const void* params_f16, uint32_t B, uint32_t n_in, uint32_t n_out,
                                  uint32_t n_hidden, uint32_t width, int hidden_act, int out_act, void* y, int y_dtype,
                                  uint32_t ld_y, void* stream);
NRF_EXPORT int nrf_mlp_forward(const void* x, int x_dtype, const void* params_f16, uint32_t B, uint32_t n_in, uint32_t n_out,
                               uint32_t n_hidden, uint32_t width, int hidden_act, int out_act, void* y, int y_dtype,
                               void* stream) {
    return nrf_mlp_forward_ex(x, x_dtype, params_f16, B, n_in, n_out, n_hidden, width, hidden_act, out_act, y, y_dtype, n_out, stream);
}
NRF_EXPORT int nrf_mlp_forward_dev(const void* x, int x_dtype, const void* params_f16, uint32_t B_cap, uint32_t n_in, uint32_t n_out,
                                   uint32_t n_hidden, uint32_t width, int hidden_act, int out_act, void* y, int y_dtype,
                                   uint32_t ld_y, const int32_t* B_dev, void* stream);
NRF_EXPORT int nrf_mlp_forward_ex(const void* x, int x_dtype, const void* params_f16, uint32_t B, uint32_t n_in, uint32_t n_out,
                                  uint32_t n_hidden, uint32_t width, int hidden_act, int out_act, void* y, int y_dtype,
                                  uint32_t ld_y, void* stream) {
    return nrf_mlp_forward_dev(x, x_dtype, params_f16, B, n_in, n_out, n_hidden, width, hidden_act, out_act, y, y_dtype, ld_y, nullptr,
                               stream);
}
// B_dev (device int32, or NULL): rows actually present (<= B_cap, which sizes the launch); tcgen05 implementation only
NRF_EXPORT int nrf_mlp_forward_dev(const void* x, int x_dtype, const void* params_f16, uint32_t B, uint32_t n_in, uint32_t n_out,
                                   uint32_t n_hidden, uint32_t width, int hidden_act, int out_act, void* y, int y_dtype,
                                   uint32_t ld_y, const int32_t* B_dev, void* stream) {
    if (B == 0) return NRF_OK;
    if (ld_y == 0) ld_y = n_out;
    if (ld_y < n_out) return NRF_E_INVALID;
    if (!x || !params_f16 || !y) return NRF_E_INVALID;
    if (width != MLP_WIDTH || n_in == 0 || n_in > 64 || n_out == 0 || n_out > 16 || n_hidden < 1 || n_hidden > 2) return NRF_E_UNSUPPORTED;
    if (hidden_act != NRF_ACT_RELU && hidden_act != NRF_ACT_NONE) return NRF_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    if (x_dtype != NRF_DTYPE_F16 && x_dtype != NRF_DTYPE_F32) return NRF_E_UNSUPPORTED;
    if (y_dtype != NRF_DTYPE_F16 && y_dtype != NRF_DTYPE_F32) return NRF_E_UNSUPPORTED;
    if (g_mlp_mode == 0) return nrf_mlp_tc_forward(x, x_dtype, params_f16, B, n_in, n_out, n_hidden, hidden_act, out_act, y, y_dtype, ld_y, B_dev, s);
    if (ld_y != n_out || out_act == NRF_ACT_TRUNC_EXP || B_dev) return NRF_E_UNSUPPORTED;      // extensions exist on the tcgen05 path only
    const int kt = (int)((n_in + 15) / 16), nt = n_out <= 8 ? 1 : 2;
#define FWD_CASE(K, H, N) if (kt == K && (int)n_hidden == H && nt == N) return launch_fwd_dt<K, H, N>(x, x_dtype, params_f16, B, n_in, n_out, hidden_act, out_act, y, y_dtype, s)
    FWD_CASE(1, 1, 1); FWD_CASE(1, 1, 2); FWD_CASE(1, 2, 1); FWD_CASE(1, 2, 2);
    FWD_CASE(2, 1, 1); FWD_CASE(2, 1, 2); FWD_CASE(2, 2, 1); FWD_CASE(2, 2, 2);
    FWD_CASE(3, 1, 1); FWD_CASE(3, 1, 2); FWD_CASE(3, 2, 1); FWD_CASE(3, 2, 2);
    FWD_CASE(4, 1, 1); FWD_CASE(4, 1, 2); FWD_CASE(4, 2, 1); FWD_CASE(4, 2, 2);
#undef FWD_CASE
    return NRF_E_UNSUPPORTED;
}

template <int IN_KT, int NH, int OUT_NT, typename XT, typename DYT, typename DXT>
static int launch_bwd_t(const void* x, const void* params, const void* dy, uint32_t B, uint32_t n_in, uint32_t n_out, int hact, int oact,
                        float loss_scale, void* dx, float* dparams, cudaStream_t s) {
    constexpr BwdSmem L = bwd_layout<IN_KT, NH>();
    const size_t smem = sizeof(__half) * (size_t)L.total;
    auto kern = k_mlp_bwd<IN_KT, NH, OUT_NT, XT, DYT, DXT>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint32_t nchunks = ceil_div_u32(B, MLP_WARPS * 16);
    const uint32_t per_sm = (uint32_t)max(1, (int)(200 * 1024 / smem));
    const uint32_t grid = (uint32_t)min((uint64_t)nchunks, (uint64_t)sm_count() * min(per_sm, 4u));
    kern<<<grid, MLP_THREADS, smem, s>>>((const XT*)x, (const __half*)params, (const DYT*)dy, B, n_in, n_out, hact, oact, loss_scale,
                                        (DXT*)dx, dparams);
    return nrf_check_launch();
}

template <int IN_KT, int NH, int OUT_NT>
static int launch_bwd_dt(const void* x, int xdt, const void* params, const void* dy, int dydt, uint32_t B, uint32_t n_in, uint32_t n_out,
                         int hact, int oact, float ls, void* dx, int dxdt, float* dparams, cudaStream_t s) {
    // dx dtype follows x dtype (the autograd contract); dy may be f16 or f32
    if (dx && dxdt != xdt) return NRF_E_UNSUPPORTED;
    if (xdt == NRF_DTYPE_F16 && dydt == NRF_DTYPE_F16) return launch_bwd_t<IN_KT, NH, OUT_NT, __half, __half, __half>(x, params, dy, B, n_in, n_out, hact, oact, ls, dx, dparams, s);
    if (xdt == NRF_DTYPE_F32 && dydt == NRF_DTYPE_F16) return launch_bwd_t<IN_KT, NH, OUT_NT, float, __half, float>(x, params, dy, B, n_in, n_out, hact, oact, ls, dx, dparams, s);
    if (xdt == NRF_DTYPE_F16 && dydt == NRF_DTYPE_F32) return launch_bwd_t<IN_KT, NH, OUT_NT, __half, float, __half>(x, params, dy, B, n_in, n_out, hact, oact, ls, dx, dparams, s);
    if (xdt == NRF_DTYPE_F32 && dydt == NRF_DTYPE_F32) return launch_bwd_t<IN_KT, NH, OUT_NT, float, float, float>(x, params, dy, B, n_in, n_out, hact, oact, ls, dx, dparams, s);
    return NRF_E_UNSUPPORTED;
}

NRF_EXPORT int nrf_mlp_backward_ex(const void* x, int x_dtype, const void* params_f16, const void* dy, int dy_dtype, uint32_t ld_dy,
                                   uint32_t B, uint32_t n_in, uint32_t n_out, uint32_t n_hidden, uint32_t width, int hidden_act,
                                   int out_act, float loss_scale, void* dx, int dx_dtype, int dx_accumulate, float* dparams,
                                   void* stream);
NRF_EXPORT int nrf_mlp_backward(const void* x, int x_dtype, const void* params_f16, const void* dy, int dy_dtype, uint32_t B,
                                uint32_t n_in, uint32_t n_out, uint32_t n_hidden, uint32_t width, int hidden_act, int out_act,
                                float loss_scale, void* dx, int dx_dtype, float* dparams, void* stream) {
    return nrf_mlp_backward_ex(x, x_dtype, params_f16, dy, dy_dtype, n_out, B, n_in, n_out, n_hidden, width, hidden_act, out_act,
                               loss_scale, dx, dx_dtype, 0, dparams, stream);
}
NRF_EXPORT int nrf_mlp_backward_ex(const void* x, int x_dtype, const void* params_f16, const void* dy, int dy_dtype, uint32_t ld_dy,
                                   uint32_t B, uint32_t n_in, uint32_t n_out, uint32_t n_hidden, uint32_t width, int hidden_act,
                                   int out_act, float loss_scale, void* dx, int dx_dtype, int dx_accumulate, float* dparams,
                                   void* stream) {
    if (B == 0) return NRF_OK;
    if (ld_dy == 0) ld_dy = n_out;
    if (ld_dy < n_out) return NRF_E_INVALID;
    if (!x || !params_f16 || !dy) return NRF_E_INVALID;
    if (width != MLP_WIDTH || n_in == 0 || n_in > 64 || n_out == 0 || n_out > 16 || n_hidden < 1 || n_hidden > 2) return NRF_E_UNSUPPORTED;
    if (hidden_act != NRF_ACT_RELU && hidden_act != NRF_ACT_NONE) return NRF_E_UNSUPPORTED;
    if (!(loss_scale > 0.0f)) return NRF_E_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (g_mlp_mode == 0) {
        if (dx && dx_dtype != x_dtype) return NRF_E_UNSUPPORTED;
        if ((x_dtype != NRF_DTYPE_F16 && x_dtype != NRF_DTYPE_F32) || (dy_dtype != NRF_DTYPE_F16 && dy_dtype != NRF_DTYPE_F32)) return NRF_E_UNSUPPORTED;
        return nrf_mlp_tc_backward(x, x_dtype, params_f16, dy, dy_dtype, ld_dy, B, n_in, n_out, n_hidden, hidden_act, out_act, loss_scale, dx,
                                   dx_accumulate, dparams, s);
    }
    if (ld_dy != n_out || dx_accumulate || out_act == NRF_ACT_TRUNC_EXP) return NRF_E_UNSUPPORTED;     // tcgen05 path only
    const int kt = (int)((n_in + 15) / 16), nt = n_out <= 8 ? 1 : 2;
#define BWD_CASE(K, H, N) if (kt == K && (int)n_hidden == H && nt == N) return launch_bwd_dt<K, H, N>(x, x_dtype, params_f16, dy, dy_dtype, B, n_in, n_out, hidden_act, out_act, loss_scale, dx, dx_dtype, dparams, s)
    BWD_CASE(1, 1, 1); BWD_CASE(1, 1, 2); BWD_CASE(1, 2, 1); BWD_CASE(1, 2, 2);
    BWD_CASE(2, 1, 1); BWD_CASE(2, 1, 2); BWD_CASE(2, 2, 1); BWD_CASE(2, 2, 2);
    BWD_CASE(3, 1, 1); BWD_CASE(3, 1, 2); BWD_CASE(3, 2, 1); BWD_CASE(3, 2, 2);
    BWD_CASE(4, 1, 1); BWD_CASE(4, 1, 2); BWD_CASE(4, 2, 1); BWD_CASE(4, 2, 2);
#undef BWD_CASE
    return NRF_E_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// tcnn.Encoding 'SphericalHarmonics' (networks/style_nerf.py:33-42, networks/tcnn_nerf.py:87-95; use_dir=True models)
// ------------------------------------------------------------------------------------------------
// tiny-cuda-nn is un-vendored; this restates its published real-SH basis (degree <= 4; the reference's dir_enc_sh_deg is 4):
// inputs in [0,1]^3 are mapped to [-1,1]^3 (x*2-1) and evaluated as polynomials.  Forward only: the encoding has no
// parameters and ray directions never carry gradients on this path.
template <typename OT>
__global__ void k_sh_encode(const float* __restrict__ in, uint32_t B, uint32_t degree, OT* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float x = __ldg(in + 3 * (size_t)i) * 2.0f - 1.0f, y = __ldg(in + 3 * (size_t)i + 1) * 2.0f - 1.0f, z = __ldg(in + 3 * (size_t)i + 2) * 2.0f - 1.0f;
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    float o[16];
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
    const uint32_t n = degree * degree;
    OT* dst = out + (size_t)i * n;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if ((uint32_t)k < n) {
            if constexpr (sizeof(OT) == 2) dst[k] = __float2half_rn(o[k]); else dst[k] = o[k];
        }
    }
}

NRF_EXPORT int nrf_sh_encode_forward(const float* inputs01, uint32_t B, uint32_t degree, void* outputs, int out_dtype, void* stream) {
    if (B == 0) return NRF_OK;
    if (!inputs01 || !outputs) return NRF_E_INVALID;
    if (degree < 1 || degree > 4) return NRF_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    if (out_dtype == NRF_DTYPE_F16) k_sh_encode<__half><<<ceil_div_u32(B, 256), 256, 0, s>>>(inputs01, B, degree, (__half*)outputs);
    else if (out_dtype == NRF_DTYPE_F32) k_sh_encode<float><<<ceil_div_u32(B, 256), 256, 0, s>>>(inputs01, B, degree, (float*)outputs);
    else return NRF_E_UNSUPPORTED;
    return nrf_check_launch();
}
