// fp32 "parity mode" of the tcnn-style MLP ("fp32 (parity mode) or with fp16 rounding of weights and layer
// inputs (perf mode)"; SURVEY 7 hard part 3).  The same network  y = act_out(W_n relu(... relu(W_1 x)))  evaluated with
// fp32 weights, fp32 activations and fp32 FMA chains on the SIMT pipes -- no rounding point anywhere -- so that gradients
// of the whole render/train path can be compared with an fp32 CPU evaluation at rel 1e-4.  It is not a performance path
// (the tcgen05 kernels of mlp_tc.cu are); it exists to separate "the kernel is wrong" from "fp16 rounded it".
//
// A block owns 64-row tiles.  Weights, the tile's activations and its gradients live in shared memory and every layer is a
// small strided product  out[m][n] = sum_k A[m][k] * B[n][k]  with runtime sizes (one kernel per input dtype: it compiles in
// seconds).  Weight gradients are accumulated per block in shared memory (each entry has one owner thread: no atomics
// inside the block) and added to global memory once at the end.
#include "common.cuh"

namespace {

constexpr int W = 64;            // hidden width
constexpr int TR = 64;           // rows per tile
constexpr int NT = 128;          // threads per block
constexpr int OUTP = 16;         // padded output width
constexpr int LD = W + 1;        // row stride of the shared activation tiles

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return (float)*p; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }

__device__ __forceinline__ float act_fwd(float z, int act) {
    switch (act) {
        case NRF_ACT_RELU: return fmaxf(z, 0.0f);
        case NRF_ACT_SIGMOID: return 1.0f / (1.0f + expf(-z));
        case NRF_ACT_EXP: case NRF_ACT_TRUNC_EXP: return expf(z);
        default: return z;
    }
}
// d act / d z given z and y = act(z)
__device__ __forceinline__ float act_bwd(float z, float y, int act) {
    switch (act) {
        case NRF_ACT_RELU: return z > 0.0f ? 1.0f : 0.0f;
        case NRF_ACT_SIGMOID: return y * (1.0f - y);
        case NRF_ACT_EXP: return y;
        case NRF_ACT_TRUNC_EXP: return expf(fminf(fmaxf(z, -15.0f), 15.0f));       // networks/tcnn_nerf.py:65-69
        default: return 1.0f;
    }
}

// out[m * ldo + n] (=|+=) sum_{k < K} A[m * sam + k * sak] * B[n * sbn + k * sbk]     m < M, n < N; one owner thread per entry
template <bool ACCUM>
__device__ __forceinline__ void prod(float* out, int ldo, const float* A, int sam, int sak, const float* Bm, int sbn, int sbk, int M, int N,
                                     int K) {
    for (int idx = threadIdx.x; idx < M * N; idx += NT) {
        const int m = idx / N, n = idx - m * N;
        const float* a = A + m * sam;
        const float* b = Bm + n * sbn;
        float s = 0.0f;
        for (int k = 0; k < K; k++) s = fmaf(a[k * sak], b[k * sbk], s);
        if (ACCUM) out[m * ldo + n] += s; else out[m * ldo + n] = s;
    }
}

struct Shape { int in_pad, nh, n_par, off2, offo; };
__device__ __forceinline__ Shape shape_of(uint32_t n_in, uint32_t n_hidden) {
    Shape s;
    s.in_pad = (int)((n_in + 15) / 16) * 16;
    s.nh = (int)n_hidden;
    s.off2 = W * s.in_pad;
    s.offo = s.off2 + (s.nh - 1) * W * W;
    s.n_par = s.offo + OUTP * W;
    return s;
}

template <typename XT, typename YT>
__global__ void __launch_bounds__(NT) k_mlp_f32_fwd(const XT* __restrict__ x, const float* __restrict__ params, uint32_t B, uint32_t n_in,
                                                    uint32_t n_out, uint32_t n_hidden, int hact, int oact, YT* __restrict__ y, uint32_t ld_y) {
    extern __shared__ float sm[];
    const Shape S = shape_of(n_in, n_hidden);
    float* Wp = sm;
    float* Xa = sm + S.n_par;            // [TR][LD]
    float* Ha = Xa + TR * LD;
    for (int i = threadIdx.x; i < S.n_par; i += NT) Wp[i] = params[i];
    const uint32_t n_tiles = (B + TR - 1) / TR;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t r0 = tile * TR;
        const int rows = (int)min((uint32_t)TR, B - r0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < TR * S.in_pad; idx += NT) {
            const int m = idx / S.in_pad, i = idx - m * S.in_pad;
            Xa[m * LD + i] = (m < rows && (uint32_t)i < n_in) ? ldf(x + (size_t)(r0 + m) * n_in + i) : 0.0f;
        }
        __syncthreads();
        prod<false>(Ha, LD, Xa, LD, 1, Wp, S.in_pad, 1, TR, W, S.in_pad);
        __syncthreads();
        for (int idx = threadIdx.x; idx < TR * W; idx += NT) { float* p = Ha + (idx / W) * LD + idx % W; *p = act_fwd(*p, hact); }
        __syncthreads();
        float* last = Ha;
        if (S.nh == 2) {
            prod<false>(Xa, LD, Ha, LD, 1, Wp + S.off2, W, 1, TR, W, W);
            __syncthreads();
            for (int idx = threadIdx.x; idx < TR * W; idx += NT) { float* p = Xa + (idx / W) * LD + idx % W; *p = act_fwd(*p, hact); }
            __syncthreads();
            last = Xa;
        }
        for (int idx = threadIdx.x; idx < rows * (int)n_out; idx += NT) {
            const int m = idx / (int)n_out, o = idx - m * (int)n_out;
            float z = 0.0f;
            for (int j = 0; j < W; j++) z = fmaf(Wp[S.offo + o * W + j], last[m * LD + j], z);
            stf(y + (size_t)(r0 + m) * ld_y + o, act_fwd(z, oact));
        }
    }
}

template <typename XT, typename DYT>
__global__ void __launch_bounds__(NT) k_mlp_f32_bwd(const XT* __restrict__ x, const float* __restrict__ params, const DYT* __restrict__ dy,
                                                    uint32_t ld_dy, uint32_t B, uint32_t n_in, uint32_t n_out, uint32_t n_hidden, int hact,
                                                    int oact, XT* __restrict__ dx, int dx_accumulate, float* __restrict__ dparams) {
    extern __shared__ float sm[];
    const Shape S = shape_of(n_in, n_hidden);
    float* Wp = sm;
    float* dWp = Wp + S.n_par;           // per-block weight-gradient accumulator
    float* Xa = dWp + S.n_par;           // x                                   [TR][LD]
    float* H1 = Xa + TR * LD;            // pre-activation of hidden layer 1
    float* H2 = H1 + TR * LD;            // pre-activation of hidden layer 2 (nh == 2), else scratch
    float* Ga = H2 + TR * LD;            // gradient w.r.t. a layer's pre-activation
    float* Gb = Ga + TR * LD;            // second gradient buffer / activated copy
    for (int i = threadIdx.x; i < S.n_par; i += NT) { Wp[i] = params[i]; dWp[i] = 0.0f; }
    const int hrelu = hact == NRF_ACT_RELU ? NRF_ACT_RELU : NRF_ACT_NONE;
    const uint32_t n_tiles = (B + TR - 1) / TR;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t r0 = tile * TR;
        const int rows = (int)min((uint32_t)TR, B - r0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < TR * S.in_pad; idx += NT) {
            const int m = idx / S.in_pad, i = idx - m * S.in_pad;
            Xa[m * LD + i] = (m < rows && (uint32_t)i < n_in) ? ldf(x + (size_t)(r0 + m) * n_in + i) : 0.0f;
        }
        __syncthreads();
        // forward, keeping the pre-activations
        prod<false>(H1, LD, Xa, LD, 1, Wp, S.in_pad, 1, TR, W, S.in_pad);
        __syncthreads();
        float* Hl = H1;                  // pre-activation of the last hidden layer
        if (S.nh == 2) {
            for (int idx = threadIdx.x; idx < TR * W; idx += NT) { const int o = (idx / W) * LD + idx % W; Gb[o] = act_fwd(H1[o], hact); }
            __syncthreads();
            prod<false>(H2, LD, Gb, LD, 1, Wp + S.off2, W, 1, TR, W, W);
            __syncthreads();
            Hl = H2;
        }
        // Gb = act(Hl) (input of the output layer); dz -> Ga[:, 0:OUTP]
        for (int idx = threadIdx.x; idx < TR * W; idx += NT) { const int o = (idx / W) * LD + idx % W; Gb[o] = act_fwd(Hl[o], hact); }
        __syncthreads();
        for (int idx = threadIdx.x; idx < TR * OUTP; idx += NT) {
            const int m = idx / OUTP, o = idx - m * OUTP;
            float dz = 0.0f;
            if (m < rows && (uint32_t)o < n_out) {
                float z = 0.0f;
                for (int j = 0; j < W; j++) z = fmaf(Wp[S.offo + o * W + j], Gb[m * LD + j], z);
                dz = ldf(dy + (size_t)(r0 + m) * ld_dy + o) * act_bwd(z, act_fwd(z, oact), oact);
            }
            Ga[m * LD + o] = dz;
        }
        __syncthreads();
        // dWo[o][j] += sum_m dz[m][o] * act(Hl)[m][j]
        if (dparams) prod<true>(dWp + S.offo, W, Ga, 1, LD, Gb, 1, LD, OUTP, W, TR);
        __syncthreads();
        // dHl[m][j] = (sum_o dz[m][o] * Wo[o][j]) * act'(Hl[m][j])   -> Gb
        prod<false>(Gb, LD, Ga, LD, 1, Wp + S.offo, 1, W, TR, W, OUTP);
        __syncthreads();
        for (int idx = threadIdx.x; idx < TR * W; idx += NT) { const int o = (idx / W) * LD + idx % W; Gb[o] *= act_bwd(Hl[o], 0.0f, hrelu); }
        __syncthreads();
        float* G1 = Gb;                  // gradient w.r.t. H1
        if (S.nh == 2) {
            // Ga = act(H1); dW2[k][j] += sum_m dH2[m][k] * act(H1)[m][j]
            for (int idx = threadIdx.x; idx < TR * W; idx += NT) { const int o = (idx / W) * LD + idx % W; Ga[o] = act_fwd(H1[o], hact); }
            __syncthreads();
            if (dparams) prod<true>(dWp + S.off2, W, Gb, 1, LD, Ga, 1, LD, W, W, TR);
            __syncthreads();
            // dH1[m][j] = (sum_k dH2[m][k] * W2[k][j]) * act'(H1[m][j])   -> Ga
            prod<false>(Ga, LD, Gb, LD, 1, Wp + S.off2, 1, W, TR, W, W);
            __syncthreads();
            for (int idx = threadIdx.x; idx < TR * W; idx += NT) { const int o = (idx / W) * LD + idx % W; Ga[o] *= act_bwd(H1[o], 0.0f, hrelu); }
            __syncthreads();
            G1 = Ga;
        }
        // dW1[j][i] += sum_m dH1[m][j] * x[m][i]
        if (dparams) prod<true>(dWp, S.in_pad, G1, 1, LD, Xa, 1, LD, W, S.in_pad, TR);
        // dx[m][i] = sum_j dH1[m][j] * W1[j][i]
        if (dx) {
            for (int idx = threadIdx.x; idx < rows * (int)n_in; idx += NT) {
                const int m = idx / (int)n_in, i = idx - m * (int)n_in;
                float s = 0.0f;
                for (int j = 0; j < W; j++) s = fmaf(G1[m * LD + j], Wp[j * S.in_pad + i], s);
                XT* p = dx + (size_t)(r0 + m) * n_in + i;
                if (dx_accumulate) s += ldf(p);
                stf(p, s);
            }
        }
    }
    __syncthreads();
    if (dparams) {
        for (int i = threadIdx.x; i < S.n_par; i += NT) {
            const bool pad_row = i >= S.offo && (uint32_t)((i - S.offo) / W) >= n_out;
            const bool pad_col = i < S.off2 && (uint32_t)(i % S.in_pad) >= n_in;
            if (!pad_row && !pad_col && dWp[i] != 0.0f) atomicAdd(dparams + i, dWp[i]);
        }
    }
}

int sm_count_f32() {
    static int n = 0;
    if (n == 0) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}
size_t n_params(uint32_t n_in, uint32_t n_hidden) { return (size_t)W * ((n_in + 15) / 16 * 16) + (n_hidden - 1) * W * W + OUTP * W; }

}  // namespace

// Same argument meaning as nrf_mlp_forward_ex / nrf_mlp_backward_ex, except: params are F32 (same tcnn layout), there is
// no loss_scale (nothing is rounded to f16); x may be f32 or f16 (dx follows it), y and dy are f32.
NRF_EXPORT int nrf_mlp_forward_f32(const void* x, int x_dtype, const float* params_f32, uint32_t B, uint32_t n_in, uint32_t n_out,
                                   uint32_t n_hidden, uint32_t width, int hidden_act, int out_act, float* y, uint32_t ld_y, void* stream) {
    if (B == 0) return NRF_OK;
    if (ld_y == 0) ld_y = n_out;
    if (!x || !params_f32 || !y || ld_y < n_out) return NRF_E_INVALID;
    if (width != W || n_in == 0 || n_in > 64 || n_out == 0 || n_out > 16 || n_hidden < 1 || n_hidden > 2) return NRF_E_UNSUPPORTED;
    if (hidden_act != NRF_ACT_RELU && hidden_act != NRF_ACT_NONE) return NRF_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * (n_params(n_in, n_hidden) + 2 * TR * LD);
    const uint32_t grid = (uint32_t)min((uint64_t)ceil_div_u32(B, TR), (uint64_t)sm_count_f32() * 2);
    if (x_dtype == NRF_DTYPE_F32) {
        auto kern = k_mlp_f32_fwd<float, float>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, NT, smem, s>>>((const float*)x, params_f32, B, n_in, n_out, n_hidden, hidden_act, out_act, y, ld_y);
    } else if (x_dtype == NRF_DTYPE_F16) {
        auto kern = k_mlp_f32_fwd<__half, float>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, NT, smem, s>>>((const __half*)x, params_f32, B, n_in, n_out, n_hidden, hidden_act, out_act, y, ld_y);
    } else return NRF_E_UNSUPPORTED;
    return nrf_check_launch();
}

NRF_EXPORT int nrf_mlp_backward_f32(const void* x, int x_dtype, const float* params_f32, const float* dy, uint32_t ld_dy, uint32_t B,
                                    uint32_t n_in, uint32_t n_out, uint32_t n_hidden, uint32_t width, int hidden_act, int out_act, void* dx,
                                    int dx_accumulate, float* dparams, void* stream) {
    if (B == 0) return NRF_OK;
    if (ld_dy == 0) ld_dy = n_out;
    if (!x || !params_f32 || !dy || ld_dy < n_out) return NRF_E_INVALID;
    if (width != W || n_in == 0 || n_in > 64 || n_out == 0 || n_out > 16 || n_hidden < 1 || n_hidden > 2) return NRF_E_UNSUPPORTED;
    if (hidden_act != NRF_ACT_RELU && hidden_act != NRF_ACT_NONE) return NRF_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * (2 * n_params(n_in, n_hidden) + 5 * TR * LD);
    const uint32_t grid = (uint32_t)min((uint64_t)ceil_div_u32(B, TR), (uint64_t)sm_count_f32());
    if (x_dtype == NRF_DTYPE_F32) {
        auto kern = k_mlp_f32_bwd<float, float>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, NT, smem, s>>>((const float*)x, params_f32, dy, ld_dy, B, n_in, n_out, n_hidden, hidden_act, out_act, (float*)dx,
                                    dx_accumulate, dparams);
    } else if (x_dtype == NRF_DTYPE_F16) {
        auto kern = k_mlp_f32_bwd<__half, float>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, NT, smem, s>>>((const __half*)x, params_f32, dy, ld_dy, B, n_in, n_out, n_hidden, hidden_act, out_act, (__half*)dx,
                                    dx_accumulate, dparams);
    } else return NRF_E_UNSUPPORTED;
    return nrf_check_launch();
}
