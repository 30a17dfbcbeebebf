// Reconstruction loss head in one launch: the tail of Renderer.render_train (white background,
// /root/reference/renderer.py:229-232) + Trainer.calc_loss (trainers/base.py:251-304: MSE on the rgb channels + class_lambda *
// cross-entropy on the class channels) + the gradients of both w.r.t. the compositing outputs.  The reference runs ~25
// elementwise / reduction / slicing kernels (and their backward twins) on N x (3 + K) floats here; N = 8192 rays per step.
#include "common.cuh"

#define LOSS_MAX_CH 35      // 3 + up to 32 classes

// Thread per ray, LOSS_BLOCK rays per block; the block sums are combined in block order by whichever block finishes last
// (deterministic: fixed reduction tree inside a block, fixed order across blocks).  image [N, Cch] = (rgb, class logits) as
// composited, weights_sum [N].
//   rgb    = image[:, :3] + (1 - weights_sum)                       (renderer.py:231)
//   mse    = mean((rgb - target_rgb)^2) over N x 3                    (base.py:272)
//   ce     = mean_n(logsumexp(logits_n) - logits_n[target_cls_n])     (nn.CrossEntropyLoss, base.py:282)
//   out    = {mse + class_lambda * ce, mse, ce}
//   grad_image [N, Cch], grad_ws [N] = d out[0] / d image, d out[0] / d weights_sum
#define LOSS_BLOCK 128

struct LossScratch {                 // device scratch: per-block partial sums + the arrival counter (zero before first use)
    unsigned int arrived;
    unsigned int pad;
    double partial[1];               // [2 * nblocks]
};

__global__ void __launch_bounds__(LOSS_BLOCK)
k_recon_loss(const float* __restrict__ image, const float* __restrict__ weights_sum, const float* __restrict__ target_rgb,
             const long long* __restrict__ target_cls, uint32_t N, uint32_t Cch, float class_lambda, float* __restrict__ out,
             float* __restrict__ grad_image, float* __restrict__ grad_ws, LossScratch* __restrict__ sc) {
    __shared__ double s_mse[LOSS_BLOCK / 32], s_ce[LOSS_BLOCK / 32];
    __shared__ bool s_last;
    const uint32_t K = Cch - 3;
    const float inv_n = 1.0f / (float)N, inv_3n = 1.0f / (3.0f * (float)N);
    double mse = 0.0, ce = 0.0;
    const uint32_t n = blockIdx.x * LOSS_BLOCK + threadIdx.x;
    if (n < N) {
        const float* row = image + (size_t)n * Cch;
        float* grow = grad_image + (size_t)n * Cch;
        const float bg = 1.0f - __ldg(weights_sum + n);
        float gws = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float d = (__ldg(row + c) + bg) - __ldg(target_rgb + 3 * (size_t)n + c);
            mse += (double)d * (double)d;
            const float g = 2.0f * d * inv_3n;
            grow[c] = g;
            gws -= g;
        }
        grad_ws[n] = gws;
        if (K > 0) {
            float mx = -3.402823466e38f;
            for (uint32_t k = 0; k < K; k++) mx = fmaxf(mx, __ldg(row + 3 + k));
            float se = 0.0f;
            for (uint32_t k = 0; k < K; k++) se += __expf(__ldg(row + 3 + k) - mx);
            const float lse = mx + __logf(se);
            const long long t = __ldg(target_cls + n);
            const float inv_se = 1.0f / se;
            for (uint32_t k = 0; k < K; k++) {
                const float l = __ldg(row + 3 + k);
                const float p = __expf(l - mx) * inv_se;
                grow[3 + k] = class_lambda * inv_n * (p - ((long long)k == t ? 1.0f : 0.0f));
                if ((long long)k == t) ce += (double)(lse - l);
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { mse += __shfl_xor_sync(NRF_FULL_MASK, mse, d); ce += __shfl_xor_sync(NRF_FULL_MASK, ce, d); }
    if ((threadIdx.x & 31) == 0) { s_mse[threadIdx.x >> 5] = mse; s_ce[threadIdx.x >> 5] = ce; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < LOSS_BLOCK / 32; w++) { a += s_mse[w]; b += s_ce[w]; }
        sc->partial[2 * blockIdx.x] = a;
        sc->partial[2 * blockIdx.x + 1] = b;
        __threadfence();
        s_last = atomicAdd(&sc->arrived, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        double a = 0.0, b = 0.0;
        for (uint32_t i = 0; i < gridDim.x; i++) { a += sc->partial[2 * i]; b += sc->partial[2 * i + 1]; }
        const float m = (float)(a / (3.0 * (double)N)), c = K > 0 ? (float)(b / (double)N) : 0.0f;
        out[0] = m + class_lambda * c;
        out[1] = m;
        out[2] = c;
        sc->arrived = 0;             // ready for the next call on the stream
    }
}

NRF_EXPORT uint64_t nrf_recon_loss_scratch_bytes(uint32_t N) {
    return 16 + 16 * (uint64_t)((N + LOSS_BLOCK - 1) / LOSS_BLOCK);
}

// scratch: nrf_recon_loss_scratch_bytes(N) bytes, 8-byte aligned, ZERO before its first use (the kernel leaves it zeroed)
NRF_EXPORT int nrf_recon_loss(const float* image, const float* weights_sum, const float* target_rgb, const int64_t* target_cls,
                              uint32_t N, uint32_t Cch, float class_lambda, float* out, float* grad_image, float* grad_ws,
                              void* scratch, void* stream) {
    if (N == 0) return NRF_E_INVALID;
    if (!image || !weights_sum || !target_rgb || !out || !grad_image || !grad_ws || !scratch) return NRF_E_INVALID;
    if (Cch < 3 || Cch > LOSS_MAX_CH || (Cch > 3 && !target_cls) || (((uintptr_t)scratch) & 7)) return NRF_E_INVALID;
    k_recon_loss<<<(N + LOSS_BLOCK - 1) / LOSS_BLOCK, LOSS_BLOCK, 0, (cudaStream_t)stream>>>(
        image, weights_sum, target_rgb, reinterpret_cast<const long long*>(target_cls), N, Cch, class_lambda, out, grad_image, grad_ws,
        reinterpret_cast<LossScratch*>(scratch));
    return nrf_check_launch();
}
