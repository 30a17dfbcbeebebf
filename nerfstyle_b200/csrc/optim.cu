// Fused optimizer step for the hash tables and MLP weights (SURVEY.md 8f NEXT-3): GradScaler unscale + inf check,
// Adam(eps=1e-15) with bias correction and the LambdaLR decay, torch_ema update and the fp16 table copy the next
// forward needs -- ONE pass over each parameter instead of torch's unscale / Adam / EMA / .to(half) passes.
// Semantics follow trainers/base.py:216-229,420-426 (torch.optim.Adam, GradScaler, LambdaLR, torch_ema).
// Everything that GradScaler decides on the host (skip on inf, scale growth/back-off, "scheduler.step() only when
// the scale did not shrink") is decided on the device from a small state block, so the step has no host sync.
#include "common.cuh"

#include "optim_state.cuh"

__global__ void k_grads_check(const float* __restrict__ g, uint64_t n, OptState* st) {
    bool bad = false;
    const uint64_t n4 = ((((uintptr_t)g) & 15) == 0) ? n / 4 : 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        bad |= !(fabsf(v.x) <= 3.402823466e38f) || !(fabsf(v.y) <= 3.402823466e38f) || !(fabsf(v.z) <= 3.402823466e38f) ||
               !(fabsf(v.w) <= 3.402823466e38f);      // inf or nan
    }
    for (uint64_t i = 4 * n4 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float v = g[i];
        bad |= !(fabsf(v) <= 3.402823466e38f);
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) st->found_inf = 1;
}

__global__ void k_adam_ema(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                           float* __restrict__ ema, __half* __restrict__ p_half, uint64_t n, const OptState* __restrict__ st,
                           float lr0, float lr_decay_steps, float beta1, float beta2, float eps, float ema_one_minus_decay,
                           uint32_t g_stride, uint32_t h_stride) {
    // g_stride / h_stride: distance, in 2-element rows, between consecutive rows of the gradient / of the fp16 copy
    // (1 = contiguous; 2 = the interleaved [row][encoder][2] buffers of the paired hash-grid kernels, pointers pre-offset
    // to this tensor's encoder slot)
    const bool skip = st->found_inf != 0;
    const int t = st->good_steps + 1;
    const float inv_scale = 1.0f / st->scale;
    const float lr = (lr_decay_steps > 0.0f) ? lr0 * exp2f(-3.3219280948873623f * ((float)st->good_steps / lr_decay_steps)) : lr0;
    const float bc1 = 1.0f - powf(beta1, (float)t);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)t));
    const float step_size = lr / bc1;
    auto update = [&](float pi, float gi_raw, float& mi, float& vi, float& e) -> float {
        if (!skip) {
            const float gi = gi_raw * inv_scale;
            mi = beta1 * mi + (1.0f - beta1) * gi;
            vi = beta2 * vi + (1.0f - beta2) * gi * gi;
            pi -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
        }
        e = e - ema_one_minus_decay * (e - pi);
        return pi;
    };
    // 8-byte vector path (every pointer 8-byte aligned, the fp16 copy 4-byte): table shards of the data-parallel optimizer
    // start at element offsets that are even but not multiples of four
    const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)ema) & 7) == 0) && ((((uintptr_t)p_half) & 3) == 0);
    const uint64_t n2 = vec ? n / 2 : 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * blockDim.x) {
        float2 pp = reinterpret_cast<float2*>(p)[i];
        const float2 gg = reinterpret_cast<const float2*>(g)[i * g_stride];
        float2 mm = reinterpret_cast<float2*>(m)[i], vv = reinterpret_cast<float2*>(v)[i];
        float2 ee = ema ? reinterpret_cast<float2*>(ema)[i] : make_float2(0.0f, 0.0f);
        pp.x = update(pp.x, gg.x, mm.x, vv.x, ee.x);
        pp.y = update(pp.y, gg.y, mm.y, vv.y, ee.y);
        if (!skip) {
            reinterpret_cast<float2*>(m)[i] = mm;
            reinterpret_cast<float2*>(v)[i] = vv;
            reinterpret_cast<float2*>(p)[i] = pp;
            if (p_half) reinterpret_cast<__half2*>(p_half)[i * h_stride] = __floats2half2_rn(pp.x, pp.y);
        }
        if (ema) reinterpret_cast<float2*>(ema)[i] = ee;
    }
    for (uint64_t i = 2 * n2 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        float mi = m[i], vi = v[i], e = ema ? ema[i] : 0.0f;
        const float pi = update(p[i], g[(i >> 1) * 2 * g_stride + (i & 1)], mi, vi, e);
        if (!skip) {
            m[i] = mi; v[i] = vi; p[i] = pi;
            if (p_half) p_half[(i >> 1) * 2 * h_stride + (i & 1)] = __float2half_rn(pi);
        }
        if (ema) ema[i] = e;
    }
}

// Both tables of an interleaved pair in ONE pass: row i of the gradient buffer [row][table][2] is one 16-byte load and row i of
// the fp16 copy one 8-byte store (the strided single-table form touches every sector of those buffers twice).
__global__ void k_adam_ema_pair(float* __restrict__ p0, float* __restrict__ p1, const float4* __restrict__ gpair,
                                float* __restrict__ m0, float* __restrict__ m1, float* __restrict__ v0, float* __restrict__ v1,
                                float* __restrict__ ema0, float* __restrict__ ema1, uint2* __restrict__ hpair, uint64_t rows,
                                const OptState* __restrict__ st, float lr0, float lr_decay_steps, float beta1, float beta2, float eps,
                                float ema_one_minus_decay) {
    const bool skip = st->found_inf != 0;
    const int t = st->good_steps + 1;
    const float inv_scale = 1.0f / st->scale;
    const float lr = (lr_decay_steps > 0.0f) ? lr0 * exp2f(-3.3219280948873623f * ((float)st->good_steps / lr_decay_steps)) : lr0;
    const float bc1 = 1.0f - powf(beta1, (float)t);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)t));
    const float step_size = lr / bc1;
    auto update = [&](float pi, float gi_raw, float& mi, float& vi, float& e) -> float {     // same operations as k_adam_ema
        if (!skip) {
            const float gi = gi_raw * inv_scale;
            mi = beta1 * mi + (1.0f - beta1) * gi;
            vi = beta2 * vi + (1.0f - beta2) * gi * gi;
            pi -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
        }
        e = e - ema_one_minus_decay * (e - pi);
        return pi;
    };
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < rows; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 g = gpair[i];
        float2 pa = reinterpret_cast<float2*>(p0)[i], pb = reinterpret_cast<float2*>(p1)[i];
        float2 ma = reinterpret_cast<float2*>(m0)[i], mb = reinterpret_cast<float2*>(m1)[i];
        float2 va = reinterpret_cast<float2*>(v0)[i], vb = reinterpret_cast<float2*>(v1)[i];
        float2 ea = ema0 ? reinterpret_cast<float2*>(ema0)[i] : make_float2(0.0f, 0.0f);
        float2 eb = ema1 ? reinterpret_cast<float2*>(ema1)[i] : make_float2(0.0f, 0.0f);
        pa.x = update(pa.x, g.x, ma.x, va.x, ea.x); pa.y = update(pa.y, g.y, ma.y, va.y, ea.y);
        pb.x = update(pb.x, g.z, mb.x, vb.x, eb.x); pb.y = update(pb.y, g.w, mb.y, vb.y, eb.y);
        if (!skip) {
            reinterpret_cast<float2*>(m0)[i] = ma; reinterpret_cast<float2*>(m1)[i] = mb;
            reinterpret_cast<float2*>(v0)[i] = va; reinterpret_cast<float2*>(v1)[i] = vb;
            reinterpret_cast<float2*>(p0)[i] = pa; reinterpret_cast<float2*>(p1)[i] = pb;
            if (hpair) {
                const __half2 ha = __floats2half2_rn(pa.x, pa.y), hb = __floats2half2_rn(pb.x, pb.y);
                hpair[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&ha), *reinterpret_cast<const uint32_t*>(&hb));
            }
        }
        if (ema0) reinterpret_cast<float2*>(ema0)[i] = ea;
        if (ema1) reinterpret_cast<float2*>(ema1)[i] = eb;
    }
}

// GradScaler.update(): back off on inf, grow after `growth_interval` finite steps; count the optimizer steps taken
__global__ void k_scaler_update(OptState* st, float growth, float backoff, int growth_interval) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (st->found_inf) {
        st->scale *= backoff;
        st->growth_tracker = 0;
    } else {
        st->good_steps += 1;
        if (++st->growth_tracker == growth_interval) { st->scale *= growth; st->growth_tracker = 0; }
    }
    st->found_inf = 0;
}

NRF_EXPORT uint64_t nrf_opt_state_bytes(void) { return sizeof(OptState); }

NRF_EXPORT int nrf_grads_check(const float* grad, uint64_t n, void* state, void* stream) {
    if (n == 0) return NRF_OK;
    if (!grad || !state) return NRF_E_INVALID;
    const uint32_t nb = (uint32_t)min((uint64_t)148 * 8, (n + 1023) / 1024);
    k_grads_check<<<nb, 256, 0, (cudaStream_t)stream>>>(grad, n, (OptState*)state);
    return nrf_check_launch();
}

NRF_EXPORT int nrf_adam_step_ex(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, void* param_half,
                                uint64_t n, const void* state, float lr0, float lr_decay_steps, float beta1, float beta2,
                                float eps, float ema_one_minus_decay, uint32_t grad_row_stride, uint32_t half_row_stride,
                                void* stream) {
    if (n == 0) return NRF_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !state) return NRF_E_INVALID;
    if (grad_row_stride == 0 || half_row_stride == 0) return NRF_E_INVALID;
    if ((grad_row_stride > 1 || half_row_stride > 1) && (n & 1)) return NRF_E_INVALID;      // strided forms address whole rows
    const uint32_t nb = (uint32_t)min((uint64_t)148 * 16, (n + 255) / 256);
    k_adam_ema<<<nb, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, ema, (__half*)param_half, n,
                                                    (const OptState*)state, lr0, lr_decay_steps, beta1, beta2, eps,
                                                    ema_one_minus_decay, grad_row_stride, half_row_stride);
    return nrf_check_launch();
}

NRF_EXPORT int nrf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, void* param_half,
                             uint64_t n, const void* state, float lr0, float lr_decay_steps, float beta1, float beta2,
                             float eps, float ema_one_minus_decay, void* stream) {
    return nrf_adam_step_ex(param, grad, exp_avg, exp_avg_sq, ema, param_half, n, state, lr0, lr_decay_steps, beta1, beta2, eps,
                            ema_one_minus_decay, 1, 1, stream);
}

NRF_EXPORT int nrf_adam_step_pair(float* param0, float* param1, const float* grad_pair, float* exp_avg0, float* exp_avg1,
                                  float* exp_avg_sq0, float* exp_avg_sq1, float* ema0, float* ema1, void* half_pair, uint64_t rows,
                                  const void* state, float lr0, float lr_decay_steps, float beta1, float beta2, float eps,
                                  float ema_one_minus_decay, void* stream) {
    if (rows == 0) return NRF_OK;
    if (!param0 || !param1 || !grad_pair || !exp_avg0 || !exp_avg1 || !exp_avg_sq0 || !exp_avg_sq1 || !state) return NRF_E_INVALID;
    if ((((uintptr_t)grad_pair) & 15) || (((uintptr_t)half_pair) & 7) ||
        (((uintptr_t)param0 | (uintptr_t)param1 | (uintptr_t)exp_avg0 | (uintptr_t)exp_avg1 | (uintptr_t)exp_avg_sq0 |
          (uintptr_t)exp_avg_sq1 | (uintptr_t)ema0 | (uintptr_t)ema1) & 7)) return NRF_E_INVALID;
    const uint32_t nb = (uint32_t)min((uint64_t)148 * 16, (rows + 255) / 256);
    k_adam_ema_pair<<<nb, 256, 0, (cudaStream_t)stream>>>(param0, param1, reinterpret_cast<const float4*>(grad_pair), exp_avg0, exp_avg1,
                                                         exp_avg_sq0, exp_avg_sq1, ema0, ema1, reinterpret_cast<uint2*>(half_pair), rows,
                                                         (const OptState*)state, lr0, lr_decay_steps, beta1, beta2, eps,
                                                         ema_one_minus_decay);
    return nrf_check_launch();
}

NRF_EXPORT int nrf_scaler_update(void* state, float growth, float backoff, int growth_interval, void* stream) {
    if (!state) return NRF_E_INVALID;
    k_scaler_update<<<1, 32, 0, (cudaStream_t)stream>>>((OptState*)state, growth, backoff, growth_interval);
    return nrf_check_launch();
}
