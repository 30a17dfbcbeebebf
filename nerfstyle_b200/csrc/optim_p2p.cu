// The data-parallel exchange of the train step fused INTO the optimizer kernels, over NVLink / NVSwitch peer memory
// (SURVEY.md 8e / 5.8): instead of  ncclReduceScatter(table gradients) -> Adam on the shard -> ncclAllGather(fp16 tables)
// plus two small all-reduces, every rank runs ONE kernel over its 1/N row shard that
//   * reduces the rows of the interleaved gradient buffers of ALL ranks straight out of their memory -- one
//     multimem.ld_reduce per 16-byte row when the buffers are bound to an NVLS multicast object (the reduction happens in the
//     switch and only the sum crosses this GPU's links), else a loop of peer loads in rank order;
//   * applies GradScaler unscale / Adam / LambdaLR / EMA exactly like k_adam_ema_pair (optim.cu);
//   * writes the updated fp16 row into the table copy of EVERY rank (one multimem.st, or a loop of peer stores) -- the
//     all-gather.
// The buffers are torch symmetric-memory allocations (torch.distributed._symmetric_memory: cuMem + fabric handles exchanged
// at rendezvous); two symmetric-memory barriers per step bracket the kernel (all ranks' gradients complete / all ranks'
// table writes landed), the second one on a side stream under the next step's ray marching.  The small tensors (MLP
// gradients, 61 KB) and the GradScaler's found-inf flag travel through a third symmetric buffer and are summed by every
// rank in rank order -- identical bits on all ranks, no NCCL call in the step.
#include "common.cuh"
#include "optim_state.cuh"

namespace {

__device__ __forceinline__ float4 mc_ld_reduce_f32x4(const float* mc_addr) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc_addr) : "memory");
    return r;
}
__device__ __forceinline__ void mc_st_b32x2(void* mc_addr, uint32_t a, uint32_t b) {
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};" ::"l"(mc_addr), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)) : "memory");
}

// out[i] = sum over ranks (in rank order) of peer[r][i], i < n; element n of every peer buffer is that rank's found-inf flag
__global__ void k_small_allreduce_p2p(const uint64_t* __restrict__ peer_ptrs, uint32_t world, uint32_t n, float* __restrict__ out,
                                      OptState* __restrict__ st) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        float s = 0.0f;
        for (uint32_t r = 0; r < world; r++) s += reinterpret_cast<const volatile float*>(peer_ptrs[r])[i];
        if (i < n) out[i] = s;
        else if (s != 0.0f) st->found_inf = 1;          // an inf / nan seen by ANY rank skips the step everywhere
    }
}

template <bool MC>
__global__ void __launch_bounds__(256)
k_adam_ema_pair_p2p(float* __restrict__ p0, float* __restrict__ p1, const uint64_t* __restrict__ grad_ptrs, const float* __restrict__ grad_mc,
                    const uint64_t* __restrict__ half_ptrs, uint8_t* __restrict__ half_mc, uint32_t world, uint64_t row_lo,
                    float* __restrict__ m0, float* __restrict__ m1, float* __restrict__ v0, float* __restrict__ v1,
                    float* __restrict__ ema0, float* __restrict__ ema1, uint64_t rows, const OptState* __restrict__ st, float lr0,
                    float lr_decay_steps, float beta1, float beta2, float eps, float ema_one_minus_decay) {
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
        const uint64_t row = row_lo + i;          // row of the full interleaved buffers; i indexes this rank's shard-local state
        float4 g;
        if (MC) {
            g = mc_ld_reduce_f32x4(grad_mc + 4 * row);
        } else {
            g = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            for (uint32_t r = 0; r < world; r++) {
                const float4 q = __ldcv(reinterpret_cast<const float4*>(grad_ptrs[r]) + row);
                g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
            }
        }
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
            const __half2 ha = __floats2half2_rn(pa.x, pa.y), hb = __floats2half2_rn(pb.x, pb.y);
            const uint32_t ua = *reinterpret_cast<const uint32_t*>(&ha), ub = *reinterpret_cast<const uint32_t*>(&hb);
            if (MC) {
                mc_st_b32x2(half_mc + 8 * row, ua, ub);
            } else {
                for (uint32_t r = 0; r < world; r++) reinterpret_cast<uint2*>(half_ptrs[r])[row] = make_uint2(ua, ub);
            }
        }
        if (ema0) reinterpret_cast<float2*>(ema0)[i] = ea;
        if (ema1) reinterpret_cast<float2*>(ema1)[i] = eb;
    }
}

}  // namespace

// peer_ptrs_dev: device array of `world` 64-bit addresses -- buffer r of the symmetric allocation as mapped into THIS
// process (torch SymmetricMemory.buffer_ptrs); every buffer holds n floats + this rank's found-inf flag (as a float) at [n].
NRF_EXPORT int nrf_small_allreduce_p2p(const uint64_t* peer_ptrs_dev, uint32_t world, uint32_t n, float* out, void* state,
                                       void* stream) {
    if (!peer_ptrs_dev || !out || !state || world == 0) return NRF_E_INVALID;
    k_small_allreduce_p2p<<<ceil_div_u32((uint64_t)n + 1, 256), 256, 0, (cudaStream_t)stream>>>(peer_ptrs_dev, world, n, out, (OptState*)state);
    return nrf_check_launch();
}

// The fused reduce-scatter + Adam/EMA + all-gather over this rank's rows [row_lo, row_lo + rows) of the interleaved pair
// buffers.  param0/param1/exp_avg*/exp_avg_sq*/ema* point at the shard (element 0 = row row_lo).  grad_mc / half_mc: the
// multicast addresses of the two symmetric buffers (both non-NULL -> NVLS path), else the peer-pointer loops are used.
NRF_EXPORT int nrf_adam_step_pair_p2p(float* param0, float* param1, const uint64_t* grad_ptrs_dev, const float* grad_mc,
                                      const uint64_t* half_ptrs_dev, void* half_mc, uint32_t world, uint64_t row_lo,
                                      float* exp_avg0, float* exp_avg1, float* exp_avg_sq0, float* exp_avg_sq1, float* ema0,
                                      float* ema1, uint64_t rows, const void* state, float lr0, float lr_decay_steps, float beta1,
                                      float beta2, float eps, float ema_one_minus_decay, void* stream) {
    if (rows == 0) return NRF_OK;
    if (!param0 || !param1 || !grad_ptrs_dev || !half_ptrs_dev || !exp_avg0 || !exp_avg1 || !exp_avg_sq0 || !exp_avg_sq1 || !state ||
        world == 0) return NRF_E_INVALID;
    const uint32_t nb = (uint32_t)min((uint64_t)148 * 8, (rows + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (grad_mc && half_mc)
        k_adam_ema_pair_p2p<true><<<nb, 256, 0, s>>>(param0, param1, grad_ptrs_dev, grad_mc, half_ptrs_dev, (uint8_t*)half_mc, world, row_lo,
                                                    exp_avg0, exp_avg1, exp_avg_sq0, exp_avg_sq1, ema0, ema1, rows, (const OptState*)state,
                                                    lr0, lr_decay_steps, beta1, beta2, eps, ema_one_minus_decay);
    else
        k_adam_ema_pair_p2p<false><<<nb, 256, 0, s>>>(param0, param1, grad_ptrs_dev, nullptr, half_ptrs_dev, nullptr, world, row_lo,
                                                     exp_avg0, exp_avg1, exp_avg_sq0, exp_avg_sq1, ema0, ema1, rows, (const OptState*)state,
                                                     lr0, lr_decay_steps, beta1, beta2, eps, ema_one_minus_decay);
    return nrf_check_launch();
}
