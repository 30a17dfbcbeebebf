// tcnn-style fully fused MLP on the 5th-generation tensor cores (tcgen05 + TMEM) -- forward and backward.
//
// Same contract as mlp.cu (call sites /root/reference/networks/style_nerf.py:44-98):
// y = act_out(W_n relu(... relu(W_1 x))), f16 operands, f32 accumulation, hidden activations rounded to f16.
//
// Design (one CTA = one 128-row tile at a time, persistent over tiles, several CTAs resident per SM):
//   * warps 0-3 are the row owners: thread t owns row t of the tile == TMEM lane t.  Warp 4 is the MMA issuer (one
//     elected lane).  Row owners and issuer hand tiles back and forth through mbarriers only (`ready`: 128 arrivals,
//     operands written / accumulator drained;  `done`: tcgen05.commit of the MMAs a row owner waits for).
//   * every operand tile lives in shared memory in the "chunked" SWIZZLE_NONE layout of tc05.cuh, which is a legal
//     K-major AND MN-major UMMA operand: the backward forms dH = dZ W, dX = dH W1 and dW = dH^T X from the very same
//     tiles with no transposed copy;
//   * accumulators sit in TMEM; a row owner pulls its row with tcgen05.ld, rounds to f16, applies relu / relu' on packed
//     halfs (HMNMX2 / HSET2+LOP3) and writes the next layer's operand back to shared memory;
//   * measured on B200 every tcgen05.mma costs >= 45 cycles whatever its shape below N=64 (tools/tc_probe.cu), so the
//     kernel minimises the NUMBER of MMAs: the weight gradients of a tile are ONE stacked product
//     [dH1 | Hlast]^T [X | dZ]  (M=128, N=in+16, K=128 rows: 8 MMAs) whose diagonal blocks are dW1 and dWo^T; it is
//     accumulated in TMEM across all tiles of the persistent CTA and flushed once with float atomics.  It runs on the
//     tensor pipe while the row owners are already in the next epilogue (own mbarrier `wdone`);
//   * the backward recomputes the hidden activations from x (nothing but x is saved by the forward); for a linear
//     output layer dZ = loss_scale * dy needs no recomputation of Z, which removes one MMA round trip;
//   * the rows of tile i+1 (x, dy) are fetched from HBM into registers while tile i is processed.
// The tensor pipe is nearly idle by design (a 64-wide MLP is ~20 FLOP/B): what the tcgen05 path buys is that the
// accumulators, transposes and weight-gradient reductions leave the register file / LSU.
#include "common.cuh"
#include "tc05.cuh"
#include "mlp_tc_common.cuh"

namespace {

struct TcSmem { uint32_t W1, Wh, Wo, X, H1, H2, dZ, dH1, dH2, total; };

template <int IN_KT, int NH>
__host__ __device__ constexpr TcSmem fwd_smem() {
    constexpr uint32_t CH = ch_for(IN_KT);
    TcSmem s{};
    uint32_t o = 0;
    s.W1 = o; o += IN_KT * 2 * CHW;
    s.Wh = o; o += (NH - 1) * 8 * CHW;
    s.Wo = o; o += 8 * CHO;
    s.X = o;  o += IN_KT * 2 * CH;
    s.H1 = o; o += 8 * CH;
    s.total = o;
    return s;
}
// The backward keeps tiles that are multiplied as ONE stacked operand adjacent (chunk stride CH continues across them):
//   NH=1:  [X | dZ]  and  [dH1 | H1]          NH=2:  [X | H1]  and  [dH1 | dH2]
template <int IN_KT, int NH>
__host__ __device__ constexpr TcSmem bwd_smem() {
    constexpr uint32_t CH = ch_for(IN_KT);
    TcSmem s{};
    uint32_t o = 0;
    s.W1 = o; o += IN_KT * 2 * CHW;
    s.Wh = o; o += (NH - 1) * 8 * CHW;
    s.Wo = o; o += 8 * CHO;
    if (NH == 1) {
        s.X = o;   o += IN_KT * 2 * CH;
        s.dZ = o;  o += 2 * CH;
        s.dH1 = o; o += 8 * CH;
        s.H1 = o;  o += 8 * CH;
        s.H2 = o; s.dH2 = o;
    } else {
        s.X = o;   o += IN_KT * 2 * CH;
        s.H1 = o;  o += 8 * CH;
        s.dZ = o;  o += 2 * CH;
        s.dH1 = o; o += 8 * CH;
        s.dH2 = o; o += 8 * CH;
        s.H2 = o;  o += 8 * CH;
    }
    s.total = o;
    return s;
}

// ------------------------------------------------------------------------------------------------ forward
template <int IN_KT, int NH>
__global__ void __launch_bounds__(TC_THREADS, 6)      // 64 registers: six CTAs really are resident (68 silently made it five)
k_mlp_fwd_tc(const void* __restrict__ x, int x_dt, const __half* __restrict__ params, uint32_t B, uint32_t n_in, uint32_t n_out,
             int hidden_act, int out_act, void* __restrict__ y, int y_dt, uint32_t ld_y, const int32_t* __restrict__ B_dev) {
    constexpr int IN_PAD = IN_KT * 16;
    if (B_dev) B = min(B, (uint32_t)*B_dev);      // device-driven inference loop: the launch is sized for the cap
    if (blockIdx.x >= (B + 127) / 128) return;    // a CTA without a tile leaves before it allocates TMEM / stages weights
    constexpr int CPR = IN_PAD / 8;                      // 16-byte chunks per input row
    constexpr uint32_t CH = ch_for(IN_KT);
    constexpr TcSmem L = fwd_smem<IN_KT, NH>();
    constexpr uint32_t TCOLS = 64;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_ready, bar_done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    stage_weight(smem + L.W1, CHW, params, 64, IN_PAD);
    if constexpr (NH == 2) stage_weight(smem + L.Wh, CHW, params + 64 * IN_PAD, 64, 64);
    stage_weight(smem + L.Wo, CHO, params + 64 * IN_PAD + (NH - 1) * 64 * 64, 16, 64);
    if (warp == 0) { tc05::tmem_alloc(&tmem_slot, TCOLS); tc05::tmem_relinquish(); }
    if (tid == 0) { tc05::mbar_init(&bar_ready, TC_ROWS); tc05::mbar_init(&bar_done, 1); tc05::fence_mbar_init(); }
    publish_and_sync();
    tc05::fence_after_sync();
    const uint32_t tacc = tmem_slot;
    const uint32_t ntiles = (B + 127) / 128;

    if (warp == 4) {
        // ================================================================ MMA issuer (one lane)
        if (lane == 0) {
            const uint64_t dX = tc05::desc_kmajor(tc05::smem_u32(smem + L.X), CH), dH = tc05::desc_kmajor(tc05::smem_u32(smem + L.H1), CH);
            const uint64_t dW1 = tc05::desc_kmajor(tc05::smem_u32(smem + L.W1), CHW), dWh = tc05::desc_kmajor(tc05::smem_u32(smem + L.Wh), CHW);
            const uint64_t dWo = tc05::desc_kmajor(tc05::smem_u32(smem + L.Wo), CHO);
            constexpr uint32_t ID_H = tc05::idesc_f16(128, 64, false, false);
            constexpr uint32_t ID_O = tc05::idesc_f16(128, 16, false, false);
            uint32_t ph = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
#pragma unroll
                for (int k = 0; k < IN_KT; k++) tc05::mma_f16(tacc, dadv(dX, k * 2 * CH), dadv(dW1, k * 2 * CHW), ID_H, k > 0);
                tc05::mma_commit(&bar_done);
                if constexpr (NH == 2) {
                    tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
#pragma unroll
                    for (int k = 0; k < 4; k++) tc05::mma_f16(tacc, dadv(dH, k * 2 * CH), dadv(dWh, k * 2 * CHW), ID_H, k > 0);
                    tc05::mma_commit(&bar_done);
                }
                tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
#pragma unroll
                for (int k = 0; k < 4; k++) tc05::mma_f16(tacc, dadv(dH, k * 2 * CH), dadv(dWo, k * 2 * CHO), ID_O, k > 0);
                tc05::mma_commit(&bar_done);
            }
        }
        __syncwarp();
    } else {
        // ================================================================ row owners
        const uint32_t tacc_lane = tacc + ((uint32_t)(warp * 32) << 16);
        const bool relu = hidden_act == NRF_ACT_RELU;
        const bool x_vec = vec_ok_for(x, n_in, n_in);
        const bool y_vec = vec_ok_for(y, n_out, ld_y);
        // the network's output precision is f16 (tcnn); an f32 `y` holds that f16 value widened, except for TRUNC_EXP
        // whose exp is evaluated in f32 on the f16-rounded pre-activation (tcnn_nerf.py:55-62)
        const bool round_y = y_dt == NRF_DTYPE_F32 && out_act != NRF_ACT_TRUNC_EXP;
        // (a 256-bit store of the 16 f16 outputs of color1_net was measured 60 % SLOWER for that launch and dropped)
        uint32_t phase = 0;
        uint8_t* hrow = smem + L.H1 + tid * 16;
        const bool coal = x_vec && x_dt == NRF_DTYPE_F16 && n_in == (uint32_t)IN_PAD && (reinterpret_cast<uintptr_t>(x) & 31) == 0;
        uint4 xr[CPR];
        load_x_tile<CPR, false>(xr, x, x_dt, (size_t)blockIdx.x * 128, true, B, n_in, x_vec, coal, warp, lane, tid);
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const size_t row = (size_t)tile * 128 + tid;
            const bool row_ok = row < B;
            stage_x_tile<CPR, false>(xr, smem + L.X, CH, coal, warp, lane, tid);
            publish(&bar_ready);
            {   // prefetch the next tile's rows (in flight until the top of the next iteration)
                const uint32_t nt = tile + gridDim.x;
                load_x_tile<CPR, false>(xr, x, x_dt, (size_t)nt * 128, nt < ntiles, B, n_in, x_vec, coal, warp, lane, tid);
            }
#pragma unroll
            for (int l = 0; l < NH; l++) {
                tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
                hidden_fwd_epilogue(tacc_lane, hrow, relu, CH);
                publish(&bar_ready);
            }
            tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
            uint32_t z[16];
            tc05::tmem_ld16(tacc_lane, z);
            tc05::tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    if (8u * c < n_out) {
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            f[j] = tact_fwd(__uint_as_float(z[8 * c + j]), out_act);
                            if (round_y) f[j] = __half2float(__float2half_rn(f[j]));
                        }
                        store_chunk(y, y_dt, row, c, n_out, ld_y, f, y_vec);
                    }
                }
            }
            // the next tile's publish() orders these TMEM reads before the issuer overwrites the accumulator
        }
    }
    publish_and_sync();
    if (warp == 0) tc05::tmem_dealloc(tacc, TCOLS);
}

// ------------------------------------------------------------------------------------------------ backward
template <int IN_KT, int NH>
__host__ __device__ constexpr uint32_t bwd_tmem_cols() {
    // accumulator (64) + stacked weight gradient (in + 16 | in + 64) + dWo^T (16, NH == 2 only)
    const uint32_t need = 64 + IN_KT * 16 + (NH == 1 ? 16 : 64 + 16);
    return need <= 128 ? 128u : 256u;
}

template <int IN_KT, int NH>
__global__ void __launch_bounds__(TC_THREADS, (NH == 2 || IN_KT == 4) ? 2 : 4)
k_mlp_bwd_tc(const void* __restrict__ x, int x_dt, const __half* __restrict__ params, const void* __restrict__ dy, int dy_dt,
             uint32_t ld_dy, uint32_t B, uint32_t n_in, uint32_t n_out, int hidden_act, int out_act, float loss_scale,
             void* __restrict__ dx, int dx_accumulate, float* __restrict__ dparams, unsigned long long* __restrict__ prof) {
    constexpr int IN_PAD = IN_KT * 16;
    constexpr int CPR = IN_PAD / 8;                      // 16-byte chunks per input row
    constexpr uint32_t CH = ch_for(IN_KT);
    constexpr TcSmem L = bwd_smem<IN_KT, NH>();
    constexpr uint32_t TCOLS = bwd_tmem_cols<IN_KT, NH>();
    constexpr uint32_t NS = IN_PAD + (NH == 1 ? 16 : 64);       // N of the stacked weight-gradient product
    constexpr uint32_t T_S = 64, T_WO = 64 + NS;                 // TMEM columns: stacked dW, dWo^T (NH == 2)
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_ready, bar_done, bar_wdone;
    __shared__ uint32_t tmem_slot;
    __shared__ long long s_prof[16];           // per-phase cycle counters of thread 0 (tools/mlp_bench.py), off by default
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    stage_weight(smem + L.W1, CHW, params, 64, IN_PAD);
    if constexpr (NH == 2) stage_weight(smem + L.Wh, CHW, params + 64 * IN_PAD, 64, 64);
    stage_weight(smem + L.Wo, CHO, params + 64 * IN_PAD + (NH - 1) * 64 * 64, 16, 64);
    if (warp == 0) { tc05::tmem_alloc(&tmem_slot, TCOLS); tc05::tmem_relinquish(); }
    if (tid == 0) {
        tc05::mbar_init(&bar_ready, TC_ROWS); tc05::mbar_init(&bar_done, 1); tc05::mbar_init(&bar_wdone, 1);
        tc05::fence_mbar_init();
    }
    publish_and_sync();
    tc05::fence_after_sync();
    const uint32_t tacc = tmem_slot;
    const uint32_t ntiles = (B + 127) / 128;
    const bool want_dw = dparams != nullptr;
    const bool linear_out = out_act == NRF_ACT_NONE;       // dZ = loss_scale * dy: no need to recompute Z
    const bool worked = blockIdx.x < ntiles;

    if (warp == 4) {
        // ================================================================ MMA issuer (one lane)
        if (lane == 0) {
            const uint32_t aX = tc05::smem_u32(smem + L.X), aH1 = tc05::smem_u32(smem + L.H1), aH2 = tc05::smem_u32(smem + L.H2);
            const uint32_t adZ = tc05::smem_u32(smem + L.dZ), adH1 = tc05::smem_u32(smem + L.dH1), adH2 = tc05::smem_u32(smem + L.dH2);
            const uint32_t aW1 = tc05::smem_u32(smem + L.W1), aWh = tc05::smem_u32(smem + L.Wh), aWo = tc05::smem_u32(smem + L.Wo);
            const uint32_t aHlast = (NH == 2) ? aH2 : aH1;
            // k = K-major view, m = MN-major view of a chunked tile
            const uint64_t kX = tc05::desc_kmajor(aX, CH), mX = tc05::desc_mnmajor(aX, CH);
            const uint64_t kH1 = tc05::desc_kmajor(aH1, CH);
            const uint64_t kHlast = tc05::desc_kmajor(aHlast, CH), mHlast = tc05::desc_mnmajor(aHlast, CH);
            const uint64_t kdZ = tc05::desc_kmajor(adZ, CH), mdZ = tc05::desc_mnmajor(adZ, CH);
            const uint64_t kdH1 = tc05::desc_kmajor(adH1, CH), mdH1 = tc05::desc_mnmajor(adH1, CH);
            const uint64_t kdH2 = tc05::desc_kmajor(adH2, CH);
            const uint64_t kW1 = tc05::desc_kmajor(aW1, CHW), mW1 = tc05::desc_mnmajor(aW1, CHW);
            const uint64_t kWh = tc05::desc_kmajor(aWh, CHW), mWh = tc05::desc_mnmajor(aWh, CHW);
            const uint64_t kWo = tc05::desc_kmajor(aWo, CHO), mWo = tc05::desc_mnmajor(aWo, CHO);
            constexpr uint32_t ID_H = tc05::idesc_f16(128, 64, false, false);        // X W1^T, H1 Wh^T      (A K-major, B K-major)
            constexpr uint32_t ID_O = tc05::idesc_f16(128, 16, false, false);        // Hlast Wo^T
            constexpr uint32_t ID_DH = tc05::idesc_f16(128, 64, false, true);        // dZ Wo, dH2 Wh        (B MN-major)
            constexpr uint32_t ID_DX = tc05::idesc_f16(128, IN_PAD, false, true);    // dH1 W1
            constexpr uint32_t ID_GS = tc05::idesc_f16(128, NS, true, true);         // stacked weight gradient
            constexpr uint32_t ID_GWO = tc05::idesc_f16(64, 16, true, true);         // H2^T dZ (= dWo^T), NH == 2
            uint32_t ph = 0;
            bool first = true;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                // ---- H1 = act(X W1^T)
                tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
#pragma unroll
                for (int k = 0; k < IN_KT; k++) tc05::mma_f16(tacc, dadv(kX, k * 2 * CH), dadv(kW1, k * 2 * CHW), ID_H, k > 0);
                tc05::mma_commit(&bar_done);
                if constexpr (NH == 2) {
                    tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
#pragma unroll
                    for (int k = 0; k < 4; k++) tc05::mma_f16(tacc, dadv(kH1, k * 2 * CH), dadv(kWh, k * 2 * CHW), ID_H, k > 0);
                    tc05::mma_commit(&bar_done);
                }
                if (!linear_out) {
                    // ---- Z = Hlast Wo^T
                    tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
#pragma unroll
                    for (int k = 0; k < 4; k++) tc05::mma_f16(tacc, dadv(kHlast, k * 2 * CH), dadv(kWo, k * 2 * CHO), ID_O, k > 0);
                    tc05::mma_commit(&bar_done);
                }
                // ---- dHlast_pre = dZ Wo   (| dWo^T += H2^T dZ when NH == 2)
                tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
                tc05::mma_f16(tacc, kdZ, mWo, ID_DH, 0);
                tc05::mma_commit(&bar_done);
                if constexpr (NH == 2) {
                    if (want_dw) {
#pragma unroll
                        for (int k = 0; k < 8; k++) tc05::mma_f16(tacc + T_WO, dadv(mHlast, k * 256), dadv(mdZ, k * 256), ID_GWO, (!first || k > 0) ? 1u : 0u);
                    }
                    // ---- dH1_pre = dH2 Wh
                    tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
#pragma unroll
                    for (int k = 0; k < 4; k++) tc05::mma_f16(tacc, dadv(kdH2, k * 2 * CH), dadv(mWh, k * 256), ID_DH, k > 0);
                    tc05::mma_commit(&bar_done);
                }
                // ---- dX = dH1 W1   |   stacked weight gradient  [dH1 | Hlast or dH2]^T [X | dZ or H1]
                tc05::mbar_wait(&bar_ready, ph); ph ^= 1; tc05::fence_after_sync();
#pragma unroll
                for (int k = 0; k < 4; k++) tc05::mma_f16(tacc, dadv(kdH1, k * 2 * CH), dadv(mW1, k * 256), ID_DX, k > 0);
                tc05::mma_commit(&bar_done);
                if (want_dw) {
#pragma unroll
                    for (int k = 0; k < 8; k++) tc05::mma_f16(tacc + T_S, dadv(mdH1, k * 256), dadv(mX, k * 256), ID_GS, (!first || k > 0) ? 1u : 0u);
                }
                tc05::mma_commit(&bar_wdone);
                first = false;
            }
        }
        __syncwarp();
    } else {
        // ================================================================ row owners
        long long plast = 0;
        if (prof && tid == 0) { for (int i = 0; i < 16; i++) s_prof[i] = 0; plast = clock64(); }
#define PROF_MARK(i) do { if (prof && tid == 0) { const long long t_ = clock64(); s_prof[i] += t_ - plast; plast = t_; } } while (0)
        const uint32_t tacc_lane = tacc + ((uint32_t)(warp * 32) << 16);
        const bool relu = hidden_act == NRF_ACT_RELU;
        const bool x_vec = vec_ok_for(x, n_in, n_in);
        const bool dx_vec = vec_ok_for(dx, n_in, n_in);
        const bool dy_vec = vec_ok_for(dy, n_out, ld_dy);
        const float inv_scale = 1.0f / loss_scale;
        // plain (non-accumulating) f16 dx whose rows are whole 32-byte sectors
        const bool dx_wide_ok = dx_vec && x_dt == NRF_DTYPE_F16 && (n_in % 16 == 0) && ((reinterpret_cast<uintptr_t>(dx) & 31) == 0);
        const bool dx_wide = dx_wide_ok && !dx_accumulate, dx_wide_acc = dx_wide_ok && dx_accumulate;
        uint32_t phase = 0, phase_w = 0;
        bool first = true;
        const bool coal = x_vec && x_dt == NRF_DTYPE_F16 && n_in == (uint32_t)IN_PAD && (reinterpret_cast<uintptr_t>(x) & 31) == 0;
        uint8_t* const h1row = smem + L.H1 + tid * 16;
        uint8_t* const h2row = smem + L.H2 + tid * 16;
        uint8_t* const dzrow = smem + L.dZ + tid * 16;
        uint8_t* const dh1row = smem + L.dH1 + tid * 16;
        uint8_t* const dh2row = smem + L.dH2 + tid * 16;
        uint8_t* const hlastrow = (NH == 2) ? h2row : h1row;
        uint8_t* const dhlastrow = (NH == 2) ? dh2row : dh1row;

        // software pipeline: rows of tile i+1 are fetched from HBM while tile i is processed
        uint4 xr[CPR];
        uint32_t dyraw[16];
        {
            const size_t row0 = (size_t)blockIdx.x * 128 + tid;
            load_x_tile<CPR, true>(xr, x, x_dt, (size_t)blockIdx.x * 128, true, B, n_in, x_vec, coal, warp, lane, tid);
            dy_load_raw(dy, dy_dt, row0, n_out, ld_dy, row0 < B, dy_vec, dyraw);
        }
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const size_t row = (size_t)tile * 128 + tid;
            const bool row_ok = row < B;
            uint32_t dyc[8];
            dy_convert(dyraw, dy_dt, dy_vec, loss_scale, dyc);
            PROF_MARK(0);
            if (!first) { tc05::mbar_wait(&bar_wdone, phase_w); phase_w ^= 1; }      // operand tiles are free again
            PROF_MARK(1);      // previous tile's weight-gradient MMAs done
            stage_x_tile<CPR, true>(xr, smem + L.X, CH, coal, warp, lane, tid);
            PROF_MARK(13);     // x rows staged (first use of the prefetched registers: exposes any load latency left)
            if (linear_out) {
                *reinterpret_cast<uint4*>(dzrow) = make_uint4(dyc[0], dyc[1], dyc[2], dyc[3]);
                *reinterpret_cast<uint4*>(dzrow + CH) = make_uint4(dyc[4], dyc[5], dyc[6], dyc[7]);
            }
            PROF_MARK(14);     // dZ rows staged
            publish(&bar_ready);
            PROF_MARK(15);     // proxy fence + arrive
            {   // prefetch the next tile's rows
                const uint32_t nt = tile + gridDim.x;
                const size_t nrow = (size_t)nt * 128 + tid;
                const bool nok = nt < ntiles && nrow < B;
                load_x_tile<CPR, true>(xr, x, x_dt, (size_t)nt * 128, nt < ntiles, B, n_in, x_vec, coal, warp, lane, tid);
                dy_load_raw(dy, dy_dt, nrow, n_out, ld_dy, nok, dy_vec, dyraw);
            }
            PROF_MARK(2);      // operands staged, prefetch issued
            tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
            PROF_MARK(3);      // H1 MMA round trip
            hidden_fwd_epilogue(tacc_lane, h1row, relu, CH);
            publish(&bar_ready);
            PROF_MARK(4);      // H1 epilogue
            if constexpr (NH == 2) {
                tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
                hidden_fwd_epilogue(tacc_lane, h2row, relu, CH);
                publish(&bar_ready);
            }
            PROF_MARK(5);      // H2 round trip + epilogue
            if (!linear_out) {
                tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
                PROF_MARK(6);  // Z MMA round trip
                uint32_t z[16];
                tc05::tmem_ld16(tacc_lane, z);
                tc05::tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    uint32_t p[4] = {0u, 0u, 0u, 0u};
                    if (8u * c < n_out) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const float2 d = __half22float2(bits2h(dyc[4 * c + q]));
                            p[q] = tpack(d.x * tact_bwd(__uint_as_float(z[8 * c + 2 * q]), out_act),
                                         d.y * tact_bwd(__uint_as_float(z[8 * c + 2 * q + 1]), out_act));
                        }
                    }
                    *reinterpret_cast<uint4*>(dzrow + c * CH) = make_uint4(p[0], p[1], p[2], p[3]);
                }
                publish(&bar_ready);
                PROF_MARK(7);  // dZ epilogue
            }
            tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
            PROF_MARK(8);      // dHlast MMA round trip
            hidden_bwd_epilogue(tacc_lane, hlastrow, dhlastrow, relu, CH);
            publish(&bar_ready);
            PROF_MARK(9);      // dHlast epilogue
            if constexpr (NH == 2) {
                tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
                hidden_bwd_epilogue(tacc_lane, h1row, dh1row, relu, CH);
                publish(&bar_ready);
            }
            tc05::mbar_wait(&bar_done, phase); phase ^= 1; tc05::fence_after_sync();
            PROF_MARK(10);     // (dH1 round trip + epilogue,) dX MMA round trip
            if (dx != nullptr) {
#pragma unroll
                for (int g = 0; g < IN_PAD / 16; g++) {
                    uint32_t v[16];
                    tc05::tmem_ld16(tacc_lane + 16 * g, v);
                    tc05::tmem_ld_wait();
                    if (row_ok && dx_wide_acc) {
                        // accumulate into an existing f16 dx: the row belongs to this thread alone and the producer of the
                        // existing values is an earlier kernel on the stream, so a 256-bit load + packed add + 256-bit store
                        // replaces four 16-byte reductions
                        __half* dst = reinterpret_cast<__half*>(dx) + row * n_in + 16 * g;
                        uint32_t old[8];
                        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=r"(old[0]), "=r"(old[1]), "=r"(old[2]), "=r"(old[3]), "=r"(old[4]), "=r"(old[5]), "=r"(old[6]), "=r"(old[7])
                                     : "l"(dst) : "memory");
                        uint32_t pk[8];
#pragma unroll
                        for (int q = 0; q < 8; q++)
                            pk[q] = h2bits(__hadd2(bits2h(old[q]), bits2h(tpack(__uint_as_float(v[2 * q]) * inv_scale, __uint_as_float(v[2 * q + 1]) * inv_scale))));
                        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                                     "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                    } else if (row_ok && dx_wide) {
                        // 16 f16 gradients = one 32-byte sector: a single 256-bit store (STG.E.256, sm_100) per lane instead of
                        // two half-sector 16-byte stores
                        uint32_t pk[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) pk[q] = tpack(__uint_as_float(v[2 * q]) * inv_scale, __uint_as_float(v[2 * q + 1]) * inv_scale);
                        __half* dst = reinterpret_cast<__half*>(dx) + row * n_in + 16 * g;
                        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                                     "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                    } else if (row_ok) {
#pragma unroll
                        for (int c = 0; c < 2; c++) {
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; j++) f[j] = __uint_as_float(v[8 * c + j]) * inv_scale;
                            if (dx_accumulate) red_chunk(dx, x_dt, row, 2 * g + c, n_in, n_in, f, dx_vec);
                            else store_chunk(dx, x_dt, row, 2 * g + c, n_in, n_in, f, dx_vec);
                        }
                    }
                }
            }
            first = false;
            PROF_MARK(11);     // dX epilogue
            // the next tile's publish() orders these TMEM reads before the issuer overwrites the accumulator
        }
        // ---- flush the weight gradients accumulated in TMEM
        if (worked) { tc05::mbar_wait(&bar_wdone, phase_w); tc05::fence_after_sync(); }
        if (want_dw && worked) {
            float* dW1 = dparams;
            float* dWh = dparams + 64 * IN_PAD;
            float* dWo = dWh + (NH - 1) * 64 * 64;
            // stacked product (M=128 layout: row m in lane m): rows 0-63 x cols [0, in) = dW1; rows 64-127 x cols [in, NS) =
            // dWo^T (NH == 1) or dWh (NH == 2).  tcgen05.ld is warp-collective: the branch below is warp-uniform.
            if (tid < 64) {
#pragma unroll 1
                for (int g = 0; g < IN_PAD / 16; g++) {
                    uint32_t v[16];
                    tc05::tmem_ld16(tacc_lane + T_S + 16 * g, v);
                    tc05::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j++) atomicAdd(dW1 + tid * IN_PAD + 16 * g + j, __uint_as_float(v[j]) * inv_scale);
                }
            } else if constexpr (NH == 1) {
                uint32_t v[16];
                tc05::tmem_ld16(tacc_lane + T_S + IN_PAD, v);
                tc05::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j++) atomicAdd(dWo + j * 64 + (tid - 64), __uint_as_float(v[j]) * inv_scale);
            } else {
#pragma unroll 1
                for (int g = 0; g < 4; g++) {
                    uint32_t v[16];
                    tc05::tmem_ld16(tacc_lane + T_S + IN_PAD + 16 * g, v);
                    tc05::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j++) atomicAdd(dWh + (tid - 64) * 64 + 16 * g + j, __uint_as_float(v[j]) * inv_scale);
                }
            }
            if constexpr (NH == 2) {
                // dWo^T (M=64 layout: row m in lane (m % 16) + 32 * (m / 16)); every lane loads, lanes 0-15 own rows
                uint32_t v[16];
                tc05::tmem_ld16(tacc_lane + T_WO, v);
                tc05::tmem_ld_wait();
                if (lane < 16) {
                    const int m = warp * 16 + lane;
#pragma unroll
                    for (int j = 0; j < 16; j++) atomicAdd(dWo + j * 64 + m, __uint_as_float(v[j]) * inv_scale);
                }
            }
        }
        PROF_MARK(12);         // dW flush
        if (prof && tid == 0 && blockIdx.x == 0) { for (int i = 0; i < 16; i++) prof[i] = (unsigned long long)s_prof[i]; }
#undef PROF_MARK
    }
    publish_and_sync();
    if (warp == 0) tc05::tmem_dealloc(tacc, TCOLS);
}

int tc_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

int g_fwd_ctas_per_sm = 6, g_bwd_ctas_per_sm = 4;
unsigned long long* g_prof = nullptr;

template <int IN_KT, int NH>
int launch_fwd(const void* x, int xdt, const void* params, uint32_t B, uint32_t n_in, uint32_t n_out, int hact, int oact, void* y, int ydt,
               uint32_t ld_y, const int32_t* B_dev, cudaStream_t s) {
    constexpr TcSmem L = fwd_smem<IN_KT, NH>();
    auto kern = k_mlp_fwd_tc<IN_KT, NH>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    const uint32_t ntiles = ceil_div_u32(B, 128);
    // NOTE per_sm must not exceed what is really resident (registers included: see the launch bounds), or the surplus CTAs
    // run as a tail wave.  (cudaOccupancyMaxActiveBlocksPerMultiprocessor under-reports here -- it ignores the opt-in
    // shared-memory carve-out -- and halved the grid when it was used as a clamp.)
    const uint32_t per_sm = (uint32_t)max(1, min(min(g_fwd_ctas_per_sm, (int)(220 * 1024 / (L.total + 1024))), 8));
    const uint32_t grid = (uint32_t)min((uint64_t)ntiles, (uint64_t)tc_sm_count() * per_sm);
    kern<<<grid, TC_THREADS, L.total, s>>>(x, xdt, (const __half*)params, B, n_in, n_out, hact, oact, y, ydt, ld_y, B_dev);
    return nrf_check_launch();
}

template <int IN_KT, int NH>
int launch_bwd(const void* x, int xdt, const void* params, const void* dy, int dydt, uint32_t ld_dy, uint32_t B, uint32_t n_in, uint32_t n_out,
               int hact, int oact, float ls, void* dx, int dx_acc, float* dparams, cudaStream_t s) {
    constexpr TcSmem L = bwd_smem<IN_KT, NH>();
    constexpr uint32_t TCOLS = bwd_tmem_cols<IN_KT, NH>();
    auto kern = k_mlp_bwd_tc<IN_KT, NH>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    const uint32_t ntiles = ceil_div_u32(B, 128);
    const uint32_t per_sm = (uint32_t)max(1, min(min(g_bwd_ctas_per_sm, (int)(220 * 1024 / (L.total + 1024))), (int)(512 / TCOLS)));
    const uint32_t grid = (uint32_t)min((uint64_t)ntiles, (uint64_t)tc_sm_count() * per_sm);
    kern<<<grid, TC_THREADS, L.total, s>>>(x, xdt, (const __half*)params, dy, dydt, ld_dy, B, n_in, n_out, hact, oact, ls, dx, dx_acc, dparams, g_prof);
    return nrf_check_launch();
}

}  // namespace

// entry points used by mlp.cu's dispatcher (same argument meaning as nrf_mlp_forward / nrf_mlp_backward)
int nrf_mlp_tc_forward(const void* x, int x_dtype, const void* params_f16, uint32_t B, uint32_t n_in, uint32_t n_out, uint32_t n_hidden,
                       int hidden_act, int out_act, void* y, int y_dtype, uint32_t ld_y, const int32_t* B_dev, cudaStream_t s) {
    const int kt = (int)((n_in + 15) / 16);
#define TC_FWD(K, H) if (kt == K && (int)n_hidden == H) return launch_fwd<K, H>(x, x_dtype, params_f16, B, n_in, n_out, hidden_act, out_act, y, y_dtype, ld_y, B_dev, s)
    TC_FWD(1, 1); TC_FWD(2, 1); TC_FWD(3, 1); TC_FWD(4, 1);
    TC_FWD(1, 2); TC_FWD(2, 2); TC_FWD(3, 2); TC_FWD(4, 2);
#undef TC_FWD
    return NRF_E_UNSUPPORTED;
}

int nrf_mlp_tc_backward(const void* x, int x_dtype, const void* params_f16, const void* dy, int dy_dtype, uint32_t ld_dy, uint32_t B,
                        uint32_t n_in, uint32_t n_out, uint32_t n_hidden, int hidden_act, int out_act, float loss_scale, void* dx,
                        int dx_accumulate, float* dparams, cudaStream_t s) {
    const int kt = (int)((n_in + 15) / 16);
#define TC_BWD(K, H) if (kt == K && (int)n_hidden == H) return launch_bwd<K, H>(x, x_dtype, params_f16, dy, dy_dtype, ld_dy, B, n_in, n_out, hidden_act, out_act, loss_scale, dx, dx_accumulate, dparams, s)
    TC_BWD(1, 1); TC_BWD(2, 1); TC_BWD(3, 1); TC_BWD(4, 1);
    TC_BWD(1, 2); TC_BWD(2, 2); TC_BWD(3, 2); TC_BWD(4, 2);
#undef TC_BWD
    return NRF_E_UNSUPPORTED;
}

void nrf_mlp_tc_set_prof(unsigned long long* buf16) { g_prof = buf16; }

void nrf_mlp_tc_set_ctas(int fwd_per_sm, int bwd_per_sm) {
    if (fwd_per_sm > 0) g_fwd_ctas_per_sm = fwd_per_sm;
    if (bwd_per_sm > 0) g_bwd_ctas_per_sm = bwd_per_sm;
}
