/*
 * nerfstyle_b200.h -- C ABI of libnerfstyle_b200.so (hand-written sm_100a CUDA).
 *
 * Drop-in boundary for the NeRF render/train hot path of hkust-vgd/nerfstyle.  Each entry point
 * replaces one pybind11 function of the reference's two extensions (and the tiny-cuda-nn calls), with
 * plain pointers + sizes instead of at::Tensor:
 *
 *   reference interface                                   replaced by
 *   ----------------------------------------------------  -------------------------------------------
 *   raymarching/src/raymarching.h:6   near_far_from_aabb   nrf_near_far_from_aabb
 *   raymarching/src/raymarching.h:7   sph_from_ray         nrf_sph_from_ray
 *   raymarching/src/raymarching.h:8   morton3D             nrf_morton3D
 *   raymarching/src/raymarching.h:9   morton3D_invert      nrf_morton3D_invert
 *   raymarching/src/raymarching.h:10  packbits             nrf_packbits
 *   raymarching/src/raymarching.h:12  march_rays_train     nrf_march_rays_train (+ _count/_write split)
 *   raymarching/src/raymarching.h:14  composite_rays_train_forward / :15 _backward
 *                                                          nrf_composite_rays_train_forward / _backward
 *   raymarching/src/raymarching.h:16  march_rays           nrf_march_rays
 *   raymarching/src/raymarching.h:17  composite_rays       nrf_composite_rays (+ nrf_compact_alive)
 *   gridencoder/src/gridencoder.h:12  grid_encode_forward  nrf_grid_encode_forward
 *   gridencoder/src/gridencoder.h:13  grid_encode_backward nrf_grid_encode_backward
 *   gridencoder/src/gridencoder.h:14  grid_initialize      nrf_grid_initialize
 *   tinycudann Network fwd/bwd (networks/style_nerf.py:44-98)   nrf_mlp_forward / nrf_mlp_backward
 *   loss.py:32-36,199-214 cosine_dists + mask + amin       nrf_nnfm_forward (backward = gather, nnfm.py)
 *   nerf_lib.py:69-142 NerfLib.generate_rays (+ RayBatch)  nrf_generate_rays
 *   renderer.py:139-194 Renderer.update_state              nrf_occ_* + nrf_packbits_dev (around the density query)
 *
 * Conventions: all pointers are DEVICE pointers owned by the caller; the callee never
 * allocates, never synchronises and never throws.  Every function launches on `stream` (a
 * cudaStream_t passed as void*) of the CURRENT device and returns 0 on success or a negative NRF_E_*
 * code; nrf_error_string() describes it.  Layouts and dtypes are the reference's unless stated.
 */
#ifndef NERFSTYLE_B200_H
#define NERFSTYLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRF_OK             0
#define NRF_E_INVALID     -1   /* bad argument (null pointer, unsupported D / C / width ...) */
#define NRF_E_UNSUPPORTED -2   /* combination the kernels do not implement */
#define NRF_E_CUDA        -3   /* a CUDA runtime call or launch failed; see nrf_last_cuda_error() */

#define NRF_DTYPE_F32 0
#define NRF_DTYPE_F16 1

/* activations of the tcnn-style MLP */
#define NRF_ACT_NONE     0
#define NRF_ACT_RELU     1
#define NRF_ACT_SIGMOID  2
#define NRF_ACT_EXP      3
#define NRF_ACT_TRUNC_EXP 4   /* y = exp(f16(z)) in f32; dy/dz = exp(clamp(f16(z), -15, 15)) (networks/tcnn_nerf.py:55-69) */

const char* nrf_error_string(int code);
int  nrf_last_cuda_error(void);          /* cudaError_t of the last NRF_E_CUDA on this thread */
int  nrf_version(void);                  /* ABI version, bumped on any signature change */
int  nrf_device_info(int* sm_count, int* cc_major, int* cc_minor, int* l2_bytes);

/* ------------------------------------------------------------------ ray-marching operator set */

/* raymarching.cu:191-244.  rays_o/rays_d [N,3] f32, aabb [6] f32 -> nears/fars [N] f32. */
int nrf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                           float min_near, float* nears, float* fars, void* stream);

/* raymarching.cu:262-297.  coords [N,2] f32. */
int nrf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                     void* stream);

/* raymarching.cu:313-359.  coords [N,3] i32 <-> indices [N] i32. */
int nrf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream);
int nrf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream);

/* raymarching.cu:367-388.  grid [8N] f32 -> bitfield [N] u8; bit i of byte n = grid[8n+i] > thresh. */
int nrf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, void* stream);

/* Size in bytes of the scratch buffer nrf_march_rays_train* need for N rays. */
uint64_t nrf_march_scratch_bytes(uint32_t N);

/* raymarching.cu:411-589, first pass.  Counts the samples of every ray, then scans: on return (stream
 * order) rays[n] = (n, offset, count) with offset = counter[0]_at_entry + exclusive scan of the counts
 * in RAY ORDER (one valid schedule of the reference's racing atomicAdd, deterministic), and
 * counter[0] += total samples, counter[1] += N (like the reference's two atomics, :506-507).
 * noises may be NULL (= all zero, raymarching.py:247).  scratch: nrf_march_scratch_bytes(N). */
int nrf_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                               float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                               const float* nears, const float* fars, const float* noises,
                               int32_t* rays, int32_t* counter, void* scratch, void* stream);

/* raymarching.cu:411-589, second pass.  Re-marches every ray and writes its samples at rays[n].offset:
 * xyzs [rows,3], dirs [rows,3], deltas [rows,4] f32.  A ray is skipped when count == 0 or
 * offset + count >= M (the reference's drop rule, :516-517); M is the reference's logical capacity
 * (N*max_steps or the aligned mean_count), `rows` (<= M) is what the caller really allocated.  Rows in
 * [zero_from, rows) are zero-filled by the kernel (the reference zero-fills the whole buffer on the
 * host side, raymarching.py:238-240); pass zero_from = rows to skip. */
int nrf_march_rays_train_write(const float* rays_o, const float* rays_d, const float* z_hats,
                               const uint8_t* grid, float bound, float dt_gamma, uint32_t max_steps, int is_ndc,
                               uint32_t N, uint32_t C, uint32_t H, uint32_t M, uint32_t rows, uint32_t zero_from,
                               const float* nears, const float* fars, const float* noises, const int32_t* rays,
                               float* xyzs, float* dirs, float* deltas, void* stream);

/* One-call form with the reference's exact native signature (pre-allocated, zero-filled M rows). */
int nrf_march_rays_train(const float* rays_o, const float* rays_d, const float* z_hats, const uint8_t* grid,
                         float bound, float dt_gamma, uint32_t max_steps, int is_ndc, uint32_t N, uint32_t C,
                         uint32_t H, uint32_t M, const float* nears, const float* fars, float* xyzs, float* dirs,
                         float* deltas, int32_t* rays, int32_t* counter, const float* noises, void* scratch,
                         void* stream);

/* Staged marching (no second walk of the occupancy grid): the count pass records the marching time of every emitted sample
 * in t_stage [N, max_steps] f32; after the host has read the total (the contract's D2H read) and allocated the outputs,
 * nrf_march_rays_train_emit produces xyzs / dirs / deltas from those times -- bit-identical to nrf_march_rays_train_write.
 * Warp-per-ray walker, non-NDC only (NRF_E_UNSUPPORTED otherwise: use _count + _write). */
int nrf_march_rays_train_count_staged(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                      float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                      const float* nears, const float* fars, const float* noises, int32_t* rays,
                                      int32_t* counter, void* scratch, float* t_stage, void* stream);
int nrf_march_rays_train_emit(const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                              uint32_t C, uint32_t H, uint32_t M, uint32_t rows, uint32_t zero_from, const float* nears,
                              const float* noises, const int32_t* rays, const float* t_stage, float* xyzs, float* dirs,
                              float* deltas, void* stream);

/* raymarching.cu:807-879.  sigmas [M], rgbs [M,C], deltas [M,4], rays [N,3] ->
 * weights_sum [N], depth [N], image [N,C] (all f32). */
int nrf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                     const int32_t* rays, uint32_t M, uint32_t N, uint32_t C, float T_thresh,
                                     int is_ndc, float* weights_sum, float* depth, float* image, void* stream);

/* raymarching.cu:905-986.  grad_sigmas [M] / grad_rgbs [M,C] must be zero-filled by the caller
 * (raymarching.py:339-340); the reference's rgbs_buf scratch is not needed. */
int nrf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                      const float* rgbs, const float* deltas, const int32_t* rays, int is_ndc,
                                      const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                      uint32_t C, float T_thresh, float* grad_sigmas, float* grad_rgbs,
                                      void* stream);
/* write_zeros != 0: the gradients need no zero fill by the caller -- every sample slot that belongs to a ray is written
 * (gradient or zero); padding rows that belong to no ray stay the caller's.  C <= 32. */
int nrf_composite_rays_train_backward_ex(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                      const float* rgbs, const float* deltas, const int32_t* rays, int is_ndc,
                                      const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                      uint32_t C, float T_thresh, float* grad_sigmas, float* grad_rgbs,
                                      int write_zeros, void* stream);

/* raymarching.cu:1005-1120.  xyzs/dirs [Mpad,3], deltas [Mpad,4]; rows [n_alive*n_step, Mpad) and the
 * unused slots of exited rays are zero-filled by the kernel when zero_fill != 0 (the reference relies on a
 * host-side torch.zeros, raymarching.py:409-412). */
int nrf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                   const float* rays_o, const float* rays_d, const float* z_hats, float bound, float dt_gamma,
                   uint32_t max_steps, int is_ndc, uint32_t C, uint32_t H, const uint8_t* grid, const float* nears,
                   const float* fars, float* xyzs, float* dirs, float* deltas, const float* noises,
                   uint32_t Mpad, int zero_fill, void* stream);

/* raymarching.cu:1134-1231.  In place on rays_alive, rays_t, weights_sum, depth, image. */
int nrf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                       const float* sigmas, const float* rgbs, const float* deltas, uint32_t C, int is_ndc,
                       float* weights_sum, float* depth, float* image, void* stream);

/* Extension (replaces `rays_alive[rays_alive >= 0]`, renderer.py:284): stable compaction of the
 * non-negative entries of in[0..n) into out; *n_out (device int32) receives the new count.
 * scratch: nrf_march_scratch_bytes(n). */
int nrf_compact_alive(const int32_t* in, uint32_t n, int32_t* out, int32_t* n_out, void* scratch, void* stream);

/* ------------------------------------------------------------------ hash-grid encoder */

/* gridencoder.cu:439-462 (kernel_grid :83-235).  inputs [B,D] f32 in [0,1]; embeddings [rows,C] f32 or
 * f16 (dtype); offsets [L+1] i32.  outputs: point_major != 0 -> [B, L*C] (what grid.py:58 produces after
 * its permute), else the reference's native [L,B,C].  dy_dx [B, L*D*C] only when calc_grad_inputs.
 * Supported: D in {2,3}, C in {1,2,4,8}. */
int nrf_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs,
                            uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                            int calc_grad_inputs, void* dy_dx, uint32_t gridtype, int align_corners, uint32_t style,
                            int dtype, int point_major, void* stream);

/* gridencoder.cu:464-494 (kernel_grid_backward :238-328, kernel_input_backward :331-357).
 * grad: [B, L*C] when point_major else [L,B,C] (dtype); grad_embeddings [rows,C] (grad_table_dtype) must be
 * zero-filled by the caller (grid.py:82).  grad_table_dtype = dtype reproduces the reference (f16 grads are
 * accumulated with __half2 atomics, :313-319); f16 grads may also be accumulated into an f32 table
 * (grad_table_dtype = NRF_DTYPE_F32), which avoids the swamping of half-precision accumulation and the
 * half->float cast the autograd engine would add.  grad_inputs [B,D] (dtype) only when calc_grad_inputs.
 * Row order never changes the result, only the speed: with f32 gradient tables, 16 levels and point-major rows the
 * scatter (here and in the dual / paired forms below) walks chunks of consecutive rows and sums the rows that stay in one
 * cell in registers, so samples should arrive ray by ray, as march_rays_train emits them. */
int nrf_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings, const int32_t* offsets,
                             void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                             uint32_t H, int calc_grad_inputs, const void* dy_dx, void* grad_inputs,
                             uint32_t gridtype, int align_corners, uint32_t style, int dtype, int grad_table_dtype,
                             int point_major, void* stream);

/* Two encoders with IDENTICAL geometry (same offsets / S / H / gridtype / align_corners) on the same points in one
 * pass: cell positions, hash rows and trilinear weights are computed once (the model's x_density_embedder and
 * x_color_embedder, networks/style_nerf.py:29-30,121-134, are such a pair).  D=3, C=2, point-major [B, L*2] outputs /
 * gradients, no input gradients.  Results are those of two nrf_grid_encode_forward / _backward calls.
 * xform (device float[7] = {min[3], size[3], bound}, or NULL): the points are first mapped by ((x - min) / size + bound) /
 * (2 bound) in f32, operation for operation what common.py:288 and grid.py:174 do in four elementwise kernels. 
 * A table that takes no gradient passes NULL for both its grad and its grad_embeddings (its share of the scatter is skipped). */
int nrf_grid_encode_forward_dual(const float* inputs, const void* embeddings0, const void* embeddings1,
                                 const int32_t* offsets, void* outputs0, void* outputs1, uint32_t B, uint32_t L, float S,
                                 uint32_t H, uint32_t gridtype, int align_corners, uint32_t style, int dtype,
                                 const float* xform, void* stream);
int nrf_grid_encode_backward_dual(const void* grad0, const void* grad1, const float* inputs, const int32_t* offsets,
                                  void* grad_embeddings0, void* grad_embeddings1, uint32_t B, uint32_t L, float S, uint32_t H,
                                  uint32_t gridtype, int align_corners, uint32_t style, int dtype, int grad_table_dtype,
                                  const float* xform, void* stream);

/* Paired forms of the dual entry points: the two tables of a same-geometry encoder pair live in ONE interleaved buffer
 * [row][encoder][2] (`table_pair`, in the tables' dtype; 16-byte aligned) and so do their f32 gradients (`grad_pair`), so a
 * corner of both encoders is one vector gather / one 16-byte reduction.  Values are those of the dual forms.
 * B_dev / row_deltas as in nrf_grid_encode_forward_dual_dev (NULL, NULL for the plain call). */
int nrf_grid_encode_forward_pair(const float* inputs, const void* table_pair, const int32_t* offsets, void* outputs0,
                                 void* outputs1, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                 int align_corners, uint32_t style, int dtype, const float* xform, const int32_t* B_dev,
                                 const float* row_deltas, void* stream);
int nrf_grid_encode_backward_pair(const void* grad0, const void* grad1, const float* inputs, const int32_t* offsets,
                                  float* grad_pair, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                  int align_corners, uint32_t style, int dtype, const float* xform, void* stream);

/* gridencoder.cu:551-571 (D=3, C=2, f32 like the reference's only use). */
int nrf_grid_initialize(const float* ref_embeddings, float* embeddings, const int32_t* ref_offsets,
                        const int32_t* offsets, uint32_t L, float S, uint32_t H, uint32_t Ns, void* stream);

/* ------------------------------------------------------------------ tcnn-style fully fused MLP */

/* Bias-free MLP  y = act_out(W_n relu(... relu(W_1 x))) with `width`-wide hidden layers (width = 64).
 * params: f16, tcnn FullyFusedMLP layout: W_1 [width, in_pad], (n_hidden-1) x [width, width],
 * W_out [out_pad, width], row-major, concatenated; in_pad / out_pad = n_in / n_out rounded up to 16.
 * x [B, n_in] (x_dtype f32 or f16, row stride n_in), y [B, n_out] (y_dtype).  Hidden activations are
 * rounded to f16 between layers, products accumulate in f32 on the tensor cores. */
int nrf_mlp_forward(const void* x, int x_dtype, const void* params_f16, uint32_t B, uint32_t n_in,
                    uint32_t n_out, uint32_t n_hidden, uint32_t width, int hidden_act, int out_act,
                    void* y, int y_dtype, void* stream);

/* Backward of the above; recomputes the hidden activations from x (nothing but x and y is saved).
 * dy [B, n_out] (dy_dtype); dx [B, n_in] (dx_dtype) or NULL; dparams f32 [same layout as params] is
 * ACCUMULATED into (caller zero-fills).  loss_scale multiplies dy before the f16 tensor-core products and
 * is divided out of dx / dparams (tcnn's loss_scale, 128 for f16). */
int nrf_mlp_backward(const void* x, int x_dtype, const void* params_f16, const void* dy, int dy_dtype,
                     uint32_t B, uint32_t n_in, uint32_t n_out, uint32_t n_hidden, uint32_t width,
                     int hidden_act, int out_act, float loss_scale, void* dx, int dx_dtype, float* dparams,
                     void* stream);

/* Extended forms used by the fused field heads of the host mirror (nerfstyle_b200/model.py): the output may be a
 * column block of a wider row-major matrix (ld_y / ld_dy = row stride in elements, 0 = n_out), which removes the
 * torch.cat of networks/style_nerf.py:139-141 and the dtype-cast / slice copies of its backward; dx_accumulate != 0
 * ADDS the input gradient into dx (vector reductions, no read-modify-write pass) so that two networks sharing an
 * input (class_net / color1_net, style_nerf.py:136-138) need no separate add.  An f32 y holds the f16 network output
 * widened (tcnn's output precision), except NRF_ACT_TRUNC_EXP which fuses networks/tcnn_nerf.py:55-69.
 * These extensions run on the tcgen05 implementation only. */
int nrf_mlp_forward_ex(const void* x, int x_dtype, const void* params_f16, uint32_t B, uint32_t n_in,
                       uint32_t n_out, uint32_t n_hidden, uint32_t width, int hidden_act, int out_act,
                       void* y, int y_dtype, uint32_t ld_y, void* stream);
int nrf_mlp_backward_ex(const void* x, int x_dtype, const void* params_f16, const void* dy, int dy_dtype,
                        uint32_t ld_dy, uint32_t B, uint32_t n_in, uint32_t n_out, uint32_t n_hidden,
                        uint32_t width, int hidden_act, int out_act, float loss_scale, void* dx, int dx_dtype,
                        int dx_accumulate, float* dparams, void* stream);

/* The four networks of the field (networks/style_nerf.py:120-142, use_dir=False: density 32->64->1 with trunc_exp,
 * class 32->64->K, color1 32->64->16, color2 16->64->64->3 with sigmoid) in ONE launch -- same arithmetic and rounding
 * points as four nrf_mlp_forward_ex calls, five MMA round trips per 128-point tile instead of nine (csrc/field_tc.cu).
 * enc_d / enc_c f16 [M, 32]; w_*: f16, tcnn layout of the respective network; sigmas f32 [M]; rgbs f32 [M, ld_rgbs]
 * (columns 0-2 rgb, 3..3+K class logits); c1_out f16 [M, 16] or NULL (color1's output, needed by the backward);
 * M_dev: device int32 row count (or NULL) for the device-driven inference loop. */
int nrf_field_forward(const void* enc_d, const void* enc_c, const void* w_density, const void* w_class, const void* w_color1,
                      const void* w_color2, uint32_t M, uint32_t n_classes, float* sigmas, float* rgbs, uint32_t ld_rgbs,
                      void* c1_out, const int32_t* M_dev, void* stream);

/* fp32 PARITY MODE of the same network (SURVEY.md 8c "fp32 (parity mode)"): fp32 weights (same tcnn layout), fp32
 * activations and fp32 FMA accumulation, nothing rounded to f16 and therefore no loss_scale.  x is f32 or f16 (dx follows
 * it), y / dy are f32; ld_y / ld_dy / dx_accumulate / NRF_ACT_TRUNC_EXP as in the _ex forms.  Not a performance path: it
 * exists so that gradients of the whole path can be checked against an fp32 CPU evaluation at rel 1e-4. */
int nrf_mlp_forward_f32(const void* x, int x_dtype, const float* params_f32, uint32_t B, uint32_t n_in, uint32_t n_out,
                        uint32_t n_hidden, uint32_t width, int hidden_act, int out_act, float* y, uint32_t ld_y,
                        void* stream);
int nrf_mlp_backward_f32(const void* x, int x_dtype, const float* params_f32, const float* dy, uint32_t ld_dy, uint32_t B,
                         uint32_t n_in, uint32_t n_out, uint32_t n_hidden, uint32_t width, int hidden_act, int out_act,
                         void* dx, int dx_accumulate, float* dparams, void* stream);

/* ------------------------------------------------------------------ device-driven inference loop (SURVEY 8f NEXT-2) */

/* The inference loop of renderer.py:237-293 with its per-iteration host logic moved to the device.  ctl = device int32[8]
 * {n_alive, n_step, n_rows = n_alive * n_step, steps_done, N, max_steps, iterations, -}; the caller initialises it to
 * {N, 1, N, 0, N, max_steps, 0, 0} and nrf_compact_alive_dev advances it at the end of every iteration (new alive count,
 * n_step = max(min(N / n_alive, 8), 1) as renderer.py:253, n_alive = 0 once max_steps is reached).  Every launch is sized
 * for the caps (n_alive_cap = N rays, B_cap >= N rows since n_alive * n_step <= N), so an iteration is shape-static and
 * can be captured in a CUDA graph and replayed with no host involvement; the host reads ctl[0] every few iterations to
 * stop.  ctl[4] is the row budget of an iteration (n_step = clamp(ctl[4] / n_alive, 1, 8); N = the reference's schedule);
 * ctl[7] = 1 selects STEP-MAJOR sample rows (sample s of alive slot n in row s * n_alive + n instead of n * n_step + s):
 * nrf_march_rays_dev then writes and nrf_composite_rays_dev reads consecutive rows from consecutive lanes.  Kernels and numerics are those of nrf_march_rays / nrf_composite_rays / nrf_compact_alive /
 * nrf_grid_encode_forward_dual / nrf_mlp_forward_ex; rows >= ctl[2] are left untouched.  row_deltas (the [B,4] deltas of
 * march_rays, or NULL) lets the encoder skip padding slots (delta == 0: composite_rays never reads them). */
/* (dirs may be NULL in nrf_march_rays_dev: a field without a direction input never reads them) */
int nrf_march_rays_dev(const int32_t* ctl, uint32_t n_alive_cap, const int32_t* rays_alive, const float* rays_t,
                       const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                       uint32_t H, const uint8_t* grid, const float* fars, float* xyzs, float* dirs, float* deltas,
                       void* stream);
int nrf_composite_rays_dev(const int32_t* ctl, uint32_t n_alive_cap, float T_thresh, int32_t* rays_alive, float* rays_t,
                           const float* sigmas, const float* rgbs, const float* deltas, uint32_t C, float* weights_sum,
                           float* depth, float* image, void* stream);
int nrf_compact_alive_dev(int32_t* ctl, uint32_t n_cap, const int32_t* in, int32_t* out, void* scratch, void* stream);
int nrf_grid_encode_forward_dual_dev(const float* inputs, const void* embeddings0, const void* embeddings1,
                                     const int32_t* offsets, void* outputs0, void* outputs1, uint32_t B_cap, uint32_t L,
                                     float S, uint32_t H, uint32_t gridtype, int align_corners, uint32_t style, int dtype,
                                     const float* xform, const int32_t* B_dev, const float* row_deltas, void* stream);
int nrf_mlp_forward_dev(const void* x, int x_dtype, const void* params_f16, uint32_t B_cap, uint32_t n_in, uint32_t n_out,
                        uint32_t n_hidden, uint32_t width, int hidden_act, int out_act, void* y, int y_dtype,
                        uint32_t ld_y, const int32_t* B_dev, void* stream);

/* tcnn.Encoding {'otype': 'SphericalHarmonics', 'degree': d} (networks/style_nerf.py:33-42, tcnn_nerf.py:87-95), d <= 4:
 * inputs01 [B,3] f32 in [0,1] (mapped to [-1,1] inside, as tiny-cuda-nn does) -> outputs [B, d*d] (f16 or f32). */
int nrf_sh_encode_forward(const float* inputs01, uint32_t B, uint32_t degree, void* outputs, int out_dtype, void* stream);

/* ------------------------------------------------------------------ nearest-neighbour feature matching */

/* loss.py:32-36,199-214.  a [N1,K] f16 (rows already L2-normalised), b [N2,K] f16 (normalised);
 * a_label [N1] i32 (class of each image feature, <0 = unconstrained), b_label [N2] i32 (cluster of each
 * style feature), match [n_class] i32 (class -> required cluster; NULL = no mask).
 * min_dist [N1] f32 = min_j (1 - a_i.b_j) over allowed j (+inf if none), argmin [N1] i32 (-1 if none).
 * The N1 x N2 matrix is never materialised. */
int nrf_nnfm_forward(const void* a_f16, const void* b_f16, uint32_t N1, uint32_t N2, uint32_t K,
                     const int32_t* a_label, const int32_t* b_label, const int32_t* match, uint32_t n_class,
                     float* min_dist, int32_t* argmin, void* scratch, void* stream);
uint64_t nrf_nnfm_scratch_bytes(uint32_t N1, uint32_t N2, uint32_t K);   /* scratch: 256-byte aligned */

/* ------------------------------------------------------------------ fused optimizer (SURVEY 8f NEXT-3) */

/* One pass per parameter tensor replacing GradScaler.unscale_/step + torch.optim.Adam(eps=1e-15) + LambdaLR +
 * torch_ema (trainers/base.py:216-229,420-426) and producing the fp16 copy the next autocast forward needs.
 * `state` is a device block of nrf_opt_state_bytes() bytes: {float scale; int found_inf; int growth_tracker;
 * int good_steps; ...}.  Per step: nrf_grads_check on every gradient, nrf_adam_step on every parameter
 * (skips itself when an inf/nan was found), nrf_scaler_update once.  No host synchronisation anywhere.
 * lr = lr0 * 0.1^(good_steps / lr_decay_steps) (LambdaLR of trainers/base.py:222-226; 0 disables the decay);
 * ema <- ema - ema_one_minus_decay * (ema - param) (torch_ema); ema / param_half may be NULL. */
uint64_t nrf_opt_state_bytes(void);
int nrf_grads_check(const float* grad, uint64_t n, void* state, void* stream);
int nrf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, void* param_half,
                  uint64_t n, const void* state, float lr0, float lr_decay_steps, float beta1, float beta2, float eps,
                  float ema_one_minus_decay, void* stream);
/* nrf_adam_step with strided gradient / fp16-copy rows: `grad` and `param_half` may point into the interleaved
 * [row][encoder][2] buffers of the paired hash-grid kernels (pre-offset to this tensor's encoder slot); the strides are in
 * 2-element rows (1 = contiguous, 2 = interleaved pair).  n must be even when a stride is > 1. */
int nrf_adam_step_ex(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, void* param_half,
                     uint64_t n, const void* state, float lr0, float lr_decay_steps, float beta1, float beta2, float eps,
                     float ema_one_minus_decay, uint32_t grad_row_stride, uint32_t half_row_stride, void* stream);
/* Both tables of an interleaved pair in one pass: grad_pair f32 [rows][2 tables][2] (16-byte aligned), half_pair f16 of the
 * same layout (8-byte aligned, may be NULL); param / exp_avg / exp_avg_sq / ema are the two tables' own contiguous
 * [rows, 2] f32 arrays (or this rank's shard of them: all pointers offset to the same first row). */
int nrf_adam_step_pair(float* param0, float* param1, const float* grad_pair, float* exp_avg0, float* exp_avg1,
                       float* exp_avg_sq0, float* exp_avg_sq1, float* ema0, float* ema1, void* half_pair, uint64_t rows,
                       const void* state, float lr0, float lr_decay_steps, float beta1, float beta2, float eps,
                       float ema_one_minus_decay, void* stream);
int nrf_scaler_update(void* state, float growth, float backoff, int growth_interval, void* stream);

/* The data-parallel exchange fused into the optimizer kernels over NVLink / NVSwitch peer memory (SURVEY.md 8e, 5.8;
 * csrc/optim_p2p.cu): replaces ncclReduceScatter -> Adam on the shard -> ncclAllGather (+ two small all-reduces) of the
 * N > 1 train step.  The *_ptrs_dev arguments are device arrays of `world` 64-bit addresses: buffer r of a symmetric
 * allocation as mapped into this process; *_mc are the NVLS multicast addresses of the same allocations (NULL: peer
 * loads / stores in rank order).  The caller brackets the calls with symmetric-memory barriers.
 *   nrf_small_allreduce_p2p: out[i] = sum_r peer[r][i], i < n; peer[r][n] != 0 on any rank sets state.found_inf.
 *   nrf_adam_step_pair_p2p : rows [row_lo, row_lo + rows) of the interleaved pair buffers: reduce the gradient rows of all
 *                            ranks (multimem.ld_reduce), Adam / EMA as nrf_adam_step_pair, write the fp16 rows to all ranks
 *                            (multimem.st); param / moment / ema pointers address the shard (element 0 = row row_lo). */
int nrf_small_allreduce_p2p(const uint64_t* peer_ptrs_dev, uint32_t world, uint32_t n, float* out, void* state, void* stream);
int nrf_adam_step_pair_p2p(float* param0, float* param1, const uint64_t* grad_ptrs_dev, const float* grad_mc,
                           const uint64_t* half_ptrs_dev, void* half_mc, uint32_t world, uint64_t row_lo, float* exp_avg0,
                           float* exp_avg1, float* exp_avg_sq0, float* exp_avg_sq1, float* ema0, float* ema1, uint64_t rows,
                           const void* state, float lr0, float lr_decay_steps, float beta1, float beta2, float eps,
                           float ema_one_minus_decay, void* stream);

/* ------------------------------------------------------------------ occupancy-grid update (SURVEY 8f NEXT-2) */

/* Renderer.update_state (renderer.py:120-194) around the density query, as device passes with no host read-back.
 * scale_hgs: device f32 [C][2] = (bound_c - half_grid_size_c, half_grid_size_c) per cascade, bound_c = min(2^c, bound),
 * half_grid_size_c = bound_c / H (renderer.py:124-127).  noise: uniform [0,1) per coordinate (torch.rand_like, :129), or NULL.
 *
 * nrf_occ_points_full (:143-155): pts [C, H^3, 3] = jittered cell centres of every cascade, in MORTON order, so that the
 * density of point (c, i) is tmp_grid[c][i] directly (the reference scatters with tmp_grid[cas, morton3D(coords)]).
 * nrf_occ_points_sparse (:157-181): per cascade, slots [0,N) = random cells rnd_cells [C,N,3] i32, slots [N,2N) = random
 * OCCUPIED cells occ_list[c][min(floor(pick * occ_count[c]), occ_count[c]-1)] (index -1 when the cascade has none);
 * writes indices [C,2N] i32 (Morton) and pts [C,2N,3].  occ_list [C,H^3] / occ_count [C] come from nrf_occ_flags +
 * nrf_compact_alive per cascade.
 * nrf_occ_scatter_max: tmp[c][indices[c][j]] = max(tmp, sigmas[c][j] * density_scale) (tmp pre-filled with -1).
 * nrf_occ_update (:183-188): grid = max(grid * decay, values * tmp_scale) where both are >= 0; state[0] = mean(max(grid,0)),
 * state[1] = min(state[0], density_thresh); scratch: nrf_occ_scratch_bytes().
 * nrf_packbits_dev (:189): nrf_packbits with the threshold read from device memory (state + 1). */
int nrf_occ_points_full(float* pts, const float* noise, uint32_t H, uint32_t C, const float* scale_hgs, void* stream);
int nrf_occ_points_sparse(float* pts, int32_t* indices, const float* noise, const int32_t* rnd_cells, const float* pick,
                          const int32_t* occ_list, const int32_t* occ_count, uint32_t N, uint32_t H, uint32_t C,
                          const float* scale_hgs, void* stream);
int nrf_occ_flags(const float* grid, uint32_t H, uint32_t C, int32_t* flags, void* stream);
int nrf_occ_scatter_max(float* tmp, const int32_t* indices, const float* sigmas, float density_scale, uint32_t n_per_cas,
                        uint32_t H, uint32_t C, void* stream);
uint64_t nrf_occ_scratch_bytes(void);
int nrf_occ_update(float* grid, const float* values, float tmp_scale, float decay, uint64_t n, float density_thresh,
                   float* state, void* scratch, void* stream);
int nrf_packbits_dev(const float* grid, uint32_t N, const float* density_thresh_dev, uint8_t* bitfield, void* stream);

/* ------------------------------------------------------------------ reconstruction loss head */

/* The tail of Renderer.render_train (white background, renderer.py:229-232) + Trainer.calc_loss (trainers/base.py:251-304) and
 * their gradients in one launch.  image [N, Cch] f32 = the composited (rgb, class logits), weights_sum [N], target_rgb [N,3],
 * target_cls [N] i64 (NULL when Cch == 3).  out[3] = {mse + class_lambda * ce, mse, ce}; grad_image [N,Cch] and grad_ws [N] are
 * d out[0] / d image and d out[0] / d weights_sum.  Deterministic (fixed reduction order).  scratch:
 * nrf_recon_loss_scratch_bytes(N) bytes, 8-byte aligned, zero before its first use (each call leaves it zeroed again). */
uint64_t nrf_recon_loss_scratch_bytes(uint32_t N);
int nrf_recon_loss(const float* image, const float* weights_sum, const float* target_rgb, const int64_t* target_cls, uint32_t N,
                   uint32_t Cch, float class_lambda, float* out, float* grad_image, float* grad_ws, void* scratch, void* stream);

/* ------------------------------------------------------------------ ray generation (SURVEY 8f NEXT-1) */

/* NerfLib.generate_rays (nerf_lib.py:69-142) + RayBatch.__post_init__ (common.py:139-147) for K rays in one launch.
 * pose: row-major 4x4 camera-to-world f32 (device).  fx, fy, cx, cy: intrinsics rounded to f32 (numpy does the same).
 * The crop window starts at pixel (x0, y0) and is win_w pixels wide (precrop / patch, nerf_lib.py:109-116); ray k is
 * window pixel id = indices ? indices[k] : k (row-major: row = id / win_w, col = id % win_w -- the `indices_1d` of
 * nerf_lib.py:132-134).  camera_flip: bit 2/1/0 negates x/y/z of the camera-frame direction (:122-123).
 * Outputs rays_o, rays_d [K,3] f32 (rays_d unit length).  img ([3,img_h,img_w] f32, or NULL): target [K,3] receives the
 * pixel under each ray (:135-137). */
int nrf_generate_rays(const float* pose, float fx, float fy, float cx, float cy, uint32_t x0, uint32_t y0, uint32_t win_w,
                      uint32_t K, const int64_t* indices, int camera_flip, const float* img, uint32_t img_w,
                      uint32_t img_h, float* rays_o, float* rays_d, float* target, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFSTYLE_B200_H */
