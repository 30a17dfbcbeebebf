// Camera-ray generation on the device (SURVEY.md 8f NEXT-1): the work of NerfLib.generate_rays
// (/root/reference/nerf_lib.py:69-142) and RayBatch.__post_init__ (common.py:139-147) in one kernel -- pixel
// centres -> camera frame -> flip -> world frame -> unit length, the tiled origins, and the target-pixel gather.
// The reference builds the whole frame's meshgrid with numpy on the host every step, uploads it, rotates the whole
// frame and only then picks `bsize` rays; here only the K selected rays are ever computed.
#include "common.cuh"

// One thread per ray.  `indices` (or NULL = 0..K-1) are flat ids over the crop window (row-major, win_w columns),
// exactly the `indices_1d` of nerf_lib.py:132-134; (x0, y0) is the window's offset in the frame (dx, dy of :110-111,
// plus the patch origin of :114-116).
__global__ void __launch_bounds__(256)
k_generate_rays(const float* __restrict__ pose, float fx, float fy, float cx, float cy, uint32_t x0, uint32_t y0,
                uint32_t win_w, uint32_t K, const long long* __restrict__ indices, float sx, float sy, float sz,
                const float* __restrict__ img, uint32_t img_w, uint32_t img_h,
                float* __restrict__ rays_o, float* __restrict__ rays_d, float* __restrict__ target) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const uint32_t id = indices ? (uint32_t)indices[k] : k;
    const uint32_t row = id / win_w, col = id - row * win_w;
    const uint32_t ix = x0 + col, iy = y0 + row;
    // np.linspace(0, w, 2w+1, float32)[1::2] is exactly ix + 0.5 (nerf_lib.py:103-104)
    const float px = (float)ix + 0.5f, py = (float)iy + 0.5f;
    // numpy float32: (i - cx) / fx with cx, fx rounded to float32 first (nerf_lib.py:119-121), then the sign flip (:122-123)
    const float d0 = __fmul_rn(__fdiv_rn(__fsub_rn(px, cx), fx), sx);
    const float d1 = __fmul_rn(__fdiv_rn(__fsub_rn(py, cy), fy), sy);
    const float d2 = sz;
    // rays_d = einsum('ij,hwj->hwi', pose_r, dirs) (:126); pose is a row-major 4x4
    float v[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float a = __fmul_rn(__ldg(pose + 4 * i), d0);
        a = __fmaf_rn(__ldg(pose + 4 * i + 1), d1, a);
        v[i] = __fmaf_rn(__ldg(pose + 4 * i + 2), d2, a);
    }
    // RayBatch: dirs / torch.norm(dirs, dim=-1, keepdim=True) (common.py:147)
    const float n = __fsqrt_rn(__fmaf_rn(v[2], v[2], __fmaf_rn(v[1], v[1], __fmul_rn(v[0], v[0]))));
#pragma unroll
    for (int i = 0; i < 3; i++) {
        rays_d[3 * (size_t)k + i] = __fdiv_rn(v[i], n);
        rays_o[3 * (size_t)k + i] = __ldg(pose + 4 * i + 3);      // torch.tile(pose_t, (K, 1)) (common.py:143-144)
    }
    if (img) {   // einops 'c h w -> h w c' then [coords_y, coords_x] (:135-137)
        const size_t plane = (size_t)img_w * img_h, off = (size_t)iy * img_w + ix;
#pragma unroll
        for (int c = 0; c < 3; c++) target[3 * (size_t)k + c] = __ldg(img + c * plane + off);
    }
}

NRF_EXPORT int nrf_generate_rays(const float* pose, float fx, float fy, float cx, float cy, uint32_t x0, uint32_t y0,
                                 uint32_t win_w, uint32_t K, const int64_t* indices, int camera_flip, const float* img,
                                 uint32_t img_w, uint32_t img_h, float* rays_o, float* rays_d, float* target, void* stream) {
    if (K == 0) return NRF_OK;
    if (!pose || !rays_o || !rays_d || win_w == 0) return NRF_E_INVALID;
    if (img && (!target || img_w == 0 || img_h == 0)) return NRF_E_INVALID;
    // flip = where([(camera_flip >> i) & 1 for i in [2, 1, 0]], -1, 1)  (nerf_lib.py:122)
    const float sx = ((camera_flip >> 2) & 1) ? -1.0f : 1.0f;
    const float sy = ((camera_flip >> 1) & 1) ? -1.0f : 1.0f;
    const float sz = (camera_flip & 1) ? -1.0f : 1.0f;
    k_generate_rays<<<ceil_div_u32(K, 256), 256, 0, (cudaStream_t)stream>>>(
        pose, fx, fy, cx, cy, x0, y0, win_w, K, reinterpret_cast<const long long*>(indices), sx, sy, sz, img, img_w, img_h,
        rays_o, rays_d, target);
    return nrf_check_launch();
}
