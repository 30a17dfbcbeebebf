/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker, never the product).
 *
 * Plain-C (C11 + OpenMP) CPU restatement of the reference's CUDA kernels for the NeRF
 * render/train hot path of hkust-vgd/nerfstyle.  Every function cites the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path (nerfstyle_b200/) never does.
 *
 * Floating point: compiled with -ffp-contract=off; the places where nvcc (-fmad=true) fuses a
 * multiply-add in the reference kernels are written as explicit fmaf() calls (SURVEY.md 8a.3,
 * read from the PTX of the reference kernels built with nvcc 12.9 for compute_100a).  Integer
 * outputs (sample counts, offsets, cell / hash indices, packed bits) are therefore bit-exact
 * restatements; float outputs that go through the GPU's ex2.approx (the compositing alpha) are
 * tolerance-checked only.
 *
 * Pinning: the reference has no tests or golden vectors of its own.  The oracle is
 * pinned (a) by the closed-form known-answer values derived from the reference formulas
 * (tests/test_oracle_kat.py) and (b) on the GPU box against the reference's own CUDA extensions
 * rebuilt for sm_100a (oracle/build_ref.sh -> oracle/_ref/, tests/test_ref_ext_gpu.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_API __attribute__((visibility("default")))

typedef _Float16 half_t;

/* ------------------------------------------------------------------------------------------
 * helpers: raymarching/src/raymarching.cu:30-81
 * ---------------------------------------------------------------------------------------- */
static inline float signf_(float x) { return copysignf(1.0f, x); }                 /* :30-32 */
static inline float clampf_(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); } /* :34-36 */

static inline int mip_from_pos(float x, float y, float z, float max_cascade) {      /* :42-47 */
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}
static inline int mip_from_dt(float dt, float H, float max_cascade) {               /* :49-54 */
    const float mx = (float)((double)(dt * H) * 0.5);
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}
static inline uint32_t expand_bits(uint32_t v) {                                    /* :56-63 */
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3D_(uint32_t x, uint32_t y, uint32_t z) {              /* :65-71 */
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3D_invert_(uint32_t x) {                               /* :73-81 */
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

ORC_API void orc_morton3D(const int32_t* coords, uint32_t N, int32_t* indices) {    /* :313-325 */
    for (uint32_t n = 0; n < N; n++)
        indices[n] = (int32_t)morton3D_((uint32_t)coords[3*n], (uint32_t)coords[3*n+1], (uint32_t)coords[3*n+2]);
}
ORC_API void orc_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords) { /* :336-353 */
    for (uint32_t n = 0; n < N; n++) {
        const int32_t ind = indices[n];
        coords[3*n+0] = (int32_t)morton3D_invert_((uint32_t)(ind >> 0));
        coords[3*n+1] = (int32_t)morton3D_invert_((uint32_t)(ind >> 1));
        coords[3*n+2] = (int32_t)morton3D_invert_((uint32_t)(ind >> 2));
    }
}

/* raymarching.cu:367-388 -- bit i of byte n = grid[8n+i] > thresh (strict) */
ORC_API void orc_packbits(const float* grid, uint32_t N, float thresh, uint8_t* bitfield) {
    #pragma omp parallel for schedule(static)
    for (uint32_t n = 0; n < N; n++) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; i++) bits |= (grid[8*(size_t)n+i] > thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* raymarching.cu:191-244 */
ORC_API void orc_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                                    uint32_t N, float min_near, float* nears, float* fars) {
    #pragma omp parallel for schedule(static)
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[3*n], oy = rays_o[3*n+1], oz = rays_o[3*n+2];
        const float dx = rays_d[3*n], dy = rays_d[3*n+1], dz = rays_d[3*n+2];
        const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, tmp;
        if (near > far) { tmp = near; near = far; far = tmp; }
        float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
        if (near_y > far_y) { tmp = near_y; near_y = far_y; far_y = tmp; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { tmp = near_z; near_z = far_z; far_z = tmp; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near; fars[n] = far;
    }
}

/* raymarching.cu:262-297 */
ORC_API void orc_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords) {
    const float RPI = 0.3183098861837907f;
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[3*n], oy = rays_o[3*n+1], oz = rays_o[3*n+2];
        const float dx = rays_d[3*n], dy = rays_d[3*n+1], dz = rays_d[3*n+2];
        const float A = dx*dx + dy*dy + dz*dz;
        const float B = ox*dx + oy*dy + oz*dz;
        const float C = ox*ox + oy*oy + oz*oz - radius*radius;
        const float t = (-B + sqrtf(B*B - A*C)) / A;
        const float x = ox + t*dx, y = oy + t*dy, z = oz + t*dz;
        const float theta = atan2f(sqrtf(x*x + z*z), y);
        const float phi = atan2f(z, x);
        coords[2*n] = 2 * theta * RPI - 1;
        coords[2*n+1] = phi * RPI;
    }
}

/* ------------------------------------------------------------------------------------------
 * The marching state machine shared by kernel_march_rays_train (raymarching.cu:411-589) and
 * kernel_march_rays (:1005-1120).  One call = one visit of the loop body at the current t.
 * FMA placement per SURVEY.md 8a.3.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, rH, H3, bound, dt_gamma, dt_min, dt_max;
    float Cf, Hf; uint32_t H;
    const uint8_t* grid;
} march_ctx;

static inline void march_ctx_init(march_ctx* c, const float* o, const float* d, const uint8_t* grid,
                                  float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    c->ox = o[0]; c->oy = o[1]; c->oz = o[2];
    c->dx = d[0]; c->dy = d[1]; c->dz = d[2];
    c->rdx = 1 / c->dx; c->rdy = 1 / c->dy; c->rdz = 1 / c->dz;
    c->rH = 1 / (float)H;
    c->H3 = (float)(H * H * H);
    c->bound = bound; c->dt_gamma = dt_gamma;
    c->dt_min = 3.4641016f / (float)max_steps;                       /* 2*SQRT3()/max_steps, :446 */
    c->dt_max = (3.4641016f * (float)(1 << (C - 1))) / (float)H;      /* :447 */
    c->Cf = (float)C; c->Hf = (float)H; c->H = H; c->grid = grid;
}

/* returns 1 when the cell at t is occupied (sample emitted: x,y,z,dt valid), else 0 and *t is
 * advanced past the empty cell exactly as the reference's do/while does (:491-499). */
static inline int march_visit(const march_ctx* c, float* t_io, float* x_, float* y_, float* z_, float* dt_) {
    const float t = *t_io;
    const float x = clampf_(fmaf(c->dx, t, c->ox), -c->bound, c->bound);
    const float y = clampf_(fmaf(c->dy, t, c->oy), -c->bound, c->bound);
    const float z = clampf_(fmaf(c->dz, t, c->oz), -c->bound, c->bound);
    const float dt = clampf_(t * c->dt_gamma, c->dt_min, c->dt_max);
    int level = mip_from_pos(x, y, z, c->Cf);
    const int l2 = mip_from_dt(dt, c->Hf, c->Cf);
    if (l2 > level) level = l2;
    const float mip_bound = fminf(scalbnf(1.0f, level), c->bound);
    const float mip_rbound = 1 / mip_bound;
    const float Hm1 = (float)(c->H - 1);
    const int nx = (int)clampf_((float)(0.5 * (double)fmaf(x, mip_rbound, 1.0f) * (double)c->H), 0.0f, Hm1);
    const int ny = (int)clampf_((float)(0.5 * (double)fmaf(y, mip_rbound, 1.0f) * (double)c->H), 0.0f, Hm1);
    const int nz = (int)clampf_((float)(0.5 * (double)fmaf(z, mip_rbound, 1.0f) * (double)c->H), 0.0f, Hm1);
    const uint32_t index = (uint32_t)fmaf(c->H3, (float)level, (float)morton3D_((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    const int occ = (c->grid[index / 8] & (1u << (index % 8))) != 0;
    *x_ = x; *y_ = y; *z_ = z; *dt_ = dt;
    if (occ) return 1;
    const float tx = (fmaf(c->rH * (0.5f * signf_(c->dx) + ((float)nx + 0.5f)), 2.0f, -1.0f) * mip_bound - x) * c->rdx;
    const float ty = (fmaf(c->rH * (0.5f * signf_(c->dy) + ((float)ny + 0.5f)), 2.0f, -1.0f) * mip_bound - y) * c->rdy;
    const float tz = (fmaf(c->rH * (0.5f * signf_(c->dz) + ((float)nz + 0.5f)), 2.0f, -1.0f) * mip_bound - z) * c->rdz;
    const float tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    float tn = t;
    do { tn += clampf_(tn * c->dt_gamma, c->dt_min, c->dt_max); } while (tn < tt);
    *t_io = tn;
    return 0;
}

/* raymarching.cu:411-589 first pass (count only).  counts[n] = num_steps of ray n. */
ORC_API void orc_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid,
        float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
        const float* nears, const float* fars, const float* noises, int32_t* counts) {
    #pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t n = 0; n < N; n++) {
        march_ctx c; march_ctx_init(&c, rays_o + 3*(size_t)n, rays_d + 3*(size_t)n, grid, bound, dt_gamma, max_steps, C, H);
        const float near = nears[n], far = fars[n], noise = noises ? noises[n] : 0.0f;
        float t = fmaf(noise, clampf_(near * dt_gamma, c.dt_min, c.dt_max), near);   /* :449-452 */
        uint32_t num_steps = 0;
        float x, y, z, dt;
        while (t < far && num_steps < max_steps) {
            if (march_visit(&c, &t, &x, &y, &z, &dt)) { num_steps++; t += dt; }
        }
        counts[n] = (int32_t)num_steps;
    }
}

/* raymarching.cu:411-589 (both passes).  The reference reserves output slots with racing
 * atomicAdd()s (:506-507); this restatement uses the schedule "rays arrive in index order", i.e.
 * offsets are the exclusive scan of the counts in ray order and rays[n] = (n, offset, count).
 * counter[0] += total samples, counter[1] += N, like the atomics.  xyzs/dirs/deltas must be
 * zero-filled [M,3],[M,3],[M,4] by the caller (raymarching.py:238-240). */
ORC_API void orc_march_rays_train(const float* rays_o, const float* rays_d, const float* z_hats, const uint8_t* grid,
        float bound, float dt_gamma, uint32_t max_steps, int is_ndc, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
        const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
        int32_t* rays, int32_t* counter, const float* noises) {
    int32_t* counts = (int32_t*)malloc(sizeof(int32_t) * (N ? N : 1));
    orc_march_rays_train_count(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, noises, counts);
    uint32_t run = (uint32_t)counter[0];
    const uint32_t ray_base = (uint32_t)counter[1];
    for (uint32_t n = 0; n < N; n++) {
        rays[3*(size_t)(ray_base + n) + 0] = (int32_t)n;
        rays[3*(size_t)(ray_base + n) + 1] = (int32_t)run;
        rays[3*(size_t)(ray_base + n) + 2] = counts[n];
        run += (uint32_t)counts[n];
    }
    counter[0] = (int32_t)run; counter[1] = (int32_t)(ray_base + N);
    #pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t num_steps = (uint32_t)counts[n];
        const uint32_t point_index = (uint32_t)rays[3*(size_t)(ray_base + n) + 1];
        if (num_steps == 0) continue;                                  /* :516 */
        if (point_index + num_steps >= M) continue;                    /* :517 (note >=) */
        march_ctx c; march_ctx_init(&c, rays_o + 3*(size_t)n, rays_d + 3*(size_t)n, grid, bound, dt_gamma, max_steps, C, H);
        const float near = nears[n], far = fars[n], noise = noises ? noises[n] : 0.0f;
        float t = fmaf(noise, clampf_(near * dt_gamma, c.dt_min, c.dt_max), near);
        float* px = xyzs + 3*(size_t)point_index; float* pd = dirs + 3*(size_t)point_index; float* pl = deltas + 4*(size_t)point_index;
        uint32_t step = 0;
        float last_t = t;
        float last_z = clampf_(fmaf(c.dz, t, c.oz), -bound, bound);
        float x, y, z, dt;
        while (t < far && step < num_steps) {
            if (march_visit(&c, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t; last_t = t;
                if (is_ndc) {                                          /* :566-571 (last_z = z quirk) */
                    const float new_z = clampf_(fmaf(c.dz, t, c.oz), -bound, bound);
                    pl[2] = (2 / (new_z - 1) - 2 / (z - 1)) / z_hats[n];
                    pl[3] = (2 / (new_z - 1) - 2 / (last_z - 1)) / z_hats[n];
                    last_z = z;
                }
                px += 3; pd += 3; pl += 4; step++;
            }
        }
    }
    free(counts);
}

/* raymarching.cu:1005-1120 */
ORC_API void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
        const float* rays_o, const float* rays_d, const float* z_hats, float bound, float dt_gamma, uint32_t max_steps,
        int is_ndc, uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars,
        float* xyzs, float* dirs, float* deltas, const float* noises) {
    #pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        const float noise = noises ? noises[n] : 0.0f;
        march_ctx c; march_ctx_init(&c, rays_o + 3*(size_t)index, rays_d + 3*(size_t)index, grid, bound, dt_gamma, max_steps, C, H);
        float t = rays_t[(size_t)index * (is_ndc ? 2 : 1)];
        const float far = fars[index];
        float* px = xyzs + 3*(size_t)n*n_step; float* pd = dirs + 3*(size_t)n*n_step; float* pl = deltas + 4*(size_t)n*n_step;
        uint32_t step = 0;
        t = fmaf(noise, clampf_(t * dt_gamma, c.dt_min, c.dt_max), t);  /* :1053 */
        float last_t = t;
        float last_z = clampf_(fmaf(c.dz, t, c.oz), -bound, bound);
        float x, y, z, dt;
        while (t < far && step < n_step) {
            if (march_visit(&c, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                if (is_ndc) {                                          /* :1094-1099 (last_z = new_z) */
                    const float new_z = clampf_(fmaf(c.dz, t, c.oz), -bound, bound);
                    pl[2] = (2 / (new_z - 1) - 2 / (z - 1)) / z_hats[index];
                    pl[3] = (2 / (new_z - 1) - 2 / (last_z - 1)) / z_hats[index];
                    last_z = new_z;
                }
                last_t = t;
                px += 3; pd += 3; pl += 4; step++;
            }
        }
    }
}

/* __expf(x) on the GPU = ex2.approx(x * log2e); the CPU restatement keeps the
 * same two multiplies and uses exp2f for the (approximate on GPU) ex2. */
static inline float alpha_of(float sigma, float delta) {
    const float p = sigma * delta;
    const float q = p * -1.4426950408889634f;
    return 1.0f - exp2f(q);
}

/* raymarching.cu:807-879 */
ORC_API void orc_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
        const int32_t* rays, uint32_t M, uint32_t N, uint32_t C, float T_thresh, int is_ndc,
        float* weights_sum, float* depth, float* image) {
    #pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[3*n], offset = (uint32_t)rays[3*n+1], num_steps = (uint32_t)rays[3*n+2];
        float* img = image + (size_t)index * C;
        for (uint32_t i = 0; i < C; i++) img[i] = 0;
        if (num_steps == 0 || offset + num_steps >= M) { weights_sum[index] = 0; depth[index] = 0; continue; }
        const float* s = sigmas + offset; const float* r = rgbs + (size_t)offset * C; const float* dl = deltas + (size_t)offset * 4;
        uint32_t step = 0;
        float T = 1.0f, ws = 0, t = 0, d = 0;
        while (step < num_steps) {
            const float alpha = alpha_of(s[0], is_ndc ? dl[2] : dl[0]);
            const float weight = alpha * T;
            for (uint32_t i = 0; i < C; i++) img[i] = fmaf(weight, r[i], img[i]);
            t += (is_ndc ? dl[3] : dl[1]);
            d = fmaf(weight, t, d);
            ws += weight;
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
            s++; r += C; dl += 4; step++;
        }
        weights_sum[index] = ws; depth[index] = d;
    }
}

/* raymarching.cu:905-986.  grad_sigmas / grad_rgbs must be zero-filled by the caller
 * (raymarching.py:339-340); rgbs_buf is internal scratch here. */
ORC_API void orc_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
        const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays, int is_ndc,
        const float* weights_sum, const float* image, uint32_t M, uint32_t N, uint32_t C, float T_thresh,
        float* grad_sigmas, float* grad_rgbs) {
    #pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t n = 0; n < N; n++) {
        const uint32_t index = (uint32_t)rays[3*n], offset = (uint32_t)rays[3*n+1], num_steps = (uint32_t)rays[3*n+2];
        if (num_steps == 0 || offset + num_steps >= M) continue;
        const float* gi = grad_image + (size_t)index * C; const float* img = image + (size_t)index * C;
        const float gws = grad_weights_sum[index];
        const float* s = sigmas + offset; const float* r = rgbs + (size_t)offset * C; const float* dl = deltas + (size_t)offset * 4;
        float* gs = grad_sigmas + offset; float* gr = grad_rgbs + (size_t)offset * C;
        float buf[64];
        for (uint32_t i = 0; i < C; i++) buf[i] = 0;
        uint32_t step = 0;
        float T = 1.0f; const float ws_final = weights_sum[index]; float ws = 0;
        while (step < num_steps) {
            const float alpha = alpha_of(s[0], is_ndc ? dl[2] : dl[0]);
            const float weight = alpha * T;
            for (uint32_t i = 0; i < C; i++) buf[i] = fmaf(weight, r[i], buf[i]);
            ws += weight;
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
            for (uint32_t i = 0; i < C; i++) gr[i] = gi[i] * weight;
            float gsum = 0;
            for (uint32_t i = 0; i < C; i++) gsum = fmaf(gi[i], fmaf(T, r[i], -(img[i] - buf[i])), gsum);
            gs[0] = (is_ndc ? dl[2] : dl[0]) * fmaf(gws, 1 - ws_final, gsum);
            s++; r += C; dl += 4; gs++; gr += C; step++;
        }
        (void)ws;
    }
}

/* raymarching.cu:1134-1231 (in place on rays_alive, rays_t, weights_sum, depth, image) */
ORC_API void orc_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
        const float* sigmas, const float* rgbs, const float* deltas, uint32_t C, int is_ndc,
        float* weights_sum, float* depth, float* image) {
    #pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t n = 0; n < n_alive; n++) {
        const int32_t index = rays_alive[n];
        const float* s = sigmas + (size_t)n * n_step; const float* r = rgbs + (size_t)n * n_step * C; const float* dl = deltas + (size_t)n * n_step * 4;
        float* rt = rays_t + (size_t)index * (is_ndc ? 2 : 1);
        float* img = image + (size_t)index * C;
        float t_rm = 0, t_phy;
        if (is_ndc) { t_rm = rt[0]; t_phy = rt[1]; } else { t_phy = rt[0]; }
        float weight_sum = weights_sum[index], d = depth[index];
        uint32_t step = 0;
        while (step < n_step) {
            if (dl[0] == 0) break;
            const float alpha = alpha_of(s[0], is_ndc ? dl[2] : dl[0]);
            const float T = 1 - weight_sum;
            const float weight = alpha * T;
            weight_sum += weight;
            if (is_ndc) { t_rm += dl[1]; t_phy += dl[3]; } else { t_phy += dl[1]; }
            d = fmaf(weight, t_phy, d);
            for (uint32_t i = 0; i < C; i++) img[i] = fmaf(weight, r[i], img[i]);
            if (T < T_thresh) break;
            s++; r += C; dl += 4; step++;
        }
        if (step < n_step) rays_alive[n] = -1;
        else { if (is_ndc) { rt[0] = t_rm; rt[1] = t_phy; } else rt[0] = t_phy; }
        weights_sum[index] = weight_sum; depth[index] = d;
    }
}

/* ------------------------------------------------------------------------------------------
 * Hash-grid encoder: gridencoder/src/gridencoder.cu
 * ---------------------------------------------------------------------------------------- */
static const uint32_t PRIMES[7] = { 1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u };

static inline uint32_t fast_hash(uint32_t D, const uint32_t* pos_grid, uint32_t style) {   /* :35-52 */
    uint32_t result = 0;
    for (uint32_t i = 0; i < D; i++) result ^= pos_grid[i] * PRIMES[i];
    result ^= style * PRIMES[D];
    return result;
}
ORC_API uint32_t orc_fast_hash3(uint32_t x, uint32_t y, uint32_t z, uint32_t style) {
    const uint32_t p[3] = { x, y, z };
    return fast_hash(3, p, style);
}
/* :55-80 ; returns the row*C + ch float offset inside the level */
static inline uint32_t get_grid_index(uint32_t D, uint32_t C, uint32_t gridtype, uint32_t ch, uint32_t hashmap_size,
                                      uint32_t resolution, const uint32_t* pos_grid, uint32_t style) {
    uint32_t stride = 1, index = 0;
    const uint32_t max_styles = 512;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pos_grid[d] * stride;
        stride *= (resolution + 1);
    }
    if (stride <= hashmap_size) { index += style * stride; stride *= max_styles; }
    if (gridtype == 0 && stride > hashmap_size) index = fast_hash(D, pos_grid, style);
    return (index % hashmap_size) * C + ch;
}
ORC_API uint32_t orc_grid_index3(uint32_t C, uint32_t gridtype, uint32_t hashmap_size, uint32_t resolution,
                                 uint32_t x, uint32_t y, uint32_t z, uint32_t style) {
    const uint32_t p[3] = { x, y, z };
    return get_grid_index(3, C, gridtype, 0, hashmap_size, resolution, p, style);
}
/* :137 -- per-level resolution as the kernel computes it */
ORC_API uint32_t orc_level_resolution(uint32_t level, float S, uint32_t H) {
    return (uint32_t)floorf(exp2f((float)level * S) * (float)H);
}

typedef struct { uint32_t pos_grid[5]; float pos[5]; int oob; uint32_t hashmap_size, resolution; float scale; } level_pt;

static inline void locate(level_pt* p, const float* in, const int32_t* offsets, uint32_t level, uint32_t D,
                          float S, uint32_t H, int align_corners) {                       /* :107-149 */
    p->oob = 0;
    for (uint32_t d = 0; d < D; d++) if (in[d] < 0 || in[d] > 1) p->oob = 1;
    p->hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    p->resolution = orc_level_resolution(level, S, H);
    p->scale = (float)(p->resolution - (align_corners ? 0u : 1u));
    for (uint32_t d = 0; d < D; d++) {
        float pos = fmaf(in[d], p->scale, align_corners ? 0.0f : 0.5f);
        p->pos_grid[d] = (uint32_t)fminf(floorf(pos), (float)(p->resolution - 1));
        p->pos[d] = pos - (float)p->pos_grid[d];
    }
}

/* kernel_grid :83-188 (+ dy_dx :191-234).  outputs are LEVEL-MAJOR [L,B,C] like the kernel writes
 * them (the permute to [B,L*C] is done by grid.py:58).  `half_tables` != 0 emulates scalar_t =
 * at::Half: tables/outputs are IEEE binary16, every `results += w * grid` rounds the float product
 * to half and the float sum to half (c10 Half operator+=).  idx_out (optional, [L,B,8] uint32)
 * records the row index of every corner for the bit-exact index test. */
ORC_API void orc_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs,
        uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs, void* dy_dx,
        uint32_t gridtype, int align_corners, uint32_t style, int half_tables, uint32_t* idx_out) {
    const uint32_t ncorner = 1u << D;
    #pragma omp parallel for collapse(2) schedule(static)
    for (uint32_t level = 0; level < L; level++) {
        for (uint32_t b = 0; b < B; b++) {
            const float* in = inputs + (size_t)b * D;
            const size_t out_off = ((size_t)level * B + b) * C;
            const size_t grid_off = (size_t)(uint32_t)offsets[level] * C;
            level_pt p; locate(&p, in, offsets, level, D, S, H, align_corners);
            if (p.oob) {
                for (uint32_t ch = 0; ch < C; ch++) { if (half_tables) ((half_t*)outputs)[out_off + ch] = 0; else ((float*)outputs)[out_off + ch] = 0; }
                if (calc_grad_inputs) for (uint32_t k = 0; k < D * C; k++) {
                    const size_t o = (size_t)b * D * L * C + (size_t)level * D * C + k;
                    if (half_tables) ((half_t*)dy_dx)[o] = 0; else ((float*)dy_dx)[o] = 0;
                }
                if (idx_out) for (uint32_t i = 0; i < ncorner; i++) idx_out[((size_t)level * B + b) * ncorner + i] = 0xFFFFFFFFu;
                continue;
            }
            float resf[8] = {0}; half_t resh[8] = {0};
            for (uint32_t idx = 0; idx < ncorner; idx++) {
                float w = 1; uint32_t pl[5];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - p.pos[d]; pl[d] = p.pos_grid[d]; }
                    else { w *= p.pos[d]; pl[d] = p.pos_grid[d] + 1; }
                }
                const uint32_t index = get_grid_index(D, C, gridtype, 0, p.hashmap_size, p.resolution, pl, style);
                if (idx_out) idx_out[((size_t)level * B + b) * ncorner + idx] = index / C;
                for (uint32_t ch = 0; ch < C; ch++) {
                    if (half_tables) {
                        const half_t prod = (half_t)(w * (float)((const half_t*)embeddings)[grid_off + index + ch]);
                        resh[ch] = (half_t)((float)resh[ch] + (float)prod);
                    } else {
                        resf[ch] = fmaf(w, ((const float*)embeddings)[grid_off + index + ch], resf[ch]);
                    }
                }
            }
            for (uint32_t ch = 0; ch < C; ch++) { if (half_tables) ((half_t*)outputs)[out_off + ch] = resh[ch]; else ((float*)outputs)[out_off + ch] = resf[ch]; }
            if (calc_grad_inputs) {
                const size_t dbase = (size_t)b * D * L * C + (size_t)level * D * C;
                for (uint32_t gd = 0; gd < D; gd++) {
                    float gf[8] = {0}; half_t gh[8] = {0};
                    for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                        float w = p.scale; uint32_t pl[5];
                        for (uint32_t nd = 0; nd < D - 1; nd++) {
                            const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                            if ((idx & (1u << nd)) == 0) { w *= 1 - p.pos[d]; pl[d] = p.pos_grid[d]; }
                            else { w *= p.pos[d]; pl[d] = p.pos_grid[d] + 1; }
                        }
                        pl[gd] = p.pos_grid[gd];
                        const uint32_t il = get_grid_index(D, C, gridtype, 0, p.hashmap_size, p.resolution, pl, style);
                        pl[gd] = p.pos_grid[gd] + 1;
                        const uint32_t ir = get_grid_index(D, C, gridtype, 0, p.hashmap_size, p.resolution, pl, style);
                        for (uint32_t ch = 0; ch < C; ch++) {
                            if (half_tables) {
                                const half_t* e = (const half_t*)embeddings;
                                const half_t diff = (half_t)((float)e[grid_off + ir + ch] - (float)e[grid_off + il + ch]);
                                const half_t prod = (half_t)(w * (float)diff);
                                gh[ch] = (half_t)((float)gh[ch] + (float)prod);
                            } else {
                                const float* e = (const float*)embeddings;
                                gf[ch] = fmaf(w, e[grid_off + ir + ch] - e[grid_off + il + ch], gf[ch]);
                            }
                        }
                    }
                    for (uint32_t ch = 0; ch < C; ch++) {
                        if (half_tables) ((half_t*)dy_dx)[dbase + gd * C + ch] = gh[ch]; else ((float*)dy_dx)[dbase + gd * C + ch] = gf[ch];
                    }
                }
            }
        }
    }
}

/* kernel_grid_backward :238-328.  grad is LEVEL-MAJOR [L,B,C] (grid.py:80).  The reference adds with
 * racing atomics; this restatement adds in point order.  Float tables: the sum is carried in double
 * and rounded once (order-independent "true" value).  Half tables: each add rounds to binary16 like
 * the __half2 atomics (:313-319), in point order (one of the many valid schedules).
 * grad_embeddings must be zero-filled [rows, C] by the caller (grid.py:82). */
ORC_API void orc_grid_encode_backward(const void* grad, const float* inputs, const int32_t* offsets, void* grad_embeddings,
        uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
        uint32_t gridtype, int align_corners, uint32_t style, int half_tables) {
    const uint32_t ncorner = 1u << D;
    #pragma omp parallel for schedule(dynamic, 1)
    for (uint32_t level = 0; level < L; level++) {
        const size_t grid_off = (size_t)(uint32_t)offsets[level] * C;
        const uint32_t rows = (uint32_t)(offsets[level + 1] - offsets[level]);
        double* acc = half_tables ? NULL : (double*)calloc((size_t)rows * C, sizeof(double));
        for (uint32_t b = 0; b < B; b++) {
            const float* in = inputs + (size_t)b * D;
            level_pt p; locate(&p, in, offsets, level, D, S, H, align_corners);
            if (p.oob) continue;
            for (uint32_t idx = 0; idx < ncorner; idx++) {
                float w = 1; uint32_t pl[5];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - p.pos[d]; pl[d] = p.pos_grid[d]; }
                    else { w *= p.pos[d]; pl[d] = p.pos_grid[d] + 1; }
                }
                const uint32_t index = get_grid_index(D, C, gridtype, 0, p.hashmap_size, p.resolution, pl, style);
                for (uint32_t ch = 0; ch < C; ch++) {
                    const size_t go = ((size_t)level * B + b) * C + ch;
                    if (half_tables) {
                        half_t* ge = (half_t*)grad_embeddings;
                        const half_t v = (half_t)(w * (float)((const half_t*)grad)[go]);
                        ge[grid_off + index + ch] = (half_t)((float)ge[grid_off + index + ch] + (float)v);
                    } else {
                        acc[index + ch] += (double)(w * ((const float*)grad)[go]);
                    }
                }
            }
        }
        if (!half_tables) {
            float* ge = (float*)grad_embeddings;
            for (size_t i = 0; i < (size_t)rows * C; i++) ge[grid_off + i] = (float)acc[i];
            free(acc);
        }
    }
}

/* kernel_input_backward :331-357 (float only) */
ORC_API void orc_grid_input_backward(const float* grad, const float* dy_dx, float* grad_inputs,
                                     uint32_t B, uint32_t D, uint32_t C, uint32_t L) {
    for (uint32_t t = 0; t < B * D; t++) {
        const uint32_t b = t / D, d = t - b * D;
        const float* dd = dy_dx + (size_t)b * L * D * C;
        float result = 0;
        for (uint32_t l = 0; l < L; l++)
            for (uint32_t ch = 0; ch < C; ch++)
                result = fmaf(grad[(size_t)l * B * C + (size_t)b * C + ch], dd[l * D * C + d * C + ch], result);
        grad_inputs[t] = result;
    }
}

/* kernel_grid_initialize + host loop :497-548 (D=3, C=2 hard-coded like the reference) */
ORC_API void orc_grid_initialize(const float* ref_embeddings, float* embeddings, const int32_t* ref_offsets,
                                 const int32_t* offsets, uint32_t L, float S, uint32_t H, uint32_t Ns) {
    for (uint32_t level = 0; level < L; level++) {
        const uint32_t resolution = orc_level_resolution(level, S, H);
        const uint32_t ext = ((resolution + 1 + 7) / 8) * 8;            /* launched threads per axis */
        const float* rg = ref_embeddings + (size_t)(uint32_t)ref_offsets[level] * 2;
        float* g = embeddings + (size_t)(uint32_t)offsets[level] * 2;
        const uint32_t ref_hs = (uint32_t)(ref_offsets[level + 1] - ref_offsets[level]);
        const uint32_t hs = (uint32_t)(offsets[level + 1] - offsets[level]);
        for (uint32_t z = 0; z < ext; z++) for (uint32_t y = 0; y < ext; y++) for (uint32_t x = 0; x < ext; x++) {
            if (x > resolution || y > resolution || z > resolution) continue;
            const uint32_t pg[3] = { x, y, z };
            const uint32_t ri = get_grid_index(3, 2, 0, 0, ref_hs, resolution, pg, 0);
            const float v0 = rg[ri], v1 = rg[ri + 1];
            for (uint32_t s = 0; s < Ns; s++) {
                const uint32_t si = get_grid_index(3, 2, 0, 0, hs, resolution, pg, s);
                g[si] = v0; g[si + 1] = v1;
            }
        }
    }
}
