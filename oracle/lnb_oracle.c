/*
 * lnb_oracle.c - CPU restatement of the reference's per-ray volume-rendering hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library; nothing under lidar-nerf_b200/ does.
 *
 * Every function restates, in scalar C, what ONE CUDA thread of the reference does (citations are
 * paths under the reference checkout), looped over the batch (OpenMP over rays/samples where the
 * iterations are independent).  Where nvcc fuses a*b+c into an FMA in the reference's device code the
 * restatement calls fmaf() explicitly and the file is compiled with -ffp-contract=off, so that the
 * integer results of the march (sample counts, cell indices) can be compared bit-for-bit.
 *
 * Pinning: the reference has no tests or golden vectors.  This oracle is pinned
 * against (a) outputs of the reference's own CUDA kernels compiled unmodified (oracle/build_ref.py ->
 * oracle/_ref) and frozen into tests/golden/ref_cuda_*.npz by tests/golden/make_golden_gpu.py on the
 * GPU box, and (b) outputs of the reference's Python code importable on CPU (encoding.FreqEncoder,
 * activation.trunc_exp, renderer.sample_pdf / NeRFRenderer.run) frozen by tests/golden/make_golden_cpu.py.
 *
 * Transcendentals: the reference uses __expf / __sinf (fast-math intrinsics); here they are libm
 * expf / sinf, so float outputs agree to ~1e-6 relative (documented tolerances live in the tests).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline float to_half_precision(float v) { return (float)(_Float16)v; }
/* bf16 (round to nearest even on the upper 16 bits of the fp32 pattern), for the bf16 build of the MLP kernels */
static inline float to_bf16_precision(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return v;               /* NaN */
    u += 0x7fffu + ((u >> 16) & 1u);
    u &= 0xffff0000u;
    memcpy(&v, &u, 4);
    return v;
}
/* element type of the fused MLPs: 0 = fp16 (default); bit 0 = bf16 activations / weights / activation gradients;
 * bit 1 = the INPUT gradient (grad_inputs) is nevertheless rounded to fp16 (the density MLP hands it to the hash-grid
 * scatter, which reads fp16).  Set by orc_set_mlp_mode; mirrors -DLNB_BF16 of lidar-nerf_b200/csrc/mlp_tiles.cuh. */
static int g_mlp_mode = 0;
ORC_API void orc_set_mlp_mode(int mode) { g_mlp_mode = mode; }
static inline float mlp_round(float v) { return (g_mlp_mode & 1) ? to_bf16_precision(v) : to_half_precision(v); }
static inline float mlp_round_dx(float v) { return ((g_mlp_mode & 1) && !(g_mlp_mode & 2)) ? to_bf16_precision(v) : to_half_precision(v); }

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * raymarching/src/raymarching.cu
 * ---------------------------------------------------------------------------------------------- */

/* :71-95 */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3_invert(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* :105-157 */
ORC_API void orc_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N,
                                    float min_near, float *nears, float *fars) {
    for (uint32_t n = 0; n < N; ++n) {
        const float *o = rays_o + 3 * (size_t)n, *d = rays_d + 3 * (size_t)n;
        float near = 0, far = 0;
        int miss = 0;
        for (int a = 0; a < 3 && !miss; ++a) {
            const float r = 1 / d[a];
            float t0 = (aabb[a] - o[a]) * r, t1 = (aabb[a + 3] - o[a]) * r;
            if (t0 > t1) { float s = t0; t0 = t1; t1 = s; }
            if (a == 0) { near = t0; far = t1; continue; }
            if (near > t1 || t0 > far) { miss = 1; break; }
            if (t0 > near) near = t0;
            if (t1 < far) far = t1;
        }
        if (miss) { nears[n] = fars[n] = 3.402823466e+38f; continue; }
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* :183-217 */
ORC_API void orc_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords) {
    const float RPI = 0.3183098861837907f;
    for (uint32_t n = 0; n < N; ++n) {
        const float *o = rays_o + 3 * (size_t)n, *d = rays_d + 3 * (size_t)n;
        const float A = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        const float B = o[0] * d[0] + o[1] * d[1] + o[2] * d[2];
        const float C = o[0] * o[0] + o[1] * o[1] + o[2] * o[2] - radius * radius;
        const float t = (-B + sqrtf(B * B - A * C)) / A;
        const float x = o[0] + t * d[0], y = o[1] + t * d[1], z = o[2] + t * d[2];
        coords[2 * n] = 2 * atan2f(sqrtf(x * x + z * z), y) * RPI - 1;
        coords[2 * n + 1] = atan2f(z, x) * RPI;
    }
}

/* :237-272 */
ORC_API void orc_morton3D(const int32_t *coords, uint32_t N, int32_t *indices) {
    for (uint32_t n = 0; n < N; ++n)
        indices[n] = (int32_t)morton3((uint32_t)coords[3 * n], (uint32_t)coords[3 * n + 1], (uint32_t)coords[3 * n + 2]);
}
ORC_API void orc_morton3D_invert(const int32_t *indices, uint32_t N, int32_t *coords) {
    for (uint32_t n = 0; n < N; ++n) {
        const int32_t v = indices[n];
        coords[3 * n] = (int32_t)morton3_invert((uint32_t)(v >> 0));
        coords[3 * n + 1] = (int32_t)morton3_invert((uint32_t)(v >> 1));
        coords[3 * n + 2] = (int32_t)morton3_invert((uint32_t)(v >> 2));
    }
}

/* :287-306 */
ORC_API void orc_packbits(const float *grid, uint32_t N, float thresh, uint8_t *bitfield) {
    for (uint32_t n = 0; n < N; ++n) {
        uint8_t b = 0;
        for (int i = 0; i < 8; ++i) b |= (grid[8 * (size_t)n + i] > thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = b;
    }
}

/* one ray of :358-533.  If xyzs == NULL only counts.  Returns the number of samples. */
typedef struct {
    const uint8_t *grid;
    float bound, dt_gamma, dt_min, dt_max, rH;
    uint32_t C, H, max_steps;
} march_cfg;

static inline int mip_clamp(int e, uint32_t C) { int hi = (int)C - 1; return e < 0 ? 0 : (e > hi ? hi : e); }

static uint32_t march_one(const march_cfg *k, const float *o, const float *d, float t0, float far, uint32_t cap,
                          float *xyzs, float *dirs, float *deltas) {
    const float rdx = 1 / d[0], rdy = 1 / d[1], rdz = 1 / d[2];
    const float rd[3] = {rdx, rdy, rdz};
    const uint32_t H = k->H, H3 = H * H * H;
    float t = t0, last_t = t0;
    uint32_t n = 0;
    while (t < far && n < cap) {
        float p[3];
        for (int a = 0; a < 3; ++a) p[a] = clampf(fmaf(t, d[a], o[a]), -k->bound, k->bound);
        const float dt = clampf(t * k->dt_gamma, k->dt_min, k->dt_max);
        int e_pos, e_dt;
        frexpf(fmaxf(fabsf(p[0]), fmaxf(fabsf(p[1]), fabsf(p[2]))), &e_pos);           /* :51-60 */
        frexpf((float)((double)(dt * (float)H) * 0.5), &e_dt);                         /* :62-69 */
        const int a_ = mip_clamp(e_pos, k->C), b_ = mip_clamp(e_dt, k->C);
        const int level = a_ > b_ ? a_ : b_;
        const float mip_bound = fminf(scalbnf(1.0f, level), k->bound);
        const float mip_rbound = 1 / mip_bound;
        int cell[3];
        for (int a = 0; a < 3; ++a) {                                                    /* :400-405 */
            const float v = (float)(0.5 * (double)fmaf(p[a], mip_rbound, 1.0f) * (double)H);
            cell[a] = (int)clampf(v, 0.0f, (float)(H - 1));
        }
        const uint32_t index = (uint32_t)level * H3 + morton3((uint32_t)cell[0], (uint32_t)cell[1], (uint32_t)cell[2]);
        const int occ = k->grid[index / 8] & (1 << (index % 8));
        if (occ) {
            const float t_new = t + dt;
            if (xyzs) {
                for (int a = 0; a < 3; ++a) { xyzs[3 * n + a] = p[a]; dirs[3 * n + a] = d[a]; }
                deltas[2 * n] = dt;
                deltas[2 * n + 1] = t_new - last_t;
            }
            last_t = t = t_new;
            ++n;
        } else {                                                                        /* :420-437 */
            float tmin = INFINITY;
            for (int a = 0; a < 3; ++a) {
                const float s = copysignf(1.0f, d[a]);
                const float c0 = fmaf(0.5f, s, (float)cell[a] + 0.5f);
                const float c1 = fmaf(c0 * k->rH, 2.0f, -1.0f);
                const float ta = fmaf(c1, mip_bound, -p[a]) * rd[a];
                tmin = fminf(tmin, ta);           /* fminf(tx, fminf(ty, tz)): order-independent incl. NaN */
            }
            const float tt = t + fmaxf(0.0f, tmin);
            do { t += clampf(t * k->dt_gamma, k->dt_min, k->dt_max); } while (t < tt);
        }
    }
    return n;
}

static march_cfg make_cfg(const uint8_t *grid, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
    march_cfg k;
    k.grid = grid; k.bound = bound; k.dt_gamma = dt_gamma; k.C = C; k.H = H; k.max_steps = max_steps;
    const float two_sqrt3 = 2 * 1.7320508075688772f;
    k.dt_min = two_sqrt3 / (float)max_steps;
    k.dt_max = two_sqrt3 * (float)(1 << (C - 1)) / (float)H;
    k.rH = 1 / (float)H;
    return k;
}

/* :332-568.  Deterministic arrival order: ray n takes slot n, samples are laid out in ray order. */
ORC_API void orc_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                                  float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                  const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                                  int32_t *rays, int32_t *counter, const float *noises) {
    const march_cfg k = make_cfg(grid, bound, dt_gamma, max_steps, C, H);
    uint32_t *counts = (uint32_t *)malloc(sizeof(uint32_t) * (N ? N : 1));
    float *t0s = (float *)malloc(sizeof(float) * (N ? N : 1));
#pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t n = 0; n < N; ++n) {
        float t0 = nears[n];
        t0 = fmaf(clampf(t0 * dt_gamma, k.dt_min, k.dt_max), noises[n], t0);              /* :375 */
        t0s[n] = t0;
        counts[n] = march_one(&k, rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, t0, fars[n], max_steps, NULL, NULL, NULL);
    }
    uint32_t *offs = (uint32_t *)malloc(sizeof(uint32_t) * (N ? N : 1));
    uint32_t run = (uint32_t)counter[0], slot0 = (uint32_t)counter[1];
    for (uint32_t n = 0; n < N; ++n) { offs[n] = run; run += counts[n]; }
    counter[0] = (int32_t)run;
    counter[1] = (int32_t)(slot0 + N);
#pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t n = 0; n < N; ++n) {
        int32_t *r = rays + 3 * (size_t)(slot0 + n);
        r[0] = (int32_t)n; r[1] = (int32_t)offs[n]; r[2] = (int32_t)counts[n];
        if (counts[n] == 0 || offs[n] + counts[n] > M) continue;                          /* :456-457 */
        march_one(&k, rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, t0s[n], fars[n], counts[n],
                  xyzs + 3 * (size_t)offs[n], dirs + 3 * (size_t)offs[n], deltas + 2 * (size_t)offs[n]);
    }
    free(counts); free(t0s); free(offs);
}

/* :809-928 */
ORC_API void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t,
                            const float *rays_o, const float *rays_d, float bound, float dt_gamma, uint32_t max_steps,
                            uint32_t C, uint32_t H, const uint8_t *grid, const float *nears, const float *fars,
                            float *xyzs, float *dirs, float *deltas, const float *noises) {
    (void)nears;
    const march_cfg k = make_cfg(grid, bound, dt_gamma, max_steps, C, H);
#pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t n = 0; n < n_alive; ++n) {
        const int32_t index = rays_alive[n];
        float t = rays_t[index];
        t = fmaf(clampf(t * dt_gamma, k.dt_min, k.dt_max), noises[n], t);                 /* :856 */
        const size_t base = (size_t)n * n_step;
        march_one(&k, rays_o + 3 * (size_t)index, rays_d + 3 * (size_t)index, t, fars[index], n_step,
                  xyzs + 3 * base, dirs + 3 * base, deltas + 2 * base);
    }
}

/* :578-655 generalised to `ch` colour channels (ch == 3 is the reference) */
ORC_API void orc_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas,
                                              const int32_t *rays, uint32_t M, uint32_t N, float T_thresh, uint32_t ch,
                                              float *weights_sum, float *depth, float *image) {
#pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t index = (uint32_t)rays[3 * n], offset = (uint32_t)rays[3 * n + 1], count = (uint32_t)rays[3 * n + 2];
        float acc[4] = {0, 0, 0, 0}, ws = 0, t = 0, d = 0, T = 1.0f;
        if (count != 0 && offset + count <= M) {
            for (uint32_t s = offset; s < offset + count; ++s) {
                const float alpha = 1.0f - expf(-sigmas[s] * deltas[2 * s]);
                const float w = alpha * T;
                for (uint32_t c = 0; c < ch; ++c) acc[c] += w * rgbs[(size_t)s * ch + c];
                t += deltas[2 * s + 1];
                d += w * t;
                ws += w;
                T *= 1.0f - alpha;
                if (T < T_thresh) break;
            }
        }
        weights_sum[index] = ws;
        depth[index] = d;
        for (uint32_t c = 0; c < ch; ++c) image[(size_t)index * ch + c] = acc[c];
    }
}

/* :691-772 (+ the depth term of SURVEY.md H1 when grad_depth != NULL) */
ORC_API void orc_composite_rays_train_backward(const float *g_ws, const float *g_depth, const float *g_img,
                                               const float *sigmas, const float *rgbs, const float *deltas,
                                               const int32_t *rays, const float *weights_sum, const float *depth,
                                               const float *image, uint32_t M, uint32_t N, float T_thresh, uint32_t ch,
                                               float *grad_sigmas, float *grad_rgbs) {
#pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t index = (uint32_t)rays[3 * n], offset = (uint32_t)rays[3 * n + 1], count = (uint32_t)rays[3 * n + 2];
        if (count == 0 || offset + count > M) continue;
        float acc[4] = {0, 0, 0, 0}, T = 1.0f, t = 0, d = 0;
        const float ws_final = weights_sum[index];
        for (uint32_t s = offset; s < offset + count; ++s) {
            const float alpha = 1.0f - expf(-sigmas[s] * deltas[2 * s]);
            const float w = alpha * T;
            for (uint32_t c = 0; c < ch; ++c) acc[c] += w * rgbs[(size_t)s * ch + c];
            t += deltas[2 * s + 1];
            d += w * t;
            T *= 1.0f - alpha;
            float g = 0;
            for (uint32_t c = 0; c < ch; ++c) {
                grad_rgbs[(size_t)s * ch + c] = g_img[(size_t)index * ch + c] * w;
                g += g_img[(size_t)index * ch + c] * (T * rgbs[(size_t)s * ch + c] - (image[(size_t)index * ch + c] - acc[c]));
            }
            g += g_ws[index] * (1 - ws_final);
            if (g_depth) g += g_depth[index] * (T * t - (depth[index] - d));
            grad_sigmas[s] = deltas[2 * s] * g;
            if (T < T_thresh) break;
        }
    }
}

/* :967-1053 */
ORC_API void orc_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive, float *rays_t,
                                const float *sigmas, const float *rgbs, const float *deltas, float *weights_sum,
                                float *depth, float *image) {
    for (uint32_t n = 0; n < n_alive; ++n) {
        const int32_t index = rays_alive[n];
        const size_t base = (size_t)n * n_step;
        float t = rays_t[index], ws = weights_sum[index], d = depth[index];
        float r = image[3 * (size_t)index], g = image[3 * (size_t)index + 1], b = image[3 * (size_t)index + 2];
        uint32_t step = 0;
        while (step < n_step) {
            const size_t s = base + step;
            if (deltas[2 * s] == 0) break;
            const float alpha = 1.0f - expf(-sigmas[s] * deltas[2 * s]);
            const float T = 1 - ws, w = alpha * T;
            ws += w;
            t += deltas[2 * s + 1];
            d += w * t;
            r += w * rgbs[3 * s]; g += w * rgbs[3 * s + 1]; b += w * rgbs[3 * s + 2];
            if (T < T_thresh) break;
            ++step;
        }
        if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
        weights_sum[index] = ws; depth[index] = d;
        image[3 * (size_t)index] = r; image[3 * (size_t)index + 1] = g; image[3 * (size_t)index + 2] = b;
    }
}

/* ------------------------------------------------------------------------------------------------
 * gridencoder/src/gridencoder.cu   (tables are passed as fp32; `half_mode` rounds every value the
 * reference would hold in at::Half)
 * ---------------------------------------------------------------------------------------------- */
static const uint32_t kPrimes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};

/* :69-93 */
static inline uint32_t grid_row(const uint32_t *p, uint32_t D, uint32_t gridtype, int align_corners,
                                uint32_t hashmap_size, uint32_t resolution) {
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; ++d) {
        index += p[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) {
        index = 0;
        for (uint32_t d = 0; d < D; ++d) index ^= p[d] * kPrimes[d];
    }
    return index % hashmap_size;
}

static inline float rh(float v, int half_mode) { return half_mode ? to_half_precision(v) : v; }

/* :95-263.  layout 0: outputs [L,B,C]; 1: [B,L*C].  dy_dx [B,L,D,C] or NULL.
 * level_scales (optional, [L]): the per-level `scale = exp2f(l*S)*H - 1` as evaluated ON THE DEVICE.  CUDA's
 * exp2f is not correctly rounded (2 ulp) and differs from glibc's by 1 ulp on some levels; one ulp of a
 * scale of ~2000 moves the interpolation weights by 1e-4, so bit-level comparisons pass the device values in. */
ORC_API void orc_grid_encode_forward(const float *inputs, const float *table, const int32_t *offsets, float *outputs,
                                     uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, float *dy_dx,
                                     uint32_t gridtype, int align_corners, uint32_t interp, int half_mode, int layout,
                                     const float *level_scales) {
#pragma omp parallel for schedule(static)
    for (uint32_t b = 0; b < B; ++b) {
        const float *x = inputs + (size_t)b * D;
        int oob = 0;
        for (uint32_t d = 0; d < D; ++d) if (x[d] < 0 || x[d] > 1) oob = 1;
        for (uint32_t l = 0; l < L; ++l) {
            float *out = layout == 0 ? outputs + ((size_t)l * B + b) * C : outputs + ((size_t)b * L + l) * C;
            float *dd = dy_dx ? dy_dx + ((size_t)b * L + l) * D * C : NULL;
            if (oob) {
                for (uint32_t c = 0; c < C; ++c) out[c] = 0;
                if (dd) for (uint32_t i = 0; i < D * C; ++i) dd[i] = 0;
                continue;
            }
            const float *tab = table + (size_t)offsets[l] * C;
            const uint32_t hsize = (uint32_t)(offsets[l + 1] - offsets[l]);
            const float scale = level_scales ? level_scales[l] : fmaf(exp2f((float)l * S), (float)H, -1.0f);
            const uint32_t res = (uint32_t)ceilf(scale) + 1;
            float pos[8], dpos[8];
            uint32_t pg[8];
            for (uint32_t d = 0; d < D; ++d) {
                float p = fmaf(x[d], scale, align_corners ? 0.0f : 0.5f);
                pg[d] = (uint32_t)floorf(p);
                p -= (float)pg[d];
                if (interp == 1) { dpos[d] = 6 * p * (1.0f - p); p = p * p * (3.0f - 2.0f * p); } else dpos[d] = 1.0f;
                pos[d] = p;
            }
            float res_c[8];
            for (uint32_t c = 0; c < C; ++c) res_c[c] = 0;
            for (uint32_t corner = 0; corner < (1u << D); ++corner) {
                float w = 1;
                uint32_t pl[8];
                for (uint32_t d = 0; d < D; ++d) {
                    if ((corner & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; } else { w *= pos[d]; pl[d] = pg[d] + 1; }
                }
                const uint32_t row = grid_row(pl, D, gridtype, align_corners, hsize, res);
                for (uint32_t c = 0; c < C; ++c) res_c[c] = rh(fmaf(w, tab[(size_t)row * C + c], res_c[c]), half_mode);
            }
            for (uint32_t c = 0; c < C; ++c) out[c] = res_c[c];
            if (dd) {
                for (uint32_t gd = 0; gd < D; ++gd) {
                    float acc[8];
                    for (uint32_t c = 0; c < C; ++c) acc[c] = 0;
                    for (uint32_t corner = 0; corner < (1u << (D - 1)); ++corner) {
                        float w = scale;
                        uint32_t pl[8];
                        for (uint32_t nd = 0; nd < D - 1; ++nd) {
                            const uint32_t d = nd >= gd ? nd + 1 : nd;
                            if ((corner & (1u << nd)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; } else { w *= pos[d]; pl[d] = pg[d] + 1; }
                        }
                        pl[gd] = pg[gd];
                        const uint32_t r0 = grid_row(pl, D, gridtype, align_corners, hsize, res);
                        pl[gd] = pg[gd] + 1;
                        const uint32_t r1 = grid_row(pl, D, gridtype, align_corners, hsize, res);
                        for (uint32_t c = 0; c < C; ++c) {
                            const float diff = rh(tab[(size_t)r1 * C + c] - tab[(size_t)r0 * C + c], half_mode);
                            acc[c] = rh(acc[c] + w * diff * dpos[gd], half_mode);
                        }
                    }
                    for (uint32_t c = 0; c < C; ++c) dd[gd * C + c] = acc[c];
                }
            }
        }
    }
}

/* :265-390.  grad_table is accumulated in fp32 (+=); in half_mode each contribution is first rounded to
 * half as the reference does before its atomicAdd (the reference's half accumulation order is
 * non-deterministic, so the sum itself is kept in fp32 here).  grad_inputs [B,D] is written if dy_dx. */
ORC_API void orc_grid_encode_backward(const float *grad, const float *inputs, const int32_t *offsets, float *grad_table,
                                      uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                      const float *dy_dx, float *grad_inputs, uint32_t gridtype, int align_corners,
                                      uint32_t interp, int half_mode, int layout, const float *level_scales) {
    for (uint32_t l = 0; l < L; ++l) {
        float *gt = grad_table + (size_t)offsets[l] * C;
        const uint32_t hsize = (uint32_t)(offsets[l + 1] - offsets[l]);
        const float scale = level_scales ? level_scales[l] : fmaf(exp2f((float)l * S), (float)H, -1.0f);
        const uint32_t res = (uint32_t)ceilf(scale) + 1;
        for (uint32_t b = 0; b < B; ++b) {
            const float *x = inputs + (size_t)b * D;
            int oob = 0;
            for (uint32_t d = 0; d < D; ++d) if (x[d] < 0 || x[d] > 1) oob = 1;
            if (oob) continue;
            const float *g = layout == 0 ? grad + ((size_t)l * B + b) * C : grad + ((size_t)b * L + l) * C;
            float pos[8];
            uint32_t pg[8];
            for (uint32_t d = 0; d < D; ++d) {
                float p = fmaf(x[d], scale, align_corners ? 0.0f : 0.5f);
                pg[d] = (uint32_t)floorf(p);
                p -= (float)pg[d];
                if (interp == 1) p = p * p * (3.0f - 2.0f * p);
                pos[d] = p;
            }
            for (uint32_t corner = 0; corner < (1u << D); ++corner) {
                float w = 1;
                uint32_t pl[8];
                for (uint32_t d = 0; d < D; ++d) {
                    if ((corner & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; } else { w *= pos[d]; pl[d] = pg[d] + 1; }
                }
                const uint32_t row = grid_row(pl, D, gridtype, align_corners, hsize, res);
                for (uint32_t c = 0; c < C; ++c) gt[(size_t)row * C + c] += rh(w * g[c], half_mode);
            }
        }
    }
    if (dy_dx && grad_inputs) {
        for (uint32_t b = 0; b < B; ++b)
            for (uint32_t d = 0; d < D; ++d) {
                float acc = 0;
                for (uint32_t l = 0; l < L; ++l) {
                    const float *g = layout == 0 ? grad + ((size_t)l * B + b) * C : grad + ((size_t)b * L + l) * C;
                    for (uint32_t c = 0; c < C; ++c)
                        acc = rh(acc + rh(g[c] * dy_dx[(((size_t)b * L + l) * D + d) * C + c], half_mode), half_mode);
                }
                grad_inputs[(size_t)b * D + d] = acc;
            }
    }
}

/* ------------------------------------------------------------------------------------------------
 * freqencoder/src/freqencoder.cu :34-101
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_freq_encode_forward(const float *inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float *outputs) {
    const float half_pi = 3.141592653589793f / 2;
#pragma omp parallel for schedule(static)
    for (uint32_t b = 0; b < B; ++b)
        for (uint32_t c = 0; c < C; ++c) {
            float *o = outputs + (size_t)b * C + c;
            if (c < D) { *o = inputs[(size_t)b * D + c]; continue; }
            const uint32_t col = c / D - 1, d = c % D, freq = col / 2;
            *o = sinf(scalbnf(inputs[(size_t)b * D + d], (int)freq) + (float)(col % 2) * half_pi);
        }
}
ORC_API void orc_freq_encode_backward(const float *grad, const float *outputs, uint32_t B, uint32_t D, uint32_t deg,
                                      uint32_t C, float *grad_inputs) {
    for (uint32_t b = 0; b < B; ++b)
        for (uint32_t d = 0; d < D; ++d) {
            const float *g = grad + (size_t)b * C, *o = outputs + (size_t)b * C;
            float r = g[d];
            g += D; o += D;
            for (uint32_t f = 0; f < deg; ++f) {
                r += scalbnf(1.0f, (int)f) * (g[d] * o[D + d] - g[D + d] * o[d]);
                g += 2 * D; o += 2 * D;
            }
            grad_inputs[(size_t)b * D + d] = r;
        }
}

/* ------------------------------------------------------------------------------------------------
 * shencoder/src/shencoder.cu :31-858.  The reference hard-codes each real SH basis function as a
 * Cartesian polynomial; they equal N_l^m Q_l^m(z) {Re,Im}(x+iy)^|m| (see csrc/shencoder.cu header).
 * Evaluated here in double precision.  dy_dx layout [B,3,C*C].
 * ---------------------------------------------------------------------------------------------- */
static double fact(int n) { double r = 1; for (int i = 2; i <= n; ++i) r *= i; return r; }

ORC_API void orc_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t deg, float *dy_dx) {
    const uint32_t C2 = deg * deg;
    for (uint32_t b = 0; b < B; ++b) {
        const double x = inputs[3 * (size_t)b], y = inputs[3 * (size_t)b + 1], z = inputs[3 * (size_t)b + 2];
        double A[10], Bm[10], Q[9][10];
        A[0] = 1; Bm[0] = 0;
        for (uint32_t m = 1; m <= deg; ++m) { A[m] = x * A[m - 1] - y * Bm[m - 1]; Bm[m] = x * Bm[m - 1] + y * A[m - 1]; }
        memset(Q, 0, sizeof(Q));
        double df = 1;
        for (uint32_t m = 0; m < deg; ++m) {
            if (m > 0) df *= (2.0 * m - 1);
            Q[m][m] = df;
            if (m + 1 < deg) Q[m + 1][m] = (2.0 * m + 1) * z * df;
            for (uint32_t l = m + 2; l < deg; ++l) Q[l][m] = ((2.0 * l - 1) * z * Q[l - 1][m] - (double)(l + m - 1) * Q[l - 2][m]) / (double)(l - m);
        }
        float *o = outputs + (size_t)b * C2;
        float *gx = dy_dx ? dy_dx + (size_t)b * 3 * C2 : NULL, *gy = gx ? gx + C2 : NULL, *gz = gy ? gy + C2 : NULL;
        for (uint32_t l = 0; l < deg; ++l)
            for (uint32_t m = 0; m <= l; ++m) {
                double N = sqrt((2.0 * l + 1) / (4 * M_PI) * fact((int)(l - m)) / fact((int)(l + m)));
                if (m > 0) N *= sqrt(2.0) * ((m & 1) ? -1.0 : 1.0);
                const double nq = N * Q[l][m], nq1 = N * Q[l][m + 1];
                const uint32_t ip = l * l + l + m, im = l * l + l - m;
                o[ip] = (float)(nq * A[m]);
                if (m > 0) o[im] = (float)(nq * Bm[m]);
                if (gx) {
                    if (m == 0) { gx[ip] = 0; gy[ip] = 0; gz[ip] = (float)nq1; }
                    else {
                        gx[ip] = (float)(nq * m * A[m - 1]); gy[ip] = (float)(-nq * m * Bm[m - 1]); gz[ip] = (float)(nq1 * A[m]);
                        gx[im] = (float)(nq * m * Bm[m - 1]); gy[im] = (float)(nq * m * A[m - 1]); gz[im] = (float)(nq1 * Bm[m]);
                    }
                }
            }
    }
}
ORC_API void orc_sh_encode_backward(const float *grad, uint32_t B, uint32_t deg, const float *dy_dx, float *grad_inputs) {
    const uint32_t C2 = deg * deg;
    for (uint32_t b = 0; b < B; ++b)
        for (uint32_t d = 0; d < 3; ++d) {
            float acc = grad_inputs[3 * (size_t)b + d];
            for (uint32_t ch = 0; ch < C2; ++ch) acc += grad[(size_t)b * C2 + ch] * dy_dx[((size_t)b * 3 + d) * C2 + ch];
            grad_inputs[3 * (size_t)b + d] = acc;
        }
}

/* ------------------------------------------------------------------------------------------------
 * ffmlp/src/ffmlp.cu (:460-576 forward, :578-733 + :1059-1264 backward; layouts :861-864) and
 * ffmlp/ffmlp.py:223-283.  Values are fp32 arrays holding fp16-representable numbers; every tensor the
 * reference stores as half is rounded to half here; dot products are accumulated in fp32.
 * weights = [hidden*in | n_hid*hidden*hidden | out*hidden] row-major, ReLU, no output activation.
 * forward_buffer [num_layers, B, hidden]; backward_buffer [num_layers, B, hidden] (index 0 = last layer).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_ffmlp_forward(const float *inputs, const float *weights, uint32_t B, uint32_t in_dim, uint32_t out_dim,
                               uint32_t hidden, uint32_t num_layers, float *forward_buffer, float *outputs) {
    const float *w_in = weights, *w_hid = weights + (size_t)hidden * in_dim;
    const float *w_out = w_hid + (size_t)(num_layers - 1) * hidden * hidden;
#pragma omp parallel for schedule(static)
    for (uint32_t b = 0; b < B; ++b) {
        float cur[256], nxt[256];
        for (uint32_t j = 0; j < hidden; ++j) {
            float acc = 0;
            for (uint32_t i = 0; i < in_dim; ++i) acc += inputs[(size_t)b * in_dim + i] * w_in[(size_t)j * in_dim + i];
            cur[j] = mlp_round(fmaxf(acc, 0.f));
        }
        if (forward_buffer) memcpy(forward_buffer + ((size_t)0 * B + b) * hidden, cur, sizeof(float) * hidden);
        for (uint32_t l = 0; l + 1 < num_layers; ++l) {
            const float *w = w_hid + (size_t)l * hidden * hidden;
            for (uint32_t j = 0; j < hidden; ++j) {
                float acc = 0;
                for (uint32_t i = 0; i < hidden; ++i) acc += cur[i] * w[(size_t)j * hidden + i];
                nxt[j] = mlp_round(fmaxf(acc, 0.f));
            }
            memcpy(cur, nxt, sizeof(float) * hidden);
            if (forward_buffer) memcpy(forward_buffer + ((size_t)(l + 1) * B + b) * hidden, cur, sizeof(float) * hidden);
        }
        for (uint32_t o = 0; o < out_dim; ++o) {
            float acc = 0;
            for (uint32_t i = 0; i < hidden; ++i) acc += cur[i] * w_out[(size_t)o * hidden + i];
            outputs[(size_t)b * out_dim + o] = mlp_round(acc);
        }
    }
}

/* grad_weights is accumulated in fp32 and NOT rounded (callers round to half if they want the reference's
 * storage type); grad_inputs / backward_buffer are rounded to half. */
ORC_API void orc_ffmlp_backward(const float *grad, const float *inputs, const float *weights, const float *forward_buffer,
                                uint32_t B, uint32_t in_dim, uint32_t out_dim, uint32_t hidden, uint32_t num_layers,
                                float *backward_buffer, float *grad_inputs, float *grad_weights) {
    const float *w_in = weights, *w_hid = weights + (size_t)hidden * in_dim;
    const float *w_out = w_hid + (size_t)(num_layers - 1) * hidden * hidden;
    float *g_in = grad_weights, *g_hid = grad_weights + (size_t)hidden * in_dim;
    float *g_out = g_hid + (size_t)(num_layers - 1) * hidden * hidden;
    for (uint32_t b = 0; b < B; ++b) {
        float d_cur[256], d_nxt[256];
        const float *g = grad + (size_t)b * out_dim;
        const float *h_last = forward_buffer + ((size_t)(num_layers - 1) * B + b) * hidden;
        for (uint32_t o = 0; o < out_dim; ++o)
            for (uint32_t j = 0; j < hidden; ++j) g_out[(size_t)o * hidden + j] += g[o] * h_last[j];
        for (uint32_t j = 0; j < hidden; ++j) {
            float acc = 0;
            for (uint32_t o = 0; o < out_dim; ++o) acc += g[o] * w_out[(size_t)o * hidden + j];
            d_cur[j] = mlp_round(h_last[j] > 0 ? acc : 0.f);
        }
        if (backward_buffer) memcpy(backward_buffer + ((size_t)0 * B + b) * hidden, d_cur, sizeof(float) * hidden);
        for (int l = (int)num_layers - 2; l >= 0; --l) {   /* W_hid[l]: h_l -> h_{l+1} */
            const float *w = w_hid + (size_t)l * hidden * hidden;
            const float *h_prev = forward_buffer + ((size_t)l * B + b) * hidden;
            float *gw = g_hid + (size_t)l * hidden * hidden;
            for (uint32_t o = 0; o < hidden; ++o)
                for (uint32_t j = 0; j < hidden; ++j) gw[(size_t)o * hidden + j] += d_cur[o] * h_prev[j];
            for (uint32_t j = 0; j < hidden; ++j) {
                float acc = 0;
                for (uint32_t o = 0; o < hidden; ++o) acc += d_cur[o] * w[(size_t)o * hidden + j];
                d_nxt[j] = mlp_round(h_prev[j] > 0 ? acc : 0.f);
            }
            memcpy(d_cur, d_nxt, sizeof(float) * hidden);
            if (backward_buffer)
                memcpy(backward_buffer + ((size_t)(num_layers - 1 - l) * B + b) * hidden, d_cur, sizeof(float) * hidden);
        }
        const float *x = inputs + (size_t)b * in_dim;
        for (uint32_t o = 0; o < hidden; ++o)
            for (uint32_t i = 0; i < in_dim; ++i) g_in[(size_t)o * in_dim + i] += d_cur[o] * x[i];
        if (grad_inputs)
            for (uint32_t i = 0; i < in_dim; ++i) {
                float acc = 0;
                for (uint32_t o = 0; o < hidden; ++o) acc += d_cur[o] * w_in[(size_t)o * in_dim + i];
                grad_inputs[(size_t)b * in_dim + i] = mlp_round_dx(acc);
            }
    }
}

/* torch.optim.Adam step (the optimiser the reference constructs, main_lidarnerf.py:389-391) */
ORC_API void orc_adam_step(float *p, const float *g, float *m, float *v, size_t n, float lr, float b1, float b2,
                           float eps, float bc1, float bc2, float gscale) {
    const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        const float gi = g[i] * gscale;
        m[i] = b1 * m[i] + (1.f - b1) * gi;
        v[i] = b2 * v[i] + (1.f - b2) * gi * gi;
        p[i] -= (lr / bc1) * (m[i] / (sqrtf(v[i]) * inv_sqrt_bc2 + eps));
    }
}

/* round an fp32 array to fp16 precision in place (helper for the Python side) */
ORC_API void orc_round_to_half(float *a, size_t n) {
    for (size_t i = 0; i < n; ++i) a[i] = to_half_precision(a[i]);
}

/* ------------------------------------------------------------------------------------------------
 * gridencoder.cu:695-808  kernel_grad_tv (fp32 tables; half tables are a no-op in the reference because its
 * at::Half atomicAdd stub is empty, gridencoder.cu:27-31)
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_grad_total_variation(const float *inputs, const float *table, float *grad, const int32_t *offsets,
                                      float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                      uint32_t gridtype, int align_corners, const float *level_scales) {
    for (uint32_t l = 0; l < L; ++l) {
        const float *tab = table + (size_t)offsets[l] * C;
        float *gt = grad + (size_t)offsets[l] * C;
        const uint32_t hsize = (uint32_t)(offsets[l + 1] - offsets[l]);
        const float scale = level_scales ? level_scales[l] : fmaf(exp2f((float)l * S), (float)H, -1.0f);
        const uint32_t res = (uint32_t)ceilf(scale) + 1;
        const float w = weight / (float)(2 * D);                                     /* :757 */
        for (uint32_t b = 0; b < B; ++b) {
            const float *x = inputs + (size_t)b * D;
            int oob = 0;
            uint32_t pg[8];
            for (uint32_t d = 0; d < D; ++d) {
                if (x[d] < 0 || x[d] > 1) oob = 1;
                pg[d] = (uint32_t)floorf(fmaf(x[d], scale, align_corners ? 0.0f : 0.5f));   /* :739-741, nvcc contracts */
            }
            if (oob) continue;
            const uint32_t index = grid_row(pg, D, gridtype, align_corners, hsize, res) * C;
            for (uint32_t c = 0; c < C; ++c) {
                float results = 0.f, idelta = 0.f;
                for (uint32_t d = 0; d < D; ++d) {
                    const uint32_t cur = pg[d];
                    if (cur < res) {                                                   /* right neighbour :764-779 */
                        pg[d] = cur + 1;
                        const float gv = tab[index + c] - tab[grid_row(pg, D, gridtype, align_corners, hsize, res) * C + c];
                        results += gv;
                        idelta = fmaf(gv, gv, idelta);
                    }
                    if (cur > 0) {                                                     /* left neighbour :782-796 */
                        pg[d] = cur - 1;
                        const float gv = tab[index + c] - tab[grid_row(pg, D, gridtype, align_corners, hsize, res) * C + c];
                        results += gv;
                        idelta = fmaf(gv, gv, idelta);
                    }
                    pg[d] = cur;
                }
                gt[index + c] += w * results * (1.0f / sqrtf(idelta + 1e-9f));          /* :803-806 */
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * extern/chamfer3D/chamfer3D.cu:9-133  NmDistanceKernel: nearest neighbour of every xyz1 point in xyz2
 * (squared distance, first minimum in index order)
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_chamfer_nn(const float *xyz1, const float *xyz2, uint32_t B, uint32_t N, uint32_t M, float *dist,
                            int32_t *idx) {
#pragma omp parallel for schedule(static) collapse(2)
    for (uint32_t b = 0; b < B; ++b)
        for (uint32_t j = 0; j < N; ++j) {
            const float *p = xyz1 + ((size_t)b * N + j) * 3;
            float best = 0.f;
            int32_t best_i = 0;
            for (uint32_t k = 0; k < M; ++k) {
                const float *q = xyz2 + ((size_t)b * M + k) * 3;
                const float x = q[0] - p[0], y = q[1] - p[1], z = q[2] - p[2];
                /* x*x + y*y + z*z as nvcc 12.9 contracts it in the reference build (SASS of NmDistanceKernel:
                 * FMUL y,y ; FFMA x,x,+ ; FFMA z,z,+): the y product is the one rounded on its own */
                const float d = fmaf(z, z, fmaf(x, x, y * y));
                if (k == 0 || d < best) { best = d; best_i = (int32_t)k; }
            }
            dist[(size_t)b * N + j] = best;
            idx[(size_t)b * N + j] = best_i;
        }
}

/* ------------------------------------------------------------------------------------------------
 * lidarnerf/convert.py:99-160  lidar_to_pano_with_intensities: spherical projection with a z-buffer (the closest
 * point of a pixel wins, the first one on ties).  fp32 arithmetic like numpy >= 2 evaluates the reference's
 * expressions on float32 scalars; libm atan2f/sqrtf.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_lidar_to_pano(const float *points, uint32_t stride, uint32_t N, uint32_t H, uint32_t W, float fov_up,
                               float fov, float max_depth, float *pano, float *intensities) {
    const float pi = 3.14159265358979323846f;
    const float fov_down = fov - fov_up;
    const float down = (float)((double)fov_down / 180.0 * 3.14159265358979323846);
    const float col_step = (float)(2.0 * 3.14159265358979323846 / (double)W);
    const float row_step = (float)((double)fov / 180.0 * 3.14159265358979323846 / (double)H);
    for (size_t i = 0; i < (size_t)H * W; ++i) pano[i] = 0.f, intensities[i] = 0.f;
    for (uint32_t n = 0; n < N; ++n) {
        const float *p = points + (size_t)n * stride;
        const float x = p[0], y = p[1], z = p[2];
        const float dist = sqrtf(x * x + y * y + z * z);
        if (dist >= max_depth) continue;
        const float beta = pi - atan2f(y, x);
        const float alpha = atan2f(z, sqrtf(x * x + y * y)) + down;
        const long c = lrintf(beta / col_step);                    /* round half to even, like Python's round() */
        const long r = lrintf((float)H - alpha / row_step);
        if (r >= (long)H || r < 0 || c >= (long)W || c < 0) continue;
        float *px = pano + (size_t)r * W + c;
        if (*px == 0.0f || *px > dist) {
            *px = dist;
            intensities[(size_t)r * W + c] = stride > 3 ? p[3] : 0.f;
        }
    }
}

/* lidarnerf/convert.py:194-235  pano_to_lidar_with_intensities: every non-empty pixel -> a point along its beam
 * direction, in row-major pixel order.  Returns the number of points written to out [H*W,4]. */
ORC_API uint32_t orc_pano_to_lidar(const float *pano, const float *intensities, uint32_t H, uint32_t W, float fov_up,
                                   float fov, float *out) {
    uint32_t n = 0;
    const float pi = 3.14159265358979323846f;
    for (uint32_t j = 0; j < H; ++j)
        for (uint32_t i = 0; i < W; ++i) {
            const float d = pano[(size_t)j * W + i];
            if (d == 0.0f) continue;
            /* numpy: float32 arrays combined with Python scalars stay float32 (evaluation order as written) */
            const float beta = -((float)i - (float)W / 2.0f) / (float)W * 2.0f * pi;
            const float alpha = (fov_up - (float)j / (float)H * fov) / 180.0f * pi;
            out[(size_t)n * 4 + 0] = cosf(alpha) * cosf(beta) * d;
            out[(size_t)n * 4 + 1] = cosf(alpha) * sinf(beta) * d;
            out[(size_t)n * 4 + 2] = sinf(alpha) * d;
            out[(size_t)n * 4 + 3] = intensities ? intensities[(size_t)j * W + i] : 0.f;
            ++n;
        }
    return n;
}
