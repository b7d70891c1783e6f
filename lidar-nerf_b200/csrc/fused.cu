// Glue kernels of the LiDAR field training step (the work the reference leaves to ~40 small torch kernels
// between its extension calls: nerf/network.py:162-237, nerf/utils.py:716-734, activation.py:6-20).
// Each is a single streaming pass over the marched samples (or the rays), vectorised to 16-byte accesses.
#include "common.cuh"

namespace lnb {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// Padding samples past the produced count are zero-filled on the device (no host sync, no full memset):
// xyzs/dirs -> 0, deltas -> 0 (alpha = 0, so they never contribute; raymarching.py:235-237 zero-fills whole
// buffers on the host side every call).
__global__ void __launch_bounds__(kThreads)
k_zero_tail(float *__restrict__ xyzs, float *__restrict__ dirs, float *__restrict__ deltas,
            int32_t *__restrict__ ray_ids, const int32_t *__restrict__ counter, uint32_t M) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= active_rows(M, counter)) return;      // only up to the end of the last partially filled 128-row tile
    const uint32_t used = (uint32_t)max(counter[0], 0);
    if (i < used) return;
    xyzs[i * 3] = xyzs[i * 3 + 1] = xyzs[i * 3 + 2] = 0.f;
    if (dirs) dirs[i * 3] = dirs[i * 3 + 1] = dirs[i * 3 + 2] = 0.f;
    if (ray_ids) ray_ids[i] = 0;
    reinterpret_cast<float2 *>(deltas)[i] = make_float2(0.f, 0.f);
}

// sigma head output -> (sigma, LiDAR-head input row).
//   sigma    = exp(h0) * density_scale                       (network.py:173 trunc_exp; renderer sigmas*scale)
//   head_in  = [ freq_enc(dir) (3 + 6*deg) | geo_feat (15) | zero pad ]      (network.py:215-216)
// One thread builds one row in shared memory (no per-element div/mod: the band loop is the outer structure),
// then the CTA streams the 128-row tile out with fully coalesced 16-byte stores.
constexpr int kHeadRows = 128;
__global__ void __launch_bounds__(kHeadRows)
k_head_input(const __half *__restrict__ sigma_out, const float *__restrict__ dirs, uint32_t M, uint32_t deg,
             uint32_t in_pad, float density_scale, float *__restrict__ sigma, __half *__restrict__ head_in,
             const int32_t *__restrict__ n_active) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    M = active_rows(M, n_active);
    if (blockIdx.x * kHeadRows >= M) return;
    __half *tile = reinterpret_cast<__half *>(smem_raw);          // [kHeadRows][in_pad + 8] (padded rows)
    const uint32_t pitch = in_pad + 8;
    const uint32_t s0 = blockIdx.x * kHeadRows;
    const uint32_t s = s0 + threadIdx.x;
    if (s < M) {
        __half *row = tile + threadIdx.x * pitch;
        const float d[3] = {dirs[s * 3], dirs[s * 3 + 1], dirs[s * 3 + 2]};
        const float half_pi = 3.141592653589793f / 2;
#pragma unroll
        for (int a = 0; a < 3; ++a) row[a] = __float2half_rn(d[a]);
        uint32_t col = 3;
        for (uint32_t f = 0; f < deg; ++f) {                      // freqencoder.cu:57-61: [sin(2^f x), cos(2^f x)]
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float arg = scalbnf(d[a], (int)f);
                row[col + a] = __float2half_rn(__sinf(arg + 0.f * half_pi));
                row[col + 3 + a] = __float2half_rn(__sinf(arg + 1.f * half_pi));
            }
            col += 6;
        }
        const uint4 *so = reinterpret_cast<const uint4 *>(sigma_out + (size_t)s * 16);
        const uint4 lo = __ldg(so), hi = __ldg(so + 1);
        const __half *h = reinterpret_cast<const __half *>(&lo);
        const __half *h2 = reinterpret_cast<const __half *>(&hi);
#pragma unroll
        for (int j = 1; j < 8; ++j) row[col + j - 1] = h[j];
#pragma unroll
        for (int j = 0; j < 8; ++j) row[col + 7 + j] = h2[j];
        for (uint32_t c = col + 15; c < in_pad; ++c) row[c] = __float2half_rn(0.f);
        sigma[s] = __expf(__half2float(h[0])) * density_scale;
    }
    __syncthreads();
    const uint32_t rows = min((uint32_t)kHeadRows, M - s0);
    const uint32_t cpr = in_pad / 8;                               // 16-byte chunks per row
    uint4 *out = reinterpret_cast<uint4 *>(head_in + (size_t)s0 * in_pad);
    for (uint32_t q = threadIdx.x; q < rows * cpr; q += kHeadRows) {
        const uint32_t r = q / cpr, c = q - r * cpr;
        out[q] = *reinterpret_cast<const uint4 *>(tile + r * pitch + c * 8);
    }
}

// LiDAR head output -> (ray-drop, intensity) = sigmoid(h[0:2])   (network.py:230)
__global__ void __launch_bounds__(kThreads)
k_head_rgb(const __half *__restrict__ head_out, uint32_t M, float *__restrict__ rgb,
           const int32_t *__restrict__ n_active) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= active_rows(M, n_active)) return;
    const __half2 h = *reinterpret_cast<const __half2 *>(head_out + (size_t)s * 16);
    const float2 f = __half22float2(h);
    reinterpret_cast<float2 *>(rgb)[s] = make_float2(1.f / (1.f + __expf(-f.x)), 1.f / (1.f + __expf(-f.y)));
}

// Per-ray LiDAR loss and its gradient w.r.t. the compositing outputs (nerf/utils.py:707-734):
//   loss = mean_n [ a_d |D m - d_gt m| + a_r (p_drop - m)^2 + a_i (p_int m - i_gt m)^2 ],   m = gt ray-drop
// with the absolute depth D = depth + t0 * weights_sum (t0 = perturbed march start).
// gt rows are (ray-drop, intensity, depth) (kitti360_dataset.py:148-157).  Gradients carry `loss_scale`.
__global__ void __launch_bounds__(kThreads)
k_lidar_loss(const float *__restrict__ ws, const float *__restrict__ depth, const float *__restrict__ image,
             const float *__restrict__ gt, const float *__restrict__ t0, uint32_t N, float a_d, float a_r,
             float a_i, float loss_scale, float *__restrict__ g_ws, float *__restrict__ g_depth,
             float *__restrict__ g_image, float *__restrict__ loss_out) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (n < N) {
        const float m = gt[n * 3], gi = gt[n * 3 + 1] * m, gd = gt[n * 3 + 2] * m;
        const float start = t0 ? t0[n] : 0.f;
        const float D = depth[n] + start * ws[n];
        const float e_d = D * m - gd;
        const float e_r = image[n * 2] - m;
        const float e_i = image[n * 2 + 1] * m - gi;
        const float inv_n = 1.f / (float)N;
        l = (a_d * fabsf(e_d) + a_r * e_r * e_r + a_i * e_i * e_i) * inv_n;
        const float s = loss_scale * inv_n;
        const float gD = a_d * m * (e_d > 0.f ? 1.f : (e_d < 0.f ? -1.f : 0.f)) * s;
        g_depth[n] = gD;
        g_ws[n] = gD * start;
        g_image[n * 2] = 2.f * a_r * e_r * s;
        g_image[n * 2 + 1] = 2.f * a_i * e_i * m * s;
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0 && l != 0.f) atomicAdd(loss_out, l);
}

// The same loss plus the patch depth-gradient term of the KITTI-360 configurations (`grad_loss = True`,
// `change_patch_size_lidar = [2, 8]`; nerf/utils.py:748-876, non-sobel branch, depth_grad_loss = l1).  Rays arrive as
// patches [np, px, py] (base_dataset.py:52-74: px rows x py columns of the range image); with P = D m / scale and
// G = d_gt m / scale (metres) the reference adds
//     alpha_grad * mean_{patch, i, j < py-1} | |P_ij - P_i,j+1| mask - (G_ij - G_i,j+1) mask |,
//     mask = m_ij * [ |G_ij - G_i,j+1| < grad_clip ]                                           (:846-866)
// - the horizontal (x) differences only (the y differences are computed at :795-800 but enter no enabled term), the
// prediction's difference in ABSOLUTE value against the SIGNED ground-truth difference, exactly as written there.
// Thread n owns ray n and evaluates the (at most two) pairs it belongs to, so no atomics on the gradients.
__device__ __forceinline__ float sgn(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

__global__ void __launch_bounds__(kThreads)
k_lidar_loss_patch(const float *__restrict__ ws, const float *__restrict__ depth, const float *__restrict__ image,
                   const float *__restrict__ gt, const float *__restrict__ t0, uint32_t N, float a_d, float a_r,
                   float a_i, float loss_scale, uint32_t px, uint32_t py, float a_grad, float inv_scale, float clip,
                   float *__restrict__ g_ws, float *__restrict__ g_depth, float *__restrict__ g_image,
                   float *__restrict__ loss_out) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (n < N) {
        const float m = gt[n * 3], gi = gt[n * 3 + 1] * m, gd = gt[n * 3 + 2] * m;
        const float start = t0 ? t0[n] : 0.f;
        const float D = depth[n] + start * ws[n];
        const float e_d = D * m - gd;
        const float e_r = image[n * 2] - m;
        const float e_i = image[n * 2 + 1] * m - gi;
        const float inv_n = 1.f / (float)N;
        l = (a_d * fabsf(e_d) + a_r * e_r * e_r + a_i * e_i * e_i) * inv_n;
        const float s = loss_scale * inv_n;
        float gD = a_d * m * sgn(e_d) * s;
        // ---- patch term ----
        const uint32_t n_patch = N / (px * py);
        if (py > 1 && n < n_patch * px * py) {
            const float inv_pairs = 1.f / (float)(n_patch * px * (py - 1));
            const uint32_t j = n % py;
            const float P = D * m * inv_scale, G = gd * inv_scale;
            auto other = [&](uint32_t q, float &Pq, float &Gq, float &mq) {
                mq = gt[q * 3];
                const float sq = t0 ? t0[q] : 0.f;
                Pq = (depth[q] + sq * ws[q]) * mq * inv_scale;
                Gq = gt[q * 3 + 2] * mq * inv_scale;
            };
            if (j + 1 < py) {              // pair (n, n + 1): this ray is the left element, the mask is its own
                float P1, G1, m1;
                other(n + 1, P1, G1, m1);
                const float dg = G - G1, dp = fabsf(P - P1);
                const float msk = m * (fabsf(dg) < clip ? 1.f : 0.f);
                const float e = dp * msk - dg * msk;
                l += a_grad * fabsf(e) * inv_pairs;
                gD += a_grad * inv_pairs * loss_scale * sgn(e) * msk * sgn(P - P1) * m * inv_scale;
            }
            if (j > 0) {                   // pair (n - 1, n): right element, the mask belongs to the left ray
                float P0, G0, m0;
                other(n - 1, P0, G0, m0);
                const float dg = G0 - G, dp = fabsf(P0 - P);
                const float msk = m0 * (fabsf(dg) < clip ? 1.f : 0.f);
                const float e = dp * msk - dg * msk;
                gD -= a_grad * inv_pairs * loss_scale * sgn(e) * msk * sgn(P0 - P) * m * inv_scale;
            }
        }
        g_depth[n] = gD;
        g_ws[n] = gD * start;
        g_image[n * 2] = 2.f * a_r * e_r * s;
        g_image[n * 2 + 1] = 2.f * a_i * e_i * m * s;
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0 && l != 0.f) atomicAdd(loss_out, l);
}

// d loss / d head_out = [ g_rgb * s (1 - s), 0 ... 0 ]  (fp16 row of 16)
__global__ void __launch_bounds__(kThreads)
k_head_out_grad(const float *__restrict__ g_rgb, const float *__restrict__ rgb, uint32_t M,
                __half *__restrict__ g_head_out, const int32_t *__restrict__ n_active) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= active_rows(M, n_active)) return;
    const float2 g = reinterpret_cast<const float2 *>(g_rgb)[s];
    const float2 p = reinterpret_cast<const float2 *>(rgb)[s];
    uint4 lo = make_uint4(pack2(g.x * p.x * (1.f - p.x), g.y * p.y * (1.f - p.y)), 0, 0, 0);
    uint4 *dst = reinterpret_cast<uint4 *>(g_head_out + (size_t)s * 16);
    dst[0] = lo;
    dst[1] = make_uint4(0, 0, 0, 0);
}

// d loss / d sigma_out = [ g_sigma * density_scale * exp(clamp(h0,-15,15)), d geo_feat (from the head's input grad) ]
__global__ void __launch_bounds__(kThreads)
k_sigma_out_grad(const float *__restrict__ g_sigma, const __half *__restrict__ sigma_out,
                 const __half *__restrict__ g_head_in, uint32_t M, uint32_t in_pad, uint32_t enc_dim,
                 float density_scale, __half *__restrict__ g_sigma_out, const int32_t *__restrict__ n_active) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= active_rows(M, n_active)) return;
    const float h0 = __half2float(sigma_out[(size_t)s * 16]);
    const float g0 = g_sigma[s] * density_scale * __expf(fminf(fmaxf(h0, -15.f), 15.f));   // activation.py:14-17
    const __half *gi = g_head_in + (size_t)s * in_pad + enc_dim;
    float v[16];
    v[0] = g0;
#pragma unroll
    for (int j = 0; j < 15; ++j) v[j + 1] = __half2float(gi[j]);
    uint4 *dst = reinterpret_cast<uint4 *>(g_sigma_out + (size_t)s * 16);
    dst[0] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
    dst[1] = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
}

// LiDAR ray generation (dataset/base_dataset.py:85-100): pixel (row j, col i) of an H x W range image ->
// direction in the sensor frame, rotated by the pose; origin = pose translation.
__global__ void __launch_bounds__(kThreads)
k_lidar_rays(const float *__restrict__ pose, const int32_t *__restrict__ inds, uint32_t N, uint32_t H, uint32_t W,
             float fov_up, float fov, float *__restrict__ rays_o, float *__restrict__ rays_d) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t idx = (uint32_t)inds[n];
    const float j = (float)(idx / W), i = (float)(idx % W);
    const float pi = 3.14159265358979323846f;
    const float beta = -(i - (float)W / 2) / (float)W * 2 * pi;
    const float alpha = (fov_up - j / (float)H * fov) / 180 * pi;
    const float dir[3] = {cosf(alpha) * cosf(beta), cosf(alpha) * sinf(beta), sinf(alpha)};
#pragma unroll
    for (int r = 0; r < 3; ++r) {   // rays_d = dir @ R^T  -> d_r = sum_c R[r][c] dir[c]
        rays_d[n * 3 + r] = pose[r * 4 + 0] * dir[0] + pose[r * 4 + 1] * dir[1] + pose[r * 4 + 2] * dir[2];
        rays_o[n * 3 + r] = pose[r * 4 + 3];
    }
}

// One training batch from a frame that lives on the device (what the reference's collate does with torch ops per step,
// kitti360_dataset.py:123-159): rays of the sampled pixels (get_lidar_rays) + their ground-truth rows gathered from the
// frame's range image [H*W, 3] = (ray-drop mask, intensity, depth * scale).
__global__ void __launch_bounds__(kThreads)
k_lidar_batch(const float *__restrict__ pose, const int32_t *__restrict__ inds, const float *__restrict__ image, uint32_t N,
              uint32_t H, uint32_t W, float fov_up, float fov, float *__restrict__ rays_o, float *__restrict__ rays_d,
              float *__restrict__ gt) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t idx = (uint32_t)inds[n];
    const float j = (float)(idx / W), i = (float)(idx % W);
    const float pi = 3.14159265358979323846f;
    const float beta = -(i - (float)W / 2) / (float)W * 2 * pi;
    const float alpha = (fov_up - j / (float)H * fov) / 180 * pi;
    const float dir[3] = {cosf(alpha) * cosf(beta), cosf(alpha) * sinf(beta), sinf(alpha)};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        rays_d[n * 3 + r] = pose[r * 4 + 0] * dir[0] + pose[r * 4 + 1] * dir[1] + pose[r * 4 + 2] * dir[2];
        rays_o[n * 3 + r] = pose[r * 4 + 3];
        gt[n * 3 + r] = __ldg(image + (size_t)idx * 3 + r);
    }
}

inline unsigned nblk(uint64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace
}  // namespace lnb

using namespace lnb;

extern "C" {

int lnb_zero_sample_tail_ex(float *xyzs, float *dirs, float *deltas, int32_t *ray_ids, const int32_t *counter,
                            uint32_t M, lnb_stream_t stream) {
    if (!xyzs || !deltas || !counter) return LNB_ERR_INVALID_ARGUMENT;
    if (M == 0) return LNB_OK;
    k_zero_tail<<<nblk(M), kThreads, 0, as_stream(stream)>>>(xyzs, dirs, deltas, ray_ids, counter, M);
    count_launch();
    return launch_status();
}

int lnb_zero_sample_tail(float *xyzs, float *dirs, float *deltas, const int32_t *counter, uint32_t M,
                         lnb_stream_t stream) {
    if (!dirs) return LNB_ERR_INVALID_ARGUMENT;
    return lnb_zero_sample_tail_ex(xyzs, dirs, deltas, nullptr, counter, M, stream);
}

int lnb_field_head_input(const void *sigma_out, const float *dirs, uint32_t M, uint32_t degree, uint32_t in_pad,
                         float density_scale, float *sigma, void *head_in, const int32_t *n_active,
                         lnb_stream_t stream) {
    if (!sigma_out || !dirs || !sigma || !head_in) return LNB_ERR_INVALID_ARGUMENT;
    if (in_pad % 8 != 0 || in_pad < 3 + 6 * degree + 15) return LNB_ERR_INVALID_ARGUMENT;
    if (M == 0) return LNB_OK;
    const size_t smem = (size_t)kHeadRows * (in_pad + 8) * sizeof(__half);
    if (smem > 48 * 1024) return LNB_ERR_UNSUPPORTED;
    k_head_input<<<(M + kHeadRows - 1) / kHeadRows, kHeadRows, smem, as_stream(stream)>>>(
        static_cast<const __half *>(sigma_out), dirs, M, degree, in_pad, density_scale, sigma,
        static_cast<__half *>(head_in), n_active);
    count_launch();
    return launch_status();
}

int lnb_field_head_rgb(const void *head_out, uint32_t M, float *rgb, const int32_t *n_active, lnb_stream_t stream) {
    if (!head_out || !rgb) return LNB_ERR_INVALID_ARGUMENT;
    if (M == 0) return LNB_OK;
    k_head_rgb<<<nblk(M), kThreads, 0, as_stream(stream)>>>(static_cast<const __half *>(head_out), M, rgb, n_active);
    count_launch();
    return launch_status();
}

int lnb_lidar_loss(const float *weights_sum, const float *depth, const float *image, const float *gt,
                   const float *t0, uint32_t N, float alpha_d, float alpha_r, float alpha_i, float loss_scale,
                   float *g_weights_sum, float *g_depth, float *g_image, float *loss_out, lnb_stream_t stream) {
    if (!weights_sum || !depth || !image || !gt || !g_weights_sum || !g_depth || !g_image || !loss_out)
        return LNB_ERR_INVALID_ARGUMENT;
    if (N == 0) return LNB_OK;
    k_lidar_loss<<<nblk(N), kThreads, 0, as_stream(stream)>>>(weights_sum, depth, image, gt, t0, N, alpha_d, alpha_r,
                                                             alpha_i, loss_scale, g_weights_sum, g_depth, g_image,
                                                             loss_out);
    count_launch();
    return launch_status();
}

int lnb_lidar_loss_ex(const float *weights_sum, const float *depth, const float *image, const float *gt,
                      const float *t0, uint32_t N, float alpha_d, float alpha_r, float alpha_i, float loss_scale,
                      uint32_t patch_x, uint32_t patch_y, float alpha_grad, float inv_scale, float grad_clip,
                      float *g_weights_sum, float *g_depth, float *g_image, float *loss_out, lnb_stream_t stream) {
    if (!weights_sum || !depth || !image || !gt || !g_weights_sum || !g_depth || !g_image || !loss_out)
        return LNB_ERR_INVALID_ARGUMENT;
    if (patch_x == 0 || patch_y == 0 || patch_x * patch_y > N + (N == 0)) return LNB_ERR_INVALID_ARGUMENT;
    if (N == 0) return LNB_OK;
    k_lidar_loss_patch<<<nblk(N), kThreads, 0, as_stream(stream)>>>(weights_sum, depth, image, gt, t0, N, alpha_d, alpha_r,
                                                                   alpha_i, loss_scale, patch_x, patch_y, alpha_grad,
                                                                   inv_scale, grad_clip, g_weights_sum, g_depth, g_image,
                                                                   loss_out);
    count_launch();
    return launch_status();
}

int lnb_field_head_out_grad(const float *g_rgb, const float *rgb, uint32_t M, void *g_head_out,
                            const int32_t *n_active, lnb_stream_t stream) {
    if (!g_rgb || !rgb || !g_head_out) return LNB_ERR_INVALID_ARGUMENT;
    if (M == 0) return LNB_OK;
    k_head_out_grad<<<nblk(M), kThreads, 0, as_stream(stream)>>>(g_rgb, rgb, M, static_cast<__half *>(g_head_out), n_active);
    count_launch();
    return launch_status();
}

int lnb_field_sigma_out_grad(const float *g_sigma, const void *sigma_out, const void *g_head_in, uint32_t M,
                             uint32_t in_pad, uint32_t degree, float density_scale, void *g_sigma_out,
                             const int32_t *n_active, lnb_stream_t stream) {
    if (!g_sigma || !sigma_out || !g_head_in || !g_sigma_out) return LNB_ERR_INVALID_ARGUMENT;
    if (M == 0) return LNB_OK;
    k_sigma_out_grad<<<nblk(M), kThreads, 0, as_stream(stream)>>>(
        g_sigma, static_cast<const __half *>(sigma_out), static_cast<const __half *>(g_head_in), M, in_pad,
        3 + 6 * degree, density_scale, static_cast<__half *>(g_sigma_out), n_active);
    count_launch();
    return launch_status();
}

int lnb_lidar_rays(const float *pose, const int32_t *inds, uint32_t N, uint32_t H, uint32_t W, float fov_up,
                   float fov, float *rays_o, float *rays_d, lnb_stream_t stream) {
    if (!pose || !inds || !rays_o || !rays_d || H == 0 || W == 0) return LNB_ERR_INVALID_ARGUMENT;
    if (N == 0) return LNB_OK;
    k_lidar_rays<<<nblk(N), kThreads, 0, as_stream(stream)>>>(pose, inds, N, H, W, fov_up, fov, rays_o, rays_d);
    count_launch();
    return launch_status();
}

int lnb_lidar_batch(const float *pose, const int32_t *inds, const float *image, uint32_t N, uint32_t H, uint32_t W,
                    float fov_up, float fov, float *rays_o, float *rays_d, float *gt, lnb_stream_t stream) {
    if (!pose || !inds || !image || !rays_o || !rays_d || !gt || H == 0 || W == 0) return LNB_ERR_INVALID_ARGUMENT;
    if (N == 0) return LNB_OK;
    k_lidar_batch<<<nblk(N), kThreads, 0, as_stream(stream)>>>(pose, inds, image, N, H, W, fov_up, fov, rays_o, rays_d, gt);
    count_launch();
    return launch_status();
}

}  // extern "C"
