// LiDAR field network as two fused tensor-core kernels: density MLP -> LiDAR head in one forward kernel, and the
// LiDAR-head backward with its glue folded in.  Same arithmetic as the unfused chain
//   ffmlp(sigma) -> [trunc_exp | concat(freq_enc(dir), geo_feat)] -> ffmlp(head) -> sigmoid
// (nerf/network.py:162-237 wiring, ffmlp.cu / freqencoder.cu kernels of the reference), restructured around one fact:
// every sample of a ray has the SAME direction.  The head's first layer is split as
//   W_in . [enc(dir) | geo | 0]  =  W_in[:, :nfreq] . enc(dir_ray)   (a per-ray 64-vector, computed once per ray)
//                                 + W_in[:, nfreq:nfreq+15] . geo     (a K=16..32 tensor-core step per sample)
// so the [M, 96] fp16 head input is never materialised: the head reads the density MLP's 16 outputs straight out of
// tensor memory, and the backward gathers the per-ray encoding rows (L2-resident, N x 96 fp16) for dW_in.
//
// What this removes from the step (per sample): 192 B head_in write + 192 B read (fwd) + 192 B read (bwd),
// 192 B g_head_in write + 30 B read, 32 B head_out, 32 B g_head_out, and five launches.
#include "common.cuh"
#include "mlp_tiles.cuh"
#include "mlp_bwd.cuh"

namespace lnb {
namespace {

struct FieldShape {
    Shape s;             // density MLP (input = hash-grid features)
    Shape h;             // LiDAR head (input = [enc(dir) | geo | pad], in_dim = in_pad)
    uint32_t nfreq;      // 3 + 6 * degree: first geo column of the head input
    uint32_t geo_tile;   // nfreq / 64   (64-column tile of W_in that holds the geo columns)
    uint32_t geo_off;    // nfreq % 64
};

// -----------------------------------------------------------------------------------------------------
// per-ray direction terms:  ray_enc[n] = fp16([freq_enc(dir_n) | 0...])  (freqencoder.cu:34-61 layout),
//                           ray_bias[n][h] = sum_{j<nfreq} W_in[h][j] * ray_enc[n][j]   (fp32)
// -----------------------------------------------------------------------------------------------------
// Real spherical harmonics of degree 4 (16 values) by the recurrence of csrc/shencoder.cu (same arithmetic, so the fused
// head sees exactly what lnb_sh_encode_forward would produce): Y_{l,+-m} = N_l^m Q_l^m(z) Re/Im (x + i y)^m.
__device__ __forceinline__ void sh4_eval(float x, float y, float z, float (&o)[16]) {
    constexpr int DEG = 4;
    const float norm[DEG][DEG] = {
        {0.28209479177387814f, 0.f, 0.f, 0.f},
        {0.48860251190291992f, -0.48860251190291998f, 0.f, 0.f},
        {0.63078313050504009f, -0.36418281019735976f, 0.18209140509867988f, 0.f},
        {0.7463526651802308f, -0.3046971996429772f, 0.096353714754685155f, -0.039336239328442907f}};
    float A[DEG + 1], Bm[DEG + 1];
    A[0] = 1.f, Bm[0] = 0.f;
#pragma unroll
    for (int m = 1; m <= DEG; ++m) {
        A[m] = x * A[m - 1] - y * Bm[m - 1];
        Bm[m] = x * Bm[m - 1] + y * A[m - 1];
    }
    float Q[DEG][DEG + 1];
#pragma unroll
    for (int l = 0; l < DEG; ++l)
#pragma unroll
        for (int m = 0; m <= DEG; ++m) Q[l][m] = 0.f;
    float dfact = 1.f;
#pragma unroll
    for (int m = 0; m < DEG; ++m) {
        if (m > 0) dfact *= (float)(2 * m - 1);
        Q[m][m] = dfact;
        if (m + 1 < DEG) Q[m + 1][m] = (float)(2 * m + 1) * z * dfact;
#pragma unroll
        for (int l = m + 2; l < DEG; ++l)
            Q[l][m] = ((float)(2 * l - 1) * z * Q[l - 1][m] - (float)(l + m - 1) * Q[l - 2][m]) * (1.0f / (float)(l - m));
    }
#pragma unroll
    for (int l = 0; l < DEG; ++l)
#pragma unroll
        for (int m = 0; m <= l; ++m) {
            const float nq = norm[l][m] * Q[l][m];
            o[l * l + l + m] = nq * A[m];
            if (m > 0) o[l * l + l - m] = nq * Bm[m];
        }
}

constexpr uint32_t kRaysPerBlock = 16;   // 4 rays at a time (one per 64-thread group), 4 rounds
__global__ void __launch_bounds__(256)
k_ray_dir_terms(const float *__restrict__ rays_d, const __half *__restrict__ w_head, uint32_t N, uint32_t deg,
                uint32_t in_pad, __half *__restrict__ ray_enc, float *__restrict__ ray_bias) {
    // the first-layer weights are staged once per block (row pitch in_pad + 2 halves: odd word stride, no bank
    // conflicts when thread h walks row h) and reused for kRaysPerBlock rays
    __shared__ __align__(16) __half w[kHid * (128 + 2)];
    __shared__ float e[4][128];
    const uint32_t nfreq = dir_code_width(deg);      // columns of the direction encoding (frequency or SH, LNB_DIR_SH)
    const bool sh = (deg & 0x100u) != 0;
    const uint32_t pitch = in_pad + 2;
    for (uint32_t q = threadIdx.x; q < kHid * in_pad / 2; q += 256) {       // in_pad is even: copy half2 words
        const uint32_t r = (2 * q) / in_pad, c = 2 * q - r * in_pad;
        *reinterpret_cast<__half2 *>(w + r * pitch + c) = reinterpret_cast<const __half2 *>(w_head)[q];
    }
    const float half_pi = 3.141592653589793f / 2;
    const uint32_t sub = threadIdx.x >> 6, h = threadIdx.x & 63u;
    for (uint32_t round = 0; round < kRaysPerBlock / 4; ++round) {
        const uint32_t n = blockIdx.x * kRaysPerBlock + round * 4 + sub;
        const bool live = n < N;
        float d[3] = {0.f, 0.f, 0.f};
        if (live) d[0] = rays_d[n * 3], d[1] = rays_d[n * 3 + 1], d[2] = rays_d[n * 3 + 2];
        __syncthreads();   // weights staged (first round) / previous round's e[] fully consumed
        float shv[16];
        if (live && sh && h < 16) sh4_eval(d[0], d[1], d[2], shv);
        if (live)
            for (uint32_t j = h; j < in_pad; j += 64) {
                float v = 0.f;
                if (sh) {
                    if (j < 16) {
#pragma unroll
                        for (uint32_t q = 0; q < 16; ++q)
                            if (q == j) v = shv[q];
                    }
                } else if (j < 3) {
                    v = d[j];
                } else if (j < nfreq) {
                    const uint32_t f = (j - 3) / 6, r = (j - 3) % 6, a = r % 3;
                    const float arg = scalbnf(d[a], (int)f);
                    v = __sinf(arg + (r >= 3 ? 1.f : 0.f) * half_pi);
                }
                const unsigned short hv = mlp_from_float(v);
                reinterpret_cast<unsigned short *>(ray_enc)[(size_t)n * in_pad + j] = hv;
                e[sub][j] = mlp_to_float(hv);
            }
        __syncthreads();
        if (live) {
            const __half *wr = w + h * pitch;
            float acc = 0.f;
            for (uint32_t j = 0; j < nfreq; ++j) acc = fmaf(mlp_to_float(wr[j]), e[sub][j], acc);
            ray_bias[(size_t)n * kHid + h] = acc;
        }
    }
}

// this thread's 32 accumulator columns -> (+bias) -> ReLU -> fp16 (four 16-byte chunks of the operand-tile row)
__device__ __forceinline__ void epilogue_relu(uint32_t d_mine, uint32_t half, const float4 *__restrict__ bias,
                                              uint4 (&pk)[4]) {
    uint32_t v[32];
    tmem_ld32(d_mine, v);
    tmem_ld_wait();
#pragma unroll
    for (uint32_t c = 0; c < 4; ++c) {
        float f[8];
#pragma unroll
        for (uint32_t e = 0; e < 8; ++e) f[e] = __uint_as_float(v[c * 8 + e]);
        if (bias) {
            const float4 b0 = __ldg(bias + half * 8 + c * 2), b1 = __ldg(bias + half * 8 + c * 2 + 1);
            f[0] += b0.x, f[1] += b0.y, f[2] += b0.z, f[3] += b0.w;
            f[4] += b1.x, f[5] += b1.y, f[6] += b1.z, f[7] += b1.w;
        }
        pk[c].x = pack_half2(fmaxf(f[0], 0.f), fmaxf(f[1], 0.f));
        pk[c].y = pack_half2(fmaxf(f[2], 0.f), fmaxf(f[3], 0.f));
        pk[c].z = pack_half2(fmaxf(f[4], 0.f), fmaxf(f[5], 0.f));
        pk[c].w = pack_half2(fmaxf(f[6], 0.f), fmaxf(f[7], 0.f));
    }
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// swizzled 128 x 64 tile -> global rows of 128 B, fully coalesced (a warp writes 512 contiguous bytes); 256 threads
__device__ __forceinline__ void store_tile_rows_256(uint32_t t, uint32_t tile, __half *__restrict__ dst) {
#pragma unroll
    for (uint32_t j = 0; j < (kRows * 8) / 256; ++j) {
        const uint32_t q = t + j * 256;
        *reinterpret_cast<uint4 *>(dst + (size_t)q * 8) = lds128(tile_chunk_addr(tile, q >> 3, q & 7));
    }
}

// =====================================================================================================
// forward: enc [M, enc_dim] -> sigma [M], rgb [M,2]  (+ saved activations of both nets and sig_out for backward)
//
// 256 epilogue threads (two per tile row, 32 accumulator columns each) + one MMA warp.  No CTA-wide barrier in the
// tile loop: the epilogue threads publish an operand tile with mbarrier `ready` (256 arrivals), the MMA warp answers
// each layer with tcgen05.commit on `done`.  Three CTAs per SM keep three tiles in flight.
// =====================================================================================================
constexpr uint32_t kFwdEpiThreads = 256;
constexpr uint32_t kFwdThreads = kFwdEpiThreads + 32;

__global__ void __launch_bounds__(kFwdThreads, 3)
k_field_fwd(const __half *__restrict__ X, const __half *__restrict__ Ws, const __half *__restrict__ Wh,
            const int32_t *__restrict__ ray_ids, const float *__restrict__ ray_bias, uint32_t B, FieldShape fs,
            float density_scale, __half *__restrict__ fb_s, __half *__restrict__ sig_out, float *__restrict__ sigma,
            __half *__restrict__ fb_h, float *__restrict__ rgb, const int32_t *__restrict__ n_active) {
    extern __shared__ uint8_t smem_raw[];
    const Shape ss = fs.s;
    const Shape sh = fs.h;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t s_ws_in = sbase;
    const uint32_t s_ws_hid = s_ws_in + ss.kt_in * kWTileBytes;
    const uint32_t s_ws_out = s_ws_hid + ss.n_hid * kWTileBytes;
    const uint32_t s_wh_geo = s_ws_out + 2048;
    const uint32_t s_wh_hid = s_wh_geo + kWTileBytes;
    const uint32_t s_wh_out = s_wh_hid + sh.n_hid * kWTileBytes;
    const uint32_t s_x = s_wh_out + 2048;
    const uint32_t s_h = s_x + ss.kt_in * kTileBytes;
    const uint32_t bar_ready = s_h + kTileBytes;
    const uint32_t bar_done = bar_ready + 8;
    const uint32_t s_slot = bar_done + 8;

    const uint32_t warp = threadIdx.x >> 5;
    constexpr uint32_t kMmaWarp = kFwdEpiThreads / 32;

    // cooperative cp.async of rows x cols halves (row-major, leading dimension ld) into swizzled 64-column tiles
    auto load = [&](uint32_t tile0, uint32_t stride, const __half *src, uint32_t rows, uint32_t cols, uint32_t ld, bool zero_pad,
                    uint32_t t, uint32_t nthr) {
        const uint32_t kt = (cols + 63) / 64, cpr = kt * 8;
        for (uint32_t q = t; q < rows * cpr; q += nthr) {
            const uint32_t r = q / cpr, c = q - r * cpr;
            const uint32_t dst = tile_chunk_addr(tile0 + (c >> 3) * stride, r, c & 7);
            if (c * 8 < cols) cp_async16(dst, src + (size_t)r * ld + c * 8);
            else if (zero_pad) cp_async16(dst, src, 0);
        }
    };

    if (warp == kMmaWarp) tmem_alloc(s_slot, 128);
    if (threadIdx.x == 0) {
        mbar_init(bar_ready, kFwdEpiThreads);
        mbar_init(bar_done, 1);
        mbar_init_fence();
    }
    {
        const uint32_t t = threadIdx.x, n = kFwdThreads;
        load(s_ws_in, kWTileBytes, Ws, kHid, ss.in_dim, ss.in_dim, true, t, n);
        for (uint32_t l = 0; l < ss.n_hid; ++l)
            load(s_ws_hid + l * kWTileBytes, kWTileBytes, Ws + ss.w_in_elems + (size_t)l * kHid * kHid, kHid, kHid, kHid, false, t, n);
        load(s_ws_out, kWOutBytes, Ws + ss.w_in_elems + (size_t)ss.n_hid * kHid * kHid, kOut, kHid, kHid, false, t, n);
        // head: only the 64-column tile of W_in that holds the geo columns (the enc(dir) columns act through ray_bias)
        load(s_wh_geo, kWTileBytes, Wh + fs.geo_tile * 64, kHid, min(64u, sh.in_dim - fs.geo_tile * 64), sh.in_dim, true, t, n);
        for (uint32_t l = 0; l < sh.n_hid; ++l)
            load(s_wh_hid + l * kWTileBytes, kWTileBytes, Wh + sh.w_in_elems + (size_t)l * kHid * kHid, kHid, kHid, kHid, false, t, n);
        load(s_wh_out, kWOutBytes, Wh + sh.w_in_elems + (size_t)sh.n_hid * kHid * kHid, kOut, kHid, kHid, false, t, n);
        cp_async_wait_all();
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = lds32(s_slot);
    const uint32_t d_hid = tmem;
    const uint32_t d_out = tmem + 64;
    const uint32_t ks_geo = (fs.geo_off + 15 + 15) / 16;     // K steps covering the geo columns inside their tile
    const uint32_t n_tiles = active_rows(B, n_active) / kRows;

    if (warp == kMmaWarp) {
        // ================= MMA warp (converged; one elected lane issues) =================
        uint32_t par = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            // density MLP: input layer, hidden layers, output layer; then the head's geo step, hidden layers, output
            for (uint32_t step = 0; step < ss.n_hid + 2 + sh.n_hid + 2; ++step) {
                mbar_wait_warp(bar_ready, par);
                par ^= 1;
                fence_after_sync();
                if (step == 0) {
                    for (uint32_t t = 0; t < ss.kt_in; ++t) {
                        const uint32_t cols = min(64u, ss.in_dim - t * 64);
                        issue_kmajor(d_hid, s_x + t * kTileBytes, s_ws_in + t * kWTileBytes, cols / 16, kIdescFwdHid, t > 0);
                    }
                } else if (step <= ss.n_hid) {
                    issue_kmajor(d_hid, s_h, s_ws_hid + (step - 1) * kWTileBytes, 4, kIdescFwdHid, false);
                } else if (step == ss.n_hid + 1) {
                    issue_kmajor(d_out, s_h, s_ws_out, 4, kIdescFwdOut, false);
                } else if (step == ss.n_hid + 2) {
                    issue_kmajor(d_hid, s_h, s_wh_geo, ks_geo, kIdescFwdHid, false);
                } else if (step <= ss.n_hid + 2 + sh.n_hid) {
                    issue_kmajor(d_hid, s_h, s_wh_hid + (step - ss.n_hid - 3) * kWTileBytes, 4, kIdescFwdHid, false);
                } else {
                    issue_kmajor(d_out, s_h, s_wh_out, 4, kIdescFwdOut, false);
                }
                mma_commit_elect(bar_done);
            }
        }
    } else {
        // ================= epilogue threads: thread = (row, column half) =================
        const uint32_t row = threadIdx.x & 127u;
        const uint32_t half = (warp >> 2) & 1u;
        const uint32_t t256 = threadIdx.x;
        const uint32_t lane_sel = ((warp & 3u) * 32u) << 16;
        const uint32_t d_mine = d_hid + lane_sel + 32 * half;
        uint32_t par = 0;
        if (blockIdx.x < n_tiles)
            load(s_x, kTileBytes, X + (size_t)blockIdx.x * kRows * ss.in_dim, kRows, ss.in_dim, ss.in_dim, false, t256, kFwdEpiThreads);
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const size_t row0 = (size_t)tile * kRows;
            const size_t r = row0 + row;
            const uint32_t rid = (uint32_t)__ldg(ray_ids + r);
            cp_async_wait_all();
            fence_proxy_async();
            fence_before_sync();
            mbar_arrive(bar_ready);                       // this thread's share of the feature tile is in place

            // ---------------- density MLP ----------------
            for (uint32_t layer = 0; layer <= ss.n_hid; ++layer) {
                mbar_wait(bar_done, par);
                par ^= 1;
                fence_after_sync();
                if (layer == 0) {
                    if (tile + gridDim.x < n_tiles)   // s_x is free: prefetch the next tile's features
                        load(s_x, kTileBytes, X + (size_t)(tile + gridDim.x) * kRows * ss.in_dim, kRows, ss.in_dim,
                             ss.in_dim, false, t256, kFwdEpiThreads);
                    // this thread's 128 B of the per-ray head bias, needed four layers from now: pull the line into L1
                    // (the profile showed the head's first epilogue stalled on exactly this L2 round trip)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(ray_bias + (size_t)rid * kHid + half * 32));
                }
                uint4 pk[4];
                epilogue_relu(d_mine, half, nullptr, pk);
                epi_sync();                                   // the previous layer's tile has been streamed out
#pragma unroll
                for (uint32_t c = 0; c < 4; ++c) sts128(tile_chunk_addr(s_h, row, half * 4 + c), pk[c]);
                fence_proxy_async();
                fence_before_sync();
                mbar_arrive(bar_ready);
                epi_sync();                                   // every row of the tile is in shared memory
                store_tile_rows_256(t256, s_h, fb_s + ((size_t)layer * B + row0) * kHid);   // while the tensor core works
            }

            // density output: sig_out (fp16, kept for backward), sigma = exp(h0) * scale, geo -> head operand tile
            mbar_wait(bar_done, par);
            par ^= 1;
            fence_after_sync();
            epi_sync();                                       // last hidden tile streamed out before s_h is reused
            if (half == 0) {
                uint32_t v[16];
                tmem_ld16(d_out + lane_sel, v);
                tmem_ld_wait();
                __align__(16) unsigned short hv[16];
#pragma unroll
                for (uint32_t k = 0; k < 16; ++k) hv[k] = mlp_from_float(__uint_as_float(v[k]));
                const uint32_t *pw = reinterpret_cast<const uint32_t *>(hv);
                uint4 *dst = reinterpret_cast<uint4 *>(sig_out + r * kOut);
                dst[0] = make_uint4(pw[0], pw[1], pw[2], pw[3]);
                dst[1] = make_uint4(pw[4], pw[5], pw[6], pw[7]);
                sigma[r] = __expf(mlp_to_float(hv[0])) * density_scale;    // activation.py:6-20 (forward)
                // head operand: zeros over the K range, geo_feat = sig_out[1..15] at columns geo_off .. geo_off+14
                for (uint32_t c = 0; c < 2 * ks_geo; ++c) sts128(tile_chunk_addr(s_h, row, c), make_uint4(0, 0, 0, 0));
#pragma unroll
                for (uint32_t k = 1; k < 16; ++k)
                    sts16(tile_elem_addr(s_h, row, fs.geo_off + k - 1), hv[k]);
            }
            fence_proxy_async();
            fence_before_sync();
            mbar_arrive(bar_ready);

            // ---------------- LiDAR head ----------------
            const float4 *bias = reinterpret_cast<const float4 *>(ray_bias + (size_t)rid * kHid);
            for (uint32_t layer = 0; layer <= sh.n_hid; ++layer) {
                mbar_wait(bar_done, par);
                par ^= 1;
                fence_after_sync();
                uint4 pk[4];
                epilogue_relu(d_mine, half, layer == 0 ? bias : nullptr, pk);
                epi_sync();
#pragma unroll
                for (uint32_t c = 0; c < 4; ++c) sts128(tile_chunk_addr(s_h, row, half * 4 + c), pk[c]);
                fence_proxy_async();
                fence_before_sync();
                mbar_arrive(bar_ready);
                epi_sync();
                store_tile_rows_256(t256, s_h, fb_h + ((size_t)layer * B + row0) * kHid);
            }

            // head output -> (ray-drop, intensity) = sigmoid(fp16(h[0:2]))   (network.py:230)
            mbar_wait(bar_done, par);
            par ^= 1;
            fence_after_sync();
            if (half == 0) {
                uint32_t v[16];
                tmem_ld16(d_out + lane_sel, v);
                tmem_ld_wait();
                const float a = mlp_to_float(mlp_from_float(__uint_as_float(v[0])));
                const float b = mlp_to_float(mlp_from_float(__uint_as_float(v[1])));
                reinterpret_cast<float2 *>(rgb)[r] = make_float2(1.f / (1.f + __expf(-a)), 1.f / (1.f + __expf(-b)));
            }
            fence_before_sync();   // orders this tile's TMEM reads before the next tile's MMAs (released by `ready`)
        }
    }

    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem, 128);
}

int make_shape(uint32_t in_dim, uint32_t nl, Shape *sh) {
    if (in_dim == 0 || in_dim % 16 != 0 || in_dim > 128 || nl < 2) return LNB_ERR_UNSUPPORTED;
    sh->in_dim = in_dim;
    sh->kt_in = (in_dim + 63) / 64;
    sh->n_hid = nl - 1;
    sh->w_in_elems = kHid * in_dim;
    return LNB_OK;
}

int make_field_shape(uint32_t enc_dim, uint32_t sigma_layers, uint32_t head_in_pad, uint32_t head_layers,
                     uint32_t degree, uint32_t hidden, FieldShape *fs) {
    if (hidden != kHid) return LNB_ERR_UNSUPPORTED;
    int rc = make_shape(enc_dim, sigma_layers, &fs->s);
    if (rc != LNB_OK) return rc;
    rc = make_shape(head_in_pad, head_layers, &fs->h);
    if (rc != LNB_OK) return rc;
    if (!dir_code_valid(degree)) return LNB_ERR_UNSUPPORTED;
    fs->nfreq = dir_code_width(degree);
    if (fs->nfreq + 15 > head_in_pad) return LNB_ERR_INVALID_ARGUMENT;
    fs->geo_tile = fs->nfreq / 64;
    fs->geo_off = fs->nfreq % 64;
    if (fs->geo_off + 15 > 64) return LNB_ERR_UNSUPPORTED;      // geo columns must sit inside one 64-column tile
    return LNB_OK;
}

int sm_count_field() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace
}  // namespace lnb

using namespace lnb;

// bf16 build of this unit (-DLNB_BF16, see mlp_tiles.cuh): every entry point gets the suffix `_bf16`
#ifdef LNB_BF16
#define lnb_field_supported lnb_field_supported_bf16
#define lnb_field_ray_terms lnb_field_ray_terms_bf16
#define lnb_field_forward lnb_field_forward_bf16
#define lnb_field_head_backward_rows lnb_field_head_backward_rows_bf16
#define lnb_field_head_backward lnb_field_head_backward_bf16
#define lnb_debug_bwd_trace_head lnb_debug_bwd_trace_head_bf16
#endif

extern "C" {

int lnb_field_supported(uint32_t enc_dim, uint32_t sigma_layers, uint32_t head_in_pad, uint32_t head_layers,
                        uint32_t degree, uint32_t hidden) {
    FieldShape fs;
    int rc = make_field_shape(enc_dim, sigma_layers, head_in_pad, head_layers, degree, hidden, &fs);
    if (rc != LNB_OK) return rc;
    if (!(fs.geo_off % 32 + 15 <= 32 && (fs.geo_off == 11 || fs.geo_off == 39 || fs.geo_off == 16))) return LNB_ERR_UNSUPPORTED;
    return LNB_OK;
}

int lnb_field_ray_terms(const float *rays_d, const void *w_head, uint32_t N, uint32_t degree, uint32_t in_pad,
                        void *ray_enc, float *ray_bias, lnb_stream_t stream) {
    if (!rays_d || !w_head || !ray_enc || !ray_bias) return LNB_ERR_INVALID_ARGUMENT;
    if (!dir_code_valid(degree)) return LNB_ERR_UNSUPPORTED;
    if (in_pad > 128 || in_pad % 8 != 0 || dir_code_width(degree) + 15 > in_pad) return LNB_ERR_INVALID_ARGUMENT;
    if (N == 0) return LNB_OK;
    k_ray_dir_terms<<<(N + kRaysPerBlock - 1) / kRaysPerBlock, 256, 0, as_stream(stream)>>>(rays_d, static_cast<const __half *>(w_head), N, degree, in_pad,
                                                      static_cast<__half *>(ray_enc), ray_bias);
    count_launch();
    return launch_status();
}

int lnb_field_forward(const void *enc, const void *w_sigma, const void *w_head, const int32_t *ray_ids,
                      const float *ray_bias, uint32_t M, uint32_t enc_dim, uint32_t sigma_layers,
                      uint32_t head_in_pad, uint32_t head_layers, uint32_t degree, uint32_t hidden,
                      float density_scale, void *fb_sigma, void *sig_out, float *sigma, void *fb_head, float *rgb,
                      const int32_t *n_active, lnb_stream_t stream) {
    if (!enc || !w_sigma || !w_head || !ray_ids || !ray_bias || !fb_sigma || !sig_out || !sigma || !fb_head || !rgb)
        return LNB_ERR_INVALID_ARGUMENT;
    if (M % kRows != 0) return LNB_ERR_INVALID_ARGUMENT;
    FieldShape fs;
    int rc = make_field_shape(enc_dim, sigma_layers, head_in_pad, head_layers, degree, hidden, &fs);
    if (rc != LNB_OK) return rc;
    if (M == 0) return LNB_OK;
    const size_t smem = 1024 + (size_t)(fs.s.kt_in + fs.s.n_hid + 1 + fs.h.n_hid) * kWTileBytes + 2 * 2048 +
                        (size_t)(fs.s.kt_in + 1) * kTileBytes + 64;
    if (smem > 220 * 1024) return LNB_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(k_field_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    const uint32_t per_sm = (uint32_t)((227 * 1024) / (smem + 1024));
    const uint32_t cap = (uint32_t)sm_count_field() * (per_sm < 1 ? 1u : (per_sm > 3 ? 3u : per_sm));
    const uint32_t tiles = M / kRows;
    k_field_fwd<<<tiles < cap ? tiles : cap, kFwdThreads, smem, as_stream(stream)>>>(
        static_cast<const __half *>(enc), static_cast<const __half *>(w_sigma), static_cast<const __half *>(w_head),
        ray_ids, ray_bias, M, fs, density_scale, static_cast<__half *>(fb_sigma), static_cast<__half *>(sig_out), sigma,
        static_cast<__half *>(fb_head), rgb, n_active);
    count_launch();
    return launch_status();
}

static int field_head_backward_impl(const float *g_rgb, const float *rgb, const float *g_sigma, const void *sig_out,
                                    const int32_t *ray_ids, const void *ray_enc, const void *w_head, const void *fb_head,
                                    uint32_t M, uint32_t head_in_pad, uint32_t head_layers, uint32_t degree,
                                    uint32_t hidden, float density_scale, void *g_sig_out, float *grad_w_head_f32,
                                    const int32_t *n_active, const int32_t *row_idx, lnb_stream_t stream) {
    if (!g_rgb || !rgb || !g_sigma || !sig_out || !ray_ids || !ray_enc || !w_head || !fb_head || !g_sig_out ||
        !grad_w_head_f32)
        return LNB_ERR_INVALID_ARGUMENT;
    if (M % kRows != 0) return LNB_ERR_INVALID_ARGUMENT;
    FieldShape fs;
    int rc = make_field_shape(32, 2, head_in_pad, head_layers, degree, hidden, &fs);
    if (rc != LNB_OK) return rc;
    if (M == 0) return LNB_OK;
    BwdArgs a = {};
    a.W = static_cast<const __half *>(w_head);
    a.fbuf = static_cast<const __half *>(fb_head);
    a.B = M;
    a.sh = fs.h;
    a.wgrad = grad_w_head_f32;
    a.n_active = n_active;
    a.row_idx = row_idx;
    a.g_rgb = g_rgb;
    a.rgb = rgb;
    a.g_sigma = g_sigma;
    a.sig_out = static_cast<const __half *>(sig_out);
    a.ray_ids = ray_ids;
    a.ray_enc = static_cast<const __half *>(ray_enc);
    a.g_sig_out = static_cast<__half *>(g_sig_out);
    a.density_scale = density_scale;
    a.nfreq = fs.nfreq;
    a.geo_tile = fs.geo_tile;
    int rc2;
    if (fs.geo_off == 11) rc2 = launch_mlp_bwd<true, 0, 11>(a, (uint32_t)sm_count_field(), as_stream(stream));        // degree 12
    else if (fs.geo_off == 39) rc2 = launch_mlp_bwd<true, 32, 7>(a, (uint32_t)sm_count_field(), as_stream(stream));   // degree 6
    else if (fs.geo_off == 16) rc2 = launch_mlp_bwd<true, 0, 16>(a, (uint32_t)sm_count_field(), as_stream(stream));   // SH degree 4
    else return LNB_ERR_UNSUPPORTED;
    if (rc2 != LNB_OK) return rc2;
    count_launch();
    return launch_status();
}

int lnb_field_head_backward(const float *g_rgb, const float *rgb, const float *g_sigma, const void *sig_out,
                            const int32_t *ray_ids, const void *ray_enc, const void *w_head, const void *fb_head,
                            uint32_t M, uint32_t head_in_pad, uint32_t head_layers, uint32_t degree, uint32_t hidden,
                            float density_scale, void *g_sig_out, float *grad_w_head_f32, const int32_t *n_active,
                            lnb_stream_t stream) {
    return field_head_backward_impl(g_rgb, rgb, g_sigma, sig_out, ray_ids, ray_enc, w_head, fb_head, M, head_in_pad,
                                    head_layers, degree, hidden, density_scale, g_sig_out, grad_w_head_f32, n_active,
                                    nullptr, stream);
}

int lnb_field_head_backward_rows(const float *g_rgb, const float *rgb, const float *g_sigma, const void *sig_out,
                                 const int32_t *ray_ids, const void *ray_enc, const void *w_head, const void *fb_head,
                                 uint32_t M, uint32_t head_in_pad, uint32_t head_layers, uint32_t degree,
                                 uint32_t hidden, float density_scale, void *g_sig_out, float *grad_w_head_f32,
                                 const int32_t *row_idx, const int32_t *n_rows, lnb_stream_t stream) {
    if (!row_idx || !n_rows) return LNB_ERR_INVALID_ARGUMENT;
    return field_head_backward_impl(g_rgb, rgb, g_sigma, sig_out, ray_ids, ray_enc, w_head, fb_head, M, head_in_pad,
                                    head_layers, degree, hidden, density_scale, g_sig_out, grad_w_head_f32, n_rows,
                                    row_idx, stream);
}

#ifdef LNB_TRACE
// diagnostic builds only (build.py --trace): timeline recorded by CTA 0 of the head backward kernel
int lnb_debug_bwd_trace_head(unsigned long long *host_out, uint32_t max_events, int reset) {
    unsigned int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_bwd_trace_n, sizeof(n));
    if (n > 16384u) n = 16384u;
    if (n > max_events) n = max_events;
    if (host_out && n) cudaMemcpyFromSymbol(host_out, g_bwd_trace, sizeof(unsigned long long) * n);
    if (reset) {
        const unsigned int zero = 0;
        cudaMemcpyToSymbol(g_bwd_trace_n, &zero, sizeof(zero));
    }
    return (int)n;
}
#endif

}  // extern "C"
