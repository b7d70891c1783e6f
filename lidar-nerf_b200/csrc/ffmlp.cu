// Fully fused fp16 MLP (forward / inference / backward) on the sm_100a tensor cores.
//
// Behavioural spec: lidarnerf/ffmlp/src/ffmlp.cu of the reference (forward :460-576,739-941; backward
// :578-733,1059-1264; weight layout :861-864) and ffmlp/ffmlp.py:187-283.
//
// Design (not a port of the wmma kernels):
//  * a CTA owns 128 batch rows at a time (UMMA M = 128) and walks tiles persistently;
//  * all weights are staged once per CTA into shared memory in the 128B-swizzled canonical layout and
//    every layer is ONE group of tcgen05.mma instructions (M128 x N64 x K16 each) issued by one thread,
//    accumulating in fp32 in tensor memory (the reference accumulates in fp16, ffmlp.cu:94);
//  * the epilogue (tcgen05.ld -> ReLU -> fp16) writes the activation tile straight back into shared
//    memory in operand layout, so it is the next layer's A operand without touching HBM; the saved
//    activations the backward pass needs leave the SM as fully coalesced 512 B-per-warp stores;
//  * backward: the same swizzled tile is a K-major operand for dgrad (dH = dH' W) and an MN-major operand
//    for wgrad (dW += dH^T H), so weight gradients are accumulated IN TENSOR MEMORY across all tiles of a
//    CTA (UMMA M = 64) and flushed once with fp32 atomics - no split-K GEMMs, side streams, or
//    [num_layers, B, hidden] backward buffer round trip (ffmlp.cu:1107-1263 in the reference).
#include "common.cuh"
#include "tcgen05.cuh"

namespace lnb {
namespace {

using namespace tc;

constexpr uint32_t kRows = 128;        // batch rows per tile (UMMA M)
constexpr uint32_t kHid = 64;          // hidden width this build implements
constexpr uint32_t kOut = 16;          // padded output width
constexpr uint32_t kTileBytes = kRows * 128;   // 128 x 64 halves
constexpr uint32_t kWTileBytes = kHid * 128;   // 64 x 64 halves
constexpr uint32_t kWOutBytes = kOut * 128;    // 16 x 64 halves
constexpr uint32_t kThreads = 128;

constexpr uint32_t kIdescFwdHid = instr_desc_f16(128, 64, 0, 0);
constexpr uint32_t kIdescFwdOut = instr_desc_f16(128, 16, 0, 0);
constexpr uint32_t kIdescDgrad = instr_desc_f16(128, 64, 0, 1);   // A K-major, B = W read MN-major
constexpr uint32_t kIdescWgrad = instr_desc_f16(64, 64, 1, 1);    // both operands MN-major, M = 64

struct Shape {
    uint32_t in_dim, kt_in, n_hid;  // n_hid = num_layers - 1 hidden-to-hidden matmuls
    uint32_t w_in_elems;            // offsets (in halves) into the flat weight vector
};

// ---- cooperative tile movers (all 128 threads) ---------------------------------------------------

// rows x cols halves, row-major in global with leading dimension ld -> swizzled tiles of 64 columns.
// Chunks beyond `cols` are zero-filled when `zero_pad` (needed when the tile is later read MN-major
// with N = 64).
__device__ __forceinline__ void load_tiles(uint32_t tile0, uint32_t tile_stride, const __half *__restrict__ src,
                                           uint32_t rows, uint32_t cols, uint32_t ld, bool zero_pad) {
    const uint32_t kt = (cols + 63) / 64;
    const uint32_t chunks_per_row = kt * 8;
    for (uint32_t q = threadIdx.x; q < rows * chunks_per_row; q += kThreads) {
        const uint32_t r = q / chunks_per_row, c = q - r * chunks_per_row;
        const uint32_t t = c >> 3, cc = c & 7;
        const uint32_t col = c * 8;
        if (col < cols) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)r * ld + col));
            sts128(tile_chunk_addr(tile0 + t * tile_stride, r, cc), v);
        } else if (zero_pad) {
            sts128(tile_chunk_addr(tile0 + t * tile_stride, r, cc), make_uint4(0, 0, 0, 0));
        }
    }
}

// swizzled 128 x 64 tile -> global rows of 64 halves (128 B), fully coalesced (a warp writes 512 B).
__device__ __forceinline__ void store_tile_rows(uint32_t tile, __half *__restrict__ dst) {
#pragma unroll
    for (uint32_t j = 0; j < (kRows * 8) / kThreads; ++j) {
        const uint32_t q = threadIdx.x + j * kThreads;
        const uint32_t r = q >> 3, c = q & 7;
        const uint4 v = lds128(tile_chunk_addr(tile, r, c));
        *reinterpret_cast<uint4 *>(dst + (size_t)r * 64 + c * 8) = v;
    }
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// ---- MMA issue helpers (one thread) -----------------------------------------------------------------

// D[128 x N] (+)= A[128 x K] * B[N x K]^T ; both K-major tiles; K = 16 * ksteps (<= 64)
__device__ __forceinline__ void issue_kmajor(uint32_t d, uint32_t a_tile, uint32_t b_tile, uint32_t ksteps,
                                             uint32_t idesc, bool accumulate_first) {
    for (uint32_t k = 0; k < ksteps; ++k)
        mma_f16(d, smem_desc_sw128(a_tile + k * 32, 16), smem_desc_sw128(b_tile + k * 32, 16), idesc,
                (accumulate_first || k > 0) ? 1u : 0u);
}
// D[128 x 64] (+)= A[128 x K] (K-major tile) * W (tile holding W[K rows][64 cols], read MN-major)
__device__ __forceinline__ void issue_dgrad(uint32_t d, uint32_t a_tile, uint32_t w_tile, uint32_t ksteps) {
    for (uint32_t k = 0; k < ksteps; ++k)
        mma_f16(d, smem_desc_sw128(a_tile + k * 32, 16), smem_desc_sw128(w_tile + k * 2048, kWTileBytes),
                kIdescDgrad, k > 0 ? 1u : 0u);
}
// D[64 x 64] (+)= A^T B over the 128 rows of two activation tiles (both read MN-major)
__device__ __forceinline__ void issue_wgrad(uint32_t d, uint32_t a_tile, uint32_t b_tile, bool accumulate_first) {
    for (uint32_t k = 0; k < kRows / 16; ++k)
        mma_f16(d, smem_desc_sw128(a_tile + k * 2048, kTileBytes), smem_desc_sw128(b_tile + k * 2048, kTileBytes),
                kIdescWgrad, (accumulate_first || k > 0) ? 1u : 0u);
}

// =====================================================================================================
// forward / inference
// =====================================================================================================
template <bool kSaveActs>
__global__ void __launch_bounds__(kThreads)
k_ffmlp_fwd(const __half *__restrict__ X, const __half *__restrict__ W, uint32_t B, Shape sh,
            __half *__restrict__ fbuf, __half *__restrict__ Y) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t s_win = sbase;
    const uint32_t s_whid = s_win + sh.kt_in * kWTileBytes;
    const uint32_t s_wout = s_whid + sh.n_hid * kWTileBytes;
    const uint32_t s_x = s_wout + 2048;                      // keeps 1024-byte alignment
    const uint32_t s_h = s_x + sh.kt_in * kTileBytes;
    const uint32_t s_bar = s_h + kTileBytes;
    const uint32_t s_slot = s_bar + 8;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row = threadIdx.x;  // tile row owned in epilogues (TMEM lane)

    // ---- one-time setup ----
    if (warp == 0) tmem_alloc(s_slot, 128);
    if (threadIdx.x == 32) {
        mbar_init(s_bar, 1);
        mbar_init_fence();
    }
    load_tiles(s_win, kWTileBytes, W, kHid, sh.in_dim, sh.in_dim, true);
    for (uint32_t l = 0; l < sh.n_hid; ++l)
        load_tiles(s_whid + l * kWTileBytes, kWTileBytes, W + sh.w_in_elems + (size_t)l * kHid * kHid, kHid, kHid,
                   kHid, false);
    load_tiles(s_wout, kWOutBytes, W + sh.w_in_elems + (size_t)sh.n_hid * kHid * kHid, kOut, kHid, kHid, false);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = lds32(s_slot);
    const uint32_t d_hid = tmem;         // columns [0,64)
    const uint32_t d_out = tmem + 64;    // columns [64,80)
    const uint32_t lane_sel = (warp * 32u) << 16;

    uint32_t phase = 0;
    const uint32_t n_tiles = B / kRows;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t row0 = (size_t)tile * kRows;
        // input tile -> smem
        load_tiles(s_x, kTileBytes, X + row0 * sh.in_dim, kRows, sh.in_dim, sh.in_dim, false);
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after_sync();
            for (uint32_t t = 0; t < sh.kt_in; ++t) {
                const uint32_t cols = min(64u, sh.in_dim - t * 64);
                issue_kmajor(d_hid, s_x + t * kTileBytes, s_win + t * kWTileBytes, cols / 16, kIdescFwdHid, t > 0);
            }
            mma_commit(s_bar);
        }

        for (uint32_t layer = 0; layer <= sh.n_hid; ++layer) {
            // wait for this layer's accumulator
            mbar_wait(s_bar, phase);
            phase ^= 1;
            fence_after_sync();
            // epilogue: ReLU, fp16, back into the operand tile (row `row`)
#pragma unroll
            for (uint32_t half_id = 0; half_id < 2; ++half_id) {
                uint32_t v[32];
                tmem_ld32(d_hid + lane_sel + half_id * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (uint32_t c = 0; c < 4; ++c) {
                    uint4 pk;
                    pk.x = pack_half2(fmaxf(__uint_as_float(v[c * 8 + 0]), 0.f), fmaxf(__uint_as_float(v[c * 8 + 1]), 0.f));
                    pk.y = pack_half2(fmaxf(__uint_as_float(v[c * 8 + 2]), 0.f), fmaxf(__uint_as_float(v[c * 8 + 3]), 0.f));
                    pk.z = pack_half2(fmaxf(__uint_as_float(v[c * 8 + 4]), 0.f), fmaxf(__uint_as_float(v[c * 8 + 5]), 0.f));
                    pk.w = pack_half2(fmaxf(__uint_as_float(v[c * 8 + 6]), 0.f), fmaxf(__uint_as_float(v[c * 8 + 7]), 0.f));
                    sts128(tile_chunk_addr(s_h, row, half_id * 4 + c), pk);
                }
            }
            fence_proxy_async();
            fence_before_sync();
            __syncthreads();
            if (threadIdx.x == 0) {
                fence_after_sync();
                if (layer < sh.n_hid)
                    issue_kmajor(d_hid, s_h, s_whid + layer * kWTileBytes, 4, kIdescFwdHid, false);
                else
                    issue_kmajor(d_out, s_h, s_wout, 4, kIdescFwdOut, false);
                mma_commit(s_bar);
            }
            // while the tensor core works: stream the activation tile out for the backward pass
            if (kSaveActs) store_tile_rows(s_h, fbuf + ((size_t)layer * B + row0) * kHid);
            __syncthreads();  // every reader of s_h is done before the next epilogue overwrites it
        }

        // output layer accumulator -> [B,16] fp16
        mbar_wait(s_bar, phase);
        phase ^= 1;
        fence_after_sync();
        {
            uint32_t v[16];
            tmem_ld16(d_out + lane_sel, v);
            tmem_ld_wait();
            uint4 lo, hi;
            lo.x = pack_half2(__uint_as_float(v[0]), __uint_as_float(v[1]));
            lo.y = pack_half2(__uint_as_float(v[2]), __uint_as_float(v[3]));
            lo.z = pack_half2(__uint_as_float(v[4]), __uint_as_float(v[5]));
            lo.w = pack_half2(__uint_as_float(v[6]), __uint_as_float(v[7]));
            hi.x = pack_half2(__uint_as_float(v[8]), __uint_as_float(v[9]));
            hi.y = pack_half2(__uint_as_float(v[10]), __uint_as_float(v[11]));
            hi.z = pack_half2(__uint_as_float(v[12]), __uint_as_float(v[13]));
            hi.w = pack_half2(__uint_as_float(v[14]), __uint_as_float(v[15]));
            uint4 *dst = reinterpret_cast<uint4 *>(Y + (row0 + row) * kOut);
            dst[0] = lo;
            dst[1] = hi;
        }
        fence_before_sync();  // orders this tile's TMEM reads before the next tile's MMAs (after the barrier)
    }

    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
    (void)lane;
}

// =====================================================================================================
// backward
// =====================================================================================================
// TMEM columns: [0,64) dgrad accumulator, [64,128) dW_out^T, [128 + 64 l) dW_hid[l], then dW_in per k-tile.
__global__ void __launch_bounds__(kThreads)
k_ffmlp_bwd(const __half *__restrict__ G, const __half *__restrict__ X, const __half *__restrict__ W,
            const __half *__restrict__ fbuf, uint32_t B, Shape sh, __half *__restrict__ bbuf,
            __half *__restrict__ dX, float *__restrict__ wgrad /* fp32, flat weight layout */) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t s_win = sbase;
    const uint32_t s_whid = s_win + sh.kt_in * kWTileBytes;
    const uint32_t s_wout = s_whid + sh.n_hid * kWTileBytes;
    const uint32_t s_x = s_wout + 2048;
    const uint32_t s_h = s_x + sh.kt_in * kTileBytes;   // saved activation of the current layer
    const uint32_t s_d = s_h + kTileBytes;              // d(pre-activation) of the current layer
    const uint32_t s_g = s_d + kTileBytes;              // upstream gradient tile (cols >= 16 stay zero)
    const uint32_t s_bar = s_g + kTileBytes;
    const uint32_t s_slot = s_bar + 8;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row = threadIdx.x;

    if (warp == 0) tmem_alloc(s_slot, 512);
    if (threadIdx.x == 32) {
        mbar_init(s_bar, 1);
        mbar_init_fence();
    }
    load_tiles(s_win, kWTileBytes, W, kHid, sh.in_dim, sh.in_dim, true);
    for (uint32_t l = 0; l < sh.n_hid; ++l)
        load_tiles(s_whid + l * kWTileBytes, kWTileBytes, W + sh.w_in_elems + (size_t)l * kHid * kHid, kHid, kHid,
                   kHid, false);
    load_tiles(s_wout, kWOutBytes, W + sh.w_in_elems + (size_t)sh.n_hid * kHid * kHid, kOut, kHid, kHid, false);
    // zero the gradient tile and the input tiles once: their padding columns are read by N = 64 wgrad MMAs
    for (uint32_t q = threadIdx.x; q < kRows * 8; q += kThreads) sts128(s_g + q * 16, make_uint4(0, 0, 0, 0));
    for (uint32_t q = threadIdx.x; q < sh.kt_in * kRows * 8; q += kThreads) sts128(s_x + q * 16, make_uint4(0, 0, 0, 0));
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = lds32(s_slot);
    const uint32_t d_acc = tmem;
    const uint32_t d_wout = tmem + 64;
    const uint32_t d_whid = tmem + 128;
    const uint32_t d_win = d_whid + 64 * sh.n_hid;
    const uint32_t lane_sel = (warp * 32u) << 16;

    uint32_t phase = 0;
    const uint32_t n_tiles = B / kRows;
    uint32_t iter = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
        const size_t row0 = (size_t)tile * kRows;
        const bool acc = iter > 0;

        // ---- output layer: dH_last = G W_out ; dW_out^T += H_last^T G ----
        // G tile: 16 valid columns = chunks 0,1 of every row
        for (uint32_t q = threadIdx.x; q < kRows * 2; q += kThreads) {
            const uint32_t r = q >> 1, c = q & 1;
            sts128(tile_chunk_addr(s_g, r, c), __ldg(reinterpret_cast<const uint4 *>(G + (row0 + r) * kOut + c * 8)));
        }
        load_tiles(s_h, kTileBytes, fbuf + ((size_t)sh.n_hid * B + row0) * kHid, kRows, kHid, kHid, false);
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after_sync();
            issue_dgrad(d_acc, s_g, s_wout, 1);            // K = 16 output channels
            issue_wgrad(d_wout, s_h, s_g, acc);             // rows: hidden unit j, cols: output o (>=16 zero)
            mma_commit(s_bar);
        }

        // ---- hidden layers, last to first ----
        for (int layer = (int)sh.n_hid; layer >= 0; --layer) {
            mbar_wait(s_bar, phase);
            phase ^= 1;
            fence_after_sync();
            // epilogue: d(pre-act) = acc * (saved activation > 0)  -> s_d  (this thread's row)
#pragma unroll
            for (uint32_t half_id = 0; half_id < 2; ++half_id) {
                uint32_t v[32];
                tmem_ld32(d_acc + lane_sel + half_id * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (uint32_t c = 0; c < 4; ++c) {
                    const uint4 hv = lds128(tile_chunk_addr(s_h, row, half_id * 4 + c));
                    const __half2 *hh = reinterpret_cast<const __half2 *>(&hv);
                    uint32_t pk[4];
#pragma unroll
                    for (uint32_t e = 0; e < 4; ++e) {
                        const float2 act = __half22float2(hh[e]);
                        const float a = act.x > 0.f ? __uint_as_float(v[c * 8 + 2 * e]) : 0.f;
                        const float b = act.y > 0.f ? __uint_as_float(v[c * 8 + 2 * e + 1]) : 0.f;
                        pk[e] = pack_half2(a, b);
                    }
                    sts128(tile_chunk_addr(s_d, row, half_id * 4 + c), make_uint4(pk[0], pk[1], pk[2], pk[3]));
                }
            }
            __syncthreads();  // all rows of s_h consumed (mask) and s_d written before s_h is reloaded
            if (layer > 0) {
                load_tiles(s_h, kTileBytes, fbuf + ((size_t)(layer - 1) * B + row0) * kHid, kRows, kHid, kHid, false);
            } else {
                load_tiles(s_x, kTileBytes, X + row0 * sh.in_dim, kRows, sh.in_dim, sh.in_dim, false);
            }
            fence_proxy_async();
            fence_before_sync();
            __syncthreads();
            if (threadIdx.x == 0) {
                fence_after_sync();
                if (layer > 0) {
                    issue_dgrad(d_acc, s_d, s_whid + (layer - 1) * kWTileBytes, 4);   // dH_{l-1} = dpre_l W_l
                    issue_wgrad(d_whid + 64 * (layer - 1), s_d, s_h, acc);          // dW_l += dpre_l^T H_{l-1}
                } else {
                    for (uint32_t t = 0; t < sh.kt_in; ++t)
                        issue_wgrad(d_win + 64 * t, s_d, s_x + t * kTileBytes, acc);  // dW_in += dpre_0^T X
                    if (dX) issue_dgrad(d_acc, s_d, s_win, 4);                       // dX[:, 0:64]
                }
                mma_commit(s_bar);
            }
            if (bbuf) store_tile_rows(s_d, bbuf + ((size_t)(sh.n_hid - layer) * B + row0) * kHid);
        }

        // ---- input gradient, 64 columns at a time ----
        for (uint32_t t = 0; t < sh.kt_in; ++t) {
            mbar_wait(s_bar, phase);
            phase ^= 1;
            fence_after_sync();
            if (dX) {
                const uint32_t cols = min(64u, sh.in_dim - t * 64);
#pragma unroll
                for (uint32_t half_id = 0; half_id < 2; ++half_id) {
                    uint32_t v[32];
                    tmem_ld32(d_acc + lane_sel + half_id * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (uint32_t c = 0; c < 4; ++c) {
                        const uint32_t col = half_id * 32 + c * 8;
                        if (col < cols) {
                            uint4 pk;
                            pk.x = pack_half2(__uint_as_float(v[c * 8 + 0]), __uint_as_float(v[c * 8 + 1]));
                            pk.y = pack_half2(__uint_as_float(v[c * 8 + 2]), __uint_as_float(v[c * 8 + 3]));
                            pk.z = pack_half2(__uint_as_float(v[c * 8 + 4]), __uint_as_float(v[c * 8 + 5]));
                            pk.w = pack_half2(__uint_as_float(v[c * 8 + 6]), __uint_as_float(v[c * 8 + 7]));
                            *reinterpret_cast<uint4 *>(dX + (row0 + row) * sh.in_dim + t * 64 + col) = pk;
                        }
                    }
                }
            }
            fence_before_sync();
            __syncthreads();
            if (t + 1 < sh.kt_in) {
                if (threadIdx.x == 0) {
                    fence_after_sync();
                    if (dX) issue_dgrad(d_acc, s_d, s_win + (t + 1) * kWTileBytes, 4);
                    mma_commit(s_bar);
                }
            }
        }
    }

    // ---- flush the tensor-memory weight-gradient accumulators (fp32 atomics into the flat layout) ----
    // UMMA M = 64 puts row m at TMEM lane (m / 16) * 32 + m % 16: warp w, lanes 0..15 own rows 16 w + lane.
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (iter > 0) {
        const uint32_t m = warp * 16 + lane;  // valid when lane < 16
        float *w_in = wgrad;
        float *w_hid = wgrad + sh.w_in_elems;
        float *w_out = w_hid + (size_t)sh.n_hid * kHid * kHid;
        // dW_out^T : row = hidden j, col = output o
        {
            uint32_t v[16];
            tmem_ld16(d_wout + lane_sel, v);
            tmem_ld_wait();
            if (lane < 16)
#pragma unroll
                for (uint32_t o = 0; o < 16; ++o) atomicAdd(w_out + o * kHid + m, __uint_as_float(v[o]));
        }
        for (uint32_t l = 0; l < sh.n_hid; ++l)
#pragma unroll
            for (uint32_t half_id = 0; half_id < 2; ++half_id) {
                uint32_t v[32];
                tmem_ld32(d_whid + 64 * l + lane_sel + half_id * 32, v);
                tmem_ld_wait();
                if (lane < 16)
#pragma unroll
                    for (uint32_t n = 0; n < 32; ++n)
                        atomicAdd(w_hid + (size_t)l * kHid * kHid + m * kHid + half_id * 32 + n, __uint_as_float(v[n]));
            }
        for (uint32_t t = 0; t < sh.kt_in; ++t)
#pragma unroll
            for (uint32_t half_id = 0; half_id < 2; ++half_id) {
                uint32_t v[32];
                tmem_ld32(d_win + 64 * t + lane_sel + half_id * 32, v);
                tmem_ld_wait();
                if (lane < 16)
#pragma unroll
                    for (uint32_t n = 0; n < 32; ++n) {
                        const uint32_t col = t * 64 + half_id * 32 + n;
                        if (col < sh.in_dim) atomicAdd(w_in + m * sh.in_dim + col, __uint_as_float(v[n]));
                    }
            }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

__global__ void k_f32_to_f16(const float *__restrict__ src, __half *__restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2half_rn(src[i]);
}

int check_shape(uint32_t B, uint32_t in_dim, uint32_t out_dim, uint32_t hidden, uint32_t nl, uint32_t act,
                uint32_t out_act, Shape *sh) {
    if (B % kRows != 0) return LNB_ERR_INVALID_ARGUMENT;          // ffmlp.py:254-262 pads to 128
    if (hidden != kHid || out_dim != kOut || in_dim == 0 || in_dim % 16 != 0 || in_dim > 128 || nl < 2)
        return LNB_ERR_UNSUPPORTED;
    if (act != 0 || out_act != 6) return LNB_ERR_UNSUPPORTED;     // ReLU hidden, no output activation
    sh->in_dim = in_dim;
    sh->kt_in = (in_dim + 63) / 64;
    sh->n_hid = nl - 1;
    sh->w_in_elems = kHid * in_dim;
    if (128 + 64 * sh->n_hid + 64 * sh->kt_in > 512) return LNB_ERR_UNSUPPORTED;   // TMEM budget (backward)
    return LNB_OK;
}

size_t fwd_smem(const Shape &sh) {
    return 1024 + (size_t)sh.kt_in * kWTileBytes + (size_t)sh.n_hid * kWTileBytes + 2048 +
           (size_t)sh.kt_in * kTileBytes + kTileBytes + 64;
}
size_t bwd_smem(const Shape &sh) { return fwd_smem(sh) + 2 * (size_t)kTileBytes; }

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

template <bool kSave>
int launch_fwd(const void *inputs, const void *weights, uint32_t B, const Shape &sh, void *fbuf, void *outputs,
               cudaStream_t st) {
    const size_t smem = fwd_smem(sh);
    auto kern = k_ffmlp_fwd<kSave>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    size_t per_sm = (227 * 1024) / (smem + 1024);
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);   // 4 x 128 TMEM columns per SM
    const uint32_t cap = (uint32_t)per_sm * (uint32_t)sm_count();
    const uint32_t grid = (B / kRows) < cap ? (B / kRows) : cap;
    kern<<<grid, kThreads, smem, st>>>(static_cast<const __half *>(inputs), static_cast<const __half *>(weights), B, sh,
                                      static_cast<__half *>(fbuf), static_cast<__half *>(outputs));
    count_launch();
    return launch_status();
}

}  // namespace
}  // namespace lnb

using namespace lnb;

extern "C" {

int lnb_ffmlp_forward(const void *inputs, const void *weights, uint32_t B, uint32_t input_dim,
                      uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                      uint32_t output_activation, void *forward_buffer, void *outputs, lnb_stream_t stream) {
    if (!inputs || !weights || !forward_buffer || !outputs) return LNB_ERR_INVALID_ARGUMENT;
    Shape sh;
    int rc = check_shape(B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, &sh);
    if (rc != LNB_OK) return rc;
    if (B == 0) return LNB_OK;
    return launch_fwd<true>(inputs, weights, B, sh, forward_buffer, outputs, as_stream(stream));
}

int lnb_ffmlp_inference(const void *inputs, const void *weights, uint32_t B, uint32_t input_dim,
                        uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                        uint32_t output_activation, void *inference_buffer, void *outputs, lnb_stream_t stream) {
    (void)inference_buffer;  // activations never leave the SM in inference mode
    if (!inputs || !weights || !outputs) return LNB_ERR_INVALID_ARGUMENT;
    Shape sh;
    int rc = check_shape(B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, &sh);
    if (rc != LNB_OK) return rc;
    if (B == 0) return LNB_OK;
    return launch_fwd<false>(inputs, weights, B, sh, nullptr, outputs, as_stream(stream));
}

size_t lnb_ffmlp_backward_workspace_bytes(uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim,
                                          uint32_t num_layers) {
    if (num_layers < 1) return 0;
    return sizeof(float) * (size_t)hidden_dim * ((size_t)input_dim + (size_t)hidden_dim * (num_layers - 1) + output_dim);
}

int lnb_ffmlp_backward(const void *grad, const void *inputs, const void *weights, const void *forward_buffer,
                       uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim,
                       uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                       int calc_grad_inputs, void *backward_buffer, void *grad_inputs, void *grad_weights,
                       void *workspace, size_t workspace_bytes, lnb_stream_t stream) {
    if (!grad || !inputs || !weights || !forward_buffer || !workspace) return LNB_ERR_INVALID_ARGUMENT;
    if (calc_grad_inputs && !grad_inputs) return LNB_ERR_INVALID_ARGUMENT;
    Shape sh;
    int rc = check_shape(B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, &sh);
    if (rc != LNB_OK) return rc;
    const size_t need = lnb_ffmlp_backward_workspace_bytes(input_dim, output_dim, hidden_dim, num_layers);
    if (workspace_bytes < need) return LNB_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 15u) != 0) return LNB_ERR_INVALID_ARGUMENT;
    if (B == 0) return LNB_OK;
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(workspace, 0, need, st);
    if (e != cudaSuccess) return (int)e;
    const size_t smem = bwd_smem(sh);
    e = cudaFuncSetAttribute(k_ffmlp_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const uint32_t cap = (uint32_t)sm_count();   // 512 TMEM columns: one CTA per SM
    const uint32_t grid = (B / kRows) < cap ? (B / kRows) : cap;
    k_ffmlp_bwd<<<grid, kThreads, smem, st>>>(
        static_cast<const __half *>(grad), static_cast<const __half *>(inputs), static_cast<const __half *>(weights),
        static_cast<const __half *>(forward_buffer), B, sh, static_cast<__half *>(backward_buffer),
        calc_grad_inputs ? static_cast<__half *>(grad_inputs) : nullptr, static_cast<float *>(workspace));
    count_launch();
    rc = launch_status();
    if (rc != LNB_OK) return rc;
    if (grad_weights) {
        const size_t n = need / sizeof(float);
        k_f32_to_f16<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(static_cast<const float *>(workspace),
                                                                static_cast<__half *>(grad_weights), n);
        count_launch();
        rc = launch_status();
    }
    return rc;
}

int lnb_allocate_splitk(size_t size) {
    (void)size;
    return LNB_OK;
}
int lnb_free_splitk(void) { return LNB_OK; }

}  // extern "C"
