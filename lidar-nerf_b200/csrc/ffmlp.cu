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
        const uint32_t dst = tile_chunk_addr(tile0 + t * tile_stride, r, cc);
        if (col < cols) cp_async16(dst, src + (size_t)r * ld + col);
        else if (zero_pad) cp_async16(dst, src, 0);
    }
    // asynchronous: the caller waits (cp_async_wait_all) before publishing the tile to the tensor core
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
            __half *__restrict__ fbuf, __half *__restrict__ Y, const int32_t *__restrict__ n_active) {
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
    cp_async_wait_all();
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = lds32(s_slot);
    const uint32_t d_hid = tmem;         // columns [0,64)
    const uint32_t d_out = tmem + 64;    // columns [64,80)
    const uint32_t lane_sel = (warp * 32u) << 16;

    uint32_t phase = 0;
    const uint32_t n_tiles = active_rows(B, n_active) / kRows;   // B (the buffer stride) stays the full batch
    // the input tile of the NEXT row tile is fetched (cp.async) as soon as the first-layer MMA has consumed the
    // current one, so its HBM latency hides behind the remaining layers
    if (blockIdx.x < n_tiles)
        load_tiles(s_x, kTileBytes, X + (size_t)blockIdx.x * kRows * sh.in_dim, kRows, sh.in_dim, sh.in_dim, false);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t row0 = (size_t)tile * kRows;
        cp_async_wait_all();
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
            if (layer == 0 && tile + gridDim.x < n_tiles)   // s_x is free: prefetch the next tile's inputs
                load_tiles(s_x, kTileBytes, X + (size_t)(tile + gridDim.x) * kRows * sh.in_dim, kRows, sh.in_dim,
                           sh.in_dim, false);
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
// Warp-specialised: two compute warpgroups (128 threads each, one 128-row tile in flight each) + one MMA warp
// whose lane 0 issues EVERY tcgen05.mma of the CTA (so all accumulations into the shared weight-gradient
// accumulators are ordered on the tensor pipe).  A warpgroup loads all inputs of its tile at once
// (G, every saved activation, X), then walks the layers: [operands ready] -> MMA warp -> [accumulator ready] ->
// epilogue (ReLU mask, fp16) -> next layer's operand.  While one warpgroup waits on memory or the tensor core,
// the other runs its epilogue.
//
// TMEM columns: [0,64) / [64,128) dgrad accumulators of warpgroup 0 / 1; then dW_out^T (64), dW_hid[l] (64 each),
// dW_in per 64-column input tile (64 each) - accumulated over ALL tiles of the CTA, flushed once at the end.
constexpr uint32_t kWG = 2;
constexpr uint32_t kBwdThreads = kWG * 128 + 32;

__device__ __forceinline__ void wg_sync(uint32_t wg) {
    asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// cooperative movers for ONE warpgroup (128 threads, tid = thread index inside the group)
__device__ __forceinline__ void wg_load_tiles(uint32_t tid, uint32_t tile0, const __half *__restrict__ src,
                                              uint32_t cols, uint32_t ld) {
    const uint32_t kt = (cols + 63) / 64;
    const uint32_t cpr = kt * 8;
    for (uint32_t q = tid; q < kRows * cpr; q += 128) {
        const uint32_t r = q / cpr, c = q - r * cpr;
        if (c * 8 < cols)
            cp_async16(tile_chunk_addr(tile0 + (c >> 3) * kTileBytes, r, c & 7), src + (size_t)r * ld + c * 8);
    }
}
__device__ __forceinline__ void wg_store_tile_rows(uint32_t tid, uint32_t tile, __half *__restrict__ dst) {
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j) {
        const uint32_t q = tid + j * 128;
        const uint32_t r = q >> 3, c = q & 7;
        *reinterpret_cast<uint4 *>(dst + (size_t)r * 64 + c * 8) = lds128(tile_chunk_addr(tile, r, c));
    }
}

__global__ void __launch_bounds__(kBwdThreads, 1)
k_ffmlp_bwd(const __half *__restrict__ G, const __half *__restrict__ X, const __half *__restrict__ W,
            const __half *__restrict__ fbuf, uint32_t B, Shape sh, __half *__restrict__ bbuf,
            __half *__restrict__ dX, float *__restrict__ wgrad /* fp32, flat weight layout */,
            const int32_t *__restrict__ n_active) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t n_act = sh.n_hid + 1;                                  // saved activation tiles per row tile
    const uint32_t s_win = sbase;
    const uint32_t s_whid = s_win + sh.kt_in * kWTileBytes;
    const uint32_t s_wout = s_whid + sh.n_hid * kWTileBytes;
    const uint32_t wg_bytes = (1 + n_act + sh.kt_in) * kTileBytes;        // G (aliased by dH), activations, X
    const uint32_t s_wg0 = s_wout + 2048;
    const uint32_t s_bar = s_wg0 + kWG * wg_bytes;                        // ready[2], done[2], fin, slot

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wg = threadIdx.x >> 7;                                 // 0,1 compute; 2 = MMA warp
    const uint32_t tid = threadIdx.x & 127;
    const uint32_t bar_ready0 = s_bar, bar_done0 = s_bar + 16, bar_fin = s_bar + 32, s_slot = s_bar + 40;

    if (warp == kWG * 4) tmem_alloc(s_slot, 512);
    if (threadIdx.x == 0) {
        for (uint32_t g = 0; g < kWG; ++g) {
            mbar_init(bar_ready0 + 8 * g, 128);
            mbar_init(bar_done0 + 8 * g, 1);
        }
        mbar_init(bar_fin, 1);
        mbar_init_fence();
    }
    // weights (all threads), zero padding of the per-group G and X tiles (their padding columns are read by N = 64 MMAs)
    {
        const uint32_t nthr = kBwdThreads, t = threadIdx.x;
        auto load_w = [&](uint32_t tile0, uint32_t stride, const __half *src, uint32_t rows, uint32_t cols, uint32_t ld) {
            const uint32_t kt = (cols + 63) / 64, cpr = kt * 8;
            for (uint32_t q = t; q < rows * cpr; q += nthr) {
                const uint32_t r = q / cpr, c = q - r * cpr;
                const bool ok = c * 8 < cols;
                cp_async16(tile_chunk_addr(tile0 + (c >> 3) * stride, r, c & 7), ok ? src + (size_t)r * ld + c * 8 : src,
                           ok ? 16u : 0u);
            }
        };
        load_w(s_win, kWTileBytes, W, kHid, sh.in_dim, sh.in_dim);
        for (uint32_t l = 0; l < sh.n_hid; ++l)
            load_w(s_whid + l * kWTileBytes, kWTileBytes, W + sh.w_in_elems + (size_t)l * kHid * kHid, kHid, kHid, kHid);
        load_w(s_wout, kWOutBytes, W + sh.w_in_elems + (size_t)sh.n_hid * kHid * kHid, kOut, kHid, kHid);
        for (uint32_t q = t; q < kWG * wg_bytes / 16; q += nthr) sts128(s_wg0 + q * 16, make_uint4(0, 0, 0, 0));
        cp_async_wait_all();
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = lds32(s_slot);
    const uint32_t d_wout = tmem + 128;
    const uint32_t d_whid = tmem + 192;
    const uint32_t d_win = d_whid + 64 * sh.n_hid;

    const uint32_t n_tiles = active_rows(B, n_active) / kRows;
    const uint32_t stride_tiles = gridDim.x * kWG;
    const uint32_t n_iter = (n_tiles + stride_tiles - 1) / stride_tiles;
    const bool want_dx = dX != nullptr;
    const uint32_t extra_dx = (want_dx && sh.kt_in > 1) ? sh.kt_in - 1 : 0;

    if (wg == kWG) {
        // ================= MMA warp =================
        if (lane == 0) {
            uint32_t par_ready[kWG] = {0, 0};
            for (uint32_t it = 0; it < n_iter; ++it) {
                const uint32_t n_phase = 2 + sh.n_hid + extra_dx;
                for (uint32_t ph = 0; ph < n_phase; ++ph) {
                    for (uint32_t g = 0; g < kWG; ++g) {
                        const uint32_t tile = (it * gridDim.x + blockIdx.x) * kWG + g;
                        if (tile >= n_tiles) continue;
                        const uint32_t base = s_wg0 + g * wg_bytes;
                        const uint32_t s_g = base, s_d = base, s_h = base + kTileBytes;   // dH overwrites G
                        const uint32_t s_x = s_h + n_act * kTileBytes;
                        const uint32_t d_acc = tmem + 64 * g;
                        mbar_wait(bar_ready0 + 8 * g, par_ready[g]);
                        par_ready[g] ^= 1;
                        fence_after_sync();
                        // group 0 of a CTA that owns any tile is served first in iteration 0: it initialises the accumulators
                        const bool accw = !(it == 0 && g == 0);
                        if (ph == 0) {
                            issue_dgrad(d_acc, s_g, s_wout, 1);
                            issue_wgrad(d_wout, s_h + sh.n_hid * kTileBytes, s_g, accw);
                        } else if (ph <= sh.n_hid) {
                            const uint32_t layer = sh.n_hid - ph + 1;        // dpre of h_layer is in s_d
                            issue_dgrad(d_acc, s_d, s_whid + (layer - 1) * kWTileBytes, 4);
                            issue_wgrad(d_whid + 64 * (layer - 1), s_d, s_h + (layer - 1) * kTileBytes, accw);
                        } else if (ph == sh.n_hid + 1) {
                            for (uint32_t t = 0; t < sh.kt_in; ++t) issue_wgrad(d_win + 64 * t, s_d, s_x + t * kTileBytes, accw);
                            if (want_dx) issue_dgrad(d_acc, s_d, s_win, 4);
                        } else {
                            const uint32_t t = ph - sh.n_hid - 1;            // 1 .. kt_in-1
                            issue_dgrad(d_acc, s_d, s_win + t * kWTileBytes, 4);
                        }
                        mma_commit(bar_done0 + 8 * g);
                    }
                }
            }
            mma_commit(bar_fin);
        }
    } else {
        // ================= compute warpgroups =================
        const uint32_t base = s_wg0 + wg * wg_bytes;
        const uint32_t s_g = base, s_d = base, s_h = base + kTileBytes;   // dH overwrites G after phase 0
        const uint32_t s_x = s_h + n_act * kTileBytes;
        const uint32_t d_acc = tmem + 64 * wg;
        const uint32_t bar_ready = bar_ready0 + 8 * wg, bar_done = bar_done0 + 8 * wg;
        const uint32_t lane_sel = ((warp & 3u) * 32u) << 16;
        const uint32_t row = tid;
        uint32_t par_done = 0;
        for (uint32_t it = 0; it < n_iter; ++it) {
            const uint32_t tile = (it * gridDim.x + blockIdx.x) * kWG + wg;
            if (tile >= n_tiles) break;
            const size_t row0 = (size_t)tile * kRows;
            // ---- all inputs of this tile at once ----
            // G: 16 valid columns (chunks 0,1); chunks 2..7 re-zeroed every tile because dH aliases this tile
            for (uint32_t q = tid; q < kRows * 8; q += 128) {
                const uint32_t r = q >> 3, c = q & 7;
                cp_async16(tile_chunk_addr(s_g, r, c), G + (row0 + r) * kOut + (c < 2 ? c * 8 : 0), c < 2 ? 16u : 0u);
            }
            for (uint32_t l = 0; l < n_act; ++l)
                wg_load_tiles(tid, s_h + l * kTileBytes, fbuf + ((size_t)l * B + row0) * kHid, kHid, kHid);
            wg_load_tiles(tid, s_x, X + row0 * sh.in_dim, sh.in_dim, sh.in_dim);
            cp_async_wait_all();          // ~40 independent 16-byte copies per thread were in flight together
            fence_proxy_async();
            fence_before_sync();
            mbar_arrive(bar_ready);

            // ---- layers, last to first: epilogue = mask with the saved activation, fp16, operand for the next MMA ----
            for (int layer = (int)sh.n_hid; layer >= 0; --layer) {
                mbar_wait(bar_done, par_done);
                par_done ^= 1;
                fence_after_sync();
                const uint32_t s_act = s_h + (uint32_t)layer * kTileBytes;
#pragma unroll
                for (uint32_t half_id = 0; half_id < 2; ++half_id) {
                    uint32_t v[32];
                    tmem_ld32(d_acc + lane_sel + half_id * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (uint32_t c = 0; c < 4; ++c) {
                        const uint4 hv = lds128(tile_chunk_addr(s_act, row, half_id * 4 + c));
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
                fence_proxy_async();
                fence_before_sync();
                mbar_arrive(bar_ready);
                if (bbuf) {
                    wg_sync(wg);   // every row of s_d written
                    wg_store_tile_rows(tid, s_d, bbuf + ((size_t)(sh.n_hid - layer) * B + row0) * kHid);
                    wg_sync(wg);   // all readers done before the next epilogue rewrites s_d
                }
            }
            // ---- input-gradient tiles ----
            for (uint32_t t = 0; t < sh.kt_in; ++t) {
                if (t > 0 && !want_dx) break;
                mbar_wait(bar_done, par_done);
                par_done ^= 1;
                fence_after_sync();
                if (want_dx) {
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
                if (t + 1 < sh.kt_in && want_dx) {
                    fence_before_sync();
                    mbar_arrive(bar_ready);   // accumulator drained -> next 64 input columns
                }
            }
            // the next tile's loads overwrite G / activations / X: every MMA that read them has completed
            // (the last `done` wait above covers all MMAs issued for this tile)
            wg_sync(wg);
        }
    }

    // ---- flush the tensor-memory weight-gradient accumulators (fp32 atomics into the flat layout) ----
    // UMMA M = 64 puts row m at TMEM lane (m / 16) * 32 + m % 16: warp w (mod 4), lanes 0..15 own rows 16 (w%4) + lane.
    if (wg == 0) {
        mbar_wait(bar_fin, 0);
        fence_after_sync();
        if (blockIdx.x * kWG < n_tiles) {   // this CTA accumulated something
            const uint32_t lane_sel = ((warp & 3u) * 32u) << 16;
            const uint32_t m = (warp & 3u) * 16 + lane;  // valid when lane < 16
            float *w_in = wgrad;
            float *w_hid = wgrad + sh.w_in_elems;
            float *w_out = w_hid + (size_t)sh.n_hid * kHid * kHid;
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
    }
    fence_before_sync();
    __syncthreads();
    if (warp == kWG * 4) tmem_dealloc(tmem, 512);
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
    if (192 + 64 * sh->n_hid + 64 * sh->kt_in > 512) return LNB_ERR_UNSUPPORTED;   // TMEM budget (backward)
    return LNB_OK;
}

size_t fwd_smem(const Shape &sh) {
    return 1024 + (size_t)sh.kt_in * kWTileBytes + (size_t)sh.n_hid * kWTileBytes + 2048 +
           (size_t)sh.kt_in * kTileBytes + kTileBytes + 64;
}
size_t bwd_smem(const Shape &sh) {
    return 1024 + (size_t)sh.kt_in * kWTileBytes + (size_t)sh.n_hid * kWTileBytes + 2048 +
           (size_t)kWG * (1 + sh.n_hid + 1 + sh.kt_in) * kTileBytes + 64;
}

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
               const int32_t *n_active, cudaStream_t st) {
    const size_t smem = fwd_smem(sh);
    auto kern = k_ffmlp_fwd<kSave>;
    if (smem > 200 * 1024) return LNB_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    size_t per_sm = (227 * 1024) / (smem + 1024);
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);   // 4 x 128 TMEM columns per SM
    const uint32_t cap = (uint32_t)per_sm * (uint32_t)sm_count();
    const uint32_t grid = (B / kRows) < cap ? (B / kRows) : cap;
    kern<<<grid, kThreads, smem, st>>>(static_cast<const __half *>(inputs), static_cast<const __half *>(weights), B, sh,
                                      static_cast<__half *>(fbuf), static_cast<__half *>(outputs), n_active);
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
    return launch_fwd<true>(inputs, weights, B, sh, forward_buffer, outputs, nullptr, as_stream(stream));
}

int lnb_ffmlp_forward_ex(const void *inputs, const void *weights, uint32_t B, uint32_t input_dim,
                         uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                         uint32_t output_activation, void *forward_buffer, void *outputs, const int32_t *n_active,
                         lnb_stream_t stream) {
    if (!inputs || !weights || !outputs) return LNB_ERR_INVALID_ARGUMENT;
    Shape sh;
    int rc = check_shape(B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, &sh);
    if (rc != LNB_OK) return rc;
    if (B == 0) return LNB_OK;
    if (forward_buffer) return launch_fwd<true>(inputs, weights, B, sh, forward_buffer, outputs, n_active, as_stream(stream));
    return launch_fwd<false>(inputs, weights, B, sh, nullptr, outputs, n_active, as_stream(stream));
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
    return launch_fwd<false>(inputs, weights, B, sh, nullptr, outputs, nullptr, as_stream(stream));
}

size_t lnb_ffmlp_backward_workspace_bytes(uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim,
                                          uint32_t num_layers) {
    if (num_layers < 1) return 0;
    return sizeof(float) * (size_t)hidden_dim * ((size_t)input_dim + (size_t)hidden_dim * (num_layers - 1) + output_dim);
}

static int ffmlp_backward_impl(const void *grad, const void *inputs, const void *weights, const void *forward_buffer,
                               uint32_t B, const Shape &sh, int calc_grad_inputs, void *backward_buffer,
                               void *grad_inputs, float *wgrad_f32, const int32_t *n_active, cudaStream_t st) {
    const size_t smem = bwd_smem(sh);
    if (smem > 200 * 1024) return LNB_ERR_UNSUPPORTED;   // num_layers too deep for the two-tile backward
    cudaError_t e = cudaFuncSetAttribute(k_ffmlp_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    const uint32_t cap = (uint32_t)sm_count();   // 512 TMEM columns: one CTA per SM, two row tiles in flight each
    const uint32_t pairs = (B / kRows + kWG - 1) / kWG;
    const uint32_t grid = pairs < cap ? pairs : cap;
    k_ffmlp_bwd<<<grid, kBwdThreads, smem, st>>>(
        static_cast<const __half *>(grad), static_cast<const __half *>(inputs), static_cast<const __half *>(weights),
        static_cast<const __half *>(forward_buffer), B, sh, static_cast<__half *>(backward_buffer),
        calc_grad_inputs ? static_cast<__half *>(grad_inputs) : nullptr, wgrad_f32, n_active);
    count_launch();
    return launch_status();
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
    rc = ffmlp_backward_impl(grad, inputs, weights, forward_buffer, B, sh, calc_grad_inputs, backward_buffer,
                             grad_inputs, static_cast<float *>(workspace), nullptr, st);
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

int lnb_ffmlp_backward_accumulate(const void *grad, const void *inputs, const void *weights,
                                  const void *forward_buffer, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                                  uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                                  uint32_t output_activation, int calc_grad_inputs, void *grad_inputs,
                                  float *grad_weights_f32, const int32_t *n_active, lnb_stream_t stream) {
    if (!grad || !inputs || !weights || !forward_buffer || !grad_weights_f32) return LNB_ERR_INVALID_ARGUMENT;
    if (calc_grad_inputs && !grad_inputs) return LNB_ERR_INVALID_ARGUMENT;
    Shape sh;
    int rc = check_shape(B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, &sh);
    if (rc != LNB_OK) return rc;
    if (B == 0) return LNB_OK;
    return ffmlp_backward_impl(grad, inputs, weights, forward_buffer, B, sh, calc_grad_inputs, nullptr, grad_inputs,
                               grad_weights_f32, n_active, as_stream(stream));
}

int lnb_allocate_splitk(size_t size) {
    (void)size;
    return LNB_OK;
}
int lnb_free_splitk(void) { return LNB_OK; }

}  // extern "C"
