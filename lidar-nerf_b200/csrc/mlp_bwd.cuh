// Backward kernel of the fused fp16 MLP, shared by ffmlp.cu (generic: lnb_ffmlp_backward*) and field.cu (LiDAR head
// with its glue folded in).  Behavioural spec: ffmlp.cu:578-733,1059-1264 of the reference (dgrad kernel + split-K
// wgrad GEMMs); here one kernel does both.
//
// Warp-specialised: kWG compute warpgroups (128 threads, one 128-row tile in flight each) + one MMA warp whose lane 0
// issues EVERY tcgen05.mma of the CTA (so all accumulations into the shared weight-gradient accumulators are ordered
// on the tensor pipe).  Per tile a warpgroup walks the layers last to first:
//   [operands ready] -> MMA warp: dgrad (M128) + wgrad (M64, accumulating in tensor memory over all tiles of the CTA)
//   -> [accumulator ready] -> epilogue (ReLU mask, fp16) -> operand of the next layer.
// Latency hiding inside a warpgroup:
//   * the ReLU masks of all layers are extracted into registers (64 bits per layer and row) when the tile lands, so an
//     epilogue touches no shared memory but its own output row;
//   * a saved-activation tile is dead as soon as the wgrad MMA that read it retires, so the NEXT tile's activations
//     are fetched (cp.async) into it layer by layer while the current tile is still being processed; the small
//     per-row inputs of the next tile travel in registers.
//
// TMEM columns: [0,64) / [64,128) dgrad accumulators of warpgroup 0 / 1; then dW_out^T (64), dW_hid[l] (64 each),
// dW_in per 64-column input tile (64 each).
#pragma once
#include "mlp_tiles.cuh"

namespace lnb {
namespace {

struct BwdArgs {
    const __half *W, *fbuf;
    uint32_t B;
    Shape sh;
    float *wgrad;               // fp32, flat weight layout, accumulated with atomics
    const int32_t *n_active;
    // Row compaction (nullable).  With row_idx, tile row r of the kernel is row row_idx[r] of every per-row INPUT
    // (saved activations, X, sig_out, g_rgb, ...), *n_active is the EXACT number of valid rows (rows beyond it inside
    // the last tile contribute nothing), and the per-row OUTPUTS (dX / g_sig_out) and the generic G are in compact order.
    const int32_t *row_idx;
    // ---- generic mode ----
    const __half *G, *X;        // [B,16] gradient of the output, [B,in_dim] inputs
    __half *bbuf, *dX;          // nullable: [n_act,B,64] pre-activation gradients / [B,in_dim] input gradient
    // ---- LiDAR-head mode ----
    const float *g_rgb, *rgb, *g_sigma;
    const __half *sig_out;      // [B,16] density-MLP output rows (h0 | geo_feat)
    const int32_t *ray_ids;     // [B]
    const __half *ray_enc;      // [N_rays, in_dim]: fp16 [freq_enc(dir) | 0]
    __half *g_sig_out;          // [B,16]
    float density_scale;
    uint32_t nfreq, geo_tile;
};

// Optional in-kernel timeline (diagnostic builds only: build.py --trace, scripts/diag_bwd_trace.py): CTA 0 records
// (tag, SM clock) pairs for one thread of each role.  Compiled out of the shipped library.
#ifdef LNB_TRACE
__device__ unsigned long long g_bwd_trace[16384];
__device__ unsigned int g_bwd_trace_n;
#define LNB_TR(role, ev, arg)                                                                                   \
    do {                                                                                                        \
        if (blockIdx.x == 0) {                                                                                  \
            const unsigned int i_ = atomicAdd(&g_bwd_trace_n, 1u);                                              \
            if (i_ < 16384u)                                                                                    \
                g_bwd_trace[i_] = ((unsigned long long)(((role) << 12) | ((ev) << 4) | ((arg) & 15u)) << 44) |  \
                                  ((unsigned long long)clock64() & ((1ull << 44) - 1));                        \
        }                                                                                                       \
    } while (0)
#else
#define LNB_TR(role, ev, arg) do { } while (0)
#endif

struct HeadRow {                // per-row inputs of the LiDAR-head mode, prefetched one tile ahead
    uint32_t rid;
    uint4 so_lo, so_hi;
    float2 gr, pr;
    float gs;
};

__device__ __forceinline__ void sts16(uint32_t addr, unsigned short v) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint32_t tile_elem_addr(uint32_t tile, uint32_t row, uint32_t col) {
    return tile_chunk_addr(tile, row, col >> 3) + (col & 7u) * 2u;
}

constexpr uint32_t kTGThreads = 256;                  // two threads per tile row: each owns 32 of the 64 columns
// kTG = row tiles in flight per CTA (tile groups).  kTG = 1: 288-thread CTAs with <= 256 tensor-memory columns, TWO
// per SM, each with its own MMA-issuing warp and its own weight-gradient accumulators (MMA issue, ~100 cycles per
// instruction from one thread, is the serial resource of this kernel).  kTG = 2: one 544-thread CTA per SM whose
// single MMA warp serves both tiles (shapes whose accumulators need more than 256 columns).
__host__ __device__ constexpr uint32_t bwd_threads(uint32_t tg) { return tg * kTGThreads + 32; }

// tensor-memory columns: dgrad accumulator(s) | dW_out^T (N = 16, 32 reserved) | dW_hid[l] (64 each) | dW_in (N_in)
__host__ __device__ inline uint32_t bwd_win_cols(const Shape &sh, bool head, uint32_t nfreq) {
    return head ? ((nfreq + 15 + 15) / 16) * 16 : sh.in_dim;
}
__host__ __device__ inline uint32_t bwd_tmem_cols(const Shape &sh, uint32_t tg, uint32_t win_cols) {
    return 64 * tg + 32 + 64 * sh.n_hid + win_cols;
}

__device__ __forceinline__ void tg_sync(uint32_t tg) {
    asm volatile("bar.sync %0, 256;" ::"r"(tg + 1) : "memory");
}
// cooperative movers for ONE tile group (256 threads, gtid = thread index inside the group)
// one 128 x 64 tile whose rows are 128 B apart in global memory (saved activations): no division, 4 copies per thread
__device__ __forceinline__ void tg_load_tile64(uint32_t gtid, uint32_t tile, const __half *__restrict__ src) {
#pragma unroll
    for (uint32_t j = 0; j < (kRows * 8) / kTGThreads; ++j) {
        const uint32_t q = gtid + j * kTGThreads;
        cp_async16(tile_chunk_addr(tile, q >> 3, q & 7), src + (size_t)q * 8);
    }
}
__device__ __forceinline__ void tg_load_tiles(uint32_t gtid, uint32_t tile0, const __half *__restrict__ src,
                                              uint32_t cols, uint32_t ld) {
    const uint32_t kt = (cols + 63) / 64;
    const uint32_t cpr = kt * 8;
    for (uint32_t q = gtid; q < kRows * cpr; q += kTGThreads) {
        const uint32_t r = q / cpr, c = q - r * cpr;
        if (c * 8 < cols)
            cp_async16(tile_chunk_addr(tile0 + (c >> 3) * kTileBytes, r, c & 7), src + (size_t)r * ld + c * 8);
    }
}
__device__ __forceinline__ void tg_store_tile_rows(uint32_t gtid, uint32_t tile, __half *__restrict__ dst) {
#pragma unroll
    for (uint32_t j = 0; j < (kRows * 8) / kTGThreads; ++j) {
        const uint32_t q = gtid + j * kTGThreads;
        const uint32_t r = q >> 3, c = q & 7;
        *reinterpret_cast<uint4 *>(dst + (size_t)r * 64 + c * 8) = lds128(tile_chunk_addr(tile, r, c));
    }
}

// ReLU mask of this thread's 32 columns of an activation tile as 16 SIMD words (0xffff per half that is > 0;
// activations are >= 0, so a signed 16-bit compare against zero is exact)
__device__ __forceinline__ void relu_mask_words(uint32_t tile, uint32_t row, uint32_t half, uint32_t (&m)[16]) {
#pragma unroll
    for (uint32_t c = 0; c < 4; ++c) {
        const uint4 hv = lds128(tile_chunk_addr(tile, row, half * 4 + c));
        m[c * 4 + 0] = __vcmpgts2(hv.x, 0u);
        m[c * 4 + 1] = __vcmpgts2(hv.y, 0u);
        m[c * 4 + 2] = __vcmpgts2(hv.z, 0u);
        m[c * 4 + 3] = __vcmpgts2(hv.w, 0u);
    }
}

template <bool kHead, uint32_t kGeoWin, uint32_t kGeoIdx, uint32_t kTG>
__global__ void __launch_bounds__(bwd_threads(kTG), kTG == 1 ? 2 : 1)
k_mlp_bwd(const BwdArgs a) {
    constexpr uint32_t kBwdThreads = bwd_threads(kTG);
    extern __shared__ uint8_t smem_raw[];
    const Shape sh = a.sh;
    const uint32_t B = a.B;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t n_act = sh.n_hid + 1;                                  // saved activation tiles per row tile (<= 3)
    const uint32_t s_win = sbase;
    const uint32_t s_whid = s_win + sh.kt_in * kWTileBytes;
    const uint32_t s_wout = s_whid + sh.n_hid * kWTileBytes;
    const uint32_t tg_bytes = (2 + n_act + sh.kt_in) * kTileBytes;        // G/dH ping-pong pair, activations, X
    const uint32_t s_tg0 = s_wout + 2048;
    const uint32_t s_bar = s_tg0 + kTG * tg_bytes;                        // ready[2], done[2], fin, slot

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_ready0 = s_bar, bar_done0 = s_bar + 16, bar_fin = s_bar + 32, s_slot = s_bar + 40;
    constexpr uint32_t kMmaWarp = kTG * kTGThreads / 32;

    const uint32_t win_cols = bwd_win_cols(sh, kHead, a.nfreq);
    const uint32_t tmem_cols = bwd_tmem_cols(sh, kTG, win_cols) <= 256 ? 256u : 512u;
    if (warp == kMmaWarp) tmem_alloc(s_slot, tmem_cols);
    if (threadIdx.x == 0) {
        for (uint32_t g = 0; g < kTG; ++g) {
            mbar_init(bar_ready0 + 8 * g, kTGThreads);
            mbar_init(bar_done0 + 8 * g, 1);
        }
        mbar_init(bar_fin, 1);
        mbar_init_fence();
    }
    // weights (all threads), zero padding of the per-group tiles (padding columns are read by N = 64 MMAs)
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
        load_w(s_win, kWTileBytes, a.W, kHid, sh.in_dim, sh.in_dim);
        for (uint32_t l = 0; l < sh.n_hid; ++l)
            load_w(s_whid + l * kWTileBytes, kWTileBytes, a.W + sh.w_in_elems + (size_t)l * kHid * kHid, kHid, kHid, kHid);
        load_w(s_wout, kWOutBytes, a.W + sh.w_in_elems + (size_t)sh.n_hid * kHid * kHid, kOut, kHid, kHid);
        for (uint32_t q = t; q < kTG * tg_bytes / 16; q += nthr) sts128(s_tg0 + q * 16, make_uint4(0, 0, 0, 0));
        cp_async_wait_all();
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = lds32(s_slot);
    const uint32_t d_wout = tmem + 64 * kTG;
    const uint32_t d_whid = d_wout + 32;
    const uint32_t d_win = d_whid + 64 * sh.n_hid;

    const int32_t *__restrict__ ridx = a.row_idx;
    uint32_t n_valid = B;                                                  // rows that exist (compact mode: exact count)
    if (ridx) {
        const int32_t nv = *a.n_active;
        n_valid = min(B, (uint32_t)(nv > 0 ? nv : 0));
    }
    const uint32_t n_tiles = ridx ? (n_valid + kRows - 1) / kRows : active_rows(B, a.n_active) / kRows;
    const uint32_t stride_tiles = gridDim.x * kTG;
    const uint32_t n_iter = (n_tiles + stride_tiles - 1) / stride_tiles;
    const bool want_dx = kHead || a.dX != nullptr;
    const uint32_t extra_dx = (!kHead && want_dx && sh.kt_in > 1) ? sh.kt_in - 1 : 0;

    if (warp == kMmaWarp) {
        // ================= MMA warp: all 32 lanes run this loop converged, one elected lane issues =================
        uint32_t par_ready[kTG] = {};
        for (uint32_t it = 0; it < n_iter; ++it) {
            const uint32_t n_phase = 2 + sh.n_hid + extra_dx;
            for (uint32_t ph = 0; ph < n_phase; ++ph) {
                for (uint32_t g = 0; g < kTG; ++g) {
                    const uint32_t tile = (it * gridDim.x + blockIdx.x) * kTG + g;
                    if (tile >= n_tiles) continue;
                    const uint32_t base = s_tg0 + g * tg_bytes;
                    // phase p reads the gradient tile (p & 1); the epilogue that follows writes tile ((p + 1) & 1), so it
                    // never waits for this phase's weight-gradient MMAs, which keep reading tile (p & 1)
                    const uint32_t s_cur = base + ((ph <= sh.n_hid + 1 ? ph : sh.n_hid + 1) & 1u) * kTileBytes;
                    const uint32_t s_h = base + 2 * kTileBytes;
                    const uint32_t s_x = s_h + n_act * kTileBytes;
                    const uint32_t d_acc = tmem + 64 * g;
                    mbar_wait_warp(bar_ready0 + 8 * g, par_ready[g]);
                    par_ready[g] ^= 1;
                    fence_after_sync();
                    if (lane == 0) LNB_TR(2u, 1u + g, ph);
                    // group 0 of a CTA that owns any tile is served first in iteration 0: it initialises the accumulators
                    const bool accw = !(it == 0 && g == 0);
                    if (ph == 0) {
                        issue_dgrad(d_acc, s_cur, s_wout, 1);
                        mma_commit_elect(bar_done0 + 8 * g);     // the epilogue needs only the dgrad accumulator
                        issue_wgrad(d_wout, s_h + sh.n_hid * kTileBytes, s_cur, accw, 16);
                    } else if (ph <= sh.n_hid) {
                        const uint32_t layer = sh.n_hid - ph + 1;        // dpre of h_layer is in s_cur
                        issue_dgrad(d_acc, s_cur, s_whid + (layer - 1) * kWTileBytes, 4);
                        mma_commit_elect(bar_done0 + 8 * g);
                        issue_wgrad(d_whid + 64 * (layer - 1), s_cur, s_h + (layer - 1) * kTileBytes, accw);
                    } else if (ph == sh.n_hid + 1) {
                        // dW_in: the X tiles are consecutive in shared memory -> ONE group of N = win_cols (<= 128).
                        // Last phase of the tile: one commit covers every MMA issued for it (tensor pipe is in order).
                        issue_wgrad(d_win, s_cur, s_x, accw, win_cols);
                        if (kHead) issue_dgrad(d_acc, s_cur, s_win + a.geo_tile * kWTileBytes, 4);   // geo tile of dX only
                        else if (want_dx) issue_dgrad(d_acc, s_cur, s_win, 4);
                        mma_commit_elect(bar_done0 + 8 * g);
                    } else {
                        const uint32_t t = ph - sh.n_hid - 1;            // 1 .. kt_in-1
                        issue_dgrad(d_acc, s_cur, s_win + t * kWTileBytes, 4);
                        mma_commit_elect(bar_done0 + 8 * g);
                    }
                    if (lane == 0) LNB_TR(2u, 3u + g, ph);
                }
            }
        }
        mma_commit_elect(bar_fin);
    } else {
        // ================= compute tile groups: 256 threads per tile, thread = (row, column half) =================
        const uint32_t tg = warp >> 3;                       // tile group
        const uint32_t half = (warp >> 2) & 1u;              // columns [32 half, 32 half + 32) of this thread's row
        const uint32_t row = threadIdx.x & 127u;             // = TMEM lane: warp % 4 selects the 32-lane quadrant
        const uint32_t gtid = threadIdx.x & 255u;
        const bool tracer = gtid == 0;
        const uint32_t base = s_tg0 + tg * tg_bytes;
        const uint32_t s_g = base;                           // gradient tiles: phase p reads base + (p & 1) tiles
        const uint32_t s_h = base + 2 * kTileBytes;
        const uint32_t s_x = s_h + n_act * kTileBytes;
        const uint32_t d_mine = tmem + 64 * tg + 32 * half + (((warp & 3u) * 32u) << 16);
        const uint32_t bar_ready = bar_ready0 + 8 * tg, bar_done = bar_done0 + 8 * tg;
        const uint32_t n_enc_chunks = kHead ? (a.nfreq + 7) / 8 : 0;
        uint32_t par_done = 0;

        // source row of tile row `rc` (compact numbering): identity without a row list; rows past the valid count
        // read row 0 (any valid address) and are neutralised where they could contribute
        auto src_row = [&](uint32_t rc) -> uint32_t { return ridx ? (rc < n_valid ? (uint32_t)__ldg(ridx + rc) : 0u) : rc; };
        auto load_head_row = [&](HeadRow &h, uint32_t tile) {
            const uint32_t rc = tile * kRows + row;
            const size_t r = src_row(rc);
            h.rid = (uint32_t)__ldg(a.ray_ids + r);
            h.so_lo = __ldg(reinterpret_cast<const uint4 *>(a.sig_out + r * kOut));
            h.so_hi = __ldg(reinterpret_cast<const uint4 *>(a.sig_out + r * kOut) + 1);
            h.gr = __ldg(reinterpret_cast<const float2 *>(a.g_rgb) + r);
            h.pr = __ldg(reinterpret_cast<const float2 *>(a.rgb) + r);
            h.gs = __ldg(a.g_sigma + r);
            if (rc >= n_valid) h.gr = make_float2(0.f, 0.f), h.gs = 0.f;     // padding row of the last compact tile
        };
        // the four tile rows this thread copies chunks of (tg_load_tile64's mapping: row (gtid >> 3) + 32 j)
        auto tile_rows4 = [&](uint32_t tile, uint32_t (&r4)[4]) {
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) r4[j] = src_row(tile * kRows + (gtid >> 3) + 32 * j);
        };
        auto load_act = [&](uint32_t layer, uint32_t tile, const uint32_t (&r4)[4]) {
            const __half *src = a.fbuf + (size_t)layer * B * kHid;
            if (!ridx) {
                tg_load_tile64(gtid, s_h + layer * kTileBytes, src + (size_t)tile * kRows * kHid);
            } else {
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
                    cp_async16(tile_chunk_addr(s_h + layer * kTileBytes, (gtid >> 3) + 32 * j, gtid & 7),
                               src + (size_t)r4[j] * kHid + (gtid & 7) * 8);
            }
        };
        // inputs that live in the X / G tiles: free once the last MMA of the previous tile has retired
        auto load_xg = [&](uint32_t tile, const HeadRow &h) {
            const size_t row0 = (size_t)tile * kRows;
            if (kHead) {
                const __half *enc_row = a.ray_enc + (size_t)h.rid * sh.in_dim;   // geo slots are zero in ray_enc
                for (uint32_t c = half; c < n_enc_chunks; c += 2)               // the row's two threads alternate chunks
                    cp_async16(tile_chunk_addr(s_x + (c >> 3) * kTileBytes, row, c & 7), enc_row + c * 8);
            } else {
                // G: 16 valid columns (chunks 0,1); chunks 2..7 re-zeroed every tile because dH reuses this tile
                for (uint32_t q = gtid; q < kRows * 8; q += kTGThreads) {
                    const uint32_t r = q >> 3, c = q & 7;
                    const bool live = c < 2 && (!ridx || row0 + r < n_valid);      // padding rows of a compact tile: zeros
                    cp_async16(tile_chunk_addr(s_g, r, c), a.G + (row0 + r) * kOut + (c < 2 ? c * 8 : 0), live ? 16u : 0u);
                }
                if (!ridx) {
                    tg_load_tiles(gtid, s_x, a.X + row0 * sh.in_dim, sh.in_dim, sh.in_dim);
                } else {
                    const uint32_t cpr = sh.kt_in * 8;
                    for (uint32_t q = gtid; q < kRows * cpr; q += kTGThreads) {
                        const uint32_t r = q / cpr, c = q - r * cpr;
                        if (c * 8 < sh.in_dim)
                            cp_async16(tile_chunk_addr(s_x + (c >> 3) * kTileBytes, r, c & 7),
                                       a.X + (size_t)src_row((uint32_t)row0 + r) * sh.in_dim + c * 8);
                    }
                }
            }
        };

        uint32_t tile = blockIdx.x * kTG + tg;
        bool have = tile < n_tiles;
        HeadRow hr = {};
        uint32_t rows_next[4] = {0, 0, 0, 0};     // source rows of the NEXT tile's activation copies (compact mode)
        if (have) {
            if (kHead) load_head_row(hr, tile);
            tile_rows4(tile, rows_next);
            for (uint32_t l = 0; l < n_act; ++l) load_act(l, tile, rows_next);
            load_xg(tile, hr);
        }
        while (have) {
            const size_t row0 = (size_t)tile * kRows;
            const size_t r = row0 + row;
            if (tracer) LNB_TR(tg, 0u, 0u);
            cp_async_wait_all();
            if (tracer) LNB_TR(tg, 1u, 0u);
            if (kHead) {
                if (half == 0) {
                    // G = d loss / d head_out: sigmoid' * g_rgb in columns 0,1 (network.py:230)
                    sts128(tile_chunk_addr(s_g, row, 0),
                           make_uint4(pack_half2(hr.gr.x * hr.pr.x * (1.f - hr.pr.x), hr.gr.y * hr.pr.y * (1.f - hr.pr.y)), 0, 0, 0));
#pragma unroll
                    for (uint32_t c = 1; c < 4; ++c) sts128(tile_chunk_addr(s_g, row, c), make_uint4(0, 0, 0, 0));
                } else {
#pragma unroll
                    for (uint32_t c = 4; c < 8; ++c) sts128(tile_chunk_addr(s_g, row, c), make_uint4(0, 0, 0, 0));
                }
            }
            // geo_feat = sig_out[1..15] at input columns nfreq .. nfreq+14, patched in by the thread of the row that
            // copied the one ray_enc chunk these columns can share (so it is ordered after its own cp.async)
            if (kHead && half == ((a.nfreq >> 3) & 1u)) {
                const unsigned short *hs = reinterpret_cast<const unsigned short *>(&hr.so_lo);
                const unsigned short *hs2 = reinterpret_cast<const unsigned short *>(&hr.so_hi);
#pragma unroll
                for (uint32_t k = 1; k < 16; ++k) {
                    const uint32_t col = a.nfreq + k - 1;
                    sts16(tile_elem_addr(s_x + (col >> 6) * kTileBytes, row, col & 63u), k < 8 ? hs[k] : hs2[k - 8]);
                }
            }
            fence_proxy_async();
            fence_before_sync();
            mbar_arrive(bar_ready);       // phase-0 operands of this thread are in place
            if (tracer) LNB_TR(tg, 2u, 0u);
            tg_sync(tg);                  // every thread's copies of this tile have landed (mask reads below)
            if (tracer) LNB_TR(tg, 8u, 0u);
            const uint32_t next = tile + stride_tiles;
            const bool have_next = next < n_tiles;
            HeadRow hn = {};
            if (kHead && have_next) load_head_row(hn, next);
            if (ridx && have_next) tile_rows4(next, rows_next);

            // ---- layers, last to first: epilogue = ReLU mask, fp16, operand for the next MMA ----
            for (int layer = (int)sh.n_hid; layer >= 0; --layer) {
                // while the tensor core works: this layer's ReLU mask (the tile is refilled only after a LATER phase)
                uint32_t m[16];
                relu_mask_words(s_h + (uint32_t)layer * kTileBytes, row, half, m);
                mbar_wait(bar_done, par_done);
                par_done ^= 1;
                fence_after_sync();
                if (tracer) LNB_TR(tg, 3u, (uint32_t)layer);
                uint32_t v[32];
                tmem_ld32(d_mine, v);
                tmem_ld_wait();
                if (tracer) LNB_TR(tg, 4u, (uint32_t)layer);
                const uint32_t ph = sh.n_hid - (uint32_t)layer;             // phase whose dgrad accumulator this is
                const uint32_t s_out = base + ((ph + 1u) & 1u) * kTileBytes;   // gradient tile the next phase reads
#pragma unroll
                for (uint32_t c = 0; c < 4; ++c) {
                    uint4 pk;
                    pk.x = pack_half2(__uint_as_float(v[c * 8 + 0]), __uint_as_float(v[c * 8 + 1])) & m[c * 4 + 0];
                    pk.y = pack_half2(__uint_as_float(v[c * 8 + 2]), __uint_as_float(v[c * 8 + 3])) & m[c * 4 + 1];
                    pk.z = pack_half2(__uint_as_float(v[c * 8 + 4]), __uint_as_float(v[c * 8 + 5])) & m[c * 4 + 2];
                    pk.w = pack_half2(__uint_as_float(v[c * 8 + 6]), __uint_as_float(v[c * 8 + 7])) & m[c * 4 + 3];
                    sts128(tile_chunk_addr(s_out, row, half * 4 + c), pk);
                }
                if (tracer) LNB_TR(tg, 5u, (uint32_t)layer);
                fence_proxy_async();
                fence_before_sync();
                mbar_arrive(bar_ready);
                if (tracer) LNB_TR(tg, 6u, (uint32_t)layer);
                if (!kHead && a.bbuf) {
                    tg_sync(tg);   // every row of the tile written
                    tg_store_tile_rows(gtid, s_out, a.bbuf + ((size_t)(sh.n_hid - layer) * B + row0) * kHid);
                    tg_sync(tg);   // all readers done before a later epilogue rewrites it
                }
                // The `done` we just consumed retires every EARLIER MMA of this group (in-order pipe), in particular
                // the previous phase's weight-gradient MMAs that read saved-activation tile layer + 1: refill it now.
                if (have_next && layer < (int)sh.n_hid) load_act((uint32_t)layer + 1, next, rows_next);
            }
            if (kHead) {
                // ---- geo gradient + density gradient -> g_sig_out row (network.py:173 trunc_exp backward) ----
                mbar_wait(bar_done, par_done);
                par_done ^= 1;
                fence_after_sync();
                if (half == kGeoWin / 32) {
                    uint32_t v[32];
                    tmem_ld32(d_mine, v);
                    tmem_ld_wait();
                    const float h0 = mlp_to_float((unsigned short)(hr.so_lo.x & 0xffffu));
                    const float g0 = hr.gs * a.density_scale * __expf(fminf(fmaxf(h0, -15.f), 15.f));   // activation.py:14-17
                    float o[16];
                    o[0] = g0;
#pragma unroll
                    for (uint32_t k = 1; k < 16; ++k) o[k] = __uint_as_float(v[kGeoIdx + k - 1]);
                    uint4 *dst = reinterpret_cast<uint4 *>(a.g_sig_out + r * kOut);
                    dst[0] = make_uint4(pack_half2(o[0], o[1]), pack_half2(o[2], o[3]), pack_half2(o[4], o[5]),
                                        pack_half2(o[6], o[7]));
                    dst[1] = make_uint4(pack_half2(o[8], o[9]), pack_half2(o[10], o[11]), pack_half2(o[12], o[13]),
                                        pack_half2(o[14], o[15]));
                }
            } else {
                // ---- input-gradient tiles ----
                for (uint32_t t = 0; t < sh.kt_in; ++t) {
                    if (t > 0 && !want_dx) break;
                    mbar_wait(bar_done, par_done);
                    par_done ^= 1;
                    fence_after_sync();
                    if (want_dx) {
                        const uint32_t cols = min(64u, sh.in_dim - t * 64);
                        uint32_t v[32];
                        tmem_ld32(d_mine, v);
                        tmem_ld_wait();
#pragma unroll
                        for (uint32_t c = 0; c < 4; ++c) {
                            const uint32_t col = half * 32 + c * 8;
                            if (col < cols) {
                                uint4 pk;
                                // (the input gradient is ALWAYS fp16: the hash-grid scatter reads it)
                                pk.x = pack_f16x2(__uint_as_float(v[c * 8 + 0]), __uint_as_float(v[c * 8 + 1]));
                                pk.y = pack_f16x2(__uint_as_float(v[c * 8 + 2]), __uint_as_float(v[c * 8 + 3]));
                                pk.z = pack_f16x2(__uint_as_float(v[c * 8 + 4]), __uint_as_float(v[c * 8 + 5]));
                                pk.w = pack_f16x2(__uint_as_float(v[c * 8 + 6]), __uint_as_float(v[c * 8 + 7]));
                                *reinterpret_cast<uint4 *>(a.dX + r * sh.in_dim + t * 64 + col) = pk;
                            }
                        }
                    }
                    if (t + 1 < sh.kt_in && want_dx) {
                        fence_before_sync();
                        mbar_arrive(bar_ready);   // accumulator drained -> next 64 input columns
                    }
                }
            }
            if (tracer) LNB_TR(tg, 7u, 0u);
            // every MMA of this tile has retired (last `done` wait above): activation tile 0 and the X / G tiles can
            // take the next tile
            if (have_next) {
                load_act(0, next, rows_next);
                load_xg(next, hn);
            }
            fence_before_sync();          // orders this tile's TMEM reads before the next tile's MMAs
            hr = hn;
            tile = next;
            have = have_next;
        }
    }

    // ---- flush the tensor-memory weight-gradient accumulators (fp32 atomics into the flat layout) ----
    // UMMA M = 64 puts row m at TMEM lane (m / 16) * 32 + m % 16: warp w (mod 4), lanes 0..15 own rows 16 (w%4) + lane.
    // Tile group 0 does it: its two column halves split the 16-column chunks of every accumulator.
    if (warp < 8) {
        mbar_wait(bar_fin, 0);
        fence_after_sync();
        if (blockIdx.x * kTG < n_tiles) {   // this CTA accumulated something
            const uint32_t half = (warp >> 2) & 1u;
            const uint32_t lane_sel = ((warp & 3u) * 32u) << 16;
            const uint32_t m = (warp & 3u) * 16 + lane;  // valid when lane < 16
            float *w_in = a.wgrad;
            float *w_hid = a.wgrad + sh.w_in_elems;
            float *w_out = w_hid + (size_t)sh.n_hid * kHid * kHid;
            if (half == 0) {
                uint32_t v[16];
                tmem_ld16(d_wout + lane_sel, v);
                tmem_ld_wait();
                if (lane < 16)
#pragma unroll
                    for (uint32_t o = 0; o < (kHead ? 2u : 16u); ++o)   // head: only outputs 0,1 carry gradient
                        atomicAdd(w_out + o * kHid + m, __uint_as_float(v[o]));
            }
            for (uint32_t l = 0; l < sh.n_hid; ++l) {
                uint32_t v[32];
                tmem_ld32(d_whid + 64 * l + lane_sel + half * 32, v);
                tmem_ld_wait();
                if (lane < 16) {
                    float4 *dst = reinterpret_cast<float4 *>(w_hid + (size_t)l * kHid * kHid + m * kHid + half * 32);
#pragma unroll
                    for (uint32_t n = 0; n < 8; ++n)      // 16-byte vector reductions (RED.v4.f32): 4x fewer L2 atomics
                        atomicAdd(dst + n, make_float4(__uint_as_float(v[4 * n]), __uint_as_float(v[4 * n + 1]),
                                                       __uint_as_float(v[4 * n + 2]), __uint_as_float(v[4 * n + 3])));
                }
            }
            const uint32_t in_used = kHead ? a.nfreq + 15 : sh.in_dim;     // head: padding columns have zero gradient
            for (uint32_t c16 = half; c16 < win_cols / 16; c16 += 2) {
                uint32_t v[16];
                tmem_ld16(d_win + 16 * c16 + lane_sel, v);
                tmem_ld_wait();
                if (lane < 16) {
                    float4 *dst = reinterpret_cast<float4 *>(w_in + m * sh.in_dim + c16 * 16);
#pragma unroll
                    for (uint32_t n = 0; n < 4; ++n)      // columns at or beyond in_used hold exact zeros (zero X padding)
                        if (c16 * 16 + 4 * n < in_used)
                            atomicAdd(dst + n, make_float4(__uint_as_float(v[4 * n]), __uint_as_float(v[4 * n + 1]),
                                                           __uint_as_float(v[4 * n + 2]), __uint_as_float(v[4 * n + 3])));
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem, tmem_cols);
}

inline size_t mlp_bwd_smem(const Shape &sh, uint32_t tg) {
    return 1024 + (size_t)sh.kt_in * kWTileBytes + (size_t)sh.n_hid * kWTileBytes + 2048 +
           (size_t)tg * (2 + sh.n_hid + 1 + sh.kt_in) * kTileBytes + 64;
}

template <bool kHead, uint32_t kGeoWin, uint32_t kGeoIdx, uint32_t kTG>
int launch_mlp_bwd_tg(const BwdArgs &a, uint32_t sm_count, cudaStream_t st) {
    const size_t smem = mlp_bwd_smem(a.sh, kTG);
    auto kern = k_mlp_bwd<kHead, kGeoWin, kGeoIdx, kTG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    const uint32_t groups = (a.B / kRows + kTG - 1) / kTG;
    const uint32_t cap = sm_count * ((kTG == 1 && bwd_tmem_cols(a.sh, 1, bwd_win_cols(a.sh, kHead, a.nfreq)) <= 256) ? 2u : 1u);
    kern<<<groups < cap ? groups : cap, bwd_threads(kTG), smem, st>>>(a);
    return LNB_OK;
}

// geo window / index pairs the LiDAR-head mode is instantiated for (kGeoIdx + 15 <= 32)
template <bool kHead, uint32_t kGeoWin, uint32_t kGeoIdx>
int launch_mlp_bwd(const BwdArgs &a, uint32_t sm_count, cudaStream_t st) {
    if (a.sh.n_hid > 2) return LNB_ERR_UNSUPPORTED;                       // three mask registers / shared memory
    const uint32_t win_cols = bwd_win_cols(a.sh, kHead, a.nfreq);
    // Measured on B200 (500 k rows, 64x2 nets): one 544-thread CTA per SM with two tiles in flight beats two
    // independent 288-thread CTAs (109/101 us vs 118/114 us for head / density MLP) - the M = 64 weight-gradient MMAs
    // (both operands from shared memory) occupy the tensor pipe ~150 cycles each, so a second issuing warp buys
    // nothing.  The single-tile variant remains for shapes whose tiles / accumulators do not fit twice.
    if (bwd_tmem_cols(a.sh, 2, win_cols) <= 512 && mlp_bwd_smem(a.sh, 2) <= 226 * 1024)
        return launch_mlp_bwd_tg<kHead, kGeoWin, kGeoIdx, 2>(a, sm_count, st);
    if (bwd_tmem_cols(a.sh, 1, win_cols) > 512 || mlp_bwd_smem(a.sh, 1) > 226 * 1024) return LNB_ERR_UNSUPPORTED;
    return launch_mlp_bwd_tg<kHead, kGeoWin, kGeoIdx, 1>(a, sm_count, st);
}

}  // namespace
}  // namespace lnb
