// One persistent kernel per ray packet: hash-grid gather -> density MLP -> LiDAR head (forward).
//
// What the two-kernel forward (k_grid_fwd -> enc [M,32] in HBM -> k_field_fwd) does in sequence - an L1/L2-gather-bound
// kernel followed by a latency-bound tensor-core kernel, each owning the whole SM while the other's units idle - runs
// here concurrently inside ONE CTA per SM (800 threads), warp-specialised:
//
//   warps  0..15  GATHER     warp <-> level (32 neighbouring samples per gather instruction, the 8 corner rows of TWO
//                            sample groups = 16 loads per lane in flight), results written as fp16 pairs STRAIGHT INTO
//                            the swizzled shared-memory operand tile of the first MLP layer (a ring of kStages tiles,
//                            full/empty mbarriers; coordinates double-buffered in shared memory);
//   warps 16..23  EPILOGUE   two groups of 128 threads (thread = tile row = TMEM lane), one 128-row tile in flight each:
//                            tcgen05.ld accumulator row, 32 columns at a time -> (+per-ray bias) -> ReLU -> fp16 ->
//                            operand tile of the next layer; the saved activations leave the SM as warp-local coalesced
//                            512-byte stores read back from that tile while the tensor core works; sigma = exp(h0),
//                            geo features -> head operand; sigmoid -> (ray-drop, intensity); one mbarrier arrival per
//                            WARP (fence.proxy.async by every lane, __syncwarp, lane 0 arrives);
//   warp   24     MMA        one elected lane issues every tcgen05.mma of the CTA (M128 x N64/N16 x K16, fp32
//                            accumulators in tensor memory, 128 columns per tile slot), polling the slots' mbarriers
//                            and consuming the operand ring strictly in tile order.
//
// MLP weights are staged ONCE per CTA by the TMA engine: `lnb_field_pack_weights` lays the six weight tiles out in global
// memory as the exact shared-memory image (128-byte rows, 16-byte chunks xor-swizzled) and the kernel pulls that image
// with cp.async.bulk (SASS UBLKCP) onto an mbarrier - no per-thread LDGSTS, no register staging.
//
// Numerics are those of k_grid_fwd + k_field_fwd (same helpers, same rounding points): tests compare the two paths
// bit-for-bit on everything the step keeps.  Measured (profiles/r02_fused_forward_diag.txt): at 385 k samples 164 us vs
// 97 + 72 us for the two kernels - the overlap is real but each role only has a share of the SM's warps and registers
// (gather alone 136 us with 16 warps vs 97 us with 48; MLP alone 99 us vs 72 us), so the gain over the sequence is small.
//
// Reference behaviour: gridencoder.cu:95-199 (gather), ffmlp.cu:460-576 (MLP), network.py:162-237 (wiring).
#include <cstdlib>

#include "common.cuh"
#include "grid_common.cuh"
#include "mlp_tiles.cuh"

namespace lnb {
namespace {

constexpr uint32_t kGatherWarps = 16;
constexpr uint32_t kGatherThreads = kGatherWarps * 32;
constexpr uint32_t kGroups = 2;                        // epilogue groups (128 threads, thread = tile row)
// Measured on B200 (385 k samples, profiles/r02_fused_forward_diag.txt): one tile per epilogue group (2 tiles in flight)
// 164 us, two per group (4 in flight) 181 us - the extra tiles in flight only lengthen the gather's wait for a free stage.
#ifndef LNB_FUSED_SLOTS_PER_GROUP
#define LNB_FUSED_SLOTS_PER_GROUP 1
#endif
#ifndef LNB_FUSED_STAGES
#define LNB_FUSED_STAGES 3
#endif
constexpr uint32_t kSlotsPerGroup = LNB_FUSED_SLOTS_PER_GROUP;   // tiles a group keeps in flight, processed phase by phase in turn
constexpr uint32_t kSlots = kGroups * kSlotsPerGroup;  // tiles in flight in the MLP part of one CTA
constexpr uint32_t kGroupThreads = 128;
constexpr uint32_t kEpiWarp0 = kGatherWarps;           // first epilogue warp (multiple of 4: TMEM lane quadrants)
constexpr uint32_t kMmaWarpIdx = kEpiWarp0 + kGroups * 4;
constexpr uint32_t kFusedThreads = (kMmaWarpIdx + 1) * 32;      // 800
constexpr uint32_t kStages = LNB_FUSED_STAGES;         // operand tiles between the gather and the first MLP layer
constexpr uint32_t kTmemColsPerSlot = 128;             // [0,64) hidden accumulator, [64,80) output accumulator
constexpr uint32_t kCoordBufs = 2;

struct FusedShape {
    uint32_t enc_dim;        // L * C (multiple of 16, <= 64)
    uint32_t n_hid_s, n_hid_h;
    uint32_t head_in;        // padded head input width (row length of W_in of the head)
    uint32_t nfreq, geo_tile, geo_off, ks_geo;
};

__host__ __device__ inline uint32_t weight_image_bytes(const FusedShape &fs) {
    return kWTileBytes * (1 + fs.n_hid_s) + 2048 + kWTileBytes * (1 + fs.n_hid_h) + 2048;
}

struct FusedArgs {
    const float *xyz;
    const __half *table;
    const int32_t *offsets;
    uint32_t L;
    float S;
    uint32_t H;
    float2 norm;
    const uint8_t *wimg;
    const int32_t *ray_ids;
    const float *ray_bias;
    uint32_t B;
    const int32_t *n_active;
    FusedShape fs;
    float density_scale;
    __half *enc, *fb_s, *sig_out, *fb_h;
    float *sigma, *rgb;
    uint32_t dbg;     // diagnostics only (LNB_FUSED_DBG, scripts/diag_fused_fwd.py): bit 0 = gather without table loads,
                      // bit 1 = no saved-activation / enc stores, bit 2 = epilogue skips the per-layer math,
                      // bit 3 = gather warps skip the cell / index / blend arithmetic as well (protocol only)
};

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts16h(uint32_t addr, unsigned short v) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint32_t elem_addr(uint32_t tile, uint32_t row, uint32_t col) {
    return tile_chunk_addr(tile, row, col >> 3) + (col & 7u) * 2u;
}
__device__ __forceinline__ void named_bar(uint32_t id, uint32_t n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
// non-blocking phase test, warp-uniform result (the MMA warp stays converged)
__device__ __forceinline__ bool mbar_test_warp(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return __all_sync(0xffffffffu, done) != 0;
}

// -----------------------------------------------------------------------------------------------------
// weight image: [Ws_in | Ws_hid x n | Ws_out (2 KB) | Wh_geo | Wh_hid x n | Wh_out (2 KB)], every tile in the
// one operand layout of tcgen05.cuh (tile_chunk_addr).  Chunks that hold no weight stay zero (the image is cleared
// when it is allocated and only valid chunks are ever written).
// -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pack_field_weights(const __half *__restrict__ Ws, const __half *__restrict__ Wh, FusedShape fs, uint8_t *__restrict__ img) {
    auto put = [&](uint32_t tile_off, const __half *src, uint32_t rows, uint32_t cols, uint32_t ld) {
        const uint32_t cpr = (cols + 7) / 8;
        for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < rows * cpr; q += gridDim.x * blockDim.x) {
            const uint32_t r = q / cpr, c = q - r * cpr;
            const uint4 v = *reinterpret_cast<const uint4 *>(src + (size_t)r * ld + c * 8);
            *reinterpret_cast<uint4 *>(img + tile_chunk_addr(tile_off, r, c)) = v;
        }
    };
    uint32_t off = 0;
    const uint32_t ws_in_elems = kHid * fs.enc_dim, wh_in_elems = kHid * fs.head_in;
    put(off, Ws, kHid, fs.enc_dim, fs.enc_dim);
    off += kWTileBytes;
    for (uint32_t l = 0; l < fs.n_hid_s; ++l, off += kWTileBytes) put(off, Ws + ws_in_elems + (size_t)l * kHid * kHid, kHid, kHid, kHid);
    put(off, Ws + ws_in_elems + (size_t)fs.n_hid_s * kHid * kHid, kOut, kHid, kHid);
    off += 2048;
    // head: only the 64-column tile of W_in that holds the geo columns (the direction columns act through ray_bias)
    put(off, Wh + fs.geo_tile * 64, kHid, min(64u, fs.head_in - fs.geo_tile * 64), fs.head_in);
    off += kWTileBytes;
    for (uint32_t l = 0; l < fs.n_hid_h; ++l, off += kWTileBytes) put(off, Wh + wh_in_elems + (size_t)l * kHid * kHid, kHid, kHid, kHid);
    put(off, Wh + wh_in_elems + (size_t)fs.n_hid_h * kHid * kHid, kOut, kHid, kHid);
}

// this thread's accumulator row (64 fp32 columns) -> (+bias) -> ReLU -> fp16 -> the row of the next layer's operand tile,
// 32 columns at a time (keeps the epilogue at 72 registers: no setmaxnreg redistribution needed; the second tcgen05.ld's
// latency is covered by the group's other tile)
__device__ __forceinline__ void row_relu_to_tile(uint32_t d_row, const float4 *__restrict__ bias, uint32_t s_h, uint32_t row) {
#pragma unroll
    for (uint32_t half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(d_row + 32 * half, v);
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
            uint4 pk;
            pk.x = pack_half2(fmaxf(f[0], 0.f), fmaxf(f[1], 0.f));
            pk.y = pack_half2(fmaxf(f[2], 0.f), fmaxf(f[3], 0.f));
            pk.z = pack_half2(fmaxf(f[4], 0.f), fmaxf(f[5], 0.f));
            pk.w = pack_half2(fmaxf(f[6], 0.f), fmaxf(f[7], 0.f));
            sts128(tile_chunk_addr(s_h, row, half * 4 + c), pk);
        }
    }
}

// (Measured and rejected: gathering with 4-byte cp.async into a per-warp shared-memory ring - no register per load in
// flight, 24+ loads per lane outstanding - made the gather 2.5x SLOWER than plain LDG (257 vs ~100 us at 385 k samples):
// LDGSTS of 4-byte elements is processed far below the LSU's gather rate.)
__global__ void __launch_bounds__(kFusedThreads, 1)
k_field_fused_fwd(const FusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const FusedShape fs = a.fs;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // weight image (same order as k_pack_field_weights)
    const uint32_t s_ws_in = sbase;
    const uint32_t s_ws_hid = s_ws_in + kWTileBytes;
    const uint32_t s_ws_out = s_ws_hid + fs.n_hid_s * kWTileBytes;
    const uint32_t s_wh_geo = s_ws_out + 2048;
    const uint32_t s_wh_hid = s_wh_geo + kWTileBytes;
    const uint32_t s_wh_out = s_wh_hid + fs.n_hid_h * kWTileBytes;
    const uint32_t wbytes = weight_image_bytes(fs);
    const uint32_t s_x0 = sbase + wbytes;                          // kStages operand tiles written by the gather
    const uint32_t s_h0 = s_x0 + kStages * kTileBytes;             // one activation operand tile per slot
    const uint32_t s_in0 = s_h0 + kSlots * kTileBytes;             // kCoordBufs x [128][3] coordinates in [0,1]
    const uint32_t s_bar = s_in0 + kCoordBufs * kRows * 3 * 4;
    const uint32_t bar_xfull = s_bar, bar_xempty = bar_xfull + 8 * kStages, bar_ready = bar_xempty + 8 * kStages;
    const uint32_t bar_done = bar_ready + 8 * kSlots, bar_w = bar_done + 8 * kSlots, s_slot = bar_w + 8;
    float *s_in = reinterpret_cast<float *>(smem_raw + (s_in0 - smem_u32(smem_raw)));

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    constexpr uint32_t kTmemCols = 512;
    static_assert(kSlots * kTmemColsPerSlot <= kTmemCols, "tensor memory");

    if (warp == kMmaWarpIdx) tmem_alloc(s_slot, kTmemCols);
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < kStages; ++s) {
            mbar_init(bar_xfull + 8 * s, kGatherWarps);           // one arrival per gather warp
            mbar_init(bar_xempty + 8 * s, 1 + 4);                 // tcgen05.commit + one arrival per epilogue warp
        }
        for (uint32_t q = 0; q < kSlots; ++q) {
#ifdef LNB_FUSED_THREAD_ARRIVE
            mbar_init(bar_ready + 8 * q, kGroupThreads);
#else
            mbar_init(bar_ready + 8 * q, 4);                      // one arrival per epilogue warp of the owning group
#endif
            mbar_init(bar_done + 8 * q, 1);
        }
        mbar_init(bar_w, 1);
        mbar_init_fence();
        // the TMA engine pulls the pre-laid-out weight image: one transaction count, <= 16 KB per bulk copy
        mbar_expect_tx(bar_w, wbytes);
        for (uint32_t o = 0; o < wbytes; o += 16384u)
            bulk_g2s(sbase + o, a.wimg + o, min(16384u, wbytes - o), bar_w);
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = lds32(s_slot);

    const uint32_t n_tiles = active_rows(a.B, a.n_active) / kRows;
    const uint32_t n_my = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t n_steps = fs.n_hid_s + 2 + fs.n_hid_h + 2;
    const uint32_t ks_in = fs.enc_dim / 16;

    if (warp < kGatherWarps) {
        // ======================= GATHER: warp <-> level =======================
        // (Registers: 25 warps -> 72 per thread for every role.  Redistributing with setmaxnreg was tried both ways -
        // gather 48 / epilogue 120 with a 64-column epilogue, gather 80 / epilogue 56 - and bought nothing: the epilogue
        // works on 32 accumulator columns at a time and the gather keeps 16 loads per lane in flight within 72.)
        const uint32_t tid = threadIdx.x;
        constexpr uint32_t D = 3, C = 2;
        const uint32_t nlv = warp < a.L ? (a.L - warp + kGatherWarps - 1) / kGatherWarps : 0;   // levels of this warp
        // level-uniform quantities of the warp's first level, once per kernel
        LevelGeo g0 = {};
        LevelIndex<D> li0 = {};
        if (nlv > 0) {
            g0 = level_geo(a.offsets, warp, a.S, a.H);
            li0 = level_index<D>(g0, 0u, false);
        }
        auto stage_coords = [&](uint32_t k) {
            if (tid < kRows * 3 && k < n_my) {
                const size_t tile = blockIdx.x + (size_t)k * gridDim.x;
                float x = __ldg(a.xyz + tile * kRows * 3 + tid);
                if (a.norm.x != 0.f) x = (x + a.norm.x) * a.norm.y;
                s_in[(k & 1u) * kRows * 3 + tid] = x;
            }
        };
        // the cell of this lane's sample `sl` of the tile whose coordinates are at `in`
        auto locate_sample = [&](const float *in, uint32_t sl, const LevelGeo &g) -> Cell<D> {
            float v[D];
            bool inside = true;
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) {
                v[d] = in[sl * D + d];
                if (v[d] < 0 || v[d] > 1) inside = false;
            }
            return locate_unit<D>(v, inside, g, false, 0u);
        };
        // all 8 corner rows of a cell (raw half2 words): 8 independent loads in flight
        auto load_corners = [&](const Cell<D> &cell, const LevelGeo &g, const LevelIndex<D> &li, const __half *tab,
                                uint32_t (&raw)[8]) {
            const CornerRows<D> cr(li, cell.base);
#pragma unroll
            for (uint32_t corner = 0; corner < 8; ++corner) {
                uint32_t row;
                if (li.generic) {
                    uint32_t pp[D];
#pragma unroll
                    for (uint32_t d = 0; d < D; ++d) pp[d] = cell.base[d] + ((corner >> d) & 1u);
                    row = cell_row<D>(pp, 0u, false, g);
                } else {
                    row = cr.row(corner);
                }
                raw[corner] = (cell.inside && !(a.dbg & 1u)) ? __ldg(reinterpret_cast<const uint32_t *>(tab + (size_t)row * C)) : 0u;
            }
        };
        // sum over the 8 corners, accumulated in fp16 in corner order exactly like interp_corners() / the reference
        // (gridencoder.cu:173-199); 0 outside the unit cube
        auto blend = [&](const Cell<D> &cell, const uint32_t (&raw)[8]) -> uint32_t {
            __half r0 = __float2half_rn(0.f), r1 = r0;
            if (cell.inside && !(a.dbg & 1u)) {
#pragma unroll
                for (uint32_t corner = 0; corner < 8; ++corner) {
                    float w = 1;
#pragma unroll
                    for (uint32_t d = 0; d < D; ++d) {
                        if ((corner & (1u << d)) == 0) w *= 1 - cell.frac[d];
                        else w *= cell.frac[d];
                    }
                    const __half2 hv = *reinterpret_cast<const __half2 *>(&raw[corner]);
                    r0 = __float2half_rn(__half2float(r0) + w * __low2float(hv));
                    r1 = __float2half_rn(__half2float(r1) + w * __high2float(hv));
                }
            }
            // the table and its interpolation are fp16 (reference semantics); the MLP operand is this unit's element type
            return (uint32_t)mlp_from_float(__half2float(r0)) | ((uint32_t)mlp_from_float(__half2float(r1)) << 16);
        };

        stage_coords(0);
        for (uint32_t k = 0; k < n_my; ++k) {
            const uint32_t stage = k % kStages, use = k / kStages;
            named_bar(1, kGatherThreads);                   // coordinates of tile k complete; everybody is done with tile k - 1
            stage_coords(k + 1);
            if (use > 0) mbar_wait(bar_xempty + 8 * stage, (use - 1) & 1u);
            const uint32_t s_x = s_x0 + stage * kTileBytes;
            const float *in = s_in + (k & 1u) * kRows * 3;
            for (uint32_t lv = 0; lv < nlv; ++lv) {
                const uint32_t level = warp + lv * kGatherWarps;
                LevelGeo g = g0;
                LevelIndex<D> li = li0;
                if (lv != 0) {
                    g = level_geo(a.offsets, level, a.S, a.H);
                    li = level_index<D>(g, 0u, false);
                }
                const __half *__restrict__ tab = a.table + (size_t)g.table_offset * C;
                // two sample groups at a time: 16 gathers in flight per lane (the 16 gather warps have to cover the
                // L2 latency that 48 warps cover in the stand-alone encoder kernel)
#pragma unroll
                for (uint32_t pair = 0; pair < kRows / 64; ++pair) {
                    if (a.dbg & 8u) break;
                    const uint32_t sl_a = pair * 64 + lane, sl_b = sl_a + 32;
                    const Cell<D> ca = locate_sample(in, sl_a, g), cb = locate_sample(in, sl_b, g);
                    uint32_t ra[8], rb[8];
                    load_corners(ca, g, li, tab, ra);
                    load_corners(cb, g, li, tab, rb);
                    sts32(elem_addr(s_x, sl_a, level * C), blend(ca, ra));
                    sts32(elem_addr(s_x, sl_b, level * C), blend(cb, rb));
                }
            }
            fence_proxy_async();                            // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_xfull + 8 * stage);
        }
    } else if (warp < kMmaWarpIdx) {
        // ======================= EPILOGUE groups: thread = tile row, two tiles in flight per group =======================
        const uint32_t g = (warp - kEpiWarp0) >> 2;
        const uint32_t wq = warp & 3u;                     // TMEM lane quadrant of this warp = rows 32 wq .. 32 wq + 31
        const uint32_t row = wq * 32u + lane;
        const uint32_t lane_sel = (wq * 32u) << 16;
        uint32_t par[kSlotsPerGroup] = {};
        size_t row0[kSlotsPerGroup] = {};
        uint32_t rid[kSlotsPerGroup] = {};

        // warp-local coalesced copy of this warp's 32 rows of a swizzled tile to global memory (row pitch `pitch_h`
        // halves, `cpr` 16-byte chunks per row): every instruction writes 512 contiguous bytes
        auto copy_rows_out = [&](uint32_t tile, __half *dst_row0, uint32_t cpr) {
            if (cpr == 8) {                                  // 64-wide rows (saved activations): no division, 8 copies
#pragma unroll
                for (uint32_t j = 0; j < 8; ++j) {
                    const uint32_t q = lane + 32 * j, rr = wq * 32u + (q >> 3), c = q & 7u;
                    *reinterpret_cast<uint4 *>(dst_row0 + ((size_t)rr * 8 + c) * 8) = lds128(tile_chunk_addr(tile, rr, c));
                }
                return;
            }
            for (uint32_t q = lane; q < 32 * cpr; q += 32) {
                const uint32_t rr = wq * 32u + q / cpr, c = q % cpr;
                *reinterpret_cast<uint4 *>(dst_row0 + ((size_t)rr * cpr + c) * 8) = lds128(tile_chunk_addr(tile, rr, c));
            }
        };
#ifdef LNB_FUSED_THREAD_ARRIVE
        auto publish = [&](uint32_t bar) {
            fence_proxy_async();
            fence_before_sync();
            mbar_arrive(bar);
        };
        auto arrive_warp = [&](uint32_t bar) { mbar_arrive(bar); };
#else
        auto publish = [&](uint32_t bar) {                  // this warp's rows of an operand tile are in place
            fence_proxy_async();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        };
        auto arrive_warp = [&](uint32_t bar) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        };
#endif

        for (uint32_t q = 0; q < kSlotsPerGroup; ++q)       // tensor memory of the group's slots is free
            if (g * kSlotsPerGroup + q < n_my) arrive_warp(bar_ready + 8 * (g * kSlotsPerGroup + q));

        for (uint32_t k0 = g * kSlotsPerGroup; k0 < n_my; k0 += kSlots) {
            const uint32_t nt = min(kSlotsPerGroup, n_my - k0);
            // ---- per tile: gathered features -> enc (the backward pass needs them), release the operand tile ----
            for (uint32_t t = 0; t < nt; ++t) {
                const uint32_t k = k0 + t, stage = k % kStages, use = k / kStages;
                row0[t] = ((size_t)blockIdx.x + (size_t)k * gridDim.x) * kRows;
                rid[t] = (uint32_t)__ldg(a.ray_ids + row0[t] + row);
                mbar_wait(bar_xfull + 8 * stage, use & 1u);
                if (!(a.dbg & 2u)) copy_rows_out(s_x0 + stage * kTileBytes, a.enc + row0[t] * fs.enc_dim, fs.enc_dim / 8);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_xempty + 8 * stage);
                // this row's 256 B of the per-ray head bias, needed four layers from now: pull the lines into L1
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a.ray_bias + (size_t)rid[t] * kHid));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a.ray_bias + (size_t)rid[t] * kHid + 32));
            }
            // ---- the network, phase by phase, the group's tiles in turn (one tile's MMA runs under the other's epilogue) ----
            for (uint32_t ph = 0; ph < n_steps; ++ph) {
                for (uint32_t t = 0; t < nt; ++t) {
                    const uint32_t slot = g * kSlotsPerGroup + t;
                    const uint32_t d_hid = tmem + slot * kTmemColsPerSlot + lane_sel, d_out = d_hid + 64;
                    const uint32_t s_h = s_h0 + slot * kTileBytes;
                    const uint32_t ready = bar_ready + 8 * slot, done = bar_done + 8 * slot;
                    const size_t r = row0[t] + row;
                    mbar_wait(done, par[t]);
                    par[t] ^= 1;
                    fence_after_sync();
                    const bool sigma_layer = ph <= fs.n_hid_s;
                    const bool head_layer = ph >= fs.n_hid_s + 2 && ph <= fs.n_hid_s + 2 + fs.n_hid_h;
                    if (sigma_layer || head_layer) {
                        // hidden layer: accumulator row -> (+per-ray bias) -> ReLU -> fp16 -> operand tile of the next layer
                        const uint32_t layer = sigma_layer ? ph : ph - (fs.n_hid_s + 2);
                        const float4 *bias = (head_layer && layer == 0)
                                                 ? reinterpret_cast<const float4 *>(a.ray_bias + (size_t)rid[t] * kHid) : nullptr;
                        __syncwarp();                        // the previous layer's copy-out reads of this warp's rows are done
                        if (!(a.dbg & 4u)) row_relu_to_tile(d_hid, bias, s_h, row);
                        publish(ready);
                        if (!(a.dbg & 6u)) {                 // saved activations leave coalesced while the tensor core works
                            __half *fb = sigma_layer ? a.fb_s : a.fb_h;
                            copy_rows_out(s_h, fb + ((size_t)layer * a.B + row0[t]) * kHid, 8);
                        }
                    } else if (ph == fs.n_hid_s + 1) {
                        // density output: sig_out (fp16, kept for backward), sigma = exp(h0) * scale, geo -> head operand
                        uint32_t v[16];
                        tmem_ld16(d_out, v);
                        tmem_ld_wait();
                        __align__(16) unsigned short hv[16];
#pragma unroll
                        for (uint32_t j = 0; j < 16; ++j) hv[j] = mlp_from_float(__uint_as_float(v[j]));
                        const uint32_t *pw = reinterpret_cast<const uint32_t *>(hv);
                        uint4 *dst = reinterpret_cast<uint4 *>(a.sig_out + r * kOut);
                        dst[0] = make_uint4(pw[0], pw[1], pw[2], pw[3]);
                        dst[1] = make_uint4(pw[4], pw[5], pw[6], pw[7]);
                        a.sigma[r] = __expf(mlp_to_float(hv[0])) * a.density_scale;    // activation.py:6-20 (forward)
                        __syncwarp();                        // copy-out of the last hidden layer is done with these rows
                        for (uint32_t c = 0; c < 2 * fs.ks_geo; ++c) sts128(tile_chunk_addr(s_h, row, c), make_uint4(0, 0, 0, 0));
#pragma unroll
                        for (uint32_t j = 1; j < 16; ++j) sts16h(elem_addr(s_h, row, fs.geo_off + j - 1), hv[j]);
                        publish(ready);
                    } else {
                        // head output -> (ray-drop, intensity) = sigmoid(fp16(h[0:2]))   (network.py:230)
                        uint32_t v[16];
                        tmem_ld16(d_out, v);
                        tmem_ld_wait();
                        const float x0 = mlp_to_float(mlp_from_float(__uint_as_float(v[0])));
                        const float x1 = mlp_to_float(mlp_from_float(__uint_as_float(v[1])));
                        reinterpret_cast<float2 *>(a.rgb)[r] = make_float2(1.f / (1.f + __expf(-x0)), 1.f / (1.f + __expf(-x1)));
                        fence_before_sync();                 // this tile's TMEM reads are ordered before the next tile's MMAs
                        if (k0 + t + kSlots < n_my) arrive_warp(ready);                // the slot's next tile may start
                    }
                }
            }
        }
    } else {
        // ======================= MMA warp (converged; one elected lane issues) =======================
        mbar_wait_warp(bar_w, 0);                           // weight image landed (TMA transaction bytes complete)
        uint32_t kk[kSlots], step[kSlots], par[kSlots];
        uint32_t left = 0;
#pragma unroll
        for (uint32_t q = 0; q < kSlots; ++q) {
            kk[q] = q, step[q] = 0, par[q] = 0;
            if (q < n_my) ++left;
        }
        uint32_t spins = 0;
        uint32_t next_x = 0;          // CTA-local index of the next tile whose first layer may be issued
        while (left > 0) {
            bool progressed = false;
#pragma unroll
            for (uint32_t q = 0; q < kSlots; ++q) {
                if (kk[q] >= n_my) continue;
                // The operand ring is consumed strictly in tile order.  (A parity test is only meaningful for the NEXT
                // completion of a barrier: asking for use u of a stage before use u - 1 has completed succeeds at once -
                // four slots over three stages would otherwise start tile 3 on the stage tile 0 is still being gathered
                // into.)
                if (step[q] == 0 && kk[q] != next_x) continue;
                if (!mbar_test_warp(bar_ready + 8 * q, par[q])) continue;
                const uint32_t stage = kk[q] % kStages, use = kk[q] / kStages;
                if (step[q] == 0 && !mbar_test_warp(bar_xfull + 8 * stage, use & 1u)) continue;
                par[q] ^= 1;
                fence_after_sync();
                const uint32_t d_hid = tmem + q * kTmemColsPerSlot, d_out = d_hid + 64;
                const uint32_t s_h = s_h0 + q * kTileBytes;
                const uint32_t st = step[q];
                if (st == 0) {
                    issue_kmajor(d_hid, s_x0 + stage * kTileBytes, s_ws_in, ks_in, kIdescFwdHid, false);
                } else if (st <= fs.n_hid_s) {
                    issue_kmajor(d_hid, s_h, s_ws_hid + (st - 1) * kWTileBytes, 4, kIdescFwdHid, false);
                } else if (st == fs.n_hid_s + 1) {
                    issue_kmajor(d_out, s_h, s_ws_out, 4, kIdescFwdOut, false);
                } else if (st == fs.n_hid_s + 2) {
                    issue_kmajor(d_hid, s_h, s_wh_geo, fs.ks_geo, kIdescFwdHid, false);
                } else if (st <= fs.n_hid_s + 2 + fs.n_hid_h) {
                    issue_kmajor(d_hid, s_h, s_wh_hid + (st - fs.n_hid_s - 3) * kWTileBytes, 4, kIdescFwdHid, false);
                } else {
                    issue_kmajor(d_out, s_h, s_wh_out, 4, kIdescFwdOut, false);
                }
                mma_commit_elect(bar_done + 8 * q);
                if (st == 0) {
                    mma_commit_elect(bar_xempty + 8 * stage);              // the operand tile has been consumed
                    ++next_x;
                }
                if (++step[q] == n_steps) {
                    step[q] = 0;
                    kk[q] += kSlots;
                    if (kk[q] >= n_my) --left;
                }
                progressed = true;
            }
            if (progressed) spins = 0;
            else if (++spins > (1u << 24)) __trap();        // a protocol bug becomes a kernel error, not a hung GPU
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarpIdx) tmem_dealloc(tmem, kTmemCols);
}

int make_fused_shape(uint32_t enc_dim, uint32_t sigma_layers, uint32_t head_in_pad, uint32_t head_layers, uint32_t degree,
                     uint32_t hidden, FusedShape *fs) {
    if (hidden != kHid) return LNB_ERR_UNSUPPORTED;
    if (enc_dim == 0 || enc_dim % 16 != 0 || enc_dim > 64) return LNB_ERR_UNSUPPORTED;
    if (head_in_pad == 0 || head_in_pad % 16 != 0 || head_in_pad > 128) return LNB_ERR_UNSUPPORTED;
    if (sigma_layers < 2 || head_layers < 2 || sigma_layers > 4 || head_layers > 4) return LNB_ERR_UNSUPPORTED;
    fs->enc_dim = enc_dim;
    fs->n_hid_s = sigma_layers - 1;
    fs->n_hid_h = head_layers - 1;
    fs->head_in = head_in_pad;
    if (!dir_code_valid(degree)) return LNB_ERR_UNSUPPORTED;
    fs->nfreq = dir_code_width(degree);
    if (fs->nfreq + 15 > head_in_pad) return LNB_ERR_INVALID_ARGUMENT;
    fs->geo_tile = fs->nfreq / 64;
    fs->geo_off = fs->nfreq % 64;
    if (fs->geo_off + 15 > 64) return LNB_ERR_UNSUPPORTED;      // geo columns must sit inside one 64-column tile
    fs->ks_geo = (fs->geo_off + 15 + 15) / 16;
    return LNB_OK;
}

size_t fused_smem_bytes(const FusedShape &fs) {
    return 1024 + weight_image_bytes(fs) + (size_t)(kStages + kSlots) * kTileBytes + kCoordBufs * kRows * 3 * 4 +
           8 * (2 * kStages + 2 * kSlots + 1) + 16;
}

int sm_count_fused() {
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
#define lnb_field_fused_weight_bytes lnb_field_fused_weight_bytes_bf16
#define lnb_field_pack_weights lnb_field_pack_weights_bf16
#define lnb_field_fused_forward lnb_field_fused_forward_bf16
#endif

extern "C" {

size_t lnb_field_fused_weight_bytes(uint32_t enc_dim, uint32_t sigma_layers, uint32_t head_in_pad, uint32_t head_layers,
                                    uint32_t degree, uint32_t hidden) {
    FusedShape fs;
    if (make_fused_shape(enc_dim, sigma_layers, head_in_pad, head_layers, degree, hidden, &fs) != LNB_OK) return 0;
    if (fused_smem_bytes(fs) > 226 * 1024) return 0;
    return weight_image_bytes(fs);
}

int lnb_field_pack_weights(const void *w_sigma, const void *w_head, uint32_t enc_dim, uint32_t sigma_layers,
                           uint32_t head_in_pad, uint32_t head_layers, uint32_t degree, uint32_t hidden, void *image,
                           lnb_stream_t stream) {
    if (!w_sigma || !w_head || !image) return LNB_ERR_INVALID_ARGUMENT;
    FusedShape fs;
    int rc = make_fused_shape(enc_dim, sigma_layers, head_in_pad, head_layers, degree, hidden, &fs);
    if (rc != LNB_OK) return rc;
    k_pack_field_weights<<<8, 256, 0, as_stream(stream)>>>(static_cast<const __half *>(w_sigma), static_cast<const __half *>(w_head),
                                                          fs, static_cast<uint8_t *>(image));
    count_launch();
    return launch_status();
}

int lnb_field_fused_forward(const float *xyzs, const void *table, const int32_t *offsets, uint32_t L, uint32_t C, float S,
                            uint32_t H, float in_bound, const void *weight_image, const int32_t *ray_ids,
                            const float *ray_bias, uint32_t M, uint32_t sigma_layers, uint32_t head_in_pad,
                            uint32_t head_layers, uint32_t degree, uint32_t hidden, float density_scale, void *enc,
                            void *fb_sigma, void *sig_out, float *sigma, void *fb_head, float *rgb,
                            const int32_t *n_active, lnb_stream_t stream) {
    if (!xyzs || !table || !offsets || !weight_image || !ray_ids || !ray_bias || !enc || !fb_sigma || !sig_out || !sigma ||
        !fb_head || !rgb)
        return LNB_ERR_INVALID_ARGUMENT;
    if (M % kRows != 0 || L == 0) return LNB_ERR_INVALID_ARGUMENT;
    if (C != 2) return LNB_ERR_UNSUPPORTED;
    FusedShape fs;
    int rc = make_fused_shape(L * C, sigma_layers, head_in_pad, head_layers, degree, hidden, &fs);
    if (rc != LNB_OK) return rc;
    const size_t smem = fused_smem_bytes(fs);
    if (smem > 226 * 1024) return LNB_ERR_UNSUPPORTED;
    if (M == 0) return LNB_OK;
    cudaError_t e = cudaFuncSetAttribute(k_field_fused_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    FusedArgs a = {};
    a.xyz = xyzs;
    a.table = static_cast<const __half *>(table);
    a.offsets = offsets;
    a.L = L;
    a.S = S;
    a.H = H;
    a.norm = in_bound > 0.f ? make_float2(in_bound, 1.0f / (2.0f * in_bound)) : make_float2(0.f, 0.f);
    a.wimg = static_cast<const uint8_t *>(weight_image);
    a.ray_ids = ray_ids;
    a.ray_bias = ray_bias;
    a.B = M;
    a.n_active = n_active;
    a.fs = fs;
    a.density_scale = density_scale;
    a.enc = static_cast<__half *>(enc);
    a.fb_s = static_cast<__half *>(fb_sigma);
    a.sig_out = static_cast<__half *>(sig_out);
    a.fb_h = static_cast<__half *>(fb_head);
    a.sigma = sigma;
    a.rgb = rgb;
    {
        static const uint32_t dbg = [] { const char *e = getenv("LNB_FUSED_DBG"); return e ? (uint32_t)atoi(e) : 0u; }();
        a.dbg = dbg;
        if (const char *e = getenv("LNB_FUSED_DBG_LIVE")) a.dbg = (uint32_t)atoi(e);      // re-read per call (diagnostics)
    }
    const uint32_t tiles = M / kRows;
    const uint32_t cap = (uint32_t)sm_count_fused();
    k_field_fused_fwd<<<tiles < cap ? tiles : cap, kFusedThreads, smem, as_stream(stream)>>>(a);
    count_launch();
    return launch_status();
}

}  // extern "C"
