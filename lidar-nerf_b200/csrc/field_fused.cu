// One persistent kernel per ray packet: hash-grid gather -> density MLP -> LiDAR head (forward).
//
// What the two-kernel forward (k_grid_fwd -> enc [M,32] in HBM -> k_field_fwd) does in sequence - an L1/L2-gather-bound
// kernel followed by a latency-bound tensor-core kernel, each owning the whole SM while the other's units idle - runs
// here concurrently inside ONE CTA per SM (832 threads), warp-specialised:
//
//   warps  0..15  GATHER     warp <-> level (32 neighbouring samples per gather instruction, the 8 corner rows of TWO
//                            sample groups = 16 loads per lane in flight; dense / power-of-two-hash / generic indexing
//                            chosen once per level, not per corner), results written as fp16 pairs STRAIGHT INTO the
//                            swizzled shared-memory operand tile of the first MLP layer (a ring of kStages tiles,
//                            full/empty mbarriers).  Sample positions arrive through a TMA ring (cp.async.bulk of
//                            1.5 KB per tile onto an mbarrier, two tiles ahead): no CTA-wide barrier in the tile loop;
//   warps 16..23  EPILOGUE   two groups of 128 threads (thread = tile row = TMEM lane), one 128-row tile in flight each:
//                            tcgen05.ld accumulator row, 16 columns per load with three loads in flight -> (+per-ray
//                            bias) -> ReLU -> fp16 -> operand tile of the next layer; the saved activations leave the SM
//                            through TMA tensor stores (cp.async.bulk.tensor.2d of each warp's 32 rows, un-swizzled by
//                            the copy engine) while the tensor core works; sigma = exp(h0), geo features -> head
//                            operand; sigmoid -> (ray-drop, intensity); one mbarrier arrival per WARP
//                            (fence.proxy.async by every lane, __syncwarp, lane 0 arrives);
//   warps 24..25  MMA        one warp per tile slot; an elected lane issues the slot's tcgen05.mma chain (M128 x N64/N16 x
//                            K16, fp32 accumulators in tensor memory, 128 columns per slot) in program order behind
//                            blocking mbarrier waits.
//
// MLP weights are staged ONCE per CTA by the TMA engine: `lnb_field_pack_weights` lays the six weight tiles out in global
// memory as the exact shared-memory image (128-byte rows, 16-byte chunks xor-swizzled) and the kernel pulls that image
// with cp.async.bulk (SASS UBLKCP) onto an mbarrier - no per-thread LDGSTS, no register staging.
//
// Numerics are those of k_grid_fwd + k_field_fwd (same helpers, same rounding points): tests compare the two paths
// bit-for-bit on everything the step keeps, with every level forced through the generic indexing path as a third leg.
// Measured (profiles/r02_fused_forward_diag.txt, r02_fused_forward_timeline.txt): at 385 k samples 150 us vs 97 + 72 us
// for the two kernels; gather alone 109 us, MLP alone 80 us - the two halves compete for the LSU / L1 / L2 path (the saved
// activations cost 28 us of L2 bandwidth the gather would otherwise have), which is what is left between 150 and 109.
//
// Reference behaviour: gridencoder.cu:95-199 (gather), ffmlp.cu:460-576 (MLP), network.py:162-237 (wiring).
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "grid_common.cuh"
#include "mlp_tiles.cuh"

namespace lnb {
struct FieldL2Window {
    const void *base;
    size_t bytes;
};
#ifndef LNB_BF16
FieldL2Window g_field_l2_window = {nullptr, 0};
#else
extern FieldL2Window g_field_l2_window;
#endif

namespace {

constexpr uint32_t kGatherWarps = 16;
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
constexpr uint32_t kEpiWarp0 = kGatherWarps;           // first epilogue warp (multiple of 4: TMEM lane quadrants)
constexpr uint32_t kMmaWarpIdx = kEpiWarp0 + kGroups * 4;
constexpr uint32_t kFusedThreads = (kMmaWarpIdx + kSlots) * 32;      // 832: one MMA warp per tile slot
constexpr uint32_t kStages = LNB_FUSED_STAGES;         // operand tiles between the gather and the first MLP layer
constexpr uint32_t kTmemColsPerSlot = 128;             // [0,64) hidden accumulator, [64,80) output accumulator
constexpr uint32_t kCoordBufs = 4;                    // ring of [128][3] coordinate blocks filled by the TMA engine

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
                      // bit 3 = gather warps skip the cell / index / blend arithmetic as well (protocol only),
                      // bit 4 = every level takes the generic indexing path (tests: it must give the same bits)
};

// Optional in-kernel timeline (diagnostic builds only: build.py --trace, scripts/diag_fwd_trace.py): one thread of each
// role of CTA 0 records (event, SM clock) pairs into its own lane of a global buffer.  Compiled out of the shipped library.
#ifdef LNB_TRACE
__device__ unsigned long long g_fwd_trace[4][4096];
#define FTR(role, ev, arg)                                                                                          \
    do {                                                                                                            \
        if (tr_on && tr_i < 4096u)                                                                                  \
            g_fwd_trace[role][tr_i++] = ((unsigned long long)((((ev) & 255u) << 8) | ((arg) & 255u)) << 44) |       \
                                        ((unsigned long long)clock64() & ((1ull << 44) - 1));                      \
    } while (0)
#define FTR_DECL(cond) const bool tr_on = (cond); uint32_t tr_i = 0; (void)tr_on; (void)tr_i
#else
#define FTR(role, ev, arg) do {} while (0)
#define FTR_DECL(cond) do {} while (0)
#endif

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t elem_addr(uint32_t tile, uint32_t row, uint32_t col) {
    return tile_chunk_addr(tile, row, col >> 3) + (col & 7u) * 2u;
}
// TMA tensor store (SASS UTMASTG): a [rows x 64-half] box of a 128-byte-swizzled shared-memory tile -> row-major global
// memory, un-swizzled by the copy engine; no LSU instruction, no register staging.  Issued by ONE lane.
__device__ __forceinline__ void tma_store_rows(const CUtensorMap *tm, uint32_t smem_src, uint32_t row) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(smem_src),
                 "r"(0), "r"(row)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

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

// this thread's accumulator row (64 fp32 columns) -> (+bias) -> ReLU -> fp16 -> the row of the next layer's operand tile.
// 16 columns per tcgen05.ld, the NEXT load issued before the current 16 columns are processed (only the first load's
// latency is exposed; 32 accumulator registers live, the epilogue stays within the 72 registers of a 25-warp CTA).
// `bias` = this row's 64 fp32 bias values (global memory), nullptr = none
__device__ __forceinline__ void relu_chunk16(const uint32_t (&v)[16], const float4 *__restrict__ bias, uint32_t q, uint32_t s_h,
                                             uint32_t row) {
#pragma unroll
    for (uint32_t c = 0; c < 2; ++c) {
        float f[8];
#pragma unroll
        for (uint32_t e = 0; e < 8; ++e) f[e] = __uint_as_float(v[c * 8 + e]);
        if (bias) {
            const float4 b0 = __ldg(bias + q * 4 + c * 2), b1 = __ldg(bias + q * 4 + c * 2 + 1);
            f[0] += b0.x, f[1] += b0.y, f[2] += b0.z, f[3] += b0.w;
            f[4] += b1.x, f[5] += b1.y, f[6] += b1.z, f[7] += b1.w;
        }
        uint4 pk;
        pk.x = pack_half2(fmaxf(f[0], 0.f), fmaxf(f[1], 0.f));
        pk.y = pack_half2(fmaxf(f[2], 0.f), fmaxf(f[3], 0.f));
        pk.z = pack_half2(fmaxf(f[4], 0.f), fmaxf(f[5], 0.f));
        pk.w = pack_half2(fmaxf(f[6], 0.f), fmaxf(f[7], 0.f));
        sts128(tile_chunk_addr(s_h, row, q * 2 + c), pk);
    }
}
__device__ __forceinline__ void row_relu_to_tile(uint32_t d_row, const float4 *__restrict__ bias, uint32_t s_h, uint32_t row) {
#if defined(LNB_FUSED_EPI_SERIAL)
#pragma unroll
    for (uint32_t q = 0; q < 4; ++q) {
        uint32_t v[16];
        tmem_ld16(d_row + 16 * q, v);
        tmem_ld_wait16(v);
        relu_chunk16(v, bias, q, s_h, row);
    }
#elif defined(LNB_FUSED_EPI_PIPE2)
    uint32_t va[16], vb[16];
    tmem_ld16(d_row, va);
    tmem_ld_wait16(va);
    tmem_ld16(d_row + 16, vb);
    relu_chunk16(va, bias, 0, s_h, row);
    tmem_ld_wait16(vb);
    tmem_ld16(d_row + 32, va);
    relu_chunk16(vb, bias, 1, s_h, row);
    tmem_ld_wait16(va);
    tmem_ld16(d_row + 48, vb);
    relu_chunk16(va, bias, 2, s_h, row);
    tmem_ld_wait16(vb);
    relu_chunk16(vb, bias, 3, s_h, row);
#else
    // three loads up front (one exposed tensor-memory latency for 48 columns), the fourth under the processing of the
    // second and third (timeline: a tcgen05.ld round trip is ~200 cycles, processing 16 columns ~70)
    uint32_t va[16], vb[16], vc[16];
    tmem_ld16(d_row, va);
    tmem_ld16(d_row + 16, vb);
    tmem_ld16(d_row + 32, vc);
    tmem_ld_wait16(va);
    tmem_ld_tie16(vb);
    tmem_ld_tie16(vc);
    relu_chunk16(va, bias, 0, s_h, row);
    tmem_ld16(d_row + 48, va);
    relu_chunk16(vb, bias, 1, s_h, row);
    relu_chunk16(vc, bias, 2, s_h, row);
    tmem_ld_wait16(va);
    relu_chunk16(va, bias, 3, s_h, row);
#endif
}

// (Measured and rejected: gathering with 4-byte cp.async into a per-warp shared-memory ring - no register per load in
// flight, 24+ loads per lane outstanding - made the gather 2.5x SLOWER than plain LDG (257 vs ~100 us at 385 k samples):
// LDGSTS of 4-byte elements is processed far below the LSU's gather rate.)
__global__ void __launch_bounds__(kFusedThreads, 1)
k_field_fused_fwd(const FusedArgs a, const __grid_constant__ CUtensorMap tm_fb_s, const __grid_constant__ CUtensorMap tm_fb_h) {
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
    const uint32_t s_in0 = s_h0 + kSlots * kTileBytes;             // kCoordBufs x [128][3] sample positions (as given)
    const uint32_t s_bar = s_in0 + kCoordBufs * kRows * 3 * 4;
    const uint32_t bar_xfull = s_bar, bar_xempty = bar_xfull + 8 * kStages, bar_ready = bar_xempty + 8 * kStages;
    const uint32_t bar_done = bar_ready + 8 * kSlots, bar_w = bar_done + 8 * kSlots, bar_c = bar_w + 8;
    const uint32_t s_slot = bar_c + 8 * kCoordBufs;

    // (read once through a volatile asm: under the 72-register cap the compiler otherwise re-reads SR_TID.X - a ~50-cycle
    // S2R - inside the per-tile loops of every role)
    uint32_t tid_x;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_x));
    const uint32_t warp = tid_x >> 5, lane = tid_x & 31u;
    constexpr uint32_t kTmemCols = 512;
    static_assert(kSlots * kTmemColsPerSlot <= kTmemCols, "tensor memory");

    if (warp == kMmaWarpIdx) tmem_alloc(s_slot, kTmemCols);
    if (tid_x == 0) {
        for (uint32_t b = 0; b < kCoordBufs; ++b) mbar_init(bar_c + 8 * b, 1);
        for (uint32_t s = 0; s < kStages; ++s) {
            mbar_init(bar_xfull + 8 * s, kGatherWarps);           // one arrival per gather warp
            mbar_init(bar_xempty + 8 * s, 1 + 4);                 // tcgen05.commit + one arrival per epilogue warp
        }
        for (uint32_t q = 0; q < kSlots; ++q) {
            mbar_init(bar_ready + 8 * q, 4);                      // one arrival per epilogue warp of the owning group
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
        constexpr uint32_t D = 3, C = 2;
        const uint32_t nlv = warp < a.L ? (a.L - warp + kGatherWarps - 1) / kGatherWarps : 0;   // levels of this warp
        // level-uniform quantities of the warp's first level, once per kernel
        LevelGeo g0 = {};
        LevelIndex<D> li0 = {};
        if (nlv > 0) {
            g0 = level_geo(a.offsets, warp, a.S, a.H);
            li0 = level_index<D>(g0, 0u, false);
        }
        // Sample positions travel global -> shared through the TMA engine (one 1.5 KB bulk copy per tile onto an mbarrier,
        // a ring of kCoordBufs blocks kept two tiles ahead by warp 0): no register staging and NO CTA-wide barrier per tile
        // - the gather warps drift apart by up to kStages tiles, bounded only by the operand ring.  (The first version
        // staged them with 384 threads behind a 512-thread named barrier per tile: barrier + exposed load latency were
        // 20 % of the gather warps' stall samples.)
        auto issue_coords = [&](uint32_t k) {
            const uint32_t bar = bar_c + 8 * (k % kCoordBufs);
            mbar_expect_tx(bar, kRows * 3 * 4);
            bulk_g2s(s_in0 + (k % kCoordBufs) * kRows * 3 * 4, a.xyz + ((size_t)blockIdx.x + (size_t)k * gridDim.x) * kRows * 3,
                     kRows * 3 * 4, bar);
        };
        // One level of one tile, the level-uniform decisions taken ONCE (template parameter), not per corner:
        //   kMode 0  dense level     row = x * 1 + y * side + z * side^2                     (can never wrap)
        //   kMode 1  hashed level    row = (x ^ y * p1 ^ z * p2) & (size - 1)                 (power-of-two table)
        //   kMode 2  anything else   the generic cell_row() with its `% hashmap_size`        (wrapping tiled grids ...)
        // Modes 0/1 also take floor() with the 2^23 trick (round-down add: the integer part lands in the mantissa; exact
        // for 0 <= pos < 2^22, which the mode selection guarantees) instead of FRND + F2I + I2F on the quarter-rate unit.
        // The first version kept `generic` / `hashed` as run-time flags inside the corner loop: 567 SASS instructions per
        // sample and level, and the 16 gather warps were ISSUE-bound (110 us with the table loads removed).
        auto gather_level = [&](auto mode_tag, uint32_t in, const LevelGeo &g, const LevelIndex<D> &li,
                                const __half *tab, uint32_t s_x, uint32_t level) {
            constexpr uint32_t kMode = decltype(mode_tag)::value;
            const uint32_t *__restrict__ tab32 = reinterpret_cast<const uint32_t *>(tab);      // C = 2 halves = one word per row
            const bool no_loads = (a.dbg & 1u) != 0;
#pragma unroll
            for (uint32_t pair = 0; pair < kRows / 64; ++pair) {
                // two sample groups at a time: 16 gathers in flight per lane (the 16 gather warps have to cover the
                // L2 latency that 48 warps cover in the stand-alone encoder kernel)
                uint32_t raw[2][8];
                float frac[2][D];
                bool live[2];
#pragma unroll
                for (uint32_t h = 0; h < 2; ++h) {
                    const uint32_t sl = pair * 64 + h * 32 + lane;
                    uint32_t base[D];
                    bool inside = true;
#pragma unroll
                    for (uint32_t d = 0; d < D; ++d) {
                        float v = lds_f32(in + (sl * D + d) * 4);
                        if (a.norm.x != 0.f) v = (v + a.norm.x) * a.norm.y;        // grid.py:213, as load_unit_coords()
                        if (v < 0 || v > 1) inside = false;
                        float pos = v * g.scale + 0.5f;                       // gridencoder.cu:160-162
                        if constexpr (kMode == 2) {
                            const float fl = floorf(pos);
                            base[d] = (uint32_t)fl;
                            pos -= (float)base[d];
                        } else {
                            const float t = __fadd_rd(pos, 8388608.0f);
                            base[d] = __float_as_uint(t) & 0x007fffffu;
                            pos -= t - 8388608.0f;
                        }
                        frac[h][d] = pos;
                    }
                    live[h] = inside && !no_loads;
                    if constexpr (kMode == 2) {
#pragma unroll
                        for (uint32_t corner = 0; corner < 8; ++corner) {
                            uint32_t pp[D];
#pragma unroll
                            for (uint32_t d = 0; d < D; ++d) pp[d] = base[d] + ((corner >> d) & 1u);
                            const uint32_t row = cell_row<D>(pp, 0u, false, g);
                            raw[h][corner] = live[h] ? __ldg(tab32 + row) : 0u;
                        }
                    } else {
                        uint32_t t0[2], t1[2], t2[2], xy[4];
                        t0[0] = base[0] * li.mul[0], t0[1] = t0[0] + li.mul[0];
                        t1[0] = base[1] * li.mul[1], t1[1] = t1[0] + li.mul[1];
                        t2[0] = base[2] * li.mul[2], t2[1] = t2[0] + li.mul[2];
#pragma unroll
                        for (uint32_t q = 0; q < 4; ++q) xy[q] = kMode == 1 ? (t0[q & 1u] ^ t1[q >> 1]) : (t0[q & 1u] + t1[q >> 1]);
#pragma unroll
                        for (uint32_t corner = 0; corner < 8; ++corner) {
                            const uint32_t row = kMode == 1 ? ((xy[corner & 3u] ^ t2[corner >> 2]) & li.mask)
                                                            : (xy[corner & 3u] + t2[corner >> 2]);
                            raw[h][corner] = live[h] ? __ldg(tab32 + row) : 0u;
                        }
                    }
                }
                // sum over the 8 corners, accumulated in fp16 in corner order exactly like interp_corners() / the
                // reference (gridencoder.cu:173-199); 0 outside the unit cube
#pragma unroll
                for (uint32_t h = 0; h < 2; ++h) {
                    __half r0 = __float2half_rn(0.f), r1 = r0;
                    if (live[h]) {
#pragma unroll
                        for (uint32_t corner = 0; corner < 8; ++corner) {
                            float w = 1;
#pragma unroll
                            for (uint32_t d = 0; d < D; ++d) {
                                if ((corner & (1u << d)) == 0) w *= 1 - frac[h][d];
                                else w *= frac[h][d];
                            }
                            const __half2 hv = *reinterpret_cast<const __half2 *>(&raw[h][corner]);
                            r0 = __float2half_rn(__half2float(r0) + w * __low2float(hv));
                            r1 = __float2half_rn(__half2float(r1) + w * __high2float(hv));
                        }
                    }
                    // the table and its interpolation are fp16 (reference semantics); the MLP operand is this unit's element type
                    const uint32_t packed = (uint32_t)mlp_from_float(__half2float(r0)) | ((uint32_t)mlp_from_float(__half2float(r1)) << 16);
                    sts32(elem_addr(s_x, pair * 64 + h * 32 + lane, level * C), packed);
                }
            }
        };
        auto level_mode = [&](const LevelGeo &g, const LevelIndex<D> &li) -> uint32_t {
            return (li.generic || !(g.scale < 4194304.0f) || (a.dbg & 16u)) ? 2u : (li.hashed ? 1u : 0u);
        };
        const uint32_t mode0 = nlv > 0 ? level_mode(g0, li0) : 0u;

        FTR_DECL(blockIdx.x == 0 && tid_x == 0);
        if (warp == 0 && lane == 0) {
            if (n_my > 0) issue_coords(0);
            if (n_my > 1) issue_coords(1);
        }
        for (uint32_t k = 0; k < n_my; ++k) {
            const uint32_t stage = k % kStages, use = k / kStages;
            FTR(3, 1, k);
            if (warp == 0 && k + 2 < n_my) {
                // block (k + 2) % 4 held tile k - 2: every gather warp has arrived on that tile's operand barrier, i.e. is
                // done reading its positions.  (Parity: this warp has not arrived for tile k + 1 yet, so the barrier of
                // that stage cannot have moved past the completion asked for.)
                if (k >= 2) mbar_wait(bar_xfull + 8 * ((k - 2) % kStages), ((k - 2) / kStages) & 1u);
                if (lane == 0) issue_coords(k + 2);
                __syncwarp();
            }
            mbar_wait(bar_c + 8 * (k % kCoordBufs), (k / kCoordBufs) & 1u);      // positions of tile k have landed
            FTR(3, 2, k);
            if (use > 0) mbar_wait(bar_xempty + 8 * stage, (use - 1) & 1u);
            FTR(3, 3, k);
            const uint32_t s_x = s_x0 + stage * kTileBytes;
            const uint32_t in = s_in0 + (k % kCoordBufs) * kRows * 3 * 4;
            for (uint32_t lv = 0; lv < nlv; ++lv) {
                const uint32_t level = warp + lv * kGatherWarps;
                LevelGeo g = g0;
                LevelIndex<D> li = li0;
                if (lv != 0) {
                    g = level_geo(a.offsets, level, a.S, a.H);
                    li = level_index<D>(g, 0u, false);
                }
                const __half *__restrict__ tab = a.table + (size_t)g.table_offset * C;
                const uint32_t mode = lv == 0 ? mode0 : level_mode(g, li);
                if (a.dbg & 8u) continue;
                if (mode == 1) gather_level(std::integral_constant<uint32_t, 1>{}, in, g, li, tab, s_x, level);
                else if (mode == 0) gather_level(std::integral_constant<uint32_t, 0>{}, in, g, li, tab, s_x, level);
                else gather_level(std::integral_constant<uint32_t, 2>{}, in, g, li, tab, s_x, level);
            }
            FTR(3, 4, k);
            fence_proxy_async();                            // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_xfull + 8 * stage);
            FTR(3, 5, k);
        }
    } else if (warp < kMmaWarpIdx) {
        // ======================= EPILOGUE groups: thread = tile row, two tiles in flight per group =======================
        const uint32_t g = (warp - kEpiWarp0) >> 2;
        const uint32_t wq = warp & 3u;                     // TMEM lane quadrant of this warp = rows 32 wq .. 32 wq + 31
        const uint32_t row = wq * 32u + lane;
        const uint32_t lane_sel = (wq * 32u) << 16;
        FTR_DECL(blockIdx.x == 0 && wq == 0 && lane == 0);
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
            if (cpr == 4) {                                  // 32-wide rows (L16 x F2 features): 4 copies, no division
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) {
                    const uint32_t q = lane + 32 * j, rr = wq * 32u + (q >> 2), c = q & 3u;
                    *reinterpret_cast<uint4 *>(dst_row0 + ((size_t)rr * 4 + c) * 8) = lds128(tile_chunk_addr(tile, rr, c));
                }
                return;
            }
            for (uint32_t q = lane; q < 32 * cpr; q += 32) {
                const uint32_t rr = wq * 32u + q / cpr, c = q % cpr;
                *reinterpret_cast<uint4 *>(dst_row0 + ((size_t)rr * cpr + c) * 8) = lds128(tile_chunk_addr(tile, rr, c));
            }
        };
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

        for (uint32_t q = 0; q < kSlotsPerGroup; ++q)       // tensor memory of the group's slots is free
            if (g * kSlotsPerGroup + q < n_my) arrive_warp(bar_ready + 8 * (g * kSlotsPerGroup + q));

        for (uint32_t k0 = g * kSlotsPerGroup; k0 < n_my; k0 += kSlots) {
            const uint32_t nt = min(kSlotsPerGroup, n_my - k0);
            // ---- per tile: gathered features -> enc (the backward pass needs them), release the operand tile ----
            for (uint32_t t = 0; t < nt; ++t) {
                const uint32_t k = k0 + t, stage = k % kStages, use = k / kStages;
                row0[t] = ((size_t)blockIdx.x + (size_t)k * gridDim.x) * kRows;
                rid[t] = (uint32_t)__ldg(a.ray_ids + row0[t] + row);
                FTR(g, 1, k);
                mbar_wait(bar_xfull + 8 * stage, use & 1u);
                FTR(g, 2, k);
                if (!(a.dbg & 2u)) copy_rows_out(s_x0 + stage * kTileBytes, a.enc + row0[t] * fs.enc_dim, fs.enc_dim / 8);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_xempty + 8 * stage);
                // this row's 256 B of the per-ray head bias, needed four layers from now: pull the lines into L1.
                // (Measured and rejected: staging the row in shared memory with 16 cp.async per thread at this point -
                // the whole kernel went from 152 to 206 us; LDGSTS with a different row per lane is far slower than the
                // exposed ld.global latency it was meant to hide.)
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
                    FTR(g, 3, ph);
                    mbar_wait(done, par[t]);
                    par[t] ^= 1;
                    fence_after_sync();
                    FTR(g, 4, ph);
                    const bool sigma_layer = ph <= fs.n_hid_s;
                    const bool head_layer = ph >= fs.n_hid_s + 2 && ph <= fs.n_hid_s + 2 + fs.n_hid_h;
                    if (sigma_layer || head_layer) {
                        // hidden layer: accumulator row -> (+per-ray bias) -> ReLU -> fp16 -> operand tile of the next layer
                        const uint32_t layer = sigma_layer ? ph : ph - (fs.n_hid_s + 2);
                        const float4 *bias = (head_layer && layer == 0)
                                                 ? reinterpret_cast<const float4 *>(a.ray_bias + (size_t)rid[t] * kHid) : nullptr;
                        if (lane == 0) tma_store_wait_read();   // the previous layer's copy-out has read this warp's rows
                        __syncwarp();
                        if (!(a.dbg & 4u)) row_relu_to_tile(d_hid, bias, s_h, row);
                        FTR(g, 5, ph);
                        publish(ready);
                        FTR(g, 6, ph);
                        if (!(a.dbg & 6u) && lane == 0)      // saved activations: this warp's 32 rows leave through the TMA engine
                            tma_store_rows(sigma_layer ? &tm_fb_s : &tm_fb_h, s_h + wq * 32u * 128u,
                                           layer * a.B + (uint32_t)row0[t] + wq * 32u);
                        FTR(g, 7, ph);
                    } else if (ph == fs.n_hid_s + 1) {
                        // density output: sig_out (fp16, kept for backward), sigma = exp(h0) * scale, geo -> head operand
                        uint32_t v[16];
                        tmem_ld16(d_out, v);
                        tmem_ld_wait();
                        unsigned short hv[16];
#pragma unroll
                        for (uint32_t j = 0; j < 16; ++j) hv[j] = mlp_from_float(__uint_as_float(v[j]));
                        uint32_t pw[8];
#pragma unroll
                        for (uint32_t j = 0; j < 8; ++j) pw[j] = (uint32_t)hv[2 * j] | ((uint32_t)hv[2 * j + 1] << 16);
                        uint4 *dst = reinterpret_cast<uint4 *>(a.sig_out + r * kOut);
                        dst[0] = make_uint4(pw[0], pw[1], pw[2], pw[3]);
                        dst[1] = make_uint4(pw[4], pw[5], pw[6], pw[7]);
                        a.sigma[r] = __expf(mlp_to_float(hv[0])) * a.density_scale;    // activation.py:6-20 (forward)
                        if (lane == 0) tma_store_wait_read();   // copy-out of the last hidden layer is done with these rows
                        __syncwarp();
                        for (uint32_t c = 0; c < 2 * fs.ks_geo; ++c) sts128(tile_chunk_addr(s_h, row, c), make_uint4(0, 0, 0, 0));
                        // the 15 geo features hv[1..15] -> columns geo_off .. geo_off + 14, as 8 word stores (the column in
                        // front of / behind them is a zero of the fill above)
                        if (fs.geo_off & 1u) {
#pragma unroll
                            for (uint32_t j = 0; j < 8; ++j)
                                sts32(elem_addr(s_h, row, fs.geo_off - 1 + 2 * j),
                                      (j == 0 ? 0u : (uint32_t)hv[2 * j]) | ((uint32_t)hv[2 * j + 1] << 16));
                        } else {
#pragma unroll
                            for (uint32_t j = 0; j < 8; ++j)
                                sts32(elem_addr(s_h, row, fs.geo_off + 2 * j),
                                      (uint32_t)hv[2 * j + 1] | (j == 7 ? 0u : ((uint32_t)hv[2 * j + 2] << 16)));
                        }
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
        if (lane == 0) tma_store_wait_all();
    } else {
        // ======================= MMA warps: one per tile slot (converged; one elected lane issues) =======================
        // Each warp walks its slot's tiles and the network's steps in program order with blocking mbarrier waits.  (The first
        // version had ONE warp poll all slots and pick the step to issue from per-slot state arrays: descriptors then live in
        // ordinary registers and every tcgen05.mma costs an ELECT / VOTEU / 4x R2UR chain - the timeline showed ~480 cycles
        // from "operand ready" to "step issued" and the single issuer as the bottleneck of the whole kernel, 62 us with all
        // other work removed.)  Everything an operand descriptor depends on is made warp-uniform IN THE COMPILER'S EYES
        // (__shfl_sync from lane 0), so descriptors are built in uniform registers.
        // Order on the operand ring: slot q takes tiles q, q + kSlots, ...; tile k's stage was last used by tile k - kStages,
        // whose gather completed before that of tile k - kSlots (kSlots <= kStages), which this warp has already consumed -
        // so the parity wait below can only be satisfied by the fill of tile k itself.
        static_assert(kSlots <= kStages, "operand ring order");
        const uint32_t q = __shfl_sync(0xffffffffu, warp - kMmaWarpIdx, 0);
        const uint32_t n_my_u = __shfl_sync(0xffffffffu, n_my, 0);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        FTR_DECL(blockIdx.x == 0 && lane == 0);
        mbar_wait_warp(bar_w, 0);                           // weight image landed (TMA transaction bytes complete)
        const uint32_t d_hid = tmem_u + q * kTmemColsPerSlot, d_out = d_hid + 64;
        const uint32_t s_h = s_h0 + q * kTileBytes;
        const uint32_t ready = bar_ready + 8 * q, done = bar_done + 8 * q;
        uint32_t par = 0, st = 0;
        auto step = [&](uint32_t d, uint32_t a_tile, uint32_t w_tile, uint32_t ksteps, uint32_t idesc, uint32_t xfull, uint32_t xpar) {
            mbar_wait_warp(ready, par);
            par ^= 1;
            if (xfull) mbar_wait_warp(xfull, xpar);
            fence_after_sync();
            FTR(2, 1 + 2 * q, st);
            issue_kmajor(d, a_tile, w_tile, ksteps, idesc, false);
            mma_commit_elect(done);
            FTR(2, 2 + 2 * q, st);
            ++st;
        };
        for (uint32_t k = q; k < n_my_u; k += kSlots) {
            const uint32_t stage = k % kStages, use = k / kStages;
            st = 0;
            step(d_hid, s_x0 + stage * kTileBytes, s_ws_in, ks_in, kIdescFwdHid, bar_xfull + 8 * stage, use & 1u);
            mma_commit_elect(bar_xempty + 8 * stage);       // the operand tile has been consumed
            for (uint32_t l = 0; l < fs.n_hid_s; ++l) step(d_hid, s_h, s_ws_hid + l * kWTileBytes, 4, kIdescFwdHid, 0, 0);
            step(d_out, s_h, s_ws_out, 4, kIdescFwdOut, 0, 0);
            step(d_hid, s_h, s_wh_geo, fs.ks_geo, kIdescFwdHid, 0, 0);
            for (uint32_t l = 0; l < fs.n_hid_h; ++l) step(d_hid, s_h, s_wh_hid + l * kWTileBytes, 4, kIdescFwdHid, 0, 0);
            step(d_out, s_h, s_wh_out, 4, kIdescFwdOut, 0, 0);
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
           8 * (2 * kStages + 2 * kSlots + 1 + kCoordBufs) + 16;
}

// [layers * B rows][64 halves] row-major activations, written in boxes of 32 rows from 128-byte-swizzled tiles
int make_rows_map(CUtensorMap *tm, void *base, uint64_t rows) {
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const EncodeTiled encode = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiled>(p);
    }();
    if (!encode) return LNB_ERR_UNSUPPORTED;
    const cuuint64_t dims[2] = {kHid, rows};
    const cuuint64_t strides[1] = {kHid * 2};
    const cuuint32_t box[2] = {kHid, 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? LNB_OK : LNB_ERR_INVALID_ARGUMENT;
}

// L2 residency of the hash table (lnb_field_set_l2_window, one setting for the fp16 and the bf16 build of this unit): the gather's launches carry an access-policy window that
// marks the table's lines as persisting and everything else the kernel touches as streaming.  Between two forward passes
// Adam streams 411 MB and the backward kernels ~500 MB through the 126 MB L2; without the window the 27 MB table is
// re-read from DRAM every step (ncu: 36 MB of DRAM reads per launch).
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
#define lnb_debug_fwd_trace_fused lnb_debug_fwd_trace_fused_bf16
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

#ifdef LNB_TRACE
// diagnostic builds only (build.py --trace): [4 roles][4096] (event << 52 | arg << 44 | SM clock) words of CTA 0, zero = unused
int lnb_debug_fwd_trace_fused(unsigned long long *host_out, int reset) {
    cudaDeviceSynchronize();
    if (host_out) cudaMemcpyFromSymbol(host_out, g_fwd_trace, sizeof(unsigned long long) * 4 * 4096);
    if (reset) {
        void *p = nullptr;
        cudaGetSymbolAddress(&p, g_fwd_trace);
        cudaMemset(p, 0, sizeof(unsigned long long) * 4 * 4096);
    }
    return (int)cudaGetLastError();
}
#endif

#ifndef LNB_BF16
// Declare [table, table + bytes) as the hash table later lnb_field_fused_forward*() launches read: reserves persisting L2
// for it (cudaLimitPersistingL2CacheSize) and makes those launches carry the access-policy window.  bytes = 0 switches
// the window off.  Process-wide, not thread-safe; returns LNB_ERR_UNSUPPORTED when the device has no persisting L2.
int lnb_field_set_l2_window(const void *table, size_t bytes) {
    FieldL2Window *w = &g_field_l2_window;
    if (!table || bytes == 0) {
        w->base = nullptr, w->bytes = 0;
        return LNB_OK;
    }
    int dev = 0, max_persist = 0, max_window = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    if (max_persist <= 0 || max_window <= 0) return LNB_ERR_UNSUPPORTED;
    const size_t want = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
    cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    w->base = table;
    w->bytes = bytes < (size_t)max_window ? bytes : (size_t)max_window;
    return LNB_OK;
}
#endif

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
    size_t smem = fused_smem_bytes(fs);
    {   // diagnostics only: LNB_FUSED_EXTRA_SMEM_KB pads the request (how much does the gather depend on the L1 share of the
        // 256 KB array?  profiles/r02_fused_forward_diag.txt)
        static const size_t extra = [] { const char *e = getenv("LNB_FUSED_EXTRA_SMEM_KB"); return e ? (size_t)atoi(e) * 1024 : (size_t)0; }();
        smem += extra;
    }
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
    CUtensorMap tm_s, tm_h;
    if ((rc = make_rows_map(&tm_s, fb_sigma, (uint64_t)(fs.n_hid_s + 1) * M)) != LNB_OK) return rc;
    if ((rc = make_rows_map(&tm_h, fb_head, (uint64_t)(fs.n_hid_h + 1) * M)) != LNB_OK) return rc;
    const uint32_t tiles = M / kRows;
    const uint32_t cap = (uint32_t)sm_count_fused();
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(tiles < cap ? tiles : cap);
    lc.blockDim = dim3(kFusedThreads);
    lc.dynamicSmemBytes = smem;
    lc.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    if (g_field_l2_window.bytes != 0 && table == g_field_l2_window.base) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<void *>(g_field_l2_window.base);
        attr[0].val.accessPolicyWindow.num_bytes = g_field_l2_window.bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        lc.attrs = attr;
        lc.numAttrs = 1;
    }
    e = cudaLaunchKernelEx(&lc, k_field_fused_fwd, a, tm_s, tm_h);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    count_launch();
    return launch_status();
}

}  // extern "C"
