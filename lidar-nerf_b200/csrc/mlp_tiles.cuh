// Tile movers, UMMA issue helpers and warpgroup utilities shared by the fused-MLP kernels (ffmlp.cu, field.cu).
// Everything here is file-local to the including translation unit (anonymous namespace).
#pragma once
#include <cuda_bf16.h>

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
constexpr uint32_t kIdescWgrad128 = instr_desc_f16(64, 128, 1, 1);

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

// ---- the MLP element type of this translation unit -------------------------------------------------------------
// Weights, activations and activation gradients are 16-bit: fp16 by default, bf16 when the unit is compiled with
// -DLNB_BF16 (build.py compiles field.cu / field_fused.cu / ffmlp.cu a second time that way and the entry points get the
// suffix `_bf16`: BASELINE config 5, "bf16 MLP on tensor cores").  Pointers stay `__half *` = "16-bit element"; only the
// conversions below and the tcgen05 operand format (tcgen05.cuh) know the difference.  fp32 accumulation either way.
#ifdef LNB_BF16
#define LNB_SYM(name) name##_bf16
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float mlp_to_float(unsigned short bits) { return __uint_as_float((uint32_t)bits << 16); }
__device__ __forceinline__ unsigned short mlp_from_float(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
#else
#define LNB_SYM(name) name
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float mlp_to_float(unsigned short bits) { return __half2float(__ushort_as_half(bits)); }
__device__ __forceinline__ unsigned short mlp_from_float(float v) { return __half_as_ushort(__float2half_rn(v)); }
#endif
__device__ __forceinline__ float mlp_to_float(__half v) { return mlp_to_float(__half_as_ushort(v)); }
// always fp16 (gradients handed to the hash-grid scatter, which reads fp16)
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// ---- MMA issue helpers: warp-collective (call from a converged warp with warp-uniform arguments; one elected lane
// issues, see mma_f16_elect).  Descriptors are built once per call; a K step only bumps the 14-bit address field. ----
__device__ __forceinline__ uint64_t desc_step(uint64_t desc, uint32_t k, uint32_t step_bytes) {
    return desc + (uint64_t)(k * (step_bytes >> 4));
}

// D[128 x N] (+)= A[128 x K] * B[N x K]^T ; both K-major tiles; K = 16 * ksteps (<= 64)
__device__ __forceinline__ void issue_kmajor(uint32_t d, uint32_t a_tile, uint32_t b_tile, uint32_t ksteps,
                                             uint32_t idesc, bool accumulate_first) {
    const uint64_t a0 = smem_desc_sw128(a_tile, 16), b0 = smem_desc_sw128(b_tile, 16);
#pragma unroll 4
    for (uint32_t k = 0; k < ksteps; ++k)
        mma_f16_elect(d, desc_step(a0, k, 32), desc_step(b0, k, 32), idesc, (accumulate_first || k > 0) ? 1u : 0u);
}
// D[128 x 64] (+)= A[128 x K] (K-major tile) * W (tile holding W[K rows][64 cols], read MN-major)
__device__ __forceinline__ void issue_dgrad(uint32_t d, uint32_t a_tile, uint32_t w_tile, uint32_t ksteps) {
    const uint64_t a0 = smem_desc_sw128(a_tile, 16), b0 = smem_desc_sw128(w_tile, kWTileBytes);
#pragma unroll 4
    for (uint32_t k = 0; k < ksteps; ++k)
        mma_f16_elect(d, desc_step(a0, k, 32), desc_step(b0, k, 2048), kIdescDgrad, k > 0 ? 1u : 0u);
}
// D[64 x N] (+)= A^T B over the 128 rows of the activation tiles (both read MN-major); N (multiple of 16, from the
// instruction descriptor) may run past one 64-column B tile into the next consecutive one, addressed through the
// descriptor's leading-dimension offset
__device__ __forceinline__ void issue_wgrad(uint32_t d, uint32_t a_tile, uint32_t b_tile, bool accumulate_first,
                                            uint32_t n_cols = 64) {
    const uint64_t a0 = smem_desc_sw128(a_tile, kTileBytes), b0 = smem_desc_sw128(b_tile, kTileBytes);
    const uint32_t idesc = instr_desc_f16(64, n_cols, 1, 1);
#pragma unroll
    for (uint32_t k = 0; k < kRows / 16; ++k)
        mma_f16_elect(d, desc_step(a0, k, 2048), desc_step(b0, k, 2048), idesc, (accumulate_first || k > 0) ? 1u : 0u);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

}  // namespace
}  // namespace lnb
