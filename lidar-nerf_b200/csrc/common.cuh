// Shared helpers for the sm_100a kernels behind include/lidarnerf_b200.h.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lidarnerf_b200.h"

namespace lnb {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kWarp = 32;

// Every launch goes through this counter so bench.py can report `gpu_launches`.
extern unsigned long long g_launch_count;
inline void count_launch(unsigned n = 1) { g_launch_count += n; }

inline int launch_status() {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear the sticky-free launch error so the next call starts clean
        return (int)e;
    }
    return LNB_OK;
}

inline cudaStream_t as_stream(lnb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ __forceinline__ T ceil_div(T a, T b) {
    return (a + b - 1) / b;
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// Width of the head's direction encoding for a `degree` argument of the lnb_field_* entry points (LNB_DIR_SH, see the
// header): frequency 3 + 6 deg, spherical harmonics deg^2.
__host__ __device__ __forceinline__ uint32_t dir_code_width(uint32_t code) {
    return (code & 0x100u) ? (code & 0xffu) * (code & 0xffu) : 3u + 6u * code;
}
__host__ __device__ __forceinline__ bool dir_code_valid(uint32_t code) {
    return (code & 0x100u) ? ((code & 0xffu) == 4u) : (code >= 1u && code <= 20u);     // SH: degree 4 only (16 columns)
}

// Number of rows a sample-parallel kernel has to process: all B of them, or - when the caller passes the device
// counter the march kernel filled - the produced samples rounded up to a 128-row tile.  Lets the training step
// size its buffers generously (no dropped rays) while the work tracks the real sample count without a host sync.
__device__ __forceinline__ uint32_t active_rows(uint32_t B, const int32_t *__restrict__ n_active) {
    if (!n_active) return B;
    const int32_t n = *n_active;
    const uint32_t up = ((uint32_t)(n > 0 ? n : 0) + 127u) & ~127u;
    return up < B ? up : B;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

// Inclusive warp scans (Kogge-Stone over shuffles).
__device__ __forceinline__ float warp_scan_add(float v) {
    const unsigned l = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float u = __shfl_up_sync(kFullMask, v, o);
        if (l >= (unsigned)o) v += u;
    }
    return v;
}
__device__ __forceinline__ float warp_scan_mul(float v) {
    const unsigned l = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float u = __shfl_up_sync(kFullMask, v, o);
        if (l >= (unsigned)o) v *= u;
    }
    return v;
}

// 10-bit-per-axis Morton code helpers (same bit layout as the reference, raymarching.cu:71-95,
// i.e. x -> bits 0,3,6..., y -> bits 1,4,7..., z -> bits 2,5,8...).
__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {
    // shift-and-add (not shift-or): identical to the reference's multiply form for EVERY uint32
    // input, including out-of-range coordinates >= 2^10 where carries make the two differ.
    v = (v + (v << 16)) & 0xFF0000FFu;
    v = (v + (v << 8)) & 0x0F00F00Fu;
    v = (v + (v << 4)) & 0xC30C30C3u;
    v = (v + (v << 2)) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t compact3(uint32_t v) {
    v &= 0x49249249u;
    v = (v | (v >> 2)) & 0xC30C30C3u;
    v = (v | (v >> 4)) & 0x0F00F00Fu;
    v = (v | (v >> 8)) & 0xFF0000FFu;
    v = (v | (v >> 16)) & 0x0000FFFFu;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton_encode(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}

}  // namespace lnb
