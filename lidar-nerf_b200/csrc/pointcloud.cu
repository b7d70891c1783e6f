// Range image <-> point cloud conversion and Chamfer nearest-neighbour search for sm_100a: the evaluation-side
// callers of the hot path.
//
// Behavioural spec:
//   lidarnerf/convert.py:99-160   lidar_to_pano_with_intensities  (Python loop over points, z-buffer)
//   lidarnerf/convert.py:194-235  pano_to_lidar_with_intensities  (numpy, row-major order of the non-empty pixels)
//   extern/chamfer3D/chamfer3D.cu:9-166 NmDistanceKernel / :167-236 NmDistanceGradKernel
#include "common.cuh"

namespace lnb {
namespace {

constexpr float kPi = 3.14159265358979323846f;

// ------------------------------------------------------------------------------------------------------------
// Chamfer: nearest neighbour of every xyz1 point in xyz2 (squared distance, FIRST minimum in index order - the strict
// `<` of the reference).  The reference walks xyz2 in 512-point shared-memory batches with 512-thread CTAs on a fixed
// 32 x 16 grid; here a CTA owns 128 query points, streams xyz2 through a 2048-point SoA tile (broadcast reads, no bank
// conflicts) and keeps four independent distance chains in flight per thread.
// ------------------------------------------------------------------------------------------------------------
// x*x + y*y + z*z (chamfer3D.cu:36-39) exactly as nvcc contracts it in the reference build - FMUL y,y; FFMA x,x,+; FFMA
// z,z,+ (SASS of NmDistanceKernel) - spelled out so that the bits do not depend on this compiler's choice
__device__ __forceinline__ float sqdist(float x, float y, float z) { return fmaf(z, z, fmaf(x, x, __fmul_rn(y, y))); }

constexpr int kNNThreads = 128;
constexpr int kNNTile = 2048;

__global__ void __launch_bounds__(kNNThreads)
k_chamfer_nn(const float *__restrict__ xyz1, const float *__restrict__ xyz2, uint32_t N, uint32_t M,
             float *__restrict__ dist, int32_t *__restrict__ idx) {
    __shared__ float sx[kNNTile], sy[kNNTile], sz[kNNTile];
    const uint32_t b = blockIdx.y;
    const uint32_t j = blockIdx.x * kNNThreads + threadIdx.x;
    const float *q = xyz2 + (size_t)b * M * 3;
    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
    if (j < N) {
        const float *p = xyz1 + ((size_t)b * N + j) * 3;
        x1 = p[0], y1 = p[1], z1 = p[2];
    }
    float best = 0.f;
    int32_t best_i = 0;
    for (uint32_t k0 = 0; k0 < M; k0 += kNNTile) {
        const uint32_t n = min((uint32_t)kNNTile, M - k0);
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < n * 3; t += kNNThreads) {      // coalesced AoS read -> SoA tile
            const float v = __ldg(q + (size_t)k0 * 3 + t);
            const uint32_t k = t / 3, c = t - k * 3;
            (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v;
        }
        __syncthreads();
        uint32_t k = 0;
        if (k0 == 0 && n > 0) {   // the reference seeds `best` with candidate 0 unconditionally (k == 0 || d < best)
            const float x2 = sx[0] - x1, y2 = sy[0] - y1, z2 = sz[0] - z1;
            best = sqdist(x2, y2, z2);
            best_i = 0;
            k = 1;
        }
        for (; k + 4 <= n; k += 4) {
            float d[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float x2 = sx[k + u] - x1, y2 = sy[k + u] - y1, z2 = sz[k + u] - z1;
                d[u] = sqdist(x2, y2, z2);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (d[u] < best) best = d[u], best_i = (int32_t)(k0 + k + u);
        }
        for (; k < n; ++k) {
            const float x2 = sx[k] - x1, y2 = sy[k] - y1, z2 = sz[k] - z1;
            const float d = sqdist(x2, y2, z2);
            if (d < best) best = d, best_i = (int32_t)(k0 + k);
        }
    }
    if (j < N) {
        dist[(size_t)b * N + j] = best;
        idx[(size_t)b * N + j] = best_i;
    }
}

// chamfer3D.cu:167-197: gradient of sum(grad_dist1 * dist1) w.r.t. both clouds through the matched pairs
__global__ void __launch_bounds__(256)
k_chamfer_grad(const float *__restrict__ xyz1, const float *__restrict__ xyz2, const float *__restrict__ grad_dist1,
               const int32_t *__restrict__ idx1, uint32_t N, uint32_t M, float *__restrict__ grad_xyz1,
               float *__restrict__ grad_xyz2) {
    const uint32_t b = blockIdx.y;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const float *p = xyz1 + ((size_t)b * N + j) * 3;
    const int32_t j2 = idx1[(size_t)b * N + j];
    const float *q = xyz2 + ((size_t)b * M + j2) * 3;
    const float g = grad_dist1[(size_t)b * N + j] * 2;
    float *g1 = grad_xyz1 + ((size_t)b * N + j) * 3;
    float *g2 = grad_xyz2 + ((size_t)b * M + j2) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = g * (p[c] - q[c]);
        atomicAdd(g1 + c, v);
        atomicAdd(g2 + c, -v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// points -> range image.  The reference loop keeps, per pixel, the closest point and the first one among equals
// (`pano == 0` or `pano > dist`): that is the minimum of the 64-bit key (dist bits << 32 | point index) - positive
// floats order like their bit patterns.  Pass 1 takes the atomic minimum per pixel, pass 2 decodes the winner.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pano_keys(const float *__restrict__ points, uint32_t stride, uint32_t N, uint32_t H, uint32_t W, float fov_up,
            float fov, float max_depth, unsigned long long *__restrict__ keys) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float *p = points + (size_t)n * stride;
    const float x = p[0], y = p[1], z = p[2];
    const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));   // np.linalg.norm, no FMA
    if (!(dist < max_depth)) return;
    const float down = (float)((double)(fov - fov_up) / 180.0 * 3.14159265358979323846);
    const float col_step = (float)(2.0 * 3.14159265358979323846 / (double)W);
    const float row_step = (float)((double)fov / 180.0 * 3.14159265358979323846 / (double)H);
    const float beta = kPi - atan2f(y, x);
    const float alpha = atan2f(z, sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)))) + down;
    const int c = __float2int_rn(__fdiv_rn(beta, col_step));                 // round half to even, like Python's round()
    const int r = __float2int_rn((float)H - __fdiv_rn(alpha, row_step));
    if (r >= (int)H || r < 0 || c >= (int)W || c < 0) return;
    if (dist == 0.0f) return;   // a zero range reads as "empty" in the reference as well
    const unsigned long long key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)n;
    atomicMin(keys + (size_t)r * W + c, key);
}

__global__ void __launch_bounds__(256)
k_pano_decode(const unsigned long long *__restrict__ keys, const float *__restrict__ points, uint32_t stride,
              uint32_t HW, float *__restrict__ pano, float *__restrict__ intensities) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const unsigned long long k = keys[i];
    if (k == ~0ull) {
        pano[i] = 0.f;
        if (intensities) intensities[i] = 0.f;
        return;
    }
    pano[i] = __uint_as_float((uint32_t)(k >> 32));
    if (intensities) intensities[i] = stride > 3 ? points[(size_t)(uint32_t)k * stride + 3] : 0.f;
}

// ------------------------------------------------------------------------------------------------------------
// range image -> points, in row-major order of the non-empty pixels (np.where order): per-1024-pixel block counts,
// then every block sums the counts before it and scans its own pixels with ballots.
// ------------------------------------------------------------------------------------------------------------
constexpr int kScanBlock = 1024;

__global__ void __launch_bounds__(kScanBlock)
k_pano_count(const float *__restrict__ pano, uint32_t HW, int32_t *__restrict__ block_counts) {
    const uint32_t i = blockIdx.x * kScanBlock + threadIdx.x;
    const int live = (i < HW) && (pano[i] != 0.0f);
    const int n = __syncthreads_count(live);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = n;
}

__global__ void __launch_bounds__(kScanBlock)
k_pano_points(const float *__restrict__ pano, const float *__restrict__ intensities, uint32_t H, uint32_t W,
              float fov_up, float fov, const int32_t *__restrict__ block_counts, float *__restrict__ out,
              int32_t *__restrict__ count_out) {
    __shared__ int s_warp[kScanBlock / 32];
    __shared__ int s_base;
    const uint32_t HW = H * W;
    const uint32_t i = blockIdx.x * kScanBlock + threadIdx.x;
    const bool live = (i < HW) && (pano[i] != 0.0f);
    // offset of this block = counts of all blocks before it (at most a few hundred values)
    int part = 0;
    for (uint32_t b = threadIdx.x; b < blockIdx.x; b += kScanBlock) part += block_counts[b];
    part = warp_sum(part);
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    if (lane == 0 && part) atomicAdd(&s_base, part);
    const unsigned ball = __ballot_sync(kFullMask, live);
    if (lane == 0) s_warp[warp] = __popc(ball);
    __syncthreads();
    int before = s_base;
    for (unsigned w = 0; w < warp; ++w) before += s_warp[w];
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanBlock - 1) {
        int total = s_base;
        for (unsigned w = 0; w < kScanBlock / 32; ++w) total += s_warp[w];
        count_out[0] = total;
    }
    if (!live) return;
    const uint32_t slot = (uint32_t)before + __popc(ball & ((1u << lane) - 1u));
    const uint32_t jrow = i / W, icol = i - jrow * W;
    // convert.py:209-218 on float32 arrays: same operation order, no contraction
    const float beta = __fmul_rn(__fmul_rn(__fdiv_rn(-((float)icol - (float)W / 2.0f), (float)W), 2.0f), kPi);
    const float alpha = __fmul_rn(__fdiv_rn(__fsub_rn(fov_up, __fmul_rn(__fdiv_rn((float)jrow, (float)H), fov)), 180.0f), kPi);
    const float d = pano[i];
    float ca, sa, cb, sb;
    sincosf(alpha, &sa, &ca);
    sincosf(beta, &sb, &cb);
    float4 o;
    o.x = __fmul_rn(__fmul_rn(ca, cb), d);
    o.y = __fmul_rn(__fmul_rn(ca, sb), d);
    o.z = __fmul_rn(sa, d);
    o.w = intensities ? intensities[i] : 0.f;
    reinterpret_cast<float4 *>(out)[slot] = o;
}

}  // namespace
}  // namespace lnb

using namespace lnb;

extern "C" {

int lnb_chamfer_forward(const float *xyz1, const float *xyz2, uint32_t B, uint32_t N, uint32_t M, float *dist1,
                        float *dist2, int32_t *idx1, int32_t *idx2, lnb_stream_t stream) {
    if (!xyz1 || !xyz2 || !dist1 || !dist2 || !idx1 || !idx2) return LNB_ERR_INVALID_ARGUMENT;
    if (B == 0 || N == 0 || M == 0) return LNB_ERR_INVALID_ARGUMENT;
    if (B > 65535) return LNB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    k_chamfer_nn<<<dim3(ceil_div<uint32_t>(N, kNNThreads), B), kNNThreads, 0, st>>>(xyz1, xyz2, N, M, dist1, idx1);
    k_chamfer_nn<<<dim3(ceil_div<uint32_t>(M, kNNThreads), B), kNNThreads, 0, st>>>(xyz2, xyz1, M, N, dist2, idx2);
    count_launch(2);
    return launch_status();
}

int lnb_chamfer_backward(const float *xyz1, const float *xyz2, float *grad_xyz1, float *grad_xyz2,
                         const float *grad_dist1, const float *grad_dist2, const int32_t *idx1, const int32_t *idx2,
                         uint32_t B, uint32_t N, uint32_t M, lnb_stream_t stream) {
    if (!xyz1 || !xyz2 || !grad_xyz1 || !grad_xyz2 || !grad_dist1 || !grad_dist2 || !idx1 || !idx2)
        return LNB_ERR_INVALID_ARGUMENT;
    if (B == 0 || N == 0 || M == 0) return LNB_ERR_INVALID_ARGUMENT;
    if (B > 65535) return LNB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    k_chamfer_grad<<<dim3(ceil_div<uint32_t>(N, 256), B), 256, 0, st>>>(xyz1, xyz2, grad_dist1, idx1, N, M, grad_xyz1,
                                                                        grad_xyz2);
    k_chamfer_grad<<<dim3(ceil_div<uint32_t>(M, 256), B), 256, 0, st>>>(xyz2, xyz1, grad_dist2, idx2, M, N, grad_xyz2,
                                                                        grad_xyz1);
    count_launch(2);
    return launch_status();
}

size_t lnb_lidar_to_pano_workspace_bytes(uint32_t H, uint32_t W) { return (size_t)H * W * sizeof(unsigned long long); }

int lnb_lidar_to_pano(const float *points, uint32_t point_stride, uint32_t N, uint32_t H, uint32_t W, float fov_up,
                      float fov, float max_depth, float *pano, float *intensities, void *workspace,
                      lnb_stream_t stream) {
    if (!pano || !workspace || (N && !points) || H == 0 || W == 0) return LNB_ERR_INVALID_ARGUMENT;
    if (point_stride < 3) return LNB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    const uint32_t HW = H * W;
    cudaError_t e = cudaMemsetAsync(workspace, 0xFF, (size_t)HW * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return (int)e;
    if (N) {
        k_pano_keys<<<ceil_div<uint32_t>(N, 256), 256, 0, st>>>(points, point_stride, N, H, W, fov_up, fov, max_depth,
                                                                static_cast<unsigned long long *>(workspace));
        count_launch();
    }
    k_pano_decode<<<ceil_div<uint32_t>(HW, 256), 256, 0, st>>>(static_cast<const unsigned long long *>(workspace), points,
                                                               point_stride, HW, pano, intensities);
    count_launch();
    return launch_status();
}

size_t lnb_pano_to_lidar_workspace_bytes(uint32_t H, uint32_t W) {
    return sizeof(int32_t) * ((size_t)ceil_div<uint32_t>(H * W, kScanBlock) + 1);
}

int lnb_pano_to_lidar(const float *pano, const float *intensities, uint32_t H, uint32_t W, float fov_up, float fov,
                      float *points_out, int32_t *count_out, void *workspace, lnb_stream_t stream) {
    if (!pano || !points_out || !count_out || !workspace || H == 0 || W == 0) return LNB_ERR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(points_out) & 15u) return LNB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = as_stream(stream);
    const uint32_t HW = H * W, nb = ceil_div<uint32_t>(HW, kScanBlock);
    int32_t *counts = static_cast<int32_t *>(workspace);
    k_pano_count<<<nb, kScanBlock, 0, st>>>(pano, HW, counts);
    k_pano_points<<<nb, kScanBlock, 0, st>>>(pano, intensities, H, W, fov_up, fov, counts, points_out, count_out);
    count_launch(2);
    return launch_status();
}

}  // extern "C"
