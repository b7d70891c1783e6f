// Microbenchmark: what is the ceiling for the hash-grid gather / scatter on this GPU?
//   (a) random 4-byte (half2 row) gathers from an L2-resident table of the KITTI configuration's size (27 MB),
//       8 independent loads in flight per thread, as a function of how clustered the addresses of a warp are;
//   (b) random 8-byte RED.v2.f32 (and 16-byte RED.v4.f32) into an L2-resident 55 MB fp32 table.
// The numbers are the roofline denominators for k_grid_fwd / k_grid_bwd (they are L1/L2 bound, not HBM bound).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/gather_peak.bin scripts/micro/gather_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// each thread: `iters` rounds of 8 independent 4-byte loads; `cluster` = log2 of the number of consecutive lanes that
// share one random base (0: every lane random; 5: the whole warp reads 32 consecutive rows)
__global__ void k_gather(const uint32_t *__restrict__ table, uint32_t mask, uint32_t iters, uint32_t cluster,
                         uint32_t *__restrict__ sink) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        uint32_t v[8];
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            const uint32_t group = (t >> cluster) * 8u + k;
            const uint32_t row = (mix(group * 0x9e3779b9u + it) + (lane & ((1u << cluster) - 1u))) & mask;
            v[k] = __ldg(table + row);
        }
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) acc += v[k];
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int kVec>
__global__ void k_scatter(float *__restrict__ table, uint32_t mask, uint32_t iters) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            const uint32_t row = mix((t * 8u + k) * 0x9e3779b9u + it) & mask;
            if (kVec == 2) atomicAdd(reinterpret_cast<float2 *>(table) + row, make_float2(1.f, 1.f));
            else atomicAdd(reinterpret_cast<float4 *>(table) + (row >> 1), make_float4(1.f, 1.f, 1.f, 1.f));
        }
    }
}

int main() {
    const uint32_t rows = 1u << 23;                       // 8 M rows: 32 MB of half2 rows / 64 MB of float2 rows
    uint32_t *table, *sink;
    float *gtable;
    cudaMalloc(&table, (size_t)rows * 4);
    cudaMalloc(&gtable, (size_t)rows * 8);
    cudaMalloc(&sink, 4);
    cudaMemset(table, 1, (size_t)rows * 4);
    cudaMemset(gtable, 0, (size_t)rows * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const uint32_t blocks = 148 * 16, threads = 256, iters = 64;
    const double n_ops = (double)blocks * threads * iters * 8;
    for (uint32_t cluster = 0; cluster <= 5; ++cluster) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            k_gather<<<blocks, threads>>>(table, rows - 1, iters, cluster, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("gather  4 B rows, %2u consecutive lanes share a base: %7.1f G lane-loads/s  (%6.1f us for %.0f M)\n",
               1u << cluster, n_ops / best / 1e6, best * 1e3, n_ops / 1e6);
    }
    for (int vec = 2; vec <= 4; vec += 2) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaMemset(gtable, 0, (size_t)rows * 8);
            cudaEventRecord(e0);
            if (vec == 2) k_scatter<2><<<blocks, threads>>>(gtable, rows - 1, iters);
            else k_scatter<4><<<blocks, threads>>>(gtable, rows - 1, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("scatter RED.v%d.f32 random rows (64 MB table):            %7.1f G reductions/s (%6.1f us for %.0f M)\n", vec,
               n_ops / best / 1e6, best * 1e3, n_ops / 1e6);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
