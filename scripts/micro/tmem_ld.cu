// Microbenchmark: tensor-memory read bandwidth seen by the MLP epilogues.  One CTA per SM; `nw` warps each read their
// 32-lane quadrant of a 128-lane x 64-column fp32 accumulator tile (tcgen05.ld 32x32b.x32, two per tile row half) `n`
// times.  Reports cycles per 128 x 64 tile and bytes per cycle per SM, for 4 and 8 reading warps and for 1-3 CTAs/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lidar-nerf_b200/csrc -o scripts/micro/tmem_ld.bin scripts/micro/tmem_ld.cu
#include <cstdio>
#include "mlp_tiles.cuh"
using namespace lnb;
using namespace lnb::tc;

__global__ void k_bench(uint32_t n, long long *out, uint32_t *sink) {
    __shared__ uint32_t s_slot;
    const uint32_t warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&s_slot), 128);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_slot;
    const uint32_t half = (warp >> 2) & 1u;
    const uint32_t addr = tmem + 32 * half + (((warp & 3u) * 32u) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t v[32];
        tmem_ld32(addr, v);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) acc ^= v[k];
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
    long long *d, h;
    uint32_t *sink;
    cudaMalloc(&d, sizeof(h));
    cudaMalloc(&sink, 4);
    const uint32_t n = 2000;
    for (uint32_t ctas_per_sm = 1; ctas_per_sm <= 3; ++ctas_per_sm)
        for (uint32_t nw = 4; nw <= 8; nw += 4) {
            k_bench<<<148 * ctas_per_sm, nw * 32>>>(n, d, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
            // per iteration the CTA reads nw x 32 lanes x 32 columns x 4 B
            const double bytes = (double)nw * 32 * 32 * 4, cyc = (double)h / n;
            printf("%u CTA/SM, %u reading warps: %6.1f cycles per x32 load round, %5.1f B/cycle/CTA, %5.1f B/cycle/SM "
                   "-> a 128x64 fp32 tile (32 KB) in %5.0f cycles\n", ctas_per_sm, nw, cyc, bytes / cyc,
                   bytes / cyc * ctas_per_sm, 32768.0 / (bytes / cyc * ctas_per_sm) );
        }
    return 0;
}
