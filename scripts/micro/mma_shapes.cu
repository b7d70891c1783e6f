// Microbenchmark: tensor-pipe time per tcgen05.mma (kind::f16, cta_group::1, operands in shared memory) for the
// shapes / operand majors the MLP backward uses.  One CTA, one issuing warp, `n` back-to-back MMAs, one commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lidar-nerf_b200/csrc -o /tmp/mma_shapes scripts/micro/mma_shapes.cu
#include <cstdio>
#include "mlp_tiles.cuh"
using namespace lnb;
using namespace lnb::tc;

__global__ void k_bench(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn, uint32_t n, uint32_t same_d, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t s_a = sbase, s_b = sbase + 4 * kTileBytes, s_bar = sbase + 8 * kTileBytes, s_slot = s_bar + 8;
    for (uint32_t q = threadIdx.x; q < 8 * kTileBytes / 16; q += blockDim.x) sts128(sbase + q * 16, make_uint4(0, 0, 0, 0));
    if (threadIdx.x < 32) tmem_alloc(s_slot, 512);
    if (threadIdx.x == 0) { mbar_init(s_bar, 1); mbar_init_fence(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = lds32(s_slot);
    if (threadIdx.x < 32) {
        const uint32_t idesc = instr_desc_f16(M, N, a_mn, b_mn);
        const uint64_t a0 = smem_desc_sw128(s_a, a_mn ? kTileBytes : 16), b0 = smem_desc_sw128(s_b, b_mn ? kTileBytes : 16);
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (uint32_t i = 0; i < n; ++i) {
                const uint32_t k = i & 3;
                mma_f16_elect(tmem + (same_d ? 0 : (i & 1) * 256), desc_step(a0, k, a_mn ? 2048 : 32), desc_step(b0, k, b_mn ? 2048 : 32),
                              idesc, i > 1 ? 1u : 0u);
            }
            const long long t1 = clock64();
            mma_commit_elect(s_bar);
            mbar_wait_warp(s_bar, rep & 1);
            const long long t2 = clock64();
            if (threadIdx.x == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    long long *d, h[6];
    cudaMalloc(&d, sizeof(h));
    const size_t smem = 8 * kTileBytes + 2048;
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    struct { uint32_t M, N, a, b; const char *what; } cfg[] = {
        {64, 64, 1, 1, "wgrad today       M64  N64  A mn  B mn"},
        {64, 96, 1, 1, "wgrad W_in today  M64  N96  A mn  B mn"},
        {64, 16, 1, 1, "wgrad W_out today M64  N16  A mn  B mn"},
        {128, 64, 1, 1, "                  M128 N64  A mn  B mn"},
        {128, 128, 1, 1, "stacked wgrad     M128 N128 A mn  B mn"},
        {128, 96, 1, 1, "stacked wgrad     M128 N96  A mn  B mn"},
        {128, 64, 0, 1, "dgrad today       M128 N64  A k   B mn"},
        {128, 64, 0, 0, "forward           M128 N64  A k   B k "},
        {64, 64, 0, 0, "                  M64  N64  A k   B k "},
        {128, 256, 1, 1, "                  M128 N256 A mn  B mn"},
    };
    for (auto &c : cfg)
        for (uint32_t same_d = 0; same_d < 2; ++same_d) {
            const uint32_t n = 64;
            k_bench<<<1, 128, smem>>>(c.M, c.N, c.a, c.b, n, same_d, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", c.what, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("%s  %s: issue %5.1f cyc/MMA, issue+drain %5.1f cyc/MMA (floor max(M,128)*N/256 = %u)\n", c.what,
                   same_d ? "one accumulator " : "two accumulators", h[4] / (double)n, h[5] / (double)n,
                   (c.M > 128 ? c.M : 128) * c.N / 256);
        }
    return 0;
}
