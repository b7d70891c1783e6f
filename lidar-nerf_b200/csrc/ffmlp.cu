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
#include "mlp_tiles.cuh"
#include "mlp_bwd.cuh"

namespace lnb {
namespace {

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
        if (warp == 0) {   // warp-collective MMA issue (one elected lane)
            fence_after_sync();
            for (uint32_t t = 0; t < sh.kt_in; ++t) {
                const uint32_t cols = min(64u, sh.in_dim - t * 64);
                issue_kmajor(d_hid, s_x + t * kTileBytes, s_win + t * kWTileBytes, cols / 16, kIdescFwdHid, t > 0);
            }
            mma_commit_elect(s_bar);
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
            if (warp == 0) {
                fence_after_sync();
                if (layer < sh.n_hid)
                    issue_kmajor(d_hid, s_h, s_whid + layer * kWTileBytes, 4, kIdescFwdHid, false);
                else
                    issue_kmajor(d_out, s_h, s_wout, 4, kIdescFwdOut, false);
                mma_commit_elect(s_bar);
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

// bf16 build of this unit (-DLNB_BF16, see mlp_tiles.cuh): every entry point gets the suffix `_bf16`
#ifdef LNB_BF16
#define lnb_ffmlp_forward_ex lnb_ffmlp_forward_ex_bf16
#define lnb_ffmlp_forward lnb_ffmlp_forward_bf16
#define lnb_ffmlp_inference lnb_ffmlp_inference_bf16
#define lnb_ffmlp_backward_workspace_bytes lnb_ffmlp_backward_workspace_bytes_bf16
#define lnb_ffmlp_backward_accumulate_rows lnb_ffmlp_backward_accumulate_rows_bf16
#define lnb_ffmlp_backward_accumulate lnb_ffmlp_backward_accumulate_bf16
#define lnb_ffmlp_backward lnb_ffmlp_backward_bf16
#define lnb_allocate_splitk lnb_allocate_splitk_bf16
#define lnb_free_splitk lnb_free_splitk_bf16
#define lnb_debug_bwd_trace_generic lnb_debug_bwd_trace_generic_bf16
#endif

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
                               void *grad_inputs, float *wgrad_f32, const int32_t *n_active, cudaStream_t st,
                               const int32_t *row_idx = nullptr) {
    BwdArgs a = {};
    a.row_idx = row_idx;
    a.W = static_cast<const __half *>(weights);
    a.fbuf = static_cast<const __half *>(forward_buffer);
    a.B = B;
    a.sh = sh;
    a.wgrad = wgrad_f32;
    a.n_active = n_active;
    a.G = static_cast<const __half *>(grad);
    a.X = static_cast<const __half *>(inputs);
    a.bbuf = static_cast<__half *>(backward_buffer);
    a.dX = calc_grad_inputs ? static_cast<__half *>(grad_inputs) : nullptr;
    const int rc = launch_mlp_bwd<false, 0, 0>(a, (uint32_t)sm_count(), st);
    if (rc != LNB_OK) return rc;
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

int lnb_ffmlp_backward_accumulate_rows(const void *grad, const void *inputs, const void *weights,
                                       const void *forward_buffer, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                                       uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                                       uint32_t output_activation, int calc_grad_inputs, void *grad_inputs,
                                       float *grad_weights_f32, const int32_t *row_idx, const int32_t *n_rows,
                                       lnb_stream_t stream) {
    if (!grad || !inputs || !weights || !forward_buffer || !grad_weights_f32 || !row_idx || !n_rows)
        return LNB_ERR_INVALID_ARGUMENT;
    if (calc_grad_inputs && !grad_inputs) return LNB_ERR_INVALID_ARGUMENT;
    Shape sh;
    int rc = check_shape(B, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, &sh);
    if (rc != LNB_OK) return rc;
    if (B == 0) return LNB_OK;
    return ffmlp_backward_impl(grad, inputs, weights, forward_buffer, B, sh, calc_grad_inputs, nullptr, grad_inputs,
                               grad_weights_f32, n_rows, as_stream(stream), row_idx);
}

int lnb_allocate_splitk(size_t size) {
    (void)size;
    return LNB_OK;
}
int lnb_free_splitk(void) { return LNB_OK; }

#ifdef LNB_TRACE
// diagnostic builds only (build.py --trace): timeline recorded by CTA 0 of the generic backward kernel
int lnb_debug_bwd_trace_generic(unsigned long long *host_out, uint32_t max_events, int reset) {
    unsigned int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_bwd_trace_n, sizeof(n));
    if (n > 16384u) n = 16384u;
    if (n > max_events) n = max_events;
    if (host_out && n) cudaMemcpyFromSymbol(host_out, g_bwd_trace, sizeof(unsigned long long) * n);
    if (reset) {
        const unsigned int zero = 0;
        cudaMemcpyToSymbol(g_bwd_trace_n, &zero, sizeof(zero));
    }
    return (int)n;
}
#endif

}  // extern "C"
