// Fused Adam for the flat parameter vector (hash table + MLP weights), sm_100a.
//
// The reference leaves the optimiser to torch.optim.Adam + GradScaler (main_lidarnerf.py:389-391,
// nerf/utils.py:1221-1223), i.e. ~10 elementwise passes over the 13.7 M-row table per step.  This is
// one streaming pass: 16 B read of (p, g, m, v) x4 lanes, 12-14 B written per element; pure HBM
// roofline work (28-30 B / parameter / step), vectorised as float4 and launched as a persistent grid
// of 148 x 8 CTAs.
#include "common.cuh"

namespace lnb {
namespace {

constexpr int kThreads = 256;

struct AdamArgs {
    float lr, b1, b2, eps, inv_bc1, inv_sqrt_bc2, gscale;
};

__device__ __forceinline__ float adam_one(float &p, float g, float &m, float &v, const AdamArgs &a) {
    g *= a.gscale;
    m = a.b1 * m + (1.f - a.b1) * g;
    v = a.b2 * v + (1.f - a.b2) * g * g;
    const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;   // torch: sqrt(v)/sqrt(bc2) + eps
    p -= a.lr * a.inv_bc1 * (m / denom);
    return p;
}

// `hyper` (nullable, device): {lr, 1/bc1, 1/sqrt(bc2), grad_scale, enable} written by k_adam_set_hyper before the
// launch - the form a CUDA graph can hold (the by-value arguments of a captured launch are frozen, the step-dependent
// bias corrections and the scheduled learning rate are not).  enable == 0 makes the launch a no-op.
template <bool kHalfCopy, bool kZeroGrad>
__global__ void __launch_bounds__(kThreads)
k_adam(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
       __half *__restrict__ ph, size_t n, AdamArgs a, const float *__restrict__ hyper) {
    if (hyper) {
        if (hyper[4] == 0.f) return;
        a.lr = hyper[0], a.inv_bc1 = hyper[1], a.inv_sqrt_bc2 = hyper[2], a.gscale = hyper[3];
    }
    const size_t n4 = n / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float4 *p4 = reinterpret_cast<float4 *>(p), *g4 = reinterpret_cast<float4 *>(g);
    float4 *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
        adam_one(pp.x, gg.x, mm.x, vv.x, a);
        adam_one(pp.y, gg.y, mm.y, vv.y, a);
        adam_one(pp.z, gg.z, mm.z, vv.z, a);
        adam_one(pp.w, gg.w, mm.w, vv.w, a);
        p4[i] = pp;
        m4[i] = mm;
        v4[i] = vv;
        if (kZeroGrad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kHalfCopy) {
            __half2 lo = __floats2half2_rn(pp.x, pp.y), hi = __floats2half2_rn(pp.z, pp.w);
            uint2 packed;
            packed.x = *reinterpret_cast<unsigned *>(&lo);
            packed.y = *reinterpret_cast<unsigned *>(&hi);
            reinterpret_cast<uint2 *>(ph)[i] = packed;
        }
    }
    // tail (n % 4 elements)
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_one(pp, g[i], mm, vv, a);
        p[i] = pp, m[i] = mm, v[i] = vv;
        if (kZeroGrad) g[i] = 0.f;
        if (kHalfCopy) ph[i] = __float2half_rn(pp);
    }
}

// ----------------------------------------------------------------------------------------------------------------
// Data-parallel exchange fused with the optimiser, over NVLink peer memory (no NCCL on the data path):
//   reduce-scatter  : this rank sums ITS shard of the flat gradient straight out of every rank's gradient buffer
//                     (peer loads over NVLink / NVSwitch, fixed rank order -> every rank computes bit-identical sums),
//   Adam            : fp32 master weights and moments of the shard live only on this rank,
//   all-gather      : the updated fp16 parameters are stored straight into every rank's parameter shadow (peer stores).
// One streaming pass; the transfers overlap the arithmetic element by element.  The caller brackets the launch with two
// cross-rank barriers (all gradients complete before / all shadows complete after) - torch symmetric memory provides
// both the peer mappings and the device-side barriers.
// ----------------------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
struct PeerPtrs {
    const float *grad[kMaxPeers];
    __half *half[kMaxPeers];
};

__global__ void __launch_bounds__(kThreads)
k_dp_adam_exchange(PeerPtrs pp, uint32_t world, float *__restrict__ p, float *__restrict__ m, float *__restrict__ v,
                   size_t lo, size_t n, AdamArgs a) {
    const size_t n4 = n / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float4 *p4 = reinterpret_cast<float4 *>(p), *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 gq[kMaxPeers];
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)        // all peer loads in flight before the first use
            if (q < (int)world) gq[q] = __ldcv(reinterpret_cast<const float4 *>(pp.grad[q] + lo) + i);
        float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
            if (q < (int)world) gg.x += gq[q].x, gg.y += gq[q].y, gg.z += gq[q].z, gg.w += gq[q].w;
        float4 pq = p4[i], mm = m4[i], vv = v4[i];
        adam_one(pq.x, gg.x, mm.x, vv.x, a);
        adam_one(pq.y, gg.y, mm.y, vv.y, a);
        adam_one(pq.z, gg.z, mm.z, vv.z, a);
        adam_one(pq.w, gg.w, mm.w, vv.w, a);
        p4[i] = pq;
        m4[i] = mm;
        v4[i] = vv;
        __half2 h0 = __floats2half2_rn(pq.x, pq.y), h1 = __floats2half2_rn(pq.z, pq.w);
        uint2 packed;
        packed.x = *reinterpret_cast<unsigned *>(&h0);
        packed.y = *reinterpret_cast<unsigned *>(&h1);
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
            if (q < (int)world) reinterpret_cast<uint2 *>(pp.half[q] + lo)[i] = packed;
    }
}

// The same exchange through the NVSwitch's multicast / in-switch reduction (NVLS): the flat gradient and the fp16 shadow
// are additionally mapped at MULTICAST addresses (torch symmetric memory: `multicast_ptr`); one `multimem.ld_reduce`
// returns the sum over all ranks of a 16-byte gradient vector (the switch reduces, 1/world of the bytes of the peer-load
// form arrive at this GPU) and one `multimem.st` writes the updated parameters into every rank's shadow (the switch
// replicates).  Per rank and parameter the wire carries 4 B / world in + 2 B / world out instead of
// (world - 1) / world x (4 + 2) B.
__device__ __forceinline__ float4 mc_ld_reduce_f32x4(const float *mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(mc)
                 : "memory");
    return r;
}
__device__ __forceinline__ void mc_st_b128(void *mc, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}

__global__ void __launch_bounds__(kThreads)
k_dp_adam_exchange_mc(const float *__restrict__ grad_mc, __half *__restrict__ half_mc, float *__restrict__ p,
                      float *__restrict__ m, float *__restrict__ v, size_t lo, size_t n, AdamArgs a) {
    const size_t n8 = n / 8;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float4 *p4 = reinterpret_cast<float4 *>(p), *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const float4 g0 = mc_ld_reduce_f32x4(grad_mc + lo + 8 * i), g1 = mc_ld_reduce_f32x4(grad_mc + lo + 8 * i + 4);
        float4 pa = p4[2 * i], pb = p4[2 * i + 1], ma = m4[2 * i], mb = m4[2 * i + 1], va = v4[2 * i], vb = v4[2 * i + 1];
        adam_one(pa.x, g0.x, ma.x, va.x, a);
        adam_one(pa.y, g0.y, ma.y, va.y, a);
        adam_one(pa.z, g0.z, ma.z, va.z, a);
        adam_one(pa.w, g0.w, ma.w, va.w, a);
        adam_one(pb.x, g1.x, mb.x, vb.x, a);
        adam_one(pb.y, g1.y, mb.y, vb.y, a);
        adam_one(pb.z, g1.z, mb.z, vb.z, a);
        adam_one(pb.w, g1.w, mb.w, vb.w, a);
        p4[2 * i] = pa, p4[2 * i + 1] = pb;
        m4[2 * i] = ma, m4[2 * i + 1] = mb;
        v4[2 * i] = va, v4[2 * i + 1] = vb;
        const __half2 h0 = __floats2half2_rn(pa.x, pa.y), h1 = __floats2half2_rn(pa.z, pa.w);
        const __half2 h2 = __floats2half2_rn(pb.x, pb.y), h3 = __floats2half2_rn(pb.z, pb.w);
        uint4 packed;
        packed.x = *reinterpret_cast<const unsigned *>(&h0);
        packed.y = *reinterpret_cast<const unsigned *>(&h1);
        packed.z = *reinterpret_cast<const unsigned *>(&h2);
        packed.w = *reinterpret_cast<const unsigned *>(&h3);
        mc_st_b128(half_mc + lo + 8 * i, packed);
    }
}

__global__ void k_adam_set_hyper(float *hyper, float lr, float inv_bc1, float inv_sqrt_bc2, float gscale, float enable) {
    hyper[0] = lr, hyper[1] = inv_bc1, hyper[2] = inv_sqrt_bc2, hyper[3] = gscale, hyper[4] = enable;
}

int launch_adam(float *params, float *grad, float *exp_avg, float *exp_avg_sq, void *params_half, size_t n,
                const AdamArgs &a, const float *hyper, int zero_grad, cudaStream_t st) {
    const uintptr_t al = reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grad) |
                         reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq);
    if ((al & 15u) || (reinterpret_cast<uintptr_t>(params_half) & 7u)) return LNB_ERR_INVALID_ARGUMENT;
    if (n == 0) return LNB_OK;
    const size_t want = (n / 4 + kThreads - 1) / kThreads;
    const unsigned blocks = (unsigned)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
    __half *ph = static_cast<__half *>(params_half);
    if (ph) {
        if (zero_grad) k_adam<true, true><<<blocks, kThreads, 0, st>>>(params, grad, exp_avg, exp_avg_sq, ph, n, a, hyper);
        else k_adam<true, false><<<blocks, kThreads, 0, st>>>(params, grad, exp_avg, exp_avg_sq, ph, n, a, hyper);
    } else {
        if (zero_grad) k_adam<false, true><<<blocks, kThreads, 0, st>>>(params, grad, exp_avg, exp_avg_sq, ph, n, a, hyper);
        else k_adam<false, false><<<blocks, kThreads, 0, st>>>(params, grad, exp_avg, exp_avg_sq, ph, n, a, hyper);
    }
    count_launch();
    return launch_status();
}

}  // namespace
}  // namespace lnb

using namespace lnb;

extern "C" int lnb_adam_set_hyper(float *hyper_dev, float lr, float bias_correction1, float bias_correction2,
                                  float grad_scale, int enable, lnb_stream_t stream) {
    if (!hyper_dev) return LNB_ERR_INVALID_ARGUMENT;
    if (enable && (bias_correction1 == 0.f || bias_correction2 <= 0.f)) return LNB_ERR_INVALID_ARGUMENT;
    k_adam_set_hyper<<<1, 1, 0, as_stream(stream)>>>(hyper_dev, lr, enable ? 1.f / bias_correction1 : 0.f,
                                                     enable ? 1.f / sqrtf(bias_correction2) : 0.f, grad_scale,
                                                     enable ? 1.f : 0.f);
    count_launch();
    return launch_status();
}

extern "C" int lnb_adam_step_dev(float *params, float *grad, float *exp_avg, float *exp_avg_sq, void *params_half,
                                 size_t n, float beta1, float beta2, float eps, const float *hyper_dev, int zero_grad,
                                 lnb_stream_t stream) {
    if (!params || !grad || !exp_avg || !exp_avg_sq || !hyper_dev) return LNB_ERR_INVALID_ARGUMENT;
    AdamArgs a{0.f, beta1, beta2, eps, 0.f, 0.f, 0.f};
    return launch_adam(params, grad, exp_avg, exp_avg_sq, params_half, n, a, hyper_dev, zero_grad, as_stream(stream));
}

extern "C" int lnb_adam_step(float *params, float *grad, float *exp_avg, float *exp_avg_sq,
                             void *params_half, size_t n, float lr, float beta1, float beta2, float eps,
                             float bias_correction1, float bias_correction2, float grad_scale,
                             int zero_grad, lnb_stream_t stream) {
    if (!params || !grad || !exp_avg || !exp_avg_sq) return LNB_ERR_INVALID_ARGUMENT;
    if (bias_correction1 == 0.f || bias_correction2 <= 0.f) return LNB_ERR_INVALID_ARGUMENT;
    AdamArgs a{lr, beta1, beta2, eps, 1.f / bias_correction1, 1.f / sqrtf(bias_correction2), grad_scale};
    return launch_adam(params, grad, exp_avg, exp_avg_sq, params_half, n, a, nullptr, zero_grad, as_stream(stream));
}

extern "C" int lnb_dp_adam_exchange(const void *const *grad_ptrs, void *const *half_ptrs, uint32_t world,
                                    float *params_shard, float *exp_avg_shard, float *exp_avg_sq_shard, size_t shard_lo,
                                    size_t shard_n, float lr, float beta1, float beta2, float eps, float bias_correction1,
                                    float bias_correction2, float grad_scale, lnb_stream_t stream) {
    if (!grad_ptrs || !half_ptrs || !params_shard || !exp_avg_shard || !exp_avg_sq_shard) return LNB_ERR_INVALID_ARGUMENT;
    if (world < 1 || world > (uint32_t)kMaxPeers) return LNB_ERR_UNSUPPORTED;
    if (shard_n % 4 != 0 || shard_lo % 4 != 0) return LNB_ERR_INVALID_ARGUMENT;
    if (bias_correction1 == 0.f || bias_correction2 <= 0.f) return LNB_ERR_INVALID_ARGUMENT;
    PeerPtrs pp = {};
    for (uint32_t q = 0; q < world; ++q) {
        if (!grad_ptrs[q] || !half_ptrs[q]) return LNB_ERR_INVALID_ARGUMENT;
        pp.grad[q] = static_cast<const float *>(grad_ptrs[q]);
        pp.half[q] = static_cast<__half *>(half_ptrs[q]);
    }
    if (shard_n == 0) return LNB_OK;
    AdamArgs a{lr, beta1, beta2, eps, 1.f / bias_correction1, 1.f / sqrtf(bias_correction2), grad_scale};
    const size_t want = (shard_n / 4 + kThreads - 1) / kThreads;
    const unsigned blocks = (unsigned)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
    k_dp_adam_exchange<<<blocks, kThreads, 0, as_stream(stream)>>>(pp, world, params_shard, exp_avg_shard, exp_avg_sq_shard,
                                                                  shard_lo, shard_n, a);
    count_launch();
    return launch_status();
}

extern "C" int lnb_dp_adam_exchange_mc(const void *grad_multicast, void *half_multicast, float *params_shard,
                                       float *exp_avg_shard, float *exp_avg_sq_shard, size_t shard_lo, size_t shard_n,
                                       float lr, float beta1, float beta2, float eps, float bias_correction1,
                                       float bias_correction2, float grad_scale, lnb_stream_t stream) {
    if (!grad_multicast || !half_multicast || !params_shard || !exp_avg_shard || !exp_avg_sq_shard)
        return LNB_ERR_INVALID_ARGUMENT;
    if (shard_n % 8 != 0 || shard_lo % 8 != 0) return LNB_ERR_INVALID_ARGUMENT;
    if (bias_correction1 == 0.f || bias_correction2 <= 0.f) return LNB_ERR_INVALID_ARGUMENT;
    if (shard_n == 0) return LNB_OK;
    AdamArgs a{lr, beta1, beta2, eps, 1.f / bias_correction1, 1.f / sqrtf(bias_correction2), grad_scale};
    const size_t want = (shard_n / 8 + kThreads - 1) / kThreads;
    const unsigned blocks = (unsigned)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
    k_dp_adam_exchange_mc<<<blocks, kThreads, 0, as_stream(stream)>>>(static_cast<const float *>(grad_multicast),
                                                                     static_cast<__half *>(half_multicast), params_shard,
                                                                     exp_avg_shard, exp_avg_sq_shard, shard_lo, shard_n, a);
    count_launch();
    return launch_status();
}
