// NeRF frequency encoding for sm_100a.
// Behavioural spec: lidarnerf/freqencoder/src/freqencoder.cu:34-101 of the reference.
//   out[b] = [ x, sin(2^0 x), cos(2^0 x), sin(2^1 x), cos(2^1 x), ... ]   (each block D wide),
//   cos is evaluated as __sinf(. + pi/2), and 2^f x is an exact exponent shift (scalbnf).
// Layout: a thread owns one (sample, input-dim) pair and produces its 1 + 2*deg outputs; the
// 2^f scaling is a running exact doubling.  Stores of a warp cover 32*D consecutive floats per
// frequency band (coalesced for D = 3).
#include "common.cuh"

namespace lnb {
namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
k_freq_fwd(const float *__restrict__ inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
           float *__restrict__ outputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float x = inputs[t];
    float *o = outputs + (size_t)b * C + d;
    o[0] = x;
    const float half_pi = 3.141592653589793f / 2;
    for (uint32_t f = 0; f < deg; ++f) {
        const float a = scalbnf(x, (int)f);
        o[(2 * f + 1) * D] = __sinf(a + 0 * half_pi);   // phase 0  (freqencoder.cu:60-61)
        o[(2 * f + 2) * D] = __sinf(a + 1 * half_pi);   // phase pi/2
    }
}

// freqencoder.cu:68-101: gx_d = g_d + sum_f 2^f (g_sin * out_cos - g_cos * out_sin)
__global__ void __launch_bounds__(kThreads)
k_freq_bwd(const float *__restrict__ grad, const float *__restrict__ outputs, uint32_t B, uint32_t D,
           uint32_t deg, uint32_t C, float *__restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float *g = grad + (size_t)b * C + d;
    const float *o = outputs + (size_t)b * C + d;
    float acc = g[0];
    for (uint32_t f = 0; f < deg; ++f) {
        const float gs = g[(2 * f + 1) * D], gc = g[(2 * f + 2) * D];
        const float os = o[(2 * f + 1) * D], oc = o[(2 * f + 2) * D];
        acc += scalbnf(1.0f, (int)f) * (gs * oc - gc * os);
    }
    grad_inputs[t] = acc;
}

}  // namespace
}  // namespace lnb

using namespace lnb;

extern "C" {

int lnb_freq_encode_forward(const float *inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                            float *outputs, lnb_stream_t stream) {
    if (!inputs || !outputs) return LNB_ERR_INVALID_ARGUMENT;
    if (C != D + 2 * D * deg || D == 0) return LNB_ERR_INVALID_ARGUMENT;
    if (B == 0) return LNB_OK;
    k_freq_fwd<<<ceil_div<uint32_t>(B * D, kThreads), kThreads, 0, as_stream(stream)>>>(inputs, B, D, deg,
                                                                                      C, outputs);
    count_launch();
    return launch_status();
}

int lnb_freq_encode_backward(const float *grad, const float *outputs, uint32_t B, uint32_t D,
                             uint32_t deg, uint32_t C, float *grad_inputs, lnb_stream_t stream) {
    if (!grad || !outputs || !grad_inputs) return LNB_ERR_INVALID_ARGUMENT;
    if (C != D + 2 * D * deg || D == 0) return LNB_ERR_INVALID_ARGUMENT;
    if (B == 0) return LNB_OK;
    k_freq_bwd<<<ceil_div<uint32_t>(B * D, kThreads), kThreads, 0, as_stream(stream)>>>(
        grad, outputs, B, D, deg, C, grad_inputs);
    count_launch();
    return launch_status();
}

}  // extern "C"
