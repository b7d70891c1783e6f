// Real spherical-harmonics encoding (degree <= 8, 64 outputs) for sm_100a.
//
// Behavioural spec: lidarnerf/shencoder/src/shencoder.cu of the reference (:31-832 forward with
// hard-coded polynomials, :835-858 backward).  The reference spells every basis function out as an
// expanded Cartesian polynomial.  Those polynomials are exactly
//     Y_{l,+m}(x,y,z) = N_l^m * Q_l^m(z) * Re (x + i y)^m
//     Y_{l,-m}(x,y,z) = N_l^m * Q_l^m(z) * Im (x + i y)^m          (index l*l + l +- m)
// with Q_l^m = d^m P_l / dz^m (a polynomial in z only) and
//     N_l^m = (-1)^m * sqrt(2 - [m==0]) * sqrt((2l+1)/(4 pi) * (l-m)!/(l+m)!),
// evaluated WITHOUT normalising (x,y,z) (checked against the reference's constants, e.g.
// Y_00 = 0.28209479, Y_1 = (-0.4886 y, 0.4886 z, -0.4886 x), Y_22 = 0.5463 (x^2 - y^2)).
// This file evaluates that closed form by recurrence, fully unrolled at compile time, and gets the
// analytic derivatives from d/dx Re(x+iy)^m = m Re(x+iy)^(m-1), d/dy Re = -m Im, d/dx Im = m Im,
// d/dy Im = m Re, d/dz Q_l^m = Q_l^(m+1).
#include "common.cuh"

namespace lnb {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxDeg = 8;

// N_l^m, row l holds m = 0..l   (values generated from the formula above in double precision)
__constant__ float kShNorm[kMaxDeg][kMaxDeg] = {
    {0.28209479177387814f},
    {0.48860251190291992f, -0.48860251190291998f},
    {0.63078313050504009f, -0.36418281019735976f, 0.18209140509867988f},
    {0.7463526651802308f, -0.3046971996429772f, 0.096353714754685155f, -0.039336239328442907f},
    {0.84628437532163447f, -0.26761861742291571f, 0.063078313050504001f, -0.016858388283618388f,
     0.0059603403376112026f},
    {0.9356025796273888f, -0.24157154730437169f, 0.045652731285460234f, -0.0093188247511476283f,
     0.0021964680580751762f, -0.00069458418713245519f},
    {1.0171072362820548f, -0.22195099524523101f, 0.03509353369580661f, -0.0058489222826344353f,
     0.0010678622237644956f, -0.00022766899107568562f, 6.5722376641838803e-05f},
    {1.0925484305920792f, -0.20647224590289676f, 0.028097313806030647f, -0.0039735602250741348f,
     0.00059903674311141165f, -9.9839457185235285e-05f, 1.9580128477462541e-05f,
     -5.233009453691466e-06f},
};

template <int DEG, bool kGrad>
__global__ void __launch_bounds__(kThreads)
k_sh_fwd(const float *__restrict__ inputs, float *__restrict__ outputs, uint32_t B,
         float *__restrict__ dy_dx) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    constexpr int C2 = DEG * DEG;
    const float x = inputs[(size_t)b * 3], y = inputs[(size_t)b * 3 + 1], z = inputs[(size_t)b * 3 + 2];

    // azimuthal part: A[m] + i B[m] = (x + i y)^m
    float A[DEG + 1], Bm[DEG + 1];
    A[0] = 1.f, Bm[0] = 0.f;
#pragma unroll
    for (int m = 1; m <= DEG; ++m) {
        A[m] = x * A[m - 1] - y * Bm[m - 1];
        Bm[m] = x * Bm[m - 1] + y * A[m - 1];
    }

    // polar part: Q[l][m] = d^m P_l / dz^m  (Q[l][m] = 0 for m > l; one extra column for d/dz)
    float Q[DEG][DEG + 1];
#pragma unroll
    for (int l = 0; l < DEG; ++l)
#pragma unroll
        for (int m = 0; m <= DEG; ++m) Q[l][m] = 0.f;
    {
        float dfact = 1.f;  // (2m-1)!!
#pragma unroll
        for (int m = 0; m < DEG; ++m) {
            if (m > 0) dfact *= (float)(2 * m - 1);
            Q[m][m] = dfact;
            if (m + 1 < DEG) Q[m + 1][m] = (float)(2 * m + 1) * z * dfact;
#pragma unroll
            for (int l = m + 2; l < DEG; ++l)
                Q[l][m] = ((float)(2 * l - 1) * z * Q[l - 1][m] - (float)(l + m - 1) * Q[l - 2][m]) *
                          (1.0f / (float)(l - m));
        }
    }

    float *o = outputs + (size_t)b * C2;
    float *gx = nullptr, *gy = nullptr, *gz = nullptr;
    if (kGrad) {
        gx = dy_dx + (size_t)b * 3 * C2;  // [B, 3, C2]: d/dx block, d/dy block, d/dz block
        gy = gx + C2;
        gz = gy + C2;
    }
#pragma unroll
    for (int l = 0; l < DEG; ++l) {
#pragma unroll
        for (int m = 0; m <= l; ++m) {
            const float nq = kShNorm[l][m] * Q[l][m];
            const float nq1 = kShNorm[l][m] * Q[l][m + 1];
            const int ip = l * l + l + m, im = l * l + l - m;
            o[ip] = nq * A[m];
            if (m > 0) o[im] = nq * Bm[m];
            if (kGrad) {
                if (m == 0) {
                    gx[ip] = 0.f;
                    gy[ip] = 0.f;
                    gz[ip] = nq1;
                } else {
                    const float fm = (float)m;
                    gx[ip] = nq * fm * A[m - 1];
                    gy[ip] = -nq * fm * Bm[m - 1];
                    gz[ip] = nq1 * A[m];
                    gx[im] = nq * fm * Bm[m - 1];
                    gy[im] = nq * fm * A[m - 1];
                    gz[im] = nq1 * Bm[m];
                }
            }
        }
    }
}

// shencoder.cu:835-858: grad_inputs[b,d] += sum_ch grad[b,ch] * dy_dx[b,d,ch]
__global__ void __launch_bounds__(kThreads)
k_sh_bwd(const float *__restrict__ grad, uint32_t B, uint32_t C2, const float *__restrict__ dy_dx,
         float *__restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 3) return;
    const uint32_t b = t / 3, d = t - b * 3;
    const float *g = grad + (size_t)b * C2;
    const float *dd = dy_dx + ((size_t)b * 3 + d) * C2;
    float acc = grad_inputs[t];
    for (uint32_t ch = 0; ch < C2; ++ch) acc += g[ch] * dd[ch];
    grad_inputs[t] = acc;
}

template <int DEG>
int run_sh(const float *in, float *out, uint32_t B, float *dy_dx, cudaStream_t st) {
    const unsigned blocks = ceil_div<uint32_t>(B, kThreads);
    if (dy_dx) k_sh_fwd<DEG, true><<<blocks, kThreads, 0, st>>>(in, out, B, dy_dx);
    else k_sh_fwd<DEG, false><<<blocks, kThreads, 0, st>>>(in, out, B, nullptr);
    count_launch();
    return launch_status();
}

}  // namespace
}  // namespace lnb

using namespace lnb;

extern "C" {

int lnb_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t C,
                          float *dy_dx, lnb_stream_t stream) {
    if (!inputs || !outputs) return LNB_ERR_INVALID_ARGUMENT;
    if (D != 3 || C < 1 || C > kMaxDeg) return LNB_ERR_UNSUPPORTED;
    if (B == 0) return LNB_OK;
    cudaStream_t st = as_stream(stream);
    switch (C) {
        case 1: return run_sh<1>(inputs, outputs, B, dy_dx, st);
        case 2: return run_sh<2>(inputs, outputs, B, dy_dx, st);
        case 3: return run_sh<3>(inputs, outputs, B, dy_dx, st);
        case 4: return run_sh<4>(inputs, outputs, B, dy_dx, st);
        case 5: return run_sh<5>(inputs, outputs, B, dy_dx, st);
        case 6: return run_sh<6>(inputs, outputs, B, dy_dx, st);
        case 7: return run_sh<7>(inputs, outputs, B, dy_dx, st);
        default: return run_sh<8>(inputs, outputs, B, dy_dx, st);
    }
}

int lnb_sh_encode_backward(const float *grad, const float *inputs, uint32_t B, uint32_t D, uint32_t C,
                           const float *dy_dx, float *grad_inputs, lnb_stream_t stream) {
    (void)inputs;
    if (!grad || !dy_dx || !grad_inputs) return LNB_ERR_INVALID_ARGUMENT;
    if (D != 3 || C < 1 || C > kMaxDeg) return LNB_ERR_UNSUPPORTED;
    if (B == 0) return LNB_OK;
    k_sh_bwd<<<ceil_div<uint32_t>(B * 3, kThreads), kThreads, 0, as_stream(stream)>>>(grad, B, C * C,
                                                                                     dy_dx, grad_inputs);
    count_launch();
    return launch_status();
}

}  // extern "C"
