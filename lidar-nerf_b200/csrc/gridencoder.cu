// Multiresolution hash / tiled grid encoding for sm_100a.
//
// Behavioural spec: lidarnerf/gridencoder/src/gridencoder.cu of the reference
//   index/hash      :53-93      forward  :95-263      backward :265-362     input grad :364-390
// Design: a CTA owns a tile of 128 samples and ALL levels.  Warp w walks levels w, w+W, ... so the
// 32 lanes of a gather instruction always hit the SAME level with 32 neighbouring samples (few
// distinct sectors on the dense coarse levels, maximal memory-level parallelism - 8 corner loads
// x 4 sample groups in flight per thread - on the hashed fine levels, whose 2-4 MiB tables live
// in B200's 126 MB L2).  Results are staged through a padded shared-memory tile and leave the SM
// as full coalesced rows, in either the reference's [L,B,C] layout or the [B,L*C] layout the MLP
// consumes (which removes the torch permute + copy the reference pays, grid.py:87,104).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "grid_common.cuh"

namespace lnb {
namespace {

constexpr int kTileB = 128;     // samples per CTA
constexpr int kFwdThreads = 512;
constexpr int kBwdThreads = 256;

// Forward.  kMinBlocks = resident CTAs per SM the register allocation is sized for (3: 40 registers, 2: 64).
template <typename T, uint32_t D, uint32_t C, int kMinBlocks = 3>
__global__ void __launch_bounds__(kFwdThreads, kMinBlocks)
k_grid_fwd(const float *__restrict__ inputs, const T *__restrict__ table,
           const int32_t *__restrict__ offsets, T *__restrict__ outputs, uint32_t B, uint32_t L,
           float S, uint32_t H, T *__restrict__ dy_dx, uint32_t gridtype, bool align_corners,
           uint32_t interp, int layout, float2 norm, const int32_t *__restrict__ n_active) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float s_in[kTileB * D];   // the tile's coordinates mapped to [0,1]: every warp (= level) reads all of them
    B = (layout == LNB_LAYOUT_LBC) ? B : active_rows(B, n_active);   // [L,B,C] addressing needs the true B
    if (blockIdx.x * kTileB >= B) return;
    T *tile = reinterpret_cast<T *>(smem_raw);  // [kTileB][pitch], only for LNB_LAYOUT_BLC
    const uint32_t F = L * C;
    const uint32_t pitch = F + (sizeof(T) == 2 ? 2 : 1);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const uint32_t b0 = blockIdx.x * kTileB;
    {
        const uint32_t n_in = min((uint32_t)kTileB, B - b0) * D;
        const float *src = inputs + (size_t)b0 * D;
        for (uint32_t i = threadIdx.x; i < kTileB * D; i += blockDim.x) {
            float x = 0.f;
            if (i < n_in) {
                x = __ldg(src + i);
                if (norm.x != 0.f) x = (x + norm.x) * norm.y;
            }
            s_in[i] = x;
        }
    }
    __syncthreads();

    for (uint32_t level = warp; level < L; level += n_warps) {
        const LevelGeo g = level_geo(offsets, level, S, H);
        const LevelIndex<D> li = level_index<D>(g, gridtype, align_corners);
        const T *__restrict__ tab = table + (size_t)g.table_offset * C;
#pragma unroll
        for (uint32_t grp = 0; grp < kTileB / 32; ++grp) {
            const uint32_t sl = grp * 32 + lane;
            const uint32_t b = b0 + sl;
            if (b >= B) continue;
            float v[D];
            bool inside = true;
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) {
                v[d] = s_in[sl * D + d];
                if (v[d] < 0 || v[d] > 1) inside = false;
            }
            const Cell<D> cell = locate_unit<D>(v, inside, g, align_corners, interp);

            T res[C];
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) res[c] = Num<T>::from_f(0.f);

            if (cell.inside) {
                if (li.generic) interp_corners<T, D, C, true>(cell, g, li, gridtype, align_corners, tab, res);
                else interp_corners<T, D, C, false>(cell, g, li, gridtype, align_corners, tab, res);
            }

            if (layout == LNB_LAYOUT_LBC) {
                T *o = outputs + ((size_t)level * B + b) * C;
#pragma unroll
                for (uint32_t c = 0; c < C; ++c) o[c] = res[c];
            } else {
                T *o = tile + (size_t)sl * pitch + level * C;
#pragma unroll
                for (uint32_t c = 0; c < C; ++c) o[c] = res[c];
            }

            if (dy_dx) {  // gridencoder.cu:214-262, layout [B, L, D, C]
                T *dd = dy_dx + ((size_t)b * L + level) * D * C;
#pragma unroll
                for (uint32_t gd = 0; gd < D; ++gd) {
                    T acc[C];
#pragma unroll
                    for (uint32_t c = 0; c < C; ++c) acc[c] = Num<T>::from_f(0.f);
                    if (cell.inside) {
#pragma unroll
                        for (uint32_t corner = 0; corner < (1u << (D - 1)); ++corner) {
                            float w = g.scale;
                            uint32_t p[D];
#pragma unroll
                            for (uint32_t nd = 0; nd < D - 1; ++nd) {
                                const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                                if ((corner & (1u << nd)) == 0) {
                                    w *= 1 - cell.frac[d];
                                    p[d] = cell.base[d];
                                } else {
                                    w *= cell.frac[d];
                                    p[d] = cell.base[d] + 1;
                                }
                            }
                            p[gd] = cell.base[gd];
                            const uint32_t r0 = cell_row<D>(p, gridtype, align_corners, g);
                            p[gd] = cell.base[gd] + 1;
                            const uint32_t r1 = cell_row<D>(p, gridtype, align_corners, g);
                            float v0[C], v1[C];
                            load_row<T, C>(tab + (size_t)r0 * C, v0);
                            load_row<T, C>(tab + (size_t)r1 * C, v1);
#pragma unroll
                            for (uint32_t c = 0; c < C; ++c) {
                                // reference: w * (right - left) * pos_deriv with (right-left) in T
                                const float diff = Num<T>::to_f(Num<T>::from_f(v1[c] - v0[c]));
                                acc[c] = Num<T>::from_f(Num<T>::to_f(acc[c]) + w * diff * cell.dfrac[gd]);
                            }
                        }
                    }
#pragma unroll
                    for (uint32_t c = 0; c < C; ++c) dd[gd * C + c] = acc[c];
                }
            }
        }
    }

    if (layout == LNB_LAYOUT_BLC) {
        __syncthreads();
        const uint32_t rows = min((uint32_t)kTileB, B - b0);
        T *out = outputs + (size_t)b0 * F;
        if ((F * sizeof(T)) % 4 == 0) {  // move 4-byte words
            const uint32_t wpr = F * sizeof(T) / 4;
            const uint32_t *tw = reinterpret_cast<const uint32_t *>(tile);
            uint32_t *ow = reinterpret_cast<uint32_t *>(out);
            const uint32_t pitch_w = pitch * sizeof(T) / 4;
            for (uint32_t i = threadIdx.x; i < rows * wpr; i += blockDim.x) {
                const uint32_t r = i / wpr, j = i - r * wpr;
                ow[i] = tw[r * pitch_w + j];
            }
        } else {
            for (uint32_t i = threadIdx.x; i < rows * F; i += blockDim.x) {
                const uint32_t r = i / F, j = i - r * F;
                out[i] = tile[r * pitch + j];
            }
        }
    }
}

template <typename T>
__device__ __forceinline__ void atomic_add_pair(T *p, float a, float b);
template <>
__device__ __forceinline__ void atomic_add_pair<__half>(__half *p, float a, float b) {
    // gridencoder.cu:346-353: each contribution is rounded to half, then added with a half2 atomic
    atomicAdd(reinterpret_cast<__half2 *>(p), __halves2half2(__float2half_rn(a), __float2half_rn(b)));
}
template <>
__device__ __forceinline__ void atomic_add_pair<float>(float *p, float a, float b) {
    atomicAdd(reinterpret_cast<float2 *>(p), make_float2(a, b));
}
__device__ __forceinline__ void atomic_add_one(__half *p, float a) { atomicAdd(p, __float2half_rn(a)); }
__device__ __forceinline__ void atomic_add_one(float *p, float a) { atomicAdd(p, a); }

// Backward: scatter-add of w * grad into the table gradient (gridencoder.cu:265-362).
// grid = (ceil(B / kBwdThreads), L); one thread per (sample, level), all C channels.
// TG = type of the incoming gradient, TA = type of the table gradient (TA = float with TG = half is the
// mixed mode of the fused training step: fp16 activations-grad, fp32 accumulation, no loss of small updates).
//
// One thread owns one sample and walks all levels: the coordinates are loaded once and the gradient row (L*C values,
// contiguous in the [B, L*C] layout) is read with 16-byte loads, instead of one dependent 4-byte load per
// (sample, level) thread (the profile of the first version showed exactly those two loads as its top stalls).
//
// Levels below `n_agg` use warp-level run aggregation: consecutive samples of a ray are a fraction of a cell apart
// on the coarse levels (37 samples per cell at level 0 of the KITTI config), so neighbouring lanes scatter into the
// SAME 2^D rows and plain atomics serialise in L2.  Lanes whose cell equals the previous lane's cell form a run; the
// 2^D corner contributions are summed over the run with a segmented shuffle scan and only the run's last lane issues
// the atomics (exact up to fp32 summation order).  A zero gradient row adds nothing and is skipped (padding samples
// all sit at one position and would otherwise serialise tens of thousands of atomics on the same rows).
constexpr uint32_t kMaxLevelsShared = 32;

template <typename TG, uint32_t C>
__device__ __forceinline__ void load_grad_row(const TG *__restrict__ gp, float (&gv)[C]) {
    if constexpr (sizeof(TG) == 2 && C % 2 == 0) {
#pragma unroll
        for (uint32_t c = 0; c < C; c += 2) {
            const float2 f = __half22float2(__ldg(reinterpret_cast<const __half2 *>(gp + c)));
            gv[c] = f.x, gv[c + 1] = f.y;
        }
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) gv[c] = Num<TG>::to_f(__ldg(gp + c));
    }
}

template <typename TG, typename TA, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(kBwdThreads)
k_grid_bwd(const TG *__restrict__ grad, const float *__restrict__ inputs,
           const int32_t *__restrict__ offsets, TA *__restrict__ grad_table, uint32_t B, uint32_t L,
           float S, uint32_t H, uint32_t gridtype, bool align_corners, uint32_t interp, int layout, float2 norm,
           uint32_t n_agg, const int32_t *__restrict__ n_active, const int32_t *__restrict__ row_idx,
           uint32_t level_lo, uint32_t level_hi) {
    __shared__ LevelGeo s_geo[kMaxLevelsShared];
    __shared__ LevelIndex<D> s_idx[kMaxLevelsShared];
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    // row_idx (nullable): gradient row b belongs to the sample at inputs[row_idx[b]], *n_active is the exact row count
    uint32_t Bact;
    if (row_idx) {
        const int32_t nv = *n_active;
        Bact = min(B, (uint32_t)(nv > 0 ? nv : 0));
    } else {
        Bact = active_rows(B, n_active);
    }
    if (blockIdx.x * blockDim.x >= Bact) return;
    if (threadIdx.x < min(L, kMaxLevelsShared)) {   // level-uniform quantities, once per CTA
        const LevelGeo g = level_geo(offsets, threadIdx.x, S, H);
        s_geo[threadIdx.x] = g;
        s_idx[threadIdx.x] = level_index<D>(g, gridtype, align_corners);
    }
    const bool in_range = b < Bact;
    const unsigned lane = lane_id();

    float v[D];
    bool inside = false;
    if (in_range) inside = load_unit_coords<D>(inputs + (size_t)(row_idx ? (uint32_t)__ldg(row_idx + b) : b) * D, norm, v);
    else {
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) v[d] = 0.f;
    }
    const bool has_grad = in_range && inside;
    const size_t g_stride = (layout == LNB_LAYOUT_LBC) ? (size_t)B * C : (size_t)C;
    const TG *gp = (layout == LNB_LAYOUT_LBC) ? grad + (size_t)b * C : grad + (size_t)b * L * C;
    float gv_next[C];
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) gv_next[c] = 0.f;
    if (has_grad) load_grad_row<TG, C>(gp, gv_next);
    __syncthreads();

    for (uint32_t level = 0; level < L; ++level) {
        float gv[C];
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) gv[c] = gv_next[c];
        if (has_grad && level + 1 < L) load_grad_row<TG, C>(gp + (size_t)(level + 1) * g_stride, gv_next);   // one level ahead

        LevelGeo g;
        LevelIndex<D> li;
        if (level < kMaxLevelsShared) {
            g = s_geo[level];
            li = s_idx[level];
        } else {
            g = level_geo(offsets, level, S, H);
            li = level_index<D>(g, gridtype, align_corners);
        }
        bool live = false;
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) live |= (gv[c] != 0.f);
        const bool agg = level < n_agg;      // warp-uniform
        if (level < level_lo || level >= level_hi) continue;   // diagnostic level window (all levels in production)
        if (!agg && !live) continue;

        const Cell<D> cell = locate_unit<D>(v, inside, g, align_corners, interp);
        float acc[1u << D][C];
#pragma unroll
        for (uint32_t corner = 0; corner < (1u << D); ++corner) {
            float w = 1;
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) {
                if ((corner & (1u << d)) == 0) w *= 1 - cell.frac[d];
                else w *= cell.frac[d];
            }
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) acc[corner][c] = w * gv[c];
        }

        bool issue = live;
        if (agg) {
            // runs = maximal groups of consecutive live lanes in the same cell
            bool head = (lane == 0) || !live;
            const int prev_live = __shfl_up_sync(kFullMask, (int)live, 1);
            if (!prev_live) head = true;
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) {
                const uint32_t pb = __shfl_up_sync(kFullMask, cell.base[d], 1);
                if (pb != cell.base[d]) head = true;
            }
            const unsigned head_mask = __ballot_sync(kFullMask, head);          // bit 0 is always set
            const unsigned run_start = 31u - (unsigned)__clz(head_mask & (kFullMask >> (31u - lane)));
            // segmented inclusive scan over the runs, only as deep as the longest run of this warp needs: the fine
            // levels (runs of one) pay a ballot and a REDUX, the coarse ones (tens of samples per cell) the full depth
            const unsigned need = __reduce_max_sync(kFullMask, lane - run_start);
            for (unsigned o = 1; o <= need; o <<= 1) {
                const bool take = lane >= run_start + o;
#pragma unroll
                for (uint32_t corner = 0; corner < (1u << D); ++corner) {
#pragma unroll
                    for (uint32_t c = 0; c < C; ++c) {
                        const float t = __shfl_up_sync(kFullMask, acc[corner][c], o);
                        if (take) acc[corner][c] += t;
                    }
                }
            }
            const bool tail = (lane == 31) || ((head_mask >> (lane + 1)) & 1u);
            issue = tail && live;
        }
        if (!issue) continue;

        TA *gt = grad_table + (size_t)g.table_offset * C;
        const CornerRows<D> cr(li, cell.base);
#pragma unroll
        for (uint32_t corner = 0; corner < (1u << D); ++corner) {
            uint32_t row;
            if (li.generic) {
                uint32_t p[D];
#pragma unroll
                for (uint32_t d = 0; d < D; ++d) p[d] = cell.base[d] + ((corner >> d) & 1u);
                row = cell_row<D>(p, gridtype, align_corners, g);
            } else {
                row = cr.row(corner);
            }
            TA *dst = gt + (size_t)row * C;
            if constexpr (C % 2 == 0) {
#pragma unroll
                for (uint32_t c = 0; c < C; c += 2) atomic_add_pair<TA>(dst + c, acc[corner][c], acc[corner][c + 1]);
            } else {
#pragma unroll
                for (uint32_t c = 0; c < C; ++c) atomic_add_one(dst + c, acc[corner][c]);
            }
        }
    }
}

// gridencoder.cu:364-390: grad_inputs[b,d] = sum_{l,c} grad[l,b,c] * dy_dx[b,l,d,c]
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(kBwdThreads)
k_grid_input_bwd(const T *__restrict__ grad, const T *__restrict__ dy_dx, T *__restrict__ grad_inputs,
                 uint32_t B, uint32_t L, int layout) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const T *dd = dy_dx + (size_t)b * L * D * C;
    T acc = Num<T>::from_f(0.f);  // accumulated in T like the reference (`scalar_t result`)
    for (uint32_t l = 0; l < L; ++l) {
        const T *gp = (layout == LNB_LAYOUT_LBC) ? grad + ((size_t)l * B + b) * C
                                                 : grad + ((size_t)b * L + l) * C;
#pragma unroll
        for (uint32_t c = 0; c < C; ++c)
            acc = Num<T>::from_f(Num<T>::to_f(acc) +
                                 Num<T>::to_f(Num<T>::from_f(Num<T>::to_f(gp[c]) *
                                                             Num<T>::to_f(dd[(l * D + d) * C + c]))));
    }
    grad_inputs[t] = acc;
}

// gridencoder.cu:695-808 of the reference: total-variation regulariser added to the table gradient.  For the cell that
// contains each input, on every level: grad[cell] += w * sum_nb (cell - nb) / sqrt(sum_nb (cell - nb)^2 + 1e-9), the
// neighbours being the <= 2 D axis neighbours inside [0, resolution]; w = weight / (2 D) in the table's type.
template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256)
k_grid_tv(const T *__restrict__ inputs, const T *__restrict__ table, T *__restrict__ grad,
          const int32_t *__restrict__ offsets, float weight, uint32_t B, uint32_t L, float S, uint32_t H,
          uint32_t gridtype, bool align_corners) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const LevelGeo g = level_geo(offsets, level, S, H);
    const T *tab = table + (size_t)g.table_offset * C;
    T *gt = grad + (size_t)g.table_offset * C;
    uint32_t pg[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        const float x = Num<T>::to_f(inputs[(size_t)b * D + d]);
        if (x < 0 || x > 1) return;
        pg[d] = (uint32_t)floorf(x * g.scale + (align_corners ? 0.0f : 0.5f));
    }
    T res[C], idelta[C];
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) res[c] = Num<T>::from_f(0.f), idelta[c] = Num<T>::from_f(0.f);
    const uint32_t index = cell_row<D>(pg, gridtype, align_corners, g) * C;
    const T w = Num<T>::from_f(weight / (2 * D));
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        const uint32_t cur = pg[d];
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (side == 0 ? !(cur < g.resolution) : !(cur > 0)) continue;
            pg[d] = side == 0 ? cur + 1 : cur - 1;
            const uint32_t nb = cell_row<D>(pg, gridtype, align_corners, g) * C;
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) {
                // every operation rounds to T, as the reference's scalar_t arithmetic does
                const T gv = Num<T>::from_f(Num<T>::to_f(tab[index + c]) - Num<T>::to_f(tab[nb + c]));
                res[c] = Num<T>::from_f(Num<T>::to_f(res[c]) + Num<T>::to_f(gv));
                idelta[c] = Num<T>::from_f(Num<T>::to_f(idelta[c]) +
                                           Num<T>::to_f(Num<T>::from_f(Num<T>::to_f(gv) * Num<T>::to_f(gv))));
            }
        }
        pg[d] = cur;
    }
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) {
        const T wr = Num<T>::from_f(Num<T>::to_f(w) * Num<T>::to_f(res[c]));
        atomic_add_one(gt + index + c, Num<T>::to_f(wr) * rsqrtf(Num<T>::to_f(idelta[c]) + 1e-9f));
    }
}

template <typename T, uint32_t D, uint32_t C>
int run_tv(const void *inputs, const void *emb, void *grad, const int32_t *offsets, float weight, uint32_t B, uint32_t L,
           float S, uint32_t H, uint32_t gridtype, bool ac, cudaStream_t st) {
    const dim3 grid(ceil_div<uint32_t>(B, 256), L, 1);
    k_grid_tv<T, D, C><<<grid, 256, 0, st>>>(static_cast<const T *>(inputs), static_cast<const T *>(emb),
                                             static_cast<T *>(grad), offsets, weight, B, L, S, H, gridtype, ac);
    count_launch();
    return launch_status();
}

inline uint32_t agg_max_resolution() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("LNB_GRID_AGG_MAX_RES");
        v = e ? atoi(e) : 1024;
        if (v < 0) v = 0;
    }
    return (uint32_t)v;
}

template <typename T, uint32_t D, uint32_t C>
int run_fwd(const float *inputs, const void *emb, const int32_t *offsets, void *out, uint32_t B,
            uint32_t L, float S, uint32_t H, void *dy_dx, uint32_t gridtype, bool ac, uint32_t interp,
            int layout, float2 norm, const int32_t *n_active, cudaStream_t st) {
    const uint32_t F = L * C;
    const size_t smem = (layout == LNB_LAYOUT_BLC)
                            ? (size_t)kTileB * (F + (sizeof(T) == 2 ? 2 : 1)) * sizeof(T) : 0;
    if (smem > 200 * 1024) return LNB_ERR_UNSUPPORTED;
    auto kern = k_grid_fwd<T, D, C>;
    if constexpr (sizeof(T) == 2 && D == 3 && C == 2) {
        static const int occ = [] { const char *e = getenv("LNB_GRID_FWD_OCC"); return e ? atoi(e) : 3; }();
        if (occ == 2) kern = k_grid_fwd<T, D, C, 2>;
    }
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const unsigned threads = (L >= 16) ? kFwdThreads : max(32u, min((unsigned)kFwdThreads, L * 32u));
    kern<<<ceil_div<uint32_t>(B, kTileB), threads, smem, st>>>(
        inputs, static_cast<const T *>(emb), offsets, static_cast<T *>(out), B, L, S, H,
        static_cast<T *>(dy_dx), gridtype, ac, interp, layout, norm, n_active);
    count_launch();
    return launch_status();
}

template <typename T, uint32_t D, uint32_t C>
int run_bwd(const void *grad, const float *inputs, const int32_t *offsets, void *grad_emb, uint32_t B,
            uint32_t L, float S, uint32_t H, const void *dy_dx, void *grad_inputs, uint32_t gridtype,
            bool ac, uint32_t interp, int layout, float2 norm, bool acc_f32, const int32_t *n_active, cudaStream_t st,
            const int32_t *row_idx = nullptr) {
    // leading (coarse) levels with resolution <= agg_max_res use the run-aggregating variant
    uint32_t n_agg = 0;
    {
        const uint32_t max_res = agg_max_resolution();
        for (uint32_t l = 0; l < L; ++l) {
            const float scale = exp2f((float)l * S) * (float)H - 1.0f;
            if ((uint32_t)ceilf(scale) + 1 <= max_res) n_agg = l + 1;
            else break;
        }
    }
    // diagnostic: LNB_GRID_BWD_LEVELS="lo,hi" restricts the scatter to levels [lo, hi) (scripts/diag_grid_levels.py)
    uint32_t lv_lo = 0, lv_hi = L;
    if (const char *e = getenv("LNB_GRID_BWD_LEVELS")) {
        unsigned a = 0, b = L;
        if (sscanf(e, "%u,%u", &a, &b) == 2) lv_lo = a, lv_hi = b < L ? b : L;
    }
    const unsigned bx = ceil_div<uint32_t>(B, kBwdThreads);
    if (acc_f32 && sizeof(T) == 2)
        k_grid_bwd<T, float, D, C><<<bx, kBwdThreads, 0, st>>>(static_cast<const T *>(grad), inputs, offsets,
                                                               static_cast<float *>(grad_emb), B, L, S, H, gridtype,
                                                               ac, interp, layout, norm, n_agg, n_active, row_idx,
                                                               lv_lo, lv_hi);
    else
        k_grid_bwd<T, T, D, C><<<bx, kBwdThreads, 0, st>>>(static_cast<const T *>(grad), inputs, offsets,
                                                           static_cast<T *>(grad_emb), B, L, S, H, gridtype, ac,
                                                           interp, layout, norm, n_agg, n_active, row_idx, lv_lo,
                                                           lv_hi);
    count_launch();
    int rc = launch_status();
    if (rc != LNB_OK) return rc;
    if (dy_dx) {
        if (!grad_inputs) return LNB_ERR_INVALID_ARGUMENT;
        k_grid_input_bwd<T, D, C><<<ceil_div<uint32_t>(B * D, kBwdThreads), kBwdThreads, 0, st>>>(
            static_cast<const T *>(grad), static_cast<const T *>(dy_dx), static_cast<T *>(grad_inputs),
            B, L, layout);
        count_launch();
        rc = launch_status();
    }
    return rc;
}

}  // namespace
}  // namespace lnb

using namespace lnb;

#define LNB_GRID_DISPATCH(FN, ...)                                                             \
    do {                                                                                       \
        if (dtype == LNB_F32) {                                                                \
            if (D == 3) {                                                                      \
                switch (C) {                                                                   \
                    case 1: return FN<float, 3, 1>(__VA_ARGS__);                               \
                    case 2: return FN<float, 3, 2>(__VA_ARGS__);                               \
                    case 4: return FN<float, 3, 4>(__VA_ARGS__);                               \
                    case 8: return FN<float, 3, 8>(__VA_ARGS__);                               \
                }                                                                              \
            } else if (D == 2) {                                                               \
                switch (C) {                                                                   \
                    case 1: return FN<float, 2, 1>(__VA_ARGS__);                               \
                    case 2: return FN<float, 2, 2>(__VA_ARGS__);                               \
                    case 4: return FN<float, 2, 4>(__VA_ARGS__);                               \
                    case 8: return FN<float, 2, 8>(__VA_ARGS__);                               \
                }                                                                              \
            }                                                                                  \
        } else if (dtype == LNB_F16) {                                                         \
            if (D == 3) {                                                                      \
                switch (C) {                                                                   \
                    case 1: return FN<__half, 3, 1>(__VA_ARGS__);                              \
                    case 2: return FN<__half, 3, 2>(__VA_ARGS__);                              \
                    case 4: return FN<__half, 3, 4>(__VA_ARGS__);                              \
                    case 8: return FN<__half, 3, 8>(__VA_ARGS__);                              \
                }                                                                              \
            } else if (D == 2) {                                                               \
                switch (C) {                                                                   \
                    case 1: return FN<__half, 2, 1>(__VA_ARGS__);                              \
                    case 2: return FN<__half, 2, 2>(__VA_ARGS__);                              \
                    case 4: return FN<__half, 2, 4>(__VA_ARGS__);                              \
                    case 8: return FN<__half, 2, 8>(__VA_ARGS__);                              \
                }                                                                              \
            }                                                                                  \
        }                                                                                      \
        return LNB_ERR_UNSUPPORTED;                                                            \
    } while (0)

extern "C" {

static float2 make_norm(float bound) {
    return bound > 0.f ? make_float2(bound, 1.0f / (2.0f * bound)) : make_float2(0.f, 0.f);
}

int lnb_grid_encode_forward_ex(const float *inputs, const void *embeddings, const int32_t *offsets,
                               void *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                               uint32_t H, void *dy_dx, uint32_t gridtype, int align_corners,
                               uint32_t interp, int dtype, int layout, float in_bound,
                               const int32_t *n_active, lnb_stream_t stream) {
    if (!inputs || !embeddings || !offsets || !outputs) return LNB_ERR_INVALID_ARGUMENT;
    if (gridtype > 1 || interp > 1 || (layout != LNB_LAYOUT_LBC && layout != LNB_LAYOUT_BLC) || L == 0)
        return LNB_ERR_INVALID_ARGUMENT;
    if (B == 0) return LNB_OK;
    cudaStream_t st = as_stream(stream);
    const bool ac = align_corners != 0;
    const float2 norm = make_norm(in_bound);
    LNB_GRID_DISPATCH(run_fwd, inputs, embeddings, offsets, outputs, B, L, S, H, dy_dx, gridtype, ac,
                      interp, layout, norm, n_active, st);
}

int lnb_grid_encode_forward(const float *inputs, const void *embeddings, const int32_t *offsets,
                            void *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                            uint32_t H, void *dy_dx, uint32_t gridtype, int align_corners,
                            uint32_t interp, int dtype, int layout, lnb_stream_t stream) {
    return lnb_grid_encode_forward_ex(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype,
                                      align_corners, interp, dtype, layout, 0.f, nullptr, stream);
}

int lnb_grid_encode_backward_ex(const void *grad, const float *inputs, const void *embeddings,
                                const int32_t *offsets, void *grad_embeddings, uint32_t B, uint32_t D,
                                uint32_t C, uint32_t L, float S, uint32_t H, const void *dy_dx,
                                void *grad_inputs, uint32_t gridtype, int align_corners,
                                uint32_t interp, int dtype, int layout, float in_bound,
                                int accumulate_f32, const int32_t *n_active, lnb_stream_t stream) {
    (void)embeddings;  // the table values are not needed for the table gradient (kept for ABI parity)
    if (!grad || !inputs || !offsets || !grad_embeddings) return LNB_ERR_INVALID_ARGUMENT;
    if (gridtype > 1 || interp > 1 || (layout != LNB_LAYOUT_LBC && layout != LNB_LAYOUT_BLC) || L == 0)
        return LNB_ERR_INVALID_ARGUMENT;
    if (B == 0) return LNB_OK;
    cudaStream_t st = as_stream(stream);
    const bool ac = align_corners != 0;
    if (accumulate_f32 && dy_dx) return LNB_ERR_UNSUPPORTED;
    const float2 norm = make_norm(in_bound);
    const bool acc32 = accumulate_f32 != 0;
    LNB_GRID_DISPATCH(run_bwd, grad, inputs, offsets, grad_embeddings, B, L, S, H, dy_dx, grad_inputs,
                      gridtype, ac, interp, layout, norm, acc32, n_active, st);
}

int lnb_grad_total_variation(const void *inputs, const void *embeddings, void *grad, const int32_t *offsets,
                             float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                             uint32_t gridtype, int align_corners, int dtype, lnb_stream_t stream) {
    if (!inputs || !embeddings || !grad || !offsets) return LNB_ERR_INVALID_ARGUMENT;
    if (gridtype > 1 || L == 0) return LNB_ERR_INVALID_ARGUMENT;
    if (B == 0) return LNB_OK;
    cudaStream_t st = as_stream(stream);
    const bool ac = align_corners != 0;
    LNB_GRID_DISPATCH(run_tv, inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, ac, st);
}

int lnb_grid_encode_backward_rows(const void *grad, const float *inputs, const void *embeddings,
                                  const int32_t *offsets, void *grad_embeddings, uint32_t B, uint32_t D,
                                  uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                  uint32_t interp, int dtype, float in_bound, int accumulate_f32,
                                  const int32_t *row_idx, const int32_t *n_rows, lnb_stream_t stream) {
    (void)embeddings;
    if (!grad || !inputs || !offsets || !grad_embeddings || !row_idx || !n_rows) return LNB_ERR_INVALID_ARGUMENT;
    if (gridtype > 1 || interp > 1 || L == 0) return LNB_ERR_INVALID_ARGUMENT;
    if (B == 0) return LNB_OK;
    cudaStream_t st = as_stream(stream);
    const bool ac = align_corners != 0;
    const float2 norm = make_norm(in_bound);
    const bool acc32 = accumulate_f32 != 0;
    const void *dy_dx = nullptr;
    void *grad_inputs = nullptr;
    const int layout = LNB_LAYOUT_BLC;
    LNB_GRID_DISPATCH(run_bwd, grad, inputs, offsets, grad_embeddings, B, L, S, H, dy_dx, grad_inputs,
                      gridtype, ac, interp, layout, norm, acc32, n_rows, st, row_idx);
}

int lnb_grid_encode_backward(const void *grad, const float *inputs, const void *embeddings,
                             const int32_t *offsets, void *grad_embeddings, uint32_t B, uint32_t D,
                             uint32_t C, uint32_t L, float S, uint32_t H, const void *dy_dx,
                             void *grad_inputs, uint32_t gridtype, int align_corners,
                             uint32_t interp, int dtype, int layout, lnb_stream_t stream) {
    return lnb_grid_encode_backward_ex(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H,
                                       dy_dx, grad_inputs, gridtype, align_corners, interp, dtype, layout, 0.f, 0,
                                       nullptr, stream);
}

}  // extern "C"
