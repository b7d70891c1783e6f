// Device-side building blocks of the multiresolution hash / tiled grid (level geometry, row indexing, cell location,
// corner interpolation) shared by the stand-alone encoder kernels (gridencoder.cu) and the fused field kernel
// (field_fused.cu).  Behavioural spec: lidarnerf/gridencoder/src/gridencoder.cu:53-199 of the reference.
// Everything is file-local to the including translation unit (anonymous namespace).
#pragma once
#include "common.cuh"

namespace lnb {
namespace {

template <typename T> struct Num;
template <> struct Num<float> {
    static __device__ __forceinline__ float to_f(float v) { return v; }
    static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Num<__half> {
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};

struct LevelGeo {
    float scale;
    uint32_t resolution, hashmap_size, table_offset;
};

// gridencoder.cu:146-148
__device__ __forceinline__ LevelGeo level_geo(const int32_t *__restrict__ offsets, uint32_t level,
                                              float S, uint32_t H) {
    LevelGeo g;
    g.table_offset = (uint32_t)offsets[level];
    g.hashmap_size = (uint32_t)offsets[level + 1] - g.table_offset;
    g.scale = exp2f(level * S) * H - 1.0f;
    g.resolution = (uint32_t)ceilf(g.scale) + 1;
    return g;
}

// gridencoder.cu:53-93: dense (strided) index while the level fits, xor-prime hash otherwise.
template <uint32_t D>
__device__ __forceinline__ uint32_t cell_row(const uint32_t (&p)[D], uint32_t gridtype,
                                             bool align_corners, const LevelGeo &g) {
    constexpr uint32_t kPrimes[7] = {1u, 2654435761u, 805459861u, 3674653429u,
                                     2097192037u, 1434869437u, 2165219737u};
    const uint32_t side = align_corners ? g.resolution : (g.resolution + 1);
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        if (stride <= g.hashmap_size) {
            index += p[d] * stride;
            stride *= side;
        }
    }
    if (gridtype == 0 && stride > g.hashmap_size) {
        index = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) index ^= p[d] * kPrimes[d];
    }
    return index % g.hashmap_size;
}

// Per-level, per-sample row indexing with the level-uniform decisions of cell_row() hoisted out of the corner loop
// (the profile of the first version showed the generic `% hashmap_size` and the stride tests, repeated for each
// of the 2^D corners, as 15-20 % of both the gather and the scatter kernel).  Three cases, same results as cell_row():
//   dense   every dimension fits and side^D <= hashmap_size: row = sum p_d * stride_d, can never wrap;
//   hashed  with a power-of-two table (every hashed level the reference's sizing produces, grid.py:179-192):
//           row = xor p_d * prime_d, masked;
//   anything else (tiled grids that wrap, odd table sizes): the generic cell_row().
template <uint32_t D>
struct LevelIndex {
    uint32_t mul[D];
    uint32_t mask;      // all ones for dense levels
    bool hashed, generic;
};

template <uint32_t D>
__device__ __forceinline__ LevelIndex<D> level_index(const LevelGeo &g, uint32_t gridtype, bool align_corners) {
    constexpr uint32_t kPrimes[7] = {1u, 2654435761u, 805459861u, 3674653429u,
                                     2097192037u, 1434869437u, 2165219737u};
    LevelIndex<D> li;
    const uint32_t side = align_corners ? g.resolution : (g.resolution + 1);
    uint32_t stride = 1;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        if (stride <= g.hashmap_size) {
            li.mul[d] = stride;
            stride *= side;
        } else {
            li.mul[d] = 0;
        }
    }
    li.hashed = (gridtype == 0 && stride > g.hashmap_size);
    const bool pow2 = (g.hashmap_size & (g.hashmap_size - 1)) == 0;
    // align_corners: side = resolution, and an input of exactly 1.0 puts the upper corner at coordinate `side`, i.e. one
    // row past the dense block (weight 0, but the ADDRESS must stay inside the level): those levels keep the reference's
    // `% hashmap_size` through the generic path
    li.generic = li.hashed ? !pow2 : (stride > g.hashmap_size || align_corners);
    li.mask = li.hashed ? g.hashmap_size - 1 : 0xffffffffu;
    if (li.hashed) {
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) li.mul[d] = kPrimes[d];
    }
    return li;
}

// the two per-dimension terms (base, base + 1) of one sample on one level; row(corner) combines D of them
template <uint32_t D>
struct CornerRows {
    uint32_t t[D][2];
    uint32_t mask;
    bool hashed;
    __device__ __forceinline__ CornerRows(const LevelIndex<D> &li, const uint32_t (&base)[D]) {
        mask = li.mask;
        hashed = li.hashed;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            t[d][0] = base[d] * li.mul[d];
            t[d][1] = t[d][0] + li.mul[d];
        }
    }
    __device__ __forceinline__ uint32_t row(uint32_t corner) const {
        uint32_t v = t[0][corner & 1u];
#pragma unroll
        for (uint32_t d = 1; d < D; ++d) {
            const uint32_t u = t[d][(corner >> d) & 1u];
            v = hashed ? (v ^ u) : (v + u);
        }
        return v & mask;
    }
};

template <uint32_t D>
struct Cell {
    uint32_t base[D];
    float frac[D];    // interpolation weight along each axis (after optional smoothstep)
    float dfrac[D];   // its derivative w.r.t. the fractional position
    bool inside;
};

// gridencoder.cu:119-167
// `norm` = (bound, 1/(2 bound)) maps world coordinates in [-bound, bound] to [0,1] exactly as the reference's
// GridEncoder.forward does in torch, (x + bound) / (2 bound) (grid.py:213; torch divides by a scalar by
// multiplying with its fp32 reciprocal); norm.x == 0 means the inputs are already in [0,1].
template <uint32_t D>
__device__ __forceinline__ bool load_unit_coords(const float *__restrict__ x, float2 norm, float (&v)[D]) {
    bool inside = true;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        v[d] = (norm.x != 0.f) ? (x[d] + norm.x) * norm.y : x[d];
        if (v[d] < 0 || v[d] > 1) inside = false;
    }
    return inside;
}

template <uint32_t D>
__device__ __forceinline__ Cell<D> locate_unit(const float (&v)[D], bool inside, const LevelGeo &g,
                                               bool align_corners, uint32_t interp) {
    Cell<D> c;
    c.inside = inside;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        float pos = v[d] * g.scale + (align_corners ? 0.0f : 0.5f);
        const float fl = floorf(pos);
        c.base[d] = (uint32_t)fl;
        pos -= (float)c.base[d];
        if (interp == 1) {
            c.dfrac[d] = 6 * pos * (1.0f - pos);
            c.frac[d] = pos * pos * (3.0f - 2.0f * pos);
        } else {
            c.dfrac[d] = 1.0f;
            c.frac[d] = pos;
        }
    }
    return c;
}

template <uint32_t D>
__device__ __forceinline__ Cell<D> locate(const float *__restrict__ x, const LevelGeo &g,
                                          bool align_corners, uint32_t interp, float2 norm) {
    float v[D];
    const bool inside = load_unit_coords<D>(x, norm, v);
    return locate_unit<D>(v, inside, g, align_corners, interp);
}

template <typename T, uint32_t C>
__device__ __forceinline__ void load_row(const T *__restrict__ p, float (&v)[C]) {
    if constexpr (sizeof(T) == 2 && C % 2 == 0) {
#pragma unroll
        for (uint32_t c = 0; c < C; c += 2) {
            const __half2 h = __ldg(reinterpret_cast<const __half2 *>(p + c));
            v[c] = __low2float(h);
            v[c + 1] = __high2float(h);
        }
    } else if constexpr (sizeof(T) == 4 && C % 2 == 0) {
#pragma unroll
        for (uint32_t c = 0; c < C; c += 2) {
            const float2 f = __ldg(reinterpret_cast<const float2 *>(p + c));
            v[c] = f.x;
            v[c + 1] = f.y;
        }
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) v[c] = Num<T>::to_f(__ldg(p + c));
    }
}

// one table row with the widest load its size allows, left in the table's type
template <typename T, uint32_t C>
__device__ __forceinline__ void load_row_raw(const T *__restrict__ p, T (&v)[C]) {
    constexpr uint32_t kBytes = sizeof(T) * C;
    if constexpr (kBytes == 4) {
        *reinterpret_cast<uint32_t *>(v) = __ldg(reinterpret_cast<const uint32_t *>(p));
    } else if constexpr (kBytes == 8) {
        *reinterpret_cast<uint2 *>(v) = __ldg(reinterpret_cast<const uint2 *>(p));
    } else if constexpr (kBytes % 16 == 0) {
#pragma unroll
        for (uint32_t i = 0; i < kBytes / 16; ++i)
            reinterpret_cast<uint4 *>(v)[i] = __ldg(reinterpret_cast<const uint4 *>(p) + i);
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) v[c] = __ldg(p + c);
    }
}

// sum over the 2^D corners of w_corner * table[row(corner)], accumulated in the table's type exactly like the
// reference (`scalar_t results[C]`, gridencoder.cu:173-199): every `+= w * grid[...]` is rounded to T.
template <typename T, uint32_t D, uint32_t C, bool kGeneric>
__device__ __forceinline__ void interp_corners(const Cell<D> &cell, const LevelGeo &g, const LevelIndex<D> &li,
                                               uint32_t gridtype, bool align_corners, const T *__restrict__ tab,
                                               T (&res)[C]) {
    const CornerRows<D> cr(li, cell.base);
    // all gathers first (2^D independent loads in flight, kept in the table's type), then the rounding-ordered
    // accumulation
    __align__(16) T v[1u << D][C];
    // (Measured on B200 and rejected: loading the two x-neighbours of a hashed level as ONE aligned 8-byte pair when the
    // base x is even - they differ only in bit 0 of the xor - 115 -> 125 us; `ld.global.nc.L1::no_allocate` for the
    // hashed levels, 115 -> 180 us; 64 instead of 40 registers at 2 CTAs/SM, no change.)
#pragma unroll
    for (uint32_t corner = 0; corner < (1u << D); ++corner) {
        uint32_t row;
        if (kGeneric) {
            uint32_t p[D];
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) p[d] = cell.base[d] + ((corner >> d) & 1u);
            row = cell_row<D>(p, gridtype, align_corners, g);
        } else {
            row = cr.row(corner);
        }
        load_row_raw<T, C>(tab + (size_t)row * C, v[corner]);
    }
#pragma unroll
    for (uint32_t corner = 0; corner < (1u << D); ++corner) {
        float w = 1;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            if ((corner & (1u << d)) == 0) w *= 1 - cell.frac[d];
            else w *= cell.frac[d];
        }
#pragma unroll
        for (uint32_t c = 0; c < C; ++c)
            res[c] = Num<T>::from_f(Num<T>::to_f(res[c]) + w * Num<T>::to_f(v[corner][c]));
    }
}

}  // namespace
}  // namespace lnb
