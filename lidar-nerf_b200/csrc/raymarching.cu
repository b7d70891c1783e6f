// Occupancy-grid ray marching + alpha compositing for sm_100a.
//
// Behavioural spec: lidarnerf/raymarching/src/raymarching.cu of the reference (cited per kernel).
// Design (not a port): the reference walks each ray with ONE thread and two serial while-loops.
// Here a WARP owns a ray.  The key observation is that every update of the ray parameter t in
// the reference, whether it is an "occupied" step or one iteration of the empty-space do/while,
// is the same recurrence  t <- t + clamp(t*dt_gamma, dt_min, dt_max)   (raymarching.cu:390,416,436).
// The sequence of candidate positions t_0, t_1, ... is therefore independent of the occupancy
// grid; the grid only decides which candidates are visited/emitted.  A warp evaluates 32
// consecutive candidates at once (32 independent bitfield probes in flight instead of one),
// resolves the visit chain with ballots + a shuffle binary search, and writes the emitted
// samples with coalesced stores.  All per-candidate arithmetic keeps the reference's operand
// order and types so that sample counts and positions are bit-identical.
//
// Compositing likewise uses a warp per ray: 32 samples per step, transmittance via a
// multiplicative warp scan, early termination via ballot.
#include <cfloat>
#include <cmath>

#include "common.cuh"

namespace lnb {

unsigned long long g_launch_count = 0;

namespace {

constexpr int kThreads = 128;  // 4 warps = 4 rays per CTA

// ------------------------------------------------------------------------------------------
// small per-ray kernels
// ------------------------------------------------------------------------------------------

// raymarching.cu:105-157: slab test against an axis-aligned box.
__global__ void __launch_bounds__(kThreads)
k_near_far(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
           const float *__restrict__ aabb, uint32_t N, float min_near, float *__restrict__ nears,
           float *__restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float lo = -FLT_MAX, hi = FLT_MAX;  // running intersection of the three slabs
    bool miss = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float o = rays_o[n * 3 + a];
        const float r = 1 / rays_d[n * 3 + a];
        float t0 = (aabb[a] - o) * r;
        float t1 = (aabb[a + 3] - o) * r;
        if (t0 > t1) {
            const float s = t0;
            t0 = t1;
            t1 = s;
        }
        if (a == 0) {
            lo = t0;
            hi = t1;
        } else if (!miss) {
            if (lo > t1 || t0 > hi) miss = true;
            else {
                if (t0 > lo) lo = t0;
                if (t1 < hi) hi = t1;
            }
        }
    }
    if (miss) {
        nears[n] = fars[n] = FLT_MAX;
        return;
    }
    if (lo < min_near) lo = min_near;
    nears[n] = lo;
    fars[n] = hi;
}

// raymarching.cu:183-217: far intersection with the background sphere -> (theta, phi) in [-1,1]^2.
__global__ void __launch_bounds__(kThreads)
k_sph_from_ray(const float *__restrict__ rays_o, const float *__restrict__ rays_d, float radius,
               uint32_t N, float *__restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float B = ox * dx + oy * dy + oz * dz;
    const float C = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-B + sqrtf(B * B - A * C)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    const float inv_pi = 0.3183098861837907f;
    coords[n * 2] = 2 * theta * inv_pi - 1;
    coords[n * 2 + 1] = phi * inv_pi;
}

__global__ void __launch_bounds__(kThreads)
k_morton(const int32_t *__restrict__ coords, uint32_t N, int32_t *__restrict__ indices) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    indices[n] = (int32_t)morton_encode((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1],
                                        (uint32_t)coords[n * 3 + 2]);
}

__global__ void __launch_bounds__(kThreads)
k_morton_invert(const int32_t *__restrict__ indices, uint32_t N, int32_t *__restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int32_t v = indices[n];  // arithmetic shifts on the signed value, as the reference does
    coords[n * 3] = (int32_t)compact3((uint32_t)(v >> 0));
    coords[n * 3 + 1] = (int32_t)compact3((uint32_t)(v >> 1));
    coords[n * 3 + 2] = (int32_t)compact3((uint32_t)(v >> 2));
}

// raymarching.cu:287-306: 8 cells -> 1 byte.  One thread per byte, two 16-byte loads.
__global__ void __launch_bounds__(256)
k_packbits(const float4 *__restrict__ grid, uint32_t N, float thresh, uint8_t *__restrict__ bits) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float4 a = __ldg(grid + 2 * (size_t)n);
    const float4 b = __ldg(grid + 2 * (size_t)n + 1);
    unsigned v = 0;
    v |= (a.x > thresh) ? 1u : 0u;
    v |= (a.y > thresh) ? 2u : 0u;
    v |= (a.z > thresh) ? 4u : 0u;
    v |= (a.w > thresh) ? 8u : 0u;
    v |= (b.x > thresh) ? 16u : 0u;
    v |= (b.y > thresh) ? 32u : 0u;
    v |= (b.z > thresh) ? 64u : 0u;
    v |= (b.w > thresh) ? 128u : 0u;
    bits[n] = (uint8_t)v;
}

// The same with the threshold read from device memory, thresh = min(*mean, cap): the density-grid refresh of the training
// engine computes the grid mean on the device and never synchronises with the host (nerf/engine.py).
__global__ void __launch_bounds__(256)
k_packbits_dev(const float4 *__restrict__ grid, uint32_t N, const float *__restrict__ mean, float cap,
               uint8_t *__restrict__ bits) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float thresh = fminf(__ldg(mean), cap);
    const float4 a = __ldg(grid + 2 * (size_t)n);
    const float4 b = __ldg(grid + 2 * (size_t)n + 1);
    unsigned v = 0;
    v |= (a.x > thresh) ? 1u : 0u;
    v |= (a.y > thresh) ? 2u : 0u;
    v |= (a.z > thresh) ? 4u : 0u;
    v |= (a.w > thresh) ? 8u : 0u;
    v |= (b.x > thresh) ? 16u : 0u;
    v |= (b.y > thresh) ? 32u : 0u;
    v |= (b.z > thresh) ? 64u : 0u;
    v |= (b.w > thresh) ? 128u : 0u;
    bits[n] = (uint8_t)v;
}

// ------------------------------------------------------------------------------------------
// warp-cooperative marcher
// ------------------------------------------------------------------------------------------

struct MarchConst {
    const uint8_t *grid;
    float bound, dt_gamma, dt_min, dt_max, rH;
    float Cf, Hf;
    float dt_const;     // the step when dt_gamma == 0 (every step is clamp(0, dt_min, dt_max))
    bool const_step;    // compile-time constant in the <kConstStep = true> kernels
    uint32_t C, H, H3;
};

struct RayGeo {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
};

__device__ __forceinline__ MarchConst make_const(const uint8_t *grid, float bound, float dt_gamma,
                                                 uint32_t max_steps, uint32_t C, uint32_t H, bool const_step = false) {
    MarchConst k;
    k.const_step = const_step;
    k.grid = grid;
    k.bound = bound;
    k.dt_gamma = dt_gamma;
    const float two_sqrt3 = 2 * 1.7320508075688772f;
    k.dt_min = two_sqrt3 / max_steps;                // raymarching.cu:369
    k.dt_max = two_sqrt3 * (1 << (C - 1)) / H;       // raymarching.cu:370
    k.dt_const = clampf(0.0f, k.dt_min, k.dt_max);
    k.rH = 1 / (float)H;
    k.Cf = (float)C;
    k.Hf = (float)H;
    k.C = C;
    k.H = H;
    k.H3 = H * H * H;
    return k;
}

__device__ __forceinline__ RayGeo load_ray(const float *__restrict__ o, const float *__restrict__ d) {
    RayGeo r;
    r.ox = o[0], r.oy = o[1], r.oz = o[2];
    r.dx = d[0], r.dy = d[1], r.dz = d[2];
    r.rdx = 1 / r.dx, r.rdy = 1 / r.dy, r.rdz = 1 / r.dz;
    return r;
}

// dt_gamma == 0 (the LiDAR configurations): t * 0 is +0 for every finite t and NaN otherwise, and fmaxf drops the NaN,
// so the clamp yields the same constant for every t - the 31-step candidate chain of a window shrinks from
// (FMUL, FMNMX, FMNMX, FADD) to one FADD per step, bit-identical.
__device__ __forceinline__ float step_len(const MarchConst &k, float t) {
    if (k.const_step) return k.dt_const;
    return clampf(t * k.dt_gamma, k.dt_min, k.dt_max);
}

__device__ __forceinline__ float unit_sign(float v) { return copysignf(1.0f, v); }

__device__ __forceinline__ int clamp_level(int e, uint32_t C) { return min((int)C - 1, max(0, e)); }

// One candidate position: sample point, step, occupancy, and (if empty) where the skip lands.
// Mirrors raymarching.cu:386-433 term by term (operand order and the double-precision 0.5 kept).
struct Cand {
    float x, y, z, dt, tt;
    bool occ;
};

__device__ __forceinline__ Cand probe(const MarchConst &k, const RayGeo &r, float t) {
    Cand c;
    c.x = clampf(r.ox + t * r.dx, -k.bound, k.bound);
    c.y = clampf(r.oy + t * r.dy, -k.bound, k.bound);
    c.z = clampf(r.oz + t * r.dz, -k.bound, k.bound);
    c.dt = step_len(k, t);

    int e_pos, e_dt;
    frexpf(fmaxf(fabsf(c.x), fmaxf(fabsf(c.y), fabsf(c.z))), &e_pos);   // raymarching.cu:51-60
    const float half_cells = c.dt * k.Hf * 0.5;                            // raymarching.cu:62-69
    frexpf(half_cells, &e_dt);
    const int level = max(clamp_level(e_pos, k.C), clamp_level(e_dt, k.C));

    const float mip_bound = fminf(scalbnf(1.0f, level), k.bound);
    const float mip_rbound = 1 / mip_bound;

    const int nx = clampf(0.5 * (c.x * mip_rbound + 1) * k.H, 0.0f, (float)(k.H - 1));
    const int ny = clampf(0.5 * (c.y * mip_rbound + 1) * k.H, 0.0f, (float)(k.H - 1));
    const int nz = clampf(0.5 * (c.z * mip_rbound + 1) * k.H, 0.0f, (float)(k.H - 1));

    const uint32_t cell = (uint32_t)level * k.H3 + morton_encode(nx, ny, nz);
    c.occ = (__ldg(k.grid + (cell >> 3)) >> (cell & 7u)) & 1u;

    // distance to the far face of this cell along the ray (only meaningful when !occ)
    const float tx = (((nx + 0.5f + 0.5f * unit_sign(r.dx)) * k.rH * 2 - 1) * mip_bound - c.x) * r.rdx;
    const float ty = (((ny + 0.5f + 0.5f * unit_sign(r.dy)) * k.rH * 2 - 1) * mip_bound - c.y) * r.rdy;
    const float tz = (((nz + 0.5f + 0.5f * unit_sign(r.dz)) * k.rH * 2 - 1) * mip_bound - c.z) * r.rdz;
    c.tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    return c;
}

// Marches one ray with the whole warp.  `emit(rank, cand, delta_real)` is called by the lane that
// owns emitted sample number `rank` (0-based along the ray).  Returns the number of samples
// (<= cap).  All lanes must call with identical arguments.
// What the counting pass remembers for the emitting pass: the windows that emitted anything, as (emit mask, t of the
// window's first candidate), window i held by lane i.  The candidate sequence does not depend on the grid, so the
// second pass only has to rebuild t inside those windows - no occupancy probes, no skip resolution, and windows that
// lie entirely in empty space are not visited at all.
struct WindowLog {
    unsigned mask = 0;        // this lane's window: which candidates were emitted
    float base = 0.f;         // this lane's window: t of candidate 0
    uint32_t n = 0;           // windows logged (uniform); > 32 = overflow, the caller re-marches instead
};

template <bool kEmit, class Emit>
__device__ __forceinline__ uint32_t march_warp(const MarchConst &k, const RayGeo &r, float t_start,
                                               float far, uint32_t cap, Emit emit, WindowLog *log = nullptr) {
    if (cap == 0) return 0;
    const unsigned lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;

    float base_t = t_start;   // candidate owned by lane 0 of the current window
    float last_t = t_start;   // t after the previously emitted sample (raymarching.cu:466,506-507)
    uint32_t emitted = 0;
    bool pending = false;     // an empty-space skip is still in flight across windows
    float skip_to = 0.f;

    for (;;) {
        // candidates: lane j holds base_t advanced j times (serial float adds, bit-exact)
        float t = base_t;
#pragma unroll
        for (int i = 0; i < 31; ++i) {
            const float tn = t + step_len(k, t);
            if ((unsigned)i < lane) t = tn;
        }
        const float next_base = __shfl_sync(kFullMask, t + step_len(k, t), 31);

        const unsigned act_mask = __ballot_sync(kFullMask, t < far);
        const int n_act = (act_mask == kFullMask) ? 32 : (__ffs(~act_mask) - 1);
        if (n_act == 0) break;

        int cur = 0;
        if (pending) {
            cur = __popc(__ballot_sync(kFullMask, t < skip_to));
            if (cur >= 32) {  // the whole window is still inside the skipped span
                base_t = next_base;
                continue;
            }
            pending = false;
        }

        Cand c;
        if (t < far) c = probe(k, r, t);
        else {
            c.occ = false;
            c.tt = 0.f;
            c.dt = 0.f;
            c.x = c.y = c.z = 0.f;
        }
        const unsigned occ_mask = __ballot_sync(kFullMask, (t < far) && c.occ);

        // where does an empty candidate jump to?  first j' > j with t_j' >= tt (do/while: at least
        // one step).  Lower bound by shuffle binary search over the sorted per-lane t.
        int pos = 0;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const float tv = __shfl_sync(kFullMask, t, pos + s - 1);
            if (tv < c.tt) pos += s;
        }
        {
            const float t31 = __shfl_sync(kFullMask, t, 31);
            if (pos == 31 && t31 < c.tt) pos = 32;
        }
        const int jump = max((int)lane + 1, pos);

        // resolve the visit chain (uniform across the warp)
        unsigned emit_mask = 0;
        bool finished = false;
        while (cur < n_act) {
            if ((occ_mask >> cur) & 1u) {
                const unsigned run = ~(occ_mask >> cur);
                int len = run ? (__ffs(run) - 1) : (32 - cur);  // consecutive occupied candidates
                const uint32_t room = cap - emitted;
                if ((uint32_t)len >= room) {
                    len = (int)room;
                    finished = true;
                }
                emit_mask |= ((len >= 32) ? kFullMask : ((1u << len) - 1u)) << cur;
                emitted += (uint32_t)len;
                cur += len;
                if (finished) break;
            } else {
                const int nj = __shfl_sync(kFullMask, jump, cur);
                if (nj >= 32) {
                    pending = true;
                    skip_to = __shfl_sync(kFullMask, c.tt, cur);
                }
                cur = nj;
            }
        }

        if (log && emit_mask) {
            if (lane == log->n) {
                log->mask = emit_mask;
                log->base = base_t;
            }
            ++log->n;
        }
        if (kEmit) {
            const float t_after = t + c.dt;
            const unsigned before = emit_mask & lt_mask;
            const int prev_lane = before ? (31 - __clz(before)) : 0;
            const float prev_after = __shfl_sync(kFullMask, t_after, prev_lane);
            if ((emit_mask >> lane) & 1u) {
                const uint32_t rank = emitted - __popc(emit_mask) + __popc(before);
                emit(rank, c, t_after - (before ? prev_after : last_t));
            }
            if (emit_mask) last_t = __shfl_sync(kFullMask, t_after, 31 - __clz(emit_mask));
        }

        if (finished || n_act < 32) break;
        base_t = next_base;
    }
    return emitted;
}

// Second pass from a WindowLog: same samples, same bits as march_warp<true> (t is rebuilt by the same serial adds,
// position / step / delta by the same expressions as probe()).
template <class Emit>
__device__ __forceinline__ void replay_windows(const MarchConst &k, const RayGeo &r, float t_start, const WindowLog &log,
                                               Emit emit) {
    const unsigned lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    float last_t = t_start;
    uint32_t emitted = 0;
    for (uint32_t w = 0; w < log.n; ++w) {
        const unsigned emit_mask = __shfl_sync(kFullMask, log.mask, w);
        float t = __shfl_sync(kFullMask, log.base, w);
#pragma unroll
        for (int i = 0; i < 31; ++i) {
            const float tn = t + step_len(k, t);
            if ((unsigned)i < lane) t = tn;
        }
        Cand c;
        c.x = clampf(r.ox + t * r.dx, -k.bound, k.bound);
        c.y = clampf(r.oy + t * r.dy, -k.bound, k.bound);
        c.z = clampf(r.oz + t * r.dz, -k.bound, k.bound);
        c.dt = step_len(k, t);
        c.tt = 0.f;
        c.occ = true;
        const float t_after = t + c.dt;
        const unsigned before = emit_mask & lt_mask;
        const int prev_lane = before ? (31 - __clz(before)) : 0;
        const float prev_after = __shfl_sync(kFullMask, t_after, prev_lane);
        if ((emit_mask >> lane) & 1u) emit(emitted + __popc(before), c, t_after - (before ? prev_after : last_t));
        last_t = __shfl_sync(kFullMask, t_after, 31 - __clz(emit_mask));
        emitted += __popc(emit_mask);
    }
}

struct NoEmit {
    __device__ __forceinline__ void operator()(uint32_t, const Cand &, float) const {}
};

struct SampleWriter {
    float *xyzs, *dirs, *deltas;   // dirs may be null (extended entry point: the fused field kernels use ray ids)
    float dx, dy, dz;
    int32_t *ray_ids = nullptr;    // optional [M]: index of the ray each sample belongs to
    int32_t ray = 0;
    __device__ __forceinline__ void operator()(uint32_t rank, const Cand &c, float delta_real) const {
        float *p = xyzs + (size_t)rank * 3;
        p[0] = c.x, p[1] = c.y, p[2] = c.z;
        if (dirs) {
            float *q = dirs + (size_t)rank * 3;
            q[0] = dx, q[1] = dy, q[2] = dz;
        }
        if (ray_ids) ray_ids[rank] = ray;
        float2 *d = reinterpret_cast<float2 *>(deltas) + rank;
        *d = make_float2(c.dt, delta_real);
    }
};

// raymarching.cu:332-534 (training march).  One warp per ray.
template <bool kConstStep>
__global__ void __launch_bounds__(kThreads)
k_march_train(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
              const uint8_t *__restrict__ grid, float bound, float dt_gamma, uint32_t max_steps,
              uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float *__restrict__ nears,
              const float *__restrict__ fars, float *__restrict__ xyzs, float *__restrict__ dirs,
              float *__restrict__ deltas, int32_t *__restrict__ rays, int32_t *counter,
              const float *__restrict__ noises, int32_t *__restrict__ ray_ids, bool zero_tail) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const MarchConst k = make_const(grid, bound, dt_gamma, max_steps, C, H, kConstStep);
    const RayGeo r = load_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
    const float far = fars[n];
    float t0 = nears[n];
    t0 += step_len(k, t0) * noises[n];  // raymarching.cu:375

    WindowLog log;
    const uint32_t count = march_warp<false>(k, r, t0, far, max_steps, NoEmit(), &log);

    uint32_t offset = 0, slot = 0;
    if (lane_id() == 0) {
        offset = (uint32_t)atomicAdd(counter, (int)count);
        if (zero_tail) __threadfence();   // the sample reservation is visible before the ray is counted (see below)
        slot = (uint32_t)atomicAdd(counter + 1, 1);
        rays[slot * 3] = (int32_t)n;
        rays[slot * 3 + 1] = (int32_t)offset;
        rays[slot * 3 + 2] = (int32_t)count;
    }
    offset = __shfl_sync(kFullMask, offset, 0);
    if (zero_tail) {
        // Extended entry point: the warp that takes the LAST ray slot sees the final sample total (every other warp
        // reserved its samples before it took its slot) and zeroes the rows between the total and the next 128-row
        // tile boundary - the padding the per-sample kernels process (raymarching.py:235-237 zero-fills the whole
        // buffers on the host before every call instead).  Nobody else writes those rows.
        slot = __shfl_sync(kFullMask, slot, 0);
        if (slot == N - 1) {
            __threadfence();
            const uint32_t total = (uint32_t)max(*(volatile int32_t *)counter, 0);
            const uint32_t hi = min((total + 127u) & ~127u, M);
            for (uint32_t i = min(total, M) + lane_id(); i < hi; i += 32) {
                xyzs[(size_t)i * 3] = xyzs[(size_t)i * 3 + 1] = xyzs[(size_t)i * 3 + 2] = 0.f;
                if (dirs) dirs[(size_t)i * 3] = dirs[(size_t)i * 3 + 1] = dirs[(size_t)i * 3 + 2] = 0.f;
                if (ray_ids) ray_ids[i] = 0;
                reinterpret_cast<float2 *>(deltas)[i] = make_float2(0.f, 0.f);
            }
        }
    }
    if (count == 0 || offset + count > M) return;

    SampleWriter w{xyzs + (size_t)offset * 3, dirs ? dirs + (size_t)offset * 3 : nullptr,
                   deltas + (size_t)offset * 2, r.dx, r.dy, r.dz, ray_ids ? ray_ids + offset : nullptr, (int32_t)n};
    if (log.n <= 32) replay_windows(k, r, t0, log, w);
    else march_warp<true>(k, r, t0, far, count, w);     // more than 32 non-empty windows: march again
}

// raymarching.cu:809-928 (inference march: up to n_step samples for each alive ray).
template <bool kConstStep>
__global__ void __launch_bounds__(kThreads)
k_march_infer(uint32_t n_alive, uint32_t n_step, const int32_t *__restrict__ rays_alive,
              const float *__restrict__ rays_t, const float *__restrict__ rays_o,
              const float *__restrict__ rays_d, float bound, float dt_gamma, uint32_t max_steps,
              uint32_t C, uint32_t H, const uint8_t *__restrict__ grid,
              const float *__restrict__ nears, const float *__restrict__ fars,
              float *__restrict__ xyzs, float *__restrict__ dirs, float *__restrict__ deltas,
              const float *__restrict__ noises) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    const MarchConst k = make_const(grid, bound, dt_gamma, max_steps, C, H, kConstStep);
    const RayGeo r = load_ray(rays_o + (size_t)index * 3, rays_d + (size_t)index * 3);
    float t = rays_t[index];
    const float far = fars[index];
    t += step_len(k, t) * noises[n];  // raymarching.cu:856
    const size_t base = (size_t)n * n_step;
    SampleWriter w{xyzs + base * 3, dirs + base * 3, deltas + base * 2, r.dx, r.dy, r.dz};
    march_warp<true>(k, r, t, far, n_step, w);
}

// ------------------------------------------------------------------------------------------
// compositing
// ------------------------------------------------------------------------------------------

// raymarching.cu:578-655.  Warp per ray; NCH colour channels.
template <int NCH>
__global__ void __launch_bounds__(kThreads)
k_composite_train_fwd(const float *__restrict__ sigmas, const float *__restrict__ rgbs,
                      const float *__restrict__ deltas, const int32_t *__restrict__ rays, uint32_t M,
                      uint32_t N, float T_thresh, float *__restrict__ weights_sum,
                      float *__restrict__ depth, float *__restrict__ image) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const unsigned lane = lane_id();
    const uint32_t index = (uint32_t)rays[n * 3];
    const uint32_t offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t count = (uint32_t)rays[n * 3 + 2];

    float acc_c[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc_c[c] = 0.f;
    float acc_w = 0.f, acc_d = 0.f;

    if (count != 0 && offset + count <= M) {
        float T_in = 1.0f, t_in = 0.f;
        for (uint32_t base = 0; base < count; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < count;
            const size_t s = (size_t)offset + i;
            float alpha = 0.f, d_real = 0.f;
            if (valid) {
                const float2 dl = __ldg(reinterpret_cast<const float2 *>(deltas) + s);
                alpha = 1.0f - __expf(-__ldg(sigmas + s) * dl.x);
                d_real = dl.y;
            }
            const float keep = 1.0f - alpha;
            const float P = warp_scan_mul(keep);                 // inclusive product
            float Pex = __shfl_up_sync(kFullMask, P, 1);
            if (lane == 0) Pex = 1.0f;
            const float T_before = T_in * Pex;
            const float T_after = T_before * keep;
            const float t_here = t_in + warp_scan_add(d_real);   // t at the END of this interval

            // the sample that drives T below the threshold still contributes (raymarching.cu:619-634)
            const unsigned stop = __ballot_sync(kFullMask, valid && (T_after < T_thresh));
            const bool use = valid && (stop == 0 || lane < (unsigned)__ffs(stop));
            if (use) {
                const float w = alpha * T_before;
#pragma unroll
                for (int c = 0; c < NCH; ++c) acc_c[c] += w * __ldg(rgbs + s * NCH + c);
                acc_d += w * t_here;
                acc_w += w;
            }
            if (stop) break;
            T_in = __shfl_sync(kFullMask, T_after, 31);
            t_in = __shfl_sync(kFullMask, t_here, 31);
        }
        acc_w = warp_sum(acc_w);
        acc_d = warp_sum(acc_d);
#pragma unroll
        for (int c = 0; c < NCH; ++c) acc_c[c] = warp_sum(acc_c[c]);
    }
    if (lane == 0) {
        weights_sum[index] = acc_w;
        depth[index] = acc_d;
#pragma unroll
        for (int c = 0; c < NCH; ++c) image[(size_t)index * NCH + c] = acc_c[c];
    }
}

// raymarching.cu:691-772 (+ optional depth term).  Samples after the early stop keep
// the caller's zero fill.
template <int NCH, bool kDepthGrad>
__global__ void __launch_bounds__(kThreads)
k_composite_train_bwd(const float *__restrict__ g_ws, const float *__restrict__ g_depth,
                      const float *__restrict__ g_img, const float *__restrict__ sigmas,
                      const float *__restrict__ rgbs, const float *__restrict__ deltas,
                      const int32_t *__restrict__ rays, const float *__restrict__ weights_sum,
                      const float *__restrict__ depth, const float *__restrict__ image, uint32_t M,
                      uint32_t N, float T_thresh, float *__restrict__ grad_sigmas,
                      float *__restrict__ grad_rgbs) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const unsigned lane = lane_id();
    const uint32_t index = (uint32_t)rays[n * 3];
    const uint32_t offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t count = (uint32_t)rays[n * 3 + 2];
    if (count == 0 || offset + count > M) return;

    float gi[NCH], c_final[NCH], c_in[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        gi[c] = g_img[(size_t)index * NCH + c];
        c_final[c] = image[(size_t)index * NCH + c];
        c_in[c] = 0.f;
    }
    const float gw = g_ws[index];
    const float ws_final = weights_sum[index];
    const float gd = kDepthGrad ? g_depth[index] : 0.f;
    const float d_final = kDepthGrad ? depth[index] : 0.f;
    float d_in = 0.f;

    float T_in = 1.0f, t_in = 0.f;
    for (uint32_t base = 0; base < count; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < count;
        const size_t s = (size_t)offset + i;
        float alpha = 0.f, d_rgb = 0.f, d_real = 0.f;
        float col[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) col[c] = 0.f;
        if (valid) {
            const float2 dl = __ldg(reinterpret_cast<const float2 *>(deltas) + s);
            d_rgb = dl.x;
            d_real = dl.y;
            alpha = 1.0f - __expf(-__ldg(sigmas + s) * d_rgb);
#pragma unroll
            for (int c = 0; c < NCH; ++c) col[c] = __ldg(rgbs + s * NCH + c);
        }
        const float keep = 1.0f - alpha;
        const float P = warp_scan_mul(keep);
        float Pex = __shfl_up_sync(kFullMask, P, 1);
        if (lane == 0) Pex = 1.0f;
        const float T_before = T_in * Pex;
        const float T_after = T_before * keep;
        const float w = alpha * T_before;

        const unsigned stop = __ballot_sync(kFullMask, valid && (T_after < T_thresh));
        const bool use = valid && (stop == 0 || lane < (unsigned)__ffs(stop));

        // inclusive running colour (and depth) sums up to and including this sample
        float acc = 0.f;
        float c_run[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            c_run[c] = c_in[c] + warp_scan_add(w * col[c]);
            acc += gi[c] * (T_after * col[c] - (c_final[c] - c_run[c]));
        }
        acc += gw * (1 - ws_final);
        float t_here = 0.f, d_run = 0.f;
        if (kDepthGrad) {
            t_here = t_in + warp_scan_add(d_real);
            d_run = d_in + warp_scan_add(w * t_here);
            acc += gd * (T_after * t_here - (d_final - d_run));
        }
        if (use) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) grad_rgbs[s * NCH + c] = gi[c] * w;
            grad_sigmas[s] = d_rgb * acc;
        }
        if (stop) break;
        T_in = __shfl_sync(kFullMask, T_after, 31);
#pragma unroll
        for (int c = 0; c < NCH; ++c) c_in[c] = __shfl_sync(kFullMask, c_run[c], 31);
        if (kDepthGrad) {
            t_in = __shfl_sync(kFullMask, t_here, 31);
            d_in = __shfl_sync(kFullMask, d_run, 31);
        }
    }
}

// Fused LiDAR compositing step of the training engine: composite forward (raymarching.cu:578-655) -> LiDAR loss and
// its per-ray gradients (nerf/utils.py:726-734: L1 depth, MSE ray-drop, MSE intensity) -> composite backward
// (raymarching.cu:691-772 + the depth term).  The warp that owns a ray keeps its forward results in registers and
// walks the ray's samples a second time for the backward pass (the re-read hits L1/L2), so the per-ray outputs and
// their gradients never make a round trip through global memory, and every sample of the ray gets a gradient
// written - zero after the early stop - which removes the zero-fill of the two gradient buffers from the step.
// Same arithmetic, in the same order, as k_composite_train_fwd + k_lidar_loss + k_composite_train_bwd<2, true>.
//
// kMode 0: all of it in one pass (per-ray losses only).  A loss that couples NEIGHBOURING rays (the patch depth-gradient
// term, nerf/utils.py:748-876) needs every ray's depth before any gradient exists, so the step is then split at the loss:
// kMode 1 = forward only (weights_sum / depth / image / t0 out), then lnb_lidar_loss_ex, then kMode 2 = backward with the
// per-ray gradients read from ext_g_ws / ext_g_depth / ext_g_image (forward results re-read from the output arrays),
// still producing the zero-filled sample gradients and the compact live-row list.
template <int kMode>
__global__ void __launch_bounds__(kThreads)
k_lidar_composite_step(const float *__restrict__ sigmas, const float *__restrict__ rgbs,
                       const float *__restrict__ deltas, const int32_t *__restrict__ rays,
                       const float *__restrict__ gt, const float *__restrict__ nears,
                       const float *__restrict__ noises, float dt_gamma, float dt_min, float dt_max,
                       const int32_t *__restrict__ counter, uint32_t M, uint32_t N, float T_thresh, float a_d,
                       float a_r, float a_i, float loss_scale, float *__restrict__ weights_sum,
                       float *__restrict__ depth, float *__restrict__ image, float *__restrict__ t0_out,
                       float *__restrict__ grad_sigmas, float *__restrict__ grad_rgbs,
                       float *__restrict__ loss_out, int32_t *__restrict__ live_idx, int32_t *__restrict__ n_live,
                       const float *__restrict__ ext_g_ws, const float *__restrict__ ext_g_depth,
                       const float *__restrict__ ext_g_image) {
    constexpr int NCH = 2;
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= N) return;
    const unsigned lane = lane_id();
    if (kMode != 1 && n == 0 && counter) {
        // rows between the produced count and the next 128-row tile boundary are processed by the per-sample
        // kernels but belong to no ray: their gradients are zero
        const uint32_t cnt = (uint32_t)max(*counter, 0);
        const uint32_t lo = min(cnt, M), hi = min((cnt + 127u) & ~127u, M);
        for (uint32_t s = lo + lane; s < hi; s += 32) {
            grad_sigmas[s] = 0.f;
            grad_rgbs[(size_t)s * NCH] = 0.f;
            grad_rgbs[(size_t)s * NCH + 1] = 0.f;
        }
    }
    const uint32_t index = (uint32_t)rays[n * 3];
    const uint32_t offset = (uint32_t)rays[n * 3 + 1];
    const uint32_t count = (uint32_t)rays[n * 3 + 2];
    const bool marched = count != 0 && offset + count <= M;

    // ---------------- forward ----------------
    float c_final[NCH] = {0.f, 0.f};
    float ws_final = 0.f, d_final = 0.f;
    if (kMode == 2) {
        ws_final = weights_sum[index];
        d_final = depth[index];
        const float2 im = reinterpret_cast<const float2 *>(image)[index];
        c_final[0] = im.x, c_final[1] = im.y;
    }
    if (kMode != 2 && marched) {
        float T_in = 1.0f, t_in = 0.f;
        for (uint32_t base = 0; base < count; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < count;
            const size_t s = (size_t)offset + i;
            float alpha = 0.f, d_real = 0.f;
            if (valid) {
                const float2 dl = __ldg(reinterpret_cast<const float2 *>(deltas) + s);
                alpha = 1.0f - __expf(-__ldg(sigmas + s) * dl.x);
                d_real = dl.y;
            }
            const float keep = 1.0f - alpha;
            const float P = warp_scan_mul(keep);
            float Pex = __shfl_up_sync(kFullMask, P, 1);
            if (lane == 0) Pex = 1.0f;
            const float T_before = T_in * Pex;
            const float T_after = T_before * keep;
            const float t_here = t_in + warp_scan_add(d_real);
            const unsigned stop = __ballot_sync(kFullMask, valid && (T_after < T_thresh));
            const bool use = valid && (stop == 0 || lane < (unsigned)__ffs(stop));
            if (use) {
                const float w = alpha * T_before;
                const float2 col = __ldg(reinterpret_cast<const float2 *>(rgbs) + s);
                c_final[0] += w * col.x;
                c_final[1] += w * col.y;
                d_final += w * t_here;
                ws_final += w;
            }
            if (stop) break;
            T_in = __shfl_sync(kFullMask, T_after, 31);
            t_in = __shfl_sync(kFullMask, t_here, 31);
        }
        ws_final = warp_sum(ws_final);
        d_final = warp_sum(d_final);
        c_final[0] = warp_sum(c_final[0]);
        c_final[1] = warp_sum(c_final[1]);
    }

    // ---------------- loss and per-ray gradients (every lane computes the same values) ----------------
    const float near = __ldg(nears + index);
    const float start = fmaf(clampf(near * dt_gamma, dt_min, dt_max), __ldg(noises + index), near);   // raymarching.cu:375
    const float m = __ldg(gt + (size_t)index * 3);
    const float gti = __ldg(gt + (size_t)index * 3 + 1) * m, gtd = __ldg(gt + (size_t)index * 3 + 2) * m;
    const float D = d_final + start * ws_final;
    const float e_d = D * m - gtd;
    const float e_r = c_final[0] - m;
    const float e_i = c_final[1] * m - gti;
    const float inv_n = 1.f / (float)N;
    const float sc = loss_scale * inv_n;
    float gd = a_d * m * (e_d > 0.f ? 1.f : (e_d < 0.f ? -1.f : 0.f)) * sc;
    float gw = gd * start;
    float gi[NCH] = {2.f * a_r * e_r * sc, 2.f * a_i * e_i * m * sc};
    if (kMode == 2) {
        gd = ext_g_depth[index];
        gw = ext_g_ws[index];
        const float2 g2 = reinterpret_cast<const float2 *>(ext_g_image)[index];
        gi[0] = g2.x, gi[1] = g2.y;
    }
    if (kMode != 2 && lane == 0) {
        weights_sum[index] = ws_final;
        depth[index] = d_final;
        reinterpret_cast<float2 *>(image)[index] = make_float2(c_final[0], c_final[1]);
        if (t0_out) t0_out[index] = start;
        if (kMode == 0) {
            const float l = (a_d * fabsf(e_d) + a_r * e_r * e_r + a_i * e_i * e_i) * inv_n;
            if (l != 0.f) atomicAdd(loss_out, l);
        }
    }
    if (kMode == 1 || !marched) return;

    // ---------------- backward ----------------
    float c_in[NCH] = {0.f, 0.f};
    float d_in = 0.f, T_in = 1.0f, t_in = 0.f;
    uint32_t done = count;     // first sample index that received no gradient (early stop)
    uint32_t n_used = count;   // samples up to and including the one that crossed the threshold
    for (uint32_t base = 0; base < count; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < count;
        const size_t s = (size_t)offset + i;
        float alpha = 0.f, d_rgb = 0.f, d_real = 0.f;
        float2 col = make_float2(0.f, 0.f);
        if (valid) {
            const float2 dl = __ldg(reinterpret_cast<const float2 *>(deltas) + s);
            d_rgb = dl.x;
            d_real = dl.y;
            alpha = 1.0f - __expf(-__ldg(sigmas + s) * d_rgb);
            col = __ldg(reinterpret_cast<const float2 *>(rgbs) + s);
        }
        const float keep = 1.0f - alpha;
        const float P = warp_scan_mul(keep);
        float Pex = __shfl_up_sync(kFullMask, P, 1);
        if (lane == 0) Pex = 1.0f;
        const float T_before = T_in * Pex;
        const float T_after = T_before * keep;
        const float w = alpha * T_before;
        const unsigned stop = __ballot_sync(kFullMask, valid && (T_after < T_thresh));
        const bool use = valid && (stop == 0 || lane < (unsigned)__ffs(stop));

        float acc = 0.f;
        float c_run[NCH];
        c_run[0] = c_in[0] + warp_scan_add(w * col.x);
        acc += gi[0] * (T_after * col.x - (c_final[0] - c_run[0]));
        c_run[1] = c_in[1] + warp_scan_add(w * col.y);
        acc += gi[1] * (T_after * col.y - (c_final[1] - c_run[1]));
        acc += gw * (1 - ws_final);
        const float t_here = t_in + warp_scan_add(d_real);
        const float d_run = d_in + warp_scan_add(w * t_here);
        acc += gd * (T_after * t_here - (d_final - d_run));
        if (valid) {
            reinterpret_cast<float2 *>(grad_rgbs)[s] = use ? make_float2(gi[0] * w, gi[1] * w) : make_float2(0.f, 0.f);
            grad_sigmas[s] = use ? d_rgb * acc : 0.f;
        }
        if (stop) {
            done = base + 32;
            n_used = base + (uint32_t)__ffs(stop);
            break;
        }
        T_in = __shfl_sync(kFullMask, T_after, 31);
        c_in[0] = __shfl_sync(kFullMask, c_run[0], 31);
        c_in[1] = __shfl_sync(kFullMask, c_run[1], 31);
        t_in = __shfl_sync(kFullMask, t_here, 31);
        d_in = __shfl_sync(kFullMask, d_run, 31);
    }
    for (uint32_t i = done + lane; i < count; i += 32) {   // samples behind the early stop (raymarching.py:338-339)
        const size_t s = (size_t)offset + i;
        reinterpret_cast<float2 *>(grad_rgbs)[s] = make_float2(0.f, 0.f);
        grad_sigmas[s] = 0.f;
    }
    if (live_idx) {
        // rows that can carry a gradient (the ray's samples up to the early stop), appended to a compact list in
        // arrival order: the backward kernels walk this list instead of all marched rows
        uint32_t at = 0;
        if (lane == 0) at = (uint32_t)atomicAdd(n_live, (int)n_used);
        at = __shfl_sync(kFullMask, at, 0);
        for (uint32_t i = lane; i < n_used; i += 32) live_idx[at + i] = (int32_t)(offset + i);
    }
}

// raymarching.cu:967-1053 (inference compositing).  n_step is small (<= 8 in the upstream
// driver), so one thread per ray is the right granularity; T = 1 - weight_sum (not a product).
__global__ void __launch_bounds__(kThreads)
k_composite_infer(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *__restrict__ rays_alive,
                  float *__restrict__ rays_t, const float *__restrict__ sigmas,
                  const float *__restrict__ rgbs, const float *__restrict__ deltas,
                  float *__restrict__ weights_sum, float *__restrict__ depth,
                  float *__restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    const size_t base = (size_t)n * n_step;
    float t = rays_t[index];
    float ws = weights_sum[index], d = depth[index];
    float r = image[(size_t)index * 3], g = image[(size_t)index * 3 + 1], b = image[(size_t)index * 3 + 2];
    uint32_t step = 0;
    for (; step < n_step; ++step) {
        const float2 dl = *(reinterpret_cast<const float2 *>(deltas) + base + step);
        if (dl.x == 0) break;  // zero-filled padding marks the end of this ray's samples
        const float alpha = 1.0f - __expf(-sigmas[base + step] * dl.x);
        const float T = 1 - ws;
        const float w = alpha * T;
        ws += w;
        t += dl.y;
        d += w * t;
        r += w * rgbs[(base + step) * 3];
        g += w * rgbs[(base + step) * 3 + 1];
        b += w * rgbs[(base + step) * 3 + 2];
        if (T < T_thresh) break;
    }
    if (step < n_step) rays_alive[n] = -1;
    else rays_t[index] = t;
    weights_sum[index] = ws;
    depth[index] = d;
    image[(size_t)index * 3] = r;
    image[(size_t)index * 3 + 1] = g;
    image[(size_t)index * 3 + 2] = b;
}

inline unsigned blocks_for_threads(uint64_t threads, unsigned per_block) {
    return (unsigned)((threads + per_block - 1) / per_block);
}

}  // namespace
}  // namespace lnb

using namespace lnb;

#define LNB_REQUIRE(cond) \
    do {                  \
        if (!(cond)) return LNB_ERR_INVALID_ARGUMENT; \
    } while (0)

extern "C" {

uint64_t lnb_launch_count(void) { return g_launch_count; }
int lnb_version(void) { return 100; }
const char *lnb_arch(void) { return "sm_100a"; }

const char *lnb_strerror(int status) {
    switch (status) {
        case LNB_OK: return "ok";
        case LNB_ERR_INVALID_ARGUMENT: return "lidarnerf_b200: invalid argument";
        case LNB_ERR_UNSUPPORTED: return "lidarnerf_b200: unsupported shape/dtype/configuration";
        case LNB_ERR_NO_DEVICE: return "lidarnerf_b200: no CUDA device";
        case LNB_ERR_WORKSPACE: return "lidarnerf_b200: workspace too small";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "lidarnerf_b200: unknown status";
}

int lnb_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N,
                           float min_near, float *nears, float *fars, lnb_stream_t stream) {
    LNB_REQUIRE(rays_o && rays_d && aabb && nears && fars);
    if (N == 0) return LNB_OK;
    k_near_far<<<blocks_for_threads(N, kThreads), kThreads, 0, as_stream(stream)>>>(
        rays_o, rays_d, aabb, N, min_near, nears, fars);
    count_launch();
    return launch_status();
}

int lnb_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N,
                     float *coords, lnb_stream_t stream) {
    LNB_REQUIRE(rays_o && rays_d && coords);
    if (N == 0) return LNB_OK;
    k_sph_from_ray<<<blocks_for_threads(N, kThreads), kThreads, 0, as_stream(stream)>>>(
        rays_o, rays_d, radius, N, coords);
    count_launch();
    return launch_status();
}

int lnb_morton3D(const int32_t *coords, uint32_t N, int32_t *indices, lnb_stream_t stream) {
    LNB_REQUIRE(coords && indices);
    if (N == 0) return LNB_OK;
    k_morton<<<blocks_for_threads(N, kThreads), kThreads, 0, as_stream(stream)>>>(coords, N, indices);
    count_launch();
    return launch_status();
}

int lnb_morton3D_invert(const int32_t *indices, uint32_t N, int32_t *coords, lnb_stream_t stream) {
    LNB_REQUIRE(coords && indices);
    if (N == 0) return LNB_OK;
    k_morton_invert<<<blocks_for_threads(N, kThreads), kThreads, 0, as_stream(stream)>>>(indices, N,
                                                                                        coords);
    count_launch();
    return launch_status();
}

int lnb_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield,
                 lnb_stream_t stream) {
    LNB_REQUIRE(grid && bitfield);
    LNB_REQUIRE((reinterpret_cast<uintptr_t>(grid) & 15u) == 0);
    if (N == 0) return LNB_OK;
    k_packbits<<<blocks_for_threads(N, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4 *>(grid), N, density_thresh, bitfield);
    count_launch();
    return launch_status();
}

int lnb_packbits_dev(const float *grid, uint32_t N, const float *mean_density_dev, float density_thresh_cap,
                     uint8_t *bitfield, lnb_stream_t stream) {
    LNB_REQUIRE(grid && bitfield && mean_density_dev);
    LNB_REQUIRE((reinterpret_cast<uintptr_t>(grid) & 15u) == 0);
    if (N == 0) return LNB_OK;
    k_packbits_dev<<<blocks_for_threads(N, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4 *>(grid), N, mean_density_dev, density_thresh_cap, bitfield);
    count_launch();
    return launch_status();
}

static int march_rays_train_impl(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                                 float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                 uint32_t M, const float *nears, const float *fars, float *xyzs, float *dirs,
                                 float *deltas, int32_t *rays, int32_t *counter, const float *noises,
                                 int32_t *ray_ids, int zero_tail, lnb_stream_t stream) {
    LNB_REQUIRE(rays_o && rays_d && grid && nears && fars && rays && counter && noises);
    LNB_REQUIRE(M == 0 || (xyzs && deltas && (dirs || ray_ids)));
    LNB_REQUIRE(C >= 1 && C <= 8 && H >= 1 && H <= 1024 && max_steps >= 1);
    if (N == 0) return LNB_OK;
    auto kern = (dt_gamma == 0.0f) ? k_march_train<true> : k_march_train<false>;
    kern<<<blocks_for_threads((uint64_t)N * 32, kThreads), kThreads, 0, as_stream(stream)>>>(
        rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas,
        rays, counter, noises, ray_ids, zero_tail != 0);
    count_launch();
    return launch_status();
}

int lnb_march_rays_train_ex(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                            float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                            uint32_t M, const float *nears, const float *fars, float *xyzs, float *dirs,
                            float *deltas, int32_t *rays, int32_t *counter, const float *noises,
                            int32_t *ray_ids, lnb_stream_t stream) {
    return march_rays_train_impl(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs,
                                 deltas, rays, counter, noises, ray_ids, 1, stream);
}

int lnb_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                         float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                         uint32_t M, const float *nears, const float *fars, float *xyzs, float *dirs,
                         float *deltas, int32_t *rays, int32_t *counter, const float *noises,
                         lnb_stream_t stream) {
    LNB_REQUIRE(M == 0 || dirs);
    return march_rays_train_impl(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs,
                                 deltas, rays, counter, noises, nullptr, 0, stream);
}

int lnb_composite_rays_train_forward_ex(const float *sigmas, const float *rgbs, const float *deltas,
                                        const int32_t *rays, uint32_t M, uint32_t N, float T_thresh,
                                        uint32_t channels, float *weights_sum, float *depth,
                                        float *image, lnb_stream_t stream) {
    LNB_REQUIRE(rays && weights_sum && depth && image);
    LNB_REQUIRE(M == 0 || (sigmas && rgbs && deltas));
    if (N == 0) return LNB_OK;
    const unsigned blocks = blocks_for_threads((uint64_t)N * 32, kThreads);
    cudaStream_t st = as_stream(stream);
#define LNB_FWD(CH)                                                                              \
    k_composite_train_fwd<CH><<<blocks, kThreads, 0, st>>>(sigmas, rgbs, deltas, rays, M, N,     \
                                                           T_thresh, weights_sum, depth, image)
    switch (channels) {
        case 1: LNB_FWD(1); break;
        case 2: LNB_FWD(2); break;
        case 3: LNB_FWD(3); break;
        case 4: LNB_FWD(4); break;
        default: return LNB_ERR_UNSUPPORTED;
    }
#undef LNB_FWD
    count_launch();
    return launch_status();
}

int lnb_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas,
                                     const int32_t *rays, uint32_t M, uint32_t N, float T_thresh,
                                     float *weights_sum, float *depth, float *image,
                                     lnb_stream_t stream) {
    return lnb_composite_rays_train_forward_ex(sigmas, rgbs, deltas, rays, M, N, T_thresh, 3,
                                               weights_sum, depth, image, stream);
}

int lnb_composite_rays_train_backward_ex(const float *grad_weights_sum, const float *grad_depth,
                                         const float *grad_image, const float *sigmas,
                                         const float *rgbs, const float *deltas, const int32_t *rays,
                                         const float *weights_sum, const float *depth,
                                         const float *image, uint32_t M, uint32_t N, float T_thresh,
                                         uint32_t channels, float *grad_sigmas, float *grad_rgbs,
                                         lnb_stream_t stream) {
    LNB_REQUIRE(grad_weights_sum && grad_image && rays && weights_sum && image);
    LNB_REQUIRE(M == 0 || (sigmas && rgbs && deltas && grad_sigmas && grad_rgbs));
    LNB_REQUIRE((grad_depth == nullptr) == (depth == nullptr) || grad_depth == nullptr);
    if (N == 0 || M == 0) return LNB_OK;
    const unsigned blocks = blocks_for_threads((uint64_t)N * 32, kThreads);
    cudaStream_t st = as_stream(stream);
    const bool dg = grad_depth != nullptr;
    LNB_REQUIRE(!dg || depth != nullptr);
#define LNB_BWD(CH)                                                                                  \
    do {                                                                                             \
        if (dg)                                                                                      \
            k_composite_train_bwd<CH, true><<<blocks, kThreads, 0, st>>>(                            \
                grad_weights_sum, grad_depth, grad_image, sigmas, rgbs, deltas, rays, weights_sum,   \
                depth, image, M, N, T_thresh, grad_sigmas, grad_rgbs);                               \
        else                                                                                         \
            k_composite_train_bwd<CH, false><<<blocks, kThreads, 0, st>>>(                           \
                grad_weights_sum, nullptr, grad_image, sigmas, rgbs, deltas, rays, weights_sum,      \
                nullptr, image, M, N, T_thresh, grad_sigmas, grad_rgbs);                             \
    } while (0)
    switch (channels) {
        case 1: LNB_BWD(1); break;
        case 2: LNB_BWD(2); break;
        case 3: LNB_BWD(3); break;
        case 4: LNB_BWD(4); break;
        default: return LNB_ERR_UNSUPPORTED;
    }
#undef LNB_BWD
    count_launch();
    return launch_status();
}

int lnb_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_image,
                                      const float *sigmas, const float *rgbs, const float *deltas,
                                      const int32_t *rays, const float *weights_sum,
                                      const float *image, uint32_t M, uint32_t N, float T_thresh,
                                      float *grad_sigmas, float *grad_rgbs, lnb_stream_t stream) {
    return lnb_composite_rays_train_backward_ex(grad_weights_sum, nullptr, grad_image, sigmas, rgbs,
                                                deltas, rays, weights_sum, nullptr, image, M, N,
                                                T_thresh, 3, grad_sigmas, grad_rgbs, stream);
}

int lnb_lidar_composite_step(const float *sigmas, const float *rgbs, const float *deltas, const int32_t *rays,
                              const float *gt, const float *nears, const float *noises, float dt_gamma,
                              uint32_t max_steps, uint32_t C, uint32_t H, const int32_t *counter, uint32_t M,
                              uint32_t N, float T_thresh, float alpha_d, float alpha_r, float alpha_i,
                              float loss_scale, float *weights_sum, float *depth, float *image, float *t0,
                              float *grad_sigmas, float *grad_rgbs, float *loss_out, int32_t *live_idx,
                              int32_t *n_live, lnb_stream_t stream) {
    LNB_REQUIRE((live_idx == nullptr) == (n_live == nullptr));
    LNB_REQUIRE(rays && gt && nears && noises && weights_sum && depth && image && loss_out);
    LNB_REQUIRE(M == 0 || (sigmas && rgbs && deltas && grad_sigmas && grad_rgbs));
    LNB_REQUIRE(C >= 1 && C <= 8 && H >= 1 && max_steps >= 1);
    if (N == 0) return LNB_OK;
    const float two_sqrt3 = 2 * 1.7320508075688772f;       // same expressions as make_const()
    const float dt_min = two_sqrt3 / max_steps;
    const float dt_max = two_sqrt3 * (1 << (C - 1)) / H;
    k_lidar_composite_step<0><<<blocks_for_threads((uint64_t)N * 32, kThreads), kThreads, 0, as_stream(stream)>>>(
        sigmas, rgbs, deltas, rays, gt, nears, noises, dt_gamma, dt_min, dt_max, counter, M, N, T_thresh, alpha_d,
        alpha_r, alpha_i, loss_scale, weights_sum, depth, image, t0, grad_sigmas, grad_rgbs, loss_out, live_idx,
        n_live, nullptr, nullptr, nullptr);
    count_launch();
    return launch_status();
}

int lnb_lidar_composite_forward(const float *sigmas, const float *rgbs, const float *deltas, const int32_t *rays,
                                const float *gt, const float *nears, const float *noises, float dt_gamma,
                                uint32_t max_steps, uint32_t C, uint32_t H, uint32_t M, uint32_t N, float T_thresh,
                                float *weights_sum, float *depth, float *image, float *t0, lnb_stream_t stream) {
    LNB_REQUIRE(rays && gt && nears && noises && weights_sum && depth && image);
    LNB_REQUIRE(M == 0 || (sigmas && rgbs && deltas));
    LNB_REQUIRE(C >= 1 && C <= 8 && H >= 1 && max_steps >= 1);
    if (N == 0) return LNB_OK;
    const float two_sqrt3 = 2 * 1.7320508075688772f;
    const float dt_min = two_sqrt3 / max_steps;
    const float dt_max = two_sqrt3 * (1 << (C - 1)) / H;
    k_lidar_composite_step<1><<<blocks_for_threads((uint64_t)N * 32, kThreads), kThreads, 0, as_stream(stream)>>>(
        sigmas, rgbs, deltas, rays, gt, nears, noises, dt_gamma, dt_min, dt_max, nullptr, M, N, T_thresh, 0.f, 0.f, 0.f,
        0.f, weights_sum, depth, image, t0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    count_launch();
    return launch_status();
}

int lnb_lidar_composite_backward(const float *g_weights_sum, const float *g_depth, const float *g_image,
                                 const float *sigmas, const float *rgbs, const float *deltas, const int32_t *rays,
                                 const float *gt, const float *nears, const float *noises, float dt_gamma,
                                 uint32_t max_steps, uint32_t C, uint32_t H, const int32_t *counter, uint32_t M,
                                 uint32_t N, float T_thresh, const float *weights_sum, const float *depth,
                                 const float *image, float *grad_sigmas, float *grad_rgbs, int32_t *live_idx,
                                 int32_t *n_live, lnb_stream_t stream) {
    LNB_REQUIRE((live_idx == nullptr) == (n_live == nullptr));
    LNB_REQUIRE(g_weights_sum && g_depth && g_image && rays && gt && nears && noises && weights_sum && depth && image);
    LNB_REQUIRE(M == 0 || (sigmas && rgbs && deltas && grad_sigmas && grad_rgbs));
    LNB_REQUIRE(C >= 1 && C <= 8 && H >= 1 && max_steps >= 1);
    if (N == 0) return LNB_OK;
    const float two_sqrt3 = 2 * 1.7320508075688772f;
    const float dt_min = two_sqrt3 / max_steps;
    const float dt_max = two_sqrt3 * (1 << (C - 1)) / H;
    k_lidar_composite_step<2><<<blocks_for_threads((uint64_t)N * 32, kThreads), kThreads, 0, as_stream(stream)>>>(
        sigmas, rgbs, deltas, rays, gt, nears, noises, dt_gamma, dt_min, dt_max, counter, M, N, T_thresh, 0.f, 0.f, 0.f,
        0.f, const_cast<float *>(weights_sum), const_cast<float *>(depth), const_cast<float *>(image), nullptr,
        grad_sigmas, grad_rgbs, nullptr, live_idx, n_live, g_weights_sum, g_depth, g_image);
    count_launch();
    return launch_status();
}

int lnb_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t,
                   const float *rays_o, const float *rays_d, float bound, float dt_gamma,
                   uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *grid,
                   const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                   const float *noises, lnb_stream_t stream) {
    LNB_REQUIRE(rays_alive && rays_t && rays_o && rays_d && grid && nears && fars && noises);
    LNB_REQUIRE(xyzs && dirs && deltas);
    LNB_REQUIRE(C >= 1 && C <= 8 && H >= 1 && H <= 1024 && max_steps >= 1);
    if (n_alive == 0 || n_step == 0) return LNB_OK;
    auto kern = (dt_gamma == 0.0f) ? k_march_infer<true> : k_march_infer<false>;
    kern<<<blocks_for_threads((uint64_t)n_alive * 32, kThreads), kThreads, 0, as_stream(stream)>>>(
        n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars,
        xyzs, dirs, deltas, noises);
    count_launch();
    return launch_status();
}

int lnb_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive,
                       float *rays_t, const float *sigmas, const float *rgbs, const float *deltas,
                       float *weights_sum, float *depth, float *image, lnb_stream_t stream) {
    LNB_REQUIRE(rays_alive && rays_t && sigmas && rgbs && deltas && weights_sum && depth && image);
    if (n_alive == 0) return LNB_OK;
    k_composite_infer<<<blocks_for_threads(n_alive, kThreads), kThreads, 0, as_stream(stream)>>>(
        n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image);
    count_launch();
    return launch_status();
}

}  // extern "C"
