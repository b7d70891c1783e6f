"""Seeded input generators shared by the golden-fixture scripts and the parity tests (numpy only)."""
import numpy as np


def unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def clumpy_bitfield(seed, cascade=1, H=128, fill=0.3):
    """A density bitfield with blob structure in Morton order: [cascade * H^3 / 8] uint8."""
    rng = np.random.default_rng(seed)
    n = cascade * H ** 3
    # runs of occupied/empty cells (Morton order keeps runs spatially compact)
    bits = np.zeros(n, dtype=bool)
    pos = 0
    while pos < n:
        run = int(rng.integers(1, 4096))
        if rng.random() < fill:
            bits[pos:pos + run] = True
        pos += run
    return np.packbits(bits.reshape(-1, 8), axis=1, bitorder="little").reshape(-1)


def march_case(seed, N=512, cascade=1, bound=1.0, H=128, fill=0.3, lidar=False, perturb=True):
    rng = np.random.default_rng(seed)
    if lidar:
        rays_o = np.tile(rng.uniform(-0.05, 0.05, size=(1, 3)), (N, 1)).astype(np.float32)
    else:
        rays_o = rng.uniform(-0.6 * bound, 0.6 * bound, size=(N, 3)).astype(np.float32)
    rays_d = unit(rng.normal(size=(N, 3))).astype(np.float32)
    # a few axis-aligned directions (zero components -> infinite reciprocals in the kernels)
    k = min(6, N)
    axes = np.eye(3, dtype=np.float32)
    rays_d[:k] = np.concatenate([axes, -axes])[:k]
    noises = rng.uniform(0, 1, size=N).astype(np.float32) if perturb else np.zeros(N, np.float32)
    bitfield = clumpy_bitfield(seed + 1, cascade, H, fill)
    return dict(rays_o=rays_o, rays_d=rays_d, noises=noises, bitfield=bitfield, cascade=cascade, bound=float(bound), H=H)


def composite_case(seed, N=300, max_count=200, ch=3, opaque_frac=0.3):
    """Ragged rays: counts include 0; some rays are dense enough to hit the early-termination threshold."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, max_count, size=N).astype(np.int32)
    counts[: min(5, N)] = [0, 1, 31, 32, 33][: min(5, N)]
    order = rng.permutation(N)  # arrival order != ray id
    offsets = np.zeros(N, np.int32)
    offsets[order] = np.concatenate([[0], np.cumsum(counts[order])[:-1]])
    M = int(counts.sum()) + 7  # a few padding rows past the end
    rays = np.stack([np.arange(N, dtype=np.int32), offsets, counts], axis=1)[order]
    sigmas = rng.gamma(1.0, 2.0, size=M).astype(np.float32)
    opaque = rng.random(N) < opaque_frac
    for n in np.nonzero(opaque)[0]:
        sigmas[offsets[n]:offsets[n] + counts[n]] *= 60.0
    deltas = np.stack([rng.uniform(0.002, 0.02, size=M), rng.uniform(0.002, 0.05, size=M)], axis=1).astype(np.float32)
    rgbs = rng.uniform(0, 1, size=(M, ch)).astype(np.float32)
    return dict(sigmas=sigmas, rgbs=rgbs, deltas=deltas, rays=np.ascontiguousarray(rays), N=N, M=M)


def grid_case(seed, B=1000, D=3, C=2, L=16, base_resolution=16, desired_resolution=2048, log2_hashmap_size=19,
              align_corners=False):
    from oracle import oracle as orc
    rng = np.random.default_rng(seed)
    per_level_scale = float(np.exp2(np.log2(desired_resolution / base_resolution) / max(L - 1, 1)))
    offsets = orc.grid_offsets(D, L, base_resolution, per_level_scale, log2_hashmap_size, align_corners)
    table = rng.uniform(-1, 1, size=(int(offsets[-1]), C)).astype(np.float32)
    x = rng.uniform(0, 1, size=(B, D)).astype(np.float32)
    # edge cases: exact corners, the [0,1] boundary, and out-of-range rows (whole output must be zero)
    x[0] = 0.0
    x[1] = 1.0
    x[2, 0] = -1e-3
    x[3, -1] = 1.0 + 1e-3
    x[4] = 0.5
    return dict(inputs=x, table=table, offsets=offsets, per_level_scale=per_level_scale,
                base_resolution=base_resolution, D=D, C=C, L=L, align_corners=align_corners)


def ffmlp_case(seed, B=256, input_dim=32, hidden_dim=64, num_layers=2, output_dim=16):
    rng = np.random.default_rng(seed)
    n_w = hidden_dim * (input_dim + hidden_dim * (num_layers - 1) + output_dim)
    bound = np.sqrt(3 / hidden_dim)
    w = rng.uniform(-bound, bound, size=n_w).astype(np.float16)
    x = rng.uniform(-1, 1, size=(B, input_dim)).astype(np.float16)
    g = (rng.normal(size=(B, output_dim)) * 0.1).astype(np.float16)
    return dict(x=x, w=w, g=g, input_dim=input_dim, hidden_dim=hidden_dim, num_layers=num_layers,
                output_dim=output_dim)
