"""Hypothesis-driven shape / edge-case sweeps of the CPU oracle ("every parity claim is pinned by
our own harness ... (iii) hypothesis-driven shape/edge-case sweeps").  The oracle is what the GPU kernels are compared
against, so its own invariants are checked here over randomly drawn configurations: ragged and empty inputs, cascades,
step laws, thresholds, table shapes.  CPU only, a few seconds in total."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import cases

SET = dict(deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow], derandomize=True)


@settings(max_examples=40, **SET)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(0, 300), top=st.sampled_from([1, 2, 128, 1024]))
def test_morton_round_trip(orc, seed, n, top):
    rng = np.random.default_rng(seed)
    c = rng.integers(0, top, size=(n, 3)).astype(np.int32)
    idx = orc.morton3D(c)
    assert idx.shape == (n,)
    np.testing.assert_array_equal(orc.morton3D_invert(idx), c)
    if n:
        # bit interleave: x -> bits 0,3,6.., y -> 1,4,7.., z -> 2,5,8..  (raymarching.cu:71-95)
        x, y, z = (int(v) for v in c[0])
        want = sum(((x >> b) & 1) << (3 * b) | ((y >> b) & 1) << (3 * b + 1) | ((z >> b) & 1) << (3 * b + 2) for b in range(10))
        assert int(idx[0]) == want


@settings(max_examples=30, **SET)
@given(seed=st.integers(0, 2 ** 31 - 1), nbytes=st.integers(1, 200), thresh=st.floats(-1, 2))
def test_packbits_is_a_little_endian_threshold(orc, seed, nbytes, thresh):
    g = np.random.default_rng(seed).uniform(-1, 2, size=nbytes * 8).astype(np.float32)
    got = orc.packbits(g, thresh)
    want = np.packbits((g > np.float32(thresh)).reshape(-1, 8), axis=1, bitorder="little").reshape(-1)
    np.testing.assert_array_equal(got, want)


@settings(max_examples=25, **SET)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 40), cascade=st.sampled_from([1, 1, 2, 3]),
       dt_gamma=st.sampled_from([0.0, 0.0, 1.0 / 256, 1.0 / 128]), max_steps=st.sampled_from([16, 64, 300, 1024]),
       fill=st.sampled_from([0.0, 0.1, 0.5, 1.0]), lidar=st.booleans())
def test_march_bookkeeping_and_geometry(orc, seed, n, cascade, dt_gamma, max_steps, fill, lidar):
    bound = float(2 ** (cascade - 1))
    c = cases.march_case(seed, n, cascade, bound, 128, fill, lidar)
    if lidar:
        nears = np.full(n, 0.0108, np.float32)
        fars = nears * np.float32(81.0)
    else:
        nears, fars = orc.near_far_from_aabb(c["rays_o"], c["rays_d"], np.array([-bound] * 3 + [bound] * 3, np.float32), 0.05)
    x, d, dl, rays, counter = orc.march_rays_train(c["rays_o"], c["rays_d"], bound, c["bitfield"], cascade, 128, nears, fars,
                                                   c["noises"], dt_gamma, max_steps)
    tot = int(counter[0])
    assert int(counter[1]) == n and tot == int(rays[:, 2].sum())
    assert (rays[:, 2] >= 0).all() and (rays[:, 2] <= max_steps).all()
    assert sorted(rays[:, 0].tolist()) == list(range(n))
    by_off = rays[np.argsort(rays[:, 1], kind="stable")]
    nz = by_off[by_off[:, 2] > 0]
    if len(nz):
        assert nz[0, 1] == 0 and (nz[1:, 1] == nz[:-1, 1] + nz[:-1, 2]).all()
    if fill == 0.0:
        assert tot == 0
    assert (np.abs(x[:tot]) <= bound).all()
    dt_min = np.float32(2 * np.sqrt(3)) / np.float32(max_steps)
    dt_max = np.float32(2 * np.sqrt(3)) * np.float32(2 ** (cascade - 1)) / np.float32(128)
    lo, hi = min(dt_min, dt_max), max(dt_min, dt_max)
    assert (dl[:tot, 0] >= lo * (1 - 1e-6)).all() and (dl[:tot, 0] <= hi * (1 + 1e-6)).all()
    # (t + dt) - last_t is evaluated in fp32 at t ~ bound: it can fall short of dt by a few ulp(t)
    assert (dl[:tot, 1] >= dl[:tot, 0] * (1 - 2e-3)).all(), "the real step is at least the sampled interval"
    for rid, off, cnt in rays:
        if cnt == 0:
            continue
        # samples of a ray advance along it: t = (x - o) . d is strictly increasing, dirs are copies of the direction
        t = ((x[off:off + cnt] - c["rays_o"][rid]) * c["rays_d"][rid]).sum(1)
        inside = (np.abs(x[off:off + cnt]) < bound).all(1)          # clamped samples leave the line
        tt = t[inside]
        assert (np.diff(tt) > 0).all()
        np.testing.assert_array_equal(d[off:off + cnt], np.tile(c["rays_d"][rid], (cnt, 1)))
    if cascade == 1 and tot:
        bits = np.unpackbits(c["bitfield"], bitorder="little")
        cell = np.clip((0.5 * (x[:tot].astype(np.float64) / bound + 1) * 128).astype(np.float32), 0, 127).astype(np.int32)
        assert bits[orc.morton3D(cell).astype(np.int64)].all(), "a sample was emitted in an empty cell"


@settings(max_examples=30, **SET)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 40), max_count=st.integers(1, 90), ch=st.sampled_from([1, 2, 3]),
       opaque=st.sampled_from([0.0, 0.5, 1.0]), thresh=st.sampled_from([0.0, 1e-4, 1e-2]))
def test_composite_invariants_and_gradient(orc, seed, n, max_count, ch, opaque, thresh):
    c = cases.composite_case(seed, N=max(n, 5), max_count=max_count + 1, ch=ch, opaque_frac=opaque)
    ws, dep, img = orc.composite_rays_train_forward(c["sigmas"], c["rgbs"], c["deltas"], c["rays"], thresh)
    assert np.isfinite(ws).all() and (ws >= 0).all() and (ws <= 1 + 1e-5).all()
    assert (img >= -1e-6).all() and (img <= ws[:, None] + 1e-5).all()          # colours are in [0, 1]
    empty = c["rays"][c["rays"][:, 2] == 0][:, 0]
    assert (ws[empty] == 0).all() and (dep[empty] == 0).all() and (img[empty] == 0).all()
    # zero density -> nothing accumulates
    ws0, dep0, img0 = orc.composite_rays_train_forward(np.zeros_like(c["sigmas"]), c["rgbs"], c["deltas"], c["rays"], thresh)
    assert (ws0 == 0).all() and (img0 == 0).all()
    # backward is the gradient of L = sum(gw * ws) + sum(gi * img) w.r.t. the colours (exact: img is linear in them)
    rng = np.random.default_rng(seed)
    gw = rng.normal(size=ws.shape).astype(np.float32)
    gi = rng.normal(size=img.shape).astype(np.float32)
    gs, gc = orc.composite_rays_train_backward(gw, gi, c["sigmas"], c["rgbs"], c["deltas"], c["rays"], ws, img, thresh)
    assert np.isfinite(gs).all() and np.isfinite(gc).all()
    rgb2 = c["rgbs"] + 0.25
    _, _, img2 = orc.composite_rays_train_forward(c["sigmas"], rgb2, c["deltas"], c["rays"], thresh)
    np.testing.assert_allclose((gi * (img2 - img)).sum(), (gc * 0.25).sum(), rtol=2e-3, atol=1e-4)


@settings(max_examples=25, **SET)
@given(seed=st.integers(0, 10 ** 6), B=st.integers(1, 60), L=st.integers(1, 8), C=st.sampled_from([1, 2, 4, 8]),
       D=st.sampled_from([2, 3]), log2=st.sampled_from([6, 10, 14]), ac=st.booleans(), interp=st.sampled_from([0, 1]))
def test_grid_encode_linearity_range_and_adjoint(orc, seed, B, L, C, D, log2, ac, interp):
    c = cases.grid_case(seed, B=max(B, 5), D=D, C=C, L=L, base_resolution=4, desired_resolution=64, log2_hashmap_size=log2,
                        align_corners=ac)
    kw = dict(gridtype=0, align_corners=ac, interp=interp)
    y = orc.grid_encode_forward(c["inputs"], c["table"], c["offsets"], c["per_level_scale"], 4, **kw)
    assert y.shape == (max(B, 5), L * C) and np.isfinite(y).all()
    oob = ((c["inputs"] < 0) | (c["inputs"] > 1)).any(1)
    assert oob.any() and (y[oob] == 0).all(), "out-of-range inputs encode to zero on every level"
    # linear in the table ...
    y2 = orc.grid_encode_forward(c["inputs"], 2 * c["table"], c["offsets"], c["per_level_scale"], 4, **kw)
    np.testing.assert_allclose(y2, 2 * y, rtol=1e-5, atol=1e-6)
    # ... interpolation weights sum to one: a constant table encodes to the constant
    yc = orc.grid_encode_forward(c["inputs"], np.full_like(c["table"], 0.75), c["offsets"], c["per_level_scale"], 4, **kw)
    np.testing.assert_allclose(yc[~oob], 0.75, rtol=1e-5)
    # ... and the backward scatter is its adjoint: <g, E(table)> == <E^T g, table>
    g = np.random.default_rng(seed + 1).normal(size=y.shape).astype(np.float32)
    gt = orc.grid_encode_backward(g, c["inputs"], c["table"].shape, c["offsets"], c["per_level_scale"], 4, **kw)
    np.testing.assert_allclose((g.astype(np.float64) * y).sum(), (gt.astype(np.float64) * c["table"]).sum(), rtol=2e-4, atol=1e-4)


@settings(max_examples=25, **SET)
@given(seed=st.integers(0, 10 ** 6), B=st.integers(1, 50), deg=st.integers(0, 12))
def test_freq_encode_closed_form(orc, seed, B, deg):
    x = np.random.default_rng(seed).uniform(-1, 1, size=(B, 3)).astype(np.float32)
    y = orc.freq_encode_forward(x, deg)
    assert y.shape == (B, 3 + 6 * deg)
    np.testing.assert_array_equal(y[:, :3], x)
    for f in range(deg):
        arg = x.astype(np.float64) * 2.0 ** f
        np.testing.assert_allclose(y[:, 3 + 6 * f:6 + 6 * f], np.sin(arg), atol=2e-3)        # __sinf-level accuracy
        np.testing.assert_allclose(y[:, 6 + 6 * f:9 + 6 * f], np.cos(arg), atol=2e-3)


@settings(max_examples=20, **SET)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 120), m=st.integers(1, 120))
def test_chamfer_properties(orc, seed, n, m):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(1, n, 3)).astype(np.float32)
    b = rng.normal(size=(1, m, 3)).astype(np.float32)
    d1, d2, i1, i2 = orc.chamfer_forward(a, b)
    assert (d1 >= 0).all() and (d2 >= 0).all() and (0 <= i1).all() and (i1 < m).all() and (i2 < n).all()
    # the reported index realises the reported distance; no other point is closer
    np.testing.assert_allclose(d1[0], ((a[0] - b[0][i1[0]]) ** 2).sum(1), rtol=1e-5, atol=1e-7)
    full = ((a[0][:, None, :].astype(np.float64) - b[0][None]) ** 2).sum(-1)
    assert (d1[0] <= full.min(1) * (1 + 1e-5) + 1e-7).all()
    # a cloud against itself: zero distance, identity matches (distinct points)
    s1, s2, j1, j2 = orc.chamfer_forward(a, a)
    assert (s1 == 0).all() and (s2 == 0).all() and (j1[0] == np.arange(n)).all()


@settings(max_examples=20, **SET)
@given(seed=st.integers(0, 10 ** 6), H=st.sampled_from([8, 16, 64]), W=st.sampled_from([32, 256, 1024]),
       density=st.sampled_from([0.0, 0.05, 0.6, 1.0]))
def test_range_image_round_trip(orc, seed, H, W, density):
    rng = np.random.default_rng(seed)
    K = (2.0, 26.9)
    pano = (rng.uniform(1, 79, size=(H, W)) * (rng.random((H, W)) < density)).astype(np.float32)
    inten = rng.uniform(0, 1, size=(H, W)).astype(np.float32) * (pano != 0)
    pts = orc.pano_to_lidar_with_intensities(pano, inten, K)
    assert pts.shape == (int((pano != 0).sum()), 4)
    np.testing.assert_allclose(np.linalg.norm(pts[:, :3], axis=1), pano[pano != 0], rtol=1e-5)
    np.testing.assert_array_equal(pts[:, 3], inten[pano != 0])                   # row-major order of the non-empty pixels
    back, inten2 = orc.lidar_to_pano_with_intensities(pts, H, W, K, max_depth=80)
    # every point lands in its own column; rows may shift by one where a beam centre sits on a rounding boundary
    assert ((back != 0).sum(0) == (pano != 0).sum(0)).mean() > 0.9 or density == 0.0
    same = np.isclose(back, pano, rtol=1e-5)
    assert same.mean() > 0.9
    far = orc.lidar_to_pano_with_intensities(pts * np.array([100, 100, 100, 1], np.float32), H, W, K, max_depth=80)[0]
    assert (far == 0).all(), "points at or beyond max_depth are dropped"
