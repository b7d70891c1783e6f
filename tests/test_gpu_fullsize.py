"""Full-size (BASELINE.json configs[1]) checks of the hot path through size-independent properties: 4096 rays of the
64x1024 synthetic pano, hash grid L16 F2 T2^19 res 16->32768, max_steps 1024 - sizes at which the CPU oracle would
take minutes, so parity is established through invariants of the reference's algorithm instead
(raymarching.cu:332-568 bookkeeping, :578-802 compositing, linearity of the backward pass, Adam's fp16 shadow)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N = 4096


@pytest.fixture(scope="module")
def full():
    from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig
    from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence
    cfg = FieldConfig(grid_update_interval=0, perturb=False)
    seq = SyntheticLidarSequence(n_frames=2, device=DEV)
    eng = LidarFieldEngine(cfg, N, device=DEV, sample_budget=N * 256)
    eng.seed_occupancy_from_points(seq.surface_points())
    gen = torch.Generator().manual_seed(7)
    ro, rd, gt = seq.sample_batch(N, frame=0, generator=gen, device=DEV)
    eng.set_batch(ro, rd, gt)
    eng.G.zero_()
    eng.loss_acc.zero_()
    # poison the per-sample buffers: the step must not depend on their previous contents
    eng.xyzs.fill_(5.0)
    eng.deltas.fill_(5.0)
    eng.ray_ids.fill_(N - 1)
    eng.g_sigma.fill_(float("nan"))
    eng.g_rgb.fill_(float("nan"))
    eng._forward_backward()
    torch.cuda.synchronize()
    return eng, seq


def test_march_bookkeeping_at_full_size(full):
    """rays[:, (index, offset, count)] tiles [0, counter[0]) without gaps or overlaps; every sample lies on its ray
    at the distance the deltas integrate to (raymarching.cu:452-457,508-524)."""
    eng, _ = full
    rays = eng.rays.cpu().numpy().astype(np.int64)
    total, n_rays = (int(x) for x in eng.counter[:2].cpu().numpy())
    assert n_rays == N and total > 50 * N, (n_rays, total)
    assert total <= eng.M, "the sample budget must hold every ray (no dropped rays in the bench configuration)"
    assert sorted(rays[:, 0].tolist()) == list(range(N)), "every ray owns exactly one record"
    by_off = rays[np.argsort(rays[:, 1], kind="stable")]
    nz = by_off[by_off[:, 2] > 0]
    assert nz[0, 1] == 0
    np.testing.assert_array_equal(nz[1:, 1], nz[:-1, 1] + nz[:-1, 2])
    assert nz[-1, 1] + nz[-1, 2] == total == rays[:, 2].sum()
    assert rays[:, 2].max() <= eng.cfg.max_steps
    # per-sample ray ids (extended march) agree with the records
    ids = eng.ray_ids[:total].cpu().numpy()
    want = np.empty(total, np.int64)
    for n, off, cnt in nz:
        want[off:off + cnt] = n
    np.testing.assert_array_equal(ids, want)
    # geometry: xyz = o + t d with t = t0 + cumsum(delta1) (t advances by dt after each emitted sample, :519-523)
    xyz = eng.xyzs[:total].cpu().numpy().astype(np.float64)
    dl = eng.deltas[:total].cpu().numpy().astype(np.float64)
    o = eng.rays_o.cpu().numpy().astype(np.float64)
    d = eng.rays_d.cpu().numpy().astype(np.float64)
    t0 = eng.t0.cpu().numpy().astype(np.float64)
    assert (dl[:, 0] > 0).all() and (dl[:, 1] >= dl[:, 0] - 1e-7).all()
    for n, off, cnt in nz[:: max(1, len(nz) // 256)]:
        t_after = t0[n] + np.cumsum(dl[off:off + cnt, 1])
        t_at = t_after - dl[off:off + cnt, 0]
        np.testing.assert_allclose(xyz[off:off + cnt], o[n] + t_at[:, None] * d[n], atol=2e-5)
    assert np.abs(xyz).max() <= eng.cfg.bound + 1e-6
    # the padding rows of the last 128-row tile are zeroed by the extended march, and carry zero gradients
    rows = (total + 127) // 128 * 128
    if rows > total:
        assert float(eng.xyzs[total:rows].abs().max()) == 0 and float(eng.deltas[total:rows].abs().max()) == 0
        assert int(eng.ray_ids[total:rows].abs().max()) == 0
        assert float(eng.g_sigma[total:rows].abs().max()) == 0 and float(eng.g_rgb[total:rows].abs().max()) == 0
    assert float(eng.xyzs[rows:rows + 128].min()) == 5.0, "rows beyond the last tile are not touched"


def test_every_sample_sits_in_an_occupied_cell(full):
    """The march emits a sample only where the bitfield is set (raymarching.cu:407-408)."""
    eng, _ = full
    from oracle import oracle as orc
    total = int(eng.counter[0].item())
    xyz = eng.xyzs[:total].cpu().numpy()
    H = eng.cfg.grid_size
    # same cell arithmetic as the kernel (Appendix B): fp64 product narrowed to fp32, truncated
    cell = np.clip((0.5 * (xyz.astype(np.float64) * 1.0 + 1) * H).astype(np.float32), 0, H - 1).astype(np.int32)
    idx = orc.morton3D(cell).astype(np.int64)
    bits = np.unpackbits(eng.bitfield.cpu().numpy(), bitorder="little")
    assert bits[idx].all()


def test_composite_invariants_at_full_size(full):
    eng, _ = full
    ws, depth, img = eng.ws.cpu().numpy(), eng.depth.cpu().numpy(), eng.image.cpu().numpy()
    assert np.isfinite(ws).all() and np.isfinite(depth).all() and np.isfinite(img).all()
    assert ws.min() >= 0 and ws.max() <= 1 + 1e-5
    assert (img >= 0).all() and (img <= ws[:, None] + 1e-5).all(), "sum_i w_i * sigmoid <= sum_i w_i"
    far = eng.cfg.min_near_lidar * eng.cfg.far_factor
    assert (depth >= 0).all() and (depth <= ws * (far + 0.01) + 1e-5).all()
    # independent fp64 recomposition of a subset of rays from the per-sample sigma / rgb the kernels produced
    rays = eng.rays.cpu().numpy()
    sig = eng.sigma.cpu().numpy().astype(np.float64)
    rgb = eng.rgb.cpu().numpy().astype(np.float64)
    dl = eng.deltas.cpu().numpy().astype(np.float64)
    for n, off, cnt in rays[:: N // 128]:
        a = 1 - np.exp(-sig[off:off + cnt] * dl[off:off + cnt, 0])
        T = np.concatenate([[1.0], np.cumprod(1 - a)[:-1]])
        stop = np.nonzero(np.cumprod(1 - a) < eng.cfg.T_thresh)[0]
        k = cnt if len(stop) == 0 else stop[0] + 1            # the sample that crosses the threshold still counts (:648-651)
        w = (a * T)[:k]
        np.testing.assert_allclose(ws[n], w.sum(), rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(img[n], (w[:, None] * rgb[off:off + k]).sum(0), rtol=2e-4, atol=2e-5)


def test_backward_is_linear_in_the_loss_scale(full):
    """Doubling the loss scale doubles every gradient (powers of two are exact in fp16/fp32 away from overflow)."""
    eng, _ = full
    g1 = eng.G.clone()
    assert torch.isfinite(g1).all() and float(g1.abs().sum()) > 0
    eng.cfg.loss_scale *= 2
    try:
        eng.G.zero_()
        eng._forward_backward()
        torch.cuda.synchronize()
        g2 = eng.G.clone()
    finally:
        eng.cfg.loss_scale /= 2
        eng.G.zero_()                # leave every buffer of the fixture in the state of the original scale
        eng.loss_acc.zero_()
        eng._forward_backward()
        torch.cuda.synchronize()
    num = float((g2.double() - 2 * g1.double()).norm())
    den = float(g1.double().norm()) * 2
    assert num / den < 2e-3, num / den     # atomics reorder fp32 sums; fp16 denormals at the tails


def test_table_gradient_touches_only_rows_the_samples_address(full):
    """Checksum-of-checksums: the per-level gradient mass equals what the feature gradient carries into that level
    (sum over corners of the trilinear weights is 1, gridencoder.cu:309-352), and levels are independent."""
    eng, _ = full
    c = eng.cfg
    # the backward pass runs on the compact list of live rows (counter[2] of them); g_enc is in that order
    rows = int(eng.counter[2].item()) if c.compact_backward else int(eng.counter[0].item())
    assert 0 < rows <= int(eng.counter[0].item())
    g_enc = eng.g_enc[:rows].float().cpu().numpy().astype(np.float64)          # [rows, L*C]
    g_tab = eng.g_table.cpu().numpy().astype(np.float64).reshape(-1, c.level_dim)
    offs = eng.offsets.cpu().numpy()
    inside = np.ones(rows, bool)                                               # marched samples are clamped into the box
    for level in range(c.num_levels):
        want = g_enc[inside, level * c.level_dim:(level + 1) * c.level_dim].sum(0)
        got = g_tab[offs[level]:offs[level + 1]].sum(0)
        scale = np.abs(g_enc[inside, level * c.level_dim:(level + 1) * c.level_dim]).sum(0) + 1e-12
        assert (np.abs(got - want) / scale < 2e-3).all(), (level, got, want)


def test_graph_replay_reproduces_eager_at_full_size(full):
    eng, _ = full
    g_eager = eng.G.clone()
    loss_eager = float(eng.loss_acc.item())
    eng._capture()
    eng.G.zero_()
    eng.loss_acc.zero_()
    eng._graph.replay()
    torch.cuda.synchronize()
    rel = float((eng.G.double() - g_eager.double()).norm() / g_eager.double().norm())
    assert rel < 1e-3, rel
    np.testing.assert_allclose(float(eng.loss_acc.item()), loss_eager, rtol=1e-4)
    eng.G.copy_(g_eager)


def test_adam_keeps_the_fp16_shadow_in_sync_at_full_size(full):
    eng, _ = full
    p0 = eng.P.clone()
    eng._optimizer()
    torch.cuda.synchronize()
    assert torch.isfinite(eng.P).all()
    assert torch.equal(eng.Ph[:eng.n_params], eng.P[:eng.n_params].to(torch.float16))
    if eng.cfg.late_grad_zero:      # the gradient is cleared by the next backward pass, right before the scatter
        assert float(eng.G.abs().max()) > 0.0
    else:
        assert float(eng.G.abs().max()) == 0.0, "Adam zeroes the gradient for the next step"
    moved = (eng.P != p0).float().mean().item()
    assert moved > 0.01, "parameters addressed by the batch must move"
    # first Adam step: |delta| <= lr for every parameter (bias-corrected m/sqrt(v) = +-1)
    assert float((eng.P - p0).abs().max()) <= eng.cfg.lr * 1.001


def test_full_size_step_matches_cpu_restatement_at_4096_rays():
    """BASELINE config 2 at full size - 4096 rays, the 2^19-entry table (13.7 M parameters), max_steps 1024, clumpy
    occupancy grid: ONE fused step (loss, per-ray outputs, per-ray sample counts, every gradient) against the CPU
    restatement.  (The restatement takes ~20 s on the host cores for this size.)"""
    from oracle import check_engine
    cfg = check_engine.small_config(log2_hashmap_size=19, desired_resolution=32768, max_steps=1024)
    eng, gpu, cpu = check_engine.run_pair(n_rays=4096, device=DEV, cfg=cfg, seed=2, fill=0.05)
    assert eng.n_params == 13693520 and gpu["n_samples"] > 100000
    check_engine.compare(gpu, cpu, eng.n_table)
