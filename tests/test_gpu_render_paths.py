"""GPU tests of `NeRFRenderer.render()`'s three paths (dense `run`, per-op `run_cuda`, fused `run_fused`): absolute
LiDAR depth is consistent between them (the eval branch double-counted the near plane, the training branch
ignored the march jitter), `render()` defaults to the fused kernels on CUDA, and the fused autograd Function agrees with
the per-op chain in outputs and gradients."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _shell_field(near):
    from lidar_nerf_b200.nerf.renderer import NeRFRenderer

    class Shell(NeRFRenderer):            # a thin opaque shell of radius 0.4 around the origin
        def __init__(self):
            super().__init__(bound=1, min_near_lidar=near)
            self.out_color_dim, self.out_lidar_color_dim = 3, 2

        def density(self, x):
            r = x.norm(dim=-1)
            return {"sigma": 4000.0 * torch.exp(-((r - 0.4) / 0.01) ** 2), "geo_feat": x[:, :1] * 0.5 + 0.5}

        def color(self, x, d, cal_lidar_color=False, mask=None, geo_feat=None, **kw):
            c = torch.stack([torch.sigmoid(3 * x[:, 0] + d[:, 2]), torch.sigmoid(geo_feat[:, 0] - d[:, 0])], -1)
            return c * mask[:, None] if mask is not None else c
    return Shell().to(DEV)


def test_absolute_depth_agrees_between_dense_march_train_and_march_eval():
    near = 0.05                                   # large on purpose: a near-plane double count would show as +0.05
    f = _shell_field(near)
    f.grid_update_interval = 0                    # all-occupied bitfield: the march visits every step
    g = torch.Generator().manual_seed(0)
    n = 512
    d = torch.randn(n, 3, generator=g)
    d = (d / d.norm(dim=-1, keepdim=True)).to(DEV)
    o = (0.05 * torch.randn(n, 3, generator=g)).to(DEV)
    # distance to the shell along each ray, analytically
    b = (o * d).sum(-1)
    t_true = -b + torch.sqrt(b * b - (o * o).sum(-1) + 0.4 ** 2)
    kw = dict(cal_lidar_color=True, staged=False, dt_gamma=0, max_steps=1024)
    f.eval()
    with torch.no_grad():
        dense = f.render(o[None], d[None], cuda_ray="dense", num_steps=1536, upsample_steps=0, perturb=False, **kw)
        ev = f.render(o[None], d[None], cuda_ray="ops", perturb=False, **kw)
    f.train()
    with torch.no_grad():
        tr0 = f.render(o[None], d[None], cuda_ray="ops", perturb=False, **kw)
        tr1 = f.render(o[None], d[None], cuda_ray="ops", perturb=True, **kw)
    # the shell becomes opaque where sigma * dt ~ 1, i.e. ~0.015 in front of its centre; one march step is
    # 2 sqrt(3) / 1024 = 0.0034, one dense step 0.0026: every path must land within a few steps of the dense answer
    # (a near-plane double count would be +0.05, a lost jitter offset up to one march step)
    ref = dense["depth_lidar"][0]
    assert float((ref - t_true).abs().max()) < 0.03
    for name, out in (("eval", ev), ("train", tr0), ("train+jitter", tr1)):
        ws = out["weights_sum_lidar"]
        assert float(ws.min()) > 0.99, (name, float(ws.min()))
        err = (out["depth_lidar"][0] - ref).abs()
        assert float(err.max()) < 0.008, (name, float(err.max()))
        assert abs(float((out["depth_lidar"][0] - ref).mean())) < 0.004, (name, float((out["depth_lidar"][0] - ref).mean()))
    assert float((ev["depth_lidar"] - tr0["depth_lidar"]).abs().max()) < 0.008


def _small_net(seed=0):
    from lidar_nerf_b200.nerf.network import NeRFNetwork
    torch.manual_seed(seed)
    net = NeRFNetwork(encoding="hashgrid", desired_resolution=2048, log2_hashmap_size=15, bound=1,
                      min_near_lidar=1 / 92.7, density_thresh=10).to(DEV)
    g = torch.Generator().manual_seed(seed + 1)
    for prm, b in ((net.encoder.embeddings, 0.5), (net.sigma_net.weights, 0.25), (net.lidar_color_net.weights, 0.25)):
        prm.data.copy_(torch.empty(prm.shape).uniform_(-b, b, generator=g))
    net.grid_update_interval = 0
    return net


def test_render_defaults_to_the_fused_kernels_on_cuda():
    from lidar_nerf_b200 import _lib
    from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence
    net = _small_net()
    assert net.fused_unsupported_reason() is None
    seq = SyntheticLidarSequence(H=16, W=256, n_frames=1, device=DEV)
    ro, rd, gt = seq.sample_batch(256, generator=torch.Generator().manual_seed(0), device=DEV)
    net.train()
    n0 = _lib.launch_count()
    out = net.render(ro[None], rd[None], cal_lidar_color=True, staged=False, perturb=True, max_steps=256)
    assert net._fused is not None and len(net._fused._engines) == 1          # the fused workspace served the call
    assert out["depth_lidar"].requires_grad and out["depth_lidar"].shape == (1, 256)
    out["depth_lidar"].sum().backward()
    assert net.encoder.embeddings.grad is not None and float(net.encoder.embeddings.grad.abs().sum()) > 0
    assert _lib.launch_count() - n0 >= 8
    # unsupported network -> per-op occupancy path, still on the sm_100a kernels
    from lidar_nerf_b200.nerf.network import NeRFNetwork
    lin = NeRFNetwork(encoding="hashgrid", desired_resolution=2048, log2_hashmap_size=15, bound=1, use_ffmlp=False,
                      min_near_lidar=1 / 92.7).to(DEV)
    assert lin.fused_unsupported_reason() is not None
    assert lin._pick_path(ro, True, None) == lin.run_cuda
    with pytest.raises(RuntimeError, match="fused"):
        lin.render(ro[None], rd[None], cal_lidar_color=True, cuda_ray="fused")


def test_fused_function_matches_the_per_op_chain():
    """Same parameters, rays, no jitter: run_fused (one Function over the fused kernels) == run_cuda (grid_encode ->
    FFMLP -> trunc_exp -> freq_encode -> FFMLP -> sigmoid -> composite, each with its own autograd node)."""
    from conftest import record_parity
    from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence
    seq = SyntheticLidarSequence(H=16, W=256, n_frames=1, device=DEV)
    ro, rd, gt = seq.sample_batch(512, generator=torch.Generator().manual_seed(2), device=DEV)
    m = gt[:, 0]
    res = {}
    for mode in ("ops", "fused"):
        net = _small_net(seed=4)
        net.train()
        with torch.autocast("cuda", dtype=torch.float16):
            out = net.render(ro[None], rd[None], cal_lidar_color=True, staged=False, perturb=False, cuda_ray=mode,
                             max_steps=256, force_all_rays=True)
        loss = (10 * (out["depth_lidar"][0] * m - gt[:, 2] * m).abs() + (out["image_lidar"][0, :, 0] - m) ** 2
                + 10 * (out["image_lidar"][0, :, 1] * m - gt[:, 1] * m) ** 2).mean()
        (loss * 64.0).backward()
        res[mode] = dict(depth=out["depth_lidar"][0].detach(), image=out["image_lidar"][0].detach(),
                         ws=out["weights_sum_lidar"].detach(), loss=float(loss),
                         g_emb=net.encoder.embeddings.grad.float().reshape(-1).clone(),
                         g_sig=net.sigma_net.weights.grad.float().clone(),
                         g_head=net.lidar_color_net.weights.grad.float().clone())
    a, b = res["fused"], res["ops"]
    for k in ("depth", "image", "ws"):
        np.testing.assert_allclose(a[k].cpu().numpy(), b[k].cpu().numpy(), rtol=5e-3, atol=2e-3, err_msg=k)
    assert abs(a["loss"] - b["loss"]) < 5e-3 * abs(b["loss"])
    for k in ("g_emb", "g_sig", "g_head"):
        rel = float((a[k] - b[k]).norm() / b[k].norm())
        cos = float((a[k] @ b[k]) / (a[k].norm() * b[k].norm()))
        record_parity(f"fused_function_vs_per_op[{k}]", rel=rel, one_minus_cos=1 - cos)
        # the per-op chain rounds the table gradient to fp16 per contribution (half2 atomics, gridencoder.cu:346-353);
        # the fused backward accumulates in fp32
        assert cos > 0.999 and rel < 5e-2, (k, rel, cos)
