"""The reference's UNMODIFIED Trainer / entry script driving this library.

  * CPU (`-m "not gpu"`): BASELINE config 1 - the reference's `Trainer.train_step` + optimiser loop over this package's
    `NeRFRenderer.run` (dense sampling) with a tiny pure-torch field (torch frequency encoding + 2-layer MLP) on a
    16 x 256 synthetic range image: the render() keyword contract (`**vars(opt)`), loss plumbing and training.
  * GPU (`-m gpu`): `main_lidarnerf.py -L` run as a script, unmodified, through `lidar_nerf_b200.compat` on a synthetic
    KITTI-360-format sequence written to disk: dataset -> Trainer -> `render()` -> fused sm_100a kernels -> GradScaler /
    torch Adam -> evaluation -> test -> mesh.  Asserts that the fused kernels were launched and the loss fell, and
    compares one `train_step` of the Trainer with the fused engine on the same rays and parameters.
The reference's Python comes from /root/reference (build container) or from the sourceless .pyc tree in oracle/_ref/pyref
(GPU box); see tests/refshim.py.
"""
import math
import os
import sys

import numpy as np
import pytest
import torch

import refshim


def _ref_or_skip():
    root, kind = refshim.install()
    if root is None:
        if torch.cuda.is_available() and os.environ.get("LNB_ALLOW_MISSING_REF") != "1":
            pytest.fail("reference Python not available (oracle/_ref/pyref missing): run `python oracle/build_ref.py pyref` "
                        "in the build container before shipping to the GPU box")
        pytest.skip("reference Python not available")
    return root, kind


def _load_main(root, kind):
    """The entry script as a module (NOT executed as __main__): gives get_arg_parser()."""
    import importlib.machinery
    import importlib.util
    path = os.path.join(root, "main_lidarnerf.py" + ("c" if kind == "pyc" else ""))
    loader = (importlib.machinery.SourcelessFileLoader if kind == "pyc" else importlib.machinery.SourceFileLoader)(
        "main_lidarnerf", path)
    spec = importlib.util.spec_from_loader("main_lidarnerf", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def _cleanup_modules():
    for k in list(sys.modules):
        if k == "lidarnerf" or k.startswith("lidarnerf.") or k == "main_lidarnerf":
            sys.modules.pop(k, None)


def test_reference_trainer_drives_dense_run_on_cpu(tmp_path):
    root, kind = _ref_or_skip()
    from lidar_nerf_b200 import compat
    from lidar_nerf_b200.nerf.renderer import NeRFRenderer
    from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence
    try:
        compat.install(networks=False)
        main = _load_main(root, kind)
        from lidarnerf.nerf.utils import Trainer            # the reference's own class

        class TinyField(NeRFRenderer):                       # BASELINE config 1: torch frequency encoding + 2-layer MLP
            def __init__(self, **kw):
                super().__init__(**kw)
                self.out_color_dim, self.out_lidar_color_dim = 3, 2
                g = torch.Generator().manual_seed(0)
                self.sig = torch.nn.ModuleList([torch.nn.Linear(3 + 6 * 4, 32), torch.nn.Linear(32, 1 + 4)])
                self.col = torch.nn.ModuleList([torch.nn.Linear(3 + 4, 32), torch.nn.Linear(32, 2)])
                for p in self.parameters():
                    p.data.uniform_(-0.3, 0.3, generator=g)

            @staticmethod
            def enc(x, deg=4):
                f = 2.0 ** torch.arange(deg, dtype=x.dtype)
                xf = (x[..., None, :] * f[:, None]).reshape(*x.shape[:-1], -1)
                return torch.cat([x, torch.sin(xf), torch.cos(xf)], -1)

            def density(self, x):
                h = self.sig[1](torch.relu(self.sig[0](self.enc(x))))
                return {"sigma": torch.exp(h[..., 0].clamp(max=8)), "geo_feat": h[..., 1:]}

            def color(self, x, d, cal_lidar_color=False, mask=None, geo_feat=None, **kw):
                return torch.sigmoid(self.col[1](torch.relu(self.col[0](torch.cat([d, geo_feat], -1)))))

            def get_params(self, lr):
                return [{"params": list(self.parameters()), "lr": lr}]

        opt = main.get_arg_parser().parse_args(["--workspace", str(tmp_path), "--num_steps", "24", "--upsample_steps", "8",
                                                "--num_rays_lidar", "256", "--scale", str(1 / 92.7), "--ckpt", "scratch"])
        opt.enable_lidar = True
        opt.min_near_lidar = opt.scale
        seq = SyntheticLidarSequence(H=16, W=256, n_frames=2, device="cpu")
        model = TinyField(bound=1, min_near_lidar=opt.scale)
        criterion = {"depth": torch.nn.L1Loss(reduction="none"), "raydrop": torch.nn.MSELoss(reduction="none"),
                     "intensity": torch.nn.MSELoss(reduction="none"), "grad": torch.nn.L1Loss(reduction="none")}
        trainer = Trainer("cfg1", opt, model, device=torch.device("cpu"), workspace=str(tmp_path), criterion=criterion,
                          optimizer=lambda m: torch.optim.Adam(m.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15),
                          fp16=False, use_checkpoint="scratch", use_tensorboardX=False, mute=True)
        gen = torch.Generator().manual_seed(1)
        losses = []
        for step in range(30):
            ro, rd, gt = seq.sample_batch(256, frame=step % 2, generator=gen)
            data = {"rays_o_lidar": ro[None], "rays_d_lidar": rd[None], "images_lidar": gt[None]}
            trainer.optimizer.zero_grad()
            _, _, pred_depth, gt_depth, loss = trainer.train_step(data)        # reference code: render(**vars(opt)) inside
            loss.backward()
            trainer.optimizer.step()
            losses.append(float(loss))
            assert pred_depth.shape == gt_depth.shape == (1, 256)
        assert np.isfinite(losses).all()
        assert np.mean(losses[-5:]) < 0.7 * np.mean(losses[:5]), losses
        # evaluation path of the same Trainer: staged render over the whole 16 x 256 image
        model.eval()
        pose = seq.poses[:1]
        from lidarnerf.dataset.base_dataset import get_lidar_rays
        rays = get_lidar_rays(pose, (seq.fov_up, seq.fov), 16, 256, -1)
        with torch.no_grad():
            rd_, in_, dp_ = trainer.test_step({"rays_o_lidar": rays["rays_o"], "rays_d_lidar": rays["rays_d"],
                                               "H_lidar": 16, "W_lidar": 256})
        assert dp_.shape == (1, 16, 256) and torch.isfinite(dp_).all()
    finally:
        _cleanup_modules()


# ------------------------------------------------------------------------------------------------------------ GPU
def _run_main(tmp_path, extra, iters=96, H=64, W=1024):
    from lidar_nerf_b200 import compat, _lib
    data = tmp_path / "kitti360"
    scale, offset = refshim.write_kitti360(str(data), "1908", H=H, W=W, n_train=4, n_val=1, n_test=1)
    cfg = tmp_path / "kitti360_1908.txt"            # same keys as the reference's configs/kitti360_1908.txt
    cfg.write_text("\n".join([
        "sequence_id = 1908", "alpha_d = 1000.0", "alpha_r = 1", "alpha_i = 1e1", "alpha_grad = 100.0", "grad_loss = True",
        "desired_resolution = 32768", "change_patch_size_lidar = [2, 8]", "num_steps = 768", "upsample_steps = 64",
        "bound = 1", f"scale = {scale}", "offset = [" + ", ".join(str(o) for o in offset) + "]"]) + "\n")
    root, kind = _ref_or_skip()
    script = os.path.join(root, "main_lidarnerf.py" + ("c" if kind == "pyc" else ""))
    argv = ["--config", str(cfg), "--path", str(data), "--workspace", str(tmp_path / "ws"), "-L", "--iters", str(iters),
            "--eval_interval", "1000", "--ckpt", "scratch", *extra]
    n0 = _lib.launch_count()
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        ns = compat.run_script(script, argv)
    finally:
        os.chdir(cwd)
    return ns, _lib.launch_count() - n0


@pytest.mark.gpu
def test_unmodified_main_lidarnerf_trains_on_the_fused_kernels(tmp_path, capfd):
    """`python main_lidarnerf.py --config configs/kitti360_1908.txt -L` (readme.md), unmodified, end to end."""
    _ref_or_skip()
    from lidar_nerf_b200.nerf import fused_render
    calls = {"fwd": 0, "bwd": 0}
    f0, b0 = fused_render._FusedLidarField.forward, fused_render._FusedLidarField.backward

    def fwd(*a, **k):
        calls["fwd"] += 1
        return f0(*a, **k)

    def bwd(*a, **k):
        calls["bwd"] += 1
        return b0(*a, **k)
    fused_render._FusedLidarField.forward = staticmethod(fwd)
    fused_render._FusedLidarField.backward = staticmethod(bwd)
    try:
        ns, launches = _run_main(tmp_path, [], iters=96)
    finally:
        fused_render._FusedLidarField.forward, fused_render._FusedLidarField.backward = staticmethod(f0), staticmethod(b0)
        _cleanup_modules()
    out = capfd.readouterr().out
    # 96 iterations over 4 training frames = 24 epochs of 4 steps; patch sampling [2, 8] + gradient loss every 2nd epoch
    assert calls["bwd"] == 96, calls
    assert calls["fwd"] >= 96 + 2 * 16, calls            # + evaluation and test of one 64 x 1024 frame in 4096-ray chunks
    assert launches > 96 * 10, launches                  # kernels of liblnb200.so (lnb::k_* : march, grid, field, ...)
    log = open(tmp_path / "ws" / "log_lidar_nerf.txt").read()
    assert "Finished Epoch 24" in log and "Finished Test" in log and "Finished saving mesh" in log
    # the Trainer logged an average loss per epoch: it must have fallen by a lot (depth L1 weighted 1e3)
    import re
    stats = [float(x) for x in re.findall(r"loss=([0-9.eE+-]+) \(", out + log)]
    ckpt = torch.load(tmp_path / "ws" / "checkpoints" / "lidar_nerf_ep0024.pth", map_location="cpu", weights_only=False)
    ep_loss = ckpt["stats"]["loss"]
    assert len(ep_loss) == 24 and np.isfinite(ep_loss).all()
    assert np.mean(ep_loss[-4:]) < 0.5 * np.mean(ep_loss[:2]), ep_loss
    # evaluation wrote the predicted point cloud of the test frame
    assert any(f.endswith("_depth_lidar.npy") for f in os.listdir(tmp_path / "ws" / "results"))
    del stats


@pytest.mark.gpu
def test_trainer_train_step_equals_fused_engine_step(tmp_path):
    """One `Trainer.train_step` (reference code: render -> torch loss) + backward through the fused Function equals the
    fused ENGINE step (loss and gradients computed inside the kernels) on the same parameters, rays and jitter."""
    root, kind = _ref_or_skip()
    from lidar_nerf_b200 import compat
    from lidar_nerf_b200.nerf.engine import LidarFieldEngine
    from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence
    try:
        compat.install()
        main = _load_main(root, kind)
        from lidarnerf.nerf.utils import Trainer
        from lidarnerf.nerf.network_tcnn import NeRFNetwork           # resolves to this library (compat.install)
        dev = torch.device("cuda:0")
        N = 2048
        seq = SyntheticLidarSequence(H=64, W=1024, n_frames=2, device=dev)
        opt = main.get_arg_parser().parse_args(["--workspace", str(tmp_path), "--scale", str(seq.scale), "--ckpt", "scratch",
                                                "--alpha_i", "10", "--dt_gamma", "0", "--bound", "1",
                                                "--desired_resolution", "32768"])
        opt.enable_lidar = True
        opt.fp16 = True
        opt.min_near_lidar = opt.scale
        model = NeRFNetwork(encoding="hashgrid", desired_resolution=opt.desired_resolution,
                            log2_hashmap_size=opt.log2_hashmap_size, n_features_per_level=2, num_layers=2, hidden_dim=64,
                            geo_feat_dim=15, bound=1, density_scale=1, min_near=opt.scale, min_near_lidar=opt.scale,
                            density_thresh=10, bg_radius=-1)
        g = torch.Generator().manual_seed(3)
        for prm, b in ((model.encoder.embeddings, 0.3), (model.sigma_net.weights, 0.2), (model.lidar_color_net.weights, 0.2)):
            prm.data = torch.empty(prm.shape).uniform_(-b, b, generator=g)   # large enough to be far from degenerate
        criterion = {"depth": torch.nn.L1Loss(reduction="none"), "raydrop": torch.nn.MSELoss(reduction="none"),
                     "intensity": torch.nn.MSELoss(reduction="none"), "grad": torch.nn.L1Loss(reduction="none")}
        trainer = Trainer("eq", opt, model, device=dev, workspace=str(tmp_path), criterion=criterion, fp16=True,
                          optimizer=lambda m: torch.optim.Adam(m.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15),
                          use_checkpoint="scratch", use_tensorboardX=False, mute=True)
        model.train()
        model.grid_update_interval = 0                                       # keep the all-occupied bitfield: same march
        assert model.fused_unsupported_reason() is None
        ro, rd, gt = seq.sample_batch(N, frame=0, generator=torch.Generator().manual_seed(5), device=dev)
        data = {"rays_o_lidar": ro[None], "rays_d_lidar": rd[None], "images_lidar": gt[None]}

        torch.manual_seed(11)                                                # the jitter noise is the first RNG draw
        trainer.optimizer.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            _, _, _, _, loss = trainer.train_step(data)
        scale = 128.0
        (loss * scale).backward()
        g_emb = model.encoder.embeddings.grad.reshape(-1) / scale
        g_sig = model.sigma_net.weights.grad / scale
        g_head = model.lidar_color_net.weights.grad / scale

        # the engine on the same parameters / rays / jitter: loss and gradients come out of its fused kernels
        cfg = model._fused.config(dt_gamma=0.0, max_steps=1024, T_thresh=1e-4, fused_composite=True, compact_backward=True,
                                  alpha_d=opt.alpha_d, alpha_r=opt.alpha_r, alpha_i=opt.alpha_i, loss_scale=scale,
                                  grid_update_interval=0)
        eng = LidarFieldEngine(cfg, N, device=dev, sample_budget=N * 260, external_params=True)
        eng.S = float(math.log2(model.encoder.per_level_scale))
        eng.load_params(model.encoder.embeddings, model.sigma_net.weights, model.lidar_color_net.weights)
        eng.set_grad_buffer(torch.zeros(eng.ex.n_padded, device=dev))
        eng.set_batch(ro, rd, gt)
        torch.manual_seed(11)
        eng._forward_backward()
        torch.cuda.synchronize()
        # the engine's loss is the mean over rays of the Trainer's per-ray terms (csrc/fused.cu: inv_n), unscaled; its
        # gradient carries loss_scale
        e_loss = float(eng.loss_acc)
        assert abs(e_loss - float(loss.detach())) <= 2e-3 * abs(float(loss.detach())), (e_loss, float(loss.detach()))

        def rel(a, b):
            return float((a - b).norm() / b.norm().clamp_min(1e-20))
        ge = eng.G / scale
        a, b = eng.n_table, eng.n_table + eng.n_sigma
        assert rel(g_emb, ge[:a]) < 1e-2, rel(g_emb, ge[:a])
        assert rel(g_sig, ge[a:b]) < 1e-2, rel(g_sig, ge[a:b])
        assert rel(g_head, ge[b:eng.n_params]) < 1e-2, rel(g_head, ge[b:eng.n_params])
    finally:
        _cleanup_modules()


@pytest.mark.gpu
def test_reference_network_class_equals_this_librarys_on_shared_weights():
    """a15: the reference's OWN `lidarnerf/nerf/network.py:NeRFNetwork` (nn.Linear stacks; its encoders resolve to this
    library through compat.install) and this library's `NeRFNetwork(use_ffmlp=False)` are the same function: identical
    state_dict keys, and on shared weights identical `density()` / `color()` outputs and identical dense `render()`."""
    root, kind = _ref_or_skip()
    from lidar_nerf_b200 import compat
    from lidar_nerf_b200.nerf.network import NeRFNetwork as Ours
    try:
        compat.install(networks=False)                  # keep the reference's network class, swap only the extensions
        import importlib
        ref_mod = importlib.import_module("lidarnerf.nerf.network")
        assert ref_mod.NeRFNetwork is not Ours
        kw = dict(encoding="hashgrid", desired_resolution=2048, log2_hashmap_size=15, num_layers=2, hidden_dim=64,
                  geo_feat_dim=15, bound=1, density_scale=1, min_near=0.05, density_thresh=10, bg_radius=-1)
        torch.manual_seed(0)
        ref = ref_mod.NeRFNetwork(**kw).cuda().eval()
        ours = Ours(use_ffmlp=False, **kw).cuda().eval()
        assert sorted(ref.state_dict().keys()) == sorted(ours.state_dict().keys())
        ours.load_state_dict(ref.state_dict())          # a reference checkpoint loads as is
        ours.encoder.embeddings.data.uniform_(-0.5, 0.5)
        ref.load_state_dict(ours.state_dict())
        g = torch.Generator().manual_seed(1)
        x = (torch.rand(4096, 3, generator=g) * 2 - 1).cuda()
        d = torch.randn(4096, 3, generator=g)
        d = (d / d.norm(dim=-1, keepdim=True)).cuda()
        with torch.no_grad():
            a, b = ref.density(x), ours.density(x)
            assert torch.equal(a["sigma"], b["sigma"]) and torch.equal(a["geo_feat"], b["geo_feat"])
            ca = ref.color(x, d, cal_lidar_color=True, geo_feat=a["geo_feat"])
            cb = ours.color(x, d, cal_lidar_color=True, geo_feat=b["geo_feat"])
            assert torch.equal(ca, cb)
            o = torch.zeros(256, 3).cuda()
            ra = ref.render(o[None], d[None, :256], cal_lidar_color=True, staged=False, perturb=False, num_steps=64,
                            upsample_steps=16)
            rb = ours.render(o[None], d[None, :256], cal_lidar_color=True, staged=False, perturb=False, num_steps=64,
                             upsample_steps=16, cuda_ray="dense")
        np.testing.assert_allclose(ra["depth_lidar"].cpu().numpy(), rb["depth_lidar"].cpu().numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(ra["image_lidar"].cpu().numpy(), rb["image_lidar"].cpu().numpy(), rtol=1e-4, atol=1e-6)
    finally:
        _cleanup_modules()
