#!/usr/bin/env python3
"""Generate tests/golden/ref_py_*.npz by importing the REFERENCE's Python code (read-only checkout at
/root/reference) on CPU.  Run in the build container only (the checkout does not exist on the GPU box);
the resulting small fixtures are committed and pin the CPU oracle / host logic:

  ref_py_freq.npz        lidarnerf/encoding.py:6-47   pure-torch FreqEncoder (exact sin/cos) fwd + autograd bwd
  ref_py_trunc_exp.npz   lidarnerf/activation.py:6-20 trunc_exp fwd/bwd
  ref_py_grid_offsets.npz lidarnerf/gridencoder/grid.py:157-195 level table of GridEncoder.__init__
  ref_py_lidar_rays.npz  lidarnerf/dataset/base_dataset.py:16-105 get_lidar_rays (full image, no sampling)
  ref_py_run.npz         lidarnerf/nerf/renderer.py:99-298 NeRFRenderer.run (LiDAR mode, perturb=False, eval-mode
                         deterministic PDF up-sampling) with an analytic density/colour field
  ref_py_sample_pdf.npz  lidarnerf/nerf/renderer.py:10-46 sample_pdf(det=True)
  ref_py_convert.npz     lidarnerf/convert.py:99-160,194-235 lidar_to_pano_with_intensities / pano_to_lidar_with_intensities
  ref_py_patch_loss.npz  lidarnerf/nerf/utils.py:697-876 Trainer.train_step: the LiDAR loss INCLUDING the patch depth-gradient
                         term (grad_loss = True, patches of 2 x 8 rays) and its autograd gradient w.r.t. the rendered
                         depth / image, on given render outputs

Usage: python tests/golden/make_golden_cpu.py
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
REF = os.environ.get("LNB_REFERENCE_ROOT", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
# the reference imports trimesh at module import time (renderer.py:2, base_dataset.py:3); it is not installed
# and not needed for the functions exercised here
sys.modules.setdefault("trimesh", types.ModuleType("trimesh"))
# grid.py does `import _gridencoder as _backend` at import; only __init__ (pure numpy) is exercised
sys.modules.setdefault("_gridencoder", types.ModuleType("_gridencoder"))

torch.manual_seed(0)
rng = np.random.default_rng(0)


def save(name, **arrays):
    np.savez_compressed(os.path.join(OUT, name), **{k: np.asarray(v) for k, v in arrays.items()})
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrays.items()})


# ---- range image <-> point cloud (lidarnerf/convert.py:99-160,194-235) ---------------------------------------
def make_convert():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_convert", os.path.join(REF, "lidarnerf", "convert.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    r = np.random.default_rng(5)
    H, W, K = 32, 256, (2.0, 26.9)
    n = 6000
    # points on a few shells + noise, some beyond max_depth, some outside the vertical field of view
    d = r.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:, 2] = d[:, 2] * 0.35 - 0.15
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rad = r.choice([5.0, 12.0, 30.0, 70.0, 95.0], size=n) * r.uniform(0.9, 1.1, size=n)
    pts = np.concatenate([d * rad[:, None], r.uniform(0, 1, size=(n, 1))], axis=1).astype(np.float32)
    pano, inten = conv.lidar_to_pano_with_intensities(pts, H, W, K, max_depth=80)
    back = conv.pano_to_lidar_with_intensities(pano.astype(np.float32), inten.astype(np.float32), K)
    save("ref_py_convert.npz", points=pts, H=H, W=W, K=np.array(K, np.float32), pano=pano, intensities=inten,
         back=back, numpy_version=np.__version__)


# ---- LiDAR loss with the patch depth-gradient term (lidarnerf/nerf/utils.py:697-876) -------------------------
def make_patch_loss():
    """Calls the reference's own Trainer.train_step on a stand-in `self` whose model returns given render outputs."""
    import argparse
    for name in ("imageio", "lpips", "mcubes", "tensorboardX", "torch_ema", "skimage", "skimage.metrics", "extern",
                 "extern.chamfer3D", "extern.chamfer3D.dist_chamfer_3D", "extern.fscore"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["torch_ema"].ExponentialMovingAverage = object
    sys.modules["skimage.metrics"].structural_similarity = None
    sys.modules["extern.chamfer3D.dist_chamfer_3D"].chamfer_3DDist = object
    sys.modules["extern.fscore"].fscore = None
    from lidarnerf.nerf.utils import Trainer
    r = np.random.default_rng(11)
    px, py, n_patch = 2, 8, 24
    N = n_patch * px * py
    scale = 0.010784853507573345                       # configs/kitti360_1908.txt:11
    # smooth ground-truth depth inside a patch (so that |dG| < 0.01 m happens) with a few jumps and dropped rays
    base = r.uniform(5, 60, size=(n_patch, 1, 1))
    slope = r.uniform(-0.004, 0.004, size=(n_patch, 1, 1)) * np.arange(py)[None, None, :]
    jump = (r.random((n_patch, px, py)) < 0.15) * r.uniform(0.5, 3.0, size=(n_patch, px, py))
    gt_depth_m = base + slope + jump + r.normal(scale=0.002, size=(n_patch, px, py))
    mask = (r.random((n_patch, px, py)) > 0.12).astype(np.float32)
    gt = np.stack([mask, r.uniform(0, 1, size=mask.shape) * mask, gt_depth_m * scale * mask], -1).reshape(N, 3).astype(np.float32)
    depth = (gt_depth_m * scale + r.normal(scale=0.004 * scale * 20, size=gt_depth_m.shape)).reshape(N).astype(np.float32)
    image = r.uniform(0.02, 0.98, size=(N, 2)).astype(np.float32)
    depth_t = torch.tensor(depth, requires_grad=True)
    image_t = torch.tensor(image, requires_grad=True)

    class Model:
        def render(self, ro, rd, **kw):
            return {"depth_lidar": depth_t[None], "image_lidar": image_t[None]}
    opt = argparse.Namespace(enable_lidar=True, patch_size=1, alpha_d=1e3, alpha_r=1.0, alpha_i=10.0, alpha_grad=100.0,
                             patch_size_lidar=[px, py], scale=scale, sobel_grad=False, grad_norm_smooth=False,
                             spatial_smooth=False, tv_loss=False, grad_loss=True, depth_grad_loss="l1")
    crit = {"depth": torch.nn.L1Loss(reduction="none"), "raydrop": torch.nn.MSELoss(reduction="none"),
            "intensity": torch.nn.MSELoss(reduction="none"), "grad": torch.nn.L1Loss(reduction="none")}
    me = types.SimpleNamespace(opt=opt, model=Model(), criterion=crit, device=torch.device("cpu"))
    data = {"rays_o_lidar": torch.zeros(1, N, 3), "rays_d_lidar": torch.zeros(1, N, 3), "images_lidar": torch.tensor(gt)[None]}
    loss = Trainer.train_step(me, data)[-1]
    loss.backward()
    opt.grad_loss = False                               # the per-ray part alone, for reference
    depth_t2, image_t2 = depth_t.detach().clone().requires_grad_(True), image_t.detach().clone().requires_grad_(True)
    Model.render = lambda self, ro, rd, **kw: {"depth_lidar": depth_t2[None], "image_lidar": image_t2[None]}
    loss0 = Trainer.train_step(me, data)[-1]
    save("ref_py_patch_loss.npz", depth=depth, image=image, gt=gt, scale=scale, patch=np.array([px, py]), alpha_d=1e3,
         alpha_r=1.0, alpha_i=10.0, alpha_grad=100.0, loss=float(loss), loss_without_grad_term=float(loss0),
         g_depth=depth_t.grad.numpy(), g_image=image_t.grad.numpy())


if sys.argv[1:] == ["patch_loss"]:   # `python tests/golden/make_golden_cpu.py patch_loss`: only this fixture
    make_patch_loss()
    sys.exit(0)

make_convert()
if sys.argv[1:] == ["convert"]:      # `python tests/golden/make_golden_cpu.py convert`: only this fixture
    sys.exit(0)

# ---- FreqEncoder -------------------------------------------------------------------------------------------
from lidarnerf.encoding import FreqEncoder  # noqa: E402

for deg in (4, 12):
    enc = FreqEncoder(input_dim=3, max_freq_log2=deg - 1, N_freqs=deg, log_sampling=True)
    x = torch.tensor(rng.uniform(-1, 1, size=(257, 3)).astype(np.float32), requires_grad=True)
    y = enc(x)
    g = torch.tensor(rng.normal(size=tuple(y.shape)).astype(np.float32))
    (gx,) = torch.autograd.grad(y, x, g)
    save(f"ref_py_freq_deg{deg}.npz", x=x.detach().numpy(), y=y.detach().numpy(), g=g.numpy(), gx=gx.numpy())

# ---- trunc_exp ---------------------------------------------------------------------------------------------
from lidarnerf.activation import trunc_exp  # noqa: E402

x = torch.tensor(np.concatenate([rng.uniform(-20, 20, 200), [-15, 15, 0, 16, -16]]).astype(np.float32),
                 requires_grad=True)
y = trunc_exp(x)
g = torch.tensor(rng.normal(size=tuple(y.shape)).astype(np.float32))
(gx,) = torch.autograd.grad(y, x, g)
save("ref_py_trunc_exp.npz", x=x.detach().numpy(), y=y.detach().numpy(), g=g.numpy(), gx=gx.numpy())

# ---- GridEncoder level table -------------------------------------------------------------------------------
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_grid", os.path.join(REF, "lidarnerf/gridencoder/grid.py"))
ref_grid = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_grid)
cases = []
for kw in (dict(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=32768),
           dict(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048),
           dict(input_dim=2, num_levels=4, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048),
           dict(input_dim=3, num_levels=8, level_dim=4, per_level_scale=2, base_resolution=8, log2_hashmap_size=14,
                align_corners=True)):
    ge = ref_grid.GridEncoder(**kw)
    cases.append((kw, ge.offsets.numpy(), float(ge.per_level_scale)))
save("ref_py_grid_offsets.npz",
     **{f"offsets{i}": c[1] for i, c in enumerate(cases)},
     **{f"scale{i}": np.float64(c[2]) for i, c in enumerate(cases)},
     kwargs=np.array([repr(c[0]) for c in cases]))

# ---- get_lidar_rays ----------------------------------------------------------------------------------------
from lidarnerf.dataset.base_dataset import get_lidar_rays  # noqa: E402

ang = 0.3
pose = np.eye(4, dtype=np.float32)
pose[:3, :3] = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], np.float32)
pose[:3, 3] = [0.1, -0.2, 0.05]
res = get_lidar_rays(torch.tensor(pose)[None], [2.0, 26.9], 16, 64, -1)
save("ref_py_lidar_rays.npz", pose=pose, intrinsics=np.array([2.0, 26.9], np.float32), H=16, W=64,
     rays_o=res["rays_o"][0].numpy(), rays_d=res["rays_d"][0].numpy())

# ---- sample_pdf + NeRFRenderer.run -------------------------------------------------------------------------
from lidarnerf.nerf.renderer import NeRFRenderer, sample_pdf  # noqa: E402

bins = torch.tensor(np.sort(rng.uniform(0.1, 2.0, size=(7, 33)).astype(np.float32), axis=1))
w = torch.tensor(rng.uniform(0, 1, size=(7, 32)).astype(np.float32))
save("ref_py_sample_pdf.npz", bins=bins.numpy(), weights=w.numpy(), samples=sample_pdf(bins, w, 16, det=True).numpy())


class AnalyticField(NeRFRenderer):
    """density = smooth shell around a sphere of radius 0.4; colour = simple functions of position/direction."""

    def __init__(self):
        super().__init__(bound=1, min_near_lidar=0.01)
        self.out_color_dim, self.out_lidar_color_dim = 3, 2

    def density(self, x):
        r = x.norm(dim=-1)
        sigma = 40.0 * torch.exp(-((r - 0.4) / 0.05) ** 2)
        return {"sigma": sigma, "geo_feat": x[:, :1] * 0.5 + 0.5}

    def color(self, x, d, cal_lidar_color=False, mask=None, geo_feat=None, **kw):
        c = torch.stack([torch.sigmoid(3 * x[:, 0] + d[:, 2]), torch.sigmoid(geo_feat[:, 0] - d[:, 0])], -1)
        if mask is not None:
            c = c * mask[:, None]
        return c


field = AnalyticField().eval()
dirs = rng.normal(size=(37, 3)).astype(np.float32)
dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
orig = np.tile(np.array([[0.02, -0.01, 0.03]], np.float32), (37, 1))
with torch.no_grad():
    out = field.render(torch.tensor(orig)[None], torch.tensor(dirs)[None], cal_lidar_color=True, staged=False,
                       perturb=False, num_steps=48, upsample_steps=16)
save("ref_py_run.npz", rays_o=orig, rays_d=dirs, num_steps=48, upsample_steps=16, min_near_lidar=0.01,
     depth=out["depth_lidar"][0].numpy(), image=out["image_lidar"][0].numpy(),
     weights_sum=out["weights_sum_lidar"].numpy())

make_patch_loss()
