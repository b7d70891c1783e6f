"""TEST INFRASTRUCTURE: run the reference's UNMODIFIED Python (Trainer, datasets, main_lidarnerf.py) on this library.

  * `reference_root()` - where the reference's Python is importable from: /root/reference in the build container, else
    the sourceless .pyc tree `oracle/build_ref.py pyref` byte-compiled from it into oracle/_ref/pyref (travels to the
    GPU box like the reference's CUDA extensions);
  * `install()`        - puts it on sys.path, registers stand-ins for the pip packages the reference imports at module
    level but this image lacks (imageio, lpips, mcubes, tensorboardX, trimesh, skimage, torch_ema, configargparse - none
    of them is on the hot path: logging, meshes, image metrics), then `lidar_nerf_b200.compat.install()`;
  * `write_kitti360(root, seq)` - a synthetic sequence in the on-disk format of `KITTI360Dataset`
    (dataset/kitti360_dataset.py:44-96: transforms_{seq}_{split}.json + [H,W,3] .npy range images).
"""
import argparse
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")


def reference_root():
    src = os.environ.get("LNB_REFERENCE_ROOT", "/root/reference")
    if os.path.isdir(os.path.join(src, "lidarnerf")):
        return src, "source"
    if os.path.isdir(os.path.join(PYREF, "lidarnerf")):
        return PYREF, "pyc"
    return None, None


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _absent(name):
    if name in sys.modules:
        return False
    try:
        __import__(name)
        return False
    except Exception:
        return True


class _ConfigArgParser(argparse.ArgumentParser):
    """The subset of configargparse the entry script uses: `is_config_file=True` options naming a `key = value` file
    whose entries become defaults (lists as `[a, b]`, booleans as `True`)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._cfg_dests = []

    def add_argument(self, *a, is_config_file=False, **k):
        act = super().add_argument(*a, **k)
        if is_config_file:
            self._cfg_dests.append(act.dest)
        return act

    def parse_args(self, args=None, namespace=None):
        args = list(sys.argv[1:] if args is None else args)
        pre, _ = super().parse_known_args(args)
        extra = []
        for dest in self._cfg_dests:
            path = getattr(pre, dest, None)
            if not path:
                continue
            if not os.path.exists(path) and path == self.get_default(dest):
                continue                      # a default config file that is not there is not an error
            for line in open(path):
                line = line.split("#")[0].strip()
                if not line or "=" not in line:
                    continue
                key, val = (x.strip() for x in line.split("=", 1))
                if val.startswith("["):
                    vals = [v.strip() for v in val.strip("[]").split(",") if v.strip()]
                    extra += [f"--{key}", *vals]
                elif val in ("True", "true"):
                    extra += [f"--{key}"]
                elif val in ("False", "false"):
                    continue
                else:
                    extra += [f"--{key}", val]
        return super().parse_args(extra + args, namespace)      # command line wins over the file


def _ssim(a, b, **_k):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    ma, mb, va, vb = a.mean(), b.mean(), a.var(), b.var()
    cov = ((a - ma) * (b - mb)).mean()
    return float(((2 * ma * mb + c1) * (2 * cov + c2)) / ((ma * ma + mb * mb + c1) * (va + vb + c2)))


def install():
    """-> (root, kind) of the reference Python, or (None, None) when it is not available on this box."""
    root, kind = reference_root()
    if root is None:
        return None, None
    if root not in sys.path:
        sys.path.insert(0, root)
    import torch

    if kind == "pyc" and not getattr(torch.jit.script, "_lnb_sourceless_ok", False):
        # nerf/utils.py:38-45 decorates two colour-space helpers (unused by the LiDAR path) with @torch.jit.script, which
        # needs source text: with the sourceless tree they stay plain Python functions
        real_script = torch.jit.script

        def script(obj, *a, **k):
            try:
                return real_script(obj, *a, **k)
            except OSError:
                return obj
        script._lnb_sourceless_ok = True
        torch.jit.script = script
    if _absent("imageio"):
        _mod("imageio", imwrite=lambda *a, **k: None, mimwrite=lambda *a, **k: None)
    if _absent("tensorboardX"):
        class SummaryWriter:
            def __init__(self, *a, **k):
                self.scalars = []

            def add_scalar(self, tag, value, step=None):
                self.scalars.append((tag, float(value), step))

            def close(self):
                pass
        _mod("tensorboardX", SummaryWriter=SummaryWriter)
    if _absent("trimesh"):
        class Trimesh:
            def __init__(self, vertices=None, faces=None, **k):
                self.vertices, self.faces = vertices, faces

            def export(self, path):
                np.save(path + ".npy", np.asarray(self.vertices))
        _mod("trimesh", Trimesh=Trimesh)
    if _absent("mcubes"):
        def marching_cubes(u, thresh):
            idx = np.argwhere(np.asarray(u) > thresh).astype(np.float64)
            return idx[:3 * (len(idx) // 3)], np.arange(3 * (len(idx) // 3)).reshape(-1, 3)
        _mod("mcubes", marching_cubes=marching_cubes)
    if _absent("lpips"):
        class LPIPS(torch.nn.Module):
            def __init__(self, net="alex", **k):
                super().__init__()

            def forward(self, a, b, normalize=False):
                return (a - b).abs().mean().reshape(1, 1, 1, 1)
        _mod("lpips", LPIPS=LPIPS)
    if _absent("skimage"):
        sk = _mod("skimage")
        sk.__path__ = []
        sk.metrics = _mod("skimage.metrics", structural_similarity=_ssim)
    if _absent("torch_ema"):
        class ExponentialMovingAverage:
            """shadow <- decay * shadow + (1 - decay) * param (the part of torch_ema the Trainer uses)."""

            def __init__(self, parameters, decay):
                self.params = [p for p in parameters if p.requires_grad]
                self.decay = decay
                self.shadow = [p.detach().clone() for p in self.params]
                self.backup = None

            @torch.no_grad()
            def update(self):
                for s, p in zip(self.shadow, self.params):
                    s.lerp_(p.detach(), 1 - self.decay)

            def store(self):
                self.backup = [p.detach().clone() for p in self.params]

            @torch.no_grad()
            def copy_to(self):
                for s, p in zip(self.shadow, self.params):
                    p.copy_(s)

            @torch.no_grad()
            def restore(self):
                for b, p in zip(self.backup, self.params):
                    p.copy_(b)
                self.backup = None

            def state_dict(self):
                return {"decay": self.decay, "shadow_params": self.shadow}

            def load_state_dict(self, sd):
                self.decay = sd["decay"]
                for s, v in zip(self.shadow, sd["shadow_params"]):
                    s.copy_(v)
        _mod("torch_ema", ExponentialMovingAverage=ExponentialMovingAverage)
    if _absent("configargparse"):
        _mod("configargparse", ArgumentParser=_ConfigArgParser)
    return root, kind


def write_kitti360(root, seq="1908", H=64, W=1024, n_train=4, n_val=1, n_test=1, seed=0):
    """Synthetic KITTI-360-format sequence (analytic scene of lidar_nerf_b200.data.synthetic) -> `root`.
    Returns (scale, offset) to pass as --scale / --offset."""
    import torch
    from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence
    n = n_train + n_val + n_test
    seq_obj = SyntheticLidarSequence(H=H, W=W, n_frames=n, seed=seed, device="cpu")
    os.makedirs(os.path.join(root, "train"), exist_ok=True)
    frames = []
    for f in range(n):
        img = seq_obj.images[f].reshape(H, W, 3).numpy()
        pc = np.zeros((H, W, 3), np.float32)
        pc[..., 1] = img[..., 1]                                    # intensity
        pc[..., 2] = img[..., 2] / seq_obj.scale * img[..., 0]      # depth in metres; 0 = dropped (kitti360_dataset.py:74-77)
        rel = os.path.join("train", f"{f:08d}.npy")
        np.save(os.path.join(root, rel), pc)
        frames.append({"lidar_file_path": rel, "lidar2world": seq_obj.poses_m[f].numpy().tolist()})
    splits = {"train": frames[:n_train], "val": frames[n_train:n_train + n_val], "test": frames[n_train + n_val:]}
    for split, fr in splits.items():
        with open(os.path.join(root, f"transforms_{seq}_{split}.json"), "w") as fh:
            json.dump({"w_lidar": W, "h_lidar": H, "aabb_scale": 2, "frames": fr}, fh)
    return seq_obj.scale, [float(x) for x in seq_obj.offset]
