"""Synthetic panoramic LiDAR sequence for benchmarks and tests (no dataset ships with the image).

Shape and conventions follow the reference's KITTI-360 loader:
  * range image H x W (64 x 1024 by default), intrinsics (fov_up, fov) = (2.0, 26.9) deg (kitti360_dataset.py:121);
  * rays as in `get_lidar_rays` (dataset/base_dataset.py:85-100): column i -> azimuth beta = -(i - W/2)/W * 2 pi,
    row j -> elevation alpha = (fov_up - j/H * fov) deg, dir = (cos a cos b, cos a sin b, sin a) rotated by the pose;
  * per-pixel ground truth rows (ray-drop mask, intensity, depth) with depth already multiplied by `scale`
    (kitti360_dataset.py:85-96), poses translated by -offset and scaled.
Scene: ground plane z = -1.73 m, 8 axis-aligned boxes and 4 vertical cylinders inside +-75 m, ray-cast analytically
from `n_frames` poses on a straight 1 m-spaced trajectory; depth valid in [1, 80] m; intensity = 0.5 + 0.5 cos(incidence)
quantised to 1/255; ray-drop ~ Bernoulli(0.08) plus everything beyond 80 m.  Everything is seeded.
"""
import math

import numpy as np
import torch


def lidar_directions(H, W, fov_up, fov, device, inds=None):
    """Sensor-frame unit directions for flat pixel indices (row-major j * W + i); all pixels if inds is None."""
    if inds is None:
        inds = torch.arange(H * W, device=device)
    j = torch.div(inds, W, rounding_mode="floor").float()
    i = (inds % W).float()
    beta = -(i - W / 2) / W * 2 * math.pi
    alpha = (fov_up - j / H * fov) / 180 * math.pi
    return torch.stack([torch.cos(alpha) * torch.cos(beta), torch.cos(alpha) * torch.sin(beta), torch.sin(alpha)], -1)


class SyntheticLidarSequence:
    def __init__(self, H=64, W=1024, n_frames=16, fov_up=2.0, fov=26.9, scale=1.0 / 92.7, seed=0, device="cpu",
                 max_range=80.0, min_range=1.0):
        self.H, self.W, self.n_frames = H, W, n_frames
        self.fov_up, self.fov = fov_up, fov
        self.scale = float(scale)
        self.device = torch.device(device)
        rng = np.random.default_rng(seed)
        dev = self.device

        # --- scene (metres) ---
        self.ground_z = -1.73
        centers = rng.uniform(-60, 60, size=(8, 2))
        centers[np.abs(centers[:, 1]) < 6, 1] += 12.0           # keep the driving corridor free
        sizes = rng.uniform(3, 12, size=(8, 3))
        self.box_min = torch.tensor(np.concatenate([centers - sizes[:, :2] / 2, np.full((8, 1), self.ground_z)], 1),
                                    dtype=torch.float32, device=dev)
        self.box_max = torch.tensor(np.concatenate([centers + sizes[:, :2] / 2, self.ground_z + sizes[:, 2:]], 1),
                                    dtype=torch.float32, device=dev)
        cyl = rng.uniform(-50, 50, size=(4, 2))
        cyl[np.abs(cyl[:, 1]) < 5, 1] -= 10.0
        self.cyl_c = torch.tensor(cyl, dtype=torch.float32, device=dev)
        self.cyl_r = torch.tensor(rng.uniform(0.3, 1.5, size=4), dtype=torch.float32, device=dev)
        self.cyl_h = torch.tensor(self.ground_z + rng.uniform(4, 12, size=4), dtype=torch.float32, device=dev)

        # --- trajectory: straight line along +x, 1 m spacing, small yaw wobble ---
        xs = (np.arange(n_frames) - (n_frames - 1) / 2.0) * 1.0
        self.offset = np.array([0.0, 0.0, 0.0], np.float32)     # trajectory centre
        poses = np.tile(np.eye(4, dtype=np.float32), (n_frames, 1, 1))
        yaw = 0.02 * np.sin(np.arange(n_frames) * 0.7)
        poses[:, 0, 0], poses[:, 0, 1] = np.cos(yaw), -np.sin(yaw)
        poses[:, 1, 0], poses[:, 1, 1] = np.sin(yaw), np.cos(yaw)
        poses[:, 0, 3] = xs
        self.poses_m = torch.tensor(poses, device=dev)           # metres
        self.poses = self.poses_m.clone()                        # scaled world units, as the reference feeds the model
        self.poses[:, :3, 3] = (self.poses_m[:, :3, 3] - torch.tensor(self.offset, device=dev)) * self.scale

        # --- ray-cast every frame ---
        gen = torch.Generator(device="cpu").manual_seed(seed + 1)
        dirs_s = lidar_directions(H, W, fov_up, fov, dev)        # [HW,3] sensor frame
        images = []
        for f in range(n_frames):
            R, t = self.poses_m[f, :3, :3], self.poses_m[f, :3, 3]
            d = dirs_s @ R.T
            depth, cosi = self._cast(t, d)
            valid = (depth >= min_range) & (depth <= max_range)
            drop = torch.rand(H * W, generator=gen).to(dev) < 0.08
            mask = (valid & ~drop).float()
            inten = torch.round((0.5 + 0.5 * cosi.clamp(0, 1)) * 255) / 255
            images.append(torch.stack([mask, inten * mask, depth.clamp(0, max_range) * self.scale * mask], -1))
        self.images = torch.stack(images)                        # [F, HW, 3] = (ray-drop mask, intensity, depth)

    # analytic ray casting; o [3], d [R,3] (metres) -> depth [R], cos(incidence) [R]
    def _cast(self, o, d):
        inf = torch.full((d.shape[0],), float("inf"), device=d.device)
        best, cosi = inf.clone(), torch.zeros_like(inf)
        # ground plane
        tz = (self.ground_z - o[2]) / d[:, 2]
        hit = (d[:, 2] < 0) & (tz > 0)
        tz = torch.where(hit, tz, inf)
        cosi = torch.where(tz < best, d[:, 2].abs(), cosi)
        best = torch.minimum(best, tz)
        # boxes (slab test)
        inv = 1.0 / d
        for b in range(self.box_min.shape[0]):
            t0 = (self.box_min[b] - o) * inv
            t1 = (self.box_max[b] - o) * inv
            tn, tf = torch.minimum(t0, t1), torch.maximum(t0, t1)
            tnear, axis = tn.max(-1)
            tfar = tf.min(-1).values
            hit = (tnear <= tfar) & (tnear > 0)
            tb = torch.where(hit, tnear, inf)
            ci = d.gather(1, axis[:, None]).squeeze(1).abs()
            cosi = torch.where(tb < best, ci, cosi)
            best = torch.minimum(best, tb)
        # vertical cylinders
        for c in range(self.cyl_c.shape[0]):
            oc = o[:2] - self.cyl_c[c]
            a = (d[:, :2] ** 2).sum(-1)
            bq = 2 * (d[:, :2] * oc).sum(-1)
            cq = (oc ** 2).sum() - self.cyl_r[c] ** 2
            disc = bq * bq - 4 * a * cq
            ok = disc > 0
            tcy = (-bq - torch.sqrt(disc.clamp(min=0))) / (2 * a)
            z = o[2] + tcy * d[:, 2]
            hit = ok & (tcy > 0) & (z >= self.ground_z) & (z <= self.cyl_h[c])
            tc = torch.where(hit, tcy, inf)
            p = o[:2] + tc[:, None].nan_to_num(posinf=0.0) * d[:, :2]
            nrm = (p - self.cyl_c[c]) / self.cyl_r[c]
            ci = (nrm * d[:, :2]).sum(-1).abs()
            cosi = torch.where(tc < best, ci, cosi)
            best = torch.minimum(best, tc)
        return best, cosi

    # ---- what the reference's collate does per step (kitti360_dataset.py:123-159) ------------------------------
    def sample_batch(self, n_rays, frame=None, generator=None, device=None):
        """-> rays_o [N,3], rays_d [N,3], gt [N,3] for `n_rays` random pixels of one frame."""
        dev = self.device if device is None else torch.device(device)
        if frame is None:
            frame = int(torch.randint(0, self.n_frames, (1,), generator=generator))
        inds = torch.randint(0, self.H * self.W, (n_rays,), generator=generator, device=generator.device if generator else "cpu").to(self.device)
        dirs = lidar_directions(self.H, self.W, self.fov_up, self.fov, self.device, inds)
        pose = self.poses[frame]
        rays_d = dirs @ pose[:3, :3].T
        rays_o = pose[:3, 3].expand_as(rays_d)
        gt = self.images[frame][inds]
        return rays_o.contiguous().to(dev), rays_d.contiguous().to(dev), gt.contiguous().to(dev)

    def surface_points(self, stride=1):
        """World-space (scaled) positions of all valid returns - the LiDAR occupancy prior."""
        pts = []
        dirs_s = lidar_directions(self.H, self.W, self.fov_up, self.fov, self.device)
        for f in range(0, self.n_frames, stride):
            img = self.images[f]
            m = img[:, 0] > 0
            d = dirs_s[m] @ self.poses[f, :3, :3].T
            pts.append(self.poses[f, :3, 3] + d * img[m, 2:3])
        return torch.cat(pts)
