"""Range image <-> point cloud on the GPU: the reference's `lidarnerf/convert.py` entry points (same names, argument
order and return values; convert.py:99-160,164-191,194-250) over `liblnb200.so`.

The reference takes and returns numpy arrays (it loops over the points in Python); these accept numpy arrays OR CUDA
tensors and answer in kind.  There is no CPU fallback: numpy inputs make a round trip through the GPU.
"""
import numpy as np
import torch

from ._lib import lib, check, u32, f32, vp

_DEV = "cuda"


def _to_dev(a):
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise RuntimeError("convert: tensors must live on a CUDA device (numpy arrays are copied there)")
        return a.to(torch.float32).contiguous(), True
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(_DEV), False


def _s():
    return vp(torch.cuda.current_stream().cuda_stream)


def lidar_to_pano_with_intensities(local_points_with_intensities, lidar_H, lidar_W, lidar_K, max_depth=80):
    """(N, 4) points (x, y, z, intensity) in the sensor frame -> (pano [H, W], intensities [H, W])."""
    pts, is_t = _to_dev(local_points_with_intensities)
    if pts.dim() != 2 or pts.shape[1] < 3:
        raise RuntimeError("points must be [N, >= 3]")
    H, W = int(lidar_H), int(lidar_W)
    pano = torch.empty(H, W, dtype=torch.float32, device=pts.device)
    inten = torch.empty(H, W, dtype=torch.float32, device=pts.device)
    ws = torch.empty(int(lib.lnb_lidar_to_pano_workspace_bytes(u32(H), u32(W))), dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        check(lib.lnb_lidar_to_pano(vp(pts.data_ptr()), u32(pts.shape[1]), u32(pts.shape[0]), u32(H), u32(W),
                                    f32(lidar_K[0]), f32(lidar_K[1]), f32(max_depth), vp(pano.data_ptr()),
                                    vp(inten.data_ptr()), vp(ws.data_ptr()), _s()), "lidar_to_pano")
    if is_t:
        return pano, inten
    return pano.cpu().numpy(), inten.cpu().numpy()


def lidar_to_pano(local_points, lidar_H, lidar_W, lidar_K, max_dpeth=80):
    """(N, 3) points -> pano [H, W]  (the reference's keyword really is spelled `max_dpeth`, convert.py:165)."""
    pts, is_t = _to_dev(local_points)
    pano, _ = lidar_to_pano_with_intensities(pts[:, :3].contiguous(), lidar_H, lidar_W, lidar_K, max_depth=max_dpeth)
    return pano if is_t else pano.cpu().numpy()


def pano_to_lidar_with_intensities(pano, intensities, lidar_K):
    """pano [H, W], intensities [H, W] -> (n, 4) points of the non-empty pixels, row-major pixel order."""
    pa, is_t = _to_dev(pano)
    it, _ = _to_dev(intensities)
    H, W = pa.shape
    out = torch.empty(H * W, 4, dtype=torch.float32, device=pa.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=pa.device)
    ws = torch.empty(int(lib.lnb_pano_to_lidar_workspace_bytes(u32(H), u32(W))), dtype=torch.uint8, device=pa.device)
    with torch.cuda.device(pa.device):
        check(lib.lnb_pano_to_lidar(vp(pa.data_ptr()), vp(it.data_ptr()), u32(H), u32(W), f32(lidar_K[0]),
                                    f32(lidar_K[1]), vp(out.data_ptr()), vp(cnt.data_ptr()), vp(ws.data_ptr()), _s()),
              "pano_to_lidar")
    pts = out[:int(cnt.item())]
    return pts if is_t else pts.cpu().numpy()


def pano_to_lidar(pano, lidar_K):
    """pano [H, W] -> (n, 3) points."""
    pa, is_t = _to_dev(pano)
    pts = pano_to_lidar_with_intensities(pa, torch.zeros_like(pa), lidar_K)[:, :3]
    return pts if is_t else pts.cpu().numpy()


__all__ = ["lidar_to_pano_with_intensities", "lidar_to_pano", "pano_to_lidar_with_intensities", "pano_to_lidar"]
