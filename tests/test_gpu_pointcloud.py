"""GPU parity tests of the evaluation-side callers of the hot path: Chamfer nearest neighbours,
range image <-> point cloud, and the grid encoder's total-variation gradient - against the CPU oracle, the fixtures
generated from the reference's Python (tests/golden/ref_py_convert.npz) and, when built, the reference's own
chamfer3D CUDA extension (oracle/_ref).

Tolerances: Chamfer distances and indices bit-exact (same expression, same FMA contraction, first minimum wins);
range-image pixels: identical except where the device's atan2f differs from libm's by an ulp exactly on a rounding
boundary (<= 0.2 % of the pixels allowed, 0 observed); reconstructed points rtol 1e-5 (sincosf vs libm);
total variation rtol 1e-4.
"""
import os

import numpy as np
import pytest
import torch

import cases
import refcuda

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("B,N,M", [(2, 3000, 2500), (1, 1, 1), (1, 129, 2049), (3, 64, 5), (1, 5000, 4099)])
def test_chamfer_forward_bit_exact_vs_oracle(orc, B, N, M):
    from lidar_nerf_b200.extern import chamfer_3DDist
    rng = np.random.default_rng(B * 1000 + N + M)
    a = rng.normal(size=(B, N, 3)).astype(np.float32)
    b = rng.normal(size=(B, M, 3)).astype(np.float32)
    if M > 10:
        b[0, 7] = b[0, 3]            # exact duplicates: the first index wins
        a[0, 0] = b[0, 3]
    d1, d2, i1, i2 = chamfer_3DDist()(T(a), T(b))
    o1, o2, j1, j2 = orc.chamfer_forward(a, b)
    np.testing.assert_array_equal(d1.cpu().numpy(), o1)
    np.testing.assert_array_equal(d2.cpu().numpy(), o2)
    np.testing.assert_array_equal(i1.cpu().numpy(), j1)
    np.testing.assert_array_equal(i2.cpu().numpy(), j2)
    assert i1.dtype == torch.int32 and d1.shape == (B, N) and d2.shape == (B, M)
    ref = refcuda.load("chamfer_3D")
    if ref is not None:
        r1, r2 = torch.zeros(B, N, device=DEV), torch.zeros(B, M, device=DEV)
        k1 = torch.zeros(B, N, dtype=torch.int32, device=DEV)
        k2 = torch.zeros(B, M, dtype=torch.int32, device=DEV)
        ref.forward(T(a), T(b), r1, r2, k1, k2)
        torch.cuda.synchronize()
        assert torch.equal(r1, d1) and torch.equal(r2, d2) and torch.equal(k1, i1) and torch.equal(k2, i2)


def test_chamfer_backward_matches_autograd_of_the_matched_pairs():
    from lidar_nerf_b200.extern import chamfer_3DDist, fscore
    rng = np.random.default_rng(1)
    a = T(rng.normal(size=(2, 400, 3)).astype(np.float32)).requires_grad_(True)
    b = T(rng.normal(size=(2, 300, 3)).astype(np.float32)).requires_grad_(True)
    d1, d2, i1, i2 = chamfer_3DDist()(a, b)
    w1, w2 = torch.rand_like(d1), torch.rand_like(d2)
    ((d1 * w1).sum() + (d2 * w2).sum()).backward()
    ga, gb = a.grad.clone(), b.grad.clone()
    a2, b2 = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    e1 = ((a2 - torch.gather(b2, 1, i1.long()[..., None].expand(-1, -1, 3))) ** 2).sum(-1)
    e2 = ((b2 - torch.gather(a2, 1, i2.long()[..., None].expand(-1, -1, 3))) ** 2).sum(-1)
    ((e1 * w1).sum() + (e2 * w2).sum()).backward()
    np.testing.assert_allclose(ga.cpu().numpy(), a2.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gb.cpu().numpy(), b2.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    f, p, r = fscore(d1.detach(), d2.detach(), threshold=0.05)
    assert f.shape == (2,) and ((0 <= f) & (f <= 1)).all()
    np.testing.assert_allclose(p.cpu().numpy(), (d1.detach() < 0.05).float().mean(1).cpu().numpy())


def test_range_image_conversion_vs_oracle_and_reference_python(orc, golden_dir):
    from lidar_nerf_b200 import convert
    z = np.load(os.path.join(golden_dir, "ref_py_convert.npz"))
    H, W, K = int(z["H"]), int(z["W"]), tuple(float(v) for v in z["K"])
    pano, inten = convert.lidar_to_pano_with_intensities(z["points"], H, W, K, max_depth=80)       # numpy in, numpy out
    assert isinstance(pano, np.ndarray) and pano.shape == (H, W)
    ref_pano = z["pano"].astype(np.float32)
    bad = (pano != ref_pano) | (inten != z["intensities"].astype(np.float32))
    assert bad.mean() <= 2e-3, f"{bad.sum()} of {H * W} pixels differ from the reference's convert.py"
    pano3 = convert.lidar_to_pano(z["points"][:, :3], H, W, K)
    np.testing.assert_array_equal(pano3, pano)
    # tensors in, tensors out; same order and values as the reference's pano_to_lidar_with_intensities
    pts = convert.pano_to_lidar_with_intensities(T(ref_pano), T(z["intensities"].astype(np.float32)), K)
    assert pts.is_cuda and tuple(pts.shape) == z["back"].shape
    np.testing.assert_allclose(pts.cpu().numpy(), z["back"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(convert.pano_to_lidar(ref_pano, K), z["back"][:, :3], rtol=1e-5, atol=2e-5)
    # empty image / no points
    e = convert.pano_to_lidar_with_intensities(np.zeros((H, W), np.float32), np.zeros((H, W), np.float32), K)
    assert e.shape == (0, 4)
    p0, i0 = convert.lidar_to_pano_with_intensities(np.zeros((0, 4), np.float32), H, W, K)
    assert (p0 == 0).all() and (i0 == 0).all()


def test_range_image_round_trip_at_full_size(orc):
    """64 x 1024 pano of the synthetic sequence (BASELINE size): image -> points -> image is the identity except on the
    rows whose beam centres sit on a rounding boundary of the row index, and agrees with the oracle everywhere."""
    from lidar_nerf_b200 import convert
    from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence
    seq = SyntheticLidarSequence(n_frames=1, device=DEV)
    img = seq.images[0].reshape(seq.H, seq.W, 3)
    pano = (img[..., 2] / seq.scale).contiguous()                 # metres, 0 = dropped
    inten = img[..., 1].contiguous()
    K = (seq.fov_up, seq.fov)
    pts = convert.pano_to_lidar_with_intensities(pano, inten, K)
    assert pts.shape[0] == int((pano != 0).sum()) > 30000
    o_pts = orc.pano_to_lidar_with_intensities(pano.cpu().numpy(), inten.cpu().numpy(), K)
    np.testing.assert_allclose(pts.cpu().numpy(), o_pts, rtol=1e-5, atol=2e-5)
    pano2, inten2 = convert.lidar_to_pano_with_intensities(pts, seq.H, seq.W, K, max_depth=81)
    o_pano2, o_inten2 = orc.lidar_to_pano_with_intensities(pts.cpu().numpy(), seq.H, seq.W, K, 81)
    assert (pano2.cpu().numpy() != o_pano2).mean() <= 2e-3
    ok = torch.isclose(pano2, pano, rtol=1e-5)
    assert float(ok.float().mean()) > 0.97


def test_grad_total_variation_vs_oracle(orc):
    from lidar_nerf_b200.gridencoder import GridEncoder
    torch.manual_seed(0)
    enc = GridEncoder(input_dim=3, num_levels=8, level_dim=2, base_resolution=16, log2_hashmap_size=14,
                      desired_resolution=512).to(DEV)
    enc.embeddings.data.uniform_(-1, 1)
    x = torch.rand(4000, 3, device=DEV) * 2 - 1
    x[0] = 1.5                                                    # out of range: contributes nothing
    enc.embeddings.grad = torch.zeros_like(enc.embeddings)
    enc.grad_total_variation(weight=1e-3, inputs=x, bound=1)
    L = 8
    ls = (torch.exp2(torch.arange(L, device=DEV, dtype=torch.float32) * torch.tensor(float(np.log2(enc.per_level_scale)), device=DEV)) * 16.0 - 1.0).cpu().numpy()
    x01 = ((x + 1) / 2).cpu().numpy()
    want = orc.grad_total_variation(x01, enc.embeddings.detach().cpu().numpy(), np.zeros(tuple(enc.embeddings.shape), np.float32),
                                    enc.offsets.cpu().numpy(), 1e-3, enc.per_level_scale, 16, level_scales=ls)
    got = enc.embeddings.grad.cpu().numpy()
    assert np.abs(want).sum() > 0
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-7)
    enc.embeddings.grad = None
    with pytest.raises(ValueError):
        enc.grad_total_variation()
