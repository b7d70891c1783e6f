"""CPU tests of the host-side mirror: renderer.run against the reference-generated fixture, ray generation and the
synthetic sequence, level tables, and the data-parallel gradient exchange on 2 gloo ranks."""
import os

import numpy as np
import pytest
import torch


def test_renderer_run_matches_reference_python(golden_dir):
    from lidar_nerf_b200.nerf.renderer import NeRFRenderer
    g = np.load(os.path.join(golden_dir, "ref_py_run.npz"))

    class AnalyticField(NeRFRenderer):   # same analytic field as tests/golden/make_golden_cpu.py
        def __init__(self):
            super().__init__(bound=1, min_near_lidar=0.01)
            self.out_color_dim, self.out_lidar_color_dim = 3, 2

        def density(self, x):
            r = x.norm(dim=-1)
            return {"sigma": 40.0 * torch.exp(-((r - 0.4) / 0.05) ** 2), "geo_feat": x[:, :1] * 0.5 + 0.5}

        def color(self, x, d, cal_lidar_color=False, mask=None, geo_feat=None, **kw):
            c = torch.stack([torch.sigmoid(3 * x[:, 0] + d[:, 2]), torch.sigmoid(geo_feat[:, 0] - d[:, 0])], -1)
            return c * mask[:, None] if mask is not None else c

    f = AnalyticField().eval()
    with torch.no_grad():
        out = f.render(torch.tensor(g["rays_o"])[None], torch.tensor(g["rays_d"])[None], cal_lidar_color=True,
                       staged=False, perturb=False, num_steps=int(g["num_steps"]), upsample_steps=int(g["upsample_steps"]))
    np.testing.assert_allclose(out["depth_lidar"][0].numpy(), g["depth"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out["image_lidar"][0].numpy(), g["image"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out["weights_sum_lidar"].numpy(), g["weights_sum"], rtol=1e-4, atol=1e-6)
    # staged rendering chunks the rays and must give the same answer
    with torch.no_grad():
        st = f.render(torch.tensor(g["rays_o"])[None], torch.tensor(g["rays_d"])[None], cal_lidar_color=True, staged=True,
                      max_ray_batch=16, perturb=False, num_steps=int(g["num_steps"]),
                      upsample_steps=int(g["upsample_steps"]))
    np.testing.assert_allclose(st["depth_lidar"][0].numpy(), g["depth"], rtol=1e-4, atol=1e-6)


def test_sample_pdf_matches_reference_python(golden_dir):
    from lidar_nerf_b200.nerf.renderer import sample_pdf
    g = np.load(os.path.join(golden_dir, "ref_py_sample_pdf.npz"))
    got = sample_pdf(torch.tensor(g["bins"]), torch.tensor(g["weights"]), 16, det=True).numpy()
    np.testing.assert_allclose(got, g["samples"], rtol=1e-5, atol=1e-6)


def test_lidar_ray_generation_matches_reference_python(golden_dir):
    from lidar_nerf_b200.data.synthetic import lidar_directions
    g = np.load(os.path.join(golden_dir, "ref_py_lidar_rays.npz"))
    H, W = int(g["H"]), int(g["W"])
    d = lidar_directions(H, W, float(g["intrinsics"][0]), float(g["intrinsics"][1]), "cpu")
    pose = torch.tensor(g["pose"])
    np.testing.assert_allclose((d @ pose[:3, :3].T).numpy(), g["rays_d"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(pose[:3, 3].expand(H * W, 3).numpy(), g["rays_o"], rtol=0, atol=0)


def test_level_table_matches_reference_python(golden_dir):
    import ast
    from lidar_nerf_b200.gridencoder import level_offsets
    g = np.load(os.path.join(golden_dir, "ref_py_grid_offsets.npz"))
    for i, kw in enumerate(g["kwargs"]):
        kw = ast.literal_eval(str(kw))
        off = level_offsets(kw["input_dim"], kw["num_levels"], kw["base_resolution"], float(g[f"scale{i}"]),
                            kw["log2_hashmap_size"], kw.get("align_corners", False))
        np.testing.assert_array_equal(off, g[f"offsets{i}"])


def test_synthetic_sequence_is_consistent():
    from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence
    s = SyntheticLidarSequence(H=16, W=128, n_frames=3, seed=1)
    assert s.images.shape == (3, 16 * 128, 3)
    m = s.images[..., 0]
    assert 0.5 < float(m.mean()) < 0.99
    depth_m = s.images[..., 2][m > 0] / s.scale
    assert float(depth_m.min()) >= 1.0 - 1e-4 and float(depth_m.max()) <= 80.0 + 1e-3
    assert (s.images[..., 1:][m == 0] == 0).all(), "dropped rays carry no intensity/depth"
    ro, rd, gt = s.sample_batch(64, frame=1, generator=torch.Generator().manual_seed(0))
    np.testing.assert_allclose(rd.norm(dim=-1).numpy(), 1.0, rtol=1e-5)
    pts = s.surface_points()
    assert pts.abs().max() <= 1.0, "the scaled scene must fit bound = 1"
    # a surface point re-projects onto its own ray
    s2 = SyntheticLidarSequence(H=16, W=128, n_frames=3, seed=1)
    assert torch.equal(s.images, s2.images), "seeded generation must be reproducible"


def test_field_config_matches_survey_numbers():
    from lidar_nerf_b200.nerf.engine import FieldConfig
    from lidar_nerf_b200.gridencoder import level_offsets
    c = FieldConfig()
    pls = float(np.exp2(np.log2(c.desired_resolution / c.base_resolution) / (c.num_levels - 1)))
    off = level_offsets(3, c.num_levels, c.base_resolution, pls, c.log2_hashmap_size, False)
    assert int(off[-1]) == 6837544 and c.head_in_dim == 96 and c.cascade == 1


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lidar_nerf_b200.nerf import dp
    torch.manual_seed(rank)
    g = torch.randn(1000)
    ref = g.clone()
    dp.allreduce_gradient_(g)
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, ref)
    assert torch.allclose(g, sum(gathered))
    assert dp.grad_scale(128.0) == pytest.approx(1.0 / (128.0 * world))
    lo, hi = dp.shard_bounds(1003, rank, world)
    sizes = [dp.shard_bounds(1003, r, world) for r in range(world)]
    assert sizes[0][0] == 0 and sizes[-1][1] == 1003 and all(a[1] == b[0] for a, b in zip(sizes[:-1], sizes[1:]))
    assert dp.rank_seed(7) == 7 + rank
    # sharded exchange: reduce-scatter -> per-shard update -> all-gather must equal all-reduce -> full update
    n = 1003
    ex = dp.ShardedExchange(n)
    assert ex.n_padded % (world * 8) == 0 and ex.n_padded >= n and ex.hi - ex.lo == ex.shard
    torch.manual_seed(100 + rank)
    grad = torch.zeros(ex.n_padded)
    grad[:n] = torch.randn(n)
    params = torch.arange(ex.n_padded, dtype=torch.float32)          # identical on every rank
    full = grad.clone()
    dp.allreduce_gradient_(full)
    want = (params - 0.1 * full).to(torch.float16)
    g_shard = torch.zeros(ex.shard)
    ex.reduce_scatter(grad, g_shard)
    assert torch.allclose(g_shard, full[ex.lo:ex.hi])
    p_shard = (params[ex.lo:ex.hi] - 0.1 * g_shard).to(torch.float16)
    got = torch.zeros(ex.n_padded, dtype=torch.float16)
    ex.all_gather(got, p_shard)
    assert torch.equal(got, want)
    # loops with collectives inside and a rank-local exit test: nobody goes on unless everybody does
    assert dp.all_ranks_agree(True) is True
    assert dp.all_ranks_agree(rank == 0) is False
    assert dp.all_ranks_agree(False) is False
    steps = 0
    while dp.all_ranks_agree(steps < 3 + 2 * rank):         # rank 1 would like two more iterations
        dp.allreduce_gradient_(torch.ones(4))               # (the collective a mismatched loop would hang in)
        steps += 1
    assert steps == 3
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


def test_data_parallel_exchange_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores) runs without a GPU and prints ONE JSON line with
    the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-rays", "32"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_bench_reference_arm_names_the_product_config_and_never_maps_the_product_library():
    """The two arms must describe the SAME workload (`config` identical), and the reference arm must not load
    liblnb200.so - the driver records which native libraries each arm mapped."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-rays', '32'];"
            "import bench; rc = bench.main(); maps = open('/proc/self/maps').read();"
            "print('MAPS', json.dumps({'lnb': 'liblnb200' in maps, 'oracle': 'lnb_oracle' in maps or 'oracle' in maps, 'rc': rc,"
            " 'cfg': bench.bench_config(2)}))")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    info = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("MAPS ")][0][5:])
    assert info["rc"] in (0, None) and info["lnb"] is False and info["oracle"] is True
    assert line["config"] == info["cfg"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_bench_product_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: the product arm must not print a result line when there is no CUDA device."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_network_tcnn_module_path_without_tinycudann():
    """`main_lidarnerf.py --tcnn` imports lidarnerf.nerf.network_tcnn.NeRFNetwork with the tcnn keyword set
    (main_lidarnerf.py:289-308); this library serves that module path without tinycudann."""
    import ast
    import inspect
    import sys
    from lidar_nerf_b200 import compat
    from lidar_nerf_b200.nerf import network_tcnn
    ours = inspect.signature(network_tcnn.NeRFNetwork.__init__).parameters
    ref_file = "/root/reference/lidarnerf/nerf/network_tcnn.py"
    if os.path.exists(ref_file):       # build container only: every constructor keyword of the reference is accepted
        tree = ast.parse(open(ref_file).read())
        init = next(n for c in tree.body if isinstance(c, ast.ClassDef) and c.name == "NeRFNetwork"
                    for n in c.body if isinstance(n, ast.FunctionDef) and n.name == "__init__")
        for a, d in zip(init.args.args[1:], init.args.defaults):
            assert a.arg in ours, a.arg
            assert ours[a.arg].default == ast.literal_eval(d), a.arg
    net = network_tcnn.NeRFNetwork(encoding="HashGrid", desired_resolution=32768, log2_hashmap_size=19,
                                   n_features_per_level=2, num_layers=2, hidden_dim=64, geo_feat_dim=15, bound=1,
                                   density_scale=1, min_near=0.2, min_near_lidar=0.0108, density_thresh=10, bg_radius=-1)
    assert tuple(net.encoder.embeddings.shape) == (6837544, 2)
    assert net.sigma_net.weights.numel() == 7168 and net.lidar_color_net.weights.numel() == 11264
    assert net.out_lidar_color_dim == 2 and len(net.get_params(1e-2)) == 6
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k in ("lidarnerf.nerf.network_tcnn",)}
    try:
        sys.modules.pop("lidarnerf.nerf.network_tcnn", None)
        compat.install(patch_lidarnerf=False)
        assert sys.modules["lidarnerf.nerf.network_tcnn"].NeRFNetwork is network_tcnn.NeRFNetwork
        assert "tinycudann" not in sys.modules
    finally:
        sys.modules.pop("lidarnerf.nerf.network_tcnn", None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


def test_ffmlp_narrow_width_padding_map_reproduces_the_narrow_network():
    """backend._widen_index: the flat weights of a hidden_dim 16 / 32 FFMLP scattered into the flat weights of the same
    network with 64 hidden units (zeros elsewhere) give bit-identical outputs, forward buffers, input gradients and
    weight gradients in the CPU restatement - the property the GPU path for those widths relies on."""
    import numpy as np
    import torch
    from oracle import oracle as orc
    from lidar_nerf_b200.backend import _widen_index
    import cases
    for hidden, ind, nl in ((16, 32, 2), (32, 48, 3)):
        c = cases.ffmlp_case(7 + hidden, 128, ind, hidden, nl, 16)
        idx, n_wide = _widen_index(torch.device("cpu"), ind, 16, hidden, nl)
        assert idx.numel() == c["w"].size and n_wide == 64 * (ind + 64 * (nl - 1) + 16)
        assert idx.unique().numel() == idx.numel()
        wide = np.zeros(n_wide, np.float16)
        wide[idx.numpy()] = c["w"]
        out_n, fb_n = orc.ffmlp_forward(c["x"], c["w"], ind, 16, hidden, nl)
        out_w, fb_w = orc.ffmlp_forward(c["x"], wide, ind, 16, 64, nl)
        assert np.array_equal(out_n, out_w) and np.array_equal(fb_n, fb_w[:, :, :hidden])
        assert not fb_w[:, :, hidden:].any()
        gi_n, gw_n, bb_n = orc.ffmlp_backward(c["g"], c["x"], c["w"], fb_n, ind, 16, hidden, nl, True)
        gi_w, gw_w, bb_w = orc.ffmlp_backward(c["g"], c["x"], wide, fb_w, ind, 16, 64, nl, True)
        assert np.array_equal(gi_n, gi_w) and np.array_equal(gw_n, gw_w[idx.numpy()]) and np.array_equal(bb_n, bb_w[:, :, :hidden])
        rest = np.ones(n_wide, bool)
        rest[idx.numpy()] = False
        assert not gw_w[rest].any()


def test_reference_linear_stack_checkpoint_maps_onto_the_ffmlp_layout():
    """NeRFNetwork.linear_stack_to_ffmlp: the reference's bias-free nn.Linear stacks (network.py:45-98) as flat FFMLP
    weights computing the same function - the two-Linear density net needs an identity layer (an FFMLP has >= 3 matmuls),
    the three-Linear LiDAR head maps one to one.  Checked through the CPU restatement of the FFMLP against the plain
    matrix products with fp16-rounded activations."""
    from oracle import oracle as orc
    from lidar_nerf_b200.nerf.network import NeRFNetwork
    rng = np.random.default_rng(5)
    h16 = lambda a: a.astype(np.float16).astype(np.float32)   # noqa: E731

    def stack(x, ws):
        h = h16(x)
        for i, w in enumerate(ws):
            h = h16(h @ h16(w).T)
            if i + 1 < len(ws):
                h = np.maximum(h, 0)
        return h

    for in_dim, pad_in, out_dim, n_lin, ff_layers in ((32, 32, 16, 2, 2), (39, 48, 2, 3, 2), (39, 48, 2, 3, 3)):
        dims = [in_dim] + [64] * (n_lin - 1) + [out_dim]
        ws = [rng.uniform(-0.3, 0.3, size=(b, a)).astype(np.float32) for a, b in zip(dims[:-1], dims[1:])]
        flat = NeRFNetwork.linear_stack_to_ffmlp(ws, pad_in, 64, ff_layers).numpy()
        assert flat.size == 64 * (pad_in + 64 * (ff_layers - 1) + 16)
        x = rng.uniform(-1, 1, size=(128, in_dim)).astype(np.float32)
        xp = np.pad(x, ((0, 0), (0, pad_in - in_dim)))
        got, _ = orc.ffmlp_forward(xp, flat, pad_in, 16, 64, ff_layers)
        np.testing.assert_allclose(got[:, :out_dim], stack(x, ws), rtol=2e-3, atol=2e-3)
        assert not got[:, out_dim:].any()
    with pytest.raises(ValueError):
        NeRFNetwork.linear_stack_to_ffmlp([np.zeros((64, 32), np.float32)] * 5, 32, 64, 2)
    # ... and through the module: a reference-shaped state_dict lands in the three flat FFMLP parameters
    net = NeRFNetwork(encoding="frequency", multires=5, use_ffmlp=True)
    g = torch.Generator().manual_seed(0)
    sd = {"sigma_net.0.weight": torch.randn(64, net.in_dim, generator=g), "sigma_net.1.weight": torch.randn(16, 64, generator=g)}
    for name, in_dim, out in (("color_net", net.in_dim_dir + 15, 3), ("lidar_color_net", net.in_dim_lidar_dir + 15, 2)):
        sd.update({f"{name}.0.weight": torch.randn(64, in_dim, generator=g), f"{name}.1.weight": torch.randn(64, 64, generator=g),
                   f"{name}.2.weight": torch.randn(out, 64, generator=g)})
    net.load_reference_state_dict(sd, strict=False)
    w = net.sigma_net.weights.detach()
    assert torch.equal(w[:64 * net.pad_in].reshape(64, net.pad_in)[:, :net.in_dim], sd["sigma_net.0.weight"])
    assert torch.equal(w[64 * net.pad_in:64 * net.pad_in + 64 * 64].reshape(64, 64), torch.eye(64))       # inserted identity
    assert torch.equal(w[-16 * 64:].reshape(16, 64), sd["sigma_net.1.weight"])
    wl = net.lidar_color_net.weights.detach()
    assert torch.equal(wl[-16 * 64:].reshape(16, 64)[:2], sd["lidar_color_net.2.weight"]) and not wl[-16 * 64:].reshape(16, 64)[2:].any()
    with pytest.raises(ValueError):
        net.load_reference_state_dict({"sigma_net.params": torch.zeros(4)})
