"""`NeRFNetwork` for the LiDAR field with the constructor/`density`/`color`/`get_params` contract of the reference's
networks (lidarnerf/nerf/network.py:10-253, network_tcnn.py:10-211), wired the way SURVEY.md section 3.2 describes:
hash grid -> FFMLP (sigma + 15 geo features) and [freq(dir) | geo] -> FFMLP (ray-drop, intensity), all on the sm_100a
kernels.  `use_ffmlp=False` swaps the fused MLPs for bias-free nn.Linear stacks (the reference's network.py layout),
which is what BASELINE config 1 (CPU plumbing) uses together with `encoding="frequency"`.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..activation import trunc_exp
from ..encoding import get_encoder
from .renderer import NeRFRenderer


def _linear_stack(in_dim, hidden, out_dim, n_layers):
    dims = [in_dim] + [hidden] * (n_layers - 1) + [out_dim]
    return nn.ModuleList([nn.Linear(a, b, bias=False) for a, b in zip(dims[:-1], dims[1:])])


def _run_stack(layers, h):
    for i, lin in enumerate(layers):
        h = lin(h)
        if i != len(layers) - 1:
            h = F.relu(h, inplace=True)
    return h


class NeRFNetwork(NeRFRenderer):
    def __init__(self, encoding="hashgrid", encoding_dir="frequency", multires=15, encoding_bg="hashgrid",
                 desired_resolution=2048, log2_hashmap_size=19, num_layers=2, hidden_dim=64, geo_feat_dim=15,
                 num_layers_color=3, hidden_dim_color=64, num_layers_bg=2, hidden_dim_bg=64, out_color_dim=3,
                 out_lidar_color_dim=2, bound=1, use_ffmlp=True, level_dim=2, **kwargs):
        # positional order and keyword set of network.py:11-30 (+ use_ffmlp / level_dim, this package's extensions)
        super().__init__(bound, **kwargs)
        self.num_layers, self.hidden_dim, self.geo_feat_dim = num_layers, hidden_dim, geo_feat_dim
        self.num_layers_color, self.hidden_dim_color = num_layers_color, hidden_dim_color
        self.out_color_dim, self.out_lidar_color_dim = out_color_dim, out_lidar_color_dim
        self.use_ffmlp = use_ffmlp

        if encoding == "frequency":
            self.encoder, self.in_dim = get_encoder("frequency", multires=multires)
        else:
            self.encoder, self.in_dim = get_encoder(encoding, desired_resolution=desired_resolution,
                                                    log2_hashmap_size=log2_hashmap_size, level_dim=level_dim)
        self.encoder_dir, self.in_dim_dir = get_encoder("sphere_harmonics")
        self.encoder_lidar_dir, self.in_dim_lidar_dir = get_encoder("frequency", multires=12)
        raw_rgb, raw_lidar = self.in_dim_dir + geo_feat_dim, self.in_dim_lidar_dir + geo_feat_dim
        if use_ffmlp:
            from ..ffmlp import FFMLP
            pad = lambda n: (n + 15) // 16 * 16   # noqa: E731
            self.pad_in, self.pad_rgb, self.pad_lidar = pad(self.in_dim), pad(raw_rgb), pad(raw_lidar)
            # An FFMLP with L layers has L + 1 matmuls and needs L >= 2 (ffmlp.py:217): a num_layers-deep nn.Linear stack
            # (num_layers matmuls) maps to max(num_layers - 1, 2) FFMLP layers - for the default num_layers = 2 that is
            # one matmul more than the reference's two-Linear density net, the smallest FFMLP there is.
            self.sigma_net = FFMLP(self.pad_in, 1 + geo_feat_dim, hidden_dim, max(num_layers - 1, 2))
            self.color_net = FFMLP(self.pad_rgb, out_color_dim, hidden_dim_color, max(num_layers_color - 1, 2))
            self.lidar_color_net = FFMLP(self.pad_lidar, out_lidar_color_dim, hidden_dim_color,
                                         max(num_layers_color - 1, 2))
        else:
            self.sigma_net = _linear_stack(self.in_dim, hidden_dim, 1 + geo_feat_dim, num_layers)
            self.color_net = _linear_stack(raw_rgb, hidden_dim_color, out_color_dim, num_layers_color)
            self.lidar_color_net = _linear_stack(raw_lidar, hidden_dim_color, out_lidar_color_dim, num_layers_color)
        if self.bg_radius > 0:
            # network.py:100-128 builds a 2-D hash grid + MLP for a background sphere; the LiDAR branch never blends a
            # background (renderer.py:277-290 is skipped for cal_lidar_color) and this package does not provide one
            raise NotImplementedError("bg_radius > 0 (background model) is not implemented in lidar-nerf_b200; the LiDAR "
                                      "branch does not use it - pass bg_radius <= 0")

    # ------------------------------------------------------------------------------------------------------------
    # checkpoints of the reference's network.py (bias-free nn.Linear stacks: `sigma_net.{i}.weight`, ...)
    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def linear_stack_to_ffmlp(weights, input_dim_padded, hidden_dim, ffmlp_layers, padded_output_dim=16):
        """[W_0 (hidden x in), ..., W_last (out x hidden)] of a bias-free ReLU nn.Linear stack -> the flat FFMLP weight
        vector [hidden*in | (L-1)*hidden^2 | 16*hidden] (ffmlp.cu:861-864) of an FFMLP with `ffmlp_layers` layers computing
        the SAME function: input columns / output rows are zero-padded, and where the FFMLP has more matmuls than the stack
        (the reference's two-Linear density net vs the smallest FFMLP, three matmuls) identity layers are inserted behind
        the first ReLU - relu(I relu(h)) = relu(h) exactly, also in fp16."""
        weights = [torch.as_tensor(w, dtype=torch.float32) for w in weights]
        n_mat = ffmlp_layers + 1
        if len(weights) > n_mat or len(weights) < 2:
            raise ValueError(f"a stack of {len(weights)} Linear layers does not fit an FFMLP with {ffmlp_layers} layers")
        w_in, w_out, mids = weights[0], weights[-1], weights[1:-1]
        if w_in.shape[0] != hidden_dim or w_out.shape[1] != hidden_dim or w_in.shape[1] > input_dim_padded or \
                w_out.shape[0] > padded_output_dim or any(tuple(m.shape) != (hidden_dim, hidden_dim) for m in mids):
            raise ValueError("layer shapes do not match the FFMLP")
        eye = torch.eye(hidden_dim)
        mids = [eye] * (n_mat - 2 - len(mids)) + mids
        flat = [F.pad(w_in, (0, input_dim_padded - w_in.shape[1])).reshape(-1)]
        flat += [m.reshape(-1) for m in mids]
        flat.append(F.pad(w_out, (0, 0, 0, padded_output_dim - w_out.shape[0])).reshape(-1))
        return torch.cat(flat)

    def load_reference_state_dict(self, state_dict, strict=True):
        """Load a checkpoint of the reference's `NeRFNetwork` (lidarnerf/nerf/network.py: `encoder.embeddings`,
        `sigma_net.{i}.weight`, `color_net.{i}.weight`, `lidar_color_net.{i}.weight`, renderer buffers).  With
        use_ffmlp=False the keys coincide and this is load_state_dict; with the fused MLPs the Linear stacks are converted
        (linear_stack_to_ffmlp).  Checkpoints of network_tcnn.py hold tiny-cuda-nn's private parameter layout and cannot
        be converted."""
        sd = dict(state_dict.get("model", state_dict))
        if any(k.endswith(".params") for k in sd):
            raise ValueError("tiny-cuda-nn checkpoint (network_tcnn.py): its parameter layout is private to tcnn")
        if not self.use_ffmlp:
            return self.load_state_dict(sd, strict=strict)
        out = {}
        for name, net, pad_in in (("sigma_net", self.sigma_net, self.pad_in), ("color_net", self.color_net, self.pad_rgb),
                                  ("lidar_color_net", self.lidar_color_net, self.pad_lidar)):
            ws, i = [], 0
            while f"{name}.{i}.weight" in sd:
                ws.append(sd.pop(f"{name}.{i}.weight"))
                i += 1
            if ws:
                out[f"{name}.weights"] = self.linear_stack_to_ffmlp(ws, pad_in, net.hidden_dim, net.num_layers,
                                                                    net.padded_output_dim)
        out.update(sd)
        return self.load_state_dict(out, strict=strict)

    def fused_unsupported_reason(self):
        from .fused_render import FusedLidarRender
        return FusedLidarRender.supported(self)

    @staticmethod
    def _pad(h, width):
        return h if h.shape[-1] == width else F.pad(h, (0, width - h.shape[-1]))

    def density(self, x):
        h = self.encoder(x, bound=self.bound) if not isinstance(self.encoder, nn.Identity) else x
        if self.use_ffmlp:
            with torch.autocast("cuda", dtype=torch.float16):
                h = self.sigma_net(self._pad(h, self.pad_in))
        else:
            h = _run_stack(self.sigma_net, h)
        # exp in fp32 whatever the autocast state (activation.py:6-20 casts its input; `custom_fwd(cast_inputs=...)` only
        # does so while autocast is ACTIVE, and the FFMLP output is fp16: exp(h > 11.09) overflows there)
        return {"sigma": trunc_exp(h[..., 0].float()), "geo_feat": h[..., 1:]}

    def color(self, x, d, cal_lidar_color=False, mask=None, geo_feat=None, **kwargs):
        out_dim = self.out_lidar_color_dim if cal_lidar_color else self.out_color_dim
        if mask is not None:
            rgbs = torch.zeros(mask.shape[0], out_dim, dtype=x.dtype, device=x.device)
            if not mask.any():
                return rgbs
            d, geo_feat = d[mask], geo_feat[mask]
        enc = self.encoder_lidar_dir(d) if cal_lidar_color else self.encoder_dir(d)
        h = torch.cat([enc, geo_feat.to(enc.dtype)], dim=-1)
        net = self.lidar_color_net if cal_lidar_color else self.color_net
        if self.use_ffmlp:
            with torch.autocast("cuda", dtype=torch.float16):
                h = net(self._pad(h, self.pad_lidar if cal_lidar_color else self.pad_rgb))
        else:
            h = _run_stack(net, h)
        h = torch.sigmoid(h.float())
        if mask is None:
            return h
        rgbs[mask] = h.to(rgbs.dtype)
        return rgbs

    def forward(self, x, d):
        dens = self.density(x)
        return dens["sigma"], self.color(x, d, geo_feat=dens["geo_feat"])

    def get_params(self, lr):
        groups = [self.encoder, self.sigma_net, self.encoder_dir, self.color_net, self.encoder_lidar_dir,
                  self.lidar_color_net]
        return [{"params": list(m.parameters()), "lr": lr} for m in groups]
