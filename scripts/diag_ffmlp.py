"""GPU diagnostic for the MLP backward variants: which outputs contain NaN / disagree with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import cases
from oracle import oracle as orc
from lidar_nerf_b200 import backend as be
DEV = "cuda:0"
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
for (B, ind, nl) in [(128, 32, 2), (256, 32, 2), (384, 32, 2), (3840, 32, 2), (128, 96, 2), (256, 96, 2), (38400, 32, 2)]:
    c = cases.ffmlp_case(50, B, ind, 64, nl, 16)
    x, w, g = T(c["x"]), T(c["w"]), T(c["g"])
    fb = torch.empty(nl, B, 64, device=DEV, dtype=torch.half); out = torch.empty(B, 16, device=DEV, dtype=torch.half)
    be._ffmlp.ffmlp_forward(x, w, B, ind, 16, 64, nl, 0, 6, fb, out)
    gi = torch.zeros(B, ind, device=DEV, dtype=torch.half); gw = torch.zeros_like(w); bb = torch.zeros(nl, B, 64, device=DEV, dtype=torch.half)
    be._ffmlp.ffmlp_backward(g, x, w, fb, B, ind, 16, 64, nl, 0, 6, True, bb, gi, gw)
    torch.cuda.synchronize()
    o_gi, o_gw, o_bb = orc.ffmlp_backward(c["g"], c["x"], c["w"], fb.float().cpu().numpy(), ind, 16, 64, nl, True)
    def rep(name, a, b):
        a = a.float().cpu().numpy()
        nan = np.isnan(a)
        bad = ~np.isclose(a, b, rtol=1e-2, atol=4e-3) | nan
        msg = f"{name}: nan={int(nan.sum())} bad={int(bad.sum())}/{a.size}"
        if bad.any():
            idx = np.argwhere(bad)
            msg += f" first={idx[0].tolist()} last={idx[-1].tolist()} rows={np.unique(idx[:, -2] if idx.shape[1] > 1 else idx[:, 0])[:8].tolist()}"
        return msg
    print(f"B={B} in={ind} nl={nl} lazy={os.environ.get('LNB_FFMLP_LAZY_WGRAD')} groups={os.environ.get('LNB_FFMLP_BWD_GROUPS')} |", rep("bb", bb, o_bb), "|", rep("gi", gi, o_gi), "|", rep("gw", gw, o_gw.reshape(-1)))
