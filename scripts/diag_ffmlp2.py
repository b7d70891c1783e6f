import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import cases
from oracle import oracle as orc
from lidar_nerf_b200 import backend as be
DEV = "cuda:0"
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
B, ind, nl = 256, 32, 2
c = cases.ffmlp_case(50, B, ind, 64, nl, 16)
x, w, g = T(c["x"]), T(c["w"]), T(c["g"])
fb = torch.empty(nl, B, 64, device=DEV, dtype=torch.half); out = torch.empty(B, 16, device=DEV, dtype=torch.half)
be._ffmlp.ffmlp_forward(x, w, B, ind, 16, 64, nl, 0, 6, fb, out)
gi = torch.zeros(B, ind, device=DEV, dtype=torch.half); gw = torch.zeros_like(w); bb = torch.zeros(nl, B, 64, device=DEV, dtype=torch.half)
be._ffmlp.ffmlp_backward(g, x, w, fb, B, ind, 16, 64, nl, 0, 6, True, bb, gi, gw)
torch.cuda.synchronize()
o_gi, o_gw, o_bb = orc.ffmlp_backward(c["g"], c["x"], c["w"], fb.float().cpu().numpy(), ind, 16, 64, nl, True)
G = gi.float().cpu().numpy(); BB = bb.float().cpu().numpy()
W = c["w"].astype(np.float32); w_in = W[:64 * ind].reshape(64, ind); w_hid = W[64 * ind:64 * ind + 4096].reshape(64, 64)
def close(a, b): return float(np.mean(np.isclose(a, b, rtol=2e-2, atol=5e-3)))
hi = slice(128, 256)
print("lazy", os.environ.get("LNB_FFMLP_LAZY_WGRAD"), "rows>=128: match expected", close(G[hi], o_gi[hi]))
print("  == gi rows 0..127           ", close(G[hi], o_gi[0:128]))
print("  == dpre0(g0 tile) @ W_in    ", close(G[hi], BB[1, 0:128] @ w_in))
print("  == unmasked dH0 of own tile (first 32 cols)", close(G[hi], (BB[0, hi] @ w_hid)[:, :32]))
print("  == masked dpre0 own (first 32 cols)", close(G[hi], BB[1, hi][:, :32]))
print("  == dpre1 own cols 0..31", close(G[hi], BB[0, hi][:, :32]))
print("  row 128 got", G[128, :8], "exp", o_gi[128, :8])
print("  zero frac", float(np.mean(G[hi] == 0)))
