import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icrl_b200 import synth
from icrl_b200.engine import A2CEngine, H
from tests.helpers import make_nets
seed, B, L = 311, 512, 12
A, R, w = make_nets(seed)
with torch.no_grad():
    A.value_network.valrnn.lstm.bias_hh_l0[512:1024] += 4.0
    R.rewrnn.gru.bias_hh_l0[512:1024] += 2.0
f, c = synth.make_inputs(seed, B, L)
u = synth.make_uniforms(seed, L - 1, B)
e1 = A2CEngine(A, R, chain_segments=1)
r1 = e1.step(f, c, uniforms=u)
Tv = r1["Tv"]
for k, p in A.named_parameters():
    g = p.grad
    print(k, "nan" if torch.isnan(g).any() else "ok", float(g[torch.isfinite(g)].abs().max()) if torch.isfinite(g).any() else None)
b = e1._bufs
for name, n in (("v_stash_h", (Tv + 1) * H), ("v_stash_c", (Tv + 1) * H), ("v_stash_g", Tv * 4 * H), ("v_dgates", Tv * 4 * H), ("v_dh_take", 11 * B * H), ("dtable", 1004 * 2048)):
    t = b[name][:n]
    print(name, "finite" if torch.isfinite(t).all() else "NONFINITE", float(t[torch.isfinite(t)].abs().max()), float(t[torch.isfinite(t)].abs().min()))
dg = b["v_dgates"][:Tv * 4 * H].view(Tv, 4 * H)
print("col absmax min/max", float(dg.abs().amax(0).min()), float(dg.abs().amax(0).max()))
print("values", float(r1["values"].abs().max()), "c max", float(b["v_stash_c"][: (Tv + 1) * H].abs().max()))
