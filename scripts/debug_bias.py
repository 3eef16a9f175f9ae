import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icrl_b200 import synth, _lib
from icrl_b200.engine import A2CEngine, H
from icrl_b200.optim import FlatAdam
from tests.helpers import make_nets
seed, B, L = 307, 1024, 20
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
A, R, w = make_nets(seed)
f, c = synth.make_inputs(seed, B, L)
u = synth.make_uniforms(seed + 999, L - 1, B)
def compare(tag):
    e1 = A2CEngine(A, R, chain_segments=1)
    r1 = e1.step(f, c, uniforms=u); g1 = e1.flat_grad.clone(); Tv = r1["Tv"]
    h1 = e1._bufs["v_stash_h"][: (Tv + 1) * H].clone(); rh1 = e1._bufs["r_stash_h"][: (r1["Tr"] + 1) * H].clone()
    d1 = e1._bufs["v_dgates"][: Tv * 4 * H].clone()
    for bf, bb in ((0.0, 0.0), (2.4e-7, 9.6e-7), (4.8e-7, 1.9e-6), (7.2e-7, 2.9e-6), (9.6e-7, 3.8e-6)):
        _lib.call("icrl_chain_tc_set_bias", bf, bb)
        ek = A2CEngine(A, R, chain_warmup=512, chain_adapt=False)
        rk = ek.step(f, c, uniforms=u)
        hk = ek._bufs["v_stash_h"][: (Tv + 1) * H]; rhk = ek._bufs["r_stash_h"][: (r1["Tr"] + 1) * H]
        dk = ek._bufs["v_dgates"][: Tv * 4 * H]
        dh = (hk - h1); drh = rhk - rh1; dd = dk - d1
        print("%s bias fwd %.1e bwd %.1e: stash_h max %.2e mean-signed %.2e rms %.2e | reward h max %.2e | values %.2e | dgates max/maxref %.2e | grads rel %.2e" % (
            tag, bf, bb, float(dh.abs().max()), float((dh * torch.sign(h1)).mean()), float(dh.pow(2).mean().sqrt()), float(drh.abs().max()),
            float((rk["values"] - r1["values"]).abs().max()), float(dd.abs().max() / d1.abs().max()),
            float((ek.flat_grad - g1).abs().max() / g1.abs().max())), flush=True)
    _lib.call("icrl_chain_tc_set_bias", 0.0, 0.0)
compare("init ")
eng = A2CEngine(A, R)
opt = FlatAdam(eng, lr=1e-4)
for i in range(nsteps):
    eng.step(f, c, uniforms=synth.make_uniforms(seed + i, L - 1, B)); opt.step()
compare("after%d" % nsteps)
