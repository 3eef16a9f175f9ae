"""Cycle split of the sharded forward chains (CTA 0, thread 0): exchange wait / GEMV+reduce / pointwise+publish."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icrl_b200 import _lib, synth
from icrl_b200.engine import A2CEngine
from bench import make_nets
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for K in (2, 4, 8):
    A, R = make_nets(0, "cuda:0")
    eng = A2CEngine(A, R, chain_shards=K)
    f, c = synth.make_inputs(100, B, 20)
    prep = eng.prepare(f, c, synth.make_uniforms(100, 19, B), plan=(1, 19))
    eng.step(prep, backward=False)
    buf = torch.zeros(8, dtype=torch.int64, device="cuda")
    _lib.call("icrl_chain_set_profile", ctypes.c_void_p(buf.data_ptr()))
    eng.step(prep, backward=False)
    torch.cuda.synchronize()
    _lib.call("icrl_chain_set_profile", None)
    t = buf.cpu().numpy()
    for name, o in (("LSTM", 0), ("GRU", 4)):
        T = max(int(t[o + 3]), 1)
        print("K=%d %s: per batch-step cycles: exchange wait %.0f, GEMV+reduce %.0f, pointwise+publish %.0f, total %.0f (%.0f per chain step)"
              % (K, name, t[o] / T, t[o + 1] / T, t[o + 2] / T, (t[o] + t[o + 1] + t[o + 2]) / T, (t[o] + t[o + 1] + t[o + 2]) / T / K))
