"""Cycle split of the segmented chains (CTA 0, thread 0), summed over the chunks of a kernel step.
forward: exchange wait (poll completion + barrier) / GEMV+reduce / pointwise+publish;
backward: coefficients+requests / poll wait / gate gradients+stores / barrier / contraction+reduce / barrier+publish."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icrl_b200 import _lib, synth
from icrl_b200.engine import A2CEngine
from bench import make_nets
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for K, Kb in ((8, 8), (16, 8), (32, 8)):
    A, R = make_nets(0, "cuda:0")
    eng = A2CEngine(A, R, chain_segments=K, chain_bwd_segments=Kb)
    f, c = synth.make_inputs(100, B, 20)
    prep = eng.prepare(f, c, synth.make_uniforms(100, 19, B), plan=(1, 19))
    eng.step(prep)
    buf = torch.zeros(16, dtype=torch.int64, device="cuda")
    _lib.call("icrl_chain_set_profile", ctypes.c_void_p(buf.data_ptr()))
    eng.step(prep)
    torch.cuda.synchronize()
    _lib.call("icrl_chain_set_profile", None)
    t = buf.cpu().numpy()
    for name, o in (("LSTM", 0), ("GRU", 4)):
        T = max(int(t[o + 3]), 1)
        tot = (t[o] + t[o + 1] + t[o + 2]) / T
        print("pieces=%d %s: per kernel step (%d steps): wait %.0f, GEMV+reduce %.0f, pointwise+publish %.0f, total %.0f cycles "
              "(%.0f per chunk of 8, %.0f per chain position)" % (K, name, T, t[o] / T, t[o + 1] / T, t[o + 2] / T, tot,
                                                                  tot / max(K // 8, 1), tot / K), flush=True)
    T = max(int(t[14]), 1)
    names = ("coeff+requests", "poll wait", "gate grads+stores", "barrier", "contraction+reduce", "barrier+publish")
    tot = float(sum(t[8:14])) / T
    print("pieces=%d backward (%d pieces, %d steps): " % (K, Kb, T) + ", ".join("%s %.0f" % (n, t[8 + i] / T) for i, n in enumerate(names))
          + ", total %.0f cycles per kernel step (%.0f per chain position and group)" % (tot, tot / (Kb / 2)), flush=True)
