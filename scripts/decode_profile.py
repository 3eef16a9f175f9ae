"""Phase timeline of the persistent decode kernel (CTA 0, first epilogue warp): clock64 stamps per cell step.
Run on the GPU box: python scripts/decode_profile.py [B]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icrl_b200 import _lib
from icrl_b200.engine import A2CEngine
from icrl_b200 import synth
from tests.helpers import make_nets

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
A, R, w = make_nets(0, "cuda:0")
eng = A2CEngine(A, R, decode="fused")
f, c = synth.make_inputs(100, B, 20)
u = synth.make_uniforms(100, 19, B)
prep = eng.prepare(f, c, u, plan=(1, 19))
for _ in range(2):
    eng.step(prep, backward=False)
buf = torch.zeros(19 * 16, dtype=torch.int64, device="cuda")
_lib.call("icrl_decode_set_profile", ctypes.c_void_p(buf.data_ptr()))
eng.phase_events = []
eng.step(prep, backward=False)
torch.cuda.synchronize()
_lib.call("icrl_decode_set_profile", None)
print("policy_fwd ms:", eng.phase_times_ms()["policy_fwd"])
t = buf.cpu().numpy().reshape(19, 16)
names = ["step start", "G acc ready", "G epilogue done", "cluster barrier", "V acc ready", "logits read+stored",
         "E1 max", "E2 sum", "E3 cdf totals", "counted", "E4 token"]
print("cycles since step start (CTA 0, warp 2 lane 0); one row per cell step")
order = [1, 11, 12, 2, 3, 4, 13, 5, 6, 7, 8, 9, 10]
names = {1: "Gacc", 11: "G.load", 12: "G.cell", 2: "G.stash", 3: "clusbar", 4: "Vacc", 13: "V.tmem", 5: "V.stash", 6: "E1max",
         7: "E2sum", 8: "E3cdf", 9: "count", 10: "E4tok"}
print("phase durations (cycles)")
print(" ".join("%8s" % names[k] for k in order), "    total")
for j in range(19):
    prev = t[j, 0]
    out = []
    for k in order:
        out.append(t[j, k] - prev)
        prev = t[j, k]
    print(" ".join("%8d" % v for v in out), " %8d" % (t[j, 10] - t[j, 0]))

g, v = t[2:, 14].mean(), t[2:, 15].mean()
step = (t[2:, 10] - t[2:, 0]).mean()
print("MMA issue spans (mean cycles): gate GEMM phase %.0f (96 MMAs N=256 = 12,288 tensor cycles -> %.0f %% busy inside the phase), "
      "vocab GEMM phase %.0f (96 MMAs N=128 = 6,144 -> %.0f %%); whole step %.0f -> %.0f %% tensor-pipe busy"
      % (g, 100 * 12288 / g, v, 100 * 6144 / v, step, 100 * (12288 + 6144) / step))
