#!/usr/bin/env python
"""Time the chain-segment variants (pieces, warm-up) on one GPU and compare their results with the serial kernels.
usage: python scripts/segments_sweep.py [batch] [cfg ...]  cfg = pieces:warmup[:backward pieces]
(The round-1 sweep of 16 forward pieces as two CTAs per SM and of four backward CTA groups -- all slower -- is kept in
profiles/r01_segments_two_ctas_per_sm.log; that code was not kept.)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from icrl_b200.engine import A2CEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cfgs = [tuple(int(x) for x in c.split(":")) for c in sys.argv[2:]] or [(8, 256), (16, 256), (32, 256)]
dev = "cuda:0"
torch.cuda.set_device(0)
A, R = bench.make_nets(0, dev)
f, c, u = bench.workload(B)
S = bench.L_CAP - 1

e1 = A2CEngine(A, R, chain_segments=1)
prep = e1.prepare(f, c, u, plan=(1, S))
r1 = e1.step(prep)
v1, w1, g1 = r1["values"].clone(), r1["rewards"].clone(), e1.flat_grad.clone()
del e1
for cfg in cfgs:
    K, warm, Kb = cfg[0], cfg[1], (cfg[2] if len(cfg) > 2 else None)
    eng = A2CEngine(A, R, chain_segments=K, chain_warmup=warm, chain_bwd_segments=Kb)
    for _ in range(2):
        r = eng.step(prep)
    ev, er = float((r["values"] - v1).abs().max()), float((r["rewards"] - w1).abs().max())
    eg = float((eng.flat_grad - g1).abs().max() / g1.abs().max())
    eng.phase_events = []
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        eng.step(prep, check=False)
    e1_.record()
    torch.cuda.synchronize()
    ph = {k: round(sum(v) / len(v), 2) for k, v in eng.phase_times_ms().items()}
    ok = eng.segments_verified()
    print("pieces %2d (backward %s) warm-up %4d: %.1f ms/step  %.0f captions/s  fwd %.1f bwd %.1f | vs serial: values %.1e rewards %.1e "
          "grads %.1e | verified %s fallbacks %d layout %s" % (K, Kb, warm, e0.elapsed_time(e1_) / 3, B / (e0.elapsed_time(e1_) / 3e3),
          ph.get("chains_fwd_fused", 0), ph.get("chain_lstm_bwd", 0), ev, er, eg, ok, eng.segment_stats["fallbacks"], eng.segment_layout), flush=True)
    del eng
