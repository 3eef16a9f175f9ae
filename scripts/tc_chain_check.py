"""Chain pieces on tcgen05 (chain_tc.cu) against the serial chain kernels, stage by stage, then timing + cycle split.
Usage: python scripts/tc_chain_check.py [stage ...]   stages: fwd bwd step time prof  (default: all)
Every stage prints one line per case; a stage that fails raises, so wrap the call in `timeout`."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icrl_b200 import _lib, synth
from icrl_b200.engine import A2CEngine, _p, H
from bench import make_nets

stages = sys.argv[1:] or ["fwd", "bwd", "step", "time", "prof"]
dev = "cuda:0"


def engines(seed, warm, pieces=None, **kw):
    A, R = make_nets(seed, dev)
    e1 = A2CEngine(A, R, chain_segments=1)
    ek = A2CEngine(A, R, chain_warmup=warm, chain_pieces=pieces, chain_adapt=False, **kw)
    return A, R, e1, ek


def run_case(seed, B, L, warm, pieces=None, level=None, backward=True):
    A, R, e1, ek = engines(seed, warm, pieces)
    f, c = synth.make_inputs(seed, B, L)
    S = (L - 1) if level is None else level
    u = synth.make_uniforms(seed, S, B)
    r1 = e1.step(f, c, uniforms=u, level=level, backward=backward)
    v1, w1, g1 = r1["values"].clone(), r1["rewards"].clone(), e1.flat_grad.clone()
    h1 = e1._bufs["v_stash_h"][: (r1["Tv"] + 1) * H].clone()
    rk = ek.step(f, c, uniforms=u, level=level, backward=backward)
    lay = ek.piece_layout
    ev, er = float((rk["values"] - v1).abs().max()), float((rk["rewards"] - w1).abs().max())
    eh = float((ek._bufs["v_stash_h"][: (r1["Tv"] + 1) * H] - h1).abs().max()) if lay else float("nan")
    eg = float((ek.flat_grad - g1).abs().max() / g1.abs().max()) if backward else 0.0
    st = ek.segment_stats
    print("B=%d L=%d level=%s warm=%d layout=%s: values %.2e rewards %.2e stash_h %.2e grads(rel) %.2e | tokens equal %s | "
          "reruns %d fallbacks %d err %s" % (B, L, level, warm, lay, ev, er, eh, eg, bool(torch.equal(rk["tokens"], r1["tokens"])),
                                            st["reruns"], st["fallbacks"], ["%.1e" % x for x in st["tc_max_err"][:14]]), flush=True)
    return ev, er, eg, lay


if "fwd" in stages:
    print("== forward only (values / rewards vs the serial kernels)", flush=True)
    for case in ((3, 64, 10, 32, None), (4, 96, 12, 64, 5), (5, 256, 20, 64, None), (6, 300, 9, 32, 200)):
        seed, B, L, warm, pieces = case
        ev, er, _, lay = run_case(seed, B, L, warm, pieces=pieces, backward=False)
        assert lay is not None, "tc layout not chosen"
        assert ev <= 2e-6 and er <= 2e-6, (ev, er)

if "bwd" in stages:
    print("== forward + backward", flush=True)
    for case in ((7, 64, 10, 32, None, None), (8, 96, 14, 64, None, 5), (9, 256, 20, 96, None, None), (10, 512, 12, 64, 300, None)):
        seed, B, L, warm, pieces, level = case
        ev, er, eg, lay = run_case(seed, B, L, warm, pieces=pieces, level=level)
        assert lay is not None
        assert ev <= 2e-6 and er <= 2e-6 and eg <= 2e-5, (ev, er, eg)

if "step" in stages:
    print("== adaptive warm-up over a few optimizer steps, B=1024", flush=True)
    A, R = make_nets(11, dev)
    eng = A2CEngine(A, R)
    opt = torch.optim.Adam(A.parameters(), lr=1e-4)
    f, c = synth.make_inputs(11, 1024, 20)
    for i in range(12):
        res = eng.step(f, c, uniforms=synth.make_uniforms(100 + i, 19, 1024))
        opt.step()
        print("step %d loss %.6f layout %s warm %s reruns %d fallbacks %d" % (i, res.loss, eng.piece_layout, eng.warm,
              eng.segment_stats["reruns"], eng.segment_stats["fallbacks"]), flush=True)
    print("tc_max_err", ["%.2e" % x for x in eng.segment_stats["tc_max_err"][:14]], "history", eng.segment_stats["warm_history"], flush=True)

if "time" in stages or "prof" in stages:
    for B in (512, 4096):
        A, R = make_nets(0, dev)
        eng = A2CEngine(A, R)
        f, c = synth.make_inputs(100, B, 20)
        prep = eng.prepare(f, c, synth.make_uniforms(100, 19, B), plan=(1, 19))
        for _ in range(3):
            eng.step(prep)
        torch.cuda.synchronize()
        eng.phase_events = []
        buf = torch.zeros(24, dtype=torch.int64, device=dev)
        n = 3
        t0 = time.time()
        for i in range(n):
            if i == n - 1:
                _lib.call("icrl_chain_tc_set_profile", ctypes.c_void_p(buf.data_ptr()))
            eng.step(prep)
        torch.cuda.synchronize()
        wall = (time.time() - t0) / n
        _lib.call("icrl_chain_tc_set_profile", None)
        ph = {k: float(np.mean(v)) for k, v in eng.phase_times_ms().items()}
        eng.phase_events = None
        print("B=%d layout %s warm %s: step wall %.1f ms (%.0f captions/s); phases ms %s" %
              (B, eng.piece_layout, eng.warm, wall * 1e3, B / wall, {k: round(v, 2) for k, v in ph.items()}), flush=True)
        t = buf.cpu().numpy()
        names = ("acc wait", "gather", "cell", "store", "barrier")
        for label, o in (("forward, cluster 0 (value LSTM when the launch is fused)", 0), ("backward", 8), ("forward, reward GRU of the fused launch", 16)):
            if o + 6 > t.size:
                continue
            T = max(int(t[o + 5]), 1)
            extra = ", gather issue %.0f, partner wait %.0f" % (t[o + 6] / T, t[o + 7] / T) if o == 8 else ""
            print("  %s: %d steps, cycles per step: %s%s, total %.0f" % (label, T, ", ".join("%s %.0f" % (nm, t[o + i] / T) for i, nm in enumerate(names)),
                  extra, float(sum(t[o:o + 5]) + (t[o + 6] + t[o + 7] if o == 8 else 0)) / T), flush=True)
        print("  reruns %d fallbacks %d tc_max_err %s" % (eng.segment_stats["reruns"], eng.segment_stats["fallbacks"],
              ["%.1e" % x for x in eng.segment_stats["tc_max_err"][:14]]), flush=True)
