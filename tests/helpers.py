"""Shared helpers for the GPU parity tests (the CUDA path vs the oracle / golden fixtures)."""
import os

import numpy as np
import torch

from oracle import synth
from oracle.gen_golden import grad_sample_index

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_nets(seed, dev="cuda:0"):
    """Drop-in modules loaded with the synthetic weights of `seed` (same as the golden generator)."""
    import icrl_b200.models as M
    w = synth.make_weights(seed)
    w2i = synth.word_to_idx()
    P, V, R = M.PolicyNetwork(w2i), M.ValueNetwork(w2i), M.RewardNetwork(w2i)
    P.load_state_dict(w["policy"])
    V.load_state_dict(w["value"])
    R.load_state_dict(w["reward"])
    R.requires_grad_(False)
    R.train(False)
    A = M.AdvantageActorCriticNetwork(V, P)
    return A.to(dev), R.to(dev), w


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    seed, B, L = int(g["seed"]), int(g["B"]), int(g["L"])
    level = int(g["level"])
    level = None if level < 0 else level
    f, c = synth.make_inputs(seed, B, L)
    S = (L - 1) if level is None else level
    return g, seed, f, c, synth.make_uniforms(seed, S, B), level


def named_grads(a2c):
    return {k: p.grad.detach().float().cpu().numpy() for k, p in a2c.named_parameters()}


def check_grads_vs_golden(grads, g, tol):
    """Sampled entries within tol * max|ref| per tensor and L2 norm within tol (relative)."""
    worst = 0.0
    for k, grad in grads.items():
        flat = grad.reshape(-1)
        ref = g["gsamp/" + k]
        got = flat[grad_sample_index(flat.size)]
        scale = max(float(np.abs(ref).max()), 1e-12)
        err = float(np.abs(got - ref).max()) / scale
        worst = max(worst, err)
        assert err <= tol, "%s: sampled grad error %.3e of max (tol %.1e)" % (k, err, tol)
        nrm = float(np.sqrt((flat.astype(np.float64) ** 2).sum()))
        ref_n = float(g["gnorm/" + k])
        assert abs(nrm - ref_n) <= tol * max(ref_n, 1e-12), "%s: grad norm %.6e vs %.6e" % (k, nrm, ref_n)
    return worst


def check_grads_vs_oracle(grads, ref_grads, tol):
    for k, grad in grads.items():
        ref = ref_grads[k].detach().numpy()
        scale = max(float(np.abs(ref).max()), 1e-12)
        err = float(np.abs(grad - ref).max()) / scale
        assert err <= tol, "%s: grad error %.3e of max (tol %.1e)" % (k, err, tol)
