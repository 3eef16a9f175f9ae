"""Drop-in replacements for the reference's ``models.py`` classes (same names, constructor
signatures, attribute names and ``state_dict`` keys), computing on B200
through ``libicrl_b200.so``.

A user of the reference puts this directory first on ``sys.path``; ``from models import *``
(``trainers.py:18``, ``utilities.py:24``) then resolves here.  Like the reference module it
re-exports ``torch, nn, F, np, device, MAX_SEQ_LEN`` and ``repackage_hidden``.

The layers are held in the same ``nn.Embedding / nn.Linear / nn.LSTM / nn.GRU`` containers as the
reference (``models.py:62-69, 114-120, 160-161, 209-215, 250-251``) so checkpoints written by
either side load in the other (``load_state_dict(strict=False)``, ``trainers.py:342-364``); the
containers' own ``forward`` is never called -- every ``forward`` below launches the CUDA kernels and
raises if the tensors are not on a CUDA device (there is no CPU fallback).

Fast training goes through ``engine.A2CEngine`` (the fused minibatch, used by our ``trainers.py``).
The per-call ``forward`` methods here reproduce the reference call semantics (growing prefix, hidden
state carried in ``valrnn.hidden_cell`` / ``rewrnn.hidden_cell`` until ``init_hidden()``) AND carry
autograd history (``torch.autograd.Function`` wrappers around the same CUDA kernels), so
the reference's own rollout loop (``trainers.py:441-480``: softmax / gather / log / stack /
``loss.backward(retain_graph=True)`` / Adam) trains these modules unmodified; like the reference,
``valrnn.hidden_cell`` keeps its graph from call to call, so gradients flow through the whole
carried-state chain.  The reward network's forward is differentiable too (GRU BPTT kernel), which is what the
reference's pretraining loops need (``train_reward_network``).  Both routes give the same numbers
(tests/test_gpu_parity.py).
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn
from torch.nn import functional as F

from . import _lib

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")   # models.py:17
MAX_SEQ_LEN = 17                                                          # models.py:18
HID = 512

__all__ = ["torch", "nn", "F", "np", "device", "MAX_SEQ_LEN", "repackage_hidden", "PolicyNetwork",
           "ValueNetworkRNN", "ValueNetwork", "RewardNetworkRNN", "RewardNetwork", "AdvantageActorCriticNetwork"]


def repackage_hidden(h):
    """Detach hidden states from their history (models.py:20-30; all reference call sites are
    commented out -- kept for import compatibility)."""
    if isinstance(h, torch.Tensor):
        return h.detach().to(device)
    return tuple(repackage_hidden(v) for v in h)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _unsupported(bidirectional, pretrained_embeddings=None):
    """Both constructor variants are built: frozen pretrained embeddings run on every path; bidirectional=True runs on
    the module (autograd) route only -- the fused A2CEngine refuses bidirectional networks and icrl_b200.trainers
    falls back to the reference-style loop for them."""
    return None


def _embedding(vocab_size, wordvec_dim, pretrained_embeddings):
    """Trainable (V, wordvec_dim) table, or the frozen pretrained vectors and their width (models.py:61-65)."""
    if pretrained_embeddings is not None:
        emb = nn.Embedding.from_pretrained(torch.FloatTensor(np.asarray(pretrained_embeddings)), freeze=True)
        return emb, int(emb.weight.shape[1])
    return nn.Embedding(vocab_size, wordvec_dim), wordvec_dim



def _st(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _sync_state(dev, cache={}):
    t = cache.get(dev)
    if t is None:
        t = torch.zeros(int(_lib.call("icrl_chain_sync_bytes")), dtype=torch.uint8, device=dev)
        cache[dev] = t
    return t


def _gemm_ws(dev, cache={}):
    t = cache.get(dev)
    if t is None:
        t = torch.empty(24 * 4 * HID * HID, dtype=torch.float32, device=dev)
        cache[dev] = t
    return t


class _PolicyFn(torch.autograd.Function):
    """PolicyNetwork.forward (models.py:71-84) with a hand-written backward: logits of all n positions
    from one incremental pass, BPTT through the n cells in backward."""

    @staticmethod
    def forward(ctx, f, captions, E, Wc, bc, W_ih, W_hh, b_ih, b_hh, Wv, bv):
        dev = f.device
        B, n = captions.shape
        V = Wv.shape[0]
        st = _st(dev)
        with torch.cuda.device(dev):
            table = torch.empty(V * 4 * HID, dtype=torch.float32, device=dev)
            _lib.call("icrl_pack_gate_table", st, V, 4 * HID, 4 * HID, E.shape[1], _p(E), _p(W_ih), _p(b_ih), _p(b_hh), _p(table), None)
            tokcm = torch.empty((n + 1) * B, dtype=torch.int32, device=dev)
            tokcm[:B] = captions[:, 0].to(torch.int32)
            forced = torch.zeros((B, n), dtype=torch.int64, device=dev)
            if n > 1:
                forced[:, :n - 1] = captions[:, 1:]
            tokens = torch.empty((B, n), dtype=torch.int64, device=dev)
            logp = torch.empty((B, n), dtype=torch.float32, device=dev)
            logits = torch.empty((n, B, V), dtype=torch.float32, device=dev)
            Hs = torch.empty((n + 1) * B * HID, dtype=torch.float32, device=dev)
            Cs = torch.empty((n + 1) * B * HID, dtype=torch.float32, device=dev)
            Gs = torch.empty(n * B * 4 * HID, dtype=torch.float32, device=dev)
            gpre = torch.empty(B * 4 * HID, dtype=torch.float32, device=dev)
            _lib.call("icrl_policy_rollout_fwd", st, B, V, 1, n, 0, _p(f), _p(Wc), _p(bc), _p(table), _p(W_hh), _p(Wv),
                      _p(bv), None, _p(forced), _p(tokcm), _p(tokens), _p(logp), _p(Hs), _p(Cs), _p(Gs), _p(logits),
                      _p(gpre), None)
        ctx.save_for_backward(f, E, W_ih, W_hh, Wv, tokcm, tokens, Hs, Cs, Gs)
        ctx.shape = (B, n, V)
        return logits.permute(1, 0, 2)

    @staticmethod
    def backward(ctx, dlogits):
        f, E, W_ih, W_hh, Wv, tokcm, tokens, Hs, Cs, Gs = ctx.saved_tensors
        B, n, V = ctx.shape
        dev = f.device
        st = _st(dev)
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            dZ = dlogits.permute(1, 0, 2).contiguous().clone()          # [n][B][V], consumed in place
            D = E.shape[1]
            dE = new(V, D) if ctx.needs_input_grad[2] else None           # frozen pretrained embedding: skipped
            dWc, dbc, dWih, dWhh = new(HID, HID), new(HID), new(4 * HID, D), new(4 * HID, HID)
            dbih, dbhh, dWv, dbv = new(4 * HID), new(4 * HID), new(V, HID), new(V)
            cs = int(_lib.call("icrl_colsum_ws_floats", max(n * B, V, B), 4 * HID)) + 2 * HID + 4 * HID * 8
            ws = _gemm_ws(dev)
            # scratch tensors are held in locals until the call returns (a temporary would be recycled at once)
            dHv, DG, dh, dc, dtable, csws = new(n * B, HID), new(n * B, 4 * HID), new(2 * B, HID), new(B, HID), new(V, 4 * HID), new(cs)
            _lib.call("icrl_policy_rollout_bwd", st, B, V, 1, n, D, _p(f), _p(E), _p(W_ih), _p(W_hh), _p(Wv), _p(tokcm),
                      _p(tokens), None, _p(Hs), _p(Cs), _p(Gs), _p(dZ), _p(dHv), _p(DG),
                      _p(dh), _p(dc), _p(dtable), _p(csws), _p(ws), ws.numel() * 4,
                      _p(dE), _p(dWc), _p(dbc), _p(dWih), _p(dWhh), _p(dbih), _p(dbhh), _p(dWv), _p(dbv), None)
        return None, None, dE, dWc, dbc, dWih, dWhh, dbih, dbhh, dWv, dbv


class _LstmSeqFn(torch.autograd.Function):
    """Teacher-forced LSTM over token columns [n][B] from h0 (c0 = 0): h after every cell, [n][B][512].  One
    direction of the bidirectional policy (models.py:59-78); the reverse direction passes the columns flipped."""

    @staticmethod
    def forward(ctx, tok_cm, h0, E, W_ih, W_hh, b_ih, b_hh):
        dev = E.device
        n, B = tok_cm.shape
        V = E.shape[0]
        st = _st(dev)
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            table = new(V, 4 * HID)
            _lib.call("icrl_pack_gate_table", st, V, 4 * HID, 4 * HID, E.shape[1], _p(E), _p(W_ih), _p(b_ih), _p(b_hh), _p(table), None)
            tok = tok_cm.contiguous()
            h0c = h0.contiguous()
            Hs, Cs, Gs, gpre = new(n + 1, B, HID), new(n + 1, B, HID), new(n, B, 4 * HID), new(B, 4 * HID)
            _lib.call("icrl_lstm_seq_fwd", st, B, n, _p(h0c), _p(tok), _p(table), _p(W_hh), _p(Hs), _p(Cs), _p(Gs), _p(gpre), None)
        ctx.save_for_backward(tok, E, W_ih, W_hh, Hs, Cs, Gs)
        return Hs[1:].clone()

    @staticmethod
    def backward(ctx, dH):
        tok, E, W_ih, W_hh, Hs, Cs, Gs = ctx.saved_tensors
        n, B = tok.shape
        V, D = E.shape
        dev = E.device
        st = _st(dev)
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            dHc = dH.contiguous()
            DG, dh, dc, dtable = new(n * B, 4 * HID), new(2 * B, HID), new(B, HID), new(V, 4 * HID)
            cs = int(_lib.call("icrl_colsum_ws_floats", max(n * B, V), 4 * HID))
            csws, ws = new(cs), _gemm_ws(dev)
            dh0 = new(B, HID)
            dE = new(V, D) if ctx.needs_input_grad[2] else None
            dWih, dWhh, dbih, dbhh = new(4 * HID, D), new(4 * HID, HID), new(4 * HID), new(4 * HID)
            _lib.call("icrl_lstm_seq_bwd", st, B, n, V, D, _p(tok), _p(Hs), _p(Cs), _p(Gs), _p(dHc), _p(W_hh), _p(E), _p(W_ih),
                      _p(DG), _p(dh), _p(dc), _p(dtable), _p(csws), _p(ws), ws.numel() * 4, _p(dh0), _p(dE), _p(dWih), _p(dWhh),
                      _p(dbih), _p(dbhh), None)
        return None, dh0, dE, dWih, dWhh, dbih, dbhh


class _ChainLSTMFn(torch.autograd.Function):
    """One ValueNetworkRNN segment: the columns of one call fed as a serial batch-1 LSTM from the carried
    state (models.py:130-135, 166-169).  Returns (h after every row of the last column, final h, final c);
    backward runs the serial BPTT kernel and hands dL/d(initial h, c) to the previous segment."""

    @staticmethod
    def forward(ctx, tok_cm, h0, c0, E, W_ih, W_hh, b_ih, b_hh):
        dev = E.device
        n, B = tok_cm.shape
        V, T = E.shape[0], n * B
        st = _st(dev)
        with torch.cuda.device(dev):
            table = torch.empty(V * 4 * HID, dtype=torch.float32, device=dev)
            _lib.call("icrl_pack_gate_table", st, V, 4 * HID, 4 * HID, E.shape[1], _p(E), _p(W_ih), _p(b_ih), _p(b_hh), _p(table), None)
            stream = tok_cm.reshape(-1).contiguous()
            stash_h = torch.empty((T + 1, HID), dtype=torch.float32, device=dev)
            stash_c = torch.empty((T + 1, HID), dtype=torch.float32, device=dev)
            stash_g = torch.empty((T, 4 * HID), dtype=torch.float32, device=dev)
            h_out = torch.empty(HID, dtype=torch.float32, device=dev)
            c_out = torch.empty(HID, dtype=torch.float32, device=dev)
            sync = _sync_state(dev)
            h0f, c0f = h0.reshape(-1).contiguous(), c0.reshape(-1).contiguous()
            _lib.call("icrl_chain_lstm_fwd", st, _p(stream), T, _p(table), _p(W_hh), _p(h0f),
                      _p(c0f), _p(stash_h), _p(stash_c), _p(stash_g), _p(h_out), _p(c_out),
                      _p(sync), None)
            _lib.call("icrl_chain_check", st, _p(sync))
        ctx.save_for_backward(stream, E, W_ih, W_hh, stash_h, stash_c, stash_g)
        ctx.dims = (n, B, V)
        ctx.state_shapes = (tuple(h0.shape), tuple(c0.shape))
        return stash_h[T - B + 1:T + 1].clone(), h_out, c_out

    @staticmethod
    def backward(ctx, dh_last, dh_out, dc_out):
        stream, E, W_ih, W_hh, stash_h, stash_c, stash_g = ctx.saved_tensors
        n, B, V = ctx.dims
        T = n * B
        dev = E.device
        st = _st(dev)
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            take = torch.full((T,), -1, dtype=torch.int32, device=dev)
            take[T - B:] = torch.arange(B, dtype=torch.int32, device=dev)
            dh_take = (dh_last if dh_last is not None else torch.zeros((B, HID), device=dev)).contiguous()
            dgates, dh0, dc0 = new(T, 4 * HID), new(HID), new(HID)
            sync = _sync_state(dev)
            dh_in = dh_out.contiguous() if dh_out is not None else None
            dc_in = dc_out.contiguous() if dc_out is not None else None
            _lib.call("icrl_chain_lstm_bwd", st, T, _p(W_hh), _p(stash_g), _p(stash_c), _p(take), _p(dh_take), _p(dgates),
                      _p(sync), _p(dh_in), _p(dc_in), _p(dh0), _p(dc0), None)
            _lib.call("icrl_chain_check", st, _p(sync))
            D = E.shape[1]
            dE = new(V, D) if ctx.needs_input_grad[3] else None
            dWih, dWhh, dbih, dbhh = new(4 * HID, D), new(4 * HID, HID), new(4 * HID), new(4 * HID)
            cs = int(_lib.call("icrl_colsum_ws_floats", max(T, V), 4 * HID))
            ws = _gemm_ws(dev)
            dtable, csws = new(V, 4 * HID), new(cs)
            _lib.call("icrl_value_chain_param_grads", st, T, V, D, _p(stream), _p(dgates), _p(stash_h), _p(E), _p(W_ih),
                      _p(dtable), _p(csws), _p(ws), ws.numel() * 4, _p(dE), _p(dWih), _p(dWhh), _p(dbih),
                      _p(dbhh), 0, 0, 0, 0, None)
        return None, dh0.view(ctx.state_shapes[0]), dc0.view(ctx.state_shapes[1]), dE, dWih, dWhh, dbih, dbhh


class _ValueHeadFn(torch.autograd.Function):
    """linear2(linear1(cat(features, h))) (models.py:175-178; no activation in between)."""

    @staticmethod
    def forward(ctx, f, h, W1, b1, W2, b2):
        dev = f.device
        B = f.shape[0]
        st = _st(dev)
        with torch.cuda.device(dev):
            weff, beff = torch.empty(2 * HID, device=dev), torch.empty(1, device=dev)
            _lib.call("icrl_pack_value_head", st, _p(W1), _p(b1), _p(W2), _p(b2), _p(weff), _p(beff), None)
            values = torch.empty((B, 1), dtype=torch.float32, device=dev)
            h = h.contiguous()
            _lib.call("icrl_value_head_fwd", st, B, 1, _p(f), _p(h), _p(weff), _p(beff), _p(values), None)
        ctx.save_for_backward(f, h, W1, b1, W2, weff)
        return values

    @staticmethod
    def backward(ctx, dv):
        f, h, W1, b1, W2, weff = ctx.saved_tensors
        dev = f.device
        B = f.shape[0]
        st = _st(dev)
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            dv = dv.reshape(-1).contiguous()
            sum_dv = dv.sum().reshape(1)
            dh, dW1, db1, dW2, db2 = new(B, HID), new(HID, 2 * HID), new(HID), new(1, HID), new(1)
            ws = new(2 * HID + int(_lib.call("icrl_colsum_ws_floats", B, HID)) + 1024)
            _lib.call("icrl_value_head_bwd", st, B, 1, _p(f), _p(h), _p(dv), _p(sum_dv), _p(W1), _p(b1), _p(W2), _p(weff),
                      _p(dh), _p(dW1), _p(db1), _p(dW2), _p(db2), _p(ws), None)
        return None, dh, dW1, db1, dW2, db2


class _ChainGRUFn(torch.autograd.Function):
    """One RewardNetworkRNN segment (models.py:223-228, 253-255): serial batch-1 GRU from the carried state.
    Returns (h after every row of the last column, final h); backward = serial GRU BPTT kernel.  Only needed to
    TRAIN the reward network (train_reward_network); the A2C step keeps it frozen."""

    @staticmethod
    def forward(ctx, tok_cm, h0, E, W_ih, W_hh, b_ih, b_hh):
        dev = E.device
        n, B = tok_cm.shape
        V, T = E.shape[0], n * B
        st = _st(dev)
        train = any(ctx.needs_input_grad)
        with torch.cuda.device(dev):
            table = torch.empty(V * 3 * HID, dtype=torch.float32, device=dev)
            _lib.call("icrl_pack_gate_table", st, V, 3 * HID, 2 * HID, E.shape[1], _p(E), _p(W_ih), _p(b_ih), _p(b_hh), _p(table), None)
            stream = tok_cm.reshape(-1).contiguous()
            stash_h = torch.empty((T + 1, HID), dtype=torch.float32, device=dev)
            stash_g = torch.empty((T, 4 * HID), dtype=torch.float32, device=dev) if train else None
            h_out = torch.empty(HID, dtype=torch.float32, device=dev)
            h0f, bhn = h0.reshape(-1).contiguous(), b_hh[2 * HID:].contiguous()
            sync = _sync_state(dev)
            _lib.call("icrl_chain_gru_fwd", st, _p(stream), T, _p(table), _p(W_hh), _p(bhn), _p(h0f), _p(stash_h), _p(h_out),
                      _p(sync), _p(stash_g), None)
            _lib.call("icrl_chain_check", st, _p(sync))
        if train:
            ctx.save_for_backward(stream, E, W_ih, W_hh, stash_h, stash_g)
        ctx.dims = (n, B, V)
        ctx.state_shape = tuple(h0.shape)
        return stash_h[T - B + 1:T + 1].clone(), h_out

    @staticmethod
    def backward(ctx, dh_last, dh_out):
        stream, E, W_ih, W_hh, stash_h, stash_g = ctx.saved_tensors
        n, B, V = ctx.dims
        T = n * B
        dev = E.device
        st = _st(dev)
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            take = torch.full((T,), -1, dtype=torch.int32, device=dev)
            take[T - B:] = torch.arange(B, dtype=torch.int32, device=dev)
            dh_take = (dh_last if dh_last is not None else torch.zeros((B, HID), device=dev)).contiguous()
            dgh, dgx, dh0 = new(T, 3 * HID), new(T, 3 * HID), new(HID)
            dh_in = dh_out.contiguous() if dh_out is not None else None
            sync = _sync_state(dev)
            _lib.call("icrl_chain_gru_bwd", st, T, _p(W_hh), _p(stash_g), _p(stash_h), _p(take), _p(dh_take), _p(dgh), _p(dgx),
                      _p(sync), _p(dh_in), _p(dh0), None)
            _lib.call("icrl_chain_check", st, _p(sync))
            D = E.shape[1]
            dE = new(V, D) if ctx.needs_input_grad[2] else None
            dWih, dWhh, dbih, dbhh = new(3 * HID, D), new(3 * HID, HID), new(3 * HID), new(3 * HID)
            cs = int(_lib.call("icrl_colsum_ws_floats", max(T, V), 3 * HID)) + 3 * HID
            dtable, csws, ws = new(V, 3 * HID), new(cs), _gemm_ws(dev)
            _lib.call("icrl_reward_chain_param_grads", st, T, V, D, _p(stream), _p(dgh), _p(dgx), _p(stash_h), _p(E), _p(W_ih),
                      _p(dtable), _p(csws), _p(ws), ws.numel() * 4, _p(dE), _p(dWih), _p(dWhh), _p(dbih), _p(dbhh), None)
        return None, dh0.view(ctx.state_shape), dE, dWih, dWhh, dbih, dbhh


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on the library's GEMM (visual_embed / semantic_embed, models.py:259-260)."""

    @staticmethod
    def forward(ctx, x, W, b):
        dev = x.device
        M, K = x.shape
        N = W.shape[0]
        with torch.cuda.device(dev):
            x = x.contiguous()
            y = torch.empty((M, N), dtype=torch.float32, device=dev)
            _lib.call("icrl_gemm_f32", _st(dev), 0, 1, M, N, K, _p(x), K, _p(W), K, _p(y), N, _p(b), 0.0, None, 0, None)
        ctx.save_for_backward(x, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dev = x.device
        M, K = x.shape
        N = W.shape[0]
        new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            dy = dy.contiguous()
            dx = new(M, K) if ctx.needs_input_grad[0] else None
            dW, db = new(N, K), new(N)
            csws, ws = new(int(_lib.call("icrl_colsum_ws_floats", M, N))), _gemm_ws(dev)
            _lib.call("icrl_linear_bwd", _st(dev), M, N, K, _p(x), _p(W), _p(dy), _p(dx), _p(dW), _p(db), _p(csws), _p(ws),
                      ws.numel() * 4, None)
        return dx, dW, db


class _KernelModule(nn.Module):
    """Workspace + stream plumbing shared by the drop-in modules."""

    def _rt_init(self):
        object.__setattr__(self, "_ws", {})
        object.__setattr__(self, "_launches", _lib.Launches())

    def _dev(self):
        d = next(self.parameters()).device
        if d.type != "cuda":
            raise _lib.IcrlError("%s.forward needs CUDA tensors: no CPU fallback exists" % type(self).__name__)
        _lib.load()
        return d

    def _buf(self, name, numel, dtype=torch.float32, dev=None):
        t = self._ws.get(name)
        numel = int(max(numel, 1))
        if t is None or t.numel() < numel or t.dtype != dtype or t.device != dev:
            t = torch.empty(numel, dtype=dtype, device=dev)
            self._ws[name] = t
        return t

    def _stream(self, dev):
        return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def _tokcm(self, captions, dev):
        caps = captions.to(dev).to(torch.int32)
        return caps.t().contiguous()            # [n][B]


class PolicyNetwork(_KernelModule):
    """Actor: LSTM over the caption prefix initialised from the image features (models.py:33-84)."""

    def __init__(self, word_to_idx, input_dim=512, wordvec_dim=512, hidden_dim=512,
                 pretrained_embeddings=None, bidirectional=False):
        super().__init__()
        _unsupported(bidirectional, pretrained_embeddings)
        if (input_dim, hidden_dim) != (HID, HID):
            raise NotImplementedError("the B200 kernels are specialised for 512-wide features / hidden state (models.py:41)")
        self.bidirectional = bidirectional
        self.word_to_idx = word_to_idx
        self.idx_to_word = {i: w for w, i in word_to_idx.items()}
        vocab_size = len(word_to_idx)
        num_dim = 2 if bidirectional else 1                    # models.py:58
        self.caption_embedding, wordvec_dim = _embedding(vocab_size, wordvec_dim, pretrained_embeddings)
        self.cnn2linear = nn.Linear(input_dim, hidden_dim * num_dim)
        self.lstm = nn.LSTM(wordvec_dim, hidden_dim, batch_first=True, bidirectional=bidirectional)
        self.linear2vocab = nn.Linear(hidden_dim * num_dim, vocab_size)
        self._rt_init()

    def forward(self, features, captions):
        """features (1,B,512), captions (B,n) int64 -> logits (B,n,V) (models.py:71-84), with autograd history
        w.r.t. the nine policy parameters."""
        dev = self._dev()
        B, n = captions.shape
        f = features.reshape(B, HID).to(dev, torch.float32).contiguous()
        if self.bidirectional:
            # models.py:76-82: cnn2linear gives both directions' initial states (first half forward, second half
            # reverse); the reverse direction reads the prefix back to front; logits from [h_fwd, h_rev].
            L = self.lstm
            hinit = _LinearFn.apply(f, self.cnn2linear.weight, self.cnn2linear.bias)
            tok = captions.to(dev).t().to(torch.int32).contiguous()
            E = self.caption_embedding.weight
            Hf = _LstmSeqFn.apply(tok, hinit[:, :HID], E, L.weight_ih_l0, L.weight_hh_l0, L.bias_ih_l0, L.bias_hh_l0)
            Hr = _LstmSeqFn.apply(tok.flip(0), hinit[:, HID:], E, L.weight_ih_l0_reverse, L.weight_hh_l0_reverse,
                                  L.bias_ih_l0_reverse, L.bias_hh_l0_reverse).flip(0)
            out = torch.cat((Hf, Hr), dim=2).reshape(n * B, 2 * HID)
            logits = _LinearFn.apply(out, self.linear2vocab.weight, self.linear2vocab.bias)
            return logits.view(n, B, -1).permute(1, 0, 2)
        return _PolicyFn.apply(f, captions.to(dev).to(torch.int64).contiguous(), self.caption_embedding.weight,
                               self.cnn2linear.weight, self.cnn2linear.bias, self.lstm.weight_ih_l0, self.lstm.weight_hh_l0,
                               self.lstm.bias_ih_l0, self.lstm.bias_hh_l0, self.linear2vocab.weight, self.linear2vocab.bias)


class _ChainRNN(_KernelModule):
    """Common part of ValueNetworkRNN / RewardNetworkRNN: embedding + carried hidden state."""

    def _common_init(self, word_to_idx, hidden_dim, pretrained_embeddings, bidirectional):
        _unsupported(bidirectional, pretrained_embeddings)
        self.bidirectional = bidirectional
        self.hidden_dim = hidden_dim
        self.word_to_idx = word_to_idx
        self.idx_to_word = {i: w for w, i in word_to_idx.items()}
        self._rt_init()

    def _run_columns(self, captions_cm, kind):
        """Feed columns [n][B] through the serial chain from the carried state; returns the hidden
        state after every row of the LAST column, (B,512), and updates ``hidden_cell``.  The LSTM
        (value) path records autograd history, including through the carried state."""
        dev = self._dev()
        n, B = captions_cm.shape
        E = self.caption_embedding.weight
        cm = captions_cm.contiguous()
        # bidirectional (models.py:120, 215): the reverse direction walks every column bottom-up from its own carried
        # state (hidden_cell[...][1]); its outputs are flipped back to row order and concatenated -> (B, 1024)
        cm_rev = cm.flip(1).contiguous() if self.bidirectional else None
        if kind == "lstm":
            rnn = self.lstm
            h0 = self.hidden_cell[0].to(dev, torch.float32)
            c0 = self.hidden_cell[1].to(dev, torch.float32)
            h_last, h_out, c_out = _ChainLSTMFn.apply(cm, h0[0], c0[0], E, rnn.weight_ih_l0, rnn.weight_hh_l0,
                                                      rnn.bias_ih_l0, rnn.bias_hh_l0)
            if not self.bidirectional:
                self.hidden_cell = (h_out.view(1, 1, HID), c_out.view(1, 1, HID))
                return h_last
            r_last, rh_out, rc_out = _ChainLSTMFn.apply(cm_rev, h0[1], c0[1], E, rnn.weight_ih_l0_reverse,
                                                        rnn.weight_hh_l0_reverse, rnn.bias_ih_l0_reverse,
                                                        rnn.bias_hh_l0_reverse)
            self.hidden_cell = (torch.stack((h_out, rh_out)).view(2, 1, HID), torch.stack((c_out, rc_out)).view(2, 1, HID))
            return torch.cat((h_last, r_last.flip(0)), dim=1)
        rnn = self.gru
        h0 = self.hidden_cell.to(dev, torch.float32)
        h_last, h_out = _ChainGRUFn.apply(cm, h0[0], E, rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0)
        if not self.bidirectional:
            self.hidden_cell = h_out.view(1, 1, HID)
            return h_last
        r_last, rh_out = _ChainGRUFn.apply(cm_rev, h0[1], E, rnn.weight_ih_l0_reverse, rnn.weight_hh_l0_reverse,
                                           rnn.bias_ih_l0_reverse, rnn.bias_hh_l0_reverse)
        self.hidden_cell = torch.stack((h_out, rh_out)).view(2, 1, HID)
        return torch.cat((h_last, r_last.flip(0)), dim=1)

    def forward(self, captions):
        """captions (B,) -> (B,1,512): the column is a length-B sequence (models.py:130-135 / 223-228)."""
        kind = "lstm" if hasattr(self, "lstm") else "gru"
        cm = captions.reshape(1, -1).to(self._dev()).to(torch.int32)
        return self._run_columns(cm, kind).unsqueeze(1)


class ValueNetworkRNN(_ChainRNN):
    def __init__(self, word_to_idx, input_dim=512, wordvec_dim=512, hidden_dim=512,
                 pretrained_embeddings=None, bidirectional=False):
        super().__init__()
        self._common_init(word_to_idx, hidden_dim, pretrained_embeddings, bidirectional)
        self.caption_embedding, wordvec_dim = _embedding(len(word_to_idx), wordvec_dim, pretrained_embeddings)
        self.init_hidden()
        self.lstm = nn.LSTM(wordvec_dim, hidden_dim, bidirectional=bidirectional)

    def init_hidden(self):
        """models.py:122-128"""
        d = 2 if self.bidirectional else 1
        self.hidden_cell = (torch.zeros(d, 1, self.hidden_dim).to(device), torch.zeros(d, 1, self.hidden_dim).to(device))


class ValueNetwork(_KernelModule):
    """Critic (models.py:138-180)."""

    def __init__(self, word_to_idx, pretrained_embeddings=None, bidirectional=False):
        super().__init__()
        _unsupported(bidirectional, pretrained_embeddings)
        self.bidirectional = bidirectional
        self.valrnn = ValueNetworkRNN(word_to_idx, pretrained_embeddings=pretrained_embeddings,
                                      bidirectional=bidirectional)
        self.linear1 = nn.Linear(1024, 512)
        self.linear2 = nn.Linear(512, 1)
        if bidirectional:
            self.rnn_linear = nn.Linear(1024, 512)           # models.py:163-164
        self._rt_init()

    def forward(self, features, captions):
        """features (B,512), captions (B,n) -> (B,1); all n columns run through the carried-state
        chain, the head uses the h after each row of the last column (models.py:166-180)."""
        dev = self._dev()
        h = self.valrnn._run_columns(self.valrnn._tokcm(captions, dev), "lstm")
        if self.bidirectional:
            h = _LinearFn.apply(h, self.rnn_linear.weight, self.rnn_linear.bias)     # models.py:171-172
        f = features.to(dev, torch.float32).contiguous()
        return _ValueHeadFn.apply(f, h, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias)


class RewardNetworkRNN(_ChainRNN):
    def __init__(self, word_to_idx, input_dim=512, wordvec_dim=512, hidden_dim=512,
                 pretrained_embeddings=None, bidirectional=False):
        super().__init__()
        self._common_init(word_to_idx, hidden_dim, pretrained_embeddings, bidirectional)
        self.caption_embedding, wordvec_dim = _embedding(len(word_to_idx), wordvec_dim, pretrained_embeddings)
        self.init_hidden()
        self.gru = nn.GRU(wordvec_dim, hidden_dim, bidirectional=bidirectional)

    def init_hidden(self):
        """models.py:217-221"""
        self.hidden_cell = torch.zeros(2 if self.bidirectional else 1, 1, self.hidden_dim).to(device)


class RewardNetwork(_KernelModule):
    """Visual-semantic embedding (models.py:231-262); returns (ve, se), GetRewards does the cosine."""

    def __init__(self, word_to_idx, pretrained_embeddings=None, bidirectional=False):
        super().__init__()
        _unsupported(bidirectional, pretrained_embeddings)
        self.bidirectional = bidirectional
        self.rewrnn = RewardNetworkRNN(word_to_idx, pretrained_embeddings=pretrained_embeddings,
                                       bidirectional=bidirectional)
        self.visual_embed = nn.Linear(512, 512)
        self.semantic_embed = nn.Linear(1024 if bidirectional else 512, 512)     # models.py:247-251
        self._rt_init()

    def forward(self, features, captions):
        """features (B,512), captions (B,n) -> (visual embedding (B,512), semantic embedding (B,512))
        (models.py:253-262), with autograd history w.r.t. the nine reward-network parameters."""
        dev = self._dev()
        h = self.rewrnn._run_columns(self.rewrnn._tokcm(captions, dev), "gru")
        f = features.to(dev, torch.float32).contiguous()
        se = _LinearFn.apply(h, self.semantic_embed.weight, self.semantic_embed.bias)
        ve = _LinearFn.apply(f, self.visual_embed.weight, self.visual_embed.bias)
        return ve, se


class AdvantageActorCriticNetwork(nn.Module):
    """models.py:265-287 (note the argument order: value network first)."""

    def __init__(self, value_network, policy_network):
        super().__init__()
        self.value_network = value_network
        self.policy_network = policy_network

    def forward(self, features, captions):
        values = self.value_network(features, captions)
        probs = self.policy_network(features.unsqueeze(0), captions)[:, -1:, :]   # logits, despite the name
        return values, probs
