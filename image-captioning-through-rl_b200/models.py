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

Training goes through ``engine.A2CEngine`` (the fused minibatch, used by our ``trainers.py``); the
per-call ``forward`` methods here reproduce the reference call semantics (growing prefix, hidden
state carried in ``valrnn.hidden_cell`` / ``rewrnn.hidden_cell`` until ``init_hidden()``) for
inference and for step-by-step parity checks.
"""
import ctypes
import warnings

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


def _unsupported(bidirectional, pretrained_embeddings):
    # variants outside the A2C hot path (no BASELINE config uses them)
    if bidirectional:
        raise NotImplementedError("bidirectional=True is outside the B200 hot path (SURVEY.md 8f)")
    if pretrained_embeddings is not None:
        raise NotImplementedError("frozen pretrained_embeddings are outside the B200 hot path (SURVEY.md 8f)")


class _KernelModule(nn.Module):
    """Workspace + stream plumbing shared by the drop-in modules."""

    def _rt_init(self):
        object.__setattr__(self, "_ws", {})
        object.__setattr__(self, "_launches", _lib.Launches())
        object.__setattr__(self, "_warned", False)

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

    def _grad_note(self):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and not self._warned:
            object.__setattr__(self, "_warned", True)
            warnings.warn("%s.forward returns tensors without autograd history; train through "
                          "icrl_b200.trainers / A2CEngine (fused rollout + hand-written backward)"
                          % type(self).__name__)

    def _tokcm(self, captions, dev):
        caps = captions.to(dev).to(torch.int32)
        return caps.t().contiguous()            # [n][B]


class PolicyNetwork(_KernelModule):
    """Actor: LSTM over the caption prefix initialised from the image features (models.py:33-84)."""

    def __init__(self, word_to_idx, input_dim=512, wordvec_dim=512, hidden_dim=512,
                 pretrained_embeddings=None, bidirectional=False):
        super().__init__()
        _unsupported(bidirectional, pretrained_embeddings)
        if (input_dim, wordvec_dim, hidden_dim) != (HID, HID, HID):
            raise NotImplementedError("the B200 kernels are specialised for 512-wide layers (models.py:41)")
        self.bidirectional = bidirectional
        self.word_to_idx = word_to_idx
        self.idx_to_word = {i: w for w, i in word_to_idx.items()}
        vocab_size = len(word_to_idx)
        self.caption_embedding = nn.Embedding(vocab_size, wordvec_dim)
        self.cnn2linear = nn.Linear(input_dim, hidden_dim)
        self.lstm = nn.LSTM(wordvec_dim, hidden_dim, batch_first=True)
        self.linear2vocab = nn.Linear(hidden_dim, vocab_size)
        self._rt_init()

    def forward(self, features, captions):
        """features (1,B,512), captions (B,n) int64 -> logits (B,n,V) (models.py:71-84)."""
        dev = self._dev()
        self._grad_note()
        B, n = captions.shape
        V = self.linear2vocab.weight.shape[0]
        st, L = self._stream(dev), self._launches.ref
        with torch.cuda.device(dev):
            f = features.reshape(B, HID).to(dev, torch.float32).contiguous()
            table = self._buf("table", V * 4 * HID, dev=dev)
            _lib.call("icrl_pack_gate_table", st, V, 4 * HID, 4 * HID, _p(self.caption_embedding.weight),
                      _p(self.lstm.weight_ih_l0), _p(self.lstm.bias_ih_l0), _p(self.lstm.bias_hh_l0), _p(table), L)
            tokcm = self._buf("tokcm", (n + 1) * B, torch.int32, dev)
            tokcm[:B] = captions[:, 0].to(dev).to(torch.int32)
            forced = torch.zeros((B, n), dtype=torch.int64, device=dev)
            if n > 1:
                forced[:, :n - 1] = captions[:, 1:].to(dev)
            tokens = torch.empty((B, n), dtype=torch.int64, device=dev)
            logp = torch.empty((B, n), dtype=torch.float32, device=dev)
            logits = torch.empty((n, B, V), dtype=torch.float32, device=dev)
            _lib.call("icrl_policy_rollout_fwd", st, B, V, 1, n, 0, _p(f), _p(self.cnn2linear.weight),
                      _p(self.cnn2linear.bias), _p(table), _p(self.lstm.weight_hh_l0), _p(self.linear2vocab.weight),
                      _p(self.linear2vocab.bias), None, _p(forced), _p(tokcm), _p(tokens), _p(logp),
                      _p(self._buf("Hs", (n + 1) * B * HID, dev=dev)), _p(self._buf("Cs", (n + 1) * B * HID, dev=dev)),
                      _p(self._buf("Gs", n * B * 4 * HID, dev=dev)), _p(logits), _p(self._buf("gpre", B * 4 * HID, dev=dev)), L)
        return logits.permute(1, 0, 2)


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
        state after every row of the LAST column, (B,512), and updates ``hidden_cell``."""
        dev = self._dev()
        n, B = captions_cm.shape
        V = self.caption_embedding.weight.shape[0]
        st, L = self._stream(dev), self._launches.ref
        with torch.cuda.device(dev):
            rnn = self.lstm if kind == "lstm" else self.gru
            G = 4 * HID if kind == "lstm" else 3 * HID
            table = self._buf("table", V * G, dev=dev)
            _lib.call("icrl_pack_gate_table", st, V, G, G if kind == "lstm" else 2 * HID,
                      _p(self.caption_embedding.weight), _p(rnn.weight_ih_l0), _p(rnn.bias_ih_l0), _p(rnn.bias_hh_l0),
                      _p(table), L)
            T = n * B
            stream = captions_cm.reshape(-1).contiguous()
            stash_h = self._buf("stash_h", (T + 1) * HID, dev=dev)
            sync = self._ws.get("sync")
            if sync is None or sync.device != dev:
                sync = torch.zeros(int(_lib.call("icrl_chain_sync_bytes")), dtype=torch.uint8, device=dev)
                self._ws["sync"] = sync
            h_out = torch.empty(HID, dtype=torch.float32, device=dev)
            if kind == "lstm":
                h0 = self.hidden_cell[0].to(dev, torch.float32).reshape(-1).contiguous()
                c0 = self.hidden_cell[1].to(dev, torch.float32).reshape(-1).contiguous()
                c_out = torch.empty(HID, dtype=torch.float32, device=dev)
                _lib.call("icrl_chain_lstm_fwd", st, _p(stream), T, _p(table), _p(rnn.weight_hh_l0), _p(h0), _p(c0),
                          _p(stash_h), None, None, _p(h_out), _p(c_out), _p(sync), L)
                self.hidden_cell = (h_out.view(1, 1, HID), c_out.view(1, 1, HID))
            else:
                h0 = self.hidden_cell.to(dev, torch.float32).reshape(-1).contiguous()
                _lib.call("icrl_chain_gru_fwd", st, _p(stream), T, _p(table), _p(rnn.weight_hh_l0),
                          _p(rnn.bias_hh_l0[2 * HID:]), _p(h0), _p(stash_h), _p(h_out), _p(sync), L)
                self.hidden_cell = h_out.view(1, 1, HID)
            _lib.call("icrl_chain_check", st, _p(sync))
            return stash_h[(T - B + 1) * HID:(T + 1) * HID].view(B, HID).clone()

    def forward(self, captions):
        """captions (B,) -> (B,1,512): the column is a length-B sequence (models.py:130-135 / 223-228)."""
        self._grad_note()
        kind = "lstm" if hasattr(self, "lstm") else "gru"
        cm = captions.reshape(1, -1).to(self._dev()).to(torch.int32)
        return self._run_columns(cm, kind).unsqueeze(1)


class ValueNetworkRNN(_ChainRNN):
    def __init__(self, word_to_idx, input_dim=512, wordvec_dim=512, hidden_dim=512,
                 pretrained_embeddings=None, bidirectional=False):
        super().__init__()
        self._common_init(word_to_idx, hidden_dim, pretrained_embeddings, bidirectional)
        self.caption_embedding = nn.Embedding(len(word_to_idx), wordvec_dim)
        self.init_hidden()
        self.lstm = nn.LSTM(wordvec_dim, hidden_dim)

    def init_hidden(self):
        """models.py:122-128"""
        self.hidden_cell = (torch.zeros(1, 1, self.hidden_dim).to(device), torch.zeros(1, 1, self.hidden_dim).to(device))


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
        self._rt_init()

    def forward(self, features, captions):
        """features (B,512), captions (B,n) -> (B,1); all n columns run through the carried-state
        chain, the head uses the h after each row of the last column (models.py:166-180)."""
        dev = self._dev()
        self._grad_note()
        B = captions.shape[0]
        h = self.valrnn._run_columns(self.valrnn._tokcm(captions, dev), "lstm")
        st, L = self._stream(dev), self._launches.ref
        with torch.cuda.device(dev):
            f = features.to(dev, torch.float32).contiguous()
            weff, beff = self._buf("weff", 2 * HID, dev=dev), self._buf("beff", 1, dev=dev)
            _lib.call("icrl_pack_value_head", st, _p(self.linear1.weight), _p(self.linear1.bias), _p(self.linear2.weight),
                      _p(self.linear2.bias), _p(weff), _p(beff), L)
            values = torch.empty((B, 1), dtype=torch.float32, device=dev)
            _lib.call("icrl_value_head_fwd", st, B, 1, _p(f), _p(h), _p(weff), _p(beff), _p(values), L)
        return values


class RewardNetworkRNN(_ChainRNN):
    def __init__(self, word_to_idx, input_dim=512, wordvec_dim=512, hidden_dim=512,
                 pretrained_embeddings=None, bidirectional=False):
        super().__init__()
        self._common_init(word_to_idx, hidden_dim, pretrained_embeddings, bidirectional)
        self.caption_embedding = nn.Embedding(len(word_to_idx), wordvec_dim)
        self.init_hidden()
        self.gru = nn.GRU(wordvec_dim, hidden_dim)

    def init_hidden(self):
        """models.py:217-221"""
        self.hidden_cell = torch.zeros(1, 1, self.hidden_dim).to(device)


class RewardNetwork(_KernelModule):
    """Visual-semantic embedding (models.py:231-262); returns (ve, se), GetRewards does the cosine."""

    def __init__(self, word_to_idx, pretrained_embeddings=None, bidirectional=False):
        super().__init__()
        _unsupported(bidirectional, pretrained_embeddings)
        self.bidirectional = bidirectional
        self.rewrnn = RewardNetworkRNN(word_to_idx, pretrained_embeddings=pretrained_embeddings,
                                       bidirectional=bidirectional)
        self.visual_embed = nn.Linear(512, 512)
        self.semantic_embed = nn.Linear(512, 512)
        self._rt_init()

    def forward(self, features, captions):
        dev = self._dev()
        self._grad_note()
        B = captions.shape[0]
        h = self.rewrnn._run_columns(self.rewrnn._tokcm(captions, dev), "gru")
        st, L = self._stream(dev), self._launches.ref
        with torch.cuda.device(dev):
            f = features.to(dev, torch.float32).contiguous()
            se = torch.empty((B, HID), dtype=torch.float32, device=dev)
            ve = torch.empty((B, HID), dtype=torch.float32, device=dev)
            for x, lin, out in ((h, self.semantic_embed, se), (f, self.visual_embed, ve)):
                _lib.call("icrl_gemm_f32", st, 0, 1, B, HID, HID, _p(x), HID, _p(lin.weight), HID, _p(out), HID,
                          _p(lin.bias), 0.0, None, 0, L)
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
