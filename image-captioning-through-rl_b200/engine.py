"""Fused A2C minibatch on the GPU: one call replaces the body of the reference's training loop
(``trainers.py:432-480`` / ``:544-594``): rollout -> rewards -> values -> loss -> backward.

Host code is PyTorch plumbing only (device memory, streams); all arithmetic runs in the CUDA
kernels behind ``include/icrl_b200.h``.  There is no CPU fallback.

Single-pass formulation (verified against the unmodified reference):
  policy   one incremental LSTM pass with sampling (instead of re-running the prefix every step)
  value    ONE serial LSTM chain over the column-major token stream, h taken at the last B
           positions of each step block; head evaluated in collapsed form.  By default the chain
           is advanced as up to 32 consecutive pieces in lockstep, each with a discarded warm-up
           that is verified against the previous piece on every step (chain segments)
  reward   ONE serial GRU chain, semantic/visual embeds batched over all steps, fused cosine
  loss     advantage = values - rewards, gradient seeds for both networks
  backward policy BPTT (batched GEMMs + S serial cell steps), value-chain BPTT kernel, weight
           gradients as contractions over all steps; gradients land in one flat bucket that the
           parameters' ``.grad`` tensors alias (ready for a single all-reduce).
"""
import ctypes

import numpy as np
import torch

from . import _lib

H = 512
END_TOKEN = 2       # trainers.py:436


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class StepResult(dict):
    """Device tensors of one minibatch; ``loss`` / ``mean_reward`` / ``mean_adv`` sync lazily."""

    host_stats = None       # (loss, mean reward, mean advantage) when the step's check already brought them to the host

    def _stat(self, i):
        if self.host_stats is not None:
            return float(self.host_stats[i])
        return float(self["stats"][i].item())

    @property
    def loss(self):
        return self._stat(0)

    @property
    def mean_reward(self):
        return self._stat(1)

    @property
    def mean_adv(self):
        return self._stat(2)


class Prepared:
    """A minibatch already resident in HBM (A2CEngine.prepare): features (B,512) f32, the prefix
    columns int32 [p0][B], uniforms (S,B) f64 or None, and the rollout plan."""

    def __init__(self, f, prefix_cm, u, B, p0, S):
        self.f, self.prefix_cm, self.u, self.B, self.p0, self.S = f, prefix_cm, u, B, p0, S


def plan_rollout(captions, level=None):
    """(p0, S) as the reference derives them: caplen = max <END> column + 1 (trainers.py:436);
    non-curriculum p0 = 1, S = caplen - 1 (:438-441); curriculum p0 = caplen - level, S = level
    (:548-554).  Raises like the reference when no row holds <END>."""
    caps = np.asarray(captions)
    cols = np.nonzero(caps == END_TOKEN)[1]
    if cols.size == 0:
        raise ValueError("no <END> (=2) token in the caption batch (trainers.py:436 would fail on max())")
    caplen = int(cols.max()) + 1
    if level is None:
        return 1, caplen - 1
    return caplen - int(level), int(level)


class A2CEngine:
    DECODE_MODES = ("fused", "tc", "simt")

    CHAIN_ENGINES = ("tc", "simt")

    def __init__(self, a2c_network, reward_network, use_tc=None, decode="fused", chain_shards=1, wgrad="tc",
                 chain_segments=32, chain_warmup=256, chain_tol=1e-5, chain_bwd_segments=None, chain_engine="tc",
                 chain_pieces=None, chain_adapt=True, chain_warmup_min=32, policy_bptt="tc", chain_fuse_fwd=True, overlap_backward=True):
        self.policy = a2c_network.policy_network
        self.value = a2c_network.value_network
        self.reward = reward_network
        self.a2c = a2c_network
        if any(getattr(m, "bidirectional", False) for m in (self.policy, self.value, self.reward)):
            raise NotImplementedError("the fused engine is built for the unidirectional networks; bidirectional=True runs "
                                      "on the module (autograd) route -- icrl_b200.trainers switches to it by itself")
        dev = self.policy.linear2vocab.weight.device
        if dev.type != "cuda":
            raise _lib.IcrlError("A2CEngine needs the networks on a CUDA device (no CPU fallback)")
        _lib.load()
        self.device = dev
        self.V = self.policy.linear2vocab.weight.shape[0]
        self._bufs = {}
        self.launches = _lib.Launches()
        self.phase_events = None          # set to [] to record (name, start, end) CUDA events per phase
        # policy decode path: "fused" = the persistent cluster kernel (decode.cu: tcgen05 GEMMs with the cell
        # update / softmax / sampling in their epilogues, one launch per rollout); "tc" = per-step kernels with
        # the two GEMMs on tcgen05 (gemm_tc.cu); "simt" = per-step kernels, fp32 CUDA-core GEMMs.
        if use_tc is not None:
            decode = "tc" if use_tc else "simt"
        if decode not in self.DECODE_MODES:
            raise ValueError("decode must be one of %s" % (self.DECODE_MODES,))
        if decode == "fused" and (self.V > 1024 or self.V % 4):
            # the persistent kernel keeps one row of logits inside a cluster (8 x 128 columns, float4 stash)
            import warnings
            warnings.warn("vocabulary of %d words does not fit the fused decode kernel (V <= 1024, V %% 4 == 0): "
                          "using the per-step kernels" % self.V)
            decode = "simt"
        self.decode = decode
        self.use_tc = decode == "tc"
        # chain_shards = K > 1 cuts the (local) batch into K contiguous row shards whose value / reward recurrences
        # each start from zero state -- the reference run on K minibatches of B/K rows with averaged gradients,
        # i.e. exactly what K data-parallel ranks compute.  K = 1 is the reference's single
        # carried-state chain over the whole batch.  The K chains advance in lockstep on the same CTAs.
        # wgrad: "tc" = the K = T weight-gradient contraction of the value chain on tcgen05 (wgrad_tc.cu; needs a
        # workspace of ~10 KB per serial step), "simt" = fp32 CUDA-core GEMM with split-K.
        if wgrad not in ("tc", "simt"):
            raise ValueError("wgrad must be 'tc' or 'simt'")
        self.wgrad = wgrad
        if chain_shards not in (1, 2, 4, 8):
            raise ValueError("chain_shards must be 1, 2, 4 or 8")
        self.chain_shards = int(chain_shards)
        # chain_segments = K > 1 keeps the reference's ONE carried-state chain but advances K consecutive pieces of it in
        # lockstep: piece k >= 1 starts from zero state `chain_warmup` positions early and throws those steps away.  A
        # gated recurrence forgets its initial state, so the piece then carries the single chain's state up to float
        # rounding.  That is verified on every step (state at the end of each warm-up vs the state the previous piece
        # computes at the same position; gate gradients at the joints for the backward recurrence); a step whose check
        # exceeds chain_tol is re-run on the serial kernels and the warm-up is lengthened.  Chains too short for K
        # pieces of >= 2 warm-ups use fewer pieces or the serial kernels.  chain_segments = 1: always serial.
        # More than 8 pieces run as two chunks inside a kernel step (16 = 2 x 8, 32 = 2 x 16): a chunk's exchange round
        # trip is covered by the arithmetic of the other chunk.  The backward recurrence uses at most 16 pieces (8 per
        # CTA group; every 2nd forward joint when the forward has 32).
        if chain_segments not in (1, 2, 4, 8, 16, 32):
            raise ValueError("chain_segments must be 1, 2, 4, 8, 16 or 32")
        # chain_bwd_segments: backward pieces (None = 16 when the forward has 16 or 32, else min(forward, 8)); experiments.
        if chain_bwd_segments not in (None, 2, 4, 8, 16):
            raise ValueError("chain_bwd_segments must be None, 2, 4, 8 or 16")
        self.chain_bwd_segments = chain_bwd_segments
        self.chain_segments = 1 if self.chain_shards > 1 else int(chain_segments)
        self.chain_warmup = int(chain_warmup)
        self.chain_tol = float(chain_tol)
        if self.chain_warmup < 1:
            raise ValueError("chain_warmup must be positive")
        # chain_engine: "tc" (default) = chain pieces on tcgen05 (chain_tc.cu): hundreds of lockstep pieces, 128 per cluster of
        # 8 CTAs, one kernel step = H_prev [pieces x 512] . W_hh^T on the tensor cores.  The warm-up is sized PER CHAIN
        # (value LSTM forward + backward, reward GRU) from the contraction rate measured on every step (errors half-way
        # through and at the end of the warm-up): it grows before the check can fail and shrinks again after
        # `_SHRINK_AFTER` clean steps; a failed check re-runs the chains with a longer warm-up, and only a chain that
        # cannot be cut at all runs on the serial kernels.  "simt" = the CUDA-core segment kernels of chain.cu
        # (chain_segments pieces, fixed warm-up, serial re-run on failure).  chain_segments = 1 forces the serial kernels.
        if chain_engine not in self.CHAIN_ENGINES:
            raise ValueError("chain_engine must be one of %s" % (self.CHAIN_ENGINES,))
        self.chain_engine = chain_engine
        self.chain_pieces = None if chain_pieces is None else int(chain_pieces)
        if self.chain_pieces is not None and self.chain_pieces < 2:
            raise ValueError("chain_pieces must be at least 2")
        # policy_bptt: "tc" = the policy's serial cell-backward steps on the tcgen05 chain-backward kernel (one launch),
        # "simt" = one pointwise kernel + one fp32 CUDA-core GEMM per cell step
        if policy_bptt not in ("tc", "simt"):
            raise ValueError("policy_bptt must be 'tc' or 'simt'")
        self.policy_bptt = policy_bptt
        # overlap_backward: the policy's backward on a second stream beside the value-chain backward (separate workspaces)
        self.overlap_backward = bool(overlap_backward)
        self._side_stream = None
        self._n_cell_max = 1
        self.chain_fuse_fwd = bool(chain_fuse_fwd)     # value + reward forward chains in one launch when that is faster
        self.chain_adapt = bool(chain_adapt)
        self.chain_warmup_min = int(chain_warmup_min)
        # current warm-up per recurrence (tc engine): value forward, reward forward, value backward -- each follows its own
        # measured contraction (the forward cell state is the slowest to converge; the backward need not pay for it)
        self.warm = {"v": self.chain_warmup, "r": self.chain_warmup, "b": self.chain_warmup}
        self._clean = {"v": 0, "r": 0, "b": 0}
        self._hold = {"v": 0, "r": 0, "b": 0}
        self._tc = None                   # {"v": (P, seg, warm) | None, "r": (P, seg, warm)} of the current step
        self.segment_stats = {"steps": 0, "segmented_steps": 0, "fallbacks": 0, "reruns": 0, "max_err": [0.0] * 5,
                              "tc_max_err": [0.0] * 16, "warm_history": []}
        self._seg = None                  # (K, seg_v, seg_r, warm) of the current step, None = serial kernels
        self._seg_strikes = 0
        self._seg_unverified = False
        with torch.cuda.device(dev):
            self._seg_ws = torch.zeros(int(_lib.call("icrl_chain_segment_ws_floats")), dtype=torch.float32, device=dev)
            self.sync_state = torch.zeros(int(_lib.call("icrl_chain_sync_bytes")), dtype=torch.uint8, device=dev)
            # [0:3] loss, mean reward, mean advantage; [4:20] joint-check words of the tc chain launches (value forward
            # 4..7, reward forward 8..11, value backward 12..17): read together in ONE device-to-host copy per step
            # [20:28] words of the policy's tcgen05 BPTT ([24] max |dL/dh|, [25] fp16-exchange overflow flag)
            self._stat_err = torch.zeros(28, dtype=torch.float32, device=dev)
        self._tc_err = self._stat_err[4:20]
        self._reward_versions = None
        self._dz_ld = None
        self._check_params()
        self._bind_flat_grads()

    # ------------------------------------------------------------------ parameters / gradients
    def _params(self):
        return list(self.a2c.parameters())

    def _check_params(self):
        for p in list(self.a2c.parameters()) + list(self.reward.parameters()):
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != self.device:
                raise _lib.IcrlError("parameters must be contiguous float32 on %s" % self.device)

    def _bind_flat_grads(self):
        """One flat fp32 bucket in a2c.parameters() order; every .grad is a view into it."""
        ps = [p for p in self._params() if p.requires_grad]
        # every tensor starts on a 256-byte boundary of the bucket (the kernels read parameters and write gradients with
        # 16-byte accesses; optim.FlatAdam lays the parameters out the same way): the padding floats stay zero
        self._flat_offsets, off = [], 0
        for p in ps:
            self._flat_offsets.append(off)
            off += -(-p.numel() // 64) * 64
        self.flat_grad = torch.zeros(off, dtype=torch.float32, device=self.device)
        self._grad_views = []
        for p, o in zip(ps, self._flat_offsets):
            self._grad_views.append((p, self.flat_grad[o:o + p.numel()].view_as(p)))
        self._attach_grads()

    def _attach_grads(self):
        for p, v in self._grad_views:
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def _g(self, p, optional=False):
        """Gradient destination for parameter p (frozen: a scratch buffer, or None where the ABI lets a gradient
        be skipped -- the frozen pretrained embeddings, models.py:61-63)."""
        for q, v in self._grad_views:
            if q is p:
                return v
        if optional:
            return None
        return self._buf("frozen_grad_%d" % id(p), p.numel())

    def _buf(self, name, numel, dtype=torch.float32):
        numel = int(max(numel, 1))
        t = self._bufs.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(numel, dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    class _Phase:
        def __init__(self, eng, name):
            self.eng, self.name = eng, name

        def __enter__(self):
            if self.eng.phase_events is not None:
                self.t0 = torch.cuda.Event(enable_timing=True)
                self.t0.record(torch.cuda.current_stream(self.eng.device))

        def __exit__(self, *a):
            if self.eng.phase_events is not None:
                t1 = torch.cuda.Event(enable_timing=True)
                t1.record(torch.cuda.current_stream(self.eng.device))
                self.eng.phase_events.append((self.name, self.t0, t1))

    def _phase(self, name):
        return A2CEngine._Phase(self, name)

    def phase_times_ms(self):
        """Sum of CUDA-event durations per phase since phase_events was last reset (syncs)."""
        torch.cuda.synchronize(self.device)
        out = {}
        for name, a, b in self.phase_events or []:
            out.setdefault(name, []).append(a.elapsed_time(b))
        return out

    @property
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ packing
    def pack_weights(self, reward=True):
        """Rebuild the derived operands from the current parameters (after every optimizer step):
        the three gate tables W_ih E + b and the collapsed value head."""
        st, L, V = self._stream, self.launches.ref, self.V
        P, Vn, R = self.policy, self.value, self.reward
        c = _lib.call
        c("icrl_pack_gate_table", st, V, 4 * H, 4 * H, P.caption_embedding.weight.shape[1], _p(P.caption_embedding.weight), _p(P.lstm.weight_ih_l0),
          _p(P.lstm.bias_ih_l0), _p(P.lstm.bias_hh_l0), _p(self._buf("p_table", V * 4 * H)), L)
        c("icrl_pack_gate_table", st, V, 4 * H, 4 * H, Vn.valrnn.caption_embedding.weight.shape[1], _p(Vn.valrnn.caption_embedding.weight),
          _p(Vn.valrnn.lstm.weight_ih_l0), _p(Vn.valrnn.lstm.bias_ih_l0), _p(Vn.valrnn.lstm.bias_hh_l0),
          _p(self._buf("v_table", V * 4 * H)), L)
        c("icrl_pack_value_head", st, _p(Vn.linear1.weight), _p(Vn.linear1.bias), _p(Vn.linear2.weight),
          _p(Vn.linear2.bias), _p(self._buf("v_weff", 2 * H)), _p(self._buf("v_beff", 1)), L)
        if self.use_tc:
            self._pack_policy_tc()
        if self.decode == "fused":
            n = int(_lib.call("icrl_decode_weight_halves"))
            _lib.call("icrl_pack_decode_weights", st, V, _p(P.lstm.weight_hh_l0), _p(P.linear2vocab.weight),
                      _p(self._buf("p_decode_pk", n, torch.float16)), L)
        if self._use_tc_chains():
            n = int(_lib.call("icrl_chain_tc_weight_halves", 0))
            _lib.call("icrl_pack_chain_tc_weights", st, 0, _p(Vn.valrnn.lstm.weight_hh_l0),
                      _p(self._buf("v_chain_pk", n, torch.float16)), L)
        if self.policy_bptt == "tc":
            n = int(_lib.call("icrl_chain_tc_weight_halves", 0))
            _lib.call("icrl_pack_chain_tc_weights", st, 0, _p(P.lstm.weight_hh_l0),
                      _p(self._buf("p_chain_pk", n, torch.float16)), L)
        if reward or self._reward_changed():
            self.pack_reward()

    def _use_tc_chains(self):
        return self.chain_engine == "tc" and self.chain_segments > 1 and self.chain_shards == 1

    def _reward_changed(self):
        """The reward network is frozen during A2C training (trainers.py:372-373) and its derived operands are packed once;
        an in-place change of its parameters (load_state_dict, more reward pretraining on the same object) is noticed
        through the tensors' version counters."""
        v = tuple(p._version for p in self.reward.parameters()) + tuple(p.data_ptr() for p in self.reward.parameters())
        return v != self._reward_versions

    def _pack_policy_tc(self):
        """3-part bf16 splits of the two decode-step weight matrices (tensor-core operands)."""
        P = self.policy
        for name, w in (("whh", P.lstm.weight_hh_l0), ("wv", P.linear2vocab.weight)):
            n = w.numel()
            _lib.call("icrl_split_bf16x3", self._stream, n, _p(w), _p(self._buf("p_%s_parts" % name, 3 * n, torch.bfloat16)),
                      self.launches.ref)

    def pack_reward(self):
        R, V = self.reward, self.V
        _lib.call("icrl_pack_gate_table", self._stream, V, 3 * H, 2 * H, R.rewrnn.caption_embedding.weight.shape[1], _p(R.rewrnn.caption_embedding.weight),
                  _p(R.rewrnn.gru.weight_ih_l0), _p(R.rewrnn.gru.bias_ih_l0), _p(R.rewrnn.gru.bias_hh_l0),
                  _p(self._buf("r_table", V * 3 * H)), self.launches.ref)
        if self._use_tc_chains():
            n = int(_lib.call("icrl_chain_tc_weight_halves", 1))
            _lib.call("icrl_pack_chain_tc_weights", self._stream, 1, _p(R.rewrnn.gru.weight_hh_l0),
                      _p(self._buf("r_chain_pk", n, torch.float16)), self.launches.ref)
        # 3-part bf16 split of the sentence-embedding weight: the [S*B x 512] x [512 x 512] projection of the reward head
        # runs on the tensor-core GEMM of the decode step (gemm_tc.cu, fp32-grade) once it is large enough to fill tiles
        n = R.semantic_embed.weight.numel()
        _lib.call("icrl_split_bf16x3", self._stream, n, _p(R.semantic_embed.weight),
                  _p(self._buf("r_se_parts", 3 * n, torch.bfloat16)), self.launches.ref)
        self._reward_versions = tuple(p._version for p in R.parameters()) + tuple(p.data_ptr() for p in R.parameters())

    _TC_LINEAR_MIN_ROWS = 4096

    def _gemm(self, ta, tb, M, N, K, A, lda, B, ldb, C, ldc, bias=None, beta=0.0):
        _lib.call("icrl_gemm_f32", self._stream, ta, tb, M, N, K, _p(A), lda, _p(B), ldb, _p(C), ldc, _p(bias),
                  beta, None, 0, self.launches.ref)

    # ------------------------------------------------------------------ inputs
    def prepare(self, features, captions, uniforms=None, level=None, plan=None):
        """Host -> HBM staging of one minibatch (the copies the end-to-end timing includes)."""
        p0, S = plan if plan is not None else plan_rollout(captions, level)
        if p0 < 1:
            return None
        dev = self.device
        caps = captions.cpu().numpy() if isinstance(captions, torch.Tensor) else np.asarray(captions)
        B = caps.shape[0]
        f = torch.as_tensor(features)
        if f.dtype != torch.float32:
            f = f.float()
        f = f.to(dev, non_blocking=True).contiguous()
        pre = torch.from_numpy(np.ascontiguousarray(caps[:, :p0].T).astype(np.int32))
        prefix_cm = pre.to(dev, non_blocking=True)
        u = None
        if uniforms is not None:
            u = torch.as_tensor(uniforms, dtype=torch.float64).to(dev, non_blocking=True).contiguous()
            assert tuple(u.shape) == (S, B), "uniforms must be (S,B) float64"
        return Prepared(f, prefix_cm, u, B, p0, S)

    def gather_rows(self, matrix, rows):
        """matrix[rows] on the device: `matrix` (M,512) f32 resident in HBM, `rows` host integers (only they cross the
        bus).  The device-side half of the reference's minibatch generator (utilities.py:172-176)."""
        idx = torch.as_tensor(np.asarray(rows), dtype=torch.int32).to(self.device, non_blocking=True)
        out = torch.empty((idx.numel(), H), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("icrl_gather_rows", self._stream, idx.numel(), _p(matrix), _p(idx), 0, _p(out), self.launches.ref)
        return out

    def _stage_inputs(self, prep, forced):
        B, p0, S = prep.B, prep.p0, prep.S
        tokcm = self._buf("tokcm", (p0 + S) * B, torch.int32)
        tokcm[:p0 * B].copy_(prep.prefix_cm.reshape(-1), non_blocking=True)
        fo = None
        if forced is not None:
            fo = torch.as_tensor(np.asarray(forced), dtype=torch.int64).to(self.device).contiguous()
            assert tuple(fo.shape) == (B, S)
        return prep.f, tokcm, prep.u, fo, B

    # ------------------------------------------------------------------ phases
    def _policy_forward(self, f, tokcm, u, fo, B, p0, S, greedy):
        V, P = self.V, self.policy
        n_cell = p0 - 1 + S
        tokens = torch.empty((B, S), dtype=torch.int64, device=self.device)
        logp = torch.empty((B, S), dtype=torch.float32, device=self.device)
        Hs = self._buf("p_Hs", (n_cell + 1) * B * H)
        Cs = self._buf("p_Cs", (n_cell + 1) * B * H)
        Gs = self._buf("p_Gs", n_cell * B * 4 * H)
        logits = self._buf("p_logits", S * B * V)
        gpre = self._buf("p_gpre", B * 4 * H)
        b = self._bufs
        with self._phase("policy_fwd"):
          if self.decode == "fused":
            _lib.call("icrl_policy_rollout_fwd_fused", self._stream, B, V, p0, S, int(bool(greedy)), _p(f),
                  _p(P.cnn2linear.weight), _p(P.cnn2linear.bias), _p(b["p_table"]), _p(b["p_decode_pk"]),
                  _p(P.linear2vocab.bias), _p(u), _p(fo), _p(tokcm), _p(tokens), _p(logp), _p(Hs), _p(Cs), _p(Gs),
                  _p(logits), None, _p(self._buf("p_hparts", 4 * B * H, torch.float16)), self.launches.ref)
          elif self.use_tc:
            _lib.call("icrl_policy_rollout_fwd_tc", self._stream, B, V, p0, S, int(bool(greedy)), _p(f),
                  _p(P.cnn2linear.weight), _p(P.cnn2linear.bias), _p(b["p_table"]), _p(b["p_whh_parts"]),
                  _p(b["p_wv_parts"]), _p(P.linear2vocab.bias), _p(u), _p(fo), _p(tokcm), _p(tokens),
                  _p(logp), _p(Hs), _p(Cs), _p(Gs), _p(logits), _p(gpre),
                  _p(self._buf("p_h_parts", 3 * B * H, torch.bfloat16)), self.launches.ref)
          else:
            _lib.call("icrl_policy_rollout_fwd", self._stream, B, V, p0, S, int(bool(greedy)), _p(f),
                  _p(P.cnn2linear.weight), _p(P.cnn2linear.bias), _p(b["p_table"]),
                  _p(P.lstm.weight_hh_l0), _p(P.linear2vocab.weight), _p(P.linear2vocab.bias), _p(u), _p(fo),
                  _p(tokcm), _p(tokens), _p(logp), _p(Hs), _p(Cs), _p(Gs), _p(logits), _p(gpre), self.launches.ref)
        return tokens, logp

    def _streams(self, tokcm, B, p0, S, serial=False):
        st, L = self._stream, self.launches.ref
        K = self.chain_shards
        i32 = torch.int32
        if K > 1:
            if B % K:
                raise ValueError("batch of %d rows does not split into %d chain shards" % (B, K))
            Tv = int(_lib.call("icrl_stream_len", B // K, p0, S, 0))
            Tr = int(_lib.call("icrl_stream_len", B // K, p0, S, 1))
            v_stream, v_take = self._buf("v_stream", K * (Tv + 1), i32), self._buf("v_take", K * (Tv + 1), i32)
            r_stream = self._buf("r_stream", K * (Tr + 1), i32)
            v_pos, r_pos = self._buf("v_pos", S * B, i32), self._buf("r_pos", S * B, i32)
            _lib.call("icrl_build_stream_sharded", st, B, p0, S, 0, K, _p(tokcm), _p(v_stream), _p(v_take), _p(v_pos), L)
            _lib.call("icrl_build_stream_sharded", st, B, p0, S, 1, K, _p(tokcm), _p(r_stream), None, _p(r_pos), L)
            return Tv, Tr
        Tv = int(_lib.call("icrl_stream_len", B, p0, S, 0))
        Tr = int(_lib.call("icrl_stream_len", B, p0, S, 1))
        self._seg, self._tc = None, None
        if not serial:
            if self._use_tc_chains():
                self._tc = self._pick_pieces(Tv, Tr)
            elif self.chain_segments > 1:
                self._seg = self._pick_segments(Tv, Tr)
        nv, nr = self._padded(Tv, 0), self._padded(Tr, 1)
        v_stream, v_take, v_pos = self._buf("v_stream", nv, i32), self._buf("v_take", nv, i32), self._buf("v_pos", S * B, i32)
        r_stream, r_pos = self._buf("r_stream", nr, i32), self._buf("r_pos", S * B, i32)
        _lib.call("icrl_build_stream", st, B, p0, S, 0, _p(tokcm), _p(v_stream), _p(v_take), _p(v_pos), L)
        _lib.call("icrl_build_stream", st, B, p0, S, 1, _p(tokcm), _p(r_stream), None, _p(r_pos), L)
        if self._seg is not None or self._tc is not None:   # the last piece runs past the end of the chain: pad with token 0 / "no output"
            v_stream[Tv:nv].zero_()
            v_take[Tv:nv].fill_(-1)
            r_stream[Tr:nr].zero_()
        return Tv, Tr

    def _pick_segments(self, Tv, Tr):
        """(K, seg_v, seg_r, warm) for chains of Tv / Tr positions, or None when they are too short."""
        warm = self.chain_warmup
        for K in (32, 16, 8, 4, 2):
            if K > self.chain_segments:
                continue
            seg_r = int(_lib.call("icrl_chain_segment_len", Tr, K, warm))
            seg_v = int(_lib.call("icrl_chain_segment_len", Tv, K, warm)) if Tv > 0 else 0
            if seg_r > 0 and (Tv == 0 or seg_v > 0):
                return (K, seg_v, seg_r, warm)
        return None

    def _pieces_for(self, T, warm, pmax=None):
        """(P, seg) for a chain of T positions cut into tcgen05 pieces with a `warm`-position warm-up, or None when it is
        too short.  As many pieces as are co-resident (whole clusters of 128 once there is more than one), each at least
        max(warm / 4, 8) live positions long: the wall time of a launch is seg + warm kernel steps."""
        pmax = self.chain_pieces or pmax or int(_lib.call("icrl_chain_tc_max_pieces"))
        min_seg = max(warm // 4, 8)
        if T <= warm + 2 * min_seg:
            return None
        P = min(pmax, (T - warm) // min_seg)
        if P > 128 and self.chain_pieces is None:
            P -= P % 128
        if P < 2:
            return None
        return P, -(-(T - warm) // P)

    # relative cost of one kernel step of the value LSTM / reward GRU forward kernels (measured: 31.1 K / 22.2 K cycles at 128
    # pieces per cluster); only their ratio matters for splitting the clusters of the fused forward launch
    _STEP_COST = {"v": 1.4, "r": 1.0}

    def _pick_pieces(self, Tv, Tr):
        """Piece layout of the two forward chains.  Sequential: each chain gets every co-resident cluster in its own launch.
        Fused ("fused": True): ONE launch, the clusters split so that the two chains end together -- fewer, longer pieces
        each, but the discarded warm-up is paid once in wall time; chosen when the cost model says it is faster (always in
        the warm-up-dominated regime of small per-rank batches)."""
        wv, wr = self.warm["v"], self.warm["r"]
        r = self._pieces_for(Tr, wr)
        v = self._pieces_for(Tv, wv) if Tv > 0 else None
        if r is None or (Tv > 0 and v is None):
            return None
        # "b": the backward recurrence of the value chain always gets every co-resident cluster (its own launch); it may
        # be cut differently from the forward as long as the arrays cover both layouts (the stash rows past the forward's
        # end are zeroed: padding positions, whose gate gradients are exactly zero)
        wb = self.warm["b"]
        bw = None if v is None else self._pieces_for(Tv, wb, int(_lib.call("icrl_chain_tc_bwd_max_pieces")))
        if v is not None and bw is None:
            bw, wb = v, wv
        lay = {"v": None if v is None else (v[0], v[1], wv), "r": (r[0], r[1], wr), "fused": False,
               "b": None if bw is None else (bw[0], bw[1], wb)}
        if v is None or not self.chain_fuse_fwd or self.chain_pieces is not None:
            return lay
        C = int(_lib.call("icrl_chain_tc_max_pieces")) // 128
        cv, cr = self._STEP_COST["v"], self._STEP_COST["r"]
        best = (cv * (v[1] + wv) + cr * (r[1] + wr), None)
        for nv in range(1, C):
            Pv, Pr = min(128 * nv, v[0]), min(128 * (C - nv), r[0])
            if Pv < 2 or Pr < 2:
                continue
            sv, sr = -(-(Tv - wv) // Pv), -(-(Tr - wr) // Pr)
            t = max(cv * (sv + wv), cr * (sr + wr))
            if t < best[0]:
                best = (t, (Pv, sv, Pr, sr))
        if best[1] is not None:
            Pv, sv, Pr, sr = best[1]
            lay = dict(lay, v=(Pv, sv, wv), r=(Pr, sr, wr), fused=True)
        return lay

    def _padded(self, T, which):
        """Positions the arrays of a chain must hold (which: 0 = value, 1 = reward)."""
        if self._tc is not None:
            lays = [self._tc["r"]] if which else [self._tc["v"], self._tc.get("b")]
            return max([T] + [l[0] * l[1] + l[2] for l in lays if l is not None])
        if self._seg is None:
            return T
        K, seg_v, seg_r, warm = self._seg
        return K * (seg_r if which else seg_v) + warm

    def _tc_scratch(self, key, P):
        """(ws, cp_state) of a tc chain launch; the tensors stay referenced by the engine."""
        ws = self._buf("tc_ws_" + key, (int(_lib.call("icrl_chain_tc_ws_bytes", P)) + 3) // 4)
        cp = self._buf("tc_cp_" + key, int(_lib.call("icrl_chain_tc_cp_floats", P)))
        return ws, cp

    def _chains_forward(self, f, B, S, Tv, Tr, train):
        st, L, b = self._stream, self.launches.ref, self._bufs
        Vn, R = self.value, self.reward
        K = self.chain_shards
        nv, nr = self._padded(Tv, 0), self._padded(Tr, 1)
        v_h = self._buf("v_stash_h", K * (nv + 1) * H)
        v_c = self._buf("v_stash_c", K * (nv + 1) * H)
        v_g = self._buf("v_stash_g", K * (nv + 1) * 4 * H)
        r_h = self._buf("r_stash_h", K * (nr + 1) * H)
        with self._phase("chains_fwd_fused"):
          if self._tc is not None:
            Pv, seg_v, warm_v = self._tc["v"]
            Pr, seg_r, warm_r = self._tc["r"]
            ws_v, cp_v = self._tc_scratch("v", Pv)
            ws_r, cp_r = self._tc_scratch("r", Pr)
            if self._tc.get("fused"):
                _lib.call("icrl_chains_tc_fwd_fused", st, Pv, seg_v, warm_v, _p(b["v_stream"]), _p(b["v_table"]), _p(b["v_chain_pk"]),
                          _p(v_h), _p(v_c), _p(v_g), _p(ws_v), _p(cp_v), _p(self._tc_err[0:]), Pr, seg_r, warm_r, _p(b["r_stream"]),
                          _p(b["r_table"]), _p(b["r_chain_pk"]), _p(R.rewrnn.gru.bias_hh_l0[2 * H:]), _p(r_h), _p(ws_r), _p(cp_r),
                          _p(self._tc_err[4:]), L)
            else:
                _lib.call("icrl_chain_tc_fwd", st, 0, Pv, seg_v, warm_v, _p(b["v_stream"]), _p(b["v_table"]), _p(b["v_chain_pk"]),
                          None, _p(v_h), _p(v_c), _p(v_g), _p(ws_v), _p(cp_v), _p(self._tc_err[0:]), L)
                _lib.call("icrl_chain_tc_fwd", st, 1, Pr, seg_r, warm_r, _p(b["r_stream"]), _p(b["r_table"]), _p(b["r_chain_pk"]),
                          _p(R.rewrnn.gru.bias_hh_l0[2 * H:]), _p(r_h), None, None, _p(ws_r), _p(cp_r), _p(self._tc_err[4:]), L)
          elif self._seg is not None:
            Ks, seg_v, seg_r, warm = self._seg
            _lib.call("icrl_chains_fwd_fused_segmented", st, Ks, warm, _p(b["v_stream"]), seg_v, _p(b["v_table"]),
                  _p(Vn.valrnn.lstm.weight_hh_l0), _p(v_h), _p(v_c), _p(v_g), _p(b["r_stream"]), seg_r, _p(b["r_table"]),
                  _p(R.rewrnn.gru.weight_hh_l0), _p(R.rewrnn.gru.bias_hh_l0[2 * H:]), _p(r_h), _p(self._seg_ws),
                  _p(self.sync_state), L)
          elif K > 1:
            _lib.call("icrl_chains_fwd_fused_sharded", st, K, _p(b["v_stream"]), Tv, _p(b["v_table"]),
                  _p(Vn.valrnn.lstm.weight_hh_l0), _p(v_h), _p(v_c), _p(v_g), _p(b["r_stream"]), Tr, _p(b["r_table"]),
                  _p(R.rewrnn.gru.weight_hh_l0), _p(R.rewrnn.gru.bias_hh_l0[2 * H:]), _p(r_h), _p(self.sync_state), L)
          else:
            _lib.call("icrl_chains_fwd_fused", st, _p(b["v_stream"]), Tv, _p(b["v_table"]), _p(Vn.valrnn.lstm.weight_hh_l0),
                  _p(v_h), _p(v_c), _p(v_g), _p(b["r_stream"]), Tr, _p(b["r_table"]), _p(R.rewrnn.gru.weight_hh_l0),
                  _p(R.rewrnn.gru.bias_hh_l0[2 * H:]), _p(r_h), _p(self.sync_state), L)
          if self._tc is not None and self._tc["v"] is not None:
            Pv, seg_v, warm_v = self._tc["v"]
            end_f = Pv * seg_v + warm_v
            if nv > end_f:                # rows only the backward layout reaches: finite (zero) stash, zero gradients
                v_g[end_f * 4 * H: nv * 4 * H].zero_()
                v_c[(end_f + 1) * H: (nv + 1) * H].zero_()
        SB = S * B
        v_take_h = self._buf("v_take_h", SB * H)
        r_take_h = self._buf("r_take_h", SB * H)
        _lib.call("icrl_gather_rows", st, SB, _p(v_h), _p(b["v_pos"]), 1, _p(v_take_h), L)
        _lib.call("icrl_gather_rows", st, SB, _p(r_h), _p(b["r_pos"]), 1, _p(r_take_h), L)
        values = torch.empty((B, S), dtype=torch.float32, device=self.device)
        rewards = torch.empty((B, S), dtype=torch.float32, device=self.device)
        _lib.call("icrl_value_head_fwd", st, B, S, _p(f), _p(v_take_h), _p(b["v_weff"]), _p(b["v_beff"]), _p(values), L)
        se = self._buf("r_se", SB * H)
        ve = self._buf("r_ve", B * H)
        if SB >= self._TC_LINEAR_MIN_ROWS:
            parts = self._buf("r_take_parts", 3 * SB * H, torch.bfloat16)
            _lib.call("icrl_split_bf16x3", st, SB * H, _p(r_take_h), _p(parts), L)
            _lib.call("icrl_gemm_bf16x3", st, SB, H, H, _p(parts), _p(b["r_se_parts"]), _p(se), H, _p(R.semantic_embed.bias), L)
        else:
            self._gemm(0, 1, SB, H, H, r_take_h, H, R.semantic_embed.weight, H, se, H, R.semantic_embed.bias)
        self._gemm(0, 1, B, H, H, f, H, R.visual_embed.weight, H, ve, H, R.visual_embed.bias)
        _lib.call("icrl_reward_cosine_fwd", st, B, S, _p(ve), _p(se), _p(rewards), L)
        return values, rewards

    def _backward(self, f, tokcm, tokens, B, p0, S, Tv):
        """Backward of the minibatch.  The policy's backward depends only on dL/dlogp; the value network's only on
        dL/dvalues: with overlap_backward the policy side runs on a second stream (own workspaces) while the main stream
        walks the value chain backwards -- the chain kernels occupy 120 of the 148 SMs (clusters of 8 cannot span GPCs),
        the policy's contractions fill the rest and whatever the chain kernel leaves in time."""
        main = torch.cuda.current_stream(self.device)
        if self.overlap_backward:
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=self.device)
            fork = torch.cuda.Event()
            fork.record(main)
            self._backward_value(f, B, p0, S, Tv, launch_only_chain=True)
            self._side_stream.wait_event(fork)
            h_packed = None
            with torch.cuda.stream(self._side_stream):
                if self.wgrad == "tc" and self.chain_shards == 1:
                    # the h stash is final since the forward: its transpose + split (B operand of dW_hh = dgates^T h) leaves
                    # the critical path and runs beside the chain backward
                    _, gemm_ws, gemm_ws_floats, _ = self._bwd_workspaces(B, S, Tv, "")
                    _lib.call("icrl_wgrad_tc_pack_b", self._stream, 4 * H, H, Tv, _p(self._bufs["v_stash_h"]), H, _p(gemm_ws),
                              gemm_ws_floats * 4, 2, self.launches.ref)
                    h_packed = torch.cuda.Event()
                    h_packed.record(self._side_stream)
                self._backward_policy(f, tokcm, tokens, B, p0, S, "_p")
                join = torch.cuda.Event()
                join.record(self._side_stream)
            if h_packed is not None:
                main.wait_event(h_packed)
            self._backward_value(f, B, p0, S, Tv, launch_only_chain=False, h_packed=h_packed is not None)
            main.wait_event(join)
        else:
            self._backward_value(f, B, p0, S, Tv, launch_only_chain=True)
            self._backward_value(f, B, p0, S, Tv, launch_only_chain=False)
            self._backward_policy(f, tokcm, tokens, B, p0, S, "")

    def _bwd_workspaces(self, B, S, Tv, tag):
        """(colsum_ws, gemm_ws, gemm_ws_floats, dtable) of the value side (tag "") or the policy side on its own stream."""
        V = self.V
        SB = S * B
        cs_rows = max(SB, V, B)
        colsum_ws = self._buf("colsum_ws" + tag, int(_lib.call("icrl_colsum_ws_floats", cs_rows, 4 * H)) + 2 * H + 4 * H * 8)
        gemm_ws_floats = 24 * 4 * H * H
        if self.wgrad == "tc":
            if tag:       # policy side: dW_v (M = 1024) and dW_hh (M = 2048) contractions over S*B resp. n_cell*B rows
                n = max(int(_lib.call("icrl_wgrad_tc_ws_bytes", 4 * H, H, (self._n_cell_max) * B, 2)),
                        int(_lib.call("icrl_wgrad_tc_ws_bytes", 1024, H, SB, 2)))
            else:
                T_total = Tv if self.chain_shards == 1 else self.chain_shards * (Tv + 1)
                n = int(_lib.call("icrl_wgrad_tc_ws_bytes", 4 * H, H, T_total, 2))
                if not self.overlap_backward:
                    n = max(n, int(_lib.call("icrl_wgrad_tc_ws_bytes", 1024, H, SB, 2)))
            gemm_ws_floats = max(gemm_ws_floats, (n + 3) // 4)
        return colsum_ws, self._buf("gemm_ws" + tag, gemm_ws_floats), gemm_ws_floats, self._buf("dtable" + tag, V * 4 * H)

    def _backward_value(self, f, B, p0, S, Tv, launch_only_chain, h_packed=False):
        st, L, b, V = self._stream, self.launches.ref, self._bufs, self.V
        Vn = self.value
        SB = S * B
        g = self._g
        colsum_ws, gemm_ws, gemm_ws_floats, dtable = self._bwd_workspaces(B, S, Tv, "")
        K = self.chain_shards
        dgates = self._buf("v_dgates", K * (self._padded(Tv, 0) + 1) * 4 * H)
        if not launch_only_chain:
            lstm = Vn.valrnn.lstm
            with self._phase("value_param_grads"):
              _lib.call("icrl_value_chain_param_grads", st, Tv if K == 1 else K * (Tv + 1), V, Vn.valrnn.caption_embedding.weight.shape[1], _p(b["v_stream"]), _p(dgates), _p(b["v_stash_h"]),
                      _p(Vn.valrnn.caption_embedding.weight), _p(lstm.weight_ih_l0), _p(dtable), _p(colsum_ws), _p(gemm_ws),
                      gemm_ws_floats * 4, _p(g(Vn.valrnn.caption_embedding.weight, True)), _p(g(lstm.weight_ih_l0)),
                      _p(g(lstm.weight_hh_l0)), _p(g(lstm.bias_ih_l0)), _p(g(lstm.bias_hh_l0)),
                      B if K == 1 else 0, p0, S, int(bool(h_packed)), L)
            return
        # value head -> dh at the take positions + head gradients
        dh_take = self._buf("v_dh_take", SB * H)
        _lib.call("icrl_value_head_bwd", st, B, S, _p(f), _p(b["v_take_h"]), _p(b["dv_sb"]), _p(b["sum_dv"]),
                  _p(Vn.linear1.weight), _p(Vn.linear1.bias), _p(Vn.linear2.weight), _p(b["v_weff"]), _p(dh_take),
                  _p(g(Vn.linear1.weight)), _p(g(Vn.linear1.bias)), _p(g(Vn.linear2.weight)), _p(g(Vn.linear2.bias)),
                  _p(colsum_ws), L)
        # value chain BPTT and (second call) its parameter gradients (contractions over all T steps)
        with self._phase("chain_lstm_bwd"):
          if self._tc is not None:
            Pv, seg_v, warm_v = self._tc["b"]
            ws, _ = self._tc_scratch("v", Pv)
            cp = self._buf("tc_cp_b", int(_lib.call("icrl_chain_tc_cp_floats", Pv)))
            _lib.call("icrl_chain_tc_lstm_bwd", st, Pv, seg_v, warm_v, _p(b["v_chain_pk"]), _p(b["v_stash_g"]), _p(b["v_stash_c"]),
                      _p(b["v_take"]), _p(dh_take), SB, _p(dgates), _p(ws), _p(cp), _p(self._tc_err[8:]), L)
          elif self._seg is not None:
            Kf, seg_v, _, warm = self._seg
            # backward pieces: 16 (8 per CTA group in one kernel step) when the forward has 16 or 32, else at most 8
            Ks = self.chain_bwd_segments or (16 if Kf in (16, 32) else min(Kf, 8))
            while Kf % Ks:
                Ks //= 2
            seg_v *= Kf // Ks                       # every (Kf/Ks)-th forward joint: same padded length
            _lib.call("icrl_chain_lstm_bwd_segmented", st, Ks, warm, seg_v, _p(Vn.valrnn.lstm.weight_hh_l0),
                  _p(b["v_stash_g"]), _p(b["v_stash_c"]), _p(b["v_take"]), _p(dh_take), SB, _p(dgates), _p(self._seg_ws),
                  _p(self.sync_state), L)
          elif K > 1:
            _lib.call("icrl_chain_lstm_bwd_sharded", st, K, Tv, _p(Vn.valrnn.lstm.weight_hh_l0), _p(b["v_stash_g"]),
                  _p(b["v_stash_c"]), _p(b["v_take"]), _p(dh_take), _p(dgates), _p(self.sync_state), L)
          else:
            _lib.call("icrl_chain_lstm_bwd", st, Tv, _p(Vn.valrnn.lstm.weight_hh_l0), _p(b["v_stash_g"]), _p(b["v_stash_c"]),
                  _p(b["v_take"]), _p(dh_take), _p(dgates), _p(self.sync_state), None, None, None, None, L)

    def _backward_policy(self, f, tokcm, tokens, B, p0, S, tag):
        # policy BPTT.  icrl_policy_rollout_bwd turns the logits into dL/dlogits IN PLACE; a step whose chains are re-run
        # (longer warm-up, serial fall-back) runs this backward again with slightly different dL/dlogp, so it works on a
        # copy and the rollout's logits stay intact.
        st, L, b, V = self._stream, self.launches.ref, self._bufs, self.V
        P = self.policy
        SB = S * B
        n_cell = p0 - 1 + S
        self._n_cell_max = max(self._n_cell_max, n_cell)
        g = self._g
        colsum_ws, gemm_ws, gemm_ws_floats, dtable = self._bwd_workspaces(B, S, 0, tag)
        pl = P.lstm
        tc = self.policy_bptt == "tc"
        ldz = int(_lib.call("icrl_vocab_pad")) if tc and V <= 1024 else V      # padded rows: vocabulary contractions on tcgen05
        dz = self._bufs.get("p_dlogits")
        if dz is None or dz.numel() < SB * ldz or self._dz_ld != ldz:
            dz = self._bufs["p_dlogits"] = torch.zeros(SB * ldz, dtype=torch.float32, device=self.device)   # padding stays zero
            self._dz_ld = ldz
        with self._phase("policy_bwd"):
          dz[:SB * ldz].view(SB, ldz)[:, :V].copy_(b["p_logits"][:SB * V].view(SB, V))
          args = [st, B, V, p0, S, P.caption_embedding.weight.shape[1], _p(f), _p(P.caption_embedding.weight), _p(pl.weight_ih_l0),
                  _p(pl.weight_hh_l0), _p(P.linear2vocab.weight), _p(tokcm), _p(tokens), _p(b["dlogp"]), _p(b["p_Hs"]),
                  _p(b["p_Cs"]), _p(b["p_Gs"]), _p(dz), _p(self._buf("p_dHv", SB * H)),
                  _p(self._buf("p_DG", n_cell * B * 4 * H)), _p(self._buf("p_dh", 2 * B * H)), _p(self._buf("p_dc", B * H)),
                  _p(dtable), _p(colsum_ws), _p(gemm_ws), gemm_ws_floats * 4, _p(g(P.caption_embedding.weight, True)),
                  _p(g(P.cnn2linear.weight)), _p(g(P.cnn2linear.bias)), _p(g(pl.weight_ih_l0)), _p(g(pl.weight_hh_l0)),
                  _p(g(pl.bias_ih_l0)), _p(g(pl.bias_hh_l0)), _p(g(P.linear2vocab.weight)), _p(g(P.linear2vocab.bias))]
          if self.policy_bptt == "tc":
            # the n_cell serial cell-backward steps as ONE launch of the tcgen05 chain-backward kernel (rows = MMA rows)
            ws = self._buf("p_bptt_ws", (int(_lib.call("icrl_policy_bwd_tc_ws_bytes", B, S, n_cell)) + 3) // 4)
            _lib.call("icrl_policy_rollout_bwd_tc", *args, _p(b["p_chain_pk"]), _p(ws), _p(self._stat_err[20:]), ldz, L)
          else:
            _lib.call("icrl_policy_rollout_bwd", *args, L)

    # ------------------------------------------------------------------ public API
    def step(self, features, captions=None, uniforms=None, level=None, greedy=False, backward=True,
             global_rows=None, forced_tokens=None, repack=True, check=True, plan=None):
        """One A2C minibatch.  `features` may be a Prepared batch (already in HBM) or host/device data:
        features (B,512) f32, captions (B,L) int (host), uniforms (S,B) f64 or None
        (None draws them from numpy's global MT19937 stream exactly as np.random.choice would,
        trainers.py:447-450).  Returns StepResult or None when the curriculum level does not fit
        (trainers.py:550).  Gradients are written to the parameters' .grad (flat bucket)."""
        with torch.cuda.device(self.device):
            if isinstance(features, Prepared):
                prep = features
            else:
                p0, S = plan if plan is not None else plan_rollout(captions, level)
                if p0 < 1:
                    return None
                if S < 1:
                    raise ValueError("rollout needs at least one sampled step")
                B = int(captions.shape[0])
                if uniforms is None and not greedy and forced_tokens is None:
                    uniforms = np.random.random_sample(S * B).reshape(S, B)
                prep = self.prepare(features, captions, uniforms, plan=(p0, S))
            p0, S = prep.p0, prep.S
            self._attach_grads()
            if repack:
                self.pack_weights(reward="r_table" not in self._bufs)
            f, tokcm, u, fo, B = self._stage_inputs(prep, forced_tokens)
            tokens, logp = self._policy_forward(f, tokcm, u, fo, B, p0, S, greedy)
            dv_sb = self._buf("dv_sb", S * B)
            dlogp = self._buf("dlogp", S * B)
            sum_dv = self._buf("sum_dv", 1)
            inv = 1.0 / float((global_rows or B) * S)

            def run_chains(serial):
                Tv, Tr = self._streams(tokcm, B, p0, S, serial=serial)
                values, rewards = self._chains_forward(f, B, S, Tv, Tr, backward)
                _lib.call("icrl_a2c_loss_fwd_bwd", self._stream, B, S, _p(values), _p(rewards), _p(logp), inv, _p(self._stat_err),
                          _p(dv_sb), _p(dlogp), _p(sum_dv), self.launches.ref)
                if backward:
                    self._backward(f, tokcm, tokens, B, p0, S, Tv)
                self.segment_stats["steps"] += 1
                return Tv, Tr, values, rewards

            # Attempts 0 .. _TC_ATTEMPTS-1 run the chains cut into pieces; a failed joint check lengthens the warm-up and
            # tries again (tc engine) or goes straight to the serial kernels (simt engine); the last attempt is always
            # the serial kernels, the guaranteed floor.  Gradients are overwritten, not accumulated, by a re-run.
            host, serial = None, False
            for attempt in range(self._TC_ATTEMPTS + 1):
                serial = serial or attempt == self._TC_ATTEMPTS
                Tv, Tr, values, rewards = run_chains(serial)
                if self._seg is None and self._tc is None:
                    break
                self.segment_stats["segmented_steps"] += 1
                if not check:
                    self._seg_unverified = True       # the caller owes a segments_verified() before trusting the steps
                    break
                if self._tc is not None:
                    ok, host = self._tc_ok()
                    if not ok:
                        host = None
                        serial = self._tc_force_serial or attempt + 1 == self._TC_ATTEMPTS
                        if serial:
                            self.segment_stats["fallbacks"] += 1
                else:
                    ok = self._segments_ok()
                    serial = True
                if ok:
                    break
            stats = self._stat_err[:3].clone()
            if check and self._tc is None:
                _lib.call("icrl_chain_check", self._stream, _p(self.sync_state))
        res = StepResult(tokens=tokens, logp=logp, values=values, rewards=rewards, stats=stats, p0=p0, S=S, B=B,
                         Tv=Tv, Tr=Tr)
        if host is not None:
            res.host_stats = host[:3]
        return res

    _TC_ATTEMPTS = 3
    _SHRINK_AFTER = 2
    _HOLD_AFTER_GROWTH = 64
    _tc_force_serial = False

    def _adapt_warm(self, key, e_half, e_full):
        """Size the next warm-up of chain `key` from the joint errors half-way through / at the end of this one.
        The two errors give the contraction rate of the recurrence under the current weights: the warm-up GROWS as soon
        as the end-of-warm-up error passes chain_tol / 4 (before the check can fail; by the measured rate, at least
        x1.25, at most x4), and SHRINKS to 5/8 once the half-way error alone has been below that target on
        _SHRINK_AFTER consecutive steps (the new end point then lies beyond a point already measured clean); after a
        growth it is held for _HOLD_AFTER_GROWTH steps.  (A finer shrink by the positions the measured rate predicts can be
        given back was tried: 2 % faster on one GPU, but with 512 rows per rank the value chain then needed growth inside
        the next ten optimizer steps and paid for it with re-runs -- the half-way rule is the one with a measured point
        behind it.)  Returns False when the step failed its check (end-of-warm-up
        error above chain_tol, or not a number)."""
        tol = self.chain_tol
        target = tol / 4.0
        warm = self.warm[key]
        failed = not (e_full <= tol)
        self._hold[key] = max(0, self._hold[key] - 1)
        if failed or (self.chain_adapt and not (e_full <= target)):
            need = 2.0 * warm
            if e_half > e_full > 0.0 and np.isfinite(e_half):
                rate = np.log(e_half / e_full) / float(warm - warm // 2)         # nats forgotten per position
                need = warm + np.log(e_full / (target / 4.0)) / rate
            new = min(max(need, (1.5 if failed else 1.25) * warm), 4.0 * warm)
            self.warm[key] = int(-(-int(np.ceil(new)) // 32) * 32)
            self._clean[key] = 0
            self._hold[key] = self._HOLD_AFTER_GROWTH
        elif self.chain_adapt and warm >= 8 and e_half <= target and self._hold[key] == 0:
            self._clean[key] += 1
            if self._clean[key] >= self._SHRINK_AFTER and warm > self.chain_warmup_min:
                self.warm[key] = max(self.chain_warmup_min, int(-(-int(0.625 * warm) // 32) * 32))
                self._clean[key] = 0
        else:
            self._clean[key] = 0
        if self.warm[key] != warm:
            self.segment_stats["warm_history"].append((self.segment_stats["steps"], key, warm, self.warm[key]))
        return not failed

    def _tc_ok(self):
        """Joint checks of the tc chain launches of this step, read together with the three step statistics in ONE
        device-to-host copy (synchronises), and re-armed.  Returns (ok, host words)."""
        host = self._stat_err.tolist()
        self._tc_err.zero_()
        self._seg_unverified = False
        e = host[4:]
        st = self.segment_stats
        st["tc_max_err"] = [max(a, b) if b == b else float("nan") for a, b in zip(st["tc_max_err"], e)]
        nanmax = lambda xs: float("nan") if any(x != x for x in xs) else max(xs)
        self._tc_force_serial = e[13] != 0.0          # the fp16 exchange of the backward recurrence overflowed
        ok_v = ok_b = True
        if self._tc["v"] is not None:
            ok_v = self._adapt_warm("v", nanmax(e[2:4]), nanmax(e[0:2]))
            if e[12] != 0.0 or any(x != 0.0 for x in e[8:12]):        # the backward recurrence ran in this step
                ok_b = self._adapt_warm("b", nanmax(e[10:12]), nanmax(e[8:10]))
        ok_r = self._adapt_warm("r", e[6], e[4])
        ok = ok_v and ok_b and ok_r and not self._tc_force_serial
        if not ok:
            import warnings
            st["reruns"] += 1
            warnings.warn("chain segments did not converge onto the single chain within %g after the warm-up (value chain "
                          "forward |dh| %.3g |dc| %.3g, backward %.3g %.3g; reward chain |dh| %.3g%s): re-running with "
                          "warm-ups %d / %d / %d" % (self.chain_tol, e[0], e[1], e[8], e[9], e[4],
                                                     "; gate gradients overflowed the fp16 exchange" if self._tc_force_serial else "",
                                                     self.warm["v"], self.warm["r"], self.warm["b"]))
        return ok, host

    def _segments_ok(self):
        """Read (and re-arm) the warm-up checks of the segmented chain launches since the last call (synchronises).
        False = some segment had not converged onto the single chain within chain_tol: the caller re-runs serially;
        the warm-up is lengthened for the next steps, and after three failures in a row segments are switched off."""
        e = self._seg_ws[:5].tolist()
        self._seg_ws[:5].zero_()
        self._seg_unverified = False
        st = self.segment_stats
        st["max_err"] = [max(a, b) if b == b else float("nan") for a, b in zip(st["max_err"], e)]
        tol = self.chain_tol
        ok = e[0] <= tol and e[1] <= tol and e[2] <= tol and e[3] <= tol * max(e[4], 1e-30)
        if ok:
            self._seg_strikes = 0
            return True
        import warnings
        st["fallbacks"] += 1
        self._seg_strikes += 1
        self.chain_warmup *= 4
        if self._seg_strikes >= 3:
            self.chain_segments = 1
        warnings.warn("chain segments did not converge onto the single chain within %g after the warm-up (|dh| %.3g, |dc| "
                      "%.3g, reward |dh| %.3g, joint gate gradients %.3g of %.3g): re-running on the serial kernels; %s"
                      % (tol, e[0], e[1], e[2], e[3], e[4],
                         "segments are now off" if self.chain_segments == 1 else "warm-up is now %d" % self.chain_warmup))
        return False

    @property
    def segment_layout(self):
        """(pieces, value-chain piece length, reward-chain piece length, warm-up) of the last step, or None when it ran on
        the serial kernels (chain too short, chain_segments = 1, or a failed warm-up check)."""
        if self._tc is not None:
            v, r = self._tc["v"], self._tc["r"]
            return (r[0] if v is None else v[0], 0 if v is None else v[1], r[1], max(r[2], 0 if v is None else v[2]))
        return self._seg

    @property
    def piece_layout(self):
        """{"v": (pieces, positions per piece, warm-up) | None, "r": (...)} of the last step on the tcgen05 chain kernels,
        or None."""
        return self._tc

    def segments_verified(self):
        """For callers that ran step(check=False): True when every segmented launch since the last check passed."""
        if not self._seg_unverified:
            return True
        with torch.cuda.device(self.device):
            if self._tc is not None:
                return self._tc_ok()[0]
            return self._segments_ok()

    def get_rewards(self, features, captions):
        """GetRewards on whole captions from zero state (trainers.py:108-121): (B,1) tensor."""
        caps = np.asarray(captions)
        B, Lc = caps.shape
        with torch.cuda.device(self.device):
            if "r_table" not in self._bufs or self._reward_changed():
                self.pack_reward()
            f, tokcm, _, _, _ = self._stage_inputs(self.prepare(features, caps, plan=(Lc, 0)), None)
            st, L, b, R = self._stream, self.launches.ref, self._bufs, self.reward
            T = int(_lib.call("icrl_stream_len", B, Lc, 1, 0))
            serial = False
            for attempt in range(self._TC_ATTEMPTS + 1):
                serial = serial or attempt == self._TC_ATTEMPTS
                self._seg, self._tc = None, None
                if not serial:
                    if self._use_tc_chains():
                        self._tc = self._pick_pieces(0, T)
                    elif self.chain_segments > 1:
                        self._seg = self._pick_segments(0, T)
                n = self._padded(T, 1)
                r_stream, r_pos = self._buf("r_stream", n, torch.int32), self._buf("r_pos", B, torch.int32)
                _lib.call("icrl_build_stream", st, B, Lc, 1, 0, _p(tokcm), _p(r_stream), None, _p(r_pos), L)
                r_h = self._buf("r_stash_h", (n + 1) * H)
                if self._tc is not None:
                    Pr, seg_r, warm_r = self._tc["r"]
                    r_stream[T:n].zero_()
                    ws, cp = self._tc_scratch("r", Pr)
                    _lib.call("icrl_chain_tc_fwd", st, 1, Pr, seg_r, warm_r, _p(r_stream), _p(b["r_table"]), _p(b["r_chain_pk"]),
                              _p(R.rewrnn.gru.bias_hh_l0[2 * H:]), _p(r_h), None, None, _p(ws), _p(cp), _p(self._tc_err[4:]), L)
                    ok, _ = self._tc_ok()
                    if ok:
                        break
                    serial = attempt + 1 == self._TC_ATTEMPTS
                    if serial:
                        self.segment_stats["fallbacks"] += 1
                elif self._seg is not None:
                    Ks, _, seg_r, warm = self._seg
                    r_stream[T:n].zero_()
                    _lib.call("icrl_chains_fwd_fused_segmented", st, Ks, warm, None, 0, None, None, None, None, None,
                              _p(r_stream), seg_r, _p(b["r_table"]), _p(R.rewrnn.gru.weight_hh_l0),
                              _p(R.rewrnn.gru.bias_hh_l0[2 * H:]), _p(r_h), _p(self._seg_ws), _p(self.sync_state), L)
                    if self._segments_ok():
                        break
                    serial = True
                else:
                    _lib.call("icrl_chain_gru_fwd", st, _p(r_stream), T, _p(b["r_table"]), _p(R.rewrnn.gru.weight_hh_l0),
                              _p(R.rewrnn.gru.bias_hh_l0[2 * H:]), None, _p(r_h), None, _p(self.sync_state), None, L)
                    break
            r_take_h = self._buf("r_take_h", B * H)
            _lib.call("icrl_gather_rows", st, B, _p(r_h), _p(r_pos), 1, _p(r_take_h), L)
            se, ve = self._buf("r_se", B * H), self._buf("r_ve", B * H)
            self._gemm(0, 1, B, H, H, r_take_h, H, R.semantic_embed.weight, H, se, H, R.semantic_embed.bias)
            self._gemm(0, 1, B, H, H, f, H, R.visual_embed.weight, H, ve, H, R.visual_embed.bias)
            rewards = torch.empty((B, 1), dtype=torch.float32, device=self.device)
            _lib.call("icrl_reward_cosine_fwd", st, B, 1, _p(ve), _p(se), _p(rewards), L)
            _lib.call("icrl_chain_check", st, _p(self.sync_state))
        return rewards

    def greedy_decode(self, features, first_col, steps=16):
        """GenerateCaptionsGreedy (trainers.py:57-70): (B, steps+1) int64 tokens and last-step logits."""
        first = np.asarray(first_col, dtype=np.int64).reshape(-1, 1)
        B = first.shape[0]
        with torch.cuda.device(self.device):
            self.pack_weights(reward=False)
            f, tokcm, _, _, _ = self._stage_inputs(self.prepare(features, first, plan=(1, steps)), None)
            tokens, _ = self._policy_forward(f, tokcm, None, None, B, 1, steps, True)
            last = self._bufs["p_logits"][(steps - 1) * B * self.V: steps * B * self.V].view(B, self.V).clone()
            out = torch.cat((torch.from_numpy(first).to(self.device), tokens), dim=1)
        return out, last
