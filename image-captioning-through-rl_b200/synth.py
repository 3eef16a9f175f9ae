"""Deterministic synthetic weights and inputs for benchmarks and tests (no arithmetic of the path lives here).

The reference's ``models_pretrained/*.pt`` blobs are absent from the mounted reference
(``.MISSING_LARGE_BLOBS``), so every parity case runs on weights drawn here.  The
draws use numpy's legacy MT19937 ``RandomState`` only (bit-portable across machines),
with PyTorch's default init *distributions* for the layers the reference builds
(``models.py:62-69, 114-120, 160-164, 209-215, 250-251``): embeddings ~ N(0,1),
LSTM/GRU/Linear ~ U(-1/sqrt(fan), 1/sqrt(fan)).  Key names and shapes are the
reference ``state_dict`` layout.
"""
import numpy as np
import torch

H = 512          # hidden / embedding / feature width (models.py:41, 160, 250)
VOCAB = 1004     # len(word_to_idx) of the committed runs
START, END = 1, 2


def word_to_idx(vocab=VOCAB):
    return {"w%d" % i: i for i in range(vocab)}


def _uni(rs, shape, fan):
    k = 1.0 / np.sqrt(fan)
    return rs.uniform(-k, k, size=shape).astype(np.float32)


def make_weights(seed=0, vocab=VOCAB, emb_scale=1.0, wordvec_dim=H, bidirectional=False):
    """Return {'policy','value','reward'} -> state_dict of float32 torch tensors.  wordvec_dim != 512 gives the
    layout of the frozen-pretrained-embedding variant (models.py:61-63): E (V, D), W_ih (gates, D)."""
    D = wordvec_dim
    rs = np.random.RandomState(1000 + seed)
    nrm = lambda *s: (emb_scale * rs.standard_normal(s)).astype(np.float32)
    policy = {
        "caption_embedding.weight": nrm(vocab, D),
        "cnn2linear.weight": _uni(rs, (H, H), H),
        "cnn2linear.bias": _uni(rs, (H,), H),
        "lstm.weight_ih_l0": _uni(rs, (4 * H, D), H),
        "lstm.weight_hh_l0": _uni(rs, (4 * H, H), H),
        "lstm.bias_ih_l0": _uni(rs, (4 * H,), H),
        "lstm.bias_hh_l0": _uni(rs, (4 * H,), H),
        "linear2vocab.weight": _uni(rs, (vocab, H), H),
        "linear2vocab.bias": _uni(rs, (vocab,), H),
    }
    value = {
        "valrnn.caption_embedding.weight": nrm(vocab, D),
        "valrnn.lstm.weight_ih_l0": _uni(rs, (4 * H, D), H),
        "valrnn.lstm.weight_hh_l0": _uni(rs, (4 * H, H), H),
        "valrnn.lstm.bias_ih_l0": _uni(rs, (4 * H,), H),
        "valrnn.lstm.bias_hh_l0": _uni(rs, (4 * H,), H),
        "linear1.weight": _uni(rs, (H, 2 * H), 2 * H),
        "linear1.bias": _uni(rs, (H,), 2 * H),
        "linear2.weight": _uni(rs, (1, H), H),
        "linear2.bias": _uni(rs, (1,), H),
    }
    reward = {
        "rewrnn.caption_embedding.weight": nrm(vocab, D),
        "rewrnn.gru.weight_ih_l0": _uni(rs, (3 * H, D), H),
        "rewrnn.gru.weight_hh_l0": _uni(rs, (3 * H, H), H),
        "rewrnn.gru.bias_ih_l0": _uni(rs, (3 * H,), H),
        "rewrnn.gru.bias_hh_l0": _uni(rs, (3 * H,), H),
        "visual_embed.weight": _uni(rs, (H, H), H),
        "visual_embed.bias": _uni(rs, (H,), H),
        "semantic_embed.weight": _uni(rs, (H, H), H),
        "semantic_embed.bias": _uni(rs, (H,), H),
    }
    if bidirectional:
        # bidirectional variant (models.py:59-69, 120, 163-164, 215, 247-251): reverse-direction RNN weights, 1024-wide
        # cnn2linear / linear2vocab / semantic_embed, value rnn_linear.  Drawn from a second stream so that the
        # unidirectional draws above stay what the committed fixtures were generated with.
        r2 = np.random.RandomState(3000 + seed)
        for sd, pre, G in ((policy, "lstm.", 4 * H), (value, "valrnn.lstm.", 4 * H), (reward, "rewrnn.gru.", 3 * H)):
            sd[pre + "weight_ih_l0_reverse"] = _uni(r2, (G, D), H)
            sd[pre + "weight_hh_l0_reverse"] = _uni(r2, (G, H), H)
            sd[pre + "bias_ih_l0_reverse"] = _uni(r2, (G,), H)
            sd[pre + "bias_hh_l0_reverse"] = _uni(r2, (G,), H)
        policy["cnn2linear.weight"] = _uni(r2, (2 * H, H), H)
        policy["cnn2linear.bias"] = _uni(r2, (2 * H,), H)
        policy["linear2vocab.weight"] = _uni(r2, (vocab, 2 * H), 2 * H)
        policy["linear2vocab.bias"] = _uni(r2, (vocab,), 2 * H)
        value["rnn_linear.weight"] = _uni(r2, (H, 2 * H), 2 * H)
        value["rnn_linear.bias"] = _uni(r2, (H,), 2 * H)
        reward["semantic_embed.weight"] = _uni(r2, (H, 2 * H), 2 * H)
        reward["semantic_embed.bias"] = _uni(r2, (H,), 2 * H)
    out = {}
    for name, sd in (("policy", policy), ("value", value), ("reward", reward)):
        out[name] = {k: torch.from_numpy(v) for k, v in sd.items()}
    return out


def a2c_state_dict(weights):
    """The a2cNetwork.pt layout: value keys then policy keys (models.py:279-280)."""
    sd = {}
    for k, v in weights["value"].items():
        sd["value_network." + k] = v
    for k, v in weights["policy"].items():
        sd["policy_network." + k] = v
    return sd


def make_inputs(seed, B, L, vocab=VOCAB):
    """features (B,512) f32 ~ N(0,1); captions (B,L) int64, col 0 = <START>, col L-1 = <END>."""
    rs = np.random.RandomState(2000 + seed)
    features = rs.standard_normal((B, H)).astype(np.float32)
    captions = rs.randint(4, vocab, size=(B, L)).astype(np.int64)
    captions[:, 0] = START
    captions[:, L - 1] = END
    return features, captions


def make_uniforms(seed, S, B):
    """Exactly the doubles np.random.choice would consume after np.random.seed(seed):
    one per call, step-major / row-minor (trainers.py:447-450)."""
    return np.random.RandomState(seed).random_sample(S * B).reshape(S, B)
