"""Single-pass restatement of the A2C minibatch with the cell arithmetic written out.

TEST INFRASTRUCTURE (see oracle/__init__).  This is the formulation the CUDA kernels
implement, checked against ``ref_port`` (the as-executed form) and against the golden
vectors produced by the unmodified reference (tests/test_oracle_golden.py):

policy   h0 = cnn2linear(f), c0 = 0 (models.py:75-78); teacher-force prefix tokens
         0..p0-2; then for s = 0..S-1: cell(E[token p0-1+s]) -> softmax(linear2vocab(h))
         -> sample -> log p[a].  Equivalent to the reference's per-step prefix re-run
         (models.py:286) because an LSTM over a prefix is causal.
value    ONE batch-1 LSTM from zero state over the token stream
         concat_s [ tokens[:, :p0+s] in column-major order ], h taken at the last B
         positions of block s (models.py:130-135, 166-169 + the carried hidden_cell).
reward   ONE batch-1 GRU over concat_s [ tokens[:, :p0+s+1] column-major ], same take
         (models.py:223-228, 253-255), then cosine(visual_embed(f), semantic_embed(h))
         with F.normalize's eps = 1e-12 clamp (trainers.py:117-120).
loss     advantage = values - rewards (sign as written, not detached);
         mean(-logp*adv) + 0.5*mean(adv^2)  (trainers.py:471-475).

Cell equations (torch.nn.LSTM / nn.GRU documentation, gate order i,f,g,o / r,z,n):
  LSTM  i,f,g,o = split(W_ih x + b_ih + W_hh h + b_hh); c' = s(f) c + s(i) tanh(g);
        h' = s(o) tanh(c')
  GRU   r = s(W_ir x + b_ir + W_hr h + b_hr); z likewise;
        n = tanh(W_in x + b_in + r * (W_hn h + b_hn)); h' = (1 - z) n + z h
"""
import numpy as np
import torch

HID = 512


def lstm_cell(xg, h, c, w_hh):
    """xg = W_ih x + b_ih + b_hh (.., 4H); h, c (.., H); w_hh (4H, H)."""
    g = xg + h @ w_hh.t()
    i, f, gg, o = g.split(HID, dim=-1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    return torch.sigmoid(o) * torch.tanh(c2), c2


def gru_cell(xg, h, w_hh, b_hh):
    """xg = W_ih x + b_ih (.., 3H); h (.., H)."""
    hg = h @ w_hh.t() + b_hh
    xr, xz, xn = xg.split(HID, dim=-1)
    hr, hz, hn = hg.split(HID, dim=-1)
    r = torch.sigmoid(xr + hr)
    z = torch.sigmoid(xz + hz)
    n = torch.tanh(xn + r * hn)
    return (1.0 - z) * n + z * h


def stream_tokens(tokens, p0, S, extra):
    """Column-major token stream and the take positions.

    tokens (B, >= p0+S-1+extra) int64.  Block s holds columns 0..p0+s-1+extra, each
    column contributing its B rows in row order.  Returns (stream (T,), take (S,B))
    where take[s, b] indexes the stream position whose hidden state is row b's output
    at rollout step s (the last column of block s)."""
    B = tokens.shape[0]
    parts, take, off = [], [], 0
    for s in range(S):
        n = p0 + s + extra
        parts.append(np.ascontiguousarray(tokens[:, :n].T).reshape(-1))
        off += n * B
        take.append(np.arange(off - B, off))
    return np.concatenate(parts), np.stack(take)


def sample_inverse_cdf(p_row_f32, u):
    cdf = np.cumsum(p_row_f32.astype(np.float64))
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


def policy_rollout(P, features, captions, p0, S, uniforms=None, greedy=False, forced=None):
    """Returns (tokens (B,S) int64 ndarray, logp (B,S) tensor, logits (B,S,V) tensor).

    P: dict of policy tensors (requires_grad as the caller set them).  ``forced`` (B,S)
    teacher-forces the sampled tokens (used to compare downstream quantities without
    letting a single near-tie flip cascade)."""
    f = torch.as_tensor(features).float()
    caps = torch.as_tensor(captions).long()
    B = caps.shape[0]
    E, w_ih, w_hh = P["caption_embedding.weight"], P["lstm.weight_ih_l0"], P["lstm.weight_hh_l0"]
    bias = P["lstm.bias_ih_l0"] + P["lstm.bias_hh_l0"]
    h = f @ P["cnn2linear.weight"].t() + P["cnn2linear.bias"]
    c = torch.zeros_like(h)
    for t in range(p0 - 1):                                   # teacher-forced prefix
        h, c = lstm_cell(E[caps[:, t]] @ w_ih.t() + bias, h, c, w_hh)
    cur = caps[:, p0 - 1]
    toks, lps, lgs = [], [], []
    for s in range(S):
        h, c = lstm_cell(E[cur] @ w_ih.t() + bias, h, c, w_hh)
        logits = h @ P["linear2vocab.weight"].t() + P["linear2vocab.bias"]
        probs = torch.softmax(logits, dim=1)
        pn = probs.detach().numpy()
        if forced is not None:
            a = np.asarray(forced[:, s], dtype=np.int64)
        elif greedy:
            a = pn.argmax(axis=1).astype(np.int64)
        else:
            a = np.array([sample_inverse_cdf(pn[b], uniforms[s, b]) for b in range(B)], dtype=np.int64)
        cur = torch.from_numpy(a)
        lps.append(torch.log(probs.gather(1, cur.view(B, 1)))[:, 0])
        toks.append(a)
        lgs.append(logits)
    return np.stack(toks, axis=1), torch.stack(lps, dim=1), torch.stack(lgs, dim=1)


def value_chain(Vw, features, all_tokens, p0, S, lib=False):
    """values (B,S).  all_tokens (B, p0+S) int64 ndarray (prefix + sampled)."""
    f = torch.as_tensor(features).float()
    B = all_tokens.shape[0]
    stream, take = stream_tokens(all_tokens, p0, S, extra=0)
    E, w_ih, w_hh = Vw["valrnn.caption_embedding.weight"], Vw["valrnn.lstm.weight_ih_l0"], Vw["valrnn.lstm.weight_hh_l0"]
    b_ih, b_hh = Vw["valrnn.lstm.bias_ih_l0"], Vw["valrnn.lstm.bias_hh_l0"]
    st = torch.from_numpy(stream)
    if lib:
        z = torch.zeros(1, 1, HID)
        hs = torch.lstm(E[st].view(-1, 1, HID), (z, z), (w_ih, w_hh, b_ih, b_hh), True, 1, 0.0, False, False, False)[0][:, 0]
    else:
        xg = E[st] @ w_ih.t() + (b_ih + b_hh)
        h = torch.zeros(HID)
        c = torch.zeros(HID)
        outs = []
        for t in range(len(stream)):
            h, c = lstm_cell(xg[t], h, c, w_hh)
            outs.append(h)
        hs = torch.stack(outs)
    ht = hs[torch.from_numpy(take.T.copy())]                  # (B,S,H)
    state = torch.cat((f.unsqueeze(1).expand(B, S, HID), ht), dim=2)
    mid = state @ Vw["linear1.weight"].t() + Vw["linear1.bias"]          # models.py:177
    return (mid @ Vw["linear2.weight"].t() + Vw["linear2.bias"])[:, :, 0], ht  # models.py:178


def reward_chain(R, features, all_tokens, p0, S, lib=False, extra=1):
    """rewards (B,S) in [-1,1] (no grad: the reward net is frozen, trainers.py:372)."""
    f = torch.as_tensor(features).float()
    stream, take = stream_tokens(all_tokens, p0, S, extra=extra)
    E, w_ih, w_hh = R["rewrnn.caption_embedding.weight"], R["rewrnn.gru.weight_ih_l0"], R["rewrnn.gru.weight_hh_l0"]
    b_ih, b_hh = R["rewrnn.gru.bias_ih_l0"], R["rewrnn.gru.bias_hh_l0"]
    st = torch.from_numpy(stream)
    with torch.no_grad():
        if lib:
            hs = torch.gru(E[st].view(-1, 1, HID), torch.zeros(1, 1, HID), (w_ih, w_hh, b_ih, b_hh),
                           True, 1, 0.0, False, False, False)[0][:, 0]
        else:
            xg = E[st] @ w_ih.t() + b_ih
            h = torch.zeros(HID)
            outs = []
            for t in range(len(stream)):
                h = gru_cell(xg[t], h, w_hh, b_hh)
                outs.append(h)
            hs = torch.stack(outs)
        ht = hs[torch.from_numpy(take.T.copy())]              # (B,S,H)
        se = ht @ R["semantic_embed.weight"].t() + R["semantic_embed.bias"]
        ve = f @ R["visual_embed.weight"].t() + R["visual_embed.bias"]
        se = se / se.norm(dim=2, keepdim=True).clamp_min(1e-12)           # F.normalize
        ve = ve / ve.norm(dim=1, keepdim=True).clamp_min(1e-12)
        return (ve.unsqueeze(1) * se).sum(dim=2), ht


def get_rewards(R, features, captions, lib=True):
    """GetRewards on whole captions from zero state (config 3): one block of L columns."""
    caps = np.asarray(captions)
    r, _ = reward_chain(R, features, caps, caps.shape[1], 1, lib=lib, extra=0)
    return r.numpy()                                           # (B,1)


def plan(captions, level=None):
    caps = np.asarray(captions)
    caplen = int(np.nonzero(caps == 2)[1].max()) + 1          # trainers.py:436 / :547
    if level is None:
        return 1, caplen - 1
    return caplen - level, level


def a2c_minibatch(weights, features, captions, uniforms, level=None, backward=True,
                  greedy=False, forced=None, lib=False, loss_scale_rows=None):
    """Same contract as ref_port.a2c_minibatch, computed in one pass.

    ``loss_scale_rows``: global row count for a data-parallel shard (gradient seeds are
    scaled by 1/(B_global*S)); default = local B."""
    p0, S = plan(captions, level)
    if p0 < 1:
        return None
    caps = np.asarray(captions)
    B = caps.shape[0]
    Pw = {k: v.clone().requires_grad_(backward) for k, v in weights["policy"].items()}
    Vw = {k: v.clone().requires_grad_(backward) for k, v in weights["value"].items()}
    toks, logp, logits = policy_rollout(Pw, features, caps, p0, S, uniforms, greedy, forced)
    all_tokens = np.concatenate((caps[:, :p0], toks), axis=1)
    values, h_val = value_chain(Vw, features, all_tokens, p0, S, lib=lib)
    rewards, h_rew = reward_chain(weights["reward"], features, all_tokens, p0, S, lib=lib)
    adv = values - rewards
    denom = float((loss_scale_rows or B) * S)
    loss = (-logp * adv).sum() / denom + 0.5 * adv.pow(2).sum() / denom
    grads = None
    if backward:
        loss.backward()
        grads = {}
        for k, v in Vw.items():
            grads["value_network." + k] = v.grad if v.grad is not None else torch.zeros_like(v)
        for k, v in Pw.items():
            grads["policy_network." + k] = v.grad if v.grad is not None else torch.zeros_like(v)
    return dict(tokens=toks, values=values.detach().numpy(), rewards=rewards.numpy(),
                logp=logp.detach().numpy(), logits=logits.detach().numpy(), loss=float(loss.detach()),
                mean_reward=float(rewards.mean()), mean_adv=float(adv.detach().mean()),
                grads=grads, p0=p0, S=S, h_val=h_val.detach().numpy(), h_rew=h_rew.numpy())


def greedy_decode(weights, features, first_col, steps=16):
    """GenerateCaptionsGreedy (trainers.py:57-70): tokens (B,steps+1), last logits (B,V)."""
    caps = np.asarray(first_col, dtype=np.int64).reshape(-1, 1)
    with torch.no_grad():
        toks, _, logits = policy_rollout(weights["policy"], features, caps, 1, steps, greedy=True)
    return np.concatenate((caps, toks), axis=1), logits[:, -1].numpy()
