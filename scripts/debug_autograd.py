import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.nn import functional as F
from tests.helpers import load_case, make_nets, named_grads
from oracle.gen_golden import grad_sample_index
import icrl_b200.trainers as T
for name in ["a2c_b8_l6", "curr_b16_l10_lv4"]:
    g, seed, f, c, u, level = load_case(name)
    A, R, w = make_nets(seed)
    features = torch.tensor(f, device="cuda").float(); captions = torch.tensor(c, device="cuda").long()
    caplen = int(np.nonzero(c == 2)[1].max() + 1)
    captions_in, steps = (captions[:, :1], caplen - 1) if level is None else (captions[:, :caplen - level], level)
    np.random.seed(seed)
    A.value_network.valrnn.init_hidden(); R.rewrnn.init_hidden()
    values, rewards, log_probs = [], [], []
    for step in range(steps):
        value, probs = A(features, captions_in)
        probs = F.softmax(probs, dim=2)
        dist = probs.cpu().detach().numpy()[:, 0]
        actions = [np.random.choice(probs.shape[-1], p=dist[i]) for i in range(captions.shape[0])]
        gen_cap = torch.from_numpy(np.array(actions)).unsqueeze(-1).to(captions_in.device)
        captions_in = torch.cat((captions_in, gen_cap), axis=1)
        log_probs.append(torch.log(probs[:, 0, :].gather(1, gen_cap)))
        rewards.append(T.GetRewards(features, captions_in, R)); values.append(value)
    values = torch.stack(values, axis=1).squeeze(); rewards = torch.stack(rewards, axis=1).squeeze(); log_probs = torch.stack(log_probs, axis=1).squeeze()
    adv = values - rewards
    loss = (-log_probs * adv).mean() + 0.5 * adv.pow(2).mean()
    loss.backward(retain_graph=True)
    print(name, "loss", float(loss), float(g["loss"]))
    for k, grad in named_grads(A).items():
        flat = grad.reshape(-1); ref = g["gsamp/" + k]; got = flat[grad_sample_index(flat.size)]
        print("  %-50s err/max %.3e   norm %.6e ref %.6e" % (k, np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-12), np.sqrt((flat.astype(np.float64) ** 2).sum()), float(g["gnorm/" + k])))
