"""Flat Adam for the fused engine: ``torch.optim.Adam(a2c.parameters(), lr=1e-4)`` of the reference
(``trainers.py:378, 480``) as ONE kernel over the flat gradient bucket the engine already fills.

The parameters are moved into one flat fp32 buffer (each ``p.data`` becomes a view of it, so modules, ``state_dict`` and
checkpoints keep working); exp_avg / exp_avg_sq are flat too.  Numerics follow torch's own Adam kernel operation by
operation (``tests/test_gpu_parity.py::test_flat_adam_matches_torch``).  Optional: every training entry point also
accepts a plain ``torch.optim`` optimizer, as the reference passes one in."""
import ctypes

import torch

from . import _lib


class FlatAdam:
    def __init__(self, engine, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        self.engine, self.lr, self.betas, self.eps, self.t = engine, float(lr), tuple(betas), float(eps), 0
        ps = [p for p, _ in engine._grad_views]
        n = engine.flat_grad.numel()
        dev = engine.device
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, off in zip(ps, engine._flat_offsets):          # the bucket's layout: every tensor on a 256-byte boundary
            view = self.flat_param[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)

    def zero_grad(self, set_to_none=False):
        pass                                    # the engine overwrites every gradient each step

    def step(self):
        self.t += 1
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(self.engine.device):
            st = ctypes.c_void_p(torch.cuda.current_stream(self.engine.device).cuda_stream)
            _lib.call("icrl_adam_flat", st, self.flat_param.numel(), p(self.flat_param), p(self.engine.flat_grad),
                      p(self.exp_avg), p(self.exp_avg_sq), self.lr, self.betas[0], self.betas[1], self.eps, self.t,
                      self.engine.launches.ref)
