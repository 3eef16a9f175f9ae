"""Data-parallel A2C: rollouts sharded by rows across ranks, one gradient all-reduce per step.

The reference is single-process; this is the new exchange step of section 8e:
rank r owns rows [r*B/N, (r+1)*B/N) of features / captions / uniforms, runs the whole rollout,
both serial chains and the backward locally (each rank's chains start from zero state, i.e. "the
reference run on that shard"), scales its gradient seeds by 1/(B_global*S), then ONE all-reduce
(sum) over the flat fp32 gradient bucket (NCCL over NVLink on GPUs; gloo in the CPU tests),
followed by the same Adam step on every rank.  The three logged scalars ride a second, 3-float
all-reduce.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_rows, rank, world):
    """Contiguous, near-equal row shards (the first n_rows % world ranks get one extra row)."""
    base, extra = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def global_plan(captions_local, level, group=None):
    """(p0, S) from the GLOBAL batch: caplen is the max <END> column over all ranks (trainers.py:436)."""
    caps = np.asarray(captions_local)
    cols = np.nonzero(caps == 2)[1]
    caplen = torch.tensor([int(cols.max()) + 1 if cols.size else 0], dtype=torch.int64)
    if dist.is_available() and dist.is_initialized():
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        caplen = caplen.to(dev)
        dist.all_reduce(caplen, op=dist.ReduceOp.MAX, group=group)
    caplen = int(caplen.item())
    if caplen == 0:
        raise ValueError("no <END> (=2) token in the caption batch")
    return (1, caplen - 1) if level is None else (caplen - int(level), int(level))


class DataParallelA2C:
    """Wraps an engine (anything with .step(...) -> result holding 'stats', and .flat_grad)."""

    def __init__(self, engine, optimizer=None, group=None):
        self.engine, self.optimizer, self.group = engine, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def _global_rows(self, features_local, captions_local):
        local = getattr(features_local, "B", None)              # a Prepared minibatch knows its rows
        if local is None:
            local = len(captions_local) if captions_local is not None else len(features_local)
        if self.world == 1:
            return int(local)
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        n = torch.tensor([int(local)], dtype=torch.int64, device=dev)
        dist.all_reduce(n, op=dist.ReduceOp.SUM, group=self.group)
        return int(n.item())

    def step(self, features_local, captions_local=None, uniforms_local=None, global_rows=None, level=None,
             plan=None, **kw):
        """One global minibatch; every rank passes its own row shard.  Returns the local StepResult
        whose 'stats' now hold the GLOBAL loss / mean reward / mean advantage."""
        if plan is None and captions_local is not None:
            plan = global_plan(captions_local, level, self.group)
        if global_rows is None:
            # the loss is a mean over the GLOBAL batch (trainers.py:472-475): every rank must scale its gradient seeds by
            # 1 / (B_global * S), so the row count is agreed on before the step (one 8-byte all-reduce)
            global_rows = self._global_rows(features_local, captions_local)
        res = self.engine.step(features_local, captions_local, uniforms=uniforms_local, global_rows=global_rows,
                               level=level, plan=plan, **kw)
        if res is None:
            return None
        if self.world > 1:
            dist.all_reduce(self.engine.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(res["stats"], op=dist.ReduceOp.SUM, group=self.group)
        if self.optimizer is not None:
            self.optimizer.step()
        return res
