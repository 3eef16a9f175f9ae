"""Throughput of the BASELINE configs that are parity cases rather than bench lines (run on the GPU box):
  configs[2]  RewardNetwork embedding-cosine reward, 8192 captions x 20 tokens, 1 GPU
  configs[4]  curriculum A2C (levels 3,6,9,12,15,16), 1024 rows per rank (global 8192 on 8 GPUs), L = 20
Usage: python scripts/other_configs.py            (1 GPU; under torchrun every rank runs its 1024-row shard)"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icrl_b200 import synth
from icrl_b200.engine import A2CEngine
from icrl_b200.dp import DataParallelA2C
from bench import make_nets

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
A, R = make_nets(0, "cuda:%d" % local)
opt = torch.optim.Adam(A.parameters(), lr=1e-4)
out = {"n_gpus": world}

def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)

for mode, kw in (("serial_chain", dict(chain_segments=1)), ("chain_segments", dict()), ("chain_shards_8", dict(chain_shards=8))):
    eng = A2CEngine(A, R, **kw)
    dp = DataParallelA2C(eng, opt)
    if mode != "chain_shards_8":          # every rank runs it: timed() contains collectives
        f, c = synth.make_inputs(200, 8192, 20)
        ms = timed(lambda: eng.get_rewards(f, c))
        out["config2_rewards_b8192_" + mode] = {"ms": ms, "captions_per_s": 8192 / (ms * 1e-3), "serial_gru_steps": 8192 * 20,
                                                "pieces": eng.segment_layout[0] if eng.segment_layout else 1, "fallbacks": eng.segment_stats["fallbacks"]}
    Bl, L = 1024, 20
    f, c = synth.make_inputs(300 + rank, Bl, L)
    per_level = {}
    for level in (3, 6, 9, 12, 15, 16):
        u = synth.make_uniforms(400 + rank + level, level, Bl)
        prep = eng.prepare(f, c, u, level=level)
        ms = timed(lambda: dp.step(prep, global_rows=Bl * world, check=False))
        per_level[str(level)] = {"ms": ms, "captions_per_s": Bl * world / (ms * 1e-3), "pieces": eng.segment_layout[0] if eng.segment_layout else 1,
                                 "verified": bool(eng.segments_verified())}
    tot = sum(v["ms"] for v in per_level.values())
    out["config4_curriculum_" + mode] = {"per_level": per_level, "aggregate_captions_per_s": 6 * Bl * world / (tot * 1e-3)}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
