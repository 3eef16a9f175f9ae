#!/usr/bin/env python
"""A2C caption-training throughput on B200 (BASELINE.json metric) and the CPU reference arm.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

Workload (config.workload): BASELINE.json configs[3] -- the full A2C training step (rollout +
embedding reward + value + loss + backward + gradient all-reduce + Adam), global batch 4096 captions,
max_len 20 (S = 19 sampled steps), rows sharded contiguously over the N ranks (strong scaling).
Synthetic data of that shape and random-init weights of the reference architecture (the
models_pretrained/*.pt blobs are absent from the reference checkout).

One JSON line on rank 0.  `value` = captions/s with the minibatch already in HBM; `e2e` = the same
step driven through the public API from pinned host buffers, including the H2D copies and the D2H
read of the loss.  Timing: CUDA events on the launching stream bracketed by barrier + synchronize,
max over ranks.  The working set per step (activation stash of the serial chains, GBs) is far larger
than the 126 MB L2, so no explicit L2 flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, L_CAP = 512, 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="global batch (captions per step)")
    ap.add_argument("--cpu-sample", type=int, default=256, help="rows of the workload timed on the host cores")
    ap.add_argument("--chain-shards", type=int, default=1, choices=[1, 2, 4, 8],
                    help="value/reward recurrences per rank (1 = the reference's single carried-state chain)")
    ap.add_argument("--chain-segments", type=int, default=32, choices=[1, 2, 4, 8, 16, 32],
                    help="lockstep pieces of the single carried-state chain (verified warm-up; 1 = serial kernels only)")
    ap.add_argument("--chain-warmup", type=int, default=256, help="warm-up positions of every chain piece")
    ap.add_argument("--sharded-leg", action="store_true", help="also time 8 zero-state row shards per rank")
    ap.add_argument("--no-serial-leg", action="store_true", help="skip the serial-kernel leg (chain_segments = 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stamps, self.proc = index, [], [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            if len(p) >= 9 and p[0] == str(self.index):
                self.rows.append(p)
                self.stamps.append(time.time())

    def stop(self, t0=None, t1=None):
        """Samples taken inside [t0, t1] (the timed region).  nvidia-smi needs ~0.5 s to start, so the sampler is
        started before the warm-up steps; if the timed region is shorter than one sampling period the samples of the
        warm-up steps (the same kernels, the same load) are used and `window` says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        window = "timed region"
        if t0 is not None:
            inside = [r for r, ts in zip(self.rows, self.stamps) if t0 <= ts <= t1 + 0.05]
            if inside:
                self.rows = inside
            else:
                window = "warm-up + timed region (timed region shorter than the 200 ms sampling period)"
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "window": window}


def workload(batch):
    from icrl_b200 import synth       # seeded synthetic inputs (product-side helper; oracle/ is not touched here)
    f, c = synth.make_inputs(100, batch, L_CAP)
    u = synth.make_uniforms(100, L_CAP - 1, batch)
    return f, c, u


def make_nets(seed, dev):
    """Drop-in modules with random-init weights of the reference architecture (the .pt blobs are absent)."""
    import icrl_b200.models as M
    from icrl_b200 import synth
    w = synth.make_weights(seed)
    w2i = synth.word_to_idx()
    P, V, R = M.PolicyNetwork(w2i), M.ValueNetwork(w2i), M.RewardNetwork(w2i)
    P.load_state_dict(w["policy"])
    V.load_state_dict(w["value"])
    R.load_state_dict(w["reward"])
    R.requires_grad_(False)
    R.train(False)
    return M.AdvantageActorCriticNetwork(V, P).to(dev), R.to(dev)


def cpu_reference(batch_sample, steps, warmup):
    """The reference algorithm as executed (oracle/ref_port: per-step prefix re-runs, batch-as-time RNN
    calls, numpy sampling, autograd backward, Adam) on the host cores; returns (captions/s, cores, s/step)."""
    from oracle import ref_port        # the one place bench.py executes oracle/: the CPU reference being timed
    from icrl_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nets = ref_port.Nets(synth.make_weights(0))
    opt = torch.optim.Adam([p for _, p in nets.named_trainable()], lr=1e-4)
    f, c, u = workload(batch_sample)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        ref_port.a2c_minibatch(nets, f, c, u)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return batch_sample / dt, torch.get_num_threads(), dt


def main_reference(args, rank):
    if rank != 0:
        return
    B = min(args.cpu_sample, args.batch)
    cps, cores, dt = cpu_reference(B, args.steps, args.warmup)
    sample = "%d of %d captions per step (cost is linear in rows: serial RNN chains), L=%d, %d steps" % (
        B, args.batch, L_CAP, args.steps)
    print(json.dumps({
        "impl": "reference", "metric": "A2C train captions/sec", "value": cps, "unit": "captions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[3]: full A2C training step, global batch %d, max_len %d" % (args.batch, L_CAP)},
        "cpu_baseline": {"value": cps, "unit": "captions/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": cps, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return main_reference(args, rank)

    import ctypes
    import torch.distributed as dist
    from icrl_b200 import _lib
    from icrl_b200.dp import DataParallelA2C, shard_bounds
    from icrl_b200.engine import A2CEngine

    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    A, R = make_nets(0, dev)
    opt = torch.optim.Adam(A.parameters(), lr=1e-4)
    eng = A2CEngine(A, R, chain_shards=args.chain_shards, chain_segments=args.chain_segments, chain_warmup=args.chain_warmup)
    dp = DataParallelA2C(eng, opt)
    B = args.batch
    S = L_CAP - 1
    f, c, u = workload(B)
    lo, hi = shard_bounds(B, rank, world)
    fl, cl, ul = f[lo:hi], c[lo:hi], np.ascontiguousarray(u[:, lo:hi])
    plan = (1, S)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run_step, steps, warmup):
        for _ in range(warmup):
            run_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launches.value
        e0.record()
        for _ in range(steps):
            run_step()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, (eng.launches.value - l0) // steps

    # ---- leg 1: minibatch resident in HBM
    prep = eng.prepare(fl, cl, ul, plan=plan)
    clocks = ClockSampler(local)
    eng.phase_events = None

    def step_resident():
        dp.step(prep, global_rows=B, check=False)

    clocks.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    eng.phase_events = []
    t_wall0 = time.time()
    ms_step, launches = timed(step_resident, args.steps, 0)
    clk = clocks.stop(t_wall0, time.time())
    phases = {k: sum(v) / len(v) for k, v in eng.phase_times_ms().items()}
    eng.phase_events = None
    _lib.call("icrl_chain_check", ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
              ctypes.c_void_p(eng.sync_state.data_ptr()))
    seg_layout = eng.segment_layout                      # (pieces, value segment, reward segment, warm-up) or None = serial kernels
    seg_ok = torch.tensor([1 if eng.segments_verified() else 0], device=dev)
    if world > 1:
        dist.all_reduce(seg_ok, op=dist.ReduceOp.MIN)
    if int(seg_ok.item()) == 0:
        # some rank's warm-up check failed inside the timed loop: those steps are not the reference's numbers.
        # Time the serial kernels instead (every rank takes this branch together).
        eng.chain_segments = 1
        eng.phase_events = []
        ms_step, launches = timed(step_resident, args.steps, 1)
        phases = {k: sum(v) / len(v) for k, v in eng.phase_times_ms().items()}
        eng.phase_events = None
        seg_layout = None

    # ---- leg 2: end to end through the public API from pinned host buffers
    e2e = None
    if not args.no_e2e:
        fh = torch.from_numpy(fl).pin_memory()
        uh = torch.from_numpy(ul).pin_memory()
        sink = []

        def step_e2e():
            res = dp.step(fh, cl, uh, global_rows=B, plan=plan)
            sink.append(res.loss)                     # D2H read of the step's result

        ms_e2e, _ = timed(step_e2e, args.steps, 1)
        e2e = {"value": B / (ms_e2e * 1e-3), "unit": "captions/s",
               "h2d_bytes_per_step": int(fl.nbytes + (hi - lo) * 4 + ul.nbytes), "d2h_bytes_per_step": 8}

    # ---- leg 3 (reported separately, never the headline): the same workload with 8 chain shards per rank
    sharded = None
    serial_leg = None
    if seg_layout is not None and not args.no_serial_leg:
        eng1 = A2CEngine(A, R, chain_segments=1)
        dp1 = DataParallelA2C(eng1, opt)
        ms1, _ = timed(lambda: dp1.step(prep, global_rows=B, check=False), min(args.steps, 2), 1)
        _lib.call("icrl_chain_check", ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
                  ctypes.c_void_p(eng1.sync_state.data_ptr()))
        serial_leg = {"value": B / (ms1 * 1e-3), "unit": "captions/s", "ms_per_step": ms1, "steps": min(args.steps, 2),
                      "note": "the same step on the serial chain kernels (chain_segments=1): one CTA group walks the whole "
                              "carried-state chain position by position"}
        del eng1, dp1
        eng._attach_grads()
    if args.chain_shards == 1 and args.sharded_leg and (hi - lo) % 8 == 0:
        eng8 = A2CEngine(A, R, chain_shards=8, chain_segments=1)
        dp8 = DataParallelA2C(eng8, opt)
        ms8, _ = timed(lambda: dp8.step(prep, global_rows=B, check=False), args.steps, 2)
        _lib.call("icrl_chain_check", ctypes.c_void_p(torch.cuda.current_stream().cuda_stream),
                  ctypes.c_void_p(eng8.sync_state.data_ptr()))
        sharded = {"chain_shards_per_rank": 8, "value": B / (ms8 * 1e-3), "unit": "captions/s", "ms_per_step": ms8,
                   "note": "value/reward recurrences restart from zero state every local_batch/8 rows and run in lockstep "
                           "(numerically = the reference on 8x%d row shards with averaged gradients, i.e. %d-rank data "
                           "parallel, SURVEY 8e); NOT the single carried-state chain of `value`" % (world, 8 * world)}
        del eng8, dp8
        eng._attach_grads()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- tensor-pipe figure of the persistent decode kernel (policy_decode_kernel: all S steps in one launch).
    # Timed live as the policy_fwd phase (h0 GEMM + h split + the decode kernel; the kernel is >95 % of it).
    decode = None
    try:
        Bl_ = hi - lo
        us = phases.get("policy_fwd", 0.0) * 1e3
        n_cell = S                                    # p0 = 1: S cell steps, each followed by a vocab projection
        alg = Bl_ * (n_cell * 2.0 * 2048 * 512 + S * 2.0 * 1004 * 512)
        rows_pad = ((Bl_ + 127) // 128) * 128
        executed = 3.0 * rows_pad * (n_cell * 2.0 * 2048 * 512 + S * 2.0 * 1024 * 512)
        decode = {"kernel": "policy_decode_kernel (gate GEMM + cell update + vocab GEMM + softmax + sampling, all %d steps)" % S,
                  "bound": "tensor", "achieved": alg / us / 1e6, "executed_tflops": executed / us / 1e6, "unit": "TFLOP/s",
                  "us_per_launch": us, "rows": Bl_, "mma_passes": 3,
                  "note": "achieved counts the algorithmic 2MNK of the recurrent gate GEMM and the vocab GEMM once; the "
                          "tensor pipe executes 3x that (2-part fp16 split) on rows padded to 128 and V padded to 1024; "
                          "ncu sm__pipe_tensor_cycles_active of this kernel: profiles/"}
    except Exception as exc:                                   # never let the side measurement kill the bench line
        decode = {"error": str(exc)}

    # ---- roofline of the dominant kernel (serial chains: latency bound -- see DESIGN.md)
    Bl = hi - lo
    Tv, Tr = Bl * 190, Bl * 209
    fwd_ms, bwd_ms = phases.get("chains_fwd_fused", 0.0), phases.get("chain_lstm_bwd", 0.0)
    ktraffic = None
    if bwd_ms >= fwd_ms:
        kname, kms, kbytes, ksteps = "chain_lstm_bwd_kernel", bwd_ms, Tv * (6 * H * 4 + 4 + 4 * H * 4), Tv
    else:
        kname, kms = "chains_fwd_fused_kernel", fwd_ms
        kbytes, ksteps = Tv * (4 * H * 4 + 6 * H * 4 + 4) + Tr * (3 * H * 4 + H * 4 + 4), max(Tv, Tr)
        # profiles/r01_chain_fwd_ncu.md: dram read+write = 695.5 MB per launch at B=256 (linear in rows)
        ktraffic = 695.5e6 * Bl / 256.0
    if seg_layout is not None:                 # lockstep pieces: one kernel step advances `pieces` chain positions
        pieces, seg_v, seg_r, warm = seg_layout
        if bwd_ms >= fwd_ms:
            bp = 16 if pieces in (16, 32) else min(pieces, 8)              # backward pieces (engine._backward)
            kname = "chain_lstm_bwd_batched8_kernel x 2 groups" if bp == 16 else "chain_lstm_bwd_batched_kernel<%d, 1> x 2 groups" % (bp // 2)
            ksteps = seg_v * pieces // bp + warm
        else:
            kname = "chains_fwd_fused_batched_kernel<%d, %d>" % ((16, 2) if pieces == 32 else (min(pieces, 8), max(pieces // 8, 1)))
            ksteps = max(seg_v, seg_r) + warm
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    if decode and "achieved" in decode:
        tp = float(peaks.get("bf16_tflops", 1590.0))
        decode.update(peak=tp, frac=decode["achieved"] / tp, frac_executed=decode["executed_tflops"] / tp,
                      peak_source="measured" if peaks else "fallback")
    ach = kbytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    roofline = {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": ktraffic, "peak_source": "measured" if peaks else "fallback",
                "ms_per_launch": kms, "serial_steps_per_launch": ksteps,
                "ns_per_serial_step": kms * 1e6 / ksteps if ksteps else None,
                "note": "serial batch-1 recurrence: latency bound, neither HBM nor tensor pipe is the limiter"
                        + ("" if seg_layout is None else "; %d pieces of the chain advance per kernel step" % seg_layout[0]),
                "latency_floor": {"one_way_l2_store_to_poll_ns": 494, "all_to_all_512_words_64_ctas_ns": 782,
                                  "register_gemv_plus_pointwise_ns": 230,
                                  "source": "profiles/r01_xchg_bench.log, profiles/r01_chain_micro.log (microbenchmarks, "
                                            "not measured in this run)"}}
    try:
        # What bounds this kernel is FP32 issue on the CUDA cores, so the same launch is also put against
        # that roof: multiply-adds of the recurrent GEMVs (forward: 4*512*512 per value position + 3*512*512 per reward
        # position; backward: the 2048 x 512 contraction per value position) over SMs x 128 lanes x 2 x the sampled SM clock.
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        mhz = float(clk.get("sm_mhz") or 1965.0)
        macs = Tv * 4.0 * H * H if kname.startswith("chain_lstm_bwd") else Tv * 4.0 * H * H + Tr * 3.0 * H * H
        peak32 = sms * 128 * 2 * mhz * 1e6 / 1e12
        ach32 = 2.0 * macs / (kms * 1e-3) / 1e12 if kms > 0 else 0.0
        roofline["fp32"] = {"achieved": ach32, "peak": peak32, "unit": "TFLOP/s", "frac": ach32 / peak32,
                            "peak_source": "nominal: %d SMs x 128 FP32 lanes x 2 x %.0f MHz" % (sms, mhz)}
    except Exception as exc:                                   # a side figure must never cost the bench line
        roofline["fp32"] = {"error": str(exc)}
    out = {
        "metric": "A2C train captions/sec", "value": B / (ms_step * 1e-3), "unit": "captions/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[3]: full A2C training step (rollout+reward+value+loss+backward+allreduce+Adam), "
                               "global batch %d, max_len %d, S=%d" % (B, L_CAP, S),
                   "global_batch": B, "local_batch": Bl, "parallelism": "dp%d" % world, "vocab": 1004,
                   "chain_shards_per_rank": args.chain_shards,
                   "chain_segments": None if seg_layout is None else
                   {"pieces": seg_layout[0], "value_segment": seg_layout[1], "reward_segment": seg_layout[2],
                    "warmup": seg_layout[3], "tolerance": eng.chain_tol,
                    "checked_max": dict(zip(("value_h", "value_c", "reward_h", "joint_dgates", "dh_take"),
                                            eng.segment_stats["max_err"])),
                    "fallbacks_to_serial": eng.segment_stats["fallbacks"],
                    "steps_checked": eng.segment_stats["segmented_steps"]},
                   "l2_flush": "not needed: per-step working set (chain stash, GBs) >> 126 MB L2"},
        "clocks": clk, "gpu_launches": int(launches), "phases_ms": phases, "roofline": roofline,
        "roofline_decode": decode,
    }
    if e2e:
        out["e2e"] = e2e
    if serial_leg:
        out["serial_chain"] = serial_leg
    if sharded:
        out["sharded"] = sharded
    if world == 1 and not args.no_cpu_baseline:
        Bs = min(args.cpu_sample, B)
        cps, cores, dt = cpu_reference(Bs, 1, 0)
        out["cpu_baseline"] = {"value": cps, "unit": "captions/s", "cores": cores, "kind": "port",
                               "sample": "1 step of %d of the %d captions (%.1f s); reference cost is linear in rows" % (Bs, B, dt)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
