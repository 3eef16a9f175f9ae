#!/bin/bash
# GPU check: (optional microbench) parity tests, smoke, small and full bench.  Everything is bounded
# by `timeout`; the chain kernels carry their own watchdog.
mkdir -p gpurun_out
if [ "$1" = "micro" ]; then
  [ -x scripts/xchg_bench ] && { timeout 300 ./scripts/xchg_bench > gpurun_out/xchg_bench.log 2>&1; echo "xchg exit $?" >> gpurun_out/xchg_bench.log; }
  [ -x scripts/chain_micro ] && { timeout 300 ./scripts/chain_micro > gpurun_out/chain_micro.log 2>&1; echo "micro exit $?" >> gpurun_out/chain_micro.log; }
  timeout 200 python scripts/tc_accuracy_probe.py > gpurun_out/tc_probe.log 2>&1
fi
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/tests_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/tests_gpu.log
tail -25 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --batch 256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b256.log 2>&1
echo "bench256 exit $?" >> gpurun_out/bench_b256.log
tail -3 gpurun_out/bench_b256.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b4096.log 2>&1
echo "bench4096 exit $?" >> gpurun_out/bench_b4096.log
tail -3 gpurun_out/bench_b4096.log
