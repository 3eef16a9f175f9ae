#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step, then a full capture of the chain kernels.
mkdir -p gpurun_out
CMD="python bench.py --batch 256 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain -s 6 -c 2 -o gpurun_out/prof_chain $CMD > gpurun_out/ncu_chain.log 2>&1
echo "chain capture exit $?"
tail -5 gpurun_out/ncu_chain.log
ls -la gpurun_out
