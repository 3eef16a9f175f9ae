#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step, then full captures of the chain kernels and of the
# persistent decode kernel.  Each ncu pass runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
CMD="python bench.py --batch 256 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-sharded-leg"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain -s 6 -c 2 -o gpurun_out/prof_chain $CMD > gpurun_out/ncu_chain.log 2>&1
echo "chain capture exit $?"
CMD4="python bench.py --batch 4096 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-sharded-leg"
$CMD4 > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:policy_decode -s 3 -c 1 -o gpurun_out/prof_decode $CMD4 > gpurun_out/ncu_decode.log 2>&1
echo "decode capture exit $?"
python scripts/decode_profile.py 4096 > gpurun_out/decode_profile_4096.log 2>&1
ls -la gpurun_out
