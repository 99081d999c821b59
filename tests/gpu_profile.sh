#!/bin/bash
# ncu evidence: launch list with per-launch device time for ~2 steady-state steps of the bench command.
mkdir -p gpurun_out
CMD="python bench.py --batch ${BATCH:-256} --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-2600} -c ${NCU_COUNT:-1700} --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"; tail -n 2 gpurun_out/plain.log | cut -c1-300
