#!/bin/bash
# ncu evidence: (1) launch list with per-launch device time, (2) one --set full capture of the GEMM kernel.
mkdir -p gpurun_out
CMD="python bench.py --batch ${BATCH:-256} --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-9000} --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"; tail -n 3 gpurun_out/plain.log | cut -c1-400
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hct_gemm_tcgen05 -s ${NCU_SKIP:-1000} -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"; ls -la gpurun_out | tail
