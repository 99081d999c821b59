#!/bin/bash
# ncu evidence for profiles/: (1) launch list with per-launch device time for one steady-state step of the bench
# command, (2) DRAM traffic of every GEMM launch of one step, (3) one full-set capture of the attention kernels and of the
# two activation-epilogue GEMMs.  Each ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
CMD="python bench.py --batch ${BATCH:-256} --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-1900} -c ${NCU_COUNT:-1100} --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:hct_gemm -s ${GEMM_SKIP:-744} -c 248 --csv --log-file gpurun_out/gemm_traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "gemm traffic exit $?"
PB=64 python tools/attn_probe.py > gpurun_out/probe_plain.log 2>&1 &&
PB=64 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 8 -c 4 -o gpurun_out/prof_attn_final -f python tools/attn_probe.py > gpurun_out/ncu_attn_final.log 2>&1
echo "attention full capture exit $?"
python tools/gemm_probe.py > gpurun_out/gprobe_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hct_gemm -s 4 -c 2 -o gpurun_out/prof_gemm_final -f python tools/gemm_probe.py > gpurun_out/ncu_gemm_final.log 2>&1
echo "gemm full capture exit $?"
tail -n 1 gpurun_out/plain.log | cut -c1-200
