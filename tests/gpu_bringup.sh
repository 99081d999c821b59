#!/bin/bash
# Run every GPU test file in its own process (a trapped kernel poisons only that file's context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rc=0
for f in tests/test_gpu_kernels.py tests/test_gpu_models.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -x --tb=short "$@" > gpurun_out/$n.log 2>&1
  r=$?; echo "$f -> exit $r"; tail -n 25 gpurun_out/$n.log
  [ $r -ne 0 ] && rc=$r
done
exit $rc
