#!/bin/bash
# Bring-up runner: separate processes so a trapped kernel poisons only its own CUDA context.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rc=0
run() {
  local name=$1; shift
  timeout 900 "$@" > gpurun_out/$name.log 2>&1
  local r=$?; echo "== $name -> exit $r"; tail -n ${TAILN:-30} gpurun_out/$name.log
  [ $r -ne 0 ] && rc=$r
}
run kernels python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short
run models python -m pytest tests/test_gpu_models.py -q -m gpu --tb=short
run smoke python __graft_entry__.py smoke
run bench64 python bench.py --batch 64 --steps 3 --warmup 3
run bench256 python bench.py --batch 256 --steps 3 --warmup 3 --no-cpu-baseline
exit $rc
