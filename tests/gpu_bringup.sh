#!/bin/bash
mkdir -p gpurun_out
rc=0
run() {
  local name=$1; shift
  timeout 900 "$@" > gpurun_out/$name.log 2>&1
  local r=$?; echo "== $name -> exit $r"; tail -n ${TAILN:-30} gpurun_out/$name.log | cut -c1-3000
  [ $r -ne 0 ] && rc=$r
}
run attn python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -k attention
run kernels python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -k "not attention"
run models python -m pytest tests/test_gpu_models.py -q -m gpu --tb=short
run bench256 python bench.py --batch 256 --steps 5 --warmup 3 --no-cpu-baseline
exit $rc
