#!/bin/bash
mkdir -p gpurun_out
rc=0
run() {
  local name=$1; shift
  timeout 900 "$@" > gpurun_out/$name.log 2>&1
  local r=$?; echo "== $name -> exit $r"; tail -n ${TAILN:-30} gpurun_out/$name.log | cut -c1-3000
  [ $r -ne 0 ] && rc=$r
}
run tests python -m pytest tests -q -m gpu --tb=short
run smoke python __graft_entry__.py smoke
run bench256 python bench.py --batch 256 --steps 10 --warmup 3
exit $rc
