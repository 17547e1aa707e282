#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2g
run() {
  local name=$1 to=$2; shift 2
  local t0=$(date +%s)
  ( timeout -k 10 $to env "$@" ) > ${O}_${name}.out 2> ${O}_${name}.err
  local rc=$?
  echo "== $name rc=$rc wall=$(( $(date +%s) - t0 ))s" | tee -a ${O}_summary.txt
  tail -c 2500 ${O}_${name}.out | tee -a ${O}_summary.txt
  grep -v "^\s*$" ${O}_${name}.err | tail -n 30 | cut -c1-300 | tee -a ${O}_summary.txt
}
: > ${O}_summary.txt
run xattn 600 X=1 python -m pytest tests/test_gpu_parity.py -q -x -k cross_attention_general
run pytest 1500 X=1 python -m pytest tests -q -m gpu --durations=5
