#!/bin/bash
# Round-2 GPU session 5: wait-log level A/B/C (0 trap only, 1 light record [default], 2 full records), fault path at levels 1 and 2,
# GPU post-processing tests, whole GPU suite.
mkdir -p gpurun_out
O=gpurun_out/r2e
run() {
  local name=$1 to=$2; shift 2
  local t0=$(date +%s)
  ( timeout -k 10 $to env "$@" ) > ${O}_${name}.out 2> ${O}_${name}.err
  local rc=$?
  echo "== $name rc=$rc wall=$(( $(date +%s) - t0 ))s" | tee -a ${O}_summary.txt
  tail -c 1300 ${O}_${name}.out | tee -a ${O}_summary.txt
  grep -v "^\s*$" ${O}_${name}.err | tail -n 30 | cut -c1-300 | tee -a ${O}_summary.txt
}
: > ${O}_summary.txt
L0=$PWD/syncfusion_b200/lib_waitlog0.so
L2=$PWD/syncfusion_b200/lib_waitlog2.so
for r in a b; do
run soak_l1_$r 200 X=1 python tools/soak.py --calls 25 --tag level1_$r
run soak_l0_$r 200 SFB_LIB=$L0 python tools/soak.py --calls 25 --tag level0_$r
run soak_l2_$r 200 SFB_LIB=$L2 python tools/soak.py --calls 25 --tag level2_$r
done
run fault_l1 120 X=1 python tools/fault_inject.py
run fault_l2 120 SFB_LIB=$L2 python tools/fault_inject.py
run pytest 1500 X=1 python -m pytest tests -q -m gpu --durations=8
run bench 600 X=1 python bench.py --gpus 1 --steps 20 --warmup 5
