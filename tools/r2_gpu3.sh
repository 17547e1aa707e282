#!/bin/bash
# Round-2 GPU session 3: new scale / soak / fault tests, cost of the wait log (A/B vs the bare loop), per-op profile, bench, ncu launch list.
mkdir -p gpurun_out
O=gpurun_out/r2c
run() {
  local name=$1 to=$2; shift 2
  local t0=$(date +%s)
  ( timeout -k 10 $to env "$@" ) > ${O}_${name}.out 2> ${O}_${name}.err
  local rc=$?
  echo "== $name rc=$rc wall=$(( $(date +%s) - t0 ))s" | tee -a ${O}_summary.txt
  tail -c 1500 ${O}_${name}.out | tee -a ${O}_summary.txt
  grep -v "^\s*$" ${O}_${name}.err | tail -n 30 | cut -c1-300 | tee -a ${O}_summary.txt
}
: > ${O}_summary.txt
run pytest 1500 X=1 python -m pytest tests -x -q -m gpu -s --durations=12
run soak_new 200 X=1 python tools/soak.py --calls 30 --tag new
run soak_simple 200 SFB_LIB=$PWD/syncfusion_b200/lib_simplewait.so python tools/soak.py --calls 30 --tag simplewait
run soak_new2 200 X=1 python tools/soak.py --calls 30 --tag new2
run soak_simple2 200 SFB_LIB=$PWD/syncfusion_b200/lib_simplewait.so python tools/soak.py --calls 30 --tag simplewait2
run opprof 300 X=1 python tools/op_profile.py
run bench 600 X=1 python bench.py --gpus 1 --steps 20 --warmup 5
run bench_cfg 600 X=1 python bench.py --gpus 1 --steps 10 --warmup 3 --scale 2.0 --no-cpu-baseline
run ncu 900 X=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/r2c_launches.csv python tools/ncu_step.py --steps 2
