#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2s
one() { # name kernel-regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o ${O}_prof_$1 python tools/ncu_step.py --steps 2 > ${O}_ncu_$1.log 2>&1; echo "ncu full $1 rc=$?"
  ncu -i ${O}_prof_$1.ncu-rep --page details --csv > ${O}_ncu_full_$1.csv 2>/dev/null
  python tools/ncu_sass.py ${O}_prof_$1.ncu-rep 45 > ${O}_ncu_sass_$1.txt 2>&1
  rm -f ${O}_prof_$1.ncu-rep
}
one d0conv1 d0_gn_conv1_kernel 2
one d0tail d0_tail_kernel 2
