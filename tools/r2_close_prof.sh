#!/bin/bash
# Round-2 closing session at the final build: full-set ncu captures of the shipped attention kernel (attn2_tc_kernel; the
# earlier script still named attn_tc_kernel and captured nothing), back-to-back soak, side configurations.
mkdir -p gpurun_out
O=gpurun_out/r2c
one() { # name kernel-regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o ${O}_prof_$1 python tools/ncu_step.py --steps 2 > ${O}_ncu_$1.log 2>&1; echo "ncu full $1 rc=$?"
  ncu -i ${O}_prof_$1.ncu-rep --page details --csv > ${O}_ncu_full_$1.csv 2>/dev/null
  python tools/ncu_sass.py ${O}_prof_$1.ncu-rep 40 > ${O}_ncu_sass_$1.txt 2>&1
  rm -f ${O}_prof_$1.ncu-rep
}
# 20 attention launches per U-Net evaluation: skip the first evaluation; launch 0 = d4, 2 = d5, 4 = d6 of the second one.
one attn_d4 attn2_tc_kernel 20
one attn_d5 attn2_tc_kernel 22
( timeout 200 python tools/soak.py --calls 100 --tag head_100 ) > ${O}_soak.out 2> ${O}_soak.err; echo "soak rc=$?"; tail -1 ${O}_soak.out
( timeout 200 python tools/soak.py --calls 40 --scale 2.0 --tag head_cfg40 ) >> ${O}_soak.out 2>> ${O}_soak.err; echo "soak cfg rc=$?"; tail -1 ${O}_soak.out
side() { # name args...
  n=$1; shift
  ( timeout 240 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline "$@" ) > ${O}_side_$n.out 2> ${O}_side_$n.err; echo "side $n rc=$?"; grep '^{' ${O}_side_$n.out | cut -c1-120
}
side cfg --scale 2.0
side fp32 --precision fp32
side transpose --upsample-mode transpose
side long --length 524288 --batch 4
side long_cfg --length 524288 --batch 4 --scale 2.0
