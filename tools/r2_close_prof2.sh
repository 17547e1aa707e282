#!/bin/bash
# Round-2 closing session, part 2: full-set ncu captures of the streaming-K launches furthest below the roofline (1 x 1
# ops at depth 4: inject = LayerNorm fold + Modulation residual, qkv, out = + fp32 residual) and of a depth-7 conv
# (one tile per CTA).  sk launch ordinals inside one evaluation follow the plan order (tools/op_profile.py table):
# d3 0-6, d4 7-17 (down, then conv1 conv2 inject qkv out per item), d5 18-28, d6 29-39, d7 40-; second evaluation = +123.
mkdir -p gpurun_out
O=gpurun_out/r2c
one() { # name kernel-regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o ${O}_prof_$1 python tools/ncu_step.py --steps 2 > ${O}_ncu_$1.log 2>&1; echo "ncu full $1 rc=$?"
  ncu -i ${O}_prof_$1.ncu-rep --page details --csv > ${O}_ncu_full_$1.csv 2>/dev/null
  python tools/ncu_sass.py ${O}_prof_$1.ncu-rep 40 > ${O}_ncu_sass_$1.txt 2>&1
  rm -f ${O}_prof_$1.ncu-rep
}
one d4inject sk_kernel 133
one d4qkv sk_kernel 134
one d4out sk_kernel 135
one d7conv1 sk_kernel 164
