#!/bin/bash
# Round-2 final profiling session: ncu launch list of the bench workload, DRAM traffic of the sk launches, full captures of
# the top kernels (CSV exports only), per-op table, bench line.
mkdir -p gpurun_out
O=gpurun_out/r2z
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 340 --csv --log-file ${O}_launches.csv python tools/ncu_step.py --steps 2 > ${O}_ncu_launches.log 2>&1; echo "launch list rc=$?"; tail -1 ${O}_ncu_launches.log
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sk_kernel -s 123 -c 123 --csv --log-file ${O}_sk_traffic.csv python tools/ncu_step.py --steps 2 > ${O}_ncu_traffic.log 2>&1; echo "traffic rc=$?"
one() { # name kernel-regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o ${O}_prof_$1 python tools/ncu_step.py --steps 2 > ${O}_ncu_$1.log 2>&1; echo "ncu full $1 rc=$?"
  ncu -i ${O}_prof_$1.ncu-rep --page details --csv > ${O}_ncu_full_$1.csv 2>/dev/null
  python tools/ncu_sass.py ${O}_prof_$1.ncu-rep 40 > ${O}_ncu_sass_$1.txt 2>&1
  rm -f ${O}_prof_$1.ncu-rep
}
one d6conv1 sk_kernel ${SK_SKIP:-153}
one attn_d4 attn_tc_kernel 20
one rk_d1inject rk_kernel 21
one d0conv1_tc d0_gn_conv1_tc_kernel 2
python tools/op_profile.py > ${O}_op_profile.txt 2>&1; head -1 ${O}_op_profile.txt
( timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > ${O}_bench.out 2> ${O}_bench.err; echo "bench rc=$?"; grep '^{' ${O}_bench.out | cut -c1-300
