#!/bin/bash
# One GPU round: parity tests, bench, ncu launch list.  Usage: bash tools/gpu_round.sh [tests] [bench] [ncu]
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests) timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -n 15 gpurun_out/pytest_gpu.log;;
bench) timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err;;
benchcfg) timeout 900 python bench.py --steps 2 --warmup 3 --scale 2.0 --no-cpu-baseline > gpurun_out/bench_cfg.json 2> gpurun_out/bench_cfg.err; echo "bench exit $?"; cat gpurun_out/bench_cfg.json; tail -n 5 gpurun_out/bench_cfg.err;;
ref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json;;
ncu) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 420 --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py --steps 2 > gpurun_out/ncu_step.log 2>&1; tail -n 3 gpurun_out/ncu_step.log;;
ncufull32) timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_tc_kernel<__nv_bfloat16, .int.32>' -s 1 -c 2 -f -o gpurun_out/prof_gemm32 python tools/ncu_step.py --steps 1 > gpurun_out/ncu_full32.log 2>&1; tail -n 2 gpurun_out/ncu_full32.log;;
ncufull256) timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_tc_kernel<__nv_bfloat16, .int.256>' -s 20 -c 2 -f -o gpurun_out/prof_gemm256 python tools/ncu_step.py --steps 1 > gpurun_out/ncu_full256.log 2>&1; tail -n 2 gpurun_out/ncu_full256.log;;
ncuattn) timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 0 -c 1 -f -o gpurun_out/prof_attn python tools/ncu_step.py --steps 1 > gpurun_out/ncu_attn.log 2>&1; tail -n 2 gpurun_out/ncu_attn.log; ls -la gpurun_out;;
opprof) timeout 600 python tools/op_profile.py > gpurun_out/op_profile.txt 2> gpurun_out/op_profile.err; tail -n 45 gpurun_out/op_profile.txt;;
ncusk) timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section SchedulerStats --clock-control none -k regex:sk_kernel -s 123 -c 123 -f -o gpurun_out/prof_sk python tools/ncu_step.py --steps 2 > gpurun_out/ncu_sk.log 2>&1; tail -n 2 gpurun_out/ncu_sk.log
  ncu -i gpurun_out/prof_sk.ncu-rep --page raw --csv > gpurun_out/prof_sk_raw.csv 2>/dev/null; rm -f gpurun_out/prof_sk.ncu-rep; ls -la gpurun_out;;
timeline) timeout 600 python tools/sk_timeline.py --ops "${SK_OPS:-4:out,6:qkv,6:conv1,7:inject}" > gpurun_out/sk_timeline.txt 2>&1; tail -n 3 gpurun_out/sk_timeline.txt;;
ncuone) # NCU_SKIP=<launch index> NCU_NAME=<tag> NCU_KERNEL=<regex>: one launch, full set + source; only CSV exports travel back
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-sk_kernel} -s ${NCU_SKIP:-134} -c 1 -f -o gpurun_out/prof_${NCU_NAME:-sk_one} python tools/ncu_step.py --steps 2 > gpurun_out/ncu_${NCU_NAME:-sk_one}.log 2>&1; tail -n 1 gpurun_out/ncu_${NCU_NAME:-sk_one}.log
  ncu -i gpurun_out/prof_${NCU_NAME:-sk_one}.ncu-rep --page details --csv > gpurun_out/ncu_full_${NCU_NAME:-sk_one}.csv 2>/dev/null
  python tools/ncu_sass.py gpurun_out/prof_${NCU_NAME:-sk_one}.ncu-rep 40 > gpurun_out/ncu_sass_${NCU_NAME:-sk_one}.txt 2>&1
  rm -f gpurun_out/prof_${NCU_NAME:-sk_one}.ncu-rep;;
smoke) timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -n 5 gpurun_out/smoke.log;;
esac
done
