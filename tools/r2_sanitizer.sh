#!/bin/bash
# compute-sanitizer over one CFG U-Net evaluation of the full 8-depth architecture (tools/san_step.py: few CTAs, many
# tiles per CTA).  The library under test is a build of the SAME sources with a 2^29-poll wait bound
# (-DSFB_WAIT_BOUND_LOG2=29: the instrumented kernels run orders of magnitude slower than the shipped bound allows):
#   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -DSFB_NO_FAST_MATH -DSFB_WAIT_BOUND_LOG2=29 \
#        -Xcompiler -fPIC -shared -o tools/bin/libsfb_san.so syncfusion_b200/csrc/sfb.cu
mkdir -p gpurun_out
O=gpurun_out/r2s
export SFB_LIB=$PWD/tools/bin/libsfb_san.so SFB_GRAPH=0
( timeout 60 python tools/san_step.py ) > ${O}_plain.out 2>&1; echo "plain rc=$?"; tail -1 ${O}_plain.out
for tool in synccheck racecheck memcheck; do
  t0=$(date +%s)
  ( timeout ${SAN_TIMEOUT:-150} compute-sanitizer --tool $tool --print-limit 40 python tools/san_step.py ) > ${O}_$tool.out 2>&1
  echo "$tool rc=$? $(( $(date +%s) - t0 )) s"; grep -E "launches|ERROR SUMMARY|RACECHECK SUMMARY" ${O}_$tool.out | tail -3
done
