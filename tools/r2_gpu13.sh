#!/bin/bash
# Round-2 GPU session 13: ring-depth variants of the 1 x 1 residual ops; read-only bulk wait at exit.
mkdir -p gpurun_out
O=gpurun_out/r2m
run() { name=$1; shift; env "$@" python tools/op_profile.py > ${O}_prof_$name.txt 2>&1; echo "== $name: $(head -1 ${O}_prof_$name.txt)"; grep -A60 "by (kind" ${O}_prof_$name.txt | grep "^sk" | awk '{t[$3]+=$6} END {printf "   "; for (k in t) printf "%s %.0f  ", k, t[k]; printf "\n"}'; }
run base X=1
run r333 SFB_SK_RINGS1R=3,3,3
run r233 SFB_SK_RINGS1R=2,3,3
run r324 SFB_SK_RINGS1R=3,2,4
run noepi12_r333 SFB_NO_EPI12=1 SFB_SK_RINGS1R=3,3,3
run noepi12 SFB_NO_EPI12=1
( timeout 300 python -m pytest tests/test_gpu_scale.py -q -x -k "full_shape or soak" ) > ${O}_pytest.out 2>&1; echo "pytest rc=$?"; tail -2 ${O}_pytest.out
