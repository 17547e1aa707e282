#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2p
( SFB_SK_PAIR=7 timeout 400 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -q -x -k "full or many_tiles or grid_limit or long_form or soak or layerwise" ) > ${O}_pytest_pair.out 2>&1
echo "pytest(pair=7) rc=$?"; tail -2 ${O}_pytest_pair.out | cut -c1-400
run() { name=$1; shift; ( timeout 120 env "$@" python tools/op_profile.py ) > ${O}_prof_$name.txt 2>&1; echo "== $name rc=$?: $(head -1 ${O}_prof_$name.txt | cut -c1-200)"; grep -A60 "by (kind" ${O}_prof_$name.txt | grep "^sk" | awk '{t[$3]+=$6} END {printf "   "; for (k in t) printf "%s %.0f  ", k, t[k]; printf "\n"}'; }
run p0 SFB_SK_PAIR=0
run p1 SFB_SK_PAIR=1
run p7 SFB_SK_PAIR=7
SFB_SK_PAIR=7 python tools/sk_timeline.py --ops 6:conv1,5:qkv,4:out > ${O}_timeline_pair.txt 2>&1
