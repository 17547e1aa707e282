#!/bin/bash
# A/B of one environment switch on the per-op table:  bash tools/ab_env.sh VAR=VALUE [tag]
mkdir -p gpurun_out
O=gpurun_out/${2:-ab}
( timeout 900 python -m pytest tests -q -x -m gpu ) > ${O}_pytest.out 2>&1; echo "pytest rc=$?"; tail -2 ${O}_pytest.out | cut -c1-200
sumk() { grep -A60 "by (kind" $1 | grep "^sk" | awk '{t[$3]+=$6} END {printf "   "; for (k in t) printf "%s %.0f  ", k, t[k]; printf "\n"}'; }
python tools/op_profile.py > ${O}_prof_on.txt 2>&1; echo "default: $(head -1 ${O}_prof_on.txt)"; sumk ${O}_prof_on.txt
env $1 python tools/op_profile.py > ${O}_prof_off.txt 2>&1; echo "$1: $(head -1 ${O}_prof_off.txt)"; sumk ${O}_prof_off.txt
( timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline ) 2>/dev/null | grep '^{' | cut -c1-140
( env $1 timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline ) 2>/dev/null | grep '^{' | cut -c1-140
