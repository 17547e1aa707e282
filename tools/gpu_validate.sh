#!/bin/bash
# One-box validation: GPU parity suite, smoke, per-op table, the driver's bench command.   bash tools/gpu_validate.sh [tag]
mkdir -p gpurun_out
O=gpurun_out/${1:-val}
( timeout 900 python -m pytest tests -q -x -m gpu ) > ${O}_pytest.out 2>&1; echo "pytest rc=$?"; tail -2 ${O}_pytest.out | cut -c1-200
( timeout 300 python __graft_entry__.py --smoke ) > ${O}_smoke.out 2>&1; echo "smoke rc=$?"; tail -3 ${O}_smoke.out
python tools/op_profile.py > ${O}_op_profile.txt 2> ${O}_op_profile.err; echo "op_profile rc=$?"; head -1 ${O}_op_profile.txt
grep -A12 "by (kind" ${O}_op_profile.txt | tail -12
( timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > ${O}_bench.out 2> ${O}_bench.err; echo "bench rc=$?"; grep '^{' ${O}_bench.out | cut -c1-250
