#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2y2
( timeout 900 python -m pytest tests -q -x -m gpu ) > ${O}_pytest.out 2>&1; echo "pytest rc=$?"; tail -4 ${O}_pytest.out | cut -c1-300
( timeout 300 python __graft_entry__.py --smoke ) > ${O}_smoke.out 2>&1; echo "smoke rc=$?"; tail -3 ${O}_smoke.out
( timeout 300 python tools/soak.py --calls 25 --tag graph ) > ${O}_soak.out 2>&1; echo "soak rc=$?"; tail -2 ${O}_soak.out | cut -c1-300
( timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > ${O}_bench.out 2> ${O}_bench.err; echo "bench rc=$?"; grep '^{' ${O}_bench.out | cut -c1-250
( SFB_GRAPH=0 timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline ) > ${O}_bench_g0.out 2>/dev/null; grep '^{' ${O}_bench_g0.out | cut -c1-120
