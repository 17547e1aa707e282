#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2y
( SFB_GRAPH=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -q -x -k "sample or soak or free_running or golden or determinism" ) > ${O}_pytest_graph.out 2>&1; echo "pytest(graph) rc=$?"; tail -3 ${O}_pytest_graph.out | cut -c1-300
for g in 0 1 0 1; do
  ( SFB_GRAPH=$g timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline ) > ${O}_bench_g$g.out 2> ${O}_bench_g$g.err
  echo "graph=$g: $(grep '^{' ${O}_bench_g$g.out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'clips/s e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],2), d.get('error'))")"
done
( SFB_GRAPH=1 timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --length 524288 --batch 4 ) > ${O}_bench_long_g1.out 2>&1; echo "long graph=1: $(grep '^{' ${O}_bench_long_g1.out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],2))")"
