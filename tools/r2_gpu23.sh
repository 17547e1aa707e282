#!/bin/bash
# Round-2 side configurations for BASELINE.md section 4 (one GPU each)
mkdir -p gpurun_out
O=gpurun_out/r2x
run() { name=$1; shift; ( timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline "$@" ) > ${O}_$name.out 2> ${O}_$name.err; echo "== $name rc=$?"; grep '^{' ${O}_$name.out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'clips/s  e2e', round(d['e2e']['value'],2) if d.get('e2e') else None, ' ms/step', round(d['ms_per_step'],1), ' e2e frac', round(d['roofline']['end_to_end_frac_of_bf16_peak'],3) if d.get('roofline') else None)"; }
run transpose --upsample-mode transpose
run cfg --scale 2.0
run fp32 --precision fp32
run long --length 524288 --batch 4
run long_cfg --length 524288 --batch 4 --scale 2.0
