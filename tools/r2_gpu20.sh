#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2u
for v in base v_poly4 v_poly3 v_poly2 v_s3o3; do
  if [ $v = base ]; then L=""; else L=$PWD/syncfusion_b200/lib_$v.so; fi
  for shape in "16 2048" "16 1024" "16 512" "16 256"; do
    set -- $shape
    echo "$v: $(SFB_LIB=$L python tools/attn_timeline.py --batch $1 --tokens $2 2>&1 | tail -1)"
  done
done
( SFB_LIB=$PWD/syncfusion_b200/lib_v_poly2.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "attention" ) > ${O}_pytest_poly2.out 2>&1; echo "poly2 attention tests rc=$?"; tail -2 ${O}_pytest_poly2.out | cut -c1-300
