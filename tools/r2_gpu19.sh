#!/bin/bash
# attention second form (lib_attn2.so) vs first form: stand-alone timing at the four depth shapes, kernel parity test, per-op table
mkdir -p gpurun_out
O=gpurun_out/r2t
L2=$PWD/syncfusion_b200/lib_attn2.so
for shape in "16 2048" "16 1024" "16 512" "16 256"; do
  set -- $shape
  echo "base : $(python tools/attn_timeline.py --batch $1 --tokens $2 2>&1 | tail -1)"
  echo "attn2: $(SFB_LIB=$L2 python tools/attn_timeline.py --batch $1 --tokens $2 2>&1 | tail -1)"
done
( SFB_LIB=$L2 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "attention" ) > ${O}_pytest_attn2.out 2>&1; echo "attn2 attention tests rc=$?"; tail -3 ${O}_pytest_attn2.out | cut -c1-300
