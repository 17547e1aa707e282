#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2w
L=$PWD/syncfusion_b200/lib_v_d0p2.so
python tools/op_profile.py > ${O}_prof_base.txt 2>&1; echo "base: $(head -1 ${O}_prof_base.txt)"; grep "^d0" ${O}_prof_base.txt
SFB_LIB=$L python tools/op_profile.py > ${O}_prof_d0p2.txt 2>&1; echo "d0p2: $(head -1 ${O}_prof_d0p2.txt)"; grep "^d0" ${O}_prof_d0p2.txt
( SFB_LIB=$L timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "layerwise or free_running or full_arch" ) > ${O}_pytest.out 2>&1; echo "pytest rc=$?"; tail -2 ${O}_pytest.out | cut -c1-200
