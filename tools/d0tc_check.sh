mkdir -p gpurun_out
( SFB_D0_TC=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -q -x -k "layerwise or full_shape or many_tiles or long_form" ) > gpurun_out/d0tc_pytest.out 2>&1; echo "pytest(d0tc) rc=$?"; tail -12 gpurun_out/d0tc_pytest.out | cut -c1-250
SFB_D0_TC=1 python tools/op_profile.py > gpurun_out/d0tc_prof.txt 2>&1; head -1 gpurun_out/d0tc_prof.txt; grep "^d0" gpurun_out/d0tc_prof.txt
