#!/bin/bash
# Round-2 GPU session 9: split A/B producers + pre-wait weight prefetch (A/B via SFB_NO_WPRE), attention back to one thread per row.
mkdir -p gpurun_out
O=gpurun_out/r2i
( timeout 900 python -m pytest tests -q -x -m gpu ) > ${O}_pytest.out 2>&1
echo "pytest rc=$?"; tail -3 ${O}_pytest.out
python tools/op_profile.py > ${O}_op_profile.txt 2> ${O}_op_profile.err; echo "op_profile rc=$?"; head -1 ${O}_op_profile.txt
SFB_NO_WPRE=1 python tools/op_profile.py > ${O}_op_profile_nowpre.txt 2>&1; head -1 ${O}_op_profile_nowpre.txt
( timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 ) > ${O}_bench.out 2> ${O}_bench.err
echo "bench rc=$?"; grep '^{' ${O}_bench.out | cut -c1-300
( SFB_NO_WPRE=1 timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 ) > ${O}_bench_nowpre.out 2> /dev/null
grep '^{' ${O}_bench_nowpre.out | cut -c1-200
python tools/sk_timeline.py --ops 6:conv1,7:conv1,4:out > ${O}_timeline.txt 2>&1
