#!/bin/bash
# Round-2 GPU session 8: fresh per-op table + device-clock timelines of the ops that bound the step.
mkdir -p gpurun_out
python tools/op_profile.py > gpurun_out/r2h_op_profile.txt 2> gpurun_out/r2h_op_profile.err
echo "op_profile rc=$?"; head -1 gpurun_out/r2h_op_profile.txt
python tools/sk_timeline.py --ops 6:conv1,7:conv1,7:conv2,4:out,4:inject,7:inject,5:qkv > gpurun_out/r2h_timeline.txt 2> gpurun_out/r2h_timeline.err
echo "timeline rc=$?"; wc -c gpurun_out/r2h_timeline.txt
