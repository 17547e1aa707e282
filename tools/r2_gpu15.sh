#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2o
SFB_SK_PAIR=1 SFB_SK_PAIR_VERBOSE=1 python tools/sk_timeline.py --ops 6:conv1,7:conv1,5:conv1 > ${O}_timeline_pair.txt 2>&1
grep "pair mode" ${O}_timeline_pair.txt
SFB_SK_PAIR=2 python tools/sk_timeline.py --ops 5:qkv > ${O}_timeline_pair_qkv.txt 2>&1
