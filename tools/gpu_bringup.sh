#!/bin/bash
# Runs the staged bring-up checks on the GPU box, one process per stage (a trapped kernel poisons its context).
mkdir -p gpurun_out
LOG=gpurun_out/bringup.log
: > $LOG
python -c "import torch; print(torch.cuda.get_device_name(0), torch.cuda.device_count())" >> $LOG 2>&1
for st in "$@"; do
  echo "=== $st" >> $LOG
  timeout 300 python tools/gpu_check.py $st >> $LOG 2>&1
  echo "=== $st exit $?" >> $LOG
done
tail -n 150 $LOG
