#!/bin/bash
# gpurun --gpus 2 helper: the four exchange modes of the 2-rank step one by one (bounded, logged), then the pytest case.
mkdir -p gpurun_out/n2
for mode in "False False" "True False" "False None" "True None"; do
  tag=$(echo $mode | tr ' ' '_')
  timeout 120 python tools/n2_modes.py $mode > gpurun_out/n2/mode_$tag.log 2>&1
  echo "mode $mode rc=$?"; tail -2 gpurun_out/n2/mode_$tag.log | cut -c1-300
done
