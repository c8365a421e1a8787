#!/bin/bash
# gpurun --gpus 2 helper: exchange modes of the 2-rank step one by one (bounded, logged), then the pytest case.
#   tools/n2_check.sh ["False False" "True False" ...]   (shard the text tower?, peer: False = NCCL only / None = try peer memory)
mkdir -p gpurun_out/n2
if [ $# -eq 0 ]; then set -- "False False" "True False" "False None" "True None"; fi
for mode in "$@"; do
  tag=$(echo $mode | tr ' ' '_')
  timeout 120 python tools/n2_modes.py $mode > gpurun_out/n2/mode_$tag.log 2>&1
  echo "mode $mode rc=$?"; tail -2 gpurun_out/n2/mode_$tag.log | cut -c1-300
done
timeout 400 python -m pytest tests/test_gpu_text_shard.py -m gpu -q -s -k two_process > gpurun_out/n2/pytest_two_process.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/n2/pytest_two_process.log | cut -c1-400
