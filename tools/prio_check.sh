#!/bin/bash
# gpurun helper: step time of the plain single-graph step under stream-priority variants (same box)
mkdir -p gpurun_out
for v in "RPO_SIDE_PRIO=0 RPO_MAIN_PRIO=0" "RPO_SIDE_PRIO=-1 RPO_MAIN_PRIO=0" "RPO_SIDE_PRIO=0 RPO_MAIN_PRIO=-1" "RPO_SIDE_PRIO=-2 RPO_MAIN_PRIO=-1" "RPO_SIDE_PRIO=0 RPO_MAIN_PRIO=0"; do
  env $v timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/prio.log 2>&1
  echo "$v: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/prio.log | head -1)"
done
