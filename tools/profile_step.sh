#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of two training steps + full ncu captures of the
# masked-attention forward kernel and the dominant GEMM.  Outputs land in gpurun_out/.
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 900 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline \
    > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches.txt > /dev/null
ncu --set full --clock-control none --import-source on -k regex:ro_attn_fwd -s 40 -c 2 -f \
    -o gpurun_out/${TAG}_attn_fwd python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline \
    > gpurun_out/${TAG}_attn_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 300 -c 8 -f \
    -o gpurun_out/${TAG}_gemm python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline \
    > gpurun_out/${TAG}_gemm_prof.log 2>&1
ls -la gpurun_out
