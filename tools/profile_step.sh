#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch list of two training steps + full ncu captures of the tcgen05
# masked-attention forward kernel and of the dominant (CTA-pair) GEMM.  Outputs land in gpurun_out/; summarise
# them here with tools/launch_summary.py / tools/ncu_summary.py and commit the text under profiles/.
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 900 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline \
    > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches.txt > /dev/null
# the attention kernel exactly as bench.py's roofline times it (vision shape, rotating buffers)
ncu --set full --clock-control none --import-source on -k regex:ro_attn_fwd_tc -s 8 -c 1 -f \
    -o gpurun_out/${TAG}_attn_tc python tools/kernel_bench.py --only attn --no-graph > gpurun_out/${TAG}_attn_prof.log 2>&1
# c_fc (M=7072 N=3072 K=768, bias + QuickGELU): launch 50 of the pair kernel in kernel_bench's order
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 50 -c 1 -f \
    -o gpurun_out/${TAG}_gemm_fc python tools/kernel_bench.py --only gemm --no-graph > gpurun_out/${TAG}_gemm_prof.log 2>&1
ls -la gpurun_out
