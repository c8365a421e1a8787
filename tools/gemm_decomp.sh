#!/bin/bash
# gpurun helper: where a big GEMM's time goes (RPO_GEMM_DEBUG switches of the CTA-pair kernel; results are garbage)
TAG=${1:-g}
mkdir -p gpurun_out
for d in 0 0x100 0x200 0x300 0x400; do
  RPO_GEMM_DEBUG=$d timeout 300 python tools/kernel_bench.py --only gemm > gpurun_out/${TAG}_kb_$d.log 2>&1
done
paste <(grep "^gemm v" gpurun_out/${TAG}_kb_0.log | awk '{print $2, $9}') <(grep "^gemm v" gpurun_out/${TAG}_kb_0x100.log | awk '{print $9}') <(grep "^gemm v" gpurun_out/${TAG}_kb_0x200.log | awk '{print $9}') <(grep "^gemm v" gpurun_out/${TAG}_kb_0x300.log | awk '{print $9}') <(grep "^gemm v" gpurun_out/${TAG}_kb_0x400.log | awk '{print $9}')
