#!/bin/bash
# Same-box A/B of the dynamic (cluster launch control) tile schedule: RPO_GEMM_DYNAMIC = 0 static, 1 pair kernel,
# 2 single-CTA kernel, 3 both.  Runs under gpurun; logs land in gpurun_out/.
TAG=${1:-d1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
i=0
for d in 0 3 1 2 0 3; do
  i=$((i+1))
  RPO_GEMM_DYNAMIC=$d timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_${i}_dyn${d}.log 2>&1
  echo "dyn=$d $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench_${i}_dyn${d}.log | head -1)"
done
RPO_GEMM_DYNAMIC=0 timeout 300 python tools/kernel_bench.py --only gemm > gpurun_out/${TAG}_kb_dyn0.log 2>&1
RPO_GEMM_DYNAMIC=3 timeout 300 python tools/kernel_bench.py --only gemm > gpurun_out/${TAG}_kb_dyn3.log 2>&1
paste <(grep "^gemm" gpurun_out/${TAG}_kb_dyn0.log | awk '{print $2, $9}') <(grep "^gemm" gpurun_out/${TAG}_kb_dyn3.log | awk '{print $9}')
