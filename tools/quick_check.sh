#!/bin/bash
# gpurun helper: GEMM parity tests, per-kernel bench and a short step bench; logs in gpurun_out/<tag>_*.
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/kernel_bench.py --only gemm --cublas > gpurun_out/${TAG}_kb.log 2>&1
grep "^gemm" gpurun_out/${TAG}_kb.log | awk '{print $2, $9, $(NF-1)}'
for i in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_${i}.log 2>&1
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench_${i}.log | head -1
done
if [ -n "$FULL" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_full.log 2>&1
  tail -3 gpurun_out/${TAG}_pytest_full.log
fi
