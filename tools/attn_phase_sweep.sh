#!/bin/bash
# gpurun helper: tcgen05 attention forward (vision shape) under the phase-delay / packed-f32x2 switches, then the
# attention parity tests with the chosen setting ($1 = delay for the parity run, $2 = f32x2 flag)
mkdir -p gpurun_out
for f in 0 1; do
  for d in 0 1500 3000 4500 6000 8000; do
    echo "f32x2=$f delay=$d: $(RPO_ATTN_F32X2=$f RPO_ATTN_PHASE_DELAY=$d timeout 120 python tools/kernel_bench.py --only attn 2>&1 | grep 'v.attn_fwd')"
  done
done
RPO_ATTN_F32X2=${2:-1} RPO_ATTN_PHASE_DELAY=${1:-3000} timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "dense" 2>&1 | tail -2
