#!/bin/bash
# gpurun helper: per-shape sweep of the GEMM tile configurations (RPO_GEMM_FORCE: diagnostics build), graph-timed
#   tools/gemm_sweep.sh TAG [kernel_bench.py arguments, e.g. --arch ViT-L/14 --prec bf16 --batch 16]
TAG=${1:-s}
shift
mkdir -p gpurun_out
RPO_DIAG=1 python -m rpo_b200.build --force > /dev/null 2>&1
CFGS=${CFGS:-"p256 p192 p128 s128 s64 l64"}
for c in $CFGS; do
  RPO_GEMM_FORCE=$c timeout 300 python tools/kernel_bench.py --only gemm "$@" > gpurun_out/${TAG}_kb_$c.log 2>&1
done
python - <<PY
import re
cfgs="$CFGS".split()
rows={}
for c in cfgs:
    for line in open(f"gpurun_out/${TAG}_kb_{c}.log"):
        if line.startswith("gemm"):
            p=line.split()
            rows.setdefault(p[1],{})[c]=p[5]  # microseconds
print("shape".ljust(12)+" ".join(c.rjust(7) for c in cfgs))
for k,v in rows.items():
    print(k.ljust(12)+" ".join(v.get(c,'-').rjust(7) for c in cfgs))
PY
