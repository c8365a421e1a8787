#!/bin/bash
# Closing evidence of a round on one B200 (run through gpurun), trimmed to a few GPU-minutes: the full bench line of the
# headline configuration, quick lines of configs 3 and 4, the launch list of a step, and `ncu --set full` captures of
# the kernels changed last (LayerNorm forward / backward, attention backward).  Lands in gpurun_out/$TAG/.
TAG=${1:-r02final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
NCU="ncu --set full --clock-control none"
KB="python tools/kernel_bench.py --no-graph"
cap() {
  local name=$1 rx=$2; shift 2
  timeout 200 $NCU -k regex:$rx -s 3 -c 1 -f -o $OUT/$name "$@" > $OUT/$name.log 2>&1
  python tools/ncu_summary.py $OUT/$name.ncu-rep > $OUT/ncu_$name.txt 2>&1
  rm -f $OUT/$name.ncu-rep
}
timeout 420 python bench.py > $OUT/bench_config2.log 2>&1; echo "bench2 rc=$?"
timeout 150 python bench.py --config 4 --quick --no-cpu-baseline > $OUT/bench_config4_quick.log 2>&1; echo "bench4 rc=$?"
timeout 150 python bench.py --config 3 --quick --no-cpu-baseline > $OUT/bench_config3_quick.log 2>&1; echo "bench3 rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
  python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline > $OUT/launches_bench.log 2>&1; echo "launches rc=$?"
cap ln_fwd ln_fwd_kernel $KB --only ln
cap ln_bwd ln_bwd_kernel $KB --only ln
cap attn_bwd_ks ro_attn_bwd_ks $KB --only attn
timeout 120 python tools/kernel_bench.py --cublas > $OUT/kernel_bench.txt 2>&1; echo "kb rc=$?"
ls -la $OUT | tail -20
