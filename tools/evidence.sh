#!/bin/bash
# Round evidence on one B200 (run through gpurun): launch list of the bench step, one `ncu --set full` capture per
# kernel family of the step, bench lines of BASELINE configs 3 / 4 / 5.  Everything lands in gpurun_out/$TAG/;
# tools/ncu_summary.py and tools/launch_summary.py turn it into the text files under profiles/.
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
NCU="ncu --set full --clock-control none"
KB="python tools/kernel_bench.py --no-graph"
cap() {  # name, kernel regex, command...: the text summary always comes back, the raw report only for the attention kernels
  local name=$1 rx=$2; shift 2
  local src=""; case $name in attn*) src="--import-source on";; esac
  timeout 240 $NCU $src -k regex:$rx -s 3 -c 1 -f -o $OUT/$name "$@" > $OUT/$name.log 2>&1
  python tools/ncu_summary.py $OUT/$name.ncu-rep > $OUT/ncu_$name.txt 2>&1
  case $name in attn*) ;; *) rm -f $OUT/$name.ncu-rep;; esac
}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
  python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
cap attn_fwd_pp ro_attn_fwd_pp $KB --only attn
cap attn_bwd_ks ro_attn_bwd_ks $KB --only attn
cap attn_fwd_vitl ro_attn_fwd_tc $KB --only attn --arch ViT-L/14 --prec bf16 --batch 16
cap gemm_small_tq gemm_tc_kernel $KB --only gemm --gemms t.q
cap gemm_small_vbsq gemm_tc_kernel $KB --only gemm --gemms vb.sq
cap gemm_small_tbdpre gemm_ $KB --only gemm --gemms tb.dpre
cap gemm_vbdh gemm_ $KB --only gemm --gemms vb.dh
cap gemm_vfc gemm_tc2_kernel $KB --only gemm --gemms v.fc
cap gemm_vout gemm_ $KB --only gemm --gemms v.out
cap gemm_vproj gemm_ $KB --only gemm --gemms v.proj
cap ln_fwd ln_fwd_kernel $KB --only ln
cap ln_bwd ln_bwd_kernel $KB --only ln
if [ -z "$SKIP_BENCH" ]; then  # SKIP_BENCH=1: captures only
  for c in 3 4 5; do
    timeout 600 python bench.py --config $c --no-cpu-baseline > $OUT/bench_config$c.log 2>&1
  done
fi
timeout 120 python tools/kernel_bench.py --cublas > $OUT/kernel_bench.txt 2>&1
timeout 120 python tools/kernel_bench.py --arch ViT-L/14 --prec bf16 --batch 16 --cublas > $OUT/kernel_bench_vitl.txt 2>&1
ls -la $OUT | tail -40
