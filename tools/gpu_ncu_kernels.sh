#!/bin/bash
# ncu --set full of the tensor-bound conv shapes and of the memory-bound helper kernels of the bench step; the reports are reduced
# to their raw-page CSV on the box (gpurun brings back at most 64 MiB).  Summarised by tools/ncu_summary.py into profiles/.
mkdir -p gpurun_out
T=${1:-ncuk}
timeout 300 python tools/conv_bench.py 23 24 15 6 --iters 20 > gpurun_out/${T}_convbench.log 2>&1
cap() {  # name, kernel regex, launch-skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none -k regex:$rx --launch-skip $skip -c 1 -f -o /tmp/${T}_$name "$@" > gpurun_out/${T}_ncu_$name.log 2>&1
  ncu -i /tmp/${T}_$name.ncu-rep --page raw --csv > gpurun_out/${T}_$name.raw.csv 2>/dev/null
  rm -f /tmp/${T}_$name.ncu-rep
}
for c in 23 24 15; do cap conv_case$c conv_tc_kernel 3 python tools/conv_bench.py $c --iters 3; done
BA="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-kernel-timing --no-secondary"
cap asa affine_silu_add 10 python bench.py $BA
cap gather gather 8 python bench.py $BA
( time timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err ) 2> gpurun_out/${T}_bench_default.time
echo done
