set -x
mkdir -p gpurun_out
T=r02q
# lite (two CTAs per SM) plans: correctness with the cost model's choice (prefers lite), then timing
MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 timeout 600 python tools/conv_diag.py fp16 > gpurun_out/${T}_conv_diag_lite.log 2>&1
MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 timeout 600 python tools/conv_diag.py fp16 2>&1 | tail -1 > gpurun_out/${T}_conv_diag_nolite.log
for c in 1 2 3 12 13; do
MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 timeout 300 python tools/conv_bench.py $c >> gpurun_out/${T}_convbench_lite.log 2>&1
MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 timeout 300 python tools/conv_bench.py $c >> gpurun_out/${T}_convbench_nolite.log 2>&1
MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 timeout 300 python tools/conv_bench.py $c >> gpurun_out/${T}_convbench_tuned.log 2>&1
MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 timeout 300 python tools/conv_bench.py $c >> gpurun_out/${T}_convbench_tuned_nolite.log 2>&1
done
run() { name=$1; shift; env MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench_${name}.json 2> gpurun_out/${T}_bench_${name}.err; cp gpurun_out/bench_layers.json gpurun_out/${T}_layers_${name}.json; }
run lite
run nolite MFC_CONV_TWO=0
echo done
