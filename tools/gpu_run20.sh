mkdir -p gpurun_out
T=r02u
MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 timeout 600 python tools/conv_diag.py fp16 > gpurun_out/${T}_conv_diag.log 2>&1
for c in 1 2 3 22 12 13; do
MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 timeout 300 python tools/conv_bench.py $c 2>&1 | cut -c1-200 >> gpurun_out/${T}_convbench_ostage.log
MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 MFC_CONV_OSTAGE=0 timeout 300 python tools/conv_bench.py $c 2>&1 | cut -c1-200 >> gpurun_out/${T}_convbench_noostage.log
done
run() { name=$1; shift; env MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench_${name}.json 2> gpurun_out/${T}_bench_${name}.err; cp gpurun_out/bench_layers.json gpurun_out/${T}_layers_${name}.json; }
run ostage
run noostage MFC_CONV_OSTAGE=0
