mkdir -p gpurun_out
T=r05d
MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 timeout 600 python tools/conv_diag.py fp16 2>&1 | tail -1 > gpurun_out/${T}_conv_diag.log
for v in 0 32768; do
for c in 21 14 4 10; do
MFC_CONV_DEBUG=$v timeout 300 python tools/conv_bench.py $c --iters 30 2>&1 | cut -c1-160 | tail -1 >> gpurun_out/${T}_convbench_$v.log
done
done
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench_${name}.json 2> gpurun_out/${T}_bench_${name}.err; cp gpurun_out/bench_layers.json gpurun_out/${T}_layers_${name}.json; }
run light MFC_X=1
run nolight MFC_CONV_DEBUG=32768
