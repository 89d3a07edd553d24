set -x
mkdir -p gpurun_out
T=r02r
MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 timeout 600 python tools/conv_diag.py fp16 > gpurun_out/${T}_conv_diag.log 2>&1
for c in 1 2 3 11 12 13 16 21; do
MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 timeout 300 python tools/conv_bench.py $c >> gpurun_out/${T}_convbench_new.log 2>&1
MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 MFC_CONV_DEBUG=32 timeout 300 python tools/conv_bench.py $c >> gpurun_out/${T}_convbench_oldepi.log 2>&1
done
run() { name=$1; shift; env MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench_${name}.json 2> gpurun_out/${T}_bench_${name}.err; cp gpurun_out/bench_layers.json gpurun_out/${T}_layers_${name}.json; }
run new
run oldepi MFC_CONV_DEBUG=32
timeout 900 python -m pytest tests/test_gpu_models.py -x -q -m gpu > gpurun_out/${T}_pytest_models.log 2>&1
echo done
