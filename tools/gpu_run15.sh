set -x
mkdir -p gpurun_out
T=r02o
python tools/conv_diag.py fp16 2>&1 | tail -1 > gpurun_out/${T}_conv_diag.log
run() { name=$1; shift; env MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench_${name}.json 2> gpurun_out/${T}_bench_${name}.err; cp gpurun_out/bench_layers.json gpurun_out/${T}_layers_${name}.json; }
run direct
run nodirect MFC_CONV_DIRECT=0
run direct2
(cd _r01 && python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ../gpurun_out/${T}_bench_round1_code.json 2> ../gpurun_out/${T}_bench_round1_code.err)
echo done
