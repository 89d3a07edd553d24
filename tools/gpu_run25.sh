mkdir -p gpurun_out
T=r03a
export MFC_CONV_TWO=0 MFC_CONV_OSTAGE=0
MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 timeout 600 python tools/conv_diag.py fp16 2>&1 | tail -1 > gpurun_out/${T}_conv_diag.log
for tag in "" nosmr; do
for c in 22 12 13 21 16; do
MFC_B200_LIB_TAG=$tag MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 timeout 300 python tools/conv_bench.py $c --iters 30 2>&1 | cut -c1-160 | tail -1 >> gpurun_out/${T}_convbench_smr_${tag}.log
done
done
run() { name=$1; shift; env MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench_${name}.json 2> gpurun_out/${T}_bench_${name}.err; cp gpurun_out/bench_layers.json gpurun_out/${T}_layers_${name}.json; }
run smr
run nosmr MFC_B200_LIB_TAG=nosmr
