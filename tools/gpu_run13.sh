set -x
mkdir -p gpurun_out
T=r02m
run() { name=$1; shift; env MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench_${name}.json 2> gpurun_out/${T}_bench_${name}.err; cp gpurun_out/bench_layers.json gpurun_out/${T}_layers_${name}.json; }
run default
run incwalk MFC_B200_LIB_TAG=incwalk
run nochunk MFC_CONV_CHUNK=1
run r1like_wide MFC_RES_AS_SOURCE=0 MFC_CONV_HEAD=0 MFC_CONV_FLAT=0 MFC_CONV_CHUNK=1 MFC_CONV_EPI_FAST=0
(cd _r01 && python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ../gpurun_out/${T}_bench_round1_code.json 2> ../gpurun_out/${T}_bench_round1_code.err)
python tools/conv_diag.py fp16 2>&1 | tail -1 > gpurun_out/${T}_conv_diag.log
echo done
