set -x
mkdir -p gpurun_out
T=r02l
run() { name=$1; shift; env MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench_${name}.json 2> gpurun_out/${T}_bench_${name}.err; cp gpurun_out/bench_layers.json gpurun_out/${T}_layers_${name}.json; }
run poll
run hint MFC_B200_LIB_TAG=hint
run noguard MFC_B200_LIB_TAG=noguard
run r1like MFC_RES_AS_SOURCE=0 MFC_CONV_HEAD=0 MFC_CONV_FLAT=0 MFC_CONV_CHUNK=1 MFC_CONV_EPI_FAST=0
run r1like_nowide MFC_RES_AS_SOURCE=0 MFC_CONV_HEAD=0 MFC_CONV_FLAT=0 MFC_CONV_CHUNK=1 MFC_CONV_EPI_FAST=0 MFC_CONV_TMA_WIDE=0
run poll2
(cd _r01 && python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ../gpurun_out/${T}_bench_round1_code.json 2> ../gpurun_out/${T}_bench_round1_code.err; cp gpurun_out/bench_layers.json ../gpurun_out/${T}_layers_round1_code.json)
echo done
