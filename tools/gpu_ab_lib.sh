#!/bin/bash
# Same-box A/B of a measurement build of the library (MFC_B200_LIB_TAG) against the product build:
#   MFC_B200_LIB_TAG=<tag> MFC_B200_NVCC_DEFS="-D..." python mfcnet-tracker_b200/build.py ; gpurun -- 'bash tools/gpu_ab_lib.sh <tag> <out-prefix>'
TAG=$1; T=${2:-ab}
mkdir -p gpurun_out
for rep in 1 2; do
for tag in "" $TAG; do
  echo "== lib tag '$tag' rep $rep" >> gpurun_out/${T}_convbench.log
  MFC_B200_LIB_TAG=$tag timeout 300 python tools/conv_bench.py 12 13 15 16 17 21 --iters 30 2>&1 | cut -c1-120 >> gpurun_out/${T}_convbench.log
  echo "== lib tag '$tag' rep $rep" >> gpurun_out/${T}_bench.log
  MFC_B200_LIB_TAG=$tag timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary 2>/dev/null | cut -c1-200 >> gpurun_out/${T}_bench.log
done
done
MFC_B200_LIB_TAG=$TAG timeout 900 python -m pytest tests/test_gpu_models.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/${T}_pytest_models.log
echo done
