set -x
mkdir -p gpurun_out
T=r02g
out=gpurun_out/${T}_timing.log; : > $out
for dbg in 8 15 14 13 11; do
  echo "== MFC_CONV_DEBUG=$dbg" >> $out
  MFC_CONV_DEBUG=$dbg python tools/conv_bench.py 22 12 21 --iters 3 2>&1 | cut -c1-260 >> $out
done
MFC_CONV_SLIDE=1 MFC_CONV_TUNE=0 python tools/conv_diag.py fp16 > gpurun_out/${T}_conv_diag_slide.log 2>&1
python -m pytest tests/test_gpu_models.py -q -k "hrnet" 2>&1 | tail -5 > gpurun_out/${T}_pytest_hrnet.log
echo done
