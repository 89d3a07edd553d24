mkdir -p gpurun_out
out=gpurun_out/r02v_ablation.log
: > $out
for v in "MFC_CONV_DEBUG=0" "MFC_CONV_DEBUG=256" "MFC_CONV_OSTAGE=0" "MFC_CONV_DEBUG=64" "MFC_CONV_DEBUG=5" "MFC_CONV_DEBUG=261"; do
  echo "== $v" >> $out
  env MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 $v timeout 300 python tools/conv_bench.py 22 12 --iters 30 2>&1 | cut -c1-180 | tail -2 >> $out
done
