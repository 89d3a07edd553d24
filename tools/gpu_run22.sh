mkdir -p gpurun_out
out=gpurun_out/r02x_ablation.log
: > $out
env MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 timeout 600 python tools/conv_diag.py fp16 2>&1 | tail -1 >> $out
for v in 0 256 64 576 1088 2112 3648 5; do
  echo "== $v" >> $out
  env MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 MFC_CONV_DEBUG=$v timeout 300 python tools/conv_bench.py 22 12 --iters 30 2>&1 | cut -c1-180 | tail -2 >> $out
done
echo "== ostage off" >> $out
env MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 MFC_CONV_OSTAGE=0 timeout 300 python tools/conv_bench.py 22 12 --iters 30 2>&1 | cut -c1-180 | tail -2 >> $out
