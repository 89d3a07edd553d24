mkdir -p gpurun_out
out=gpurun_out/r02y_trace.log
: > $out
for v in 4096 4103 4101 4099; do
for c in 22 12; do
  echo "== debug $v case $c" >> $out
  env MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 MFC_CONV_OSTAGE=0 MFC_CONV_DEBUG=$v timeout 300 python tools/conv_bench.py $c --iters 3 2>&1 | cut -c1-260 | tail -9 >> $out
done
done
