mkdir -p gpurun_out
out=gpurun_out/r05f_trace.log
: > $out
for c in 12 22; do
  echo "== case $c" >> $out
  env MFC_B200_LIB_TAG=trace MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 MFC_CONV_DEBUG=4096 timeout 300 python tools/conv_bench.py $c --iters 3 2>&1 | cut -c1-260 | tail -9 >> $out
done
