mkdir -p gpurun_out
out=gpurun_out/r02t_ablation.log
: > $out
for dbg in 0 64 128 192 5 69 133 197; do
  echo "== MFC_CONV_DEBUG=$dbg" >> $out
  MFC_CONV_TWO=0 MFC_CONV_DEBUG=$dbg timeout 300 python tools/conv_bench.py 22 12 --iters 30 2>&1 | cut -c1-120 | tail -2 >> $out
done
