set -x
mkdir -p gpurun_out
out=gpurun_out/r02s_ablation.log
: > $out
for dbg in 0 2 4 3 6 5 7 32 34 8; do
  echo "== MFC_CONV_DEBUG=$dbg" >> $out
  MFC_CONV_TWO=0 MFC_CONV_DEBUG=$dbg timeout 300 python tools/conv_bench.py 22 12 --iters 30 2>&1 | cut -c1-220 | grep -v "^mfc conv timing: warp 1[1-5]" | tail -12 >> $out
done
echo done
