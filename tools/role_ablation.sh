# Role ablation of conv_tc_kernel (MFC_CONV_DEBUG: bit0 no producer copies, bit1 no epilogue body, bit2 no MMAs): which role
# bounds a layer?  Tilings come from the committed table, so every run uses the same plan.
out=gpurun_out/${1:-ablation}.log
: > $out
for dbg in 0 1 2 4 3 6 5 7; do
  echo "== MFC_CONV_DEBUG=$dbg" >> $out
  MFC_CONV_DEBUG=$dbg python tools/conv_bench.py 11 12 13 21 22 --iters 30 2>&1 | cut -c1-200 >> $out
done
