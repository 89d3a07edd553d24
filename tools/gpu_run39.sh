mkdir -p gpurun_out
T=r05j
for tag in "" u6 u8; do
for c in 12 13 21; do
MFC_B200_LIB_TAG=$tag timeout 300 python tools/conv_bench.py $c --iters 30 2>&1 | cut -c1-160 | tail -1 >> gpurun_out/${T}_convbench_${tag}.log
done
done
