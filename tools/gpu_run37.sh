mkdir -p gpurun_out
T=r05g
MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 timeout 600 python tools/conv_diag.py fp16 2>&1 | tail -1 > gpurun_out/${T}_conv_diag.log
for c in 12 13 16 17 21; do
timeout 300 python tools/conv_bench.py $c --iters 30 2>&1 | cut -c1-160 | tail -1 >> gpurun_out/${T}_convbench.log
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
