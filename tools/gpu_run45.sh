mkdir -p gpurun_out
T=r06i
MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 timeout 600 python tools/conv_diag.py fp16 2>&1 | tail -1 > gpurun_out/${T}_conv_diag.log
for c in 22 1 2 12; do
timeout 300 python tools/conv_bench.py $c --iters 30 2>&1 | cut -c1-160 | tail -1 >> gpurun_out/${T}_convbench.log
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 900 python -m pytest tests/test_gpu_models.py -x -q -m gpu > gpurun_out/${T}_pytest_models.log 2>&1
