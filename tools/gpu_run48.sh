mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "corr" > gpurun_out/r06n_pytest_corr.log 2>&1
timeout 600 python tools/bench_corr.py > gpurun_out/r06n_bench_corr.log 2>&1
