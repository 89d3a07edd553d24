mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q -m gpu -k "resample or hrnet or HRNet" > gpurun_out/r06k_pytest.log 2>&1
timeout 600 python tools/layer_table.py hrnet 16 2>&1 | head -8 > gpurun_out/r06k_hrnet16.log
timeout 900 python tools/bench_stream.py --model hrnet --k 5 --frames 3000 --clips 16 > gpurun_out/r06k_cfg4.log 2>&1
