mkdir -p gpurun_out
MFC_CONV_TUNE=0 timeout 600 python tools/conv_diag.py fp16 2>&1 | tail -7 > gpurun_out/r04d_conv_diag.log
timeout 900 python -m pytest tests/test_gpu_raft.py -x -q -m gpu > gpurun_out/r04d_pytest_raft.log 2>&1
timeout 600 python tools/bench_raft.py 240 320 1 2 8 > gpurun_out/r04d_bench_raft.log 2>&1
timeout 600 python tools/raft_layers.py 240 320 2 > gpurun_out/r04d_raft_layers.log 2>&1
