mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raft.py -x -q -m gpu > gpurun_out/r04g_pytest.log 2>&1
timeout 600 python tools/bench_raft.py 240 320 1 2 8 > gpurun_out/r04g_bench_raft.log 2>&1
MFC_LANES=0 timeout 600 python tools/bench_raft.py 240 320 1 2 8 > gpurun_out/r04g_bench_raft_nolanes.log 2>&1
