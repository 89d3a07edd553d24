mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raft.py -x -q -m gpu > gpurun_out/r04b_pytest_raft.log 2>&1
timeout 600 python tools/bench_raft.py 240 320 1 2 8 > gpurun_out/r04b_bench_raft.log 2>&1
timeout 600 python tools/bench_raft.py 480 640 1 2 > gpurun_out/r04b_bench_raft_full.log 2>&1
