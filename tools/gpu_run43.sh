mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raft.py -x -q -m gpu > gpurun_out/r06e_pytest_raft.log 2>&1
timeout 600 python tools/bench_raft.py 240 320 1 2 8 > gpurun_out/r06e_bench_raft.log 2>&1
MFC_RAFT_LOOKUP_OLD=1 timeout 600 python tools/bench_raft.py 240 320 1 2 8 > gpurun_out/r06e_bench_raft_oldlookup.log 2>&1
timeout 900 python tools/bench_stream.py --model resunet --k 3 --frames 1600 --clips 8 --online-flow > gpurun_out/r06e_stream_online.log 2>&1
timeout 600 python tools/raft_layers.py 240 320 16 2>&1 | grep "raft2\|==" > gpurun_out/r06e_layers.log
