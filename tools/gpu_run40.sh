mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raft.py -x -q -m gpu > gpurun_out/r06b_pytest_raft.log 2>&1
timeout 900 python tools/bench_stream.py --model resunet --k 3 --frames 1600 --clips 8 --online-flow > gpurun_out/r06b_stream_online.log 2>&1
MFC_FLOW_REUSE=0 timeout 900 python tools/bench_stream.py --model resunet --k 3 --frames 1600 --clips 8 --online-flow > gpurun_out/r06b_stream_online_noreuse.log 2>&1
timeout 900 python tools/bench_stream.py --model resunet --k 3 --frames 1600 --clips 16 --online-flow > gpurun_out/r06b_stream_online16.log 2>&1
