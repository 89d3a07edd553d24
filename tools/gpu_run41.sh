mkdir -p gpurun_out
timeout 1500 python tools/tune_table.py --only raft > gpurun_out/r06c_tune_raft.log 2>&1
cp gpurun_out/b200.tbl mfcnet-tracker_b200/tuning/b200.tbl
timeout 900 python -m pytest tests/test_gpu_raft.py -x -q -m gpu > gpurun_out/r06c_pytest_raft.log 2>&1
timeout 600 python tools/bench_raft.py 240 320 1 2 8 > gpurun_out/r06c_bench_raft.log 2>&1
timeout 900 python tools/bench_stream.py --model resunet --k 3 --frames 1600 --clips 8 --online-flow > gpurun_out/r06c_stream_online.log 2>&1
timeout 900 python tools/bench_stream.py --model resunet --k 3 --frames 1600 --clips 16 --online-flow > gpurun_out/r06c_stream_online16.log 2>&1
timeout 600 python tools/bench_video.py --frames 300 --online-flow > gpurun_out/r06c_video_online.log 2>&1
