mkdir -p gpurun_out
timeout 900 python tools/bench_stream.py --model resunet --k 3 --frames 1600 --clips 8 --online-flow > gpurun_out/r05h_stream_online.log 2>&1
timeout 900 python tools/bench_stream.py --model resunet --k 3 --frames 1600 --clips 8 > gpurun_out/r05h_stream.log 2>&1
timeout 900 python tools/bench_stream.py --model resunet --k 3 --frames 1600 --clips 16 --online-flow > gpurun_out/r05h_stream_online16.log 2>&1
