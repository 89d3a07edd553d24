set -x
mkdir -p gpurun_out
T=r02j
python tools/bench_models.py > gpurun_out/${T}_bench_models.log 2>&1
for b in 8 16; do python tools/bench_stream.py --model hrnet --k 5 --frames 3000 --clips $b > gpurun_out/${T}_stream_hrnet_b$b.log 2>&1; done
python tools/bench_stream.py --model resunet --k 3 --frames 6000 --clips 16 > gpurun_out/${T}_stream_resunet_b16.log 2>&1
python tools/bench_keypoints.py > gpurun_out/${T}_bench_keypoints.log 2>&1
python tools/bench_tracking.py > gpurun_out/${T}_bench_tracking.log 2>&1
python tools/bench_video.py > gpurun_out/${T}_bench_video.log 2>&1
python tools/bench_train.py > gpurun_out/${T}_bench_train.log 2>&1
echo done
