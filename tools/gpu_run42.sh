mkdir -p gpurun_out
timeout 600 python tools/bench_raft.py 240 320 1 2 8 > gpurun_out/r06d_bench_raft.log 2>&1
MFC_RAFT_LANES=1 timeout 600 python tools/bench_raft.py 240 320 1 2 8 > gpurun_out/r06d_bench_raft_lanes.log 2>&1
