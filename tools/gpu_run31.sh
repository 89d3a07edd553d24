mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ingest.py tests/test_gpu_raft.py -x -q -m gpu > gpurun_out/r04f_pytest.log 2>&1
