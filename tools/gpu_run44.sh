mkdir -p gpurun_out
T=r06f
MFC_CONV_TUNE=0 timeout 600 python tools/conv_diag.py bf16 2>&1 | tail -1 > gpurun_out/${T}_conv_diag_bf16.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${T}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cp gpurun_out/bench_layers.json gpurun_out/${T}_layers.json
