set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/parity_accurate.jsonl
python -m oracle.make_golden_corr > gpurun_out/r02a_golden_corr.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02a_pytest.log
MFC_SILU_ACCURATE=1 MFC_PARITY_REPORT=gpurun_out/parity_accurate.jsonl python -m pytest tests/test_gpu_models.py -q -k "matches_reference or full_size or benchmarked or realistic or remaining" 2>&1 | tail -15 > gpurun_out/r02a_pytest_accurate.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_model.json 2> gpurun_out/r02a_bench_model.err
cp gpurun_out/bench_layers.json gpurun_out/r02a_layers_model.json
MFC_SILU_ACCURATE=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_model_accurate.json 2> gpurun_out/r02a_bench_model_accurate.err
python tools/tune_table.py --fresh > gpurun_out/r02a_tune.log 2>&1
mkdir -p mfcnet-tracker_b200/tuning && cp gpurun_out/b200.tbl mfcnet-tracker_b200/tuning/b200.tbl
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_table.json 2> gpurun_out/r02a_bench_table.err
cp gpurun_out/bench_layers.json gpurun_out/r02a_layers_table.json
python tools/bench_corr.py > gpurun_out/r02a_bench_corr.log 2>&1
python tools/bench_stream.py --model resunet --k 3 --frames 3000 --clips 8 > gpurun_out/r02a_stream_resunet_b8.log 2>&1
python tools/bench_stream.py --model hrnet --k 5 --frames 2000 --clips 4 > gpurun_out/r02a_stream_hrnet_b4.log 2>&1
echo done
