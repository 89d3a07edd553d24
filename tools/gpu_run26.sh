set -x
mkdir -p gpurun_out
T=r05e
rm -f gpurun_out/parity_report.jsonl
MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 timeout 600 python tools/conv_diag.py fp16 2>&1 | tail -1 > gpurun_out/${T}_conv_diag.log
timeout 1500 python tools/tune_table.py --fresh > gpurun_out/${T}_tune.log 2>&1
mkdir -p mfcnet-tracker_b200/tuning && cp gpurun_out/b200.tbl mfcnet-tracker_b200/tuning/b200.tbl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${T}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cp gpurun_out/bench_layers.json gpurun_out/${T}_layers.json
(cd _r01 && timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ../gpurun_out/${T}_bench_round1_code.json 2> ../gpurun_out/${T}_bench_round1_code.err)
MFC_CONV_SETMAXNREG=0 timeout 600 python bench.py --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/${T}_bench_nosmr.json 2> gpurun_out/${T}_bench_nosmr.err
echo done
