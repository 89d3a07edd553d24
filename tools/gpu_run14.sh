set -x
mkdir -p gpurun_out
T=r02n
rm -f gpurun_out/parity_report.jsonl
python tools/tune_table.py --fresh > gpurun_out/${T}_tune.log 2>&1
mkdir -p mfcnet-tracker_b200/tuning && cp gpurun_out/b200.tbl mfcnet-tracker_b200/tuning/b200.tbl
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${T}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cp gpurun_out/bench_layers.json gpurun_out/${T}_layers.json
(cd _r01 && python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ../gpurun_out/${T}_bench_round1_code.json 2> ../gpurun_out/${T}_bench_round1_code.err)
python bench.py --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err
echo done
