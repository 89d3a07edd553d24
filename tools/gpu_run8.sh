set -x
mkdir -p gpurun_out
T=r02h
rm -f gpurun_out/parity_report.jsonl
python tools/conv_diag.py fp16 > gpurun_out/${T}_conv_diag.log 2>&1
tail -1 gpurun_out/${T}_conv_diag.log
python tools/tune_table.py --fresh > gpurun_out/${T}_tune.log 2>&1
mkdir -p mfcnet-tracker_b200/tuning && cp gpurun_out/b200.tbl mfcnet-tracker_b200/tuning/b200.tbl
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cp gpurun_out/bench_layers.json gpurun_out/${T}_layers.json
python tools/bench_corr.py > gpurun_out/${T}_bench_corr.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-kernel-timing --no-secondary > gpurun_out/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 75 --launch-count 1 -f -o gpurun_out/${T}_conv16 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-kernel-timing --no-secondary > gpurun_out/${T}_ncu_full.log 2>&1
echo done
