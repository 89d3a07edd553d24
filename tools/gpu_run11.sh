set -x
mkdir -p gpurun_out
T=r02k
rm -f gpurun_out/parity_report.jsonl
# same-box A/B against the round-1 code (scratch copy of commit 381fe82 under _r01/, live autotuning as round 1 did)
(cd _r01 && python bench.py --steps 20 --warmup 5 --no-cpu-baseline > ../gpurun_out/${T}_bench_round1_code.json 2> ../gpurun_out/${T}_bench_round1_code.err)
python tools/tune_table.py --fresh > gpurun_out/${T}_tune.log 2>&1
mkdir -p mfcnet-tracker_b200/tuning && cp gpurun_out/b200.tbl mfcnet-tracker_b200/tuning/b200.tbl
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${T}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cp gpurun_out/bench_layers.json gpurun_out/${T}_layers.json
python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 75 --launch-count 1 -f -o gpurun_out/${T}_conv16 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-kernel-timing --no-secondary > gpurun_out/${T}_ncu_full.log 2>&1
echo done
