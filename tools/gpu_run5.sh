set -x
mkdir -p gpurun_out
T=r02e
rm -f gpurun_out/parity_report.jsonl
python tools/conv_diag.py fp16 > gpurun_out/${T}_conv_diag.log 2>&1
tail -2 gpurun_out/${T}_conv_diag.log
python tools/tune_table.py > gpurun_out/${T}_tune.log 2>&1      # incremental: keeps the committed entries, adds new geometries
mkdir -p mfcnet-tracker_b200/tuning && cp gpurun_out/b200.tbl mfcnet-tracker_b200/tuning/b200.tbl
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
bash tools/role_ablation.sh ${T}_ablation
python tools/bench_stream.py --model hrnet --k 5 --frames 2000 --clips 8 > gpurun_out/${T}_stream_hrnet_b8.log 2>&1
python tools/bench_models.py > gpurun_out/${T}_bench_models.log 2>&1
echo done
