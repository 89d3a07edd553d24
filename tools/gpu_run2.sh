set -x
mkdir -p gpurun_out
T=r02b
rm -f gpurun_out/parity_report.jsonl
python tools/conv_diag.py fp16 > gpurun_out/${T}_conv_diag.log 2>&1
tail -3 gpurun_out/${T}_conv_diag.log
# A/B of this round's kernel / plan changes, each with live tuning (no table)
for v in new "nofast:MFC_CONV_EPI_FAST=0" "nores:MFC_RES_AS_SOURCE=0"; do
  name=${v%%:*}; envs=""; [ "$v" != "$name" ] && envs=${v#*:}
  env MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 $envs python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_${name}.json 2> gpurun_out/${T}_bench_${name}.err
  cp gpurun_out/bench_layers.json gpurun_out/${T}_layers_${name}.json
done
python tools/tune_table.py --fresh > gpurun_out/${T}_tune.log 2>&1
mkdir -p mfcnet-tracker_b200/tuning && cp gpurun_out/b200.tbl mfcnet-tracker_b200/tuning/b200.tbl
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cp gpurun_out/bench_layers.json gpurun_out/${T}_layers.json
python tools/bench_stream.py --model resunet --k 3 --frames 3000 --clips 8 > gpurun_out/${T}_stream_resunet_b8.log 2>&1
python tools/bench_stream.py --model resunet --k 3 --frames 1500 --clips 4 > gpurun_out/${T}_stream_resunet_b4.log 2>&1
python tools/bench_stream.py --model hrnet --k 5 --frames 2000 --clips 4 > gpurun_out/${T}_stream_hrnet_b4.log 2>&1
python tools/bench_stream.py --model hrnet --k 5 --frames 2000 --clips 8 > gpurun_out/${T}_stream_hrnet_b8.log 2>&1
# ncu: launch list of one bench step, then one full capture of the dominant layer shape (2nd conv launch of the 3rd step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-kernel-timing > gpurun_out/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 77 --launch-count 1 -f -o gpurun_out/${T}_conv16 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-kernel-timing > gpurun_out/${T}_ncu_full.log 2>&1
echo done
