set -x
mkdir -p gpurun_out
T=r02c
rm -f gpurun_out/parity_report.jsonl
python tools/conv_diag.py fp16 > gpurun_out/${T}_conv_diag.log 2>&1
tail -2 gpurun_out/${T}_conv_diag.log
for v in new "noflat:MFC_CONV_FLAT=0" "nochunk:MFC_CONV_CHUNK=1" "chunk8:MFC_CONV_CHUNK=8" "nores:MFC_RES_AS_SOURCE=0"; do
  name=${v%%:*}; envs=""; [ "$v" != "$name" ] && envs=${v#*:}
  env MFC_CONV_TUNE=1 MFC_CONV_TABLE=0 $envs python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench_${name}.json 2> gpurun_out/${T}_bench_${name}.err
  cp gpurun_out/bench_layers.json gpurun_out/${T}_layers_${name}.json
done
python tools/tune_table.py --fresh > gpurun_out/${T}_tune.log 2>&1
mkdir -p mfcnet-tracker_b200/tuning && cp gpurun_out/b200.tbl mfcnet-tracker_b200/tuning/b200.tbl
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cp gpurun_out/bench_layers.json gpurun_out/${T}_layers.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
echo done
