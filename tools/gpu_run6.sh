set -x
mkdir -p gpurun_out
T=r02f
rm -f gpurun_out/parity_report.jsonl
python tools/conv_diag.py fp16 > gpurun_out/${T}_conv_diag.log 2>&1
tail -2 gpurun_out/${T}_conv_diag.log
out=gpurun_out/${T}_wide.log; : > $out
for w in 1 0; do for dbg in 0 6; do
  echo "== MFC_CONV_TMA_WIDE=$w MFC_CONV_DEBUG=$dbg" >> $out
  MFC_CONV_TMA_WIDE=$w MFC_CONV_DEBUG=$dbg python tools/conv_bench.py 11 12 13 21 22 --iters 30 2>&1 | cut -c1-200 >> $out
done; done
python tools/tune_table.py --fresh > gpurun_out/${T}_tune.log 2>&1
mkdir -p mfcnet-tracker_b200/tuning && cp gpurun_out/b200.tbl mfcnet-tracker_b200/tuning/b200.tbl
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cp gpurun_out/bench_layers.json gpurun_out/${T}_layers.json
MFC_CONV_TMA_WIDE=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${T}_bench_nowide.json 2> gpurun_out/${T}_bench_nowide.err
echo done
