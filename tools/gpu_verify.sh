#!/bin/bash
# The round's verification recipe on a B200 box:  gpurun --timeout 2400 -- 'bash tools/gpu_verify.sh r02z'
# GPU test suite, smoke, the bench line (both arms), the ncu launch list and one full capture of the most frequent conv
# shape, the correlation bench.  Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
set -x
mkdir -p gpurun_out
T=${1:-verify}
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${T}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cp gpurun_out/bench_layers.json gpurun_out/${T}_layers.json
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-kernel-timing --no-secondary > gpurun_out/${T}_plain.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-kernel-timing --no-secondary > gpurun_out/${T}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 75 --launch-count 1 -f -o gpurun_out/${T}_conv16 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-kernel-timing --no-secondary > gpurun_out/${T}_ncu_full.log 2>&1
timeout 600 python tools/bench_corr.py > gpurun_out/${T}_bench_corr.log 2>&1
echo done
