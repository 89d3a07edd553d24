mkdir -p gpurun_out
T=r03d
for v in "MFC_X=1" "MFC_CONV_SETMAXNREG=0" "MFC_CONV_DEBUG=16384" "MFC_CONV_TABLE=0"; do
echo "== $v" >> gpurun_out/${T}_ab.log
env $v timeout 300 python tools/conv_bench.py 22 1 --iters 30 2>&1 | cut -c1-200 | tail -2 >> gpurun_out/${T}_ab.log
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/${T}_launches.csv python tools/conv_bench.py 22 --iters 3 > /dev/null 2>&1
