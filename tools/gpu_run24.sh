mkdir -p gpurun_out
T=r02z
export MFC_CONV_TUNE=0 MFC_CONV_TABLE=0 MFC_CONV_TWO=0 MFC_CONV_OSTAGE=0
timeout 300 python tools/conv_bench.py 12 --iters 5 > gpurun_out/${T}_plain12.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/${T}_aff12 python tools/conv_bench.py 12 --iters 5 > gpurun_out/${T}_ncu12.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/${T}_stats22 python tools/conv_bench.py 22 --iters 5 > gpurun_out/${T}_ncu22.log 2>&1
