mkdir -p gpurun_out
for n in 2 3 4; do
MFC_CORR_NST=$n timeout 300 python tools/bench_corr.py 2>&1 | head -2 | cut -c1-200 >> gpurun_out/r06m_corr_nst.log
done
