set -x
mkdir -p gpurun_out
T=r02i
python -m pytest tests/test_gpu_models.py -q -x -k "unflow" 2>&1 | tail -25 > gpurun_out/${T}_pytest_unflow.log
python -m pytest tests/test_gpu_kernels.py -q -k "correlation" 2>&1 | tail -8 > gpurun_out/${T}_pytest_corr.log
python -m pytest tests/test_gpu_models.py -q -k "conv_formulations" 2>&1 | tail -8 > gpurun_out/${T}_pytest_form.log
echo done
