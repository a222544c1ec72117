set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "koopman or regions or table or readme or lamba_process" > gpurun_out/r2_pytest_gpu_k.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_k.log
python scripts/koopman_time.py 0:0 > gpurun_out/r2_koopman_shapes_c.log 2>&1; cat gpurun_out/r2_koopman_shapes_c.log
