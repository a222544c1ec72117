set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu_n2.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest_gpu_n2.log
timeout 900 python scripts/all_variants_smoke.py > gpurun_out/all_variants_n2.log 2>&1; echo "variants rc=$?"; grep -c "^2 " gpurun_out/all_variants_n2.log; tail -2 gpurun_out/all_variants_n2.log
