set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
cp gpurun_out/parity_observed.json gpurun_out/r2_parity_observed_full.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --fp32 --no-cpu --no-solves --no-strong > gpurun_out/r2_bench_linear2_fp32.json 2>/dev/null; echo "fp32 rc=$?"
bash scripts/gpu_profile.sh r2
