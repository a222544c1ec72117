set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_a.err
python scripts/sanitize_run.py > gpurun_out/r2_sanitize_plain.log 2>&1; echo "sanitize plain rc=$?"; tail -3 gpurun_out/r2_sanitize_plain.log
