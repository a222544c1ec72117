set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_gpu.log
python scripts/koopman_time.py 0:0,128:4 > gpurun_out/r2_koopman_shapes_b.log 2>&1; cat gpurun_out/r2_koopman_shapes_b.log
python bench.py > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_e.err
