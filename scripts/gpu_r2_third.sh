set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest_gpu.log
python bench.py --no-cpu > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_c.err
B200E_COPY_THREADS=4 python bench.py --no-cpu --no-solves --no-strong > gpurun_out/r2_bench_c_copy4.json 2>/dev/null
B200E_COPY_THREADS=12 python bench.py --no-cpu --no-solves --no-strong > gpurun_out/r2_bench_c_copy12.json 2>/dev/null
