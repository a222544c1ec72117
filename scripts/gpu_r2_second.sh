set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_b.err
B200E_COPY_THREADS=4 python bench.py --no-cpu --no-solves --no-strong > gpurun_out/r2_bench_b_copy4.json 2>/dev/null
B200E_COPY_THREADS=16 python bench.py --no-cpu --no-solves --no-strong > gpurun_out/r2_bench_b_copy16.json 2>/dev/null
python scripts/sanitize_run.py > gpurun_out/r2_sanitize_plain.log 2>&1 && \
timeout 1200 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck.log python scripts/sanitize_run.py > gpurun_out/r2_sanitize_memcheck_stdout.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r2_sanitizer_memcheck.log; tail -3 gpurun_out/r2_sanitize_memcheck_stdout.log
