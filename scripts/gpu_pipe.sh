set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "split_reduce or reproducible_dynamic or static_schedule or layouts" > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_pipe.log
timeout 900 python bench.py > gpurun_out/bench_pipe.json 2> gpurun_out/bench_pipe.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_pipe.err; cut -c1-260 gpurun_out/bench_pipe.json
timeout 600 python bench.py --workload lorenz63_mc --no-cpu --no-solves --steps 10 > gpurun_out/bench_pipe_lorenz.json 2>&1; cut -c1-200 gpurun_out/bench_pipe_lorenz.json
