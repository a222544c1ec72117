set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lamba_em.py -m gpu -q --durations=5 > gpurun_out/pytest_lamba.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_lamba.log
timeout 600 python bench.py --workload gbm4_lamba_mc --no-solves > gpurun_out/bench_lamba.json 2> gpurun_out/bench_lamba.err; echo "bench rc=$?"; cat gpurun_out/bench_lamba.json
timeout 600 bash scripts/gpu_profile_wl.sh gbm4_lamba_mc lamba; tail -3 gpurun_out/ncu_lamba.log
