set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lamba_em.py -m gpu -q --durations=3 > gpurun_out/pytest_lamba.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_lamba.log
timeout 600 python scripts/sweep.py gbm4_lamba 2e6 0:0:0,256:4:0,256:3:0,256:2:0,128:4:0,128:6:0,256:2:0:1,256:2:0:2 > gpurun_out/sweep_lamba.log 2>&1
cat gpurun_out/sweep_lamba.log
timeout 300 python scripts/nodes_time.py gbm4_lamba 2e6 0:0,256:2,256:3,256:4 > gpurun_out/nodes_lamba.log 2>&1; cat gpurun_out/nodes_lamba.log
timeout 600 python bench.py --workload gbm4_lamba_mc --no-solves > gpurun_out/bench_lamba.json 2> gpurun_out/bench_lamba.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_lamba.json
timeout 600 bash scripts/gpu_profile_wl.sh gbm4_lamba_mc lamba; tail -1 gpurun_out/ncu_lamba.log | cut -c1-200
