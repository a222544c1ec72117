set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -10 gpurun_out/pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_final.json
timeout 600 python bench.py --workload gbm4_lamba_mc --no-solves > gpurun_out/bench_lamba.json 2> gpurun_out/bench_lamba.err; echo "bench lamba rc=$?"; cut -c1-200 gpurun_out/bench_lamba.json
timeout 300 python scripts/lamba_koopman.py > gpurun_out/lamba_koopman.json 2>&1; cat gpurun_out/lamba_koopman.json
timeout 600 bash scripts/gpu_profile_wl.sh gbm4_lamba_mc lamba; tail -1 gpurun_out/ncu_lamba.log | cut -c1-120
