set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lamba_em.py -m gpu -q > gpurun_out/pytest_lamba.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_lamba.log
timeout 600 python scripts/sweep.py gbm4_lamba 2e6 0:0:0,256:4:0,256:3:0,256:2:0,128:4:0,256:2:32,256:2:8,256:2:0:1,256:2:0:2 > gpurun_out/sweep_lamba.log 2>&1
cat gpurun_out/sweep_lamba.log
