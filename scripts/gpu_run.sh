set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; echo "bench rc=$?"
python scripts/sweep.py linear2 1e7 256:4:0,256:3:0,256:2:0,128:8:0,128:6:0,256:4:32,256:3:32 > gpurun_out/sweep_v7_linear2.log 2>&1
python scripts/sweep.py lotka_volterra,lorenz63 4e6 256:4:0,256:3:0,256:2:0,128:6:0 > gpurun_out/sweep_v7_others.log 2>&1
cat gpurun_out/sweep_v7_linear2.log gpurun_out/sweep_v7_others.log
