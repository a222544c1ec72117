set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_v18.json 2> gpurun_out/bench_v18.err; echo "bench rc=$?"
python scripts/sweep.py linear2 1e7 256:4:0,256:3:0,128:8:0,256:4:32,256:4:0:1,256:3:0:1 > gpurun_out/sweep_v18_linear2.log 2>&1
python scripts/sweep.py lotka_volterra,lorenz63 4e6 256:4:0,256:3:0,256:4:32,256:4:8,256:4:0:1 > gpurun_out/sweep_v18_others.log 2>&1
cat gpurun_out/sweep_v18_linear2.log gpurun_out/sweep_v18_others.log
