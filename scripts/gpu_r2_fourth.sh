set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_d.err
for wl in lorenz63_mc lotka_volterra_mc oscillator_noise_mc gbm4_lamba_mc; do
  python bench.py --workload $wl --no-cpu --no-solves --no-strong > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; echo "bench $wl rc=$?"
done
python bench.py --fp32 --no-cpu --no-solves --no-strong > gpurun_out/r2_bench_linear2_fp32.json 2>/dev/null; echo "fp32 rc=$?"
