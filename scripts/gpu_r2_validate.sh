set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
cp gpurun_out/parity_observed.json gpurun_out/r2_parity_observed_full.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
bash scripts/gpu_fuzz.sh
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "ref rc=$?"
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
for wl in lorenz63_mc lotka_volterra_mc oscillator_noise_mc gbm4_lamba_mc; do
  python bench.py --workload $wl --no-cpu --no-solves --no-strong > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; echo "bench $wl rc=$?"
done
python bench.py --fp32 --no-cpu --no-solves --no-strong > gpurun_out/r2_bench_linear2_fp32.json 2>/dev/null; echo "fp32 rc=$?"
bash scripts/gpu_profile.sh r2
