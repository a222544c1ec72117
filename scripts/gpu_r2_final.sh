set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
cp gpurun_out/parity_observed.json gpurun_out/r2_parity_observed_full.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "ref rc=$?"
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_final.err
python bench.py --workload oscillator_noise_mc --no-cpu --no-solves --no-strong > gpurun_out/r2_bench_oscillator_noise_mc.json 2>/dev/null; echo "bench em rc=$?"
