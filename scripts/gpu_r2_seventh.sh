set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest_gpu.log
cp gpurun_out/parity_observed.json gpurun_out/r2_parity_observed_full.json
python scripts/nodes_latency.py > gpurun_out/r2_nodes_latency_b.log 2>&1; cat gpurun_out/r2_nodes_latency_b.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
