set -x
mkdir -p gpurun_out
timeout 1200 python scripts/fuzz_parity.py 0 20 > gpurun_out/fuzz_parity.log 2>&1; echo "fuzz rc=$?"; grep -v '^{' gpurun_out/fuzz_parity.log | tail -30
timeout 900 python scripts/all_variants_smoke.py > gpurun_out/all_variants.log 2>&1; echo "variants rc=$?"; tail -4 gpurun_out/all_variants.log
