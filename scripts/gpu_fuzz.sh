set -x
mkdir -p gpurun_out
timeout 900 python scripts/fuzz_parity.py 0 14 noise > gpurun_out/fuzz_noise.log 2>&1; echo "noise rc=$?"; grep -v '^{' gpurun_out/fuzz_noise.log | tail -20
timeout 1200 python scripts/fuzz_parity.py 0 5 large > gpurun_out/fuzz_large.log 2>&1; echo "large rc=$?"; grep -v '^{' gpurun_out/fuzz_large.log | tail -10
