set -x
mkdir -p gpurun_out
timeout 1200 python scripts/fuzz_parity.py 100 40 > gpurun_out/r2_fuzz_ode.log 2>&1; echo "ode rc=$?"; grep -v '^[{#]' gpurun_out/r2_fuzz_ode.log | tail -5
timeout 600 python scripts/fuzz_parity.py 100 20 noise > gpurun_out/r2_fuzz_noise.log 2>&1; echo "noise rc=$?"; grep -v '^[{#]' gpurun_out/r2_fuzz_noise.log | tail -5
timeout 600 python scripts/fuzz_parity.py 100 4 large > gpurun_out/r2_fuzz_large.log 2>&1; echo "large rc=$?"; grep -v '^[{#]' gpurun_out/r2_fuzz_large.log | tail -5
