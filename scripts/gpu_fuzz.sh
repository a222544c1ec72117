set -x
mkdir -p gpurun_out
timeout 1500 python scripts/fuzz_parity.py 20 60 > gpurun_out/fuzz_ode2.log 2>&1; echo "ode rc=$?"; grep -v '^[{#]' gpurun_out/fuzz_ode2.log | tail -10
timeout 600 python scripts/fuzz_parity.py 14 30 noise > gpurun_out/fuzz_noise2.log 2>&1; echo "noise rc=$?"; grep -v '^[{#]' gpurun_out/fuzz_noise2.log | tail -10
