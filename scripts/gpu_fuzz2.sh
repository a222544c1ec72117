# second differential run of the final round-2 kernels: NEW seeds (500...) for the three families of gpu_fuzz.sh plus the
# log-polynomial density factors (tests/fuzz_util.py with_logpoly_factors)
set -x
mkdir -p gpurun_out
timeout 900 python scripts/fuzz_parity.py 500 30 > gpurun_out/r2_fuzz2_ode.log 2>&1; echo "ode rc=$?"; grep -v '^[{#]' gpurun_out/r2_fuzz2_ode.log | tail -3
timeout 900 python scripts/fuzz_parity.py 600 40 logpoly > gpurun_out/r2_fuzz2_logpoly.log 2>&1; echo "logpoly rc=$?"; grep -v '^[{#]' gpurun_out/r2_fuzz2_logpoly.log | tail -3
timeout 600 python scripts/fuzz_parity.py 500 16 noise > gpurun_out/r2_fuzz2_noise.log 2>&1; echo "noise rc=$?"; grep -v '^[{#]' gpurun_out/r2_fuzz2_noise.log | tail -3
timeout 600 python scripts/fuzz_parity.py 500 3 large > gpurun_out/r2_fuzz2_large.log 2>&1; echo "large rc=$?"; grep -v '^[{#]' gpurun_out/r2_fuzz2_large.log | tail -3
