# A/B of B200E_STAY_IN_LOOP (v34 candidate, measured and NOT taken: the macro is no longer in the tree; kept for the record): a lane finishing below the refill threshold is handled next to the step loop (a ballot
# and a counter) instead of by a pass through the refill phase
set -x
mkdir -p gpurun_out
export B200E_DEV_HOOKS=1
python scimlexpectations.jl_b200/build.py > /dev/null 2>&1
for rep in 1 2; do
for defs in "" "#define B200E_STAY_IN_LOOP 0"; do
  echo "== defs: $defs"
  B200E_EXTRA_DEFINES="$defs" timeout 600 python scripts/sweep.py linear2 1e7 0:0:0 2>&1 | cut -c1-220
  B200E_EXTRA_DEFINES="$defs" timeout 600 python scripts/sweep.py lorenz63 8e6 0:0:0 2>&1 | cut -c1-220
  B200E_EXTRA_DEFINES="$defs" timeout 600 python scripts/sweep.py lotka_volterra 8e6 0:0:0 2>&1 | cut -c1-220
  B200E_EXTRA_DEFINES="$defs" timeout 600 python scripts/sweep.py gbm4_lamba 2e6 0:0:0 2>&1 | cut -c1-220
done
done > gpurun_out/r2_sweep_stay_in_loop.log 2>&1
cat gpurun_out/r2_sweep_stay_in_loop.log
