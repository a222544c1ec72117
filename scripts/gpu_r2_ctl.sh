# A/B of the control-path forms of round 2 (v33): failure as a lane state (always on), and -- switchable through the
# developer hook -- the controller on the undivided sum of squares (B200E_SCALED_E2), the logarithm's mantissa taken from
# the argument's side (B200E_LOG_MANTISSA), the clamp at qmax by the high word (B200E_HIWORD_CLAMP)
set -x
mkdir -p gpurun_out
export B200E_DEV_HOOKS=1
python scimlexpectations.jl_b200/build.py > /dev/null 2>&1
OLD="#undef B200E_SCALED_E2;#define B200E_SCALED_E2 0;#undef B200E_LOG_MANTISSA;#define B200E_LOG_MANTISSA 0;#undef B200E_HIWORD_CLAMP;#define B200E_HIWORD_CLAMP 0"
for rep in 1 2; do
for defs in "" "$OLD" "#undef B200E_SCALED_E2;#define B200E_SCALED_E2 0" "#undef B200E_LOG_MANTISSA;#define B200E_LOG_MANTISSA 0" "#undef B200E_HIWORD_CLAMP;#define B200E_HIWORD_CLAMP 0"; do
  echo "== defs: $defs"
  B200E_EXTRA_DEFINES="$defs" timeout 600 python scripts/sweep.py linear2 1e7 0:0:0 2>&1 | cut -c1-220
  B200E_EXTRA_DEFINES="$defs" timeout 600 python scripts/sweep.py lorenz63 8e6 0:0:0 2>&1 | cut -c1-220
  B200E_EXTRA_DEFINES="$defs" timeout 600 python scripts/sweep.py lotka_volterra 8e6 0:0:0 2>&1 | cut -c1-220
  B200E_EXTRA_DEFINES="$defs" timeout 600 python scripts/sweep.py gbm4_lamba 2e6 0:0:0 2>&1 | cut -c1-220
done
done > gpurun_out/r2_sweep_control_forms.log 2>&1
cat gpurun_out/r2_sweep_control_forms.log
