set -x
mkdir -p gpurun_out
export B200E_DEV_HOOKS=1
python scimlexpectations.jl_b200/build.py > /dev/null 2>&1
for defs in "#define B200E_EM_PREFETCH 0" "#define B200E_EM_PREFETCH 1"; do
  echo "== defs: $defs"
  B200E_EXTRA_DEFINES="$defs" timeout 600 python scripts/sweep.py oscillator_noise,gbm4 2e6 0:0:0,128:3:0,128:2:0,256:2:0 2>&1 | cut -c1-220
done > gpurun_out/r2_sweep_em_prefetch.log 2>&1
cat gpurun_out/r2_sweep_em_prefetch.log
