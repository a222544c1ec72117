set -x
mkdir -p gpurun_out
export B200E_DEV_HOOKS=1
python scimlexpectations.jl_b200/build.py > /dev/null 2>&1
for defs in "#define B200E_ZC_SMEM 0;#define B200E_LAZY_CONTROLLER 0" "#define B200E_ZC_SMEM 1;#define B200E_LAZY_CONTROLLER 0" "#define B200E_ZC_SMEM 0;#define B200E_LAZY_CONTROLLER 1" "#define B200E_ZC_SMEM 1;#define B200E_LAZY_CONTROLLER 1"; do
  echo "== defs: $defs"
  B200E_EXTRA_DEFINES="$defs" timeout 600 python scripts/sweep.py gbm4_lamba 2e6 0:0:0,256:4:0,256:3:0 2>&1 | cut -c1-220
done > gpurun_out/r2_sweep_lamba.log 2>&1
cat gpurun_out/r2_sweep_lamba.log
