set -x
mkdir -p gpurun_out
export B200E_DEV_HOOKS=1
python scimlexpectations.jl_b200/build.py > /dev/null 2>&1
for defs in "" "#define B200E_RCP2 1" "#define B200E_SHORT_POLY 1" "#define B200E_RCP2 1;#define B200E_SHORT_POLY 1"; do
  echo "== defs: $defs"
  B200E_EXTRA_DEFINES="$defs" timeout 600 python scripts/sweep.py linear2 1e7 0:0:0 2>&1 | cut -c1-220
  B200E_EXTRA_DEFINES="$defs" timeout 600 python scripts/sweep.py lorenz63 8e6 0:0:0 2>&1 | cut -c1-220
done > gpurun_out/r2_sweep_mathvar.log 2>&1
cat gpurun_out/r2_sweep_mathvar.log
