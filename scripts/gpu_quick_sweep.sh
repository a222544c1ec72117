set -x
mkdir -p gpurun_out
{
python scripts/sweep.py linear2 1e7 0:0:0,0:0:0 2>&1 | cut -c1-220
python scripts/sweep.py lotka_volterra 8e6 0:0:0 2>&1 | cut -c1-220
python scripts/sweep.py lorenz63 8e6 0:0:0 2>&1 | cut -c1-220
python scripts/sweep.py gbm4_lamba 2e6 0:0:0 2>&1 | cut -c1-220
} > gpurun_out/r2_quick_sweep.log 2>&1
cat gpurun_out/r2_quick_sweep.log
