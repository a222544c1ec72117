set -x
mkdir -p gpurun_out
python scripts/koopman_time.py 0:0,256:2,128:0,128:4,64:0,64:8,32:0,32:16 > gpurun_out/r2_koopman_shapes.log 2>&1; cat gpurun_out/r2_koopman_shapes.log
python scripts/sweep.py lotka_volterra 8e6 256:4:0,256:4:4,256:4:6,256:4:8,256:4:12,256:4:16,256:4:32,256:3:8,128:8:8 > gpurun_out/r2_sweep_lv.log 2>&1; cat gpurun_out/r2_sweep_lv.log
python scripts/sweep.py lorenz63 8e6 256:4:0,256:4:8,256:4:16,256:4:32,256:3:0 > gpurun_out/r2_sweep_lorenz.log 2>&1; cat gpurun_out/r2_sweep_lorenz.log
