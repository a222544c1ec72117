python scripts/sweep.py linear2 1e7 256:4:0,256:3:0 2>&1
B200E_EXTRA_DEFINES="#define B200E_DYNAMIC 1" python scripts/sweep.py linear2 1e7 256:4:0,256:3:0,256:4:32 2>&1
python scripts/sweep.py linear2 4e7 256:4:0 2>&1
B200E_EXTRA_DEFINES="#define B200E_DYNAMIC 1" python scripts/sweep.py linear2 4e7 256:4:0 2>&1
python scripts/sweep.py lotka_volterra,lorenz63 4e6 256:4:0 2>&1
B200E_EXTRA_DEFINES="#define B200E_DYNAMIC 1" python scripts/sweep.py lotka_volterra,lorenz63 4e6 256:4:0,256:4:32 2>&1
