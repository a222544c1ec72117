python scripts/sweep.py oscillator_noise 2e6 256:0:0,256:2:0,256:3:0,128:4:0,128:6:0 2>&1
B200E_EXTRA_DEFINES="#define B200E_EM_SPT 3" python scripts/sweep.py oscillator_noise 2e6 256:2:0,128:4:0,128:3:0 2>&1
B200E_EXTRA_DEFINES="#define B200E_EM_SPT 4" python scripts/sweep.py oscillator_noise 2e6 256:2:0,128:4:0,128:2:0,256:1:0 2>&1
B200E_EXTRA_DEFINES="#define B200E_EM_SPT 1" python scripts/sweep.py oscillator_noise 2e6 256:4:0 2>&1
python scripts/sweep.py gbm4 2e6 256:0:0,256:3:0 2>&1
