for m in oscillator_noise gbm4; do
for spt in 2 3; do
B200E_EXTRA_DEFINES="#define B200E_EM_SPT $spt" python scripts/sweep.py $m 2e6 256:2:0,128:4:0,128:3:0,96:4:0 2>&1 | sed "s/^/spt=$spt /"
done; done
