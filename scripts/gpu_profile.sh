# ncu evidence for the default bench workload and for the Koopman whole solve (each ncu run only after the same
# command line exited 0 without ncu)
set -x
TAG=${1:-r2}
mkdir -p gpurun_out
ARGS="--steps 3 --warmup 3 --no-cpu --no-solves --no-strong"
python bench.py $ARGS > gpurun_out/plain_${TAG}.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py $ARGS > gpurun_out/ncu_launches_${TAG}.log 2>&1
python bench.py $ARGS > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:b200e_reduce -s 2 -c 1 -f -o gpurun_out/prof_reduce_${TAG} \
    python bench.py $ARGS > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/prof_reduce_${TAG}.ncu-rep
python scripts/koopman_time.py 0:0 > gpurun_out/plain_koopman_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_koopman_${TAG}.csv \
    python scripts/koopman_time.py 0:0 > gpurun_out/ncu_launches_koopman_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_launches_koopman_${TAG}.log
