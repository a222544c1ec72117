# ncu evidence for the default bench workload (run only after the plain bench exited 0)
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu --no-solves > gpurun_out/plain_${TAG}.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-solves > gpurun_out/ncu_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:b200e_reduce -s 2 -c 1 -f -o gpurun_out/prof_reduce_${TAG} \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-solves > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/prof_reduce_${TAG}.ncu-rep
