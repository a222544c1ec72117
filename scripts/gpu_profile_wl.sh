# ncu full capture of the reduce kernel of another workload: gpu_profile_wl.sh <workload> <tag> [extra bench args]
WL=${1:-lotka_volterra_mc}; TAG=${2:-lv}; shift 2
python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu --no-solves "$@" > gpurun_out/plain_${TAG}.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:b200e_reduce -s 2 -c 1 -f -o gpurun_out/prof_${TAG} python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu --no-solves "$@" > gpurun_out/ncu_${TAG}.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
