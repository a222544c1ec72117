set -x
mkdir -p gpurun_out
N=${1:-2}
python -m pytest tests -m gpu -q -k "multi_device or sharding or pageable" > gpurun_out/r2_pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n$N.err
