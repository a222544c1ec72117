set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "sampler or generated or twin or lamba_sampler or full_size_properties_through" > gpurun_out/pytest_sampler.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_sampler.log
timeout 600 python bench.py --no-cpu --no-solves > gpurun_out/bench_v26.json 2> gpurun_out/bench_v26.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_v26.json').read()); print(d['value'], d['roofline']['kernel_ms'], d['e2e_device_sampling']['value'], d['e2e_device_sampling']['kernel_ms'], d['e2e']['value'])"
timeout 300 python bench.py --workload lorenz63_mc --no-cpu --no-solves --steps 10 > gpurun_out/bench_v26_lorenz.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_v26_lorenz.json').read()); print(d['value'], d['roofline']['kernel_ms'], d['e2e_device_sampling']['value'], d['e2e_device_sampling']['kernel_ms'])"
