ncu --set full --clock-control none --import-source on -k regex:b200e_reduce -s 2 -c 1 -f -o gpurun_out/prof_em_v16 python bench.py --workload oscillator_noise_mc --steps 3 --warmup 3 --no-cpu --no-solves > gpurun_out/ncu_em.log 2>&1
ls -la gpurun_out/prof_em_v16.ncu-rep
