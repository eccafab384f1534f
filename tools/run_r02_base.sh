#!/bin/bash
# round-2 baseline on one B200: gpu tests, bench (pihm + fbr), launch list, full-set capture of the RHS kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_suite.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02_gpu_suite.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.log 2> gpurun_out/r02_bench.err; echo "bench rc $?"
timeout 300 python bench.py --steps 20 --warmup 5 --fbr --no-cpu --no-strong > gpurun_out/r02_bench_fbr.log 2> gpurun_out/r02_bench_fbr.err; echo "bench fbr rc $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_1M_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu list rc $?"
NREP=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_pre|k_main" -s 4 -c 2 -o gpurun_out/r02_rhs -f python tools/rhs_probe.py 1M > gpurun_out/r02_ncu_rhs.log 2>&1; echo "ncu rhs rc $?"
NREP=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_pre|k_main" -s 4 -c 2 -o gpurun_out/r02_rhs_fbr -f python tools/rhs_probe.py 1M fbr > gpurun_out/r02_ncu_rhs_fbr.log 2>&1; echo "ncu rhs fbr rc $?"
head -c 1500 gpurun_out/r02_bench.log
