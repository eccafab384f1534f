#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_driver_gpu.py -x -q -s > gpurun_out/r02b_driver.log 2>&1; echo "driver test rc $?"; tail -12 gpurun_out/r02b_driver.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong > gpurun_out/r02b_bench.log 2> gpurun_out/r02b_bench.err; echo "bench rc $?"; tail -c 2500 gpurun_out/r02b_bench.log; tail -3 gpurun_out/r02b_bench.err
