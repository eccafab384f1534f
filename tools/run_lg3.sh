#!/bin/bash
for i in 1 2 3 4 5 6; do PIHM_B200_LIB=build_exp/OLD/libpihm_b200.so timeout 45 python -m pytest tests/test_localgroup_gpu.py -x -q 2>&1 | tail -1 | sed 's/^/[old] /'; done
for i in 1 2 3 4 5 6; do timeout 45 python -m pytest tests/test_localgroup_gpu.py -x -q -o faulthandler_timeout=25 > gpurun_out/lg_new_$i.log 2>&1; tail -1 gpurun_out/lg_new_$i.log | sed 's/^/[new] /'; done
grep -h -A12 "Thread 0x\|most recent call first" gpurun_out/lg_new_*.log | grep -v "^--" | grep "File\|Thread" | grep -v "site-packages\|threading.py" | head -40
