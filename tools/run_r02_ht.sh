#!/bin/bash
timeout 200 python -m pytest tests/test_localgroup_gpu.py tests/test_multigpu_gpu.py -x -q 2>&1 | tail -3
P="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
NREP=200 timeout 120 $P --master-port 29540 tools/mgpu_rhs_probe.py 2>&1 | grep "rhs us\|Error\|error" | head -3
HT=1 PIHM_B200_LIB=build_exp/HT/libpihm_b200.so NREP=100 timeout 120 $P --master-port 29541 tools/mgpu_rhs_probe.py 2>&1 | grep "rhs us\|rank \|Error\|error" | head -20
