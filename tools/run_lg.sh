#!/bin/bash
# each test in its own process with a hard limit (a deadlocked rank spins for minutes)
T() { timeout -k 5 ${2:-100} python -m pytest "$1" -x -q --timeout 80 --timeout-method thread -s 2>&1 | tail -${3:-15}; echo "== rc $? $1"; }
{
T "tests/test_localgroup_gpu.py::test_peer_memory_halo_rhs_bitwise[2-False]"
T "tests/test_localgroup_gpu.py::test_peer_memory_halo_rhs_bitwise[3-True]"
T "tests/test_localgroup_gpu.py::test_peer_memory_integrator_lockstep" 150 40
} > gpurun_out/lg.log 2>&1
