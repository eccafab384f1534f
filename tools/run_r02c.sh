python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/rhs_probe.py 1M > gpurun_out/probe_c.log 2>&1
NREP=2 ncu --set full --clock-control none --import-source on -k regex:"k_pre|k_main" -s 4 -c 2 -o gpurun_out/r02c_rhs -f python tools/rhs_probe.py 1M > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/probe_c.log gpurun_out/ncu_c.log
