NREP=100 python tools/rhs_probe.py 1M 2>&1 | grep us/eval
NREP=100 python tools/rhs_probe.py 1M fbr 2>&1 | grep us/eval
NREP=2 ncu --set full --clock-control none --import-source on -k regex:"k_pre|k_main" -s 4 -c 2 -o gpurun_out/r02n_rhs -f python tools/rhs_probe.py 1M > gpurun_out/ncu_n.log 2>&1
tail -2 gpurun_out/ncu_n.log
