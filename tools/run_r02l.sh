PIHM_B200_LIB=build_exp/M11/libpihm_b200.so NREP=2 ncu --set full --clock-control none --import-source on -k regex:"k_pre|k_main" -s 4 -c 2 -o gpurun_out/r02l_rhs -f python tools/rhs_probe.py 1M > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
