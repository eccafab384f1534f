nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
P() { PIHM_B200_LIB=build_exp/$1/libpihm_b200.so timeout 300 python tools/rhs_parity_probe.py "${@:2}" 2>&1 | grep "^\[" | cut -c1-200; }
P G11 1M; PARITY=0 P G11s10 1M
P G11 1M fbr; PARITY=0 P G11f4 1M fbr
P G11 100k; P G11 100k fbr
PIHM_B200_LIB=build_exp/G11/libpihm_b200.so python -m pytest tests/test_rhs_gpu.py tests/test_multigpu_gpu.py -x -q 2>&1 | tail -5
PIHM_B200_LIB=build_exp/G11/libpihm_b200.so NREP=2 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio --clock-control none -k regex:"k_pre|k_main" -s 4 -c 2 python tools/rhs_probe.py 1M 2>&1 | grep -E "k_pre|k_main|duration|inst_exec|long_sc"
