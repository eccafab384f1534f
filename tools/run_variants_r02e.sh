nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
K() { PIHM_B200_LIB=build_exp/$1/libpihm_b200.so NREP=100 timeout 120 python tools/rhs_probe.py "${@:2}" 2>&1 | grep us/eval | sed "s/^/[$*] /"
PIHM_B200_LIB=build_exp/$1/libpihm_b200.so NREP=2 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.per_cycle_active --clock-control none -k regex:"k_pre|k_main" -s 4 -c 2 python tools/rhs_probe.py "${@:2}" 2>&1 | grep -E "duration|inst_exec|long_sc|issue_active" | awk '{printf "%s ", $NF} END {print ""}' | sed "s/^/[$*] pre(us,issue,longsb,inst) main(..): /"; }
K H11 1M; K H11fake 1M; K H11noorder 1M
K H11 1M fbr; K H11fake 1M fbr
