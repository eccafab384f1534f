K() { PIHM_B200_LIB=build_exp/$1/libpihm_b200.so NREP=100 timeout 120 python tools/rhs_probe.py "${@:2}" 2>&1 | grep us/eval | sed "s/^/[$*] /"
PIHM_B200_LIB=build_exp/$1/libpihm_b200.so NREP=2 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.per_cycle_active --clock-control none -k regex:"k_pre|k_main" -s 4 -c 2 python tools/rhs_probe.py "${@:2}" 2>&1 | grep -E "duration|inst_exec|issue_active" | awk '{printf "%s ", $NF} END {print ""}' | sed "s/^/[$*] pre(us,inst,issue) main(..): /"; }
for v in $VARIANTS; do K $v 1M; done
for v in $FBRV; do K $v 1M fbr; done
for v in $PARV; do PIHM_B200_LIB=build_exp/$v/libpihm_b200.so timeout 300 python tools/rhs_parity_probe.py 1M 2>&1 | grep "max err" | cut -c1-150; done
