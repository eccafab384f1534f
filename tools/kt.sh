#!/bin/bash
# per-kernel durations of the two RHS kernels for a variant (profiling helper): tools/kt.sh <variant> [env...]
v=$1; shift
env "$@" PIHM_B200_LIB=build_exp/$v/libpihm_b200.so NREP=100 timeout 120 python tools/rhs_probe.py 1M 2>&1 | grep us/eval | sed "s/^/[$v $*] /"
env "$@" PIHM_B200_LIB=build_exp/$v/libpihm_b200.so NREP=2 timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none -k regex:"k_pre|k_main" -s 4 -c 2 python tools/rhs_probe.py 1M 2>&1 | grep -E "duration|hit_rate|inst_exec" | awk '{printf "%s ", $NF} END {print ""}' | sed "s/^/[$v $*] ncu: pre(us,l1,inst) main(us,l1,inst): /"
