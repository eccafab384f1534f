#!/bin/bash
# run tools/rhs_probe.py for each built variant (profiling helper): tools/variants.sh A B C ...
for v in "$@"; do
  PIHM_B200_LIB=build_exp/$v/libpihm_b200.so NREP=${NREP:-100} python tools/rhs_probe.py ${SIZE:-1M} 2>&1 | sed "s/^/[$v] /"
done
