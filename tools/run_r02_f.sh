#!/bin/bash
for v in default F0 F2; do
  if [ $v = default ]; then unset PIHM_B200_LIB; else export PIHM_B200_LIB=build_exp/$v/libpihm_b200.so; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong 2>/dev/null | python -c "
import json,sys
p=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('[$v]', 'ms/step %.3f'%p['ms_per_step'], 'evals', p['rhs_evals'], 'us/eval %.1f'%(1e3*p['ms_per_rhs_eval']), 'rhs %.1f in situ %.1f'%(1e3*p['rhs_ms'],1e3*p['rhs_ms_in_situ']), 'e2e %.3f'%p['e2e']['ms_per_step'], 'vec all frac %.3f'%p['vector_roofline']['all']['frac'])
"
done
unset PIHM_B200_LIB
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
