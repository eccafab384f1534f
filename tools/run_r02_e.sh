#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_gpu_suite.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r02e_gpu_suite.log; grep -c "stalled" gpurun_out/r02e_gpu_suite.log
for f in 1 0; do
PIHM_B200_FOLD=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-strong 2>/dev/null | python -c "
import json,sys
p=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('[fold $f] ms/step %.3f'%p['ms_per_step'], 'evals', p['rhs_evals'], 'us/eval %.1f'%(1e3*p['ms_per_rhs_eval']), 'rhs %.1f in situ %.1f'%(1e3*p['rhs_ms'],1e3*p['rhs_ms_in_situ']), 'frac %.3f'%p['roofline']['frac'], 'e2e %.3f'%p['e2e']['ms_per_step'], 'launches', p['gpu_launches'])
for k,d in p['vector_roofline']['kernels'].items(): print(f\"  {k:20s} n={d['launches']:4d} {d['us']:6.1f} us  {d['frac']:.2f}\")
"
done
