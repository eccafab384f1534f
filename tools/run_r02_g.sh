#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_gpu_suite.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02g_gpu_suite.log; grep -c "stalled" gpurun_out/r02g_gpu_suite.log
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-strong > gpurun_out/r02g_bench.log 2> gpurun_out/r02g_bench.err; echo "bench rc $?"
python - <<'EOF'
import json
for l in open('gpurun_out/r02g_bench.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('ms/step %.3f'%p['ms_per_step'], 'value %.4f'%p['value'], 'evals', p['rhs_evals'], 'us/eval %.1f'%(1e3*p['ms_per_rhs_eval']), 'rhs %.1f in situ %.1f'%(1e3*p['rhs_ms'],1e3*p['rhs_ms_in_situ']), 'frac %.3f'%p['roofline']['frac'], 'e2e %.3f'%p['e2e']['ms_per_step'], 'vec %.3f'%p['vector_roofline']['all']['frac'], 'launches', p['gpu_launches'])
EOF
