#!/bin/bash
# final state of round 2 on one B200: the whole GPU suite (as the driver runs it) and the driver's bench command
mkdir -p gpurun_out; rm -f gpurun_out/parity_record.jsonl
S=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_gpu_suite.log 2>&1; echo "pytest rc $? at $(( $(date +%s) - S )) s"; tail -3 gpurun_out/r02j_gpu_suite.log; grep -c "stalled" gpurun_out/r02j_gpu_suite.log
timeout 100 python bench.py --gpus 1 --steps 20 --warmup 5 --no-strong --no-cpu > gpurun_out/r02j_bench.log 2> gpurun_out/r02j_bench.err; echo "bench rc $? at $(( $(date +%s) - S )) s"
python - <<'PY'
import json
for l in open('gpurun_out/r02j_bench.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('ms/step %.3f'%p['ms_per_step'], 'value %.4f'%p['value'], 'evals', p['rhs_evals'], 'rhs %.1f in situ %.1f'%(1e3*p['rhs_ms'],1e3*p['rhs_ms_in_situ']), 'frac %.3f'%p['roofline']['frac'], 'e2e %.3f'%p['e2e']['ms_per_step'], 'vec %.3f'%p['vector_roofline']['all']['frac'], 'launches', p['gpu_launches'])
PY
