#!/bin/bash
# last GPU seconds of round 2: pipelined host transfers (csrc/transfer.cu) -- bitwise test, then the bench line with --e2e pipelined
mkdir -p gpurun_out
S=$(date +%s)
timeout 16 python -m pytest tests/test_transfer_gpu.py -x -q > gpurun_out/r02k_transfer_tests.log 2>&1; echo "pytest rc $? at $(( $(date +%s) - S )) s"; tail -4 gpurun_out/r02k_transfer_tests.log
timeout 16 python bench.py --gpus 1 --steps 20 --warmup 5 --no-strong --no-cpu --e2e pipelined > gpurun_out/r02k_bench_pipelined.log 2> gpurun_out/r02k_bench_pipelined.err; echo "bench rc $? at $(( $(date +%s) - S )) s"
python - <<'PY'
import json
for l in open('gpurun_out/r02k_bench_pipelined.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('ms/step %.3f'%p['ms_per_step'], 'e2e %.3f'%p['e2e']['ms_per_step'], p['e2e'].get('transfers'), 'frac %.3f'%p['roofline']['frac'])
PY
tail -2 gpurun_out/r02k_bench_pipelined.err
