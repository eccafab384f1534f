#!/bin/bash
timeout 150 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -2
P="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 $P --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-strong > gpurun_out/r02e_bench_n2.log 2> gpurun_out/r02e_bench_n2.err; echo "bench n2 rc $?"
python - <<'EOF'
import json
for l in open('gpurun_out/r02e_bench_n2.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('ms/step %.3f'%p['ms_per_step'], 'value %.4f'%p['value'], 'evals', p['rhs_evals'], 'us/eval %.1f'%(1e3*p['ms_per_rhs_eval']), 'rhs %.1f in situ %.1f'%(1e3*p['rhs_ms'],1e3*p['rhs_ms_in_situ']), 'e2e %.3f'%p['e2e']['ms_per_step'])
EOF
tail -n 3 gpurun_out/r02e_bench_n2.err | cut -c1-300
