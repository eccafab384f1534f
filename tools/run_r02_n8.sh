#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.log 2> gpurun_out/r02_bench_n$N.err; echo "bench n$N rc $?"
python - <<EOF
import json
for l in open('gpurun_out/r02_bench_n$N.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('ms/step %.3f'%p['ms_per_step'], 'value %.4f'%p['value'], 'evals', p['rhs_evals'], 'us/eval %.1f'%(1e3*p['ms_per_rhs_eval']), 'rhs %.1f in situ %.1f'%(1e3*p['rhs_ms'],1e3*p['rhs_ms_in_situ']), 'e2e %.3f'%p['e2e']['ms_per_step'], p['halo_path'], p['strong_scaling_8M'])
EOF
