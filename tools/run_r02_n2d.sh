#!/bin/bash
timeout 300 python -m pytest tests/test_localgroup_gpu.py tests/test_multigpu_gpu.py -x -q 2>&1 | tail -3
P="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $P --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-strong > gpurun_out/r02d_bench_n2.log 2> gpurun_out/r02d_bench_n2.err; echo "bench n2 rc $?"
python - <<'EOF'
import json
for l in open('gpurun_out/r02d_bench_n2.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('ms/step %.3f'%p['ms_per_step'], 'evals', p['rhs_evals'], 'us/eval %.1f'%(1e3*p['ms_per_rhs_eval']), 'rhs %.1f in situ %.1f'%(1e3*p['rhs_ms'],1e3*p['rhs_ms_in_situ']), 'e2e %.3f'%p['e2e']['ms_per_step'])
        for k,d in p['vector_roofline']['kernels'].items(): print(f"  {k:20s} {d['us']:6.1f} us  {d['frac']:.2f}")
EOF
