#!/bin/bash
# BASELINE config[3]: pihm-fbr, 1M triangles per GPU, on 2 B200 (the 1-GPU line is profiles/r02_bench_1M_fbr.json)
mkdir -p gpurun_out
P="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
S=$(date +%s)
timeout 140 $P --master-port 29547 bench.py --gpus 2 --steps 20 --warmup 5 --fbr --no-strong > gpurun_out/r02i_bench_fbr_n2.log 2> gpurun_out/r02i_bench_fbr_n2.err; echo "bench fbr n2 rc $? in $(( $(date +%s) - S )) s"
python - <<'PY'
import json
for l in open('gpurun_out/r02i_bench_fbr_n2.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('ms/step %.3f'%p['ms_per_step'], 'value %.4f'%p['value'], 'evals', p['rhs_evals'], 'us/eval %.1f'%(1e3*p['ms_per_rhs_eval']), 'rhs %.1f in situ %.1f'%(1e3*p['rhs_ms'],1e3*p['rhs_ms_in_situ']), 'e2e %.3f'%p['e2e']['ms_per_step'], p['halo_path'])
PY
tail -n 3 gpurun_out/r02i_bench_fbr_n2.err | cut -c1-300
