#!/bin/bash
# final state of round 2 on one B200: smoke, gpu tests, the driver's bench command, launch list
mkdir -p gpurun_out; rm -f gpurun_out/parity_record.jsonl
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_gpu_suite.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02f_gpu_suite.log; grep -c "stalled" gpurun_out/r02f_gpu_suite.log
S=$(date +%s); timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02f_bench.log 2> gpurun_out/r02f_bench.err; echo "bench rc $? in $(( $(date +%s) - S )) s"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02f_bench_1M_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > gpurun_out/r02f_ncu_bench.log 2>&1; echo "ncu list rc $?"
python - <<'EOF'
import json
for l in open('gpurun_out/r02f_bench.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('ms/step %.3f'%p['ms_per_step'], 'value %.4f'%p['value'], 'evals', p['rhs_evals'], 'us/eval %.1f'%(1e3*p['ms_per_rhs_eval']), 'rhs %.1f in situ %.1f'%(1e3*p['rhs_ms'],1e3*p['rhs_ms_in_situ']), 'frac %.3f'%p['roofline']['frac'], 'e2e %.3f'%p['e2e']['ms_per_step'], 'vec %.3f'%p['vector_roofline']['all']['frac'])
        print(p['strong_scaling_8M']); print(p['cpu_baseline'])
EOF
