#!/bin/bash
# round 2, late: the new GPU tests (one simulated day, drivers on a synthetic project), smoke, and the driver's bench
# command after the bench.py refactoring (e2e leg as a function) -- short, the round's GPU budget is nearly spent
mkdir -p gpurun_out; rm -f gpurun_out/parity_record.jsonl
S=$(date +%s)
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke done at $(( $(date +%s) - S )) s"
timeout 150 python -m pytest "tests/test_cvode_gpu.py::test_one_simulated_day" "tests/test_driver_gpu.py::test_unchanged_driver_on_a_synthetic_project" -q -s > gpurun_out/r02h_tests.log 2>&1; echo "pytest rc $? at $(( $(date +%s) - S )) s"
grep -E "model steps:|output files|passed|failed|Error|error" gpurun_out/r02h_tests.log | tail -14
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --no-strong --no-cpu > gpurun_out/r02h_bench.log 2> gpurun_out/r02h_bench.err; echo "bench rc $? at $(( $(date +%s) - S )) s"
python - <<'PY'
import json
for l in open('gpurun_out/r02h_bench.log'):
    if l.startswith('{'):
        p=json.loads(l)
        print('ms/step %.3f'%p['ms_per_step'], 'value %.4f'%p['value'], 'evals', p['rhs_evals'], 'rhs %.1f in situ %.1f'%(1e3*p['rhs_ms'],1e3*p['rhs_ms_in_situ']), 'frac %.3f'%p['roofline']['frac'], 'e2e %.3f'%p['e2e']['ms_per_step'], 'launches', p['gpu_launches'])
PY
tail -3 gpurun_out/r02h_bench.err
