#!/bin/bash
# 2 GPUs: back-to-back RHS of the partitioned 2M mesh per rank (with / without the flag wait), then the bench
mkdir -p gpurun_out
for d in 0 1; do
PIHM_B200_HALO_DEBUG=$d NREP=200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/mgpu_rhs_probe.py 2>&1 | grep "rhs us"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-strong > gpurun_out/r02_bench_n2.log 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc $?"
tail -c 3000 gpurun_out/r02_bench_n2.log | cut -c1-2200; tail -n 3 gpurun_out/r02_bench_n2.err
