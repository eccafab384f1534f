#!/bin/bash
# multi-GPU bench + parity on N GPUs of one box: tools/run_mgpu.sh N [extra bench flags]
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
tail -c 2500 gpurun_out/bench_n$N.log; tail -3 gpurun_out/bench_n$N.err
