#!/bin/bash
# usage: run_bench_multi.sh N [extra bench args]  -- launches bench.py the way the driver does for N > 1
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"
