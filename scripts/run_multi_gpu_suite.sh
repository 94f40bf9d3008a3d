#!/bin/bash
# Developer script for a `gpurun --gpus 8` call: the 1/2/4/8 scaling of the headline workload the way the driver launches
# it, BASELINE config 4 (10M-instanced scene, sample streams over 1/2/4/8 GPUs, sliced NVLink resolve and NCCL reduce) and
# config 5 (4K, 4096 spp, 4 bands x 2 streams on 8 GPUs, parity vs the reference CPU engine). Output: gpurun_out/r02_*.json
OUT=gpurun_out
mkdir -p $OUT
run() { # N workload steps reduce tag extra...
  N=$1; WL=$2; STEPS=$3; RED=$4; TAG=$5; shift 5
  if [ "$N" = "1" ]; then
    CUDA_VISIBLE_DEVICES=0 python bench.py --gpus 1 --steps $STEPS --warmup 5 --workload $WL --no-aux --no-cpu-baseline --no-dropin "$@" 2> $OUT/r02_$TAG.err | tail -1 > $OUT/r02_$TAG.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29$((600+N)) bench.py --gpus $N --steps $STEPS --warmup 5 --workload $WL --reduce $RED "$@" 2> $OUT/r02_$TAG.err | tail -1 > $OUT/r02_$TAG.json
  fi
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/r02_$TAG.json").read())
    print("$TAG", "N=%d" % d["n_gpus"], "%.1f Mrays/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.1f" % d["e2e"]["value"], d["run"])
except Exception as e:
    print("$TAG failed", e)
PY
}
for N in 1 2 4 8; do run $N heightfield_1m_1080p 20 ipc scale_hf_n$N; done
for N in 1 2 4 8; do run $N instancing_10m_1080p 64 ipc config4_ipc_n$N; done
run 8 instancing_10m_1080p 64 nccl config4_nccl_n8
run 2 instancing_10m_1080p 64 nccl config4_nccl_n2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 tests/tools/config5_parity.py --spp 4096 2> $OUT/r02_config5.err | tail -1 > $OUT/r02_config5.json
python -c "
import json; d=json.loads(open('$OUT/r02_config5.json').read()); print({k:v for k,v in d.items() if k not in ('expected','tolerance','reference')})"
