OUT=gpurun_out
for N in 1 2 4 8; do
  if [ "$N" = "1" ]; then
    CUDA_VISIBLE_DEVICES=0 python bench.py --gpus 1 --steps 20 --warmup 5 --no-aux --no-cpu-baseline --no-dropin 2> $OUT/r02_final_scale_n$N.err | tail -1 > $OUT/r02_final_scale_n$N.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29$((700+N)) bench.py --gpus $N --steps 20 --warmup 5 2> $OUT/r02_final_scale_n$N.err | tail -1 > $OUT/r02_final_scale_n$N.json
  fi
  python -c "
import json; d=json.loads(open('$OUT/r02_final_scale_n$N.json').read()); print('N=%d' % d['n_gpus'], '%.1f Mrays/s' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d['run']['resolve_ms'])"
done
