python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q -x -k "ordering or mean_samples or rejects or resolve" 2>&1 | tail -8; python scripts/gpu_variants.py --passes 96 --workloads heightfield_1m_1080p,materials_1080p --variants "default:;r16s8:RZB200_TRACE=refill;r8s8:RZB200_TRACE=refill,RZB200_REFILL_THRESH=8;r24s8:RZB200_TRACE=refill,RZB200_REFILL_THRESH=24;r16s4:RZB200_TRACE=refill,RZB200_REFILL_SLICE=4;r16s16:RZB200_TRACE=refill,RZB200_REFILL_SLICE=16;r16s64:RZB200_TRACE=refill,RZB200_REFILL_SLICE=64;r4s16:RZB200_TRACE=refill,RZB200_REFILL_THRESH=4,RZB200_REFILL_SLICE=16" 2>/dev/null > gpurun_out/r2_var8.jsonl; python - <<PY
import json
for l in open("gpurun_out/r2_var8.jsonl"):
    d=json.loads(l); print(d["workload"][:12], d["variant"], "%.1f Mrays/s trace %.3f shade %.3f sort %.3f shadow %.3f" % (d["Mrays_s"], d["trace_ms"], d["shade_ms"], d["sort_ms"], d["shadow_ms"]), d["mean_radiance_rel_to_first"][0])
PY
