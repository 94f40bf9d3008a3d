#!/usr/bin/env python3
"""bench.py -- the headline measurement of the B200 render path (BASELINE.json: Mrays/s at 1080p).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A *step* is one pass of the hot path over one frame: one path segment per pixel (closest-hit trace, shade + NEE,
shadow trace), W*H rays -- the reference's own unit (cuda_render_kernel.cu:122-129: ray_count += W*H per pass).
One JSON line is printed by rank 0:
  value      Mrays/s, whole job, scene and path state resident in HBM, K passes timed on the device (CUDA events on
             the launching stream), barrier + synchronize on both sides, max over ranks
  e2e        the same metric through the reference-facing boundary with HOST buffers: every e2e step is one
             Engine::renderWorld-equivalent frame = rzb_set_scene (host arrays -> device) + rzb_set_camera + rzb_reset
             + rzb_render(rpp passes) + rzb_resolve (tone map, RGBA8 + depth -> pinned host buffers)
  roofline   HBM roofline from ALGORITHMIC bytes of both traversal kernels; the one with the larger share of
             the step is reported at the top level
  cpu_baseline  the reference's own CPU engine (oracle/_ref/rz_ref_tool, built from /root/reference) on this box's
             host cores on a bounded sample of the same workload
`--impl reference` times that CPU engine alone and prints the same line with "impl": "reference".
N > 1 (torchrun): sample streams are sharded over the ranks (same frame, disjoint RNG streams, weak scaling), no
collective while rendering; the accumulators are combined once, inside the timed region, by the fused
IPC/NVLink sum + tone-map kernel on rank 0 (NCCL reduce with --reduce nccl).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on (fits one GPU)
    "materials_1080p": ("materials", dict(resolution=(1920, 1080), res=64)),
    # configs[2] geometry (1M triangles, maps, DoF), kept as a secondary line in "aux"
    "heightfield_1m_1080p": ("heightfield_1m", dict(resolution=(1920, 1080))),
    "instancing_10m_1080p": ("instancing_10m", dict(resolution=(1920, 1080))),
    "cornell_512": ("cornell", dict(resolution=(512, 512))),
    # configs[4]: the 1M-triangle scene at 3840x2160 (use with --split hybrid on 8 GPUs: 4 row bands x 2 sample streams)
    "heightfield_1m_4k": ("heightfield_1m", dict(resolution=(3840, 2160))),
}
# reference-layout byte constants for the algorithmic traffic figure (SURVEY.md 8d / DESIGN.md)
B_NODE, B_TRI, B_STATE, B_HIT = 48, 144, 57, 24
B_SHADOW_REC, B_SHADOW_ACC = 48, 12  # shadow queue record read, accumulator update
MAX_DEPTH = 16          # Tracing::maxDepth default, engine_parts.hpp:101
RPP_E2E = 64            # passes per renderWorld call the headless auto-tuner converges to (headless.cpp:287-295)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML: the same counters nvidia-smi prints)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


BVH = "reference"  # --bvh: "reference" = the reference's own trees (parity mode), "sah" = the optional SAH builder


def build_world(workload):
    from rayzath_b200 import scenes
    name, kw = WORKLOADS[workload]
    w = scenes.CONFIGS[name](**kw)
    if BVH != "reference":
        for m in w.meshes:
            m.bvh_builder = (BVH, int(os.environ.get("RZB200_SAH_MAX_LEAF", "4")))
    return w


# ---------------------------------------------------------------------------------------------- reference arm
def reference_run(workload, steps, warmup, budget_s=240.0):
    """The reference's own CPU engine (oracle/_ref/rz_ref_tool render) on this box's cores. A step = one
    Engine::renderWorld(CPU) = one pass over the frame. The frame is the workload's own (full resolution) whenever
    (steps + warmup) passes fit the time budget; otherwise a bounded sample of it: same scene and camera at the
    largest resolution from a list of 128x128-tile-friendly sizes (the CPU engine hands 128x128 tiles to its worker
    threads, cpu_engine_renderer.cpp:186-205; odd small frames would leave threads idle and understate it)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rz_oracle as O
    from rayzath_b200 import scenes
    name, kw = WORKLOADS[workload]
    full = tuple(kw["resolution"])
    tmp = tempfile.mkdtemp(prefix="rzb_bench_ref_")

    def run(res, passes, wu, timeout):
        kw2 = dict(kw)
        kw2["resolution"] = res
        w = scenes.CONFIGS[name](**kw2)
        path = w.save_reference(os.path.join(tmp, "%dx%d" % res))
        return O.ref_tool("render", path, passes, "-", MAX_DEPTH, 1, 1, wu, timeout=timeout, attempts=2)

    probe = run(full, 3, 1, 300.0)  # 2 timed full-resolution passes
    s_per_px = probe["seconds"] / max(probe["timed_rays"], 1)
    aspect = full[0] / full[1]
    options = [full] + [r for r in ((1536, 896), (1280, 768), (1024, 640), (768, 512), (512, 384), (384, 256), (256, 128))
                        if r[0] * r[1] < full[0] * full[1]]
    res = options[-1]
    for r in options:
        if (steps + warmup) * r[0] * r[1] * s_per_px <= budget_s:
            res = r
            break
    info = run(res, steps + warmup, max(warmup, 1), max(120.0, 4.0 * budget_s))
    timed_passes = max(info["timed_passes"], 1)
    value = info["timed_rays"] / info["seconds"] / 1e6
    frac = (res[0] * res[1]) / float(full[0] * full[1])
    return {
        "value": value, "ms_per_step": info["seconds"] / timed_passes * 1e3, "cores": info["threads"],
        "sample": "%d passes of the %s scene at %dx%d (%.0f %% of the pixels of %dx%d), max depth %d, reference CPU engine"
                  % (timed_passes, workload, res[0], res[1], 100.0 * frac, full[0], full[1], MAX_DEPTH),
        "kind": "reference",
    }


# ---------------------------------------------------------------------------------------------- our arm
def pin(arrays):
    """Page-lock the host arrays the boundary reads from (cudaHostRegister through torch)."""
    import torch
    rt = torch.cuda.cudart()
    for a in arrays:
        if a.size:
            rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--workload", default="materials_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reduce", default="ipc", choices=["ipc", "nccl"])
    ap.add_argument("--split", default="samples", choices=["samples", "tiles", "hybrid"],
                    help="N>1: samples = every rank renders the whole frame with its own RNG stream (weak scaling, default); "
                         "tiles = rank r renders row band r of N (strong scaling); hybrid = N/2 bands x 2 streams")
    ap.add_argument("--bvh", default="reference", choices=["reference", "sah", "lbvh"],
                    help="triangle trees: the reference's own (parity mode, default), the optional SAH builder (host) or the "
                         "optional linear-BVH builder (GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global BVH
    BVH = args.bvh

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    base = {"metric": "Mrays/s at 1080p (path segments per second, passes*W*H/t)", "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak" if args.split == "samples" else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = reference_run(args.workload, args.steps, args.warmup)
        line = dict(base)
        line.update({"impl": "reference", "value": r["value"], "ms_per_step": r["ms_per_step"],
                     "config": {"workload": args.workload, "resolution": list(WORKLOADS[args.workload][1]["resolution"]),
                                "max_depth": MAX_DEPTH, "light_samples": [1, 1],
                                "bvh": "reference trees (built by the reference itself)", "sharding": "host threads"},
                     "cpu_baseline": {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"],
                                      "sample": r["sample"]},
                     "e2e": {"value": r["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})
        print(json.dumps(line), flush=True)
        return 0

    import numpy as np
    import torch
    from rayzath_b200 import capi, parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        dist = parallel.init_process_group("nccl")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    world_obj = build_world(args.workload)
    flat = world_obj.flatten()
    cam = world_obj.camera_struct()
    W, H = int(cam[0]["width"]), int(cam[0]["height"])
    n_px = W * H
    peak, peak_src = load_peaks()

    ctx = capi.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_scene(flat)
    ctx.set_camera(cam)
    # how the frame is sharded over the ranks (no data-path collective in any mode; accumulators are summed at resolve)
    bands, streams = 1, world
    if world > 1 and args.split == "tiles":
        bands, streams = world, 1
    elif world > 1 and args.split == "hybrid" and world % 2 == 0:
        bands, streams = world // 2, 2
    band, stream_id = rank // streams, rank % streams
    if bands > 1:
        ctx.set_row_interleave(band, bands)  # 16-row chunk rows dealt round-robin: balances sky rows and geometry rows
    n_px = W * sum(min(r * 16 + 16, H) - r * 16 for r in range((H + 15) // 16) if r % bands == band)  # pixels per pass on this rank
    seed = parallel.stream_seed(20261018, stream_id)
    ctx.set_config(1, 1, MAX_DEPTH, capi.FLAG_NONE, seed)
    ctx.reset()
    rgba_host = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy()
    depth_host = torch.empty((H, W), dtype=torch.float32, pin_memory=True).numpy()

    fused = parallel.FusedResolve(ctx) if (world > 1 and args.reduce == "ipc") else None

    def combine():
        """the one exchange step of the N>1 path: sum the accumulators onto rank 0 (and tone-map there)"""
        if world == 1:
            return
        if fused is not None:
            fused(want_depth=True)
        else:
            parallel.reduce_accum(ctx.accum_tensor().clone(), dst=0)  # a copy: rendering continues on the original

    # ---- warm-up, then the timed region: K passes (+ the exchange step), device-timed on the launching stream
    ctx.render(args.warmup)
    combine()
    barrier()
    alpha0 = float(ctx.read_accum()[..., 3].mean())
    launches0 = int(ctx.render_stats()["kernel_launches"])
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record(stream)
    ctx.render(args.steps)
    combine()
    ev1.record(stream)
    ev1.synchronize()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    # per-kernel device time per launch: CUDA events recorded around each stage of (up to 256 of) the timed passes
    st = ctx.render_stats()
    trace_ms, shade_ms, shadow_ms = float(st["last_trace_ms"]), float(st["last_shade_ms"]), float(st["last_shadow_ms"])
    launches = int(st["kernel_launches"]) - launches0
    t_all = torch.tensor([ms, (float(ctx.read_accum()[..., 3].mean()) - alpha0), float(args.steps) * n_px],
                         dtype=torch.float64, device="cuda")
    ms_max, spp_sum, rays_sum = float(t_all[0].item()), float(t_all[1].item()), float(t_all[2].item())
    if dist is not None:
        t_max = t_all.clone()
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_all, op=dist.ReduceOp.SUM)
        ms_max, spp_sum, rays_sum = float(t_max[0].item()), float(t_all[1].item()), float(t_all[2].item())
    value = rays_sum / (ms_max * 1e-3) / 1e6
    spp_per_s = spp_sum / (ms_max * 1e-3)  # completed camera paths per pixel per second, whole job

    # ---- algorithmic bytes of the dominant kernel: replay the same passes with the counting kernels
    # (the RNG is counter-based on (seed, pixel, pass): after a reset the same pass indices retrace the same rays)
    ctx.reset()
    ctx.render(args.warmup)
    ctx.set_config(1, 1, MAX_DEPTH, capi.FLAG_COUNT_WORK, seed)
    ctx.render(args.steps)
    wc = ctx.work_counters()
    seg = max(int(wc["segments"]), 1)
    nodes_per_seg = float(wc["closest_top_nodes"] + wc["closest_mesh_nodes"]) / seg
    tris_per_seg = float(wc["closest_triangles"]) / seg
    shadow_per_seg = float(wc["shadow_rays"]) / seg
    # share of lane-time the whole-warp batches keep busy (work of a ray = pair steps + triangle tests)
    lane_util = {"closest": float(wc["closest_lane_work"]) / max(int(wc["closest_batch_work"]), 1),
                 "shadow": float(wc["shadow_lane_work"]) / max(int(wc["shadow_batch_work"]), 1),
                 "dropped_non_finite_samples": int(wc["invalid_rays"])}
    bytes_per_seg = B_STATE + nodes_per_seg * B_NODE + tris_per_seg * B_TRI + B_HIT
    # the shadow kernel in the same currency: per shadow ray the 48-byte queue record, the box and triangle tests of
    # the any-hit walk (reference layout: 48 B nodes, 144 B triangles) and the 12-byte accumulator update
    n_shadow = max(int(wc["shadow_rays"]), 1)
    sh_nodes = float(wc["shadow_top_nodes"] + wc["shadow_mesh_nodes"]) / n_shadow
    sh_tris = float(wc["shadow_triangles"]) / n_shadow
    bytes_per_shadow_ray = B_SHADOW_REC + sh_nodes * B_NODE + sh_tris * B_TRI + B_SHADOW_ACC
    prof = {}
    prof_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof_path):
        try:
            prof = json.load(open(prof_path)).get(args.workload if BVH == "reference" else args.workload + "_sah", {})
        except Exception:
            prof = {}
    kernels = {
        "k_trace_paths": {"ms": trace_ms, "units_per_launch": n_px, "algorithmic_bytes_per_unit": bytes_per_seg,
                          "achieved": bytes_per_seg * n_px / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else None,
                          "traffic": prof.get("k_trace_paths_dram_bytes_per_launch")},
        "k_trace_shadow": {"ms": shadow_ms, "units_per_launch": shadow_per_seg * n_px,
                           "algorithmic_bytes_per_unit": bytes_per_shadow_ray,
                           "achieved": bytes_per_shadow_ray * shadow_per_seg * n_px / (shadow_ms * 1e-3) / 1e9 if shadow_ms > 0 else None,
                           "traffic": prof.get("k_trace_shadow_dram_bytes_per_launch")},
    }
    dominant = max(kernels, key=lambda k: kernels[k]["ms"])
    achieved, traffic = kernels[dominant]["achieved"], kernels[dominant]["traffic"]
    ctx.set_config(1, 1, MAX_DEPTH, capi.FLAG_NONE, seed)

    # ---- e2e through the boundary with host buffers (every step = one renderWorld-equivalent frame)
    host_arrays = [np.ascontiguousarray(v) for v in flat.values()]
    flat_pinned = dict(zip(flat.keys(), host_arrays))
    try:
        pin(host_arrays)
    except Exception:
        pass
    h2d = sum(a.nbytes for a in host_arrays) + cam.nbytes + capi.config_dtype.itemsize
    d2h = rgba_host.nbytes + depth_host.nbytes
    e2e_frames = max(2, min(8, args.steps // RPP_E2E))

    def frame():
        ctx.set_scene(flat_pinned)
        ctx.set_camera(cam)
        ctx.set_config(1, 1, MAX_DEPTH, capi.FLAG_NONE, seed)
        ctx.reset()
        ctx.render(RPP_E2E)
        if world > 1:
            if fused is not None:
                fused(want_depth=True)
            else:
                parallel.reduce_accum(ctx.accum_tensor(), dst=0)
                if rank == 0:
                    ctx.resolve(rgba_host, depth_host)
        else:
            ctx.resolve(rgba_host, depth_host)

    frame()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_frames):
        frame()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_rays = torch.tensor([float(e2e_frames) * RPP_E2E * n_px], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_rays, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_rays.item()) / float(e2e_t.item()) / 1e6

    # ---- secondary workload (config 3 geometry) and the CPU baseline: rank 0, N=1 only
    aux = None
    cpu = None
    if rank == 0 and world == 1:
        if not args.no_aux and args.workload != "heightfield_1m_1080p":
            try:
                w2 = build_world("heightfield_1m_1080p")
                ctx.set_scene(w2.flatten())
                ctx.set_camera(w2.camera_struct())
                ctx.reset()
                ctx.render(args.warmup)
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n2 = min(args.steps, 256)
                a0.record(stream)
                ctx.render(n2)
                a1.record(stream)
                a1.synchronize()
                st2 = ctx.render_stats()
                aux = {"workload": "heightfield_1m_1080p", "triangles": int(w2.flatten()["triangles"].shape[0]),
                       "value": n2 * n_px / (a0.elapsed_time(a1) * 1e-3) / 1e6, "unit": "Mrays/s", "steps": n2,
                       "trace_ms": float(st2["last_trace_ms"]), "shade_ms": float(st2["last_shade_ms"]),
                       "shadow_ms": float(st2["last_shadow_ms"])}
            except Exception as e:  # the headline line must still be printed
                aux = {"workload": "heightfield_1m_1080p", "error": repr(e)}
        # the same workload on the optional SAH trees (RZB_SCENE_OWN_TREES; records equal except exact ties)
        if not args.no_aux and BVH == "reference":
            own = {}
            try:
                BVH = "sah"
                for wl in (args.workload, "heightfield_1m_1080p"):
                    w3 = build_world(wl)
                    ctx.set_scene(w3.flatten())
                    ctx.set_camera(w3.camera_struct())
                    ctx.reset()
                    ctx.render(args.warmup)
                    torch.cuda.synchronize()
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    n3 = min(args.steps, 256)
                    a0.record(stream)
                    ctx.render(n3)
                    a1.record(stream)
                    a1.synchronize()
                    st3 = ctx.render_stats()
                    own[wl] = {"value": n3 * n_px / (a0.elapsed_time(a1) * 1e-3) / 1e6, "unit": "Mrays/s", "steps": n3,
                               "trace_ms": float(st3["last_trace_ms"]), "shade_ms": float(st3["last_shade_ms"]),
                               "shadow_ms": float(st3["last_shadow_ms"])}
            except Exception as e:
                own["error"] = repr(e)
            finally:
                BVH = "reference"
            if aux is not None:
                aux["own_trees_sah"] = own
            else:
                aux = {"own_trees_sah": own}
        if not args.no_cpu_baseline:
            try:
                r = reference_run(args.workload, 8, 1, budget_s=25.0)
                cpu = {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            except Exception as e:
                cpu = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % e}

    if rank == 0:
        working_set = (flat["triangles"].shape[0] * 128 + flat["mesh_nodes"].shape[0] * 32 + n_px * (40 + 20 + 16 + 48)) / 1e6
        line = dict(base)
        line.update({
            "value": value, "ms_per_step": ms_max / args.steps,
            "config": {"workload": args.workload, "resolution": [W, H], "triangles": int(flat["triangles"].shape[0]),
                       "instances": int(flat["instances"].shape[0]), "max_depth": MAX_DEPTH, "light_samples": [1, 1],
                       "bvh": {"reference": "reference trees", "sah": "optional SAH builder (leaf <= 4)",
                               "lbvh": "optional GPU linear-BVH builder (leaf <= 4)"}[BVH],
                       "sharding": ("%d row band(s) x %d sample stream(s)" % (bands, streams)) if world > 1 else "single GPU",
                       "reduce": args.reduce if world > 1 else None,
                       "l2": "no flush: per-pass working set %.0f MB (path state + queues + accumulator + scene) > 126 MB L2"
                             % working_set,
                       "spp_per_s": spp_per_s},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "step": "set_scene + set_camera + reset + render(%d passes) + resolve to pinned host buffers" % RPP_E2E,
                    "frames": e2e_frames},
            "gpu_launches": launches,
            "stage_ms_per_pass": {"k_trace_paths": trace_ms, "k_shade": shade_ms, "k_trace_shadow": shadow_ms},
            # top level = the kernel with the largest share of the step (the dominant one); both traversal kernels below
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "kernels": {k: dict(v, frac=(v["achieved"] / peak) if v["achieved"] else None) for k, v in kernels.items()},
                         "algorithmic_bytes_per_segment": bytes_per_seg, "nodes_per_segment": nodes_per_seg,
                         "triangles_per_segment": tris_per_seg, "shadow_rays_per_segment": shadow_per_seg,
                         "nodes_per_shadow_ray": sh_nodes, "triangles_per_shadow_ray": sh_tris,
                         "segments_per_launch": n_px, "batch_lane_utilisation": lane_util},
            "cpu_baseline": cpu, "aux": aux,
        })
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
