#!/usr/bin/env python3
"""bench.py -- the headline measurement of the B200 render path (BASELINE.json: Mrays/s at 1080p).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

Default workload: heightfield_1m_1080p -- the 1M-triangle scene (texture / normal / roughness maps, depth of field) that
north_star's target is stated on; materials_1080p (BASELINE configs[1]) rides along under "aux".
A *step* is one pass of the hot path over one frame: one path segment per pixel (closest-hit trace, shade + NEE, ray
ordering, shadow trace), W*H rays -- the reference's own unit (cuda_render_kernel.cu:122-129: ray_count += W*H per pass).
Before anything is timed the renderer runs 2 x max-depth untimed passes (pre-roll): the first passes after a reset trace
only coherent camera rays. One JSON line is printed by rank 0:
  value      Mrays/s, whole job, scene and path state resident in HBM, K passes timed on the device (CUDA events on
             the launching stream), barrier + synchronize on both sides, max over ranks
  stage_ms_per_pass  exclusive device time per launch of the three kernels of a pass: the timed passes replayed with
             RZB_FLAG_SERIAL_STAGES (in the headline region a pass's shadow kernel shares the GPU with the next pass's
             closest-hit kernel on a second stream, which is worth ~2 %); the roofline figures use these times
  e2e        the same metric through the reference-facing boundary with HOST buffers: every e2e step is one
             Engine::renderWorld-equivalent frame = rzb_set_scene (host arrays -> device) + rzb_set_camera + rzb_reset
             + rzb_render(rpp passes) + rzb_resolve (tone map, RGBA8 + depth -> pinned host buffers)
  e2e_dropin the reference's OWN headless runner (Application/headless.cpp + json_loader + World, compiled in place) on this
             engine (rayzath_b200/host/_build/rz_b200_headless), same workload: the rps of its own report.txt
  roofline   both roofs of the dominant kernel: "issue" (warp instructions per launch from the committed ncu summary /
             live kernel time against SMs x 4 x SM clock; lane_frac = x active threads per instruction / 32) -- the one
             that binds -- and the HBM figure from ALGORITHMIC bytes, a throughput normalisation
  cpu_baseline  the reference's own CPU engine (oracle/_ref/rz_ref_tool, built from /root/reference) on this box's
             host cores on a bounded sample of the same workload
  config     the workload alone, identical in both arms; "run" holds what is specific to this arm (sharding, reduce ...)
`--impl reference` times that CPU engine alone and prints the same line with "impl": "reference".
N > 1 (torchrun): sample streams are sharded over the ranks (same frame, disjoint RNG streams, weak scaling), no
collective while rendering; the accumulators are combined once, inside the timed region, by the sliced NVLink resolve
(parallel.SlicedResolve: one kernel per rank, flag barriers in peer memory, no NCCL; NCCL reduce with --reduce nccl).
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
    # BASELINE.json configs[2] geometry: the 1M-triangle scene (texture / normal / roughness maps, depth of field) that
    # north_star's target (">= 1 Grays/s per B200 on a 1M-triangle scene") is stated on -- the headline workload
    "heightfield_1m_1080p": ("heightfield_1m", dict(resolution=(1920, 1080))),
    # configs[1]: materials + lights scene (23k triangles), kept as a secondary line in "aux"
    "materials_1080p": ("materials", dict(resolution=(1920, 1080), res=64)),
    "instancing_10m_1080p": ("instancing_10m", dict(resolution=(1920, 1080))),
    "cornell_512": ("cornell", dict(resolution=(512, 512))),
    # configs[4]: the 1M-triangle scene at 3840x2160 (use with --split hybrid on 8 GPUs: 4 row bands x 2 sample streams)
    "heightfield_1m_4k": ("heightfield_1m", dict(resolution=(3840, 2160))),
}
# reference-layout byte constants for the algorithmic traffic figure (SURVEY.md 8d / DESIGN.md)
B_NODE, B_TRI, B_STATE, B_HIT = 48, 144, 57, 24
B_SHADOW_REC, B_SHADOW_ACC = 48, 12  # shadow queue record read, accumulator update
MAX_DEPTH = 16          # Tracing::maxDepth default, engine_parts.hpp:101
PREROLL = 2 * MAX_DEPTH  # untimed passes after a reset before anything is timed: paths of all depths are in flight
DEFAULT_WORKLOAD = "heightfield_1m_1080p"
SECONDARY_WORKLOAD = "materials_1080p"
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


# ---------------------------------------------------------------------------------------------- drop-in leg
def dropin_run(workload, budget_s=20.0):
    """e2e through the REFERENCE-FACING host: the reference's own headless runner (Application/headless.cpp, compiled
    in place) with RayZath::Cuda::Engine implemented by rayzath_b200/host/cuda_engine_b200.cpp -- scene file loaded by
    the reference's json_loader, World flattened by world_flatten.hpp, frames through Engine::renderWorld. The number
    is the reference's own report line (`traced N rays (N rps)`, headless.cpp:317-320). None when the binary was not
    built (it needs /root/reference at build time and travels to the GPU box prebuilt)."""
    binary = os.path.join(ROOT, "rayzath_b200", "host", "_build", "rz_b200_headless")
    if not os.path.exists(binary):
        return None
    import re
    from rayzath_b200 import scenes
    name, kw = WORKLOADS[workload]
    tmp = tempfile.mkdtemp(prefix="rzb_bench_dropin_")
    try:
        w = scenes.CONFIGS[name](**kw)
        w.save_reference(tmp, "scene")
        json.dump({"tasks": [{"scene path": "scene.json", "engine": ["CUDAGPU"], "rpp": 1000000, "timeout": budget_s,
                              "max depth": MAX_DEPTH}]}, open(os.path.join(tmp, "tasks.json"), "w"))
        os.makedirs(os.path.join(tmp, "report"))
        t0 = time.perf_counter()
        r = subprocess.run([binary, "--headless", "tasks.json", "report", "-r"], cwd=tmp, capture_output=True, text=True,
                           timeout=budget_s * 6 + 240, env=dict(os.environ, RZB200_SEED="20261018"))
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": (r.stderr or r.stdout)[-400:]}
        text = ""
        for dp, _, fs in os.walk(os.path.join(tmp, "report")):
            for f in fs:
                if f == "report.txt":
                    text = open(os.path.join(dp, f)).read()
        m = re.search(r"traced\s+([0-9.]+)\s*([kMGT]?)\s*rays\s*\(([0-9.]+)\s*([kMGT]?)\s*rps\)", text)
        if not m:
            return {"error": "no report line: " + text[-300:]}
        mult = {"": 1.0, "k": 1e3, "M": 1e6, "G": 1e9, "T": 1e12}
        return {"value": float(m.group(3)) * mult[m.group(4)] / 1e6, "unit": "Mrays/s",
                "rays": float(m.group(1)) * mult[m.group(2)], "wall_s": wall,
                "step": "rz_b200_headless (reference headless.cpp + json_loader + World flatten + renderWorld on the B200 path), "
                        "task: engine CUDAGPU, timeout %.0f s, max depth %d; number = the reference's own report.txt rps" % (budget_s, MAX_DEPTH)}
    except Exception as e:
        return {"error": repr(e)}
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)


# ---------------------------------------------------------------------------------------------- our arm
def pin(arrays):
    """Page-lock the host arrays the boundary reads from (cudaHostRegister through torch)."""
    import torch
    rt = torch.cuda.cudart()
    for a in arrays:
        if a.size:
            rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)


def describe_config(workload, world_obj, n_px_state):
    """`config` of the JSON line: the workload only, identical (keys AND values) in both arms so that the driver's
    same-config check compares like with like; arm-specific facts (sharding, reduce, spp/s) are under "run"."""
    tris = int(sum(m.tris.shape[0] for m in world_obj.meshes))
    nodes_est = 0.33 * tris
    W, H = (int(x) for x in world_obj.cameras[0].resolution)
    working_set = (tris * 128 + nodes_est * 32 + n_px_state * (40 + 20 + 16 + 48)) / 1e6
    return {"workload": workload, "resolution": [W, H], "triangles": tris, "instances": len(world_obj.instances),
            "max_depth": MAX_DEPTH, "light_samples": [1, 1], "bvh": "reference trees",
            "preroll_passes": PREROLL,
            "l2": "no flush: per-pass working set ~%.0f MB (path state + queues + accumulator + scene) > 126 MB L2"
                  % working_set}


def load_profile(key):
    prof_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof_path):
        try:
            return json.load(open(prof_path)).get(key, {})
        except Exception:
            pass
    return {}


def work_and_roofline(ctx, capi, steps, warm, seed, n_px, stage_ms, profile_key, peak, clocks, sm_count):
    """Replays the timed passes with the counting kernels (the RNG is counter-based on (seed, slot, pass): after a
    reset the same pass indices retrace the same rays) and turns the counters into the two roofs of DESIGN.md:
      hbm    ALGORITHMIC bytes in reference-layout constants / measured kernel time / measured HBM peak.
             A throughput normalisation: the hot scene is L2-resident, `traffic` (ncu DRAM bytes) is far smaller.
      issue  the roof that binds: warp instructions per launch (ncu smsp__inst_executed of the same command,
             profiles/ncu_summary.json) / measured kernel time, against SMs x 4 schedulers x SM clock; `lane_frac`
             multiplies by the active threads per instruction / 32 (SIMT efficiency)."""
    trace_ms, shade_ms, shadow_ms = stage_ms
    ctx.reset()
    ctx.render(warm)
    ctx.set_config(1, 1, MAX_DEPTH, capi.FLAG_COUNT_WORK, seed)
    ctx.render(steps)
    wc = ctx.work_counters()
    ctx.set_config(1, 1, MAX_DEPTH, capi.FLAG_NONE, seed)
    seg = max(int(wc["segments"]), 1)
    nodes_per_seg = float(wc["closest_top_nodes"] + wc["closest_mesh_nodes"]) / seg
    tris_per_seg = float(wc["closest_triangles"]) / seg
    shadow_per_seg = float(wc["shadow_rays"]) / seg
    lane_util = {"closest": float(wc["closest_lane_work"]) / max(int(wc["closest_batch_work"]), 1),
                 "shadow": float(wc["shadow_lane_work"]) / max(int(wc["shadow_batch_work"]), 1),
                 "dropped_non_finite_samples": int(wc["invalid_rays"])}
    bytes_per_seg = B_STATE + nodes_per_seg * B_NODE + tris_per_seg * B_TRI + B_HIT
    n_shadow = max(int(wc["shadow_rays"]), 1)
    sh_nodes = float(wc["shadow_top_nodes"] + wc["shadow_mesh_nodes"]) / n_shadow
    sh_tris = float(wc["shadow_triangles"]) / n_shadow
    bytes_per_shadow_ray = B_SHADOW_REC + sh_nodes * B_NODE + sh_tris * B_TRI + B_SHADOW_ACC
    prof = load_profile(profile_key)
    f_mhz = float((clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)
    warp_peak = sm_count * 4 * f_mhz * 1e6  # warp instructions per second the chip can issue at the sampled clock

    def issue(kernel, ms):
        wi, tpi = prof.get(kernel + "_warp_inst_per_launch"), prof.get(kernel + "_threads_per_inst")
        if not wi or not ms:
            return None
        rate = float(wi) / (ms * 1e-3)
        out = {"warp_inst_per_launch": float(wi), "threads_per_inst": tpi, "warp_inst_per_s": rate,
               "peak_warp_inst_per_s": warp_peak, "frac": rate / warp_peak,
               "source": "ncu smsp__inst_executed.sum of the same workload (profiles/ncu_summary.json) / live kernel time; "
                         "peak = %d SMs x 4 x %.0f MHz" % (sm_count, f_mhz)}
        if tpi:
            out["thread_inst_per_s"] = rate * float(tpi)
            out["lane_frac"] = rate * float(tpi) / (warp_peak * 32.0)
        return out

    kernels = {
        "k_trace_paths": {"ms": trace_ms, "units_per_launch": n_px, "algorithmic_bytes_per_unit": bytes_per_seg,
                          "achieved": bytes_per_seg * n_px / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else None,
                          "traffic": prof.get("k_trace_paths_dram_bytes_per_launch"), "issue": issue("k_trace_paths", trace_ms)},
        "k_trace_shadow": {"ms": shadow_ms, "units_per_launch": shadow_per_seg * n_px,
                           "algorithmic_bytes_per_unit": bytes_per_shadow_ray,
                           "achieved": bytes_per_shadow_ray * shadow_per_seg * n_px / (shadow_ms * 1e-3) / 1e9 if shadow_ms > 0 else None,
                           "traffic": prof.get("k_trace_shadow_dram_bytes_per_launch"), "issue": issue("k_trace_shadow", shadow_ms)},
    }
    dominant = max(kernels, key=lambda k: kernels[k]["ms"])
    achieved = kernels[dominant]["achieved"]
    return {
        # `bound` names the roof that binds the dominant kernel: instruction issue under low SIMT efficiency (ncu: DRAM
        # ~1 % of peak, issue slots 65-70 %). achieved / peak / frac keep the contract's HBM form in algorithmic bytes.
        "bound": "issue", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None, "traffic": kernels[dominant]["traffic"],
        "hbm_note": "achieved = algorithmic bytes (reference layout: 48 B nodes, 144 B triangles) / kernel time; the hot "
                    "scene is L2-resident so `traffic` (ncu DRAM bytes per launch) is far below it -- a normalisation, not the binding roof",
        "issue": kernels[dominant]["issue"],
        "kernels": {k: dict(v, frac=(v["achieved"] / peak) if v["achieved"] else None) for k, v in kernels.items()},
        "algorithmic_bytes_per_segment": bytes_per_seg, "nodes_per_segment": nodes_per_seg,
        "triangles_per_segment": tris_per_seg, "shadow_rays_per_segment": shadow_per_seg,
        "nodes_per_shadow_ray": sh_nodes, "triangles_per_shadow_ray": sh_tris,
        "segments_per_launch": n_px, "batch_lane_utilisation": lane_util,
    }


def timed_passes(ctx, torch, stream, steps, warm):
    """reset + `warm` untimed passes + `steps` passes between CUDA events on the launching stream"""
    ctx.reset()
    ctx.render(warm)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(stream)
    ctx.render(steps)
    a1.record(stream)
    a1.synchronize()
    st = ctx.render_stats()
    return a0.elapsed_time(a1), (float(st["last_trace_ms"]), float(st["last_shade_ms"]), float(st["last_shadow_ms"]))


STAGE_NOTE = ("exclusive kernel times: replay of the same passes with every kernel in stream order (RZB_FLAG_SERIAL_STAGES); in the "
              "headline region the shadow kernel of pass p runs on a second stream beside k_trace_paths of pass p + 1")


def serial_stage_times(ctx, capi, torch, stream, steps, warm, seed):
    """Per-kernel device time per launch. The headline region overlaps a pass's shadow kernel with the next pass's closest-hit
    kernel, so a stage's events there bracket two kernels sharing the GPU; the same passes (counter-based RNG: same rays) are
    therefore replayed with RZB_FLAG_SERIAL_STAGES and the CUDA events around every stage of (up to 256 of) them are read.
    Returns (ms of the serial replay, (trace, shade, shadow) ms per launch)."""
    ctx.set_config(1, 1, MAX_DEPTH, capi.FLAG_SERIAL_STAGES, seed)
    out = timed_passes(ctx, torch, stream, steps, warm)
    ctx.set_config(1, 1, MAX_DEPTH, capi.FLAG_NONE, seed)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reduce", default="ipc", choices=["ipc", "nccl"])
    ap.add_argument("--split", default="samples", choices=["samples", "tiles", "hybrid"],
                    help="N>1: samples = every rank renders the whole frame with its own RNG stream (weak scaling, default); "
                         "tiles = rank r renders row band r of N (strong scaling); hybrid = N/2 bands x 2 streams")
    ap.add_argument("--bvh", default="reference", choices=["reference", "sah", "sah4", "lbvh"],
                    help="triangle trees: the reference's own (parity mode, default), the optional SAH builder (host) or the "
                         "optional linear-BVH builder (GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true")
    ap.add_argument("--no-dropin", action="store_true", help="skip the e2e_dropin leg (the C++ drop-in's headless run)")
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
        wobj = build_world(args.workload)
        W0, H0 = (int(x) for x in wobj.cameras[0].resolution)
        line = dict(base)
        line.update({"impl": "reference", "value": r["value"], "ms_per_step": r["ms_per_step"],
                     "config": describe_config(args.workload, wobj, ((W0 + 15) // 16) * ((H0 + 15) // 16) * 256),
                     "run": {"sharding": "host threads", "engine": "reference CPU engine (oracle/_ref/rz_ref_tool render)"},
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
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
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

    resolver = None
    if world > 1 and args.reduce == "ipc":
        resolver = parallel.SlicedResolve(ctx, rgba_host, depth_host)

    def combine():
        """the one exchange step of the N>1 path: every rank sums and tone-maps its slice of the frame over NVLink
        (parallel.SlicedResolve), or an NCCL reduce onto rank 0 with --reduce nccl"""
        if world == 1:
            return
        if resolver is not None:
            resolver()
        else:
            parallel.reduce_accum(ctx.accum_tensor().clone(), dst=0)  # a copy: rendering continues on the original

    # ---- pre-roll + warm-up (untimed), then the timed region: K passes (+ the exchange step), device-timed on the
    # launching stream. The pre-roll (2 x max depth passes) puts paths of every depth in flight: the first passes after
    # a reset trace only coherent camera rays and would flatter the number.
    # The W warm-up steps run IMMEDIATELY in front of the timed region: whatever leaves the GPU idle for milliseconds (NVML
    # set-up of the clock sampler, reading an accumulator back) happens before them, and the completed-paths figure in front
    # of the timed region comes from the device-side reduction (rzb_mean_samples, 8 bytes back). With the driver's 20-step
    # window (24 ms) an idle gap in front of it cost 4 % (1694 against 1765 Mrays/s for 512 steps).
    warm = PREROLL + args.warmup
    ctx.render(PREROLL)
    combine()
    barrier()
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    frame_share = float(n_px) / float(W * H)  # this rank's pixels / frame pixels (1 unless the frame is split into bands)
    ctx.render(args.warmup)
    combine()
    alpha0 = ctx.mean_samples() * frame_share
    launches0 = int(ctx.render_stats()["kernel_launches"])
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
    st = ctx.render_stats()
    launches = int(st["kernel_launches"]) - launches0
    if resolver is not None:
        resolver.wait()
    t_all = torch.tensor([ms, (ctx.mean_samples() * frame_share - alpha0), float(args.steps) * n_px],
                         dtype=torch.float64, device="cuda")
    ms_max, spp_sum, rays_sum = float(t_all[0].item()), float(t_all[1].item()), float(t_all[2].item())
    if dist is not None:
        t_max = t_all.clone()
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_all, op=dist.ReduceOp.SUM)
        ms_max, spp_sum, rays_sum = float(t_max[0].item()), float(t_all[1].item()), float(t_all[2].item())
    value = rays_sum / (ms_max * 1e-3) / 1e6
    spp_per_s = spp_sum / (ms_max * 1e-3)  # completed camera paths per pixel per second, whole job

    # ---- per-kernel device time per launch (serial replay of the timed passes, see serial_stage_times)
    serial_ms, stage_ms = serial_stage_times(ctx, capi, torch, stream, args.steps, warm, seed)

    # ---- both roofs of the dominant kernel (replay of the same passes with the counting kernels)
    roofline = work_and_roofline(ctx, capi, args.steps, warm, seed, n_px, stage_ms,
                                 args.workload if BVH == "reference" else args.workload + "_" + BVH, peak, clocks, sm_count)
    roofline["peak_source"] = peak_src

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
            if resolver is not None:
                resolver()
                resolver.wait()
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

    # ---- secondary workload, own-tree variants, the drop-in's headless run and the CPU baseline: rank 0, N=1 only
    aux = None
    cpu = None
    dropin = None
    if rank == 0 and world == 1:
        n2 = min(args.steps, 256)
        secondary = SECONDARY_WORKLOAD if args.workload != SECONDARY_WORKLOAD else DEFAULT_WORKLOAD
        if not args.no_aux:
            aux = {}
            try:
                w2 = build_world(secondary)
                f2 = w2.flatten()
                ctx.set_scene(f2)
                ctx.set_camera(w2.camera_struct())
                ms2, _ = timed_passes(ctx, torch, stream, n2, warm)
                _, stage2 = serial_stage_times(ctx, capi, torch, stream, n2, warm, seed)
                aux[secondary] = {"triangles": int(f2["triangles"].shape[0]), "instances": int(f2["instances"].shape[0]),
                                  "value": n2 * n_px / (ms2 * 1e-3) / 1e6, "unit": "Mrays/s", "steps": n2,
                                  "stage_ms_per_pass": {"k_trace_paths": stage2[0], "k_shade": stage2[1], "k_trace_shadow": stage2[2]},
                                  "roofline": work_and_roofline(ctx, capi, n2, warm, seed, n_px, stage2, secondary, peak, clocks, sm_count)}
            except Exception as e:  # the headline line must still be printed
                aux[secondary] = {"error": repr(e)}
            # the same workloads on the optional SAH trees (RZB_SCENE_OWN_TREES; records equal except exact ties)
            if BVH == "reference":
                own = {}
                try:
                    BVH = "sah"
                    for wl in (args.workload, secondary):
                        w3 = build_world(wl)
                        ctx.set_scene(w3.flatten())
                        ctx.set_camera(w3.camera_struct())
                        ms3, _ = timed_passes(ctx, torch, stream, n2, warm)
                        _, stage3 = serial_stage_times(ctx, capi, torch, stream, n2, warm, seed)
                        own[wl] = {"value": n2 * n_px / (ms3 * 1e-3) / 1e6, "unit": "Mrays/s", "steps": n2,
                                   "stage_ms_per_pass": {"k_trace_paths": stage3[0], "k_shade": stage3[1], "k_trace_shadow": stage3[2]}}
                except Exception as e:
                    own["error"] = repr(e)
                finally:
                    BVH = "reference"
                aux["own_trees_sah"] = own
        if not args.no_dropin:
            dropin = dropin_run(args.workload)
        if not args.no_cpu_baseline:
            try:
                r = reference_run(args.workload, 8, 1, budget_s=25.0)
                cpu = {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            except Exception as e:
                cpu = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % e}

    if rank == 0:
        line = dict(base)
        line.update({
            "value": value, "ms_per_step": ms_max / args.steps,
            "config": describe_config(args.workload, world_obj, ((W + 15) // 16) * ((H + 15) // 16) * 256),
            "run": {"bvh": {"reference": "reference trees", "sah": "optional SAH builder (leaf <= 4)",
                            "sah4": "optional SAH builder (leaf <= 4), collapsed to 4-ary trees",
                            "lbvh": "optional GPU linear-BVH builder (leaf <= 4)"}[BVH],
                    "sharding": ("%d row band(s) x %d sample stream(s)" % (bands, streams)) if world > 1 else "single GPU",
                    "reduce": args.reduce if world > 1 else None, "spp_per_s": spp_per_s,
                    "resolve_ms": resolver.last_ms() if resolver is not None else None},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "step": "set_scene + set_camera + reset + render(%d passes) + resolve to pinned host buffers" % RPP_E2E,
                    "frames": e2e_frames},
            "e2e_dropin": dropin,
            "gpu_launches": launches,
            "stage_ms_per_pass": {"k_trace_paths": stage_ms[0], "k_shade": stage_ms[1], "k_trace_shadow": stage_ms[2],
                                  "serial_ms_per_step": serial_ms / args.steps, "note": STAGE_NOTE},
            "roofline": roofline,
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
