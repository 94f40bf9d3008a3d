"""ctypes binding of oracle/_ref/liboracle.so (oracle/rz_oracle.c): the CPU restatement of the hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm -- never by rayzath_b200/. Also wraps oracle/_ref/rz_ref_tool (the reference's own CPU engine compiled in
place) for the checks that need the real thing.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from rayzath_b200 import capi  # noqa: E402  (struct layouts only)

LIB_PATH = os.path.join(HERE, "_ref", "liboracle.so")
REF_TOOL = os.path.join(HERE, "_ref", "rz_ref_tool")
ORDER_CPU, ORDER_CUDA = 0, 1
MINMAX_SELECT, MINMAX_FMINF = 0, 1

_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(LIB_PATH)
        P = C.c_void_p
        l.rzo_trace_closest.argtypes = [P, P, P, P, C.c_uint32, C.c_int, C.c_int, P, P]
        l.rzo_trace_closest.restype = None
        l.rzo_trace_any.argtypes = [P, P, P, P, C.c_uint32, C.c_int, P]
        l.rzo_trace_any.restype = None
        l.rzo_camera_rays.argtypes = [P, P, P, P]
        l.rzo_camera_rays.restype = None
        l.rzo_tonemap.argtypes = [P, C.c_uint32, C.c_float, C.c_float, P]
        l.rzo_tonemap.restype = None
        l.rzo_threads.restype = C.c_int
        _lib = l
    return _lib


class Scene:
    """Holds a flattened scene (dict of arrays) as an rzb_scene for the oracle."""

    def __init__(self, flat):
        s = capi.SceneStruct()
        self._keep = []

        def put(field, count_field, name, dtype):
            a = np.ascontiguousarray(flat[name]).reshape(-1)
            if a.dtype != dtype:
                a = a.view(np.uint8).view(dtype)
            self._keep.append(a)
            setattr(s, field, a.ctypes.data if a.size else None)
            if count_field:
                setattr(s, count_field, a.shape[0])

        put("mesh_nodes", "mesh_node_count", "mesh_nodes", capi.node_dtype)
        put("triangles", "triangle_count", "triangles", capi.triangle_dtype)
        put("tri_host_index", None, "tri_host_index", np.dtype(np.uint32))
        put("meshes", "mesh_count", "meshes", capi.mesh_dtype)
        put("instance_nodes", "instance_node_count", "instance_nodes", capi.node_dtype)
        put("instances", "instance_count", "instances", capi.instance_dtype)
        self.struct = s


def _rays(origins, directions, near_far):
    o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
    nf = np.ascontiguousarray(near_far, dtype=np.float32).reshape(-1, 2)
    return o, d, nf


def trace_closest(scene: Scene, origins, directions, near_far, order=ORDER_CUDA, minmax=MINMAX_FMINF, stats=False):
    o, d, nf = _rays(origins, directions, near_far)
    hits = np.zeros(o.shape[0], dtype=capi.hit_dtype)
    st = np.zeros(1, dtype=capi.trace_stats_dtype)
    lib().rzo_trace_closest(C.addressof(scene.struct), o.ctypes.data, d.ctypes.data, nf.ctypes.data, o.shape[0],
                            order, minmax, hits.ctypes.data, st.ctypes.data)
    return (hits, st[0]) if stats else hits


def trace_any(scene: Scene, origins, directions, near_far, minmax=MINMAX_FMINF):
    o, d, nf = _rays(origins, directions, near_far)
    masks = np.zeros((o.shape[0], 4), dtype=np.float32)
    lib().rzo_trace_any(C.addressof(scene.struct), o.ctypes.data, d.ctypes.data, nf.ctypes.data, o.shape[0], minmax,
                        masks.ctypes.data)
    return masks


def camera_rays(camera: np.ndarray):
    cam = np.ascontiguousarray(camera).view(capi.camera_dtype).reshape(-1)[:1]
    n = int(cam[0]["width"]) * int(cam[0]["height"])
    o, d, nf = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 2), np.float32)
    lib().rzo_camera_rays(cam.ctypes.data, o.ctypes.data, d.ctypes.data, nf.ctypes.data)
    return o, d, nf


def tonemap(accum: np.ndarray, aperture: float, exposure_time: float) -> np.ndarray:
    a = np.ascontiguousarray(accum, dtype=np.float32).reshape(-1, 4)
    out = np.zeros((a.shape[0], 4), dtype=np.uint8)
    lib().rzo_tonemap(a.ctypes.data, a.shape[0], aperture, exposure_time, out.ctypes.data)
    return out.reshape(accum.shape[:-1] + (4,))


def threads() -> int:
    return int(lib().rzo_threads())


def have_ref_tool() -> bool:
    return os.path.exists(REF_TOOL) and os.access(REF_TOOL, os.X_OK)


def ref_tool(*args, timeout=600.0, attempts=3):
    """Run oracle/_ref/rz_ref_tool; returns the JSON object of its last stdout line.
    The reference CPU engine's worker-thread gates occasionally deadlock (seen here as a render that sleeps forever
    at 0 % CPU), so every call runs under a timeout and is retried."""
    last = None
    for _ in range(attempts):
        try:
            r = subprocess.run([REF_TOOL, *map(str, args)], capture_output=True, text=True, timeout=timeout)
        except subprocess.TimeoutExpired as e:
            last = e
            continue
        if r.returncode != 0:
            raise RuntimeError("rz_ref_tool %s failed (%d): %s" % (args[0], r.returncode, r.stderr[-800:]))
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else {}
    raise RuntimeError("rz_ref_tool %s timed out %d times (%s)" % (args[0], attempts, last))
