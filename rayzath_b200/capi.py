"""ctypes binding of the C ABI in include/rzb200.h (rayzath_b200/librzb200.so).

This is the Python host of the B200 render path: plumbing for tests and bench.py. The product is the
shared library; there is no Python or CPU fallback here. Importing this module only loads the
library; creating a `Context` needs a CUDA device and raises `RzbError` without one.

Struct layouts are mirrored as numpy dtypes (checked against the static_asserts of csrc/rzb_api.cu
by tests/test_abi.py).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

LIB_PATH = os.environ.get("RZB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "librzb200.so")

NO_INDEX = 0xFFFFFFFF
FLAG_NONE = 0
FLAG_CPU_SEMANTICS = 1
FLAG_COUNT_WORK = 2
FLAG_TEMPORAL_REPROJECTION = 4
FLAG_SERIAL_STAGES = 8  # every kernel of a pass in stream order: exclusive per-stage times (rzb200.h)
SCENE_REFERENCE_TREES, SCENE_OWN_TREES, SCENE_KEEP_GEOMETRY, SCENE_WIDE_TREES = 0, 1, 2, 4
MAP_RGBA8, MAP_R8, MAP_R32F = 0, 1, 2
FILTER_POINT, FILTER_LINEAR = 0, 1
ADDRESS_WRAP, ADDRESS_CLAMP, ADDRESS_MIRROR, ADDRESS_BORDER = 0, 1, 2, 3

f4, u4, u8 = np.float32, np.uint32, np.uint64

node_dtype = np.dtype([("bb_min", f4, 3), ("bb_max", f4, 3), ("begin", u4), ("type_count", u4)])
triangle_dtype = np.dtype([("v", f4, (3, 3)), ("n", f4, (3, 3)), ("face_normal", f4, 3), ("uv", f4, (3, 2)),
                           ("material_slot", u4)])
mesh_dtype = np.dtype([("node_offset", u4), ("node_count", u4), ("tri_offset", u4), ("tri_count", u4)])
instance_dtype = np.dtype([("position", f4, 3), ("scale", f4, 3), ("axis_x", f4, 3), ("axis_y", f4, 3),
                           ("axis_z", f4, 3), ("bb_min", f4, 3), ("bb_max", f4, 3), ("mesh", u4),
                           ("material_offset", u4), ("material_count", u4), ("host_index", u4)])
material_dtype = np.dtype([("color", f4, 4), ("metalness", f4), ("roughness", f4), ("emission", f4), ("ior", f4),
                           ("scattering", f4), ("texture", u4), ("normal_map", u4), ("metalness_map", u4),
                           ("roughness_map", u4), ("emission_map", u4), ("_pad", u4, 2)])
map_dtype = np.dtype([("format", u4), ("width", u4), ("height", u4), ("filter", u4), ("address", u4),
                      ("scale", f4, 2), ("rotation", f4), ("translation", f4, 2), ("_pad", u4), ("pixels", u8)],
                     align=True)
direct_light_dtype = np.dtype([("direction", f4, 3), ("angular_size", f4), ("color", f4, 3), ("emission", f4)])
spot_light_dtype = np.dtype([("position", f4, 3), ("size", f4), ("direction", f4, 3), ("beam_angle", f4),
                             ("color", f4, 3), ("emission", f4)])
camera_dtype = np.dtype([("width", u4), ("height", u4), ("position", f4, 3), ("axis_x", f4, 3), ("axis_y", f4, 3),
                         ("axis_z", f4, 3), ("fov", f4), ("near_far", f4, 2), ("focal_distance", f4),
                         ("aperture", f4), ("exposure_time", f4), ("temporal_blend", f4), ("raycast_pixel", u4, 2)])
config_dtype = np.dtype([("spot_light_samples", u4), ("direct_light_samples", u4), ("max_depth", u4), ("flags", u4),
                         ("seed", u8)])
hit_dtype = np.dtype([("instance", u4), ("triangle", u4), ("t", f4), ("b1", f4), ("b2", f4), ("external", u4)])
trace_stats_dtype = np.dtype([("rays", u8), ("top_nodes", u8), ("instances_entered", u8), ("mesh_nodes", u8),
                              ("triangles", u8)])
render_stats_dtype = np.dtype([("passes", u8), ("ray_count", u8), ("shadow_rays", u8), ("kernel_launches", u8),
                               ("last_render_ms", f4), ("last_trace_ms", f4), ("last_shade_ms", f4),
                               ("last_shadow_ms", f4), ("last_sort_ms", f4), ("last_exchange_ms", f4)])

work_counters_dtype = np.dtype([(n, u8) for n in (
    "closest_top_nodes", "closest_instances", "closest_mesh_nodes", "closest_triangles",
    "shadow_top_nodes", "shadow_instances", "shadow_mesh_nodes", "shadow_triangles", "shadow_rays", "segments",
    "invalid_rays", "closest_lane_work", "closest_batch_work", "shadow_lane_work", "shadow_batch_work")])

EXPECTED_SIZES = {
    "rzb_node": (node_dtype, 32), "rzb_triangle": (triangle_dtype, 112), "rzb_mesh": (mesh_dtype, 16),
    "rzb_instance": (instance_dtype, 100), "rzb_material": (material_dtype, 64), "rzb_map": (map_dtype, 56),
    "rzb_direct_light": (direct_light_dtype, 32), "rzb_spot_light": (spot_light_dtype, 48),
    "rzb_camera": (camera_dtype, 92), "rzb_config": (config_dtype, 24), "rzb_hit": (hit_dtype, 24),
}


class SceneStruct(C.Structure):
    """rzb_scene"""
    _fields_ = [
        ("mesh_nodes", C.c_void_p), ("mesh_node_count", C.c_uint32),
        ("triangles", C.c_void_p), ("triangle_count", C.c_uint32),
        ("tri_host_index", C.c_void_p),
        ("meshes", C.c_void_p), ("mesh_count", C.c_uint32),
        ("instance_nodes", C.c_void_p), ("instance_node_count", C.c_uint32),
        ("instances", C.c_void_p), ("instance_count", C.c_uint32),
        ("instance_materials", C.c_void_p), ("instance_material_count", C.c_uint32),
        ("materials", C.c_void_p), ("material_count", C.c_uint32),
        ("maps", C.c_void_p), ("map_count", C.c_uint32),
        ("direct_lights", C.c_void_p), ("direct_light_count", C.c_uint32),
        ("spot_lights", C.c_void_p), ("spot_light_count", C.c_uint32),
        ("world_material", C.c_uint8 * 64),
        ("default_material", C.c_uint32),
        ("flags", C.c_uint32),
    ]


# every symbol include/rzb200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "rzb_abi_version": (C.c_int, []),
    "rzb_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "rzb_destroy": (None, [_P]),
    "rzb_last_error": (C.c_char_p, [_P]),
    "rzb_set_stream": (C.c_int, [_P, _P, C.c_int]),
    "rzb_set_scene": (C.c_int, [_P, C.POINTER(SceneStruct)]),
    "rzb_set_camera": (C.c_int, [_P, _P]),
    "rzb_set_config": (C.c_int, [_P, _P]),
    "rzb_set_rows": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "rzb_set_row_interleave": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "rzb_reset": (C.c_int, [_P]),
    "rzb_render": (C.c_int, [_P, C.c_uint32]),
    "rzb_resolve": (C.c_int, [_P, _P, _P, C.POINTER(C.c_uint64)]),
    "rzb_read_accum": (C.c_int, [_P, _P]),
    "rzb_mean_samples": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "rzb_accum_device_ptr": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "rzb_accum_add_device": (C.c_int, [_P, _P, C.c_size_t]),
    "rzb_resolve_peers": (C.c_int, [_P, C.POINTER(_P), C.c_uint32, _P, _P, C.POINTER(C.c_uint64)]),
    "rzb_accum_ipc_handle": (C.c_int, [_P, _P]),
    "rzb_resolve_ipc": (C.c_int, [_P, _P, C.c_uint32, _P, _P]),
    "rzb_exchange_ipc_handle": (C.c_int, [_P, _P]),
    "rzb_resolve_sliced": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P, _P, _P]),
    "rzb_resolve_sliced_wait": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "rzb_get_work_counters": (C.c_int, [_P, _P]),
    "rzb_raycast": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rzb_synchronize": (C.c_int, [_P]),
    "rzb_get_render_stats": (C.c_int, [_P, _P]),
    "rzb_timings": (C.c_int, [_P, C.c_char_p, C.c_size_t]),
    "rzb_trace_closest": (C.c_int, [_P, _P, _P, _P, C.c_uint32, _P, _P]),
    "rzb_trace_closest_device": (C.c_int, [_P, _P, _P, C.c_uint32, _P, C.POINTER(C.c_float)]),
    "rzb_trace_closest_device_counted": (C.c_int, [_P, _P, _P, C.c_uint32, _P]),
    "rzb_trace_any": (C.c_int, [_P, _P, _P, _P, C.c_uint32, _P]),
    "rzb_generate_camera_rays": (C.c_int, [_P, _P, _P, _P]),
    "rzb_resolve_async": (C.c_int, [_P, C.c_uint32, _P, _P, C.POINTER(C.c_uint64)]),
    "rzb_resolve_wait": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rzb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "rzb_host_free": (C.c_int, [_P]),
    "rzb_build_mesh_bvh": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32, _P, C.c_uint32, C.POINTER(C.c_uint32), _P]),
    "rzb_build_mesh_bvh_lbvh": (C.c_int, [C.c_int, _P, C.c_uint32, _P, C.c_uint32, C.c_uint32, _P, C.c_uint32, C.POINTER(C.c_uint32), _P,
                                          C.POINTER(C.c_float)]),
    "rzb_build_mesh_bvh_sah": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32, C.c_uint32, _P, C.c_uint32, C.POINTER(C.c_uint32), _P]),
    "rzb_build_instance_bvh": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32, C.POINTER(C.c_uint32), _P]),
    "rzb_refit_mesh_bvh": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32, _P, C.c_uint32, _P]),
    "rzb_rotation_axes": (C.c_int, [_P, C.c_int, _P]),
    "rzb_instance_bbox": (C.c_int, [_P, C.c_uint32, _P, _P, _P, _P]),
    "rzb_face_normals": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32, _P]),
}


class RzbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("rzb error %d: %s" % (code, message))
        self.code = code


_lib = None


def lib():
    """Load librzb200.so (fails loudly when it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` or "
                              "`make -C rayzath_b200/csrc`" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


class PinnedArray:
    """numpy view of page-locked host memory from rzb_host_alloc (for rzb_resolve_async); free() or use as a context."""

    def __init__(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p(0)
        rc = lib().rzb_host_alloc(nbytes, C.byref(p))
        if rc:
            raise RzbError(rc, "rzb_host_alloc failed")
        self._p = p
        self.array = np.frombuffer((C.c_uint8 * nbytes).from_address(p.value), dtype=dtype).reshape(shape)

    def free(self):
        if self._p is not None:
            self.array = None
            lib().rzb_host_free(self._p)
            self._p = None

    def __enter__(self):
        return self.array

    def __exit__(self, *exc):
        self.free()


# ---------------------------------------------------------------- host utilities (no device needed)
def build_mesh_bvh(vertices: np.ndarray, tris: np.ndarray):
    """rzb_build_mesh_bvh: the reference's triangle BVH. Returns (nodes[rzb_node], order[u32])."""
    v = _c(vertices, f4).reshape(-1, 3)
    t = _c(tris, u4).reshape(-1, 3)
    nt = t.shape[0]
    nodes = np.zeros(2 * nt + 1, dtype=node_dtype)
    order = np.zeros(nt, dtype=u4)
    count = C.c_uint32(0)
    rc = lib().rzb_build_mesh_bvh(v.ctypes.data, v.shape[0], t.ctypes.data, nt, nodes.ctypes.data, nodes.shape[0],
                                  C.byref(count), order.ctypes.data)
    if rc:
        raise RzbError(rc, "rzb_build_mesh_bvh failed")
    return nodes[:count.value].copy(), order


def build_mesh_bvh_sah(vertices: np.ndarray, tris: np.ndarray, max_leaf: int = 8):
    """rzb_build_mesh_bvh_sah: the optional SAH triangle BVH (not the reference's tree). Returns (nodes, order)."""
    v = _c(vertices, f4).reshape(-1, 3)
    t = _c(tris, u4).reshape(-1, 3)
    nt = t.shape[0]
    nodes = np.zeros(2 * nt + 1, dtype=node_dtype)
    order = np.zeros(nt, dtype=u4)
    count = C.c_uint32(0)
    rc = lib().rzb_build_mesh_bvh_sah(v.ctypes.data, v.shape[0], t.ctypes.data, nt, int(max_leaf), nodes.ctypes.data,
                                      nodes.shape[0], C.byref(count), order.ctypes.data)
    if rc:
        raise RzbError(rc, "rzb_build_mesh_bvh_sah failed")
    return nodes[:count.value].copy(), order


def refit_mesh_bvh(vertices: np.ndarray, tris: np.ndarray, nodes: np.ndarray, order: np.ndarray) -> np.ndarray:
    """rzb_refit_mesh_bvh: new boxes for an existing tree after the vertices moved (topology and order kept).
    Returns the refitted copy of `nodes`."""
    v = _c(vertices, f4).reshape(-1, 3)
    t = _c(tris, u4).reshape(-1, 3)
    out = np.ascontiguousarray(nodes, dtype=node_dtype).copy()
    o = _c(order, u4)
    rc = lib().rzb_refit_mesh_bvh(v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], out.ctypes.data, out.shape[0], o.ctypes.data)
    if rc:
        raise RzbError(rc, "rzb_refit_mesh_bvh failed")
    return out


def build_mesh_bvh_lbvh(vertices: np.ndarray, tris: np.ndarray, max_leaf: int = 4, device: int = 0, timing=False):
    """rzb_build_mesh_bvh_lbvh: the optional GPU linear-BVH builder. Returns (nodes, order[, device ms])."""
    v = _c(vertices, f4).reshape(-1, 3)
    t = _c(tris, u4).reshape(-1, 3)
    nt = t.shape[0]
    nodes = np.zeros(2 * nt + 1, dtype=node_dtype)
    order = np.zeros(nt, dtype=u4)
    count, ms = C.c_uint32(0), C.c_float(0.0)
    rc = lib().rzb_build_mesh_bvh_lbvh(int(device), v.ctypes.data, v.shape[0], t.ctypes.data, nt, int(max_leaf),
                                       nodes.ctypes.data, nodes.shape[0], C.byref(count), order.ctypes.data, C.byref(ms))
    if rc:
        raise RzbError(rc, "rzb_build_mesh_bvh_lbvh failed")
    out = (nodes[:count.value].copy(), order)
    return out + (ms.value,) if timing else out


def build_instance_bvh(boxes: np.ndarray):
    b = _c(boxes, f4).reshape(-1, 6)
    n = b.shape[0]
    nodes = np.zeros(2 * n + 2, dtype=node_dtype)
    order = np.zeros(max(n, 1), dtype=u4)
    count = C.c_uint32(0)
    rc = lib().rzb_build_instance_bvh(b.ctypes.data, n, nodes.ctypes.data, nodes.shape[0], C.byref(count),
                                      order.ctypes.data)
    if rc:
        raise RzbError(rc, "rzb_build_instance_bvh failed")
    return nodes[:count.value].copy(), order[:n]


def rotation_axes(rotation, order: int) -> np.ndarray:
    r = _c(rotation, f4).reshape(3)
    out = np.zeros(9, dtype=f4)
    rc = lib().rzb_rotation_axes(r.ctypes.data, int(order), out.ctypes.data)
    if rc:
        raise RzbError(rc, "rzb_rotation_axes failed")
    return out.reshape(3, 3)


def instance_bbox(vertices, position, scale, axes) -> np.ndarray:
    v = _c(vertices, f4).reshape(-1, 3)
    p, s, a = _c(position, f4).reshape(3), _c(scale, f4).reshape(3), _c(axes, f4).reshape(9)
    out = np.zeros(6, dtype=f4)
    rc = lib().rzb_instance_bbox(v.ctypes.data, v.shape[0], p.ctypes.data, s.ctypes.data, a.ctypes.data,
                                 out.ctypes.data)
    if rc:
        raise RzbError(rc, "rzb_instance_bbox failed")
    return out


def face_normals(vertices, tris) -> np.ndarray:
    v = _c(vertices, f4).reshape(-1, 3)
    t = _c(tris, u4).reshape(-1, 3)
    out = np.zeros((t.shape[0], 3), dtype=f4)
    rc = lib().rzb_face_normals(v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], out.ctypes.data)
    if rc:
        raise RzbError(rc, "rzb_face_normals failed")
    return out


# ---------------------------------------------------------------- context
class Context:
    """One rzb_ctx = one GPU's mirror of the world + per-camera frame state."""

    def __init__(self, device: int = 0):
        self._l = lib()
        h = _P()
        rc = self._l.rzb_create(int(device), C.byref(h))
        if rc:
            raise RzbError(rc, (self._l.rzb_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self.width = self.height = 0
        self._keep = None
        self.caller_stream = None  # the CUDA stream handed to set_stream (None = the context's private stream)

    def close(self):
        if getattr(self, "_h", None):
            self._l.rzb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc:
            raise RzbError(rc, (self._l.rzb_last_error(self._h) or b"").decode())

    def set_stream(self, cuda_stream: Optional[int]):
        """Run on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream, 0 = legacy default stream);
        None = back to the context's private stream."""
        if cuda_stream is None:
            self._check(self._l.rzb_set_stream(self._h, None, 0))
        else:
            self._check(self._l.rzb_set_stream(self._h, cuda_stream if cuda_stream else None, 1))
        self.caller_stream = cuda_stream

    # -- world mirror
    def set_scene(self, scene: Dict[str, np.ndarray]):
        """scene: dict of arrays as produced by rayzath_b200.scenes.FlatScene.arrays() / rzs.read()."""
        s = SceneStruct()
        keep = []

        def arr(name, dtype):
            a = scene.get(name)
            if a is None:
                a = np.zeros(0, dtype=dtype)
            a = np.ascontiguousarray(a)
            if a.dtype != dtype:
                a = a.reshape(-1).view(np.uint8).view(dtype)  # raw bytes / same-layout records
            keep.append(a)
            return a

        def put(field, count_field, name, dtype):
            a = arr(name, dtype).reshape(-1)
            setattr(s, field, a.ctypes.data if a.size else None)
            if count_field:
                setattr(s, count_field, a.shape[0])
            return a

        put("mesh_nodes", "mesh_node_count", "mesh_nodes", node_dtype)
        tris = put("triangles", "triangle_count", "triangles", triangle_dtype)
        thi = scene.get("tri_host_index")
        if thi is not None and len(thi) == tris.shape[0] and tris.shape[0]:
            put("tri_host_index", None, "tri_host_index", np.dtype(u4))
        put("meshes", "mesh_count", "meshes", mesh_dtype)
        put("instance_nodes", "instance_node_count", "instance_nodes", node_dtype)
        put("instances", "instance_count", "instances", instance_dtype)
        put("instance_materials", "instance_material_count", "instance_materials", np.dtype(u4))
        put("materials", "material_count", "materials", material_dtype)
        maps = arr("maps", map_dtype).reshape(-1).copy()
        for i in range(maps.shape[0]):
            px = np.ascontiguousarray(scene["map_pixels_%d" % i])
            keep.append(px)
            maps[i]["pixels"] = px.ctypes.data
        keep.append(maps)
        s.maps = maps.ctypes.data if maps.size else None
        s.map_count = maps.shape[0]
        put("direct_lights", "direct_light_count", "direct_lights", direct_light_dtype)
        put("spot_lights", "spot_light_count", "spot_lights", spot_light_dtype)
        wm = arr("world_material", material_dtype).reshape(-1)
        C.memmove(s.world_material, wm.ctypes.data, 64)
        s.default_material = int(np.asarray(scene["default_material"]).reshape(-1)[0])
        s.flags = int(np.asarray(scene.get("scene_flags", 0)).reshape(-1)[0])
        self._check(self._l.rzb_set_scene(self._h, C.byref(s)))
        self._keep = None

    def update_scene(self, scene: Dict[str, np.ndarray]):
        """Incremental update (RZB_SCENE_KEEP_GEOMETRY): instances, instance tree, materials, maps and lights are replaced,
        the geometry of the last full set_scene stays on the device."""
        s2 = {k: v for k, v in scene.items() if k not in ("mesh_nodes", "triangles", "tri_host_index", "meshes")}
        flags = int(np.asarray(scene.get("scene_flags", 0)).reshape(-1)[0]) | SCENE_KEEP_GEOMETRY
        s2["scene_flags"] = np.array([flags], dtype=np.uint32)
        self.set_scene(s2)

    def set_camera(self, camera: np.ndarray):
        cam = np.ascontiguousarray(camera).view(camera_dtype).reshape(-1)[:1].copy()
        self._check(self._l.rzb_set_camera(self._h, cam.ctypes.data))
        self.width, self.height = int(cam[0]["width"]), int(cam[0]["height"])
        self.camera = cam

    def set_config(self, spot_light_samples=1, direct_light_samples=1, max_depth=16, flags=FLAG_NONE, seed=0):
        cfg = np.zeros(1, dtype=config_dtype)
        cfg[0] = (spot_light_samples, direct_light_samples, max_depth, flags, seed)
        self._check(self._l.rzb_set_config(self._h, cfg.ctypes.data))

    def set_rows(self, row_begin: int, row_end: int):
        """Tile split: render only image rows [row_begin, row_end)."""
        self._check(self._l.rzb_set_rows(self._h, int(row_begin), int(row_end)))

    def set_row_interleave(self, index: int, count: int):
        """Interleaved tile split: render the 16-row chunk rows r with r % count == index."""
        self._check(self._l.rzb_set_row_interleave(self._h, int(index), int(count)))

    # -- frame
    def reset(self):
        self._check(self._l.rzb_reset(self._h))

    def render(self, passes: int):
        self._check(self._l.rzb_render(self._h, int(passes)))

    def synchronize(self):
        self._check(self._l.rzb_synchronize(self._h))

    def resolve(self, rgba8: Optional[np.ndarray] = None, depth: Optional[np.ndarray] = None, want_depth=False):
        """Tone-map + copy to HOST buffers. Returns (rgba8[h,w,4], depth[h,w] or None, ray_count)."""
        n = self.width * self.height
        if rgba8 is None:
            rgba8 = np.empty((self.height, self.width, 4), dtype=np.uint8)
        if depth is None and want_depth:
            depth = np.empty((self.height, self.width), dtype=f4)
        rays = C.c_uint64(0)
        self._check(self._l.rzb_resolve(self._h, rgba8.ctypes.data, _ptr(depth), C.byref(rays)))
        assert rgba8.size == n * 4
        return rgba8, depth, rays.value

    def resolve_async(self, slot: int, rgba8_pinned: Optional[np.ndarray], depth_pinned: Optional[np.ndarray] = None) -> int:
        """rzb_resolve_async: enqueue tone map + copies into PINNED arrays (pinned_array) + the ray-cast pick; returns the
        ray count at once. The arrays are valid after resolve_wait(slot)."""
        rays = C.c_uint64(0)
        self._check(self._l.rzb_resolve_async(self._h, int(slot), _ptr(rgba8_pinned), _ptr(depth_pinned), C.byref(rays)))
        return rays.value

    def resolve_wait(self, slot: int):
        """rzb_resolve_wait: block until the asynchronous resolve of `slot` is done; returns its (instance, material slot) pick."""
        inst, mat = C.c_uint32(0), C.c_uint32(0)
        self._check(self._l.rzb_resolve_wait(self._h, int(slot), C.byref(inst), C.byref(mat)))
        return inst.value, mat.value

    def resolve_peers(self, peers, want_depth=False):
        rgba8 = np.empty((self.height, self.width, 4), dtype=np.uint8)
        depth = np.empty((self.height, self.width), dtype=f4) if want_depth else None
        rays = C.c_uint64(0)
        arr = (_P * max(len(peers), 1))(*[p._h for p in peers])
        self._check(self._l.rzb_resolve_peers(self._h, arr, len(peers), rgba8.ctypes.data, _ptr(depth), C.byref(rays)))
        return rgba8, depth, rays.value

    def read_accum(self) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), dtype=f4)
        self._check(self._l.rzb_read_accum(self._h, out.ctypes.data))
        return out

    def mean_samples(self) -> float:
        """Completed paths per pixel so far (mean of the accumulator's alpha channel over this context's pixels)."""
        m = C.c_double(0.0)
        self._check(self._l.rzb_mean_samples(self._h, C.byref(m)))
        return float(m.value)

    def accum_device_ptr(self):
        p, n = _P(), C.c_size_t(0)
        self._check(self._l.rzb_accum_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def accum_add_device(self, device_ptr: int, pixel_count: int):
        self._check(self._l.rzb_accum_add_device(self._h, device_ptr, pixel_count))

    def accum_ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self._l.rzb_accum_ipc_handle(self._h, buf))
        return buf.raw

    def resolve_ipc(self, handles, want_depth=False):
        """Fused sum-over-NVLink + tone map of this context's and the peers' accumulators (peers = 64-byte IPC handles)."""
        rgba8 = np.empty((self.height, self.width, 4), dtype=np.uint8)
        depth = np.empty((self.height, self.width), dtype=f4) if want_depth else None
        blob = b"".join(handles)
        self._check(self._l.rzb_resolve_ipc(self._h, blob if blob else None, len(handles), rgba8.ctypes.data, _ptr(depth)))
        return rgba8, depth

    def exchange_ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self._l.rzb_exchange_ipc_handle(self._h, buf))
        return buf.raw

    def resolve_sliced(self, rank: int, world: int, accum_handles, exchange_handles, rgba8_pinned=None, depth_pinned=None):
        """rzb_resolve_sliced: the one-kernel-per-rank exchange step (flag barrier in peer memory, slice sum over NVLink,
        tone map into rank 0's staging image). Asynchronous; resolve_sliced_wait() completes it."""
        a, e = b"".join(accum_handles), b"".join(exchange_handles)
        self._check(self._l.rzb_resolve_sliced(self._h, int(rank), int(world), a if a else None, e if e else None,
                                               _ptr(rgba8_pinned), _ptr(depth_pinned)))

    def resolve_sliced_wait(self) -> float:
        ms = C.c_float(0.0)
        self._check(self._l.rzb_resolve_sliced_wait(self._h, C.byref(ms)))
        return float(ms.value)

    def accum_tensor(self):
        """The device accumulator as a torch tensor [h, w, 4] float32 sharing memory (for NCCL collectives)."""
        import torch
        ptr, nbytes = self.accum_device_ptr()

        class _Wrap:
            __cuda_array_interface__ = {"shape": (self.height, self.width, 4), "typestr": "<f4", "data": (ptr, False),
                                        "version": 3, "strides": None}
        return torch.as_tensor(_Wrap(), device="cuda:%d" % self.device)

    def work_counters(self) -> np.ndarray:
        w = np.zeros(1, dtype=work_counters_dtype)
        self._check(self._l.rzb_get_work_counters(self._h, w.ctypes.data))
        return w[0]

    def raycast(self):
        inst, slot = C.c_uint32(0), C.c_uint32(0)
        self._check(self._l.rzb_raycast(self._h, C.byref(inst), C.byref(slot)))
        return inst.value, slot.value

    def render_stats(self) -> np.ndarray:
        st = np.zeros(1, dtype=render_stats_dtype)
        self._check(self._l.rzb_get_render_stats(self._h, st.ctypes.data))
        return st[0]

    def timings(self) -> str:
        buf = C.create_string_buffer(1024)
        self._check(self._l.rzb_timings(self._h, buf, 1024))
        return buf.value.decode()

    # -- ray sets
    def trace_closest(self, origins, directions, near_far, stats=False):
        o, d, nf = _c(origins, f4).reshape(-1, 3), _c(directions, f4).reshape(-1, 3), _c(near_far, f4).reshape(-1, 2)
        n = o.shape[0]
        hits = np.zeros(n, dtype=hit_dtype)
        st = np.zeros(1, dtype=trace_stats_dtype) if stats else None
        self._check(self._l.rzb_trace_closest(self._h, o.ctypes.data, d.ctypes.data, nf.ctypes.data, n,
                                              hits.ctypes.data, _ptr(st)))
        return (hits, st[0]) if stats else hits

    def trace_closest_device(self, o_near_ptr: int, d_far_ptr: int, n: int, hits_ptr: int, timed=False) -> float:
        ms = C.c_float(0.0)
        self._check(self._l.rzb_trace_closest_device(self._h, o_near_ptr, d_far_ptr, n, hits_ptr,
                                                     C.byref(ms) if timed else None))
        return ms.value

    def trace_closest_device_counted(self, o_near_ptr: int, d_far_ptr: int, n: int, hits_ptr: int):
        self._check(self._l.rzb_trace_closest_device_counted(self._h, o_near_ptr, d_far_ptr, n, hits_ptr))

    def trace_any(self, origins, directions, near_far) -> np.ndarray:
        o, d, nf = _c(origins, f4).reshape(-1, 3), _c(directions, f4).reshape(-1, 3), _c(near_far, f4).reshape(-1, 2)
        n = o.shape[0]
        masks = np.zeros((n, 4), dtype=f4)
        self._check(self._l.rzb_trace_any(self._h, o.ctypes.data, d.ctypes.data, nf.ctypes.data, n, masks.ctypes.data))
        return masks

    def generate_camera_rays(self):
        n = self.width * self.height
        o, d, nf = np.zeros((n, 3), f4), np.zeros((n, 3), f4), np.zeros((n, 2), f4)
        self._check(self._l.rzb_generate_camera_rays(self._h, o.ctypes.data, d.ctypes.data, nf.ctypes.data))
        return o, d, nf
