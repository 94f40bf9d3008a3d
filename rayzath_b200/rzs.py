"""Reader/writer of the "RZSC" container (oracle/rzs_io.hpp): a flat list of named raw arrays used to move
flattened scenes, ray sets, hit dumps and accumulators between the C++ tools and Python.

layout: magic "RZSC" | u32 version=1 | u32 n | n x { char name[32] | u32 elem_size | u32 pad | u64 count | bytes, padded to 8 }
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

from . import capi

# arrays whose element type is known by name; everything else comes back as raw bytes [count, elem_size]
KNOWN = {
    "mesh_nodes": capi.node_dtype, "triangles": capi.triangle_dtype, "tri_host_index": np.dtype(np.uint32),
    "meshes": capi.mesh_dtype, "instance_nodes": capi.node_dtype, "instances": capi.instance_dtype,
    "instance_materials": np.dtype(np.uint32), "materials": capi.material_dtype, "maps": capi.map_dtype,
    "direct_lights": capi.direct_light_dtype, "spot_lights": capi.spot_light_dtype,
    "world_material": capi.material_dtype, "default_material": np.dtype(np.uint32), "camera": capi.camera_dtype,
    "ray_origins": np.dtype((np.float32, 3)), "ray_directions": np.dtype((np.float32, 3)),
    "ray_near_far": np.dtype((np.float32, 2)), "hits": capi.hit_dtype, "masks": np.dtype((np.float32, 4)),
    "accum": np.dtype((np.float32, 4)), "rgba8": np.dtype((np.uint8, 4)), "depth": np.dtype(np.float32),
    "resolution": np.dtype(np.uint32),
}


def read(path: str) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"RZSC":
        raise ValueError("bad magic in " + path)
    version, n = struct.unpack_from("<II", data, 4)
    if version != 1:
        raise ValueError("unsupported RZSC version %d" % version)
    off = 12
    for _ in range(n):
        name = data[off:off + 32].split(b"\0", 1)[0].decode()
        elem_size, _pad, count = struct.unpack_from("<IIQ", data, off + 32)
        off += 48
        nbytes = elem_size * count
        raw = np.frombuffer(data, dtype=np.uint8, count=nbytes, offset=off).copy()
        off += nbytes + (8 - nbytes % 8) % 8
        dt = KNOWN.get(name)
        if name.startswith("map_pixels_"):
            out[name] = raw
        elif dt is not None and dt.itemsize == elem_size:
            out[name] = raw.view(dt.base if dt.subdtype else dt).reshape((count,) + (dt.shape if dt.subdtype else ()))
        else:
            out[name] = raw.reshape(count, elem_size) if elem_size else raw
    return out


def write(path: str, arrays: Dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(b"RZSC")
        f.write(struct.pack("<II", 1, len(arrays)))
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            dt = KNOWN.get(name)
            if dt is not None:
                elem_size = dt.itemsize
            elif a.ndim >= 2:
                elem_size = a.dtype.itemsize * int(np.prod(a.shape[1:]))
            else:
                elem_size = a.dtype.itemsize
            raw = a.reshape(-1).view(np.uint8)
            count = raw.size // elem_size if elem_size else 0
            f.write(name.encode()[:31].ljust(32, b"\0"))
            f.write(struct.pack("<IIQ", elem_size, 0, count))
            f.write(raw.tobytes())
            f.write(b"\0" * ((8 - raw.size % 8) % 8))
