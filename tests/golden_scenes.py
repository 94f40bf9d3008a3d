"""Small instances of BASELINE.json's configurations shared by tests/tools/make_golden.py and the tests.
Changing anything here invalidates tests/golden/*.rzs (regenerate with tests/tools/make_golden.py)."""
import hashlib

import numpy as np

from rayzath_b200 import scenes

GOLDEN_SCENES = {
    "cornell": lambda: scenes.cornell(resolution=(40, 40)),
    "materials": lambda: scenes.materials_scene(resolution=(64, 36), res=16, cpu_comparable=True),
    "heightfield": lambda: scenes.heightfield_scene(resolution=(64, 36), nx=48, nz=40, map_size=32),
    "instancing": lambda: scenes.instancing_scene(resolution=(64, 36), n_instances=12, nx=10, nz=10),
}
# reference CPU renders committed per scene: (passes, max_depth)
RENDER_SETTINGS = {"cornell": (768, 8), "materials": (384, 8), "heightfield": (256, 8)}

MAP_FIELDS = ("format", "width", "height", "filter", "address", "scale", "rotation", "translation")


def shadow_rays(origins, directions, hits):
    """Shadow rays from every hit point towards a fixed direction (offset like the integrator's 1e-4*t step)."""
    hit = hits["instance"] != 0xFFFFFFFF
    ldir = np.array([0.4, 1.0, -0.5], dtype=np.float32)
    ldir = (ldir / np.float32(np.sqrt((ldir * ldir).sum(dtype=np.float32)))).astype(np.float32)
    p = origins[hit] + directions[hit] * hits["t"][hit][:, None]
    p = (p + ldir * np.float32(1e-3)).astype(np.float32)
    d = np.tile(ldir, (p.shape[0], 1)).astype(np.float32)
    nf = np.tile(np.array([0.0, 3.0e38], dtype=np.float32), (p.shape[0], 1))
    return p, d, nf


def array_digest(name, a):
    a = np.ascontiguousarray(a)
    if name == "maps":  # pointers and padding are not content
        if a.dtype.names is None:
            return hashlib.sha256(b"").digest() if a.size == 0 else hashlib.sha256(a.tobytes()).digest()
        return hashlib.sha256(b"".join(np.ascontiguousarray(a[f]).tobytes() for f in MAP_FIELDS)).digest()
    return hashlib.sha256(a.tobytes()).digest()
