import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import rz_oracle as O
from rayzath_b200 import capi, rzs
from tests.golden_scenes import GOLDEN_SCENES
for name, make in GOLDEN_SCENES.items():
    w = make(); flat = w.flatten()
    g = rzs.read(os.path.join(ROOT, "tests", "golden", name + ".rzs"))
    with capi.Context(0) as c:
        c.set_scene(flat); c.set_camera(w.camera_struct())
        hits, st = c.trace_closest(g["ray_origins"], g["ray_directions"], g["ray_near_far"], stats=True)
    ref, rst = O.trace_closest(O.Scene(flat), g["ray_origins"], g["ray_directions"], g["ray_near_far"], order=1, minmax=1, stats=True)
    print(name, "equal", np.array_equal(hits.view(np.uint8), ref.view(np.uint8)), "gpu", st, "oracle", rst, "inst nodes", flat["instance_nodes"].shape)
