import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayzath_b200 import capi, scenes
import numpy as np
for name, w in (("materials_small", scenes.materials_scene(resolution=(320, 180), res=24)), ("cornell", scenes.cornell(resolution=(256, 256))),
                ("heightfield_small", scenes.heightfield_scene(resolution=(320,180), nx=100, nz=100, map_size=64)), ("materials", scenes.materials_scene())):
    try:
        with capi.Context(0) as ctx:
            ctx.set_scene(w.flatten()); ctx.set_camera(w.camera_struct()); ctx.set_config(1, 1, 16, 0, 5); ctx.reset()
            for i in range(40):
                ctx.render(1)
            ctx.synchronize()
            acc = ctx.read_accum()
            print(name, "ok", float(acc[..., 3].mean()), np.isfinite(acc).all(), acc[..., :3].mean())
    except Exception as e:
        print(name, "FAILED", e)
        break
