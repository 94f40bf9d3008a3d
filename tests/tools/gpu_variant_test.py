import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
code = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "oracle"))
import numpy as np, torch
import rz_oracle as O
from rayzath_b200 import capi, rzs, scenes
from tests.golden_scenes import GOLDEN_SCENES
bad = 0
for rep in range(3):
  for name, make in GOLDEN_SCENES.items():
    w = make(); flat = w.flatten()
    g = rzs.read(os.path.join(%r, "tests", "golden", name + ".rzs"))
    with capi.Context(0) as c:
        c.set_scene(flat); c.set_camera(w.camera_struct())
        hits, st = c.trace_closest(g["ray_origins"], g["ray_directions"], g["ray_near_far"], stats=True)
    ref, rst = O.trace_closest(O.Scene(flat), g["ray_origins"], g["ray_directions"], g["ray_near_far"], order=1, minmax=1, stats=True)
    ok = np.array_equal(hits.view(np.uint8), ref.view(np.uint8)) and all(int(st[k]) == int(rst[k]) for k in ("top_nodes", "instances_entered", "mesh_nodes", "triangles"))
    bad += (not ok)
    if not ok: print("  MISMATCH", name, st, rst)
print("counter mismatches:", bad)
w = scenes.materials_scene()
ctx = capi.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_scene(w.flatten()); ctx.set_camera(w.camera_struct()); ctx.set_config(1, 1, 16, 0, 5); ctx.reset()
ctx.render(64); ctx.synchronize()
acc = ctx.read_accum()
print("render ok", float(acc[..., 3].mean()), float(acc[..., :3].mean()))
''' % (ROOT, ROOT, ROOT)
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["RZB200_LIB"] = os.path.join(ROOT, lib)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("=== variant", lib, "rc", r.returncode)
    print(r.stdout[-1500:])
    print(r.stderr[-600:])
