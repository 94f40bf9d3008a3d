#!/usr/bin/env python3
"""Generate tests/golden/*.rzs from the REFERENCE ITSELF (oracle/_ref/rz_ref_tool = the reference's CPU engine
compiled in place from /root/reference by oracle/Makefile). Run in the build container; the vectors travel.

Per scene (small instances of BASELINE.json's configurations, rayzath_b200/scenes.py):
  ray_origins/directions/near_far   the reference's own pixel-centre rays (Kernel::generateSimpleRay)
  hits                              reference closest hit per ray (CPU::Kernel::traverseWorld)
  shadow_* / masks                  shadow rays from the hit points and the reference's any-hit answer
  sha_<array>                       sha256 of every flattened scene array as dumped from the reference's World
  accum_a / accum_b                 two independent reference CPU renders (float accumulators) at equal passes
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rz_oracle as O  # noqa: E402
from rayzath_b200 import rzs  # noqa: E402
from tests.golden_scenes import GOLDEN_SCENES, RENDER_SETTINGS, array_digest, shadow_rays  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="rzb_golden_")
    for name, make in GOLDEN_SCENES.items():
        world = make()
        d = os.path.join(tmp, name)
        path = world.save_reference(d)
        O.ref_tool("dumpscene", path, os.path.join(d, "ref.rzs"))
        ref = rzs.read(os.path.join(d, "ref.rzs"))
        O.ref_tool("trace", path, os.path.join(d, "ref.rzs"), os.path.join(d, "hits.rzs"))
        hits = rzs.read(os.path.join(d, "hits.rzs"))["hits"]
        so, sd, snf = shadow_rays(ref["ray_origins"], ref["ray_directions"], hits)
        rzs.write(os.path.join(d, "shadow.rzs"), {"ray_origins": so, "ray_directions": sd, "ray_near_far": snf})
        O.ref_tool("traceany", path, os.path.join(d, "shadow.rzs"), os.path.join(d, "masks.rzs"))
        masks = rzs.read(os.path.join(d, "masks.rzs"))["masks"]
        out = {"ray_origins": ref["ray_origins"], "ray_directions": ref["ray_directions"],
               "ray_near_far": ref["ray_near_far"], "hits": hits, "masks": masks, "camera": ref["camera"]}
        for k, v in ref.items():
            if k.startswith("ray_") or k == "camera":
                continue
            digest = array_digest(k, v)
            out["sha_" + k] = np.frombuffer(digest, dtype=np.uint8)
        if name in RENDER_SETTINGS:
            passes, depth = RENDER_SETTINGS[name]
            for tag in ("a", "b"):
                O.ref_tool("render", path, passes, os.path.join(d, "render_%s.rzs" % tag), depth, 1, 1, timeout=120.0)
                r = rzs.read(os.path.join(d, "render_%s.rzs" % tag))
                out["accum_" + tag] = r["accum"]
            out["depth"] = r["depth"]
            out["render_settings"] = np.array([passes, depth], dtype=np.uint32)
        rzs.write(os.path.join(OUT, name + ".rzs"), out)
        print(name, {k: v.shape for k, v in out.items() if not k.startswith("sha_")})


if __name__ == "__main__":
    main()
