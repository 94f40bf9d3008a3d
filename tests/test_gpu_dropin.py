"""The C++ drop-in: the reference's own headless runner (Application/headless.cpp, compiled in place) with
RayZath::Cuda::Engine implemented by rayzath_b200/host/cuda_engine_b200.cpp on the C ABI. Needs the prebuilt
rayzath_b200/host/_build/rz_b200_headless (built where /root/reference exists; it travels to the GPU box)."""
import json
import os
import re
import subprocess

import pytest

from rayzath_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "rayzath_b200", "host", "_build", "rz_b200_headless")


@pytest.mark.skipif(not os.path.exists(BIN), reason="drop-in binary not built (needs /root/reference at build time)")
@pytest.mark.parametrize("bvh", ["reference", "sah"])
def test_headless_runner_renders_on_the_b200_path(tmp_path, bvh):
    w = scenes.materials_scene(resolution=(320, 180), res=24)
    w.save_reference(str(tmp_path), "scene")
    json.dump({"tasks": [{"scene path": "scene.json", "engine": ["CUDAGPU"], "rpp": 200, "timeout": 30.0}]},
              open(tmp_path / "tasks.json", "w"))
    os.makedirs(tmp_path / "report")
    env = dict(os.environ, RZB200_VERBOSE="1", RZB200_SEED="7", RZB200_BVH=bvh)
    r = subprocess.run([BIN, "--headless", "tasks.json", "report", "-r"], cwd=tmp_path, env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "[rzb200]" in r.stderr and "B200 wavefront engine" in r.stderr, "the CUDA engine did not run: " + r.stderr[-500:]
    reports = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp_path / "report") for f in fs if f == "report.txt"]
    assert len(reports) == 1
    text = open(reports[0]).read()
    assert "CUDAGPU" in text
    m = re.search(r"traced\s+([0-9.]+)([kMGT]?)", text)
    assert m, text
    images = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp_path / "report") for f in fs if f.lower().endswith((".png", ".jpg"))]
    assert images, "no rendered image saved"
    assert os.path.getsize(images[0]) > 1000


SELFTEST = os.path.join(ROOT, "rayzath_b200", "host", "_build", "rz_b200_dropin_selftest")


@pytest.mark.skipif(not os.path.exists(SELFTEST), reason="drop-in self-test not built (needs /root/reference at build time)")
def test_world_edits_between_frames(tmp_path):
    """The reference's host API with edits between renderWorld calls (rz_b200_dropin_selftest.cpp): moving instances and
    recolouring a material restarts accumulation through the incremental upload (RZB_SCENE_KEEP_GEOMETRY); the frames
    equal those of a run that re-uploads everything, and the pipelined sync == false calls keep converging."""
    import numpy as np
    w = scenes.materials_scene(resolution=(320, 180), res=24)
    w.save_reference(str(tmp_path), "scene")
    frames = {}
    for mode in ("incremental", "full"):
        out = tmp_path / (mode + ".raw")
        env = dict(os.environ, RZB200_SEED="11", RZB200_FULL_UPLOAD="1" if mode == "full" else "0")
        r = subprocess.run([SELFTEST, "scene.json", str(out)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        info = json.loads(r.stdout.strip().splitlines()[-1])
        assert (info["width"], info["height"]) == (320, 180)
        frames[mode] = np.fromfile(out, dtype=np.uint8).reshape(3, 180, 320, 4)[..., :3].astype(np.int32)
        assert info["rays"] == 9 * 8 * 320 * 180  # 4 + 5 calls of 8 passes since the restart
    a, b = frames["incremental"], frames["full"]
    assert (np.abs(a - b) > 1).mean() < 1e-3        # same seed, same rays: equal up to atomic-add order in the last bit
    assert np.abs(a[1] - a[0]).mean() > 1.0         # the edit is visible
    assert abs(a[2].mean() - a[1].mean()) / a[1].mean() < 0.05 and np.abs(a[2] - a[1]).mean() < np.abs(a[1] - a[0]).mean()


@pytest.mark.skipif(not os.path.exists(SELFTEST), reason="drop-in self-test not built (needs /root/reference at build time)")
def test_instance_tree_growth_falls_back_to_full_upload(tmp_path):
    """adding instances without touching a mesh keeps the upload incremental (RZB_SCENE_KEEP_GEOMETRY) until
    the instance tree outgrows the node range reserved for it (max(2 x nodes, 1024)); rzb_set_scene then answers
    RZB_ERR_STATE and the engine must re-send the whole scene instead of throwing. 700 extra instances of an existing
    mesh -> an instance tree of > 1024 nodes; the frame renders and the new geometry is visible."""
    import numpy as np
    w = scenes.materials_scene(resolution=(320, 180), res=24)
    w.save_reference(str(tmp_path), "scene")
    out = tmp_path / "grow.raw"
    env = dict(os.environ, RZB200_SEED="11")
    r = subprocess.run([SELFTEST, "scene.json", str(out), "700"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["instances_after_growth"] >= 700
    assert info["rays_after_growth"] == 4 * 8 * 320 * 180  # accumulation restarted with the edit
    f = np.fromfile(out, dtype=np.uint8).reshape(4, 180, 320, 4)[..., :3].astype(np.int32)
    assert np.abs(f[3] - f[2]).mean() > 0.5  # the grown instances are in the picture


@pytest.mark.skipif(not os.path.exists(BIN), reason="drop-in binary not built (needs /root/reference at build time)")
def test_headless_task_keys(tmp_path):
    """the task keys the reference's runner lacks ("max depth", "seed", "devices", "spp",
    "accumulator"; linux_port/patch_headless_cpp.py) through the reference-facing entry: BASELINE config 1 (Cornell box,
    depth 8, 64 spp) reproduced from a task file alone -- the run stops on the spp target, not on the pass budget or the
    timeout, reports depth 8, and leaves the float accumulator whose alpha channel is the sample count."""
    import numpy as np
    w = scenes.cornell(resolution=(256, 256))
    w.save_reference(str(tmp_path), "scene")
    json.dump({"tasks": [{"scene path": "scene.json", "engine": ["CUDAGPU"], "rpp": 1000000, "timeout": 120.0,
                          "max depth": 8, "seed": 5, "devices": "0", "spp": 64, "accumulator": "accum.f32"}]},
              open(tmp_path / "tasks.json", "w"))
    os.makedirs(tmp_path / "report")
    env = dict(os.environ, RZB200_VERBOSE="1")
    env.pop("RZB200_SEED", None)
    r = subprocess.run([BIN, "--headless", "tasks.json", "report", "-r"], cwd=tmp_path, env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "seed 5\n" in r.stderr, r.stderr[-400:]
    reports = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp_path / "report") for f in fs if f == "report.txt"]
    text = open(reports[0]).read()
    assert "max depth: 8" in text, text
    m = re.search(r"duration: ([0-9.]+)s", text)
    assert m and float(m.group(1)) < 60.0, text  # stopped by "spp", long before the timeout
    acc_files = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp_path / "report") for f in fs if f == "accum.f32"]
    assert len(acc_files) == 1
    acc = np.fromfile(acc_files[0], dtype=np.float32).reshape(256, 256, 4)
    spp = float(acc[..., 3].mean())
    assert 64.0 <= spp < 64.0 * 4, spp  # the check runs between renderWorld calls of up to 1024 passes (auto-tuned)
    assert np.isfinite(acc).all() and acc[..., :3].sum() > 0
