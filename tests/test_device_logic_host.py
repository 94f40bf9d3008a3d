"""CPU check of the DEVICE traversal's control flow: rayzath_b200/csrc/rzb_traverse.cuh is compiled for the host
(tests/host_sim/traverse_host.cpp maps the CUDA intrinsics to plain fp32 operations, -ffp-contract=off) and run one
lane at a time. Phases, deferred-node stack, reciprocal slab test with exact fallback, instance transitions and the
work counters must reproduce the oracle bit for bit. (No GPU involved; the kernels themselves are tested with -m gpu.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import rz_oracle as O
from rayzath_b200 import capi
from tests.golden_scenes import GOLDEN_SCENES, shadow_rays

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_sim", "traverse_host.cpp")
OUT = os.path.join(ROOT, "tests", "host_sim", "_build", "libtrav_host.so")
NAMES = list(GOLDEN_SCENES)


@pytest.fixture(scope="module")
def sim():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, "rayzath_b200", "csrc", f)
                    for f in ("rzb_traverse.cuh", "rzb_traverse_mr.cuh", "rzb_device.cuh", "rzb_wide.hpp")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-I/usr/local/cuda/include",
                        "-o", OUT, SRC], check=True)
    lib = C.CDLL(OUT)
    P = C.c_void_p
    lib.trav_host_run.argtypes = [P, P, P, P, C.c_uint32, C.c_int, P, P, P]
    lib.trav_host_run_mr.argtypes = [P, P, P, P, C.c_uint32, P, P]
    lib.trav_host_run_wide.argtypes = [P, P, P, P, C.c_uint32, P, P]
    return lib


def _closest(lib, scene, o, d, nf):
    o, d, nf = (np.ascontiguousarray(x, dtype=np.float32) for x in (o, d, nf))
    hits = np.zeros(len(o), dtype=capi.hit_dtype)
    cnt = np.zeros(4, np.uint64)
    lib.trav_host_run(C.addressof(scene.struct), o.ctypes.data, d.ctypes.data, nf.ctypes.data, len(o), 0,
                      hits.ctypes.data, None, cnt.ctypes.data)
    return hits, cnt


@pytest.mark.parametrize("name", NAMES)
def test_device_traversal_logic_matches_oracle(name, sim, golden, flats):
    g = golden[name]
    scene = O.Scene(flats[name])
    hits, cnt = _closest(sim, scene, g["ray_origins"], g["ray_directions"], g["ray_near_far"])
    ref, rst = O.trace_closest(scene, g["ray_origins"], g["ray_directions"], g["ray_near_far"], order=O.ORDER_CUDA,
                               minmax=O.MINMAX_FMINF, stats=True)
    assert np.array_equal(hits.view(np.uint8), ref.view(np.uint8))
    assert cnt.tolist() == [int(rst[k]) for k in ("top_nodes", "instances_entered", "mesh_nodes", "triangles")]


@pytest.mark.parametrize("name", NAMES)
def test_device_traversal_logic_incoherent_rays(name, sim, flats):
    rng = np.random.default_rng(3)
    n = 20000
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    o[:, 1] = rng.uniform(0.05, 3, n)
    d = rng.normal(0, 1, (n, 3)).astype(np.float32)
    d = (d / np.sqrt((d * d).sum(1, keepdims=True, dtype=np.float32))).astype(np.float32)
    nf = np.zeros((n, 2), np.float32)
    nf[:, 1] = 3e38
    nf[::4, 1] = rng.uniform(0.5, 5, n)[::4]
    nf[1::7, 0] = 0.7
    # axis-aligned directions: zero components -> infinities and NaNs in the slab arithmetic
    d[::50] = np.array([0, -1, 0], np.float32)
    d[1::50] = np.array([1, 0, 0], np.float32)
    scene = O.Scene(flats[name])
    hits, cnt = _closest(sim, scene, o, d, nf)
    ref, rst = O.trace_closest(scene, o, d, nf, order=O.ORDER_CUDA, minmax=O.MINMAX_FMINF, stats=True)
    assert np.array_equal(hits.view(np.uint8), ref.view(np.uint8))
    assert cnt.tolist() == [int(rst[k]) for k in ("top_nodes", "instances_entered", "mesh_nodes", "triangles")]


@pytest.mark.parametrize("name", NAMES)
def test_device_any_hit_logic_matches_reference(name, sim, golden, flats):
    g = golden[name]
    so, sd, snf = (np.ascontiguousarray(x) for x in shadow_rays(g["ray_origins"], g["ray_directions"], g["hits"]))
    masks = np.zeros((len(so), 4), np.float32)
    scene = O.Scene(flats[name])
    sim.trav_host_run(C.addressof(scene.struct), so.ctypes.data, sd.ctypes.data, snf.ctypes.data, len(so), 1, None,
                      masks.ctypes.data, None)
    assert np.array_equal(masks, g["masks"])


def _closest_mr(lib, scene, o, d, nf):
    o, d, nf = (np.ascontiguousarray(x, dtype=np.float32) for x in (o, d, nf))
    hits = np.zeros(len(o), dtype=capi.hit_dtype)
    cnt = np.zeros(4, np.uint64)
    lib.trav_host_run_mr(C.addressof(scene.struct), o.ctypes.data, d.ctypes.data, nf.ctypes.data, len(o), hits.ctypes.data,
                         cnt.ctypes.data)
    return hits, cnt


@pytest.mark.parametrize("name", NAMES)
def test_multi_ray_lane_walk_matches_oracle(name, sim, golden, flats):
    """rzb_traverse_mr.cuh (several rays per lane, phase voting) on the host: one lane with kMrRays rays in flight. Every
    ray must perform exactly the oracle's operation sequence -- records bit-equal, box / instance / triangle counts equal --
    on the golden primary rays and on 20,000 incoherent rays."""
    g = golden[name]
    scene = O.Scene(flats[name])
    rng = np.random.default_rng(5)
    n = 20000
    io = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    io[:, 1] = rng.uniform(0.05, 3, n)
    idd = rng.normal(size=(n, 3)).astype(np.float32)
    idd = (idd / np.sqrt((idd * idd).sum(axis=1, dtype=np.float32))[:, None]).astype(np.float32)
    inf = np.tile(np.array([0.0, 3.0e38], dtype=np.float32), (n, 1))
    o = np.concatenate([g["ray_origins"], io])
    d = np.concatenate([g["ray_directions"], idd])
    nf = np.concatenate([g["ray_near_far"], inf])
    hits, cnt = _closest_mr(sim, scene, o, d, nf)
    ref, rst = O.trace_closest(scene, o, d, nf, order=O.ORDER_CUDA, minmax=O.MINMAX_FMINF, stats=True)
    assert np.array_equal(hits.view(np.uint8), ref.view(np.uint8))
    assert cnt.tolist() == [int(rst[k]) for k in ("top_nodes", "instances_entered", "mesh_nodes", "triangles")]


@pytest.mark.parametrize("name", NAMES)
def test_wide_tree_walk_matches_oracle_except_ties(name, sim, golden, flats):
    """RZB_SCENE_WIDE_TREES on the host: the SAH tree of every mesh collapsed to a 4-ary tree (rzb_wide.hpp, as
    rzb_set_scene does) and walked by the own-tree traversal (conservative boxes, nearest of four first). Same closest
    hits as the oracle on the REFERENCE tree except exact-distance ties."""
    from tests.golden_scenes import GOLDEN_SCENES
    w = GOLDEN_SCENES[name]()
    for m in w.meshes:
        m.bvh_builder = ("sah4", 4)
    scene = O.Scene(w.flatten())
    g = golden[name]
    rng = np.random.default_rng(9)
    n = 20000
    io = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    io[:, 1] = rng.uniform(0.05, 3, n)
    idd = rng.normal(size=(n, 3)).astype(np.float32)
    idd = (idd / np.sqrt((idd * idd).sum(axis=1, dtype=np.float32))[:, None]).astype(np.float32)
    inf = np.tile(np.array([0.0, 3.0e38], dtype=np.float32), (n, 1))
    o = np.ascontiguousarray(np.concatenate([g["ray_origins"], io]), dtype=np.float32)
    d = np.ascontiguousarray(np.concatenate([g["ray_directions"], idd]), dtype=np.float32)
    nf = np.ascontiguousarray(np.concatenate([g["ray_near_far"], inf]), dtype=np.float32)
    hits = np.zeros(len(o), dtype=capi.hit_dtype)
    cnt = np.zeros(4, np.uint64)
    rc = sim.trav_host_run_wide(C.addressof(scene.struct), o.ctypes.data, d.ctypes.data, nf.ctypes.data, len(o), hits.ctypes.data,
                                cnt.ctypes.data)
    assert rc == 0
    ref = O.trace_closest(O.Scene(flats[name]), o, d, nf, order=O.ORDER_CUDA, minmax=O.MINMAX_FMINF)
    same = (hits["instance"] == ref["instance"]) & (hits["triangle"] == ref["triangle"])
    ties = np.flatnonzero(~same)
    assert (hits["t"][ties] == ref["t"][ties]).all(), "a different triangle at a different distance: %s" % ties[:10].tolist()
    assert ties.size <= 0.001 * len(o) + 1
    assert np.array_equal(hits[same].view(np.uint8).reshape(-1, 24), ref[same].view(np.uint8).reshape(-1, 24))
