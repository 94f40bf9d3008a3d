"""The Python host flattens a world into exactly the arrays the C++ drop-in flattens from the reference's own
World (rayzath_b200/host/world_flatten.hpp through rz_ref_tool dumpscene): same trees from the host BVH builders
(csrc/rzb_bvh_build.cpp vs component_container.hpp / bvh_tree_node.hpp), same triangle order, same axes, boxes,
normals and camera. Checked against committed digests and, when the reference build is present, live."""
import numpy as np
import pytest

import rz_oracle as O
from rayzath_b200 import capi, rzs, scenes
from tests.golden_scenes import GOLDEN_SCENES, array_digest

NAMES = list(GOLDEN_SCENES)


@pytest.mark.parametrize("name", NAMES)
def test_flatten_matches_reference_digests(name, golden, flats, worlds):
    g, flat = golden[name], flats[name]
    checked = 0
    for key, value in flat.items():
        sha = g.get("sha_" + key)
        assert sha is not None, key
        assert bytes(sha.tobytes()) == array_digest(key, value), "%s: %s differs from the reference's World" % (name, key)
        checked += 1
    assert checked >= 13
    assert np.array_equal(worlds[name].camera_struct().view(np.uint8), g["camera"].view(np.uint8))


@pytest.mark.skipif(not O.have_ref_tool(), reason="oracle/_ref/rz_ref_tool not built")
def test_flatten_live_larger_scene(tmp_path):
    """A 20k-triangle mesh through the OBJ path plus maps: trees, triangle order and maps byte-equal."""
    w = scenes.heightfield_scene(resolution=(32, 18), nx=100, nz=100, map_size=64)
    path = w.save_reference(str(tmp_path / "hf"))
    O.ref_tool("dumpscene", path, str(tmp_path / "ref.rzs"))
    ref = rzs.read(str(tmp_path / "ref.rzs"))
    flat = w.flatten()
    for key, value in flat.items():
        assert array_digest(key, value) == array_digest(key, ref[key]), key


def test_bvh_invariants():
    """Structural properties of the reference's tree as rebuilt by rzb_build_mesh_bvh: every triangle in exactly one
    leaf, children adjacent at odd indices, boxes contain their triangles, leaves <= 8 unless unsplittable."""
    v, t, uv, n = scenes.heightfield_mesh(60, 50)
    nodes, order = capi.build_mesh_bvh(v, t)
    assert sorted(order.tolist()) == list(range(t.shape[0]))
    count = nodes["type_count"] & 0x3FFFFFFF
    leaf = count != 0
    assert count[leaf].sum() == t.shape[0]
    inner = np.flatnonzero(~leaf)
    assert (nodes["begin"][inner] % 2 == 1).all() and (nodes["begin"][inner] + 1 < nodes.shape[0]).all()
    # each non-root node is referenced exactly once
    refs = np.concatenate([nodes["begin"][inner], nodes["begin"][inner] + 1])
    assert sorted(refs.tolist()) == list(range(1, nodes.shape[0]))
    tri_v = v[t[order]]
    for i in np.flatnonzero(leaf)[:200]:
        b, c = nodes["begin"][i], count[i]
        pts = tri_v[b:b + c].reshape(-1, 3)
        assert (pts >= nodes["bb_min"][i]).all() and (pts <= nodes["bb_max"][i]).all()
    # parents contain children
    for i in inner[:200]:
        for ch in (nodes["begin"][i], nodes["begin"][i] + 1):
            assert (nodes["bb_min"][ch] >= nodes["bb_min"][i]).all() and (nodes["bb_max"][ch] <= nodes["bb_max"][i]).all()


def test_bvh_edge_cases():
    # empty mesh, single triangle, 32 coplanar triangles (root leaf), 33 coplanar (unsplittable -> one leaf)
    nodes, order = capi.build_mesh_bvh(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    assert nodes.shape[0] == 0
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1]], np.float32)
    nodes, order = capi.build_mesh_bvh(v, np.array([[0, 1, 2]], np.uint32))
    assert nodes.shape[0] == 1 and (nodes["type_count"][0] & 0x3FFFFFFF) == 1
    gv, gt, _, _ = scenes.grid_mesh(4, 4)
    nodes, _ = capi.build_mesh_bvh(gv, gt)
    assert nodes.shape[0] == 1 and (nodes["type_count"][0] & 0x3FFFFFFF) == 32
    gv, gt, _, _ = scenes.grid_mesh(6, 6)
    nodes, _ = capi.build_mesh_bvh(gv, gt)  # flat: no triangle is strictly smaller than the node on y
    assert nodes.shape[0] == 1 and (nodes["type_count"][0] & 0x3FFFFFFF) == 72
    inodes, iorder = capi.build_instance_bvh(np.zeros((0, 6), np.float32))
    assert inodes.shape[0] == 1 and (inodes["type_count"][0] & 0x3FFFFFFF) == 0
