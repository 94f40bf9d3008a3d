"""The optional SAH triangle-BVH builder (rzb_build_mesh_bvh_sah). CPU part: structure of the
tree and -- through the oracle, which walks any tree in rzb_node format -- the same closest hits as on the reference's
own tree except exact-distance ties. GPU part (-m gpu): the CUDA kernels on the SAH tree (conservative box test,
RZB_SCENE_OWN_TREES) against the oracle on the REFERENCE tree, the shadow masks, and the rendered image."""
import numpy as np
import pytest

import rz_oracle as O
from rayzath_b200 import capi, scenes
from tests.golden_scenes import GOLDEN_SCENES, RENDER_SETTINGS, shadow_rays

NAMES = list(GOLDEN_SCENES)


def _sah_world(name, max_leaf=4, builder="sah"):
    w = GOLDEN_SCENES[name]()
    for m in w.meshes:
        m.bvh_builder = (builder, max_leaf)
    return w


def _depth(nodes):
    depth = np.zeros(nodes.shape[0], np.int32)
    count = nodes["type_count"] & 0x3FFFFFFF
    for i in range(nodes.shape[0]):  # children always follow their parent
        if count[i] == 0:
            depth[nodes["begin"][i]] = depth[nodes["begin"][i] + 1] = depth[i] + 1
    return int(depth.max())


def _incoherent_rays(n=20000, seed=11):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    o[:, 1] = rng.uniform(0.05, 3.0, n).astype(np.float32)
    d = rng.normal(0, 1, (n, 3)).astype(np.float32)
    d = (d / np.sqrt((d * d).sum(1, keepdims=True, dtype=np.float32))).astype(np.float32)
    nf = np.zeros((n, 2), np.float32)
    nf[:, 1] = 3.0e38
    nf[::4, 1] = rng.uniform(0.5, 5.0, n)[::4].astype(np.float32)
    return o, d, nf


def _assert_equal_except_ties(hits, ref, max_tie_fraction=0.001):
    """Whole records equal; where the triangle differs the distance must be exactly the same (a tie)."""
    same = (hits["instance"] == ref["instance"]) & (hits["triangle"] == ref["triangle"])
    ties = np.flatnonzero(~same)
    assert (hits["t"][ties] == ref["t"][ties]).all(), "a different triangle at a different distance"
    assert ties.size <= max_tie_fraction * hits.shape[0] + 1, "listed ties: %s" % ties[:20].tolist()
    assert np.array_equal(hits[same].view(np.uint8).reshape(-1, 24), ref[same].view(np.uint8).reshape(-1, 24))


def test_sah_tree_structure():
    v, t, uv, n = scenes.heightfield_mesh(60, 50)
    for max_leaf in (1, 4, 8):
        nodes, order = capi.build_mesh_bvh_sah(v, t, max_leaf)
        assert sorted(order.tolist()) == list(range(t.shape[0]))
        count = nodes["type_count"] & 0x3FFFFFFF
        leaf = count != 0
        assert count[leaf].sum() == t.shape[0] and count.max() <= max_leaf
        inner = np.flatnonzero(~leaf)
        assert (nodes["begin"][inner] % 2 == 1).all()
        refs = np.concatenate([nodes["begin"][inner], nodes["begin"][inner] + 1])
        assert sorted(refs.tolist()) == list(range(1, nodes.shape[0]))
        assert ((nodes["type_count"][inner] >> 30) <= 2).all()  # split axis Z/Y/X, never the reference's Size type
        # leaves tile the triangle order without gaps
        lb = np.sort(nodes["begin"][leaf])
        assert lb[0] == 0 and (np.diff(lb) > 0).all()
        tri_v = v[t[order]]
        for i in np.flatnonzero(leaf)[:300]:
            pts = tri_v[nodes["begin"][i]:nodes["begin"][i] + count[i]].reshape(-1, 3)
            assert (pts >= nodes["bb_min"][i]).all() and (pts <= nodes["bb_max"][i]).all()
        for i in inner[:300]:
            for ch in (nodes["begin"][i], nodes["begin"][i] + 1):
                assert (nodes["bb_min"][ch] >= nodes["bb_min"][i]).all() and (nodes["bb_max"][ch] <= nodes["bb_max"][i]).all()
        assert _depth(nodes) <= 31


def test_sah_tree_edge_cases():
    nodes, order = capi.build_mesh_bvh_sah(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    assert nodes.shape[0] == 0
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1]], np.float32)
    nodes, order = capi.build_mesh_bvh_sah(v, np.array([[0, 1, 2]], np.uint32))
    assert nodes.shape[0] == 1 and (nodes["type_count"][0] & 0x3FFFFFFF) == 1
    # 4096 coincident triangles: no split plane exists -> halved by index down to the depth limit, depth stays <= 31
    t = np.tile(np.array([[0, 1, 2]], np.uint32), (4096, 1))
    nodes, order = capi.build_mesh_bvh_sah(v, t, 4)
    assert sorted(order.tolist()) == list(range(4096)) and _depth(nodes) <= 31
    with pytest.raises(capi.RzbError):
        capi.build_mesh_bvh_sah(v, np.array([[0, 1, 7]], np.uint32))


@pytest.mark.parametrize("name", NAMES)
def test_sah_tree_same_hits_as_reference_tree_oracle(name, golden, flats):
    """CPU only: the oracle walks the SAH tree and the reference's tree; equal records except exact ties."""
    g = golden[name]
    flat = _sah_world(name).flatten()
    assert flat["mesh_nodes"].shape[0] != flats[name]["mesh_nodes"].shape[0] or name == "cornell"
    for rays in ((g["ray_origins"], g["ray_directions"], g["ray_near_far"]), _incoherent_rays(5000)):
        ref = O.trace_closest(O.Scene(flats[name]), *rays)
        own = O.trace_closest(O.Scene(flat), *rays)
        _assert_equal_except_ties(own, ref)


def test_refit_keeps_topology_and_bounds_the_deformed_mesh():
    """rzb_refit_mesh_bvh: after the vertices move, the refitted tree keeps children / ranges / order,
    every leaf box is the exact min / max of its triangles, every inner box the union of its children -- and the oracle
    finds the same closest hits on it as on a tree built from scratch for the deformed mesh (exact ties excepted)."""
    v, t, uv, n = scenes.heightfield_mesh(60, 50)
    rng = np.random.default_rng(8)
    v2 = v.copy()
    v2[:, 1] += (0.4 * np.sin(v[:, 0] * 1.7) * np.cos(v[:, 2] * 1.3) + 0.05 * rng.standard_normal(v.shape[0])).astype(np.float32)
    v2[:, 0] += (0.03 * rng.standard_normal(v.shape[0])).astype(np.float32)
    for build in (capi.build_mesh_bvh, lambda a, b: capi.build_mesh_bvh_sah(a, b, 4)):
        nodes, order = build(v, t)
        refit = capi.refit_mesh_bvh(v2, t, nodes, order)
        assert np.array_equal(refit["begin"], nodes["begin"]) and np.array_equal(refit["type_count"], nodes["type_count"])
        count = refit["type_count"] & 0x3FFFFFFF
        tri_v = v2[t[order]]
        for i in range(refit.shape[0]):
            if count[i]:
                pts = tri_v[refit["begin"][i]:refit["begin"][i] + count[i]].reshape(-1, 3)
                assert np.array_equal(refit["bb_min"][i], pts.min(axis=0)) and np.array_equal(refit["bb_max"][i], pts.max(axis=0))
            else:
                a, b = refit[refit["begin"][i]], refit[refit["begin"][i] + 1]
                assert np.array_equal(refit["bb_min"][i], np.minimum(a["bb_min"], b["bb_min"]))
                assert np.array_equal(refit["bb_max"][i], np.maximum(a["bb_max"], b["bb_max"]))
        # the same hits as a fresh build of the deformed mesh
        from rayzath_b200.world import World
        worlds = []
        for custom in (None, (refit, order)):
            w = World()
            mat = w.create_material("m", color=(200, 200, 200, 255))
            m = w.create_mesh("terrain", v2, t)
            if custom is not None:
                m.bvh_builder = "refit"
                m._bvh = ("refit", custom)
            w.create_instance("terrain", m, [mat])
            w.create_camera(name="cam", position=(0.0, 4.5, -11.0), rotation=(-0.3, 0.0, 0.0), resolution=(48, 27))
            worlds.append(w.flatten())
        rays = _incoherent_rays(4000)
        _assert_equal_except_ties(O.trace_closest(O.Scene(worlds[1]), *rays), O.trace_closest(O.Scene(worlds[0]), *rays))
    with pytest.raises(capi.RzbError):
        capi.refit_mesh_bvh(v2[:10], t, nodes, order)  # vertex ids out of range


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("builder", ["sah", "sah4"])
@pytest.mark.parametrize("name", NAMES)
def test_gpu_sah_tree_hits_vs_oracle_on_reference_tree(name, builder, golden, flats):
    """`sah4` = the same SAH tree collapsed into a 4-ary one by rzb_set_scene (RZB_SCENE_WIDE_TREES): one fetch and four
    conservative box tests per step."""
    g = golden[name]
    w = _sah_world(name, builder=builder)
    flat = w.flatten()
    assert int(flat["scene_flags"][0]) == capi.SCENE_OWN_TREES | (capi.SCENE_WIDE_TREES if builder == "sah4" else 0)
    with capi.Context(0) as c:
        c.set_scene(flat)
        c.set_camera(w.camera_struct())
        for rays in ((g["ray_origins"], g["ray_directions"], g["ray_near_far"]), _incoherent_rays()):
            hits = c.trace_closest(*rays)
            ref = O.trace_closest(O.Scene(flats[name]), *rays, order=O.ORDER_CUDA, minmax=O.MINMAX_FMINF)
            _assert_equal_except_ties(hits, ref)
        # shadow queries, CPU semantics: 0 / 1 masks are order-independent
        c.set_config(flags=capi.FLAG_CPU_SEMANTICS)
        so, sd, snf = shadow_rays(g["ray_origins"], g["ray_directions"], g["hits"])
        masks = c.trace_any(so, sd, snf)
        refm = O.trace_any(O.Scene(flats[name]), so, sd, snf)
        assert np.array_equal(masks, refm)


@pytest.mark.gpu
def test_gpu_sah_tree_full_size_primary_rays():
    """1,001,112 triangles x 2,073,600 pixel-centre rays: SAH tree + conservative boxes against the reference tree +
    exact boxes (itself byte-equal to the oracle, test_gpu_parity.py), both on the GPU."""
    w = scenes.heightfield_scene(resolution=(1920, 1080))
    cam = w.camera_struct()
    res = {}
    for builder in ("reference", ("sah", 4), ("sah4", 4)):
        for m in w.meshes:
            m.bvh_builder = builder
        with capi.Context(0) as c:
            c.set_scene(w.flatten())
            c.set_camera(cam)
            o, d, nf = c.generate_camera_rays()
            res[builder if builder == "reference" else builder[0]] = c.trace_closest(o, d, nf)
    assert (res["reference"]["instance"] != capi.NO_INDEX).mean() > 0.5
    _assert_equal_except_ties(res["sah"], res["reference"], max_tie_fraction=1e-4)
    _assert_equal_except_ties(res["sah4"], res["reference"], max_tie_fraction=1e-4)


@pytest.mark.gpu
def test_gpu_sah_tree_same_image():
    """Same seed, same passes: the paths are the same rays, so the accumulators agree up to the order of the shadow
    kernel's atomic adds and exact-tie winners."""
    name = "materials"
    passes, depth = RENDER_SETTINGS[name]
    acc = {}
    for label, w in (("reference", GOLDEN_SCENES[name]()), ("sah", _sah_world(name)), ("sah4", _sah_world(name, builder="sah4"))):
        with capi.Context(0) as c:
            c.set_scene(w.flatten())
            c.set_camera(w.camera_struct())
            c.set_config(max_depth=depth, seed=77)
            c.reset()
            c.render(passes)
            acc[label] = c.read_accum()
    for other in ("sah", "sah4"):
        a, b = acc["reference"], acc[other]
        assert np.array_equal(a[..., 3], b[..., 3])  # the same paths ended in the same passes
        close = np.isclose(a[..., :3], b[..., :3], rtol=1e-4, atol=1e-6).all(axis=2)
        assert close.mean() > 0.999, (other, close.mean())
        assert abs(a[..., :3].mean() - b[..., :3].mean()) / a[..., :3].mean() < 1e-3


# ---------------------------------------------------------------------------------------------- GPU builder (linear BVH)
def _check_tree(nodes, order, v, t, max_leaf):
    assert sorted(order.tolist()) == list(range(t.shape[0]))
    count = nodes["type_count"] & 0x3FFFFFFF
    leaf = count != 0
    assert count[leaf].sum() == t.shape[0] and count.max() <= max_leaf
    inner = np.flatnonzero(~leaf)
    assert (nodes["begin"][inner] % 2 == 1).all()
    refs = np.concatenate([nodes["begin"][inner], nodes["begin"][inner] + 1])
    assert sorted(refs.tolist()) == list(range(1, nodes.shape[0]))
    lb = np.sort(nodes["begin"][leaf])
    assert lb[0] == 0 and (np.diff(lb) > 0).all()
    tri_v = v[t[order]]
    for i in np.flatnonzero(leaf)[::max(1, int(leaf.sum()) // 300)]:
        pts = tri_v[nodes["begin"][i]:nodes["begin"][i] + count[i]].reshape(-1, 3)
        assert (pts >= nodes["bb_min"][i]).all() and (pts <= nodes["bb_max"][i]).all()
    for i in inner[::max(1, inner.size // 300)]:
        for ch in (nodes["begin"][i], nodes["begin"][i] + 1):
            assert (nodes["bb_min"][ch] >= nodes["bb_min"][i]).all() and (nodes["bb_max"][ch] <= nodes["bb_max"][i]).all()


@pytest.mark.gpu
def test_gpu_lbvh_builder_structure_and_speed():
    v, t, uv, n = scenes.heightfield_mesh(300, 300)
    for max_leaf in (1, 4, 8):
        nodes, order = capi.build_mesh_bvh_lbvh(v, t, max_leaf)
        _check_tree(nodes, order, v, t, max_leaf)
    # degenerate inputs: a single triangle, fewer triangles than a leaf holds, coincident triangles (-> SAH fallback)
    v1 = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1]], np.float32)
    nodes, order = capi.build_mesh_bvh_lbvh(v1, np.array([[0, 1, 2]], np.uint32), 4)
    assert nodes.shape[0] == 1 and (nodes["type_count"][0] & 0x3FFFFFFF) == 1
    same = np.tile(np.array([[0, 1, 2]], np.uint32), (4096, 1))
    nodes, order = capi.build_mesh_bvh_lbvh(v1, same, 4)
    _check_tree(nodes, order, v1, same, 4)
    with pytest.raises(capi.RzbError):
        capi.build_mesh_bvh_lbvh(v1, np.array([[0, 1, 9]], np.uint32), 4)
    # 1M triangles: device time of the build (keys, sort, radix tree, boxes, collapse, emission)
    v, t, uv, n = scenes.heightfield_mesh(708, 707)
    capi.build_mesh_bvh_lbvh(v, t, 4)
    nodes, order, ms = capi.build_mesh_bvh_lbvh(v, t, 4, timing=True)
    print("LBVH 1,001,112 triangles: %d nodes, %.2f ms on the device" % (nodes.shape[0], ms))
    assert 0.0 < ms < 100.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_lbvh_tree_hits_vs_oracle_on_reference_tree(name, golden, flats):
    g = golden[name]
    w = GOLDEN_SCENES[name]()
    for m in w.meshes:
        m.bvh_builder = ("lbvh", 4)
    flat = w.flatten()
    assert int(flat["scene_flags"][0]) == capi.SCENE_OWN_TREES
    with capi.Context(0) as c:
        c.set_scene(flat)
        c.set_camera(w.camera_struct())
        for rays in ((g["ray_origins"], g["ray_directions"], g["ray_near_far"]), _incoherent_rays()):
            hits = c.trace_closest(*rays)
            ref = O.trace_closest(O.Scene(flats[name]), *rays, order=O.ORDER_CUDA, minmax=O.MINMAX_FMINF)
            _assert_equal_except_ties(hits, ref)
