"""Procedural scenes of BASELINE.json's five configurations, built on rayzath_b200.world.

Every scene is deterministic (fixed seeds) and can be (a) flattened for the C ABI and (b) saved in the
reference's own scene format so that the reference CPU engine renders the identical world.
Sizes are parameters: tests use small instances of the same generators, bench.py the full ones.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .world import World, Mesh

f4 = np.float32


# ---------------------------------------------------------------------- mesh generators (all fp32)
def _orient_outward(vertices, tris, center=None):
    """Flip triangles whose front face ((v2-v1)x(v3-v1), see mesh_component.cpp:19-26) looks at `center`."""
    v = vertices.astype(np.float64)
    c = v.mean(axis=0) if center is None else np.asarray(center, dtype=np.float64)
    a, b, d = v[tris[:, 0]], v[tris[:, 1]], v[tris[:, 2]]
    n = np.cross(b - a, d - a)
    out = (a + b + d) / 3.0 - c
    flip = (n * out).sum(axis=1) < 0
    t = tris.copy()
    t[flip] = t[flip][:, [0, 2, 1]]
    return t, flip


def quad_mesh(p0, eu, ev):
    """Two triangles spanning p0 + s*eu + t*ev (callers orient the winding)."""
    p0, eu, ev = (np.asarray(x, dtype=f4) for x in (p0, eu, ev))
    v = np.stack([p0, p0 + eu, p0 + ev, p0 + eu + ev]).astype(f4)
    uv = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=f4)
    tris = np.array([[0, 1, 2], [1, 3, 2]], dtype=np.uint32)
    return v, tris, uv


def box_mesh():
    """Unit cube centred at the origin, 12 triangles, outward faces, per-face uv."""
    v = np.array([[x, y, z] for x in (-0.5, 0.5) for y in (-0.5, 0.5) for z in (-0.5, 0.5)], dtype=f4)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris, tuv = [], []
    for a, b, c, d in quads:
        tris += [[a, b, c], [a, c, d]]
        tuv += [[0, 1, 2], [0, 2, 3]]
    tris = np.array(tris, dtype=np.uint32)
    tuv = np.array(tuv, dtype=np.uint32)
    tris2, flip = _orient_outward(v, tris, center=(0, 0, 0))
    tuv[flip] = tuv[flip][:, [0, 2, 1]]
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=f4)
    return v, tris2, uv, tuv


def grid_mesh(nx, nz, size_x=1.0, size_z=1.0, height=None, uv_scale=1.0):
    """(nx x nz cells) grid in the XZ plane facing +Y, optional height function h(x, z) -> y; smooth normals."""
    xs = np.linspace(-0.5 * size_x, 0.5 * size_x, nx + 1, dtype=np.float64)
    zs = np.linspace(-0.5 * size_z, 0.5 * size_z, nz + 1, dtype=np.float64)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    Y = np.zeros_like(X) if height is None else height(X, Z)
    v = np.stack([X, Y, Z], axis=-1).reshape(-1, 3).astype(f4)
    uv = np.stack([(X / size_x + 0.5) * uv_scale, (Z / size_z + 0.5) * uv_scale], axis=-1).reshape(-1, 2).astype(f4)
    i, j = np.meshgrid(np.arange(nx), np.arange(nz), indexing="ij")
    v00 = (i * (nz + 1) + j).reshape(-1)
    v01, v10, v11 = v00 + 1, v00 + (nz + 1), v00 + (nz + 1) + 1
    tris = np.empty((nx * nz * 2, 3), dtype=np.uint32)
    tris[0::2] = np.stack([v00, v01, v10], axis=1)
    tris[1::2] = np.stack([v10, v01, v11], axis=1)
    # smooth normals from the height field gradient
    if height is None:
        n = np.tile(np.array([0, 1, 0], dtype=f4), (v.shape[0], 1))
    else:
        gx = np.gradient(Y, xs, axis=0)
        gz = np.gradient(Y, zs, axis=1)
        n = np.stack([-gx, np.ones_like(gx), -gz], axis=-1).reshape(-1, 3)
        n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(f4)
    return v, tris, uv, n


def uv_sphere_mesh(res=64):
    """Unit sphere, `res` longitudes x res/2 latitudes, smooth normals, uv."""
    nlat, nlon = max(res // 2, 2), max(res, 3)
    th = np.linspace(0.0, np.pi, nlat + 1)
    ph = np.linspace(0.0, 2.0 * np.pi, nlon + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    v = np.stack([np.sin(T) * np.cos(P), np.cos(T), np.sin(T) * np.sin(P)], axis=-1).reshape(-1, 3).astype(f4)
    uv = np.stack([P / (2 * np.pi), 1.0 - T / np.pi], axis=-1).reshape(-1, 2).astype(f4)
    i, j = np.meshgrid(np.arange(nlat), np.arange(nlon), indexing="ij")
    a = (i * (nlon + 1) + j).reshape(-1)
    b, c, d = a + 1, a + (nlon + 1), a + (nlon + 1) + 1
    tris = np.concatenate([np.stack([a, b, c], 1), np.stack([c, b, d], 1)]).astype(np.uint32)
    # drop degenerate triangles at the poles
    vv = v.astype(np.float64)
    area = np.linalg.norm(np.cross(vv[tris[:, 1]] - vv[tris[:, 0]], vv[tris[:, 2]] - vv[tris[:, 0]]), axis=1)
    tris = tris[area > 1e-9]
    tris, _ = _orient_outward(v, tris, center=(0, 0, 0))
    return v, tris, uv, v.copy()


def torus_mesh(major_res=128, minor_res=64, R=1.0, r=0.35):
    ua = np.linspace(0.0, 2.0 * np.pi, major_res + 1)
    va = np.linspace(0.0, 2.0 * np.pi, minor_res + 1)
    U, V = np.meshgrid(ua, va, indexing="ij")
    cx, cz = np.cos(U), np.sin(U)
    v = np.stack([(R + r * np.cos(V)) * cx, r * np.sin(V), (R + r * np.cos(V)) * cz], axis=-1).reshape(-1, 3).astype(f4)
    n = np.stack([np.cos(V) * cx, np.sin(V), np.cos(V) * cz], axis=-1).reshape(-1, 3).astype(f4)
    uv = np.stack([U / (2 * np.pi), V / (2 * np.pi)], axis=-1).reshape(-1, 2).astype(f4)
    i, j = np.meshgrid(np.arange(major_res), np.arange(minor_res), indexing="ij")
    a = (i * (minor_res + 1) + j).reshape(-1)
    b, c, d = a + 1, a + (minor_res + 1), a + (minor_res + 1) + 1
    tris = np.concatenate([np.stack([a, b, c], 1), np.stack([c, b, d], 1)]).astype(np.uint32)
    # orient by the analytic normal
    vv, nn = v.astype(np.float64), n.astype(np.float64)
    fn = np.cross(vv[tris[:, 1]] - vv[tris[:, 0]], vv[tris[:, 2]] - vv[tris[:, 0]])
    flip = (fn * nn[tris[:, 0]]).sum(1) < 0
    tris[flip] = tris[flip][:, [0, 2, 1]]
    return v, tris, uv, n


def cylinder_mesh(res=64, radius=0.5, height=1.0, rings=None):
    """Capped cylinder: `res` segments around, `rings` segments along the axis (default res/4: roughly square quads --
    full-height side triangles would be unsplittable for the reference's builder and end up in one huge leaf)."""
    rings = max(res // 4, 1) if rings is None else rings
    ang = np.linspace(0.0, 2.0 * np.pi, res + 1)
    ys = np.linspace(-0.5 * height, 0.5 * height, rings + 1)
    A, Y = np.meshgrid(ang, ys, indexing="ij")
    side_v = np.stack([radius * np.cos(A), Y, radius * np.sin(A)], axis=-1).reshape(-1, 3)
    side_n = np.stack([np.cos(A), np.zeros_like(A), np.sin(A)], axis=-1).reshape(-1, 3)
    side_uv = np.stack([A / (2 * np.pi), (Y / height) + 0.5], axis=-1).reshape(-1, 2)
    ring = np.stack([radius * np.cos(ang), np.zeros_like(ang), radius * np.sin(ang)], axis=-1)
    bot, top = ring.copy(), ring.copy()
    bot[:, 1], top[:, 1] = -0.5 * height, 0.5 * height
    cap_uv = np.stack([0.5 + 0.5 * np.cos(ang), 0.5 + 0.5 * np.sin(ang)], -1)
    ns = side_v.shape[0]
    k = res + 1
    v = np.concatenate([side_v, bot, top, [[0, -0.5 * height, 0]], [[0, 0.5 * height, 0]]]).astype(f4)
    n = np.concatenate([side_n, np.tile([0, -1, 0], (k, 1)), np.tile([0, 1, 0], (k, 1)), [[0, -1, 0]], [[0, 1, 0]]]).astype(f4)
    uv = np.concatenate([side_uv, cap_uv, cap_uv, [[0.5, 0.5]], [[0.5, 0.5]]]).astype(f4)
    i, j = np.meshgrid(np.arange(res), np.arange(rings), indexing="ij")
    a = (i * (rings + 1) + j).reshape(-1)
    b, c, d = a + 1, a + (rings + 1), a + (rings + 1) + 1
    side = np.concatenate([np.stack([a, b, c], 1), np.stack([c, b, d], 1)])
    jj = np.arange(res)
    cb = np.stack([np.full(res, ns + 2 * k), ns + jj, ns + jj + 1], 1)
    ct = np.stack([np.full(res, ns + 2 * k + 1), ns + k + jj, ns + k + jj + 1], 1)
    tris = np.concatenate([side, cb, ct]).astype(np.uint32)
    vv, nn = v.astype(np.float64), n.astype(np.float64)
    fn = np.cross(vv[tris[:, 1]] - vv[tris[:, 0]], vv[tris[:, 2]] - vv[tris[:, 0]])
    flip = (fn * nn[tris[:, 1]]).sum(1) < 0
    tris[flip] = tris[flip][:, [0, 2, 1]]
    return v, tris, uv, n


def lcg_uniform(n, seed):
    """Deterministic uniform [0,1) stream (numpy PCG64 with a fixed seed)."""
    return np.random.Generator(np.random.PCG64(int(seed))).random(n)


def heightfield_mesh(nx, nz, size=20.0, amplitude=0.6, jitter=0.02, seed=0xB200):
    """Displaced grid: deterministic sin height field + seeded jitter (config 3)."""
    noise = lcg_uniform((nx + 1) * (nz + 1), seed).reshape(nx + 1, nz + 1)

    def h(X, Z):
        return amplitude * (np.sin(X * 0.9) * np.cos(Z * 0.7) + 0.35 * np.sin(X * 2.3 + Z * 1.7)
                            + 0.15 * np.sin(X * 5.1) * np.sin(Z * 4.3)) + jitter * (noise - 0.5)

    return grid_mesh(nx, nz, size, size, height=h, uv_scale=8.0)


# ---------------------------------------------------------------------- procedural maps
def checker_texture(size=256, cells=8, seed=1):
    rng = np.random.Generator(np.random.PCG64(seed))
    y, x = np.mgrid[0:size, 0:size]
    c = ((x * cells // size) + (y * cells // size)) % 2
    base = np.where(c[..., None] == 0, np.array([230, 225, 210]), np.array([70, 110, 160]))
    noise = rng.integers(-12, 13, size=(size, size, 1))
    rgb = np.clip(base + noise, 0, 255).astype(np.uint8)
    return np.concatenate([rgb, np.full((size, size, 1), 255, np.uint8)], axis=-1)


def bump_normal_map(size=256, waves=6.0, strength=0.6):
    y, x = np.mgrid[0:size, 0:size].astype(np.float64) / size
    dx = strength * np.cos(x * 2 * np.pi * waves) * 0.5
    dy = strength * np.cos(y * 2 * np.pi * waves) * 0.5
    n = np.stack([-dx, -dy, np.ones_like(dx)], axis=-1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    rgb = np.clip((n * 0.5 + 0.5) * 255.0, 0, 255).astype(np.uint8)
    return np.concatenate([rgb, np.full((size, size, 1), 255, np.uint8)], axis=-1)


def roughness_map(size=256, seed=2):
    rng = np.random.Generator(np.random.PCG64(seed))
    y, x = np.mgrid[0:size, 0:size].astype(np.float64) / size
    r = 0.35 + 0.3 * np.sin(x * 12.0) * np.sin(y * 9.0) + 0.05 * rng.random((size, size))
    return np.clip(r * 255.0, 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------- config 1: Cornell box
def cornell(resolution=(512, 512)) -> World:
    """Config 1: Cornell box from generated quads and cubes, one emissive panel, all diffuse, no lights."""
    w = World()
    white = w.create_material("white", color=(200, 200, 200, 255), roughness=1.0)
    red = w.create_material("red", color=(200, 40, 40, 255), roughness=1.0)
    green = w.create_material("green", color=(40, 200, 40, 255), roughness=1.0)
    lamp = w.create_material("lamp", color=(255, 255, 255, 255), roughness=1.0, emission=60.0)
    blue = w.create_material("blue", color=(90, 110, 220, 255), roughness=1.0)
    S = 2.0

    def wall(name, p0, eu, ev, mat):
        v, t, uv = quad_mesh(p0, eu, ev)
        # front face must look into the box centre (0, 1, 0)
        t2, flip = _orient_outward(v, t, center=(0.0, 1.0, 0.0))
        t2 = t2[:, [0, 2, 1]]  # inward
        m = w.create_mesh(name, v, t2, texcrds=uv, tri_texcrds=t2)
        w.create_instance(name, m, [mat])

    wall("floor", (-1, 0, -1), (S, 0, 0), (0, 0, S), white)
    wall("ceiling", (-1, 2, -1), (S, 0, 0), (0, 0, S), white)
    wall("back", (-1, 0, 1), (S, 0, 0), (0, S, 0), white)
    wall("left", (-1, 0, -1), (0, 0, S), (0, S, 0), red)
    wall("right", (1, 0, -1), (0, 0, S), (0, S, 0), green)
    wall("panel", (-0.35, 1.995, -0.35), (0.7, 0, 0), (0, 0, 0.7), lamp)
    bv, bt, buv, btuv = box_mesh()
    cube = w.create_mesh("cube", bv, bt, texcrds=buv, tri_texcrds=btuv)
    w.create_instance("tall", cube, [white], position=(-0.35, 0.6, 0.3), rotation=(0.0, 0.3, 0.0), scale=(0.6, 1.2, 0.6))
    w.create_instance("short", cube, [blue], position=(0.4, 0.3, -0.25), rotation=(0.0, -0.35, 0.0), scale=(0.6, 0.6, 0.6))
    w.create_camera(name="cam", position=(0.0, 1.0, -4.4), rotation=(0.0, 0.0, 0.0), resolution=resolution, fov=0.75,
                    near_far=(0.01, 100.0), focal_distance=4.4, aperture=0.005, exposure_time=0.05)
    # the world material is the medium camera rays start in: fully transparent (alpha 0), no sky light
    w.world_material.color = (255, 255, 255, 0)
    w.world_material.emission = 0.0
    return w


# ---------------------------------------------------------------------- config 2: materials + lights
def materials_scene(resolution=(1920, 1080), res=64, cpu_comparable=False) -> World:
    """Config 2: UV-sphere / torus / cylinder with mirror, glossy, refractive and scattering materials,
    one direct and two spot lights (NEE + MIS). `cpu_comparable` drops what the CPU engine cannot do the same way
    (scattering medium; SURVEY.md §8a divergences) so the image can be compared with cpu_engine_kernel."""
    w = World()
    ground = w.create_material("ground", color=(180, 180, 170, 255), roughness=0.9)
    mirror = w.create_material("mirror", color=(240, 240, 240, 255), metalness=0.9, roughness=0.0)
    glossy = w.create_material("glossy", color=(200, 60, 50, 255), metalness=0.0, roughness=0.1, ior=1.5)
    glass = w.create_material("glass", color=(255, 255, 255, 0), ior=1.45)
    fog = w.create_material("fog", color=(235, 235, 255, 0), ior=1.0, scattering=0.0 if cpu_comparable else 0.5)
    gold = w.create_material("gold", color=(255, 200, 80, 255), metalness=1.0, roughness=0.25)
    # a barely undulating ground: an exactly flat mesh makes the reference's builder keep all triangles in one leaf
    gv, gt, guv, gn = grid_mesh(16, 16, 24.0, 24.0, height=lambda X, Z: 0.004 * np.sin(X * 1.3) * np.cos(Z * 1.1))
    gm = w.create_mesh("ground", gv, gt, texcrds=guv, normals=gn)
    w.create_instance("ground", gm, [ground])
    sv, st, suv, sn = uv_sphere_mesh(res)
    sm = w.create_mesh("sphere", sv, st, texcrds=suv, normals=sn)
    tv, tt, tuv, tn = torus_mesh(2 * res, res)
    tm = w.create_mesh("torus", tv, tt, texcrds=tuv, normals=tn)
    cv, ct, cuv, cn = cylinder_mesh(res)
    cm = w.create_mesh("cylinder", cv, ct, texcrds=cuv, normals=cn)
    w.create_instance("mirror ball", sm, [mirror], position=(-2.6, 1.0, 1.0))
    w.create_instance("glossy torus", tm, [glossy], position=(0.0, 0.6, 0.0), rotation=(0.5, 0.2, 0.0), scale=(1.2, 1.2, 1.2))
    w.create_instance("glass cylinder", cm, [glass], position=(2.6, 1.0, 0.6), scale=(1.4, 2.0, 1.4))
    w.create_instance("fog ball", sm, [fog], position=(1.0, 0.7, -2.2), scale=(0.7, 0.7, 0.7))
    w.create_instance("gold ball", sm, [gold], position=(-1.0, 0.5, -2.0), scale=(0.5, 0.5, 0.5))
    w.create_direct_light("sun", direction=(-0.4, -1.0, 0.5), color=(255, 244, 230), emission=1500.0, angular_size=0.02)
    w.create_spot_light("spot a", position=(-3.0, 5.0, -3.0), direction=(0.5, -1.0, 0.5), color=(255, 255, 255),
                        size=0.5, emission=400.0, beam_angle=0.6)
    w.create_spot_light("spot b", position=(4.0, 4.0, -2.0), direction=(-0.7, -1.0, 0.3), color=(200, 220, 255),
                        size=0.5, emission=300.0, beam_angle=0.6)
    w.create_camera(name="cam", position=(0.0, 3.0, -8.5), rotation=(-0.2, 0.0, 0.0), resolution=resolution, fov=1.15,
                    near_far=(0.01, 1000.0), focal_distance=8.5, aperture=0.01, exposure_time=0.006)
    # transparent medium (alpha 0); its colour times emission is the sky radiance (cuda_render_kernel.cu:174-186)
    w.world_material.color = (255, 255, 255, 0)
    w.world_material.emission = 1.0
    return w


# ---------------------------------------------------------------------- config 3: 1M-triangle textured height field
def heightfield_scene(resolution=(1920, 1080), nx=708, nz=707, map_size=2048, with_maps=True) -> World:
    """Config 3: one procedural mesh of 2*nx*nz triangles (708 x 707 -> 1,001,112) with texture, normal and
    roughness maps, camera with depth of field."""
    w = World()
    kw = {}
    if with_maps:
        kw["texture"] = w.create_map("texture", "terrain texture", checker_texture(map_size, 64, 1))
        kw["normal_map"] = w.create_map("normal", "terrain normal", bump_normal_map(map_size, 48.0, 0.6))
        kw["roughness_map"] = w.create_map("roughness", "terrain roughness", roughness_map(map_size, 2))
    terrain = w.create_material("terrain", color=(255, 255, 255, 255), metalness=0.0, roughness=0.6, **kw)
    v, t, uv, n = heightfield_mesh(nx, nz)
    m = w.create_mesh("terrain", v, t, texcrds=uv, normals=n)
    w.create_instance("terrain", m, [terrain])
    w.create_direct_light("sun", direction=(-0.5, -1.0, 0.3), color=(255, 240, 220), emission=1200.0, angular_size=0.03)
    w.create_camera(name="cam", position=(0.0, 4.5, -11.0), rotation=(-0.3, 0.0, 0.0), resolution=resolution, fov=1.0,
                    near_far=(0.01, 1000.0), focal_distance=10.0, aperture=0.05, exposure_time=0.0004)
    w.world_material.color = (255, 255, 255, 0)
    w.world_material.emission = 1.5
    return w


# ---------------------------------------------------------------------- config 4: instancing stress
def instancing_scene(resolution=(1920, 1080), n_instances=100, nx=224, nz=224, seed=4) -> World:
    """Config 4: n_instances x (2*nx*nz)-triangle mesh (100 x 100,352 = 10.0M effective triangles) with varied
    position / rotation / non-uniform scale."""
    w = World()
    rng = np.random.Generator(np.random.PCG64(seed))
    mats = [w.create_material("m%d" % i, color=tuple(int(x) for x in rng.integers(60, 255, 3)) + (255,),
                              metalness=float(rng.random() < 0.3) * 0.8, roughness=float(rng.uniform(0.05, 0.9)))
            for i in range(8)]
    v, t, uv, n = heightfield_mesh(nx, nz, size=2.0, amplitude=0.12, jitter=0.004, seed=seed)
    m = w.create_mesh("tile", v, t, texcrds=uv, normals=n)
    side = int(np.ceil(np.sqrt(n_instances)))
    for i in range(n_instances):
        gx, gz = i % side, i // side
        pos = ((gx - 0.5 * (side - 1)) * 2.3 + rng.uniform(-0.2, 0.2), rng.uniform(-0.3, 0.3),
               (gz - 0.5 * (side - 1)) * 2.3 + rng.uniform(-0.2, 0.2))
        rot = (rng.uniform(-0.25, 0.25), rng.uniform(0, 6.283), rng.uniform(-0.25, 0.25))
        scl = (rng.uniform(0.8, 1.3), rng.uniform(0.6, 2.0), rng.uniform(0.8, 1.3))
        w.create_instance("tile %d" % i, m, [mats[i % len(mats)]], position=pos, rotation=rot, scale=scl)
    w.create_direct_light("sun", direction=(-0.3, -1.0, 0.4), color=(255, 245, 230), emission=1000.0, angular_size=0.03)
    w.create_camera(name="cam", position=(0.0, 9.0, -0.62 * side * 2.3 - 6.0), rotation=(-0.5, 0.0, 0.0),
                    resolution=resolution, fov=1.0, near_far=(0.01, 1000.0), focal_distance=14.0, aperture=0.001,
                    exposure_time=0.0004)
    w.world_material.color = (255, 255, 255, 0)
    w.world_material.emission = 1.5
    return w


CONFIGS = {
    "cornell": cornell,
    "materials": materials_scene,
    "heightfield_1m": heightfield_scene,
    "instancing_10m": instancing_scene,
}
