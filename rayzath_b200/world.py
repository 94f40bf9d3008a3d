"""Python host-side world for the B200 render path: a small mirror of the reference's scene graph
(`RayZath::Engine::World`, /root/reference/RayZath/world.hpp) with the same object kinds and setter
semantics, which

  * flattens itself into the C-ABI arrays of include/rzb200.h exactly like the C++ drop-in does from the
    reference's own World (rayzath_b200/host/world_flatten.hpp) -- tests/test_flatten.py checks the two
    byte for byte against a dump of the reference World;
  * saves itself in the reference's scene format (`scene.json` + `.obj` + `.png`; json_loader.cpp:1064-1097,
    loader.cpp:735-1030) so the reference's CPU engine renders the identical scene.

All geometry is float32 and every derived quantity (axes, boxes, face normals, BVHs) comes from the
rounding-exact host utilities of librzb200.so (csrc/rzb_host_utils.cpp, csrc/rzb_bvh_build.cpp).
"""
from __future__ import annotations

import json
import os
import struct
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import capi

f4 = np.float32
EPS = float(np.finfo(np.float32).eps)
NPOS = 0xFFFFFFFF


def _normalize_f32(v: np.ndarray) -> np.ndarray:
    """Math::vec3::Normalize as restated in oracle/shim/vec3.h: v / sqrtf(x*x + y*y + z*z), all fp32."""
    v = np.asarray(v, dtype=f4).reshape(-1, 3)
    m = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2], dtype=f4)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (v / m[:, None]).astype(f4)


@dataclass
class Map:
    name: str
    kind: str                     # texture | normal | metalness | roughness | emission
    pixels: np.ndarray            # as held by the reference's World after loading ([h,w,4] u8, [h,w] u8 or [h,w] f32)
    filter: int = capi.FILTER_POINT
    address: int = capi.ADDRESS_WRAP
    scale: Sequence[float] = (1.0, 1.0)
    rotation: float = 0.0
    translation: Sequence[float] = (0.0, 0.0)
    index: int = 0


@dataclass
class Material:
    name: str
    color: Sequence[int] = (200, 200, 200, 255)   # RGBA u8, alpha 255 = opaque
    metalness: float = 0.0
    roughness: float = 0.0
    emission: float = 0.0
    ior: float = 1.0
    scattering: float = 0.0
    texture: Optional[Map] = None
    normal_map: Optional[Map] = None
    metalness_map: Optional[Map] = None
    roughness_map: Optional[Map] = None
    emission_map: Optional[Map] = None
    index: int = 0


@dataclass
class Mesh:
    name: str
    vertices: np.ndarray                       # [nv,3] f32
    tris: np.ndarray                           # [nt,3] u32 vertex ids
    texcrds: Optional[np.ndarray] = None       # [ntc,2] f32
    tri_texcrds: Optional[np.ndarray] = None   # [nt,3] u32
    normals: Optional[np.ndarray] = None       # [nn,3] f32 (normalised at creation like Mesh::createNormal)
    tri_normals: Optional[np.ndarray] = None   # [nt,3] u32
    tri_materials: Optional[np.ndarray] = None  # [nt] u32
    index: int = 0
    normals_input: Optional[np.ndarray] = None  # what was passed to createNormal (written to scene files)
    _bvh: Optional[tuple] = None
    # "reference" = the reference's own tree (parity mode); ("sah", max_leaf) = the optional SAH builder;
    # ("lbvh", max_leaf) = the optional GPU linear-BVH builder; ("sah4", max_leaf) = the SAH tree, collapsed to a 4-ary
    # tree by rzb_set_scene (RZB_SCENE_WIDE_TREES)
    bvh_builder: object = "reference"

    def bvh(self):
        if self._bvh is None or self._bvh[0] != self.bvh_builder:
            if self.bvh_builder == "reference":
                built = capi.build_mesh_bvh(self.vertices, self.tris)
            elif self.bvh_builder[0] == "lbvh":
                built = capi.build_mesh_bvh_lbvh(self.vertices, self.tris, int(self.bvh_builder[1]))
            else:
                built = capi.build_mesh_bvh_sah(self.vertices, self.tris, int(self.bvh_builder[1]))
            self._bvh = (self.bvh_builder, built)
        return self._bvh[1]


@dataclass
class Instance:
    name: str
    mesh: Optional[Mesh]
    materials: List[Optional[Material]]
    position: Sequence[float] = (0.0, 0.0, 0.0)
    rotation: Sequence[float] = (0.0, 0.0, 0.0)
    scale: Sequence[float] = (1.0, 1.0, 1.0)
    index: int = 0


@dataclass
class Camera:
    name: str = "camera"
    position: Sequence[float] = (0.0, 0.0, -10.0)
    rotation: Sequence[float] = (0.0, 0.0, 0.0)
    resolution: Sequence[int] = (1280, 720)
    fov: float = 1.5707964
    near_far: Sequence[float] = (1.0e-2, 1.0e3)
    focal_distance: float = 10.0
    aperture: float = 0.02
    exposure_time: float = 1.0 / 60.0
    temporal_blend: float = 0.75


@dataclass
class DirectLight:
    name: str
    direction: Sequence[float] = (0.0, -1.0, 0.0)
    color: Sequence[int] = (255, 255, 255)
    emission: float = 100.0
    angular_size: float = 0.1


@dataclass
class SpotLight:
    name: str
    position: Sequence[float] = (0.0, 5.0, 0.0)
    direction: Sequence[float] = (0.0, -1.0, 0.0)
    color: Sequence[int] = (255, 255, 255)
    size: float = 0.5
    emission: float = 100.0
    beam_angle: float = 1.0


class World:
    """Container of scene objects in creation order (= the reference's container index order)."""

    def __init__(self):
        self.maps: Dict[str, List[Map]] = {k: [] for k in ("texture", "normal", "metalness", "roughness", "emission")}
        self.materials: List[Material] = []
        self.meshes: List[Mesh] = []
        self.instances: List[Instance] = []
        self.cameras: List[Camera] = []
        self.direct_lights: List[DirectLight] = []
        self.spot_lights: List[SpotLight] = []
        self.world_material = Material("world", color=(255, 255, 255, 0), ior=1.0)  # world.cpp:33-38
        self.default_material = Material("default", color=(192, 192, 192, 255), ior=1.0)

    # -- creation (names follow World::container<T>().create / Mesh::create*)
    def create_map(self, kind: str, name: str, pixels: np.ndarray, **kw) -> Map:
        m = Map(name, kind, np.ascontiguousarray(pixels), **kw)
        m.index = len(self.maps[kind])
        self.maps[kind].append(m)
        return m

    def create_material(self, name: str, **kw) -> Material:
        m = Material(name, **kw)
        m.index = len(self.materials)
        self.materials.append(m)
        return m

    def create_mesh(self, name: str, vertices, tris, texcrds=None, tri_texcrds=None, normals=None, tri_normals=None,
                    tri_materials=None) -> Mesh:
        m = Mesh(name, np.ascontiguousarray(vertices, dtype=f4).reshape(-1, 3),
                 np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1, 3))
        if texcrds is not None:
            m.texcrds = np.ascontiguousarray(texcrds, dtype=f4).reshape(-1, 2)
            m.tri_texcrds = np.ascontiguousarray(tri_texcrds if tri_texcrds is not None else tris, dtype=np.uint32).reshape(-1, 3)
        if normals is not None:
            m.normals_input = np.ascontiguousarray(normals, dtype=f4).reshape(-1, 3)
            m.normals = _normalize_f32(m.normals_input)  # Mesh::createNormal stores normal.Normalized() (mesh.cpp:76-79)
            m.tri_normals = np.ascontiguousarray(tri_normals if tri_normals is not None else tris, dtype=np.uint32).reshape(-1, 3)
        if tri_materials is not None:
            m.tri_materials = np.ascontiguousarray(tri_materials, dtype=np.uint32).reshape(-1)
        m.index = len(self.meshes)
        self.meshes.append(m)
        return m

    def create_instance(self, name: str, mesh: Optional[Mesh], materials, position=(0, 0, 0), rotation=(0, 0, 0),
                        scale=(1, 1, 1)) -> Instance:
        if isinstance(materials, Material):
            materials = [materials]
        i = Instance(name, mesh, list(materials), tuple(map(float, position)), tuple(map(float, rotation)),
                     tuple(map(float, scale)))
        i.index = len(self.instances)
        self.instances.append(i)
        return i

    def create_camera(self, **kw) -> Camera:
        c = Camera(**kw)
        self.cameras.append(c)
        return c

    def create_direct_light(self, name: str, **kw) -> DirectLight:
        l = DirectLight(name, **kw)
        self.direct_lights.append(l)
        return l

    def create_spot_light(self, name: str, **kw) -> SpotLight:
        l = SpotLight(name, **kw)
        self.spot_lights.append(l)
        return l

    # ------------------------------------------------------------------ flatten -> C-ABI arrays
    def _flat_material(self, m: Material, map_base: Dict[str, int]) -> np.ndarray:
        r = np.zeros(1, dtype=capi.material_dtype)
        c = np.asarray(m.color, dtype=np.uint8)
        r["color"][0] = c.astype(f4) / f4(255.0)
        r["metalness"], r["roughness"], r["emission"] = f4(m.metalness), f4(m.roughness), f4(m.emission)
        r["ior"], r["scattering"] = f4(m.ior), f4(m.scattering)
        for fld, kind, mp in (("texture", "texture", m.texture), ("normal_map", "normal", m.normal_map),
                              ("metalness_map", "metalness", m.metalness_map),
                              ("roughness_map", "roughness", m.roughness_map),
                              ("emission_map", "emission", m.emission_map)):
            r[fld] = capi.NO_INDEX if mp is None else map_base[kind] + mp.index
        return r[0]

    def flatten(self) -> Dict[str, np.ndarray]:
        out: Dict[str, np.ndarray] = {}
        # maps: textures, normal maps, metalness, roughness, emission (WorldFlattener::run order)
        fmt = {"texture": capi.MAP_RGBA8, "normal": capi.MAP_RGBA8, "metalness": capi.MAP_R8,
               "roughness": capi.MAP_R8, "emission": capi.MAP_R32F}
        map_base, maps = {}, []
        for kind in ("texture", "normal", "metalness", "roughness", "emission"):
            map_base[kind] = len(maps)
            for mp in self.maps[kind]:
                r = np.zeros(1, dtype=capi.map_dtype)
                px = mp.pixels
                r["format"], r["height"], r["width"] = fmt[kind], px.shape[0], px.shape[1]
                r["filter"], r["address"] = mp.filter, mp.address
                r["scale"][0] = np.asarray(mp.scale, dtype=f4)
                r["rotation"] = f4(mp.rotation)
                r["translation"][0] = np.asarray(mp.translation, dtype=f4)
                out["map_pixels_%d" % len(maps)] = np.ascontiguousarray(px).reshape(-1).view(np.uint8)
                maps.append(r[0])
        out["maps"] = np.array(maps, dtype=capi.map_dtype) if maps else np.zeros(0, dtype=capi.map_dtype)

        mats = [self._flat_material(m, map_base) for m in self.materials]
        default_id = len(mats)
        mats.append(self._flat_material(self.default_material, map_base))
        out["materials"] = np.array(mats, dtype=capi.material_dtype)
        out["world_material"] = np.array([self._flat_material(self.world_material, map_base)], dtype=capi.material_dtype)
        out["default_material"] = np.array([default_id], dtype=np.uint32)

        # meshes
        all_nodes, all_tris, all_thi, mesh_recs = [], [], [], []
        node_off = tri_off = 0
        for mesh in self.meshes:
            nt = mesh.tris.shape[0]
            rec = np.zeros(1, dtype=capi.mesh_dtype)
            rec["node_offset"], rec["tri_offset"] = node_off, tri_off
            if nt:
                nodes, order = mesh.bvh()
                fn = capi.face_normals(mesh.vertices, mesh.tris)
                t = np.zeros(nt, dtype=capi.triangle_dtype)
                vid = mesh.tris[order]
                t["v"] = mesh.vertices[vid]
                t["face_normal"] = fn[order]
                if mesh.normals is not None:
                    t["n"] = mesh.normals[mesh.tri_normals[order]]
                else:
                    t["n"] = fn[order][:, None, :]
                if mesh.texcrds is not None:
                    t["uv"] = mesh.texcrds[mesh.tri_texcrds[order]]
                else:
                    t["uv"] = np.array([[0, 0], [0, 1], [1, 0]], dtype=f4)[None]
                if mesh.tri_materials is not None:
                    t["material_slot"] = mesh.tri_materials[order] & 0x3F
                all_nodes.append(nodes)
                all_tris.append(t)
                all_thi.append(order.astype(np.uint32))
                rec["node_count"], rec["tri_count"] = nodes.shape[0], nt
                node_off += nodes.shape[0]
                tri_off += nt
            mesh_recs.append(rec[0])
        out["mesh_nodes"] = np.concatenate(all_nodes) if all_nodes else np.zeros(0, dtype=capi.node_dtype)
        out["triangles"] = np.concatenate(all_tris) if all_tris else np.zeros(0, dtype=capi.triangle_dtype)
        out["tri_host_index"] = np.concatenate(all_thi) if all_thi else np.zeros(0, dtype=np.uint32)
        out["meshes"] = np.array(mesh_recs, dtype=capi.mesh_dtype) if mesh_recs else np.zeros(0, dtype=capi.mesh_dtype)
        if any(m.bvh_builder != "reference" for m in self.meshes):
            wide = any(isinstance(m.bvh_builder, tuple) and m.bvh_builder[0] == "sah4" for m in self.meshes)
            out["scene_flags"] = np.array([capi.SCENE_OWN_TREES | (capi.SCENE_WIDE_TREES if wide else 0)],
                                          dtype=np.uint32)  # absent = the reference's trees

        # lights
        dl = np.zeros(len(self.direct_lights), dtype=capi.direct_light_dtype)
        for i, l in enumerate(self.direct_lights):
            dl[i]["direction"] = _normalize_f32(l.direction)[0]
            dl[i]["angular_size"] = min(max(f4(l.angular_size), f4(0.0)), f4(np.pi))
            dl[i]["color"] = np.asarray(l.color[:3], dtype=np.uint8).astype(f4) / f4(255.0)
            dl[i]["emission"] = max(f4(l.emission), f4(0.0))
        out["direct_lights"] = dl
        sl = np.zeros(len(self.spot_lights), dtype=capi.spot_light_dtype)
        for i, l in enumerate(self.spot_lights):
            sl[i]["position"] = np.asarray(l.position, dtype=f4)
            sl[i]["direction"] = _normalize_f32(l.direction)[0]
            sl[i]["size"] = max(f4(l.size), np.finfo(f4).tiny)
            sl[i]["beam_angle"] = min(max(f4(l.beam_angle), f4(0.0)), f4(3.14159))
            sl[i]["color"] = np.asarray(l.color[:3], dtype=np.uint8).astype(f4) / f4(255.0)
            sl[i]["emission"] = max(f4(l.emission), f4(0.0))
        out["spot_lights"] = sl

        # instances in top-level BVH order
        n = len(self.instances)
        recs = np.zeros(n, dtype=capi.instance_dtype)
        inst_mats: List[int] = []
        boxes = np.zeros((n, 6), dtype=f4)
        tmp = []
        for i, inst in enumerate(self.instances):
            r = np.zeros(1, dtype=capi.instance_dtype)[0]
            axes = capi.rotation_axes(inst.rotation, 0)
            r["position"] = np.asarray(inst.position, dtype=f4)
            r["scale"] = np.asarray(inst.scale, dtype=f4)
            r["axis_x"], r["axis_y"], r["axis_z"] = axes[0], axes[1], axes[2]
            if inst.mesh is not None and inst.mesh.vertices.shape[0]:
                bb = capi.instance_bbox(inst.mesh.vertices, r["position"], r["scale"], axes)
            else:
                bb = np.zeros(6, dtype=f4)
            r["bb_min"], r["bb_max"] = bb[:3], bb[3:]
            boxes[i] = bb
            r["mesh"] = capi.NO_INDEX if inst.mesh is None else inst.mesh.index
            r["host_index"] = i
            tmp.append(r)
        inodes, iorder = capi.build_instance_bvh(boxes) if n else (np.zeros(0, dtype=capi.node_dtype), np.zeros(0, np.uint32))
        for k, hi in enumerate(iorder):
            r = tmp[int(hi)].copy()
            inst = self.instances[int(hi)]
            used = 0
            for s, m in enumerate(inst.materials[:64]):
                if m is not None:
                    used = s + 1
            r["material_offset"] = len(inst_mats)
            r["material_count"] = used
            for s in range(used):
                m = inst.materials[s]
                inst_mats.append(default_id if m is None else m.index)
            recs[k] = r
        out["instance_nodes"] = inodes
        out["instances"] = recs
        out["instance_materials"] = np.array(inst_mats, dtype=np.uint32)
        return out

    def camera_struct(self, index: int = 0) -> np.ndarray:
        """rzb_camera of camera `index` (flattenCamera in host/world_flatten.hpp; clamps of camera.cpp:104-160)."""
        c = self.cameras[index]
        r = np.zeros(1, dtype=capi.camera_dtype)
        w, h = max(int(c.resolution[0]), 1), max(int(c.resolution[1]), 1)
        r["width"], r["height"] = w, h
        r["position"][0] = np.asarray(c.position, dtype=f4)
        axes = capi.rotation_axes(c.rotation, 1)
        r["axis_x"][0], r["axis_y"][0], r["axis_z"][0] = axes[0], axes[1], axes[2]
        fov = f4(c.fov)
        pi = f4(np.pi)
        eps = f4(EPS)
        if fov < eps:
            fov = eps
        elif fov > pi - eps:
            fov = pi - eps
        r["fov"] = fov
        near, far = f4(c.near_far[0]), f4(c.near_far[1])
        if near < eps:
            near = eps
        if far < near + eps:
            far = near + eps
        r["near_far"][0] = (near, far)
        r["focal_distance"] = max(f4(c.focal_distance), eps)
        r["aperture"] = max(f4(c.aperture), eps)
        r["exposure_time"] = max(f4(c.exposure_time), eps)
        r["temporal_blend"] = min(max(f4(c.temporal_blend), f4(0.0)), f4(1.0))
        r["raycast_pixel"][0] = (w // 2, h // 2)
        return r

    # ------------------------------------------------------------------ save in the reference's scene format
    def save_reference(self, directory: str, name: str = "scene", inline_limit: int = 2000) -> str:
        """Write <directory>/<name>.json (+ .obj per large mesh, .png per map). Returns the json path."""
        os.makedirs(directory, exist_ok=True)

        def fl(x):
            return float(f4(x))

        def v3(v):
            return [fl(v[0]), fl(v[1]), fl(v[2])]

        objects: Dict[str, list] = {}
        key_of = {"texture": "Texture", "normal": "NormalMap", "metalness": "MetalnessMap",
                  "roughness": "RoughnessMap", "emission": "EmissionMap"}
        filt = {capi.FILTER_POINT: "point", capi.FILTER_LINEAR: "linear"}
        addr = {capi.ADDRESS_WRAP: "wrap", capi.ADDRESS_CLAMP: "clamp", capi.ADDRESS_MIRROR: "mirror",
                capi.ADDRESS_BORDER: "border"}
        for kind, lst in self.maps.items():
            for mp in lst:
                if kind == "emission":
                    raise NotImplementedError("emission maps need .hdr files; not used by the parity scenes")
                fname = "%s_%s_%d.png" % (name, kind, mp.index)
                px = mp.pixels
                if kind == "normal":
                    px = px.copy()
                    px[..., 1] = (256 - px[..., 1].astype(np.int32)).astype(np.uint8)  # loader negates green (loader.cpp:54-66)
                write_png(os.path.join(directory, fname), px)
                objects.setdefault(key_of[kind], []).append({
                    "name": mp.name, "file": fname, "filter mode": filt[mp.filter], "address mode": addr[mp.address],
                    "scale": [fl(mp.scale[0]), fl(mp.scale[1])], "rotation": fl(mp.rotation),
                    "translation": [fl(mp.translation[0]), fl(mp.translation[1])]})

        def mat_json(m: Material, with_name=True):
            j = {"color": [int(x) for x in m.color], "metalness": fl(m.metalness), "roughness": fl(m.roughness),
                 "emission": fl(m.emission), "ior": fl(m.ior), "scattering": fl(m.scattering)}
            if with_name:
                j["name"] = m.name
            for key, mp in (("texture", m.texture), ("normal map", m.normal_map), ("metalness map", m.metalness_map),
                            ("roughness map", m.roughness_map), ("emission map", m.emission_map)):
                if mp is not None:
                    j[key] = mp.name
            return j

        objects["Material"] = [mat_json(m) for m in self.materials]

        meshes = []
        for mesh in self.meshes:
            nt = mesh.tris.shape[0]
            if nt > inline_limit:
                fname = "%s_%s.obj" % (name, mesh.name)
                write_obj(os.path.join(directory, fname), mesh)
                meshes.append({"file": fname})
                continue
            j = {"name": mesh.name, "vertices": [v3(v) for v in mesh.vertices]}
            if mesh.texcrds is not None:
                j["texcrds"] = [[fl(t[0]), fl(t[1])] for t in mesh.texcrds]
            if mesh.normals is not None:
                j["normals"] = [v3(nrm) for nrm in mesh.normals_input]
            tris = []
            for i in range(nt):
                t = {"v": [int(x) for x in mesh.tris[i]]}
                if mesh.texcrds is not None:
                    t["t"] = [int(x) for x in mesh.tri_texcrds[i]]
                if mesh.normals is not None:
                    t["n"] = [int(x) for x in mesh.tri_normals[i]]
                if mesh.tri_materials is not None:
                    t["m"] = int(mesh.tri_materials[i])
                tris.append(t)
            j["triangles"] = tris
            meshes.append(j)
        objects["Mesh"] = meshes

        objects["Camera"] = [{
            "name": c.name, "position": v3(c.position), "rotation": v3(c.rotation),
            "resolution": [int(c.resolution[0]), int(c.resolution[1])], "fov": fl(c.fov),
            "near far": [fl(c.near_far[0]), fl(c.near_far[1])], "focal distance": fl(c.focal_distance),
            "aperture": fl(c.aperture), "exposure time": fl(c.exposure_time), "temporal blend": fl(c.temporal_blend),
            "enabled": True} for c in self.cameras]
        objects["SpotLight"] = [{
            "name": l.name, "position": v3(l.position), "direction": v3(l.direction),
            "color": [int(x) for x in l.color[:3]], "size": fl(l.size), "emission": fl(l.emission),
            "angle": fl(l.beam_angle)} for l in self.spot_lights]
        objects["DirectLight"] = [{
            "name": l.name, "direction": v3(l.direction), "color": [int(x) for x in l.color[:3]],
            "emission": fl(l.emission), "size": fl(l.angular_size)} for l in self.direct_lights]
        insts = []
        for inst in self.instances:
            j = {"name": inst.name, "position": v3(inst.position), "rotation": v3(inst.rotation), "scale": v3(inst.scale)}
            if any(m is None for m in inst.materials):
                raise ValueError("the reference's scene format cannot express empty material slots")
            j["Material"] = [m.name for m in inst.materials]
            if inst.mesh is not None:
                j["Mesh"] = inst.mesh.name
            insts.append(j)
        objects["Instance"] = insts
        objects = {k: v for k, v in objects.items() if v}
        scene = {"Objects": objects, "Material": mat_json(self.world_material, False),
                 "DefaultMaterial": mat_json(self.default_material, False)}
        path = os.path.join(directory, name + ".json")
        with open(path, "w") as f:
            json.dump(scene, f)
        return path


# ---------------------------------------------------------------------- file writers
def write_png(path: str, pixels: np.ndarray) -> None:
    """Minimal PNG encoder (8-bit grey or RGBA), enough for stb_image to read back the exact bytes."""
    px = np.ascontiguousarray(pixels, dtype=np.uint8)
    if px.ndim == 2:
        h, w = px.shape
        color_type, row = 0, px.reshape(h, w)
    else:
        h, w, c = px.shape
        assert c == 4
        color_type, row = 6, px.reshape(h, w * 4)
    raw = np.zeros((h, row.shape[1] + 1), dtype=np.uint8)
    raw[:, 1:] = row

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, color_type, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw.tobytes(), 1)))
        f.write(chunk(b"IEND", b""))


def write_obj(path: str, mesh: Mesh) -> None:
    """Wavefront OBJ the reference's OBJLoader::parseOBJ turns back into exactly `mesh`:
    z is negated on load (loader.cpp:810,832) and a face `f a b c` becomes triangle (a, c, b) (:1005-1012)."""
    def fmt(a, sign):
        a = np.asarray(a, dtype=f4).astype(np.float64) * sign
        return a

    lines = ["o " + mesh.name]
    v = fmt(mesh.vertices, np.array([1.0, 1.0, -1.0]))
    lines.append("\n".join("v %.9g %.9g %.9g" % (a, b, c) for a, b, c in v))
    if mesh.texcrds is not None:
        lines.append("\n".join("vt %.9g %.9g" % (a, b) for a, b in mesh.texcrds.astype(np.float64)))
    if mesh.normals is not None:
        nrm = fmt(mesh.normals_input, np.array([1.0, 1.0, -1.0]))
        lines.append("\n".join("vn %.9g %.9g %.9g" % (a, b, c) for a, b, c in nrm))
    t = mesh.tris.astype(np.int64) + 1
    tt = mesh.tri_texcrds.astype(np.int64) + 1 if mesh.texcrds is not None else None
    tn = mesh.tri_normals.astype(np.int64) + 1 if mesh.normals is not None else None
    if mesh.tri_materials is not None and mesh.tri_materials.max(initial=0) != 0:
        raise NotImplementedError("per-triangle material slots in OBJ export")
    faces = []
    for i in range(t.shape[0]):
        parts = []
        for k in (0, 2, 1):
            s = str(t[i, k])
            if tt is not None or tn is not None:
                s += "/" + (str(tt[i, k]) if tt is not None else "")
                if tn is not None:
                    s += "/" + str(tn[i, k])
            parts.append(s)
        faces.append("f " + " ".join(parts))
    lines.append("\n".join(faces))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
