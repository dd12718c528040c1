"""Render / Scene / Frame descriptions and the loader for the reference's JSON format.

Host-side mirror of the reference's data model (src/rt.rs:10-190) and of the part of
src/parser.rs that turns a JSON description into it: serde defaults (parser.rs:188-271),
hex colours (parser.rs:713-733), inline base64+gzip textures / meshes (parser.rs:620-628,
674-682), texture image files (parser.rs:660-672), .obj meshes (parser.rs:602-618) and the
instance expansion rules (parser.rs:838-853).  `Render.pack()` flattens it into the C-ABI
structs of include/mrt.h.  Nothing here computes a pixel.
"""
from __future__ import annotations

import base64
import ctypes as C
import gzip
import json
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import abi


class SceneError(ValueError):
    """≙ the `Err(String)` of parser.rs:12-14."""


# ----------------------------------------------------------------------------- data model
@dataclass
class Texture:  # rt.rs:81-86
    w: int
    h: int
    dat: Optional[np.ndarray]  # (w*h, 3) float32, row-major

    def key(self):
        return (self.w, self.h, None if self.dat is None else self.dat.tobytes())


@dataclass
class Material:  # rt.rs:88-103, defaults parser.rs:242-259
    albedo: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    rough: float = 0.0
    metal: float = 0.0
    glass: float = 0.0
    opacity: float = 1.0
    emit: float = 0.0
    tex: Optional[Texture] = None
    rmap: Optional[Texture] = None
    mmap: Optional[Texture] = None
    gmap: Optional[Texture] = None
    omap: Optional[Texture] = None
    emap: Optional[Texture] = None


BACKWARD = (-0.0, -0.0, -1.0, -0.0)  # Vec4f::backward(), lin.rs:143-145, as (w, x, y, z)


@dataclass
class Renderer:  # rt.rs:152-158
    kind: str  # sphere | plane | box | triangle | mesh
    r: float = 0.0
    n: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    sizes: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    vtx: Optional[np.ndarray] = None   # (3,3) for triangle
    mesh: Optional[np.ndarray] = None  # (n,3,3) float32 for mesh
    mat: Material = field(default_factory=Material)
    instance: List[Tuple[Tuple[float, float, float], Tuple[float, float, float, float]]] = field(default_factory=list)
    name: Optional[str] = None


@dataclass
class Light:  # rt.rs:170-175, defaults parser.rs:261-271
    kind: str = "point"  # point | dir
    v: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    pwr: float = 0.5
    color: Tuple[float, float, float] = (1.0, 1.0, 1.0)


@dataclass
class Sky:  # rt.rs:177-181, defaults parser.rs:222-229
    color: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    pwr: float = 0.5


@dataclass
class Scene:  # rt.rs:183-190
    renderer: Optional[List[Renderer]] = None
    light: Optional[List[Light]] = None
    sky: Sky = field(default_factory=Sky)


@dataclass
class Camera:  # rt.rs:63-72, defaults parser.rs:198-210
    pos: Tuple[float, float, float] = (-0.0, -1.0, -0.0)
    dir: Tuple[float, float, float, float] = (0.0, 0.0, 1.0, 0.0)  # (w, x, y, z)
    fov: float = 70.0
    gamma: float = 0.8
    exp: float = 0.2
    aprt: float = 0.001
    foc: float = 100.0


@dataclass
class Frame:  # rt.rs:74-79, defaults parser.rs:212-220
    res: Tuple[int, int] = (1280, 720)
    ssaa: float = 1.0
    cam: Camera = field(default_factory=Camera)

    def film_size(self) -> Tuple[int, int]:
        """(nw, nh) of the supersampled film, sampler.rs:29-30 (f32 product, truncated)."""
        nw = int(np.float32(self.res[0]) * np.float32(self.ssaa))
        nh = int(np.float32(self.res[1]) * np.float32(self.ssaa))
        return nw, nh

    def pack(self) -> abi.MrtFrame:
        f = abi.MrtFrame()
        f.res[0], f.res[1] = int(self.res[0]), int(self.res[1])
        f.ssaa = float(self.ssaa)
        f.cam_pos[:] = [float(v) for v in self.cam.pos]
        f.cam_dir[:] = [float(v) for v in self.cam.dir]
        f.fov, f.gamma, f.exp = float(self.cam.fov), float(self.cam.gamma), float(self.cam.exp)
        f.aprt, f.foc = float(self.cam.aprt), float(self.cam.foc)
        return f


@dataclass
class RayTracer:  # rt.rs:16-22, defaults parser.rs:188-196
    bounce: int = 8
    sample: int = 16
    loss: float = 0.15


@dataclass
class Render:  # rt.rs:9-14
    rt: RayTracer = field(default_factory=RayTracer)
    frame: Frame = field(default_factory=Frame)
    scene: Scene = field(default_factory=Scene)


# ----------------------------------------------------------------------------- packing
class PackedScene:
    """Owns the ctypes arrays an `mrt_scene` points into."""

    KIND = {"sphere": abi.MRT_SPHERE, "plane": abi.MRT_PLANE, "box": abi.MRT_BOX,
            "triangle": abi.MRT_TRIANGLE, "mesh": abi.MRT_MESH}

    def __init__(self, scene: Scene):
        objs = scene.renderer or []
        lights = scene.light or []
        tex_ids = {}
        textures: List[Texture] = []

        def tex_id(t: Optional[Texture]) -> int:
            if t is None:
                return -1
            k = t.key()
            if k not in tex_ids:
                tex_ids[k] = len(textures)
                textures.append(t)
            return tex_ids[k]

        n_inst = sum(len(o.instance) for o in objs)
        self.objects = (abi.MrtObject * max(1, len(objs)))()
        self.instances = (abi.MrtInstance * max(1, n_inst))()
        meshes: List[np.ndarray] = []
        ii = 0
        for k, o in enumerate(objs):
            po = self.objects[k]
            po.kind = self.KIND[o.kind]
            if o.kind == "sphere":
                po.param[0] = float(o.r)
            elif o.kind == "plane":
                po.param[0:3] = [float(v) for v in o.n]
            elif o.kind == "box":
                po.param[0:3] = [float(v) for v in o.sizes]
            elif o.kind == "triangle":
                po.param[0:9] = [float(v) for v in np.asarray(o.vtx, dtype=np.float32).reshape(9)]
            else:
                po.mesh = len(meshes)
                meshes.append(np.ascontiguousarray(o.mesh, dtype=np.float32).reshape(-1, 9))
            po.first_inst, po.n_inst = ii, len(o.instance)
            for pos, d in o.instance:
                self.instances[ii].pos[:] = [float(v) for v in pos]
                self.instances[ii].dir[:] = [float(v) for v in d]
                ii += 1
            m, pm = o.mat, po.mat
            pm.albedo[:] = [float(v) for v in m.albedo]
            pm.rough, pm.metal, pm.glass = float(m.rough), float(m.metal), float(m.glass)
            pm.opacity, pm.emit = float(m.opacity), float(m.emit)
            pm.tex, pm.rmap, pm.mmap = tex_id(m.tex), tex_id(m.rmap), tex_id(m.mmap)
            pm.gmap, pm.omap, pm.emap = tex_id(m.gmap), tex_id(m.omap), tex_id(m.emap)

        self.textures = (abi.MrtTexture * max(1, len(textures)))()
        chunks = []
        off = 0
        for k, t in enumerate(textures):
            pt = self.textures[k]
            pt.w, pt.h, pt.first_texel = int(t.w), int(t.h), off
            pt.has_dat = 0 if t.dat is None else 1
            if t.dat is not None:
                d = np.ascontiguousarray(t.dat, dtype=np.float32).reshape(-1, 3)
                chunks.append(d)
                off += d.shape[0]
        self.texels = np.concatenate(chunks) if chunks else np.zeros((1, 3), np.float32)
        self.n_texels = off

        self.meshes = (abi.MrtMesh * max(1, len(meshes)))()
        t0 = 0
        for k, m in enumerate(meshes):
            self.meshes[k].first_tri, self.meshes[k].n_tri = t0, m.shape[0]
            t0 += m.shape[0]
        self.triangles = np.concatenate(meshes) if meshes else np.zeros((1, 9), np.float32)
        self.n_triangles = t0

        self.lights = (abi.MrtLight * max(1, len(lights)))()
        for k, l in enumerate(lights):
            pl = self.lights[k]
            pl.kind = abi.MRT_LIGHT_POINT if l.kind == "point" else abi.MRT_LIGHT_DIR
            pl.v[:] = [float(v) for v in l.v]
            pl.pwr = float(l.pwr)
            pl.color[:] = [float(v) for v in l.color]

        s = abi.MrtScene()
        s.objects, s.n_objects = self.objects, len(objs)
        s.instances, s.n_instances = self.instances, n_inst
        s.textures, s.n_textures = self.textures, len(textures)
        s.texels = self.texels.ctypes.data_as(C.POINTER(C.c_float))
        s.n_texels = self.n_texels
        s.meshes, s.n_meshes = self.meshes, len(meshes)
        s.triangles = self.triangles.ctypes.data_as(C.POINTER(C.c_float))
        s.n_triangles = self.n_triangles
        s.lights, s.n_lights = self.lights, len(lights)
        s.sky_color[:] = [float(v) for v in scene.sky.color]
        s.sky_pwr = float(scene.sky.pwr)
        self.c = s

    def objects_array(self):
        return [self.objects[k] for k in range(self.c.n_objects)]

    def instances_array(self):
        return [self.instances[k] for k in range(self.c.n_instances)]

    def lights_array(self):
        return [self.lights[k] for k in range(self.c.n_lights)]

    def nbytes(self) -> int:
        """Host bytes an mrt_set_scene call copies (for bench.py's h2d accounting)."""
        s = self.c
        return (s.n_objects * C.sizeof(abi.MrtObject) + s.n_instances * C.sizeof(abi.MrtInstance)
                + s.n_textures * C.sizeof(abi.MrtTexture) + self.n_texels * 12
                + s.n_meshes * C.sizeof(abi.MrtMesh) + self.n_triangles * 36
                + s.n_lights * C.sizeof(abi.MrtLight) + C.sizeof(abi.MrtScene))


def pack_scene(scene: Scene) -> PackedScene:
    return PackedScene(scene)


# ----------------------------------------------------------------------------- JSON loader
def _vec(v, n, what):
    if not isinstance(v, (list, tuple)) or len(v) != n:
        raise SceneError(f"{what}: expected {n} numbers")
    return tuple(float(x) for x in v)


def _color(v, what="color"):
    """ColorWrapper::unwrap, parser.rs:713-733."""
    if isinstance(v, str):
        if not v.startswith("#"):
            raise SceneError(f"{v} is not a hex color!")
        try:
            n = int(v[1:7], 16)
        except ValueError as e:
            raise SceneError(str(e))
        f32 = np.float32
        return tuple(float(f32((n >> s) & 0xFF) / f32(255.0)) for s in (16, 8, 0))
    return _vec(v, 3, what)


def _inline_json(s: str):
    try:
        return json.loads(gzip.decompress(base64.b64decode(s)).decode("utf-8"))
    except Exception as e:  # noqa: BLE001 - any decode error is a scene error
        raise SceneError(f"inline asset: {e}")


def _texture_buffer(d) -> Texture:
    w, h = int(d.get("w", 0)), int(d.get("h", 0))
    dat = d.get("dat")
    if dat is not None:
        dat = np.asarray(dat, dtype=np.float32).reshape(-1, 3)
    return Texture(w, h, dat)


def _texture_file(path: str, base_dir: Optional[str]) -> Texture:
    """TextureWrapper::load, parser.rs:660-672: RGB8 pixels / 255."""
    from PIL import Image
    p = path if os.path.isabs(path) or base_dir is None else os.path.join(base_dir, path)
    try:
        img = Image.open(p)
    except OSError as e:
        raise SceneError(str(e))
    if img.mode != "RGB":
        raise SceneError("is not rgb888 image!")
    a = np.asarray(img, dtype=np.uint8)
    dat = (a.reshape(-1, 3).astype(np.float32) / np.float32(255.0)).astype(np.float32)
    return Texture(img.size[0], img.size[1], dat)


def _texture(v, base_dir) -> Optional[Texture]:
    """TextureWrapper (untagged: buffer | inline base64 | file), parser.rs:86-92, 674-696."""
    if v is None:
        return None
    if isinstance(v, dict):
        return _texture_buffer(v)
    if isinstance(v, str):
        if "." in v:
            return _texture_file(v, base_dir)
        return _texture(_inline_json(v), base_dir)
    raise SceneError("bad texture")


def _mesh_obj(path: str, base_dir: Optional[str]) -> np.ndarray:
    """MeshWrapper::load, parser.rs:602-618: first object, first group, position indices,
    first three vertices of every polygon."""
    p = path if os.path.isabs(path) or base_dir is None else os.path.join(base_dir, path)
    pos, tris = [], []
    group_open = True
    n_groups = 0
    try:
        with open(p) as fh:
            for line in fh:
                t = line.split()
                if not t:
                    continue
                if t[0] == "v":
                    pos.append([float(t[1]), float(t[2]), float(t[3])])
                elif t[0] in ("o", "g"):
                    n_groups += 1
                    group_open = n_groups <= 1 or not tris
                elif t[0] == "f" and group_open:
                    idx = []
                    for tok in t[1:4]:
                        i = int(tok.split("/")[0])
                        idx.append(i - 1 if i > 0 else len(pos) + i)
                    tris.append(idx)
    except OSError as e:
        raise SceneError(str(e))
    pos = np.asarray(pos, dtype=np.float32)
    return pos[np.asarray(tris, dtype=np.int64)].astype(np.float32)


def _mesh(v, base_dir) -> np.ndarray:
    if isinstance(v, str):
        if "." in v:
            return _mesh_obj(v, base_dir)
        return _mesh(_inline_json(v), base_dir)
    return np.asarray(v, dtype=np.float32).reshape(-1, 3, 3)


def _material(d, base_dir) -> Material:
    d = d or {}
    m = Material()
    if "albedo" in d:
        m.albedo = _color(d["albedo"], "albedo")
    for k in ("rough", "metal", "glass", "opacity", "emit"):
        if k in d:
            setattr(m, k, float(d[k]))
    for k in ("tex", "rmap", "mmap", "gmap", "omap", "emap"):
        if d.get(k) is not None:
            setattr(m, k, _texture(d[k], base_dir))
    return m


def _renderer(d, base_dir) -> Renderer:
    """RendererWrapper + unwrap, parser.rs:130-150, 826-864."""
    kind = d.get("type")
    r = Renderer(kind=kind)
    if kind == "sphere":
        r.r = float(d["r"])
    elif kind == "plane":
        r.n = _vec(d["n"], 3, "n")
    elif kind == "box":
        r.sizes = _vec(d["sizes"], 3, "sizes")
    elif kind == "triangle":
        r.vtx = np.asarray(d["vtx"], dtype=np.float32).reshape(3, 3)
    elif kind == "mesh":
        r.mesh = _mesh(d["mesh"], base_dir)
    else:
        raise SceneError(f"unknown renderer type {kind!r}")
    r.mat = _material(d.get("mat"), base_dir)
    r.name = d.get("name")
    pos = d.get("pos")
    dr = d.get("dir")
    inst = d.get("inst")
    if inst is not None:
        lst = [(_vec(p, 3, "inst pos"), _vec(q, 4, "inst dir")) for p, q in inst]
        if pos is not None or dr is not None:  # parser.rs:841-843: prepended
            lst.insert(0, (_vec(pos, 3, "pos") if pos is not None else (0.0, 0.0, 0.0),
                           _vec(dr, 4, "dir") if dr is not None else BACKWARD))
        r.instance = lst
    else:
        r.instance = [(_vec(pos, 3, "pos") if pos is not None else (0.0, 0.0, 0.0),
                       _vec(dr, 4, "dir") if dr is not None else BACKWARD)]
    return r


def _light(d) -> Light:
    l = Light()
    kind = d.get("type")
    if kind == "point":
        l.kind, l.v = "point", _vec(d["pos"], 3, "pos")
    elif kind == "dir":
        l.kind, l.v = "dir", _vec(d["dir"], 3, "dir")
    else:
        raise SceneError(f"unknown light type {kind!r}")
    if "pwr" in d:
        l.pwr = float(d["pwr"])
    if "color" in d:
        l.color = _color(d["color"])
    return l


def render_from_dict(d: dict, base_dir: Optional[str] = None) -> Render:
    """RenderWrapper (serde, every key optional) → Render, parser.rs:160-166, 929-937."""
    out = Render()
    rt = d.get("rt") or {}
    out.rt = RayTracer(int(rt.get("bounce", 8)), int(rt.get("sample", 16)), float(rt.get("loss", 0.15)))
    fr = d.get("frame") or {}
    cam_d = fr.get("cam") or {}
    cam = Camera()
    if "pos" in cam_d:
        cam.pos = _vec(cam_d["pos"], 3, "cam pos")
    if "dir" in cam_d:
        cam.dir = _vec(cam_d["dir"], 4, "cam dir")
    for k in ("fov", "gamma", "exp", "aprt", "foc"):
        if k in cam_d:
            setattr(cam, k, float(cam_d[k]))
    res = fr.get("res", (1280, 720))
    if not (0 <= int(res[0]) <= 65535 and 0 <= int(res[1]) <= 65535):
        raise SceneError("res does not fit u16")
    out.frame = Frame((int(res[0]), int(res[1])), float(fr.get("ssaa", 1.0)), cam)
    sc = d.get("scene") or {}
    scene = Scene()
    if sc.get("renderer") is not None:
        scene.renderer = [_renderer(o, base_dir) for o in sc["renderer"]]
    if sc.get("light") is not None:
        scene.light = [_light(l) for l in sc["light"]]
    sky = sc.get("sky") or {}
    scene.sky = Sky(_color(sky["color"]) if "color" in sky else (0.0, 0.0, 0.0), float(sky.get("pwr", 0.5)))
    out.scene = scene
    return out


def load_render(path: str) -> Render:
    with open(path) as fh:
        try:
            d = json.load(fh)
        except json.JSONDecodeError as e:
            raise SceneError(str(e))
    return render_from_dict(d, os.path.dirname(os.path.abspath(path)))
