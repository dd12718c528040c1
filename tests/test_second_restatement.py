"""The C++ oracle against a SECOND restatement of the reference (tests/second_restatement.py: numpy, written from
/root/reference/src alone) on exactly the features the reference's published renders never reach — box-atlas uv,
directional lights, instance lists, the mesh octree — i.e. on Minecraft.json, Instance.json and Mesh.json.
Both are restatements of the same Rust source; where they agree ray for ray, only a misreading SHARED by two
independently written texts could still be wrong.  CPU only."""
import json
import os

import numpy as np
import pytest

import micro_raytracer_b200 as mrt
import oracle_lib
import second_restatement as sr
from util import SCENES, load


def _primary(name, res, aprt0=True, **rt):
    r = load(name, res, 1.0, **rt)
    if aprt0:
        r.frame.cam.aprt = 0.0
    cpu = oracle_lib.OracleSampler()
    cpu._bind(r.scene, r.frame, r.rt)
    h = cpu.trace_primary().reshape(-1)
    desc = json.load(open(os.path.join(SCENES, name + ".json")))
    objs = sr.load_objects(desc, SCENES)
    trees = {i: sr.build_octree(o[1]) for i, o in enumerate(objs) if o[0] == "mesh"}
    return r, cpu, h, desc, objs, trees


def test_instance_expansion_rules():
    """parser.rs:838-853 against the product's loader (scene.py) on the example that uses instance lists and on the
    corner cases: pos/dir given next to `inst` are PREPENDED as one more instance; defaults are (0,0,0) and backward."""
    for name in ("Instance", "Minecraft", "CornellBox2"):
        desc = json.load(open(os.path.join(SCENES, name + ".json")))
        r = load(name)
        for o, ro in zip(desc["scene"]["renderer"], r.scene.renderer):
            want = sr.expand_instances(o)
            assert len(want) == len(ro.instance)
            for (p, d), (rp, rd) in zip(want, ro.instance):
                assert np.array_equal(np.float32(p), np.float32(rp)) and np.array_equal(np.float32(d), np.float32(rd))
    o = {"type": "sphere", "r": 1, "pos": [1, 2, 3], "inst": [[[0, 0, 0], [0, 0, 1, 0]]]}
    r = mrt.render_from_dict({"scene": {"renderer": [o]}})
    assert sr.expand_instances(o) == [([1, 2, 3], [-0.0, -0.0, -1.0, -0.0]), ([0, 0, 0], [0, 0, 1, 0])]
    assert len(r.scene.renderer[0].instance) == 2 and tuple(r.scene.renderer[0].instance[0][0]) == (1.0, 2.0, 3.0)
    o = {"type": "sphere", "r": 1, "dir": [0.5, 0, 1, 0], "inst": []}
    assert sr.expand_instances(o) == [([0.0, 0.0, 0.0], [0.5, 0, 1, 0])]
    assert len(mrt.render_from_dict({"scene": {"renderer": [o]}}).scene.renderer[0].instance) == 1


@pytest.mark.parametrize("name,res", [("Minecraft", (160, 90)), ("Instance", (128, 72)), ("Mesh", (96, 54)), ("dof", (96, 54))])
def test_primary_hits_normals_uv(name, res):
    """closest_hit over (object, instance) order, every kind's intersect / normal / uv, per-instance transforms with the
    negated dir (incl. the rotated Minecraft torch), octree candidate sets — ray for ray."""
    r, cpu, h, desc, objs, trees = _primary(name, res)
    o, d = h["orig"].astype(np.float32), h["dir"].astype(np.float32)
    got = sr.closest_hit(objs, o, d, trees)
    hit_o = h["obj"] >= 0
    same = (got["hit"] == hit_o) & (~hit_o | ((got["obj"] == h["obj"]) & (got["inst"] == h["inst"])))
    assert same.mean() >= 0.9995, f"{name}: ids differ on {(~same).sum()} of {same.size} rays"
    m = same & hit_o
    assert m.sum() > 0.2 * m.size
    dt = np.abs(got["t0"][m] - h["t0"][m]) / np.maximum(1.0, np.abs(h["t0"][m]))
    assert (dt <= 1e-5).mean() >= 0.999 and dt.max() < 1e-3, (name, dt.max())
    dt1 = np.abs(got["t1"][m] - h["t1"][m]) / np.maximum(1.0, np.abs(h["t1"][m]))
    assert (dt1 <= 1e-5).mean() >= 0.999, (name, dt1.max())
    if name == "Mesh":
        mm = m & (h["tri0"] >= 0)
        assert mm.sum() > 200
        assert (got["tri0"][mm] == h["tri0"][mm]).mean() >= 0.999 and (got["tri1"][mm] == h["tri1"][mm]).mean() >= 0.999
    n = sr.hit_normal(objs, got, o, d)
    dn = np.abs(n[m] - h["n0"][m]).max(axis=1)
    assert (dn <= 1e-4).mean() >= 0.998, (name, np.sort(dn)[-5:])
    uv = sr.hit_uv(objs, got, o, d)
    nm = m & np.array([objs[k][0] != "mesh" for k in np.maximum(h["obj"], 0)])
    du = np.abs(uv[nm] - h["uv"][nm])
    du = np.minimum(du, 1.0 - du).max(axis=1)  # plane uv wraps
    assert (du <= 2e-4).mean() >= 0.998, (name, np.sort(du)[-5:])


def test_octree_membership_and_holes():
    """The depth-3 octree lists a triangle in a leaf iff one of its VERTICES lies inside the leaf (rt.rs:227-248):
    large triangles leave holes.  The second restatement's tree — built recursively as BVH::construct does — must give
    the same candidate sets as the oracle's flattened leaf grid: compared through the hits of rays shot from many
    directions, with the ancestors' boxes tested on the way down as rt.rs:707-723 does."""
    desc = json.load(open(os.path.join(SCENES, "Mesh.json")))
    objs = sr.load_objects(desc, SCENES)
    tris = objs[0][1]
    tree = sr.build_octree(tris)

    def leaves(nd):
        return [nd] if nd["content"] is not None else [l for k in (nd["childs"] or []) for l in leaves(k)]
    lv = leaves(tree)
    # 2 158 memberships in the reference's f32 arithmetic (SURVEY.md's 2 164 came from a float64 probe: six vertices sit
    # within an ulp of a leaf face)
    assert len(tris) == 967 and len(lv) == 144 and sum(len(l["content"]) for l in lv) == 2158 and max(len(l["content"]) for l in lv) <= 58
    listed = np.zeros(len(tris), bool)
    for l in lv:
        listed[l["content"]] = True
    assert listed.all()  # every triangle has its vertices somewhere
    rng = np.random.default_rng(5)
    n = 4000
    o = rng.normal(size=(n, 3)).astype(np.float32)
    o = o / np.linalg.norm(o, axis=1, keepdims=True) * np.float32(2.0) + np.float32([0, 0.5, 0])
    tgt = rng.uniform(-0.4, 0.4, (n, 3)).astype(np.float32) + np.float32([0, 0.5, 0])
    d = sr.norm(tgt - o)
    r = load("Mesh", (8, 8), 1.0)
    cpu = oracle_lib.OracleSampler()
    cpu._bind(r.scene, r.frame, r.rt)
    got = sr.closest_hit(objs[:1], o, d, {0: tree})
    # the oracle has no free-ray probe: route the rays through its per-path entry point is not possible either, so the
    # comparison of free rays is against a brute-force walk of the SAME second restatement with the octree switched off —
    # holes must show up as rays that hit a triangle by brute force but not through the octree
    hits_bf = np.zeros(n, bool)
    for ti in range(len(tris)):
        ok, t = sr.tri_intersect(tris[ti], np.float32(objs[0][3][0][0]), sr.to_object(*objs[0][3][0], o), sr.to_object(*objs[0][3][0], d, is_dir=True))
        hits_bf |= ok
    assert (got["hit"] & ~hits_bf).sum() == 0
    holes = (hits_bf & ~got["hit"]).sum()
    assert 0 <= holes < 0.2 * hits_bf.sum()


@pytest.mark.parametrize("name,res", [("Minecraft", (160, 90)), ("Instance", (96, 54)), ("dof", (96, 54))])
def test_direct_light_of_point_and_dir_lights(name, res):
    """rt.bounce = 0, no lens jitter: one deterministic path per pixel.  Radiance of the oracle's literal reduce_light
    against the second restatement's direct-light formula (dir light: -norm(dir); point light: pos - hit; shadow rays
    without a distance limit; texture fetches through the box atlas)."""
    r, cpu, h, desc, objs, trees = _primary(name, res, bounce=0)
    o, d = h["orig"].astype(np.float32), h["dir"].astype(np.float32)
    got = sr.closest_hit(objs, o, d, trees)
    n = sr.hit_normal(objs, got, o, d)
    uv = sr.hit_uv(objs, got, o, d)
    rad, valid = sr.direct_light(desc, objs, got, o, d, n, uv, trees)
    cpu.set_mode(oracle_lib.LITERAL)
    cpu.execute(r.scene, r.frame, r.rt, 1)
    acc = cpu.accum()[0].reshape(-1, 3)
    m = valid & (h["obj"] == got["obj"]) & (h["inst"] == got["inst"])
    assert m.sum() > 0.15 * m.size
    err = np.abs(rad[m] - acc[m]).max(axis=1) / (1e-3 + np.abs(acc[m]).max(axis=1))
    assert (err <= 2e-3).mean() >= 0.995, (name, np.sort(err)[-8:])
    assert acc[m].max() > 0.05


def test_material_maps_on_every_box_face_and_sphere():
    """tex / rmap / mmap fetched through the box atlas (all six faces, rotated instance included) and the sphere's
    atan2 mapping, lit by a dir AND a point light: oracle vs second restatement on a synthetic scene."""
    rng = np.random.default_rng(11)

    def tex(w, h):
        return {"w": w, "h": h, "dat": rng.uniform(0.05, 1.0, (w * h, 3)).round(3).tolist()}
    objs = [
        {"type": "box", "sizes": [0.8, 0.6, 0.5], "pos": [-0.7, 1.2, 0.1], "dir": [0.2, 0.5, 1, 0.1],
         "mat": {"tex": tex(16, 12), "rmap": tex(8, 6), "mmap": tex(4, 3), "albedo": [0.9, 0.8, 0.7]}},
        {"type": "box", "sizes": [0.7, 0.7, 0.7], "inst": [[[0.6, 1.4, -0.2], [0, 0, -1, 0]], [[0.1, 2.6, 0.6], [0, 0.3, -1, 0.2]]],
         "mat": {"tex": tex(12, 9), "rough": 0.4}},
        {"type": "sphere", "r": 0.35, "pos": [0.0, 0.9, -0.45], "mat": {"tex": tex(32, 16), "mmap": tex(5, 5)}},
        {"type": "plane", "n": [0, 0, 1], "pos": [0, 0, -0.8], "mat": {"tex": tex(8, 8), "rough": 1}},
    ]
    desc = {"rt": {"bounce": 0}, "frame": {"res": [120, 80], "cam": {"pos": [0, -1.2, 0.3], "aprt": 0, "fov": 75}},
            "scene": {"renderer": objs, "sky": {"color": [0.3, 0.5, 0.7], "pwr": 0.4},
                      "light": [{"type": "dir", "dir": [0.2, 0.6, -1], "pwr": 0.6, "color": "#ffeecc"},
                                {"type": "point", "pos": [-1.5, -0.5, 1.5], "pwr": 0.4}]}}
    r = mrt.render_from_dict(desc)
    cpu = oracle_lib.OracleSampler()
    cpu._bind(r.scene, r.frame, r.rt)
    h = cpu.trace_primary().reshape(-1)
    objects = sr.load_objects(desc, SCENES)
    o, d = h["orig"].astype(np.float32), h["dir"].astype(np.float32)
    got = sr.closest_hit(objects, o, d)
    same = (got["hit"] == (h["obj"] >= 0)) & ((h["obj"] < 0) | ((got["obj"] == h["obj"]) & (got["inst"] == h["inst"])))
    assert same.mean() >= 0.999
    n = sr.hit_normal(objects, got, o, d)
    uv = sr.hit_uv(objects, got, o, d)
    m = same & (h["obj"] >= 0)
    du = np.abs(uv[m] - h["uv"][m]); du = np.minimum(du, 1 - du).max(axis=1)
    assert (du <= 2e-4).mean() >= 0.998
    assert len(np.unique(np.round(n[m & (h["obj"] == 0)], 3), axis=0)) >= 2  # more than one face of the rotated box is seen
    assert len(np.unique(np.round(n[m & (h["obj"] <= 1)], 3), axis=0)) >= 5  # and several orientations over all boxes
    rad, valid = sr.direct_light(desc, objects, got, o, d, n, uv)
    cpu.set_mode(oracle_lib.LITERAL)
    cpu.execute(r.scene, r.frame, r.rt, 1)
    acc = cpu.accum()[0].reshape(-1, 3)
    mm = m & valid
    err = np.abs(rad[mm] - acc[mm]).max(axis=1) / (1e-3 + np.abs(acc[mm]).max(axis=1))
    # texel-boundary pixels may fetch a neighbouring texel after a last-ulp difference in uv
    assert (err <= 2e-3).mean() >= 0.985, np.sort(err)[-10:]
