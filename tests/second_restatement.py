"""A SECOND, independent restatement of the parts of the reference that its published renders never reach —
TEST INFRASTRUCTURE, written from /root/reference/src alone (numpy, float32), NOT from oracle/mrt_oracle.cpp:

  * instance expansion of a renderer description              parser.rs:838-853
  * per-instance transform rotate_y(-dir) * lookat(-dir)       rt.rs:726-733, lin.rs:175-183, 197-209
  * Box / Sphere / Plane / Triangle intersect                  rt.rs:299-412
  * the depth-3 mesh octree: construction by vertex containment, recursive traversal with the
    ancestors' boxes, adjacent dedup, first-min / last-max     rt.rs:227-248, 261-270, 630-723, 740-772, parser.rs:805-824
  * closest_hit's first-minimum rule over (object, instance)   rt.rs:867-872
  * Box::normal incl. the missing `else`, Renderer::normal     rt.rs:414-466, 776-793
  * Box / Sphere / Plane uv, Texture::get_color                rt.rs:468-541, 618-628
  * direct light of point AND dir lights, unbounded shadows    rt.rs:973-987, 1027-1045

The C++ oracle is pinned by doc/out0-4.png only for spheres, planes, identity boxes, point lights and one plane
texture; everything above is "parity unpinned by the reference".  Two restatements written separately from the same
source must agree — a shared misreading is the only way both can be wrong the same way.
Everything is vectorised over rays; arithmetic is float32 in the reference's operation order where it matters.
"""
import json
import os

import numpy as np

F = np.float32
E = F(0.0001)  # rt.rs:7


def v3(x):
    return np.asarray(x, dtype=F)


def dot(a, b):  # Vec3f * Vec3f, lin.rs:259-264: x*x + y*y + z*z in that order
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def cross(a, b):  # lin.rs:52-58
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def norm(a):  # lin.rs:60-66: self * mag().recip()
    mag = np.sqrt(a[..., 0] ** 2 + a[..., 1] ** 2 + a[..., 2] ** 2, dtype=F)
    with np.errstate(divide="ignore", invalid="ignore"):
        return a * (F(1) / mag)[..., None]


def rotate_y(dir4):  # lin.rs:175-183 — uses only w
    w = F(dir4[0])
    cw = np.sqrt(F(1) - w * w, dtype=F)
    return np.array([[cw, 0, w], [0, 1, 0], [-w, 0, cw]], dtype=F)


def lookat(dir4):  # lin.rs:197-209 with up = (0,0,1); the 3x3 part the Mat4f * Vec3f product uses (lin.rs:355-365)
    fwd = norm(v3(dir4[1:4]))
    right = norm(cross(fwd, v3([0, 0, 1])))
    up = cross(right, fwd)
    return np.array([[right[0], -right[1], right[2]], [-fwd[0], fwd[1], -fwd[2]], [up[0], -up[1], up[2]]], dtype=F)


def matvec(m, v):  # lin.rs:344-353: row . v, summed left to right
    return np.stack([m[i, 0] * v[..., 0] + m[i, 1] * v[..., 1] + m[i, 2] * v[..., 2] for i in range(3)], axis=-1)


def to_object(pos, dir4, p, is_dir=False):
    """rt.rs:726-733 / 779-782: rot_y * (look * x) with the NEGATED instance dir; points go around inst.pos."""
    nd = [-F(c) for c in dir4]
    ry, lk = rotate_y(nd), lookat(nd)
    if is_dir:
        return matvec(ry, matvec(lk, p))
    return pos + matvec(ry, matvec(lk, p - pos))


# ----------------------------------------------------------------------------- description -> objects
def expand_instances(obj):  # parser.rs:838-853
    backward = [-0.0, -0.0, -1.0, -0.0]  # lin.rs:143-145
    if obj.get("inst") is not None:
        inst = [(list(p), list(d)) for p, d in obj["inst"]]
        if obj.get("pos") is not None or obj.get("dir") is not None:
            inst.insert(0, (list(obj.get("pos") or [0.0, 0.0, 0.0]), list(obj.get("dir") or backward)))
        return inst
    return [(list(obj.get("pos") or [0.0, 0.0, 0.0]), list(obj.get("dir") or backward))]


def hex_color(c):  # parser.rs:713-733
    if isinstance(c, str):
        return [int(c[1 + 2 * k:3 + 2 * k], 16) / 255.0 for k in range(3)]
    return list(c)


def load_texture(v, base):
    if v is None:
        return None
    if isinstance(v, dict):
        return (int(v["w"]), int(v["h"]), None if v.get("dat") is None else v3(v["dat"]).reshape(-1, 3))
    if "." in v:  # a file: RGB8 / 255 (parser.rs:660-672)
        from PIL import Image
        im = np.asarray(Image.open(os.path.join(base, v)).convert("RGB"))
        return (im.shape[1], im.shape[0], (im.reshape(-1, 3).astype(F) / F(255.0)))
    import base64, gzip
    return load_texture(json.loads(gzip.decompress(base64.b64decode(v))), base)


def load_obj_mesh(path):  # parser.rs:601-618: position indices of the polygons (triangles) of the first object/group
    vs, tris = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            vs.append([float(x) for x in t[1:4]])
        elif t[0] == "f":
            idx = [int(x.split("/")[0]) for x in t[1:4]]
            tris.append([vs[i - 1] if i > 0 else vs[i] for i in idx])
    return v3(tris)


def load_objects(desc, base):
    """[(kind, params, material dict, instances)] in declaration order."""
    out = []
    for o in desc["scene"].get("renderer") or []:
        m = dict(o.get("mat") or {})
        mat = {"albedo": v3(hex_color(m.get("albedo", [1, 1, 1]))), "rough": F(m.get("rough", 0)), "metal": F(m.get("metal", 0)),
               "glass": F(m.get("glass", 0)), "opacity": F(m.get("opacity", 1)), "emit": F(m.get("emit", 0))}
        for k in ("tex", "rmap", "mmap", "gmap", "omap", "emap"):
            mat[k] = load_texture(m.get(k), base)
        kind = o["type"]
        if kind == "sphere":
            par = F(o["r"])
        elif kind == "plane":
            par = v3(o["n"])
        elif kind == "box":
            par = v3(o["sizes"])
        elif kind == "mesh":
            mesh = o["mesh"]
            par = load_obj_mesh(os.path.join(base, mesh)) if isinstance(mesh, str) else v3(mesh)
        else:
            raise ValueError(kind)
        out.append((kind, par, mat, [(v3(p), [F(c) for c in d]) for p, d in expand_instances(o)]))
    return out


# ----------------------------------------------------------------------------- intersect, rt.rs:299-412
def box_intersect(size, pos, o, d):
    with np.errstate(divide="ignore"):
        m = F(1) / d
    m = np.where(np.isinf(m), F(1) / E, m)  # rt.rs:303-316 (the sign is lost)
    n = (o - pos) * m
    k = (F(0.5) * size) * np.abs(m)
    a, b = -n - k, -n + k
    t0 = np.maximum(np.maximum(a[..., 0], a[..., 1]), a[..., 2])
    t1 = np.minimum(np.minimum(b[..., 0], b[..., 1]), b[..., 2])
    return ~((t0 > t1) | (t1 < 0)), t0, t1


def sphere_intersect(r, pos, o, d):  # rt.rs:335-359
    oc = o - pos
    a = dot(d, d)
    b = F(2) * dot(oc, d)
    c = dot(oc, oc) - r * r
    disc = b * b - F(4) * a * c
    with np.errstate(invalid="ignore"):
        sq = np.sqrt(disc, dtype=F)
        t0 = (-b - sq) / (F(2) * a)
        t1 = (-b + sq) / (F(2) * a)
    return ~((disc < 0) | (t0 < 0) | np.isnan(t0)), t0, t1


def plane_intersect(n, pos, o, d):  # rt.rs:400-412
    nh = norm(n)
    dd = -dot(nh, pos)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = -(dot(o, nh) + dd) / dot(d, nh)
    return ~(t <= 0), t


def tri_intersect(tri, pos, o, d):  # rt.rs:361-398
    e0, e1 = tri[1] - tri[0], tri[2] - tri[0]
    p = cross(d, e1)
    det = dot(e0, p)
    ok = ~((det < E) & (det > -E))
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = F(1) / det
        t = o - (tri[0] + pos)
        u = dot(t, p) * inv
        ok &= ~((u < 0) | (u > 1))
        q = cross(t, e0)
        v = dot(d, q) * inv
        ok &= ~((v < 0) | ((u + v) > 1))
        tt = dot(e1, q) * inv
        ok &= ~(tt < 0)
    return ok, tt


# ----------------------------------------------------------------------------- mesh octree, rt.rs:630-723
GEN_POS = v3([[1, 1, 1], [-1, 1, 1], [-1, -1, 1], [1, -1, 1], [1, 1, -1], [-1, 1, -1], [-1, -1, -1], [1, -1, -1]])


def mesh_aabb(tris):  # rt.rs:261-270
    a = np.abs(tris.reshape(-1, 3))
    return F(2) * a.max(axis=0)


def build_octree(tris, deep=3):
    """BVH::construct as nested dicts {aabb, rel, content | childs}; empty children are dropped (rt.rs:662-666)."""
    def construct(aabb, rel, d):
        node = {"aabb": aabb, "rel": rel, "content": None, "childs": None}
        if d >= deep:
            hi, lo = rel + F(0.5) * aabb, rel - F(0.5) * aabb
            inside = np.all(~(tris > hi), axis=2) & np.all(~(tris < lo), axis=2)  # per vertex, rt.rs:231-241
            idx = np.nonzero(inside.any(axis=1))[0]
            if len(idx):
                node["content"] = idx
            return node
        kids = [construct(F(0.5) * aabb, rel + aabb * (g * F(0.25)), d + 1) for g in GEN_POS]
        kids = [k for k in kids if k["content"] is not None or k["childs"] is not None]
        if kids:
            node["childs"] = kids
        return node
    return construct(mesh_aabb(tris), v3([0, 0, 0]), 0)


def mesh_intersect(tris, tree, pos, o, d):
    """rt.rs:707-723 + 740-772, vectorised: the candidate SEQUENCE of a ray is the concatenation, in depth-first
    order, of the lists of the leaves whose whole chain of boxes the ray pierces; adjacent duplicates are dropped
    (`dedup`), entry = first minimum, exit = last maximum."""
    n = o.shape[0]
    best0, best1 = np.full(n, np.inf, F), np.full(n, -np.inf, F)
    i0, i1 = np.full(n, -1, np.int64), np.full(n, -1, np.int64)
    last = np.full(n, -1, np.int64)  # previous candidate of each ray (for dedup)

    def walk(node, active):
        hit, _, _ = box_intersect(node["aabb"], pos + node["rel"], o, d)
        active = active & hit
        if not active.any():
            return
        if node["content"] is not None:
            for ti in node["content"]:
                todo = active & (last != ti)
                last[active] = ti
                if not todo.any():
                    continue
                ok, t = tri_intersect(tris[ti], pos, o, d)
                ok &= todo
                lt = ok & ((i0 < 0) | (t < best0))     # min_by keeps the FIRST minimum
                best0[lt], i0[lt] = t[lt], ti
                ge = ok & ((i1 < 0) | (t >= best1))    # max_by keeps the LAST maximum
                best1[ge], i1[ge] = t[ge], ti
            return
        for k in node["childs"]:  # a node without content and without children panics in the reference (unwrap)
            walk(k, active)

    walk(tree, np.ones(n, bool))
    return i0 >= 0, best0, best1, i0, i1


# ----------------------------------------------------------------------------- closest hit, normal, uv
def closest_hit(objects, o, d, trees=None):
    """rt.rs:867-872: brute force in (object, instance) order, first minimum of t0."""
    n = o.shape[0]
    best = np.full(n, np.inf, F)
    res = {"hit": np.zeros(n, bool), "t0": np.zeros(n, F), "t1": np.zeros(n, F), "obj": np.full(n, -1), "inst": np.full(n, -1),
           "tri0": np.full(n, -1), "tri1": np.full(n, -1)}
    for oi, (kind, par, _mat, insts) in enumerate(objects):
        for ii, (pos, dir4) in enumerate(insts):
            ol, dl = to_object(pos, dir4, o), to_object(pos, dir4, d, is_dir=True)
            tri0 = tri1 = None
            if kind == "sphere":
                ok, t0, t1 = sphere_intersect(par, pos, ol, dl)
            elif kind == "plane":
                ok, t0 = plane_intersect(par, pos, ol, dl)
                t1 = t0
            elif kind == "box":
                ok, t0, t1 = box_intersect(par, pos, ol, dl)
            else:
                ok, t0, t1, tri0, tri1 = mesh_intersect(par, trees[oi], pos, ol, dl)
            better = ok & (~res["hit"] | (t0 < best))   # strict: the first of equal minima stays
            best[better] = t0[better]
            res["hit"] |= better
            for k, v in (("t0", t0), ("t1", t1)):
                res[k][better] = v[better]
            res["obj"][better], res["inst"][better] = oi, ii
            res["tri0"][better] = tri0[better] if tri0 is not None else -1
            res["tri1"][better] = tri1[better] if tri1 is not None else -1
    return res


def box_normal(size, hit, pos):  # rt.rs:414-445 — note the missing `else` before the z test
    p = (hit - pos) * ((F(1) / size) * F(2))
    lo_p, hi_p, lo_n, hi_n = F(1) - E, F(1) + E, F(-1) - E, F(-1) + E
    inp = lambda c: (c >= lo_p) & (c < hi_p)
    inn = lambda c: (c >= lo_n) & (c < hi_n)
    n = np.zeros_like(p)
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    c1 = inp(x); c2 = ~c1 & inn(x); c3 = ~c1 & ~c2 & inp(y); c4 = ~c1 & ~c2 & ~c3 & inn(y)
    n[c1] = [1, 0, 0]; n[c2] = [-1, 0, 0]; n[c3] = [0, 1, 0]; n[c4] = [0, -1, 0]
    z1 = inp(z); z2 = ~z1 & inn(z)
    n[z1] = [0, 0, 1]; n[z2] = [0, 0, -1]
    return n


def hit_normal(objects, res, o, d, which="t0"):
    """Renderer::normal, rt.rs:776-793: kind normal of the object-space hit point, pushed through the FORWARD transform."""
    n = np.zeros_like(o)
    hp = o + d * res[which][:, None]
    for oi, (kind, par, _mat, insts) in enumerate(objects):
        for ii, (pos, dir4) in enumerate(insts):
            m = res["hit"] & (res["obj"] == oi) & (res["inst"] == ii)
            if not m.any():
                continue
            nh = to_object(pos, dir4, hp[m])
            if kind == "sphere":
                kn = nh - pos
            elif kind == "plane":
                kn = np.broadcast_to(par, nh.shape).copy()
            elif kind == "box":
                kn = box_normal(par, nh, pos)
            else:
                ti = res["tri0" if which == "t0" else "tri1"][m]
                kn = cross(par[ti, 1] - par[ti, 0], par[ti, 2] - par[ti, 0])
            n[m] = norm(to_object(pos, dir4, kn, is_dir=True))
    return n


def box_uv(size, hit, pos):  # rt.rs:468-516: x and y faces return first
    p = (hit - pos) * ((F(1) / size) * F(2))
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    inp = lambda c: (c >= F(1) - E) & (c < F(1) + E)
    inn = lambda c: (c >= F(-1) - E) & (c < F(-1) + E)
    h, q, th = F(0.5), F(4.0), F(3.0)
    cases = [
        (inp(x), (h + h * y) / q + F(2.0) / q, (h - h * z) / th + F(1.0) / th),
        (inn(x), (h - h * y) / q, (h - h * z) / th + F(1.0) / th),
        (inp(y), (h - h * x) / q + F(3.0) / q, (h - h * z) / th + F(1.0) / th),
        (inn(y), (h + h * x) / q + F(1.0) / q, (h - h * z) / th + F(1.0) / th),
        (inp(z), (h + h * x) / q + F(1.0) / q, (h - h * y) / th),
        (inn(z), (h + h * x) / q + F(1.0) / q, (h + h * y) / th + F(2.0) / th),
    ]
    u, v, done = np.zeros_like(x), np.zeros_like(x), np.zeros(x.shape, bool)
    for c, uu, vv in cases:
        c = c & ~done
        u[c], v[c] = uu[c], vv[c]
        done |= c
    return np.stack([u, v], axis=-1)


def sphere_uv(hit, pos):  # rt.rs:518-526
    v = norm(hit - pos)
    return np.stack([F(0.5) + F(0.5) * np.arctan2(v[..., 0], -v[..., 1]).astype(F) / F(np.pi), F(0.5) - F(0.5) * v[..., 2]], axis=-1)


def plane_uv(hit):  # rt.rs:528-542: fract of x and y whatever the plane's normal
    def fr(c):
        c = c + F(0.5)
        f = c - np.trunc(c)
        return np.where(f < 0, F(1) + f, f)
    return np.stack([fr(hit[..., 0]), fr(hit[..., 1])], axis=-1)


def hit_uv(objects, res, o, d):
    uv = np.zeros((o.shape[0], 2), F)
    hp = o + d * res["t0"][:, None]
    for oi, (kind, par, _mat, insts) in enumerate(objects):
        for ii, (pos, dir4) in enumerate(insts):
            m = res["hit"] & (res["obj"] == oi) & (res["inst"] == ii)
            if not m.any() or kind == "mesh":
                continue
            nh = to_object(pos, dir4, hp[m])
            uv[m] = sphere_uv(nh, pos) if kind == "sphere" else plane_uv(nh) if kind == "plane" else box_uv(par, nh, pos)
    return uv


def tex_fetch(tex, uv):  # rt.rs:618-628: truncating casts, linear index x + y*w
    w, h, dat = tex
    if dat is None:
        return np.zeros((uv.shape[0], 3), F)
    x = np.maximum(uv[:, 0] * F(w), 0).astype(np.int64)
    y = np.maximum(uv[:, 1] * F(h), 0).astype(np.int64)
    return dat[np.minimum(x + y * w, w * h - 1)]  # the reference panics past the end; the product clamps


# ----------------------------------------------------------------------------- direct light at the first hit
def direct_light(desc, objects, res, o, d, normals, uv, trees=None):
    """What one path returns with rt.bounce = 0 and no lens jitter, for hits on OPAQUE, non-emissive materials
    (everything is deterministic there): fold seed sky.color * sky.pwr (rt.rs:964), visibility from the hit point with
    no distance limit (rt.rs:1027-1045), then (0.5 col + color (.) col + l_col) * pwr with pwr = 1 (rt.rs:973-992).
    Returns (radiance, mask of the pixels this holds for)."""
    sky = desc["scene"].get("sky") or {}
    tail = v3(hex_color(sky.get("color", [0, 0, 0]))) * F(sky.get("pwr", 0.5))
    n = o.shape[0]
    hp = o + d * res["t0"][:, None]
    color = np.zeros((n, 3), F)
    rough, metal = np.zeros(n, F), np.zeros(n, F)
    valid = res["hit"].copy()
    for oi, (_kind, _par, mat, _insts) in enumerate(objects):
        m = res["hit"] & (res["obj"] == oi)
        if not m.any():
            continue
        if mat["omap"] is not None or mat["emap"] is not None or mat["opacity"] != 1 or mat["emit"] != 0:
            valid[m] = False
            continue
        c = np.broadcast_to(mat["albedo"], (int(m.sum()), 3)).copy()
        if mat["tex"] is not None:
            c = c * tex_fetch(mat["tex"], uv[m])
        color[m] = c
        rough[m] = tex_fetch(mat["rmap"], uv[m])[:, 0] if mat["rmap"] is not None else mat["rough"]
        metal[m] = tex_fetch(mat["mmap"], uv[m])[:, 0] if mat["mmap"] is not None else mat["metal"]
    l_col = np.zeros((n, 3), F)
    for light in desc["scene"].get("light") or []:
        if light["type"] == "point":
            l = v3(light["pos"]) - hp
        else:
            l = np.broadcast_to(-norm(v3(light["dir"])), hp.shape)
        ln = norm(l)
        sh = closest_hit(objects, hp + ln * E, ln, trees)  # Ray::cast_default offsets the origin by dir * E (rt.rs:555-557)
        vis = ~sh["hit"]
        diff = np.maximum(dot(ln, normals), F(0))
        refl = ln - normals * (F(2) * dot(ln, normals))[:, None]  # lin.rs:68-70
        spec = np.maximum(dot(d, refl), F(0)) ** 32 * (F(1) - rough)
        o_col = color * (F(1) - metal)[:, None]
        term = ((o_col * diff[:, None]) * v3(hex_color(light.get("color", [1, 1, 1]))) + spec[:, None]) * F(light.get("pwr", 0.5))
        l_col += np.where(vis[:, None], term, F(0)).astype(F)
    rad = (F(0.5) * tail + color * tail + l_col) * F(1)
    return rad, valid
