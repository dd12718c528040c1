"""Seeded random scenes in the reference's own JSON vocabulary (parser.rs), for parity fuzzing:
every primitive kind, rotated instances (yaw + roll), instance lists, every material field and
map, both light kinds, sky, DOF, fractional ssaa.  Shared by the CPU and GPU fuzz tests."""
import numpy as np

import micro_raytracer_b200 as mrt


def _tex(rng, w, h, scalar=False):
    dat = rng.uniform(0.05, 1.0, size=(w * h, 3)).round(3)
    if scalar:
        dat[:, 1:] = dat[:, :1]
    return {"w": w, "h": h, "dat": dat.tolist()}


def _dir(rng, allow_roll=True):
    a = rng.uniform(0, 2 * np.pi)
    z = rng.uniform(-0.4, 0.4)
    w = float(rng.uniform(-0.5, 0.5)) if allow_roll and rng.random() < 0.5 else 0.0
    return [w, float(np.cos(a)), float(np.sin(a)), float(z)]


def _mat(rng, textured_ok=True):
    m = {"albedo": rng.uniform(0.2, 1.0, 3).round(3).tolist()}
    r = rng.random()
    if r < 0.2:
        m["metal"] = 1.0
        m["rough"] = float(rng.choice([0.0, 0.1, 0.5]))
    elif r < 0.35:
        m["opacity"] = float(rng.choice([0.0, 0.3, 0.7]))
        m["glass"] = float(rng.uniform(0.0, 0.5))
    elif r < 0.5:
        m["emit"] = float(rng.choice([1.0, 0.5, 0.25]))
    else:
        m["rough"] = float(rng.uniform(0.0, 1.0))
    if textured_ok and rng.random() < 0.45:
        for key, scalar in (("tex", False), ("rmap", True), ("mmap", True), ("gmap", True), ("omap", True), ("emap", True)):
            if rng.random() < 0.3:
                m[key] = _tex(rng, int(rng.integers(2, 9)), int(rng.integers(2, 9)), scalar)
    return m


def _mesh(rng, n):
    c = rng.uniform(-0.35, 0.35, size=(n, 1, 3))
    return (c + rng.uniform(-0.15, 0.15, size=(n, 3, 3))).round(4).tolist()


def random_scene(seed, res=(56, 40), many=False):
    """many=True: a few objects with 40..160 instances each (> 128 primitives: the scene-level BVH)."""
    rng = np.random.default_rng(seed + (100000 if many else 0))
    objs = []
    for oi in range(int(rng.integers(2, 9)) if not many else int(rng.integers(2, 5))):
        kind = rng.choice(["sphere", "box", "plane", "mesh"], p=[0.35, 0.4, 0.15, 0.1])
        if many and oi == 0:
            kind = rng.choice(["sphere", "box"])
        o = {"type": str(kind)}
        if kind == "sphere":
            o["r"] = float(rng.uniform(0.1, 0.5))
        elif kind == "box":
            o["sizes"] = rng.uniform(0.15, 0.9, 3).round(3).tolist()
        elif kind == "plane":
            n = rng.normal(size=3)
            n[2] = abs(n[2]) + 0.5
            o["n"] = n.round(3).tolist()
        else:
            o["mesh"] = _mesh(rng, int(rng.integers(4, 40)))
        o["mat"] = _mat(rng, textured_ok=kind != "mesh")
        pos = [float(rng.uniform(-1.2, 1.2)), float(rng.uniform(0.6, 3.0)), float(rng.uniform(-0.8, 0.8))]
        if kind == "plane":
            pos = [0.0, 0.0, float(rng.uniform(-1.2, -0.6))]
        if many and kind != "plane":
            if kind == "sphere":
                o["r"] *= 0.4
            elif kind == "box":
                o["sizes"] = [v * 0.4 for v in o["sizes"]]
            n_inst = (int(rng.integers(140, 201)) if oi == 0 else int(rng.integers(40, 161))) if kind != "mesh" else int(rng.integers(2, 6))
            o["inst"] = [[[float(rng.uniform(-2.5, 2.5)), float(rng.uniform(0.5, 6.0)), float(rng.uniform(-1.5, 1.5))],
                          _dir(rng) if rng.random() < 0.4 else [0, 0, -1, 0]] for _ in range(n_inst)]
        elif rng.random() < 0.25 and kind != "plane":
            o["inst"] = [[[float(pos[0] + rng.uniform(-1, 1)), float(pos[1] + rng.uniform(0, 1)), float(pos[2] + rng.uniform(-.5, .5))], _dir(rng)]
                         for _ in range(int(rng.integers(1, 4)))]
            if rng.random() < 0.5:
                o["pos"] = pos
        else:
            o["pos"] = pos
            if rng.random() < 0.5:
                o["dir"] = _dir(rng)
        objs.append(o)
    lights = []
    for _ in range(int(rng.integers(0, 3))):
        if rng.random() < 0.5:
            lights.append({"type": "point", "pos": rng.uniform(-2, 2, 3).round(3).tolist(), "pwr": float(rng.uniform(0.3, 1.5)),
                           "color": rng.uniform(0.3, 1, 3).round(3).tolist()})
        else:
            d = rng.normal(size=3)
            d[2] = -abs(d[2]) - 0.2
            lights.append({"type": "dir", "dir": d.round(3).tolist(), "pwr": float(rng.uniform(0.3, 1.0))})
    cam = {"pos": [float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-1.5, -0.5)), float(rng.uniform(-0.2, 0.4))],
           "fov": float(rng.uniform(40, 90)), "aprt": float(rng.choice([0.0, 0.001, 0.02])), "foc": float(rng.choice([100.0, 2.0]))}
    if rng.random() < 0.4:
        cam["dir"] = [float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.3, 0.3)), 1.0, float(rng.uniform(-0.2, 0.2))]
    d = {"rt": {"bounce": int(rng.integers(0, 7)), "sample": 2, "loss": float(rng.choice([0.0, 0.15, 0.6, 1.5]))},
         "frame": {"res": list(res), "ssaa": float(rng.choice([1.0, 1.5, 2.0])), "cam": cam},
         "scene": {"renderer": objs, "light": lights or None,
                   "sky": {"color": rng.uniform(0, 1, 3).round(3).tolist(), "pwr": float(rng.uniform(0.1, 1.0))}}}
    return mrt.render_from_dict(d)
