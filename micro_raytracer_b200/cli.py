"""`raytrace` front-end: the reference's command line (src/cli.rs, src/bin/raytrace.rs) over the
CUDA path.  `python -m micro_raytracer_b200 [FILE.json] [flags]` accepts the reference's flags and
its `key: v v v` mini-grammar for --cam / --obj / --light / --sky (src/parser.rs:274-598),
merges them with the same precedence as CLI::parse_render (cli.rs:78-153), and drives
Sampler::{execute, img} exactly like CLI::raytrace (cli.rs:155-177).  SURVEY.md §8(f) "next #2".

Everything is turned into the JSON description first, so the flags and the files share one
path into scene.render_from_dict.  --http ADDRESS starts the endpoint of http.py.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from typing import Iterator, List, Optional

import numpy as np

from .scene import Render, SceneError, render_from_dict

OBJ_TYPES = ["sphere", "sph", "plane", "pln", "box", "tri", "triangle", "mesh"]  # cli.rs:125
LIGHT_TYPES = ["pt:", "point:", "dir:"]                                           # cli.rs:135
BACKWARD = [-0.0, -0.0, -1.0, -0.0]                                               # Vec4f::backward(), lin.rs:143


class CliError(ValueError):
    """≙ Err(String) of the reference's parsers."""


def _f32(it: Iterator[str]) -> float:  # parser.rs:276-280
    try:
        tok = next(it)
    except StopIteration:
        raise CliError("unexpected ends!")
    try:
        return float(np.float32(float(tok)))
    except ValueError:
        raise CliError("should be <f32>!")


def _vec(it, n):
    return [_f32(it) for _ in range(n)]


def _color(it: "_Peek"):  # parser.rs:312-323: '#rrggbb' or three floats
    tok = it.peek()
    if tok is None:
        raise CliError("unexpected ends!")
    if tok.startswith("#"):
        next(it)
        return tok
    return _vec(it, 3)


class _Peek:
    def __init__(self, toks: List[str]):
        self.t, self.i = toks, 0

    def __iter__(self):
        return self

    def __next__(self):
        if self.i >= len(self.t):
            raise StopIteration
        self.i += 1
        return self.t[self.i - 1]

    def peek(self) -> Optional[str]:
        return self.t[self.i] if self.i < len(self.t) else None


def split_args(args: List[str], pat: List[str]) -> List[List[str]]:
    """ParseFromArgs::parse_args, parser.rs:584-598: the argument list is REVERSED, split after
    every type token, and each piece reversed back — so the objects come out in reverse order of
    the command line (why CornellBox2.json lists the README command's objects backwards)."""
    out, cur = [], []
    for tok in reversed(args):
        cur.append(tok)
        if tok in pat:
            out.append(list(reversed(cur)))
            cur = []
    if cur:
        out.append(list(reversed(cur)))
    return out


def camera_from_args(args: List[str]) -> dict:  # parser.rs:330-349: a fresh default camera + the given keys
    cam, it = {}, _Peek(args)
    for p in it:
        if p == "pos:":
            cam["pos"] = _vec(it, 3)
        elif p == "dir:":
            cam["dir"] = _vec(it, 4)
        elif p in ("fov:", "gamma:", "exp:", "aprt:", "foc:"):
            cam[p[:-1]] = _f32(it)
        else:
            raise CliError(f"`{p}` param for `cam` is unxpected!")
    return cam


def light_from_args(args: List[str]) -> dict:  # parser.rs:352-403
    t = args[0]
    if t in ("pt:", "point:"):
        light = {"type": "point", "pos": [0.0, 0.0, 0.0]}
    elif t == "dir:":
        light = {"type": "dir", "dir": [0.0, 1.0, 0.0]}
    else:
        raise CliError(f"`{t}` type is unxpected!")
    it = _Peek(args)
    for p in it:
        if light["type"] == "point" and p in ("pt:", "point:"):
            light["pos"] = _vec(it, 3)
        elif light["type"] == "dir" and p == "dir:":
            x, y, z = (np.float32(c) for c in _vec(it, 3))
            r = np.float32(1.0) / np.sqrt(x * x + y * y + z * z)  # Vec3f::norm = self * mag().recip(), lin.rs:60-66 (parser.rs:383)
            light["dir"] = [float(x * r), float(y * r), float(z * r)]
        elif p == "col:":
            light["color"] = _color(it)
        elif p == "pwr:":
            light["pwr"] = _f32(it)
        else:
            raise CliError(f"`{p}` param for `light` is unxpected!")
    return light


_TRI = [[0.5, 0.0, -0.25], [0.0, 0.0, 0.5], [-0.5, 0.0, -0.25]]  # parser.rs:417-421


def renderer_from_args(args: List[str]) -> dict:  # parser.rs:406-582
    t = args[0]
    if t in ("sph", "sphere"):
        obj = {"type": "sphere", "r": 0.5}
    elif t in ("pln", "plane"):
        obj = {"type": "plane", "n": [0.0, 0.0, 1.0]}
    elif t == "box":
        obj = {"type": "box", "sizes": [0.5, 0.5, 0.5]}
    elif t in ("tri", "triangle"):
        obj = {"type": "triangle", "vtx": [list(v) for v in _TRI]}
    elif t == "mesh":
        obj = {"type": "mesh", "mesh": [[list(v) for v in _TRI]]}
    else:
        raise CliError(f"`{t}` type is unxpected!")
    obj.update({"pos": [0.0, 0.0, 0.0], "dir": list(BACKWARD), "mat": {}})
    it = _Peek(args[1:])
    for p in it:
        kind = obj["type"]
        if kind == "sphere" and p == "r:":
            obj["r"] = _f32(it)
        elif kind == "plane" and p == "n:":
            obj["n"] = _vec(it, 3)
        elif kind == "box" and p == "size:":  # CLI spelling; the JSON key is `sizes`
            obj["sizes"] = _vec(it, 3)
        elif kind == "triangle" and p == "vtx:":
            obj["vtx"] = [_vec(it, 3) for _ in range(3)]
        elif kind == "mesh" and p == "mesh:":
            tris = [[_vec(it, 3) for _ in range(3)]]
            while True:  # parser.rs:493-503: triangles until the numbers run out (consumed tokens stay consumed)
                try:
                    tris.append([_vec(it, 3) for _ in range(3)])
                except CliError:
                    break
            obj["mesh"] = tris
        elif p == "name:":
            obj["name"] = next(it, None)
        elif p == "pos:":
            obj["pos"] = _vec(it, 3)
        elif p == "dir:":
            obj["dir"] = _vec(it, 4)
        elif p == "albedo:":
            obj["mat"]["albedo"] = _color(it)
        elif p in ("rough:", "metal:", "glass:", "opacity:", "emit:"):
            obj["mat"][p[:-1]] = _f32(it)
        elif p in ("tex:", "rmap:", "mmap:", "gmap:", "omap:", "emap:"):
            s = next(it, None)
            if s is None:
                raise CliError("unexpected ended!")
            obj["mat"][p[:-1]] = s  # a string with a '.' is a file, else inline base64 (parser.rs:521-527)
        else:
            raise CliError(f"`{p}` param for `{t}` is unxpected!")
    return obj


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="raytrace", description="Tiny raytracing microservice (B200 path).")
    ap.add_argument("full", nargs="?", metavar="FILE.json", help="Full render description json input filename")
    ap.add_argument("-v", "--verbose", action="store_true", help="Enable logging")
    ap.add_argument("--pretty", action="store_true", help="Print full render info in json with prettifier")
    ap.add_argument("-d", "--dry", action="store_true", help="Dry run (useful with verbose)")
    ap.add_argument("-o", "--output", metavar="FILE.EXT", help="Final image output filename")
    ap.add_argument("--http", metavar="address", help="Launch http server")
    ap.add_argument("--bounce", type=int, help="Max ray bounce")
    ap.add_argument("--sample", type=int, help="Max path-tracing samples")
    ap.add_argument("--loss", type=float, help="Ray bounce energy loss")
    ap.add_argument("-u", "--update", action="store_true", help="Save output on each sample")
    ap.add_argument("-w", "--worker", type=int, help="Parallel workers count (accepted, ignored: the CUDA grid replaces the pool)")
    ap.add_argument("--dim", type=int, help="Parallel jobs count on each dimension (accepted, ignored)")
    ap.add_argument("-s", "--scene", metavar="FILE.json", help="Scene description json input filename")
    ap.add_argument("-f", "--frame", metavar="FILE.json", help="Frame description json input filename")
    ap.add_argument("--res", nargs=2, type=int, metavar=("w", "h"), help="Frame output image resolution")
    ap.add_argument("--ssaa", type=float, help="Output image SSAAx antialiasing")
    ap.add_argument("--cam", nargs="+", help="Add camera to the scene: pos: dir: fov: gamma: exp: aprt: foc:")
    ap.add_argument("--obj", nargs="*", action="extend", help="Add renderer to the scene: type name: <param> pos: dir: albedo: rough: metal: glass: opacity: emit: tex: ...")
    ap.add_argument("--light", nargs="*", action="extend", help="Add light source to the scene: pt:|dir: <f32 f32 f32> pwr: col:")
    ap.add_argument("--sky", nargs="+", action="extend", help="Scene sky color: <f32 f32 f32> pwr")
    ap.add_argument("--device", type=int, default=0, help="CUDA device (extension)")
    ap.add_argument("--seed", type=int, default=0x5EED, help="RNG seed (extension: the reference is unseedable)")
    return ap


def _read_json(path):
    try:
        with open(path) as fh:
            return json.load(fh)
    except (OSError, json.JSONDecodeError) as e:
        raise CliError(str(e))


def parse_render_dict(ns: argparse.Namespace) -> dict:
    """CLI::parse_render, cli.rs:78-153, on the JSON dict (same precedence: full json -> --bounce/
    --sample/--loss -> --frame -> --res/--ssaa/--cam -> --scene -> --obj/--light -> --sky)."""
    d = _read_json(ns.full) if ns.full else {}
    rt = dict(d.get("rt") or {})
    if ns.bounce is not None:
        rt["bounce"] = ns.bounce
    if ns.sample is not None:
        rt["sample"] = ns.sample
    if ns.loss is not None:
        rt["loss"] = ns.loss
    frame = dict(d.get("frame") or {})
    if ns.frame:
        frame = _read_json(ns.frame)          # replaces the whole frame, cli.rs:101-104
    if ns.res:
        frame["res"] = list(ns.res)
    if ns.ssaa is not None:
        frame["ssaa"] = ns.ssaa
    if ns.cam:
        frame["cam"] = camera_from_args(ns.cam)  # replaces the whole camera, cli.rs:117-119
    scene = dict(d.get("scene") or {})
    if ns.scene:
        scene = _read_json(ns.scene)          # replaces the whole scene, cli.rs:122-125
    if ns.obj is not None:
        new = [renderer_from_args(a) for a in split_args(ns.obj, OBJ_TYPES)]
        scene["renderer"] = list(scene.get("renderer") or []) + new
    if ns.light is not None:
        new = [light_from_args(a) for a in split_args(ns.light, LIGHT_TYPES)]
        scene["light"] = list(scene.get("light") or []) + new
    if ns.sky:
        it = _Peek(ns.sky)
        scene["sky"] = {"color": _vec(it, 3), "pwr": _f32(it)}  # three floats then pwr; hex is not accepted here (cli.rs:146-150)
    return {"rt": rt, "frame": frame, "scene": scene}


def parse_render(argv: List[str]):
    ns = build_parser().parse_args(argv)
    d = parse_render_dict(ns)
    base = os.path.dirname(os.path.abspath(ns.full or ns.scene or "."))
    return ns, d, render_from_dict(d, base)


def _save(img: np.ndarray, path: str):
    from PIL import Image
    Image.fromarray(img).save(path)


def raytrace(ns: argparse.Namespace, render: Render, log=None) -> float:
    """CLI::raytrace, cli.rs:155-177: one Sampler, rt.sample passes, optional save per pass, final save."""
    from .sampler import Sampler
    sampler = Sampler(ns.worker or 24, ns.dim or 64, device=ns.device, seed=ns.seed)
    out = ns.output or "out.png"
    t0 = time.perf_counter()
    if ns.update:
        for n in range(render.rt.sample):
            dt = sampler.execute(render.scene, render.frame, render.rt)
            if log:
                log(f"cli:sample:{n}: {dt:.6f}s")
            _save(sampler.img(render.frame), out)
    elif render.rt.sample > 0:
        # without --update nothing observes the accumulator between passes: one call renders them all
        dt = sampler.execute(render.scene, render.frame, render.rt, render.rt.sample)
        if log:
            log(f"cli:sample:0..{render.rt.sample - 1}: {dt:.6f}s")
    _save(sampler.img(render.frame), out)
    return time.perf_counter() - t0


def main(argv: Optional[List[str]] = None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    try:
        ns, d, render = parse_render(argv)
        if ns.http:  # raytrace.rs:22-30: blocks forever
            from .http import serve
            serve(ns.http, ns.device)
            return 0
        log = (lambda m: print(m, flush=True)) if ns.verbose else None
        if ns.verbose:
            print(json.dumps(d, indent=2 if ns.pretty else None))
        if ns.dry:
            return 0
        dt = raytrace(ns, render, log)
        if log:
            log(f"cli:done: {dt:.3f}s")
        return 0
    except (CliError, SceneError) as e:
        print(f"cli: {e}", file=sys.stderr)  # raytrace.rs:55
        return 1
