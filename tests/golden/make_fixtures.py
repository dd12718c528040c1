#!/usr/bin/env python
"""Regenerates tests/golden/ from the read-only reference checkout (run in the build container).

  scenes/<name>.json   the reference's example/<name>.json with every inline base64+gzip
                       asset resolved into a side file in the reference's own file-asset
                       forms (textures: RGB8 PNG, meshes: .obj; parser.rs:602-618,660-672),
                       checked to pack bit-identically to the original description
  ref_renders/outN.png the reference's own renders doc/out0..out4.png (the only golden
                       outputs the reference ships)

/root/reference does not exist on the GPU box; tests only read the files written here.
"""
import ctypes as C
import json
import os
import shutil
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from micro_raytracer_b200 import scene as S  # noqa: E402

REF = "/root/reference"


def same_pack(a: S.PackedScene, b: S.PackedScene) -> bool:
    def raw(x):
        return bytes(memoryview(x))
    ok = raw(a.objects) == raw(b.objects) and raw(a.instances) == raw(b.instances)
    ok &= raw(a.textures) == raw(b.textures) and raw(a.meshes) == raw(b.meshes) and raw(a.lights) == raw(b.lights)
    ok &= a.texels.tobytes() == b.texels.tobytes() and a.triangles.tobytes() == b.triangles.tobytes()
    return ok


def main():
    out_sc = os.path.join(HERE, "scenes")
    out_rr = os.path.join(HERE, "ref_renders")
    os.makedirs(out_sc, exist_ok=True)
    os.makedirs(out_rr, exist_ok=True)
    for name in sorted(os.listdir(os.path.join(REF, "example"))):
        stem = name[:-5]
        d = json.load(open(os.path.join(REF, "example", name)))
        n_asset = 0
        for k, o in enumerate(d.get("scene", {}).get("renderer") or []):
            for key in ("tex", "rmap", "mmap", "gmap", "omap", "emap"):
                v = (o.get("mat") or {}).get(key)
                if isinstance(v, str) and "." not in v:
                    t = S._texture(v, None)
                    u8 = np.rint(t.dat * 255.0).astype(np.uint8)
                    assert np.array_equal((u8.astype(np.float32) / np.float32(255.0)), t.dat), "texture is not u8/255"
                    fn = f"{stem}_{k}_{key}.png"
                    Image.fromarray(u8.reshape(t.h, t.w, 3), "RGB").save(os.path.join(out_sc, fn))
                    o["mat"][key] = fn
                    n_asset += 1
            v = o.get("mesh")
            if isinstance(v, str) and "." not in v:
                m = S._mesh(v, None)
                fn = f"{stem}_{k}.obj"
                with open(os.path.join(out_sc, fn), "w") as fh:
                    fh.write("o mesh\n")
                    for tri in m:
                        for p in tri:
                            fh.write("v %s %s %s\n" % tuple(repr(float(np.float32(c))) for c in p))
                    for i in range(len(m)):
                        fh.write("f %d %d %d\n" % (3 * i + 1, 3 * i + 2, 3 * i + 3))
                o["mesh"] = fn
                n_asset += 1
        dst = os.path.join(out_sc, name)
        with open(dst, "w") as fh:
            json.dump(d, fh, indent=1)
        a = S.load_render(os.path.join(REF, "example", name))
        b = S.load_render(dst)
        assert a.rt == b.rt and a.frame == b.frame, name
        assert same_pack(S.pack_scene(a.scene), S.pack_scene(b.scene)), name
        print(f"{name}: {n_asset} assets resolved, packs identically")
    for k in range(5):
        shutil.copyfile(os.path.join(REF, "doc", f"out{k}.png"), os.path.join(out_rr, f"out{k}.png"))
    print("ref_renders copied")


if __name__ == "__main__":
    main()
