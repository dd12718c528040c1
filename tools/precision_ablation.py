"""Precision ablation: how much of the deterministic mismatch budget against the oracle is owed
to the kernel's approximate units (MUFU.RCP / MUFU.RSQ / MUFU.SIN, -prec-div=false -prec-sqrt=false) and how much to the
reference's own ill-conditioned arithmetic?  Runs the per-ray probe (mrt_trace_primary, lens jitter off) and a
shared-random-number render of every example scene through the library named by MRT_LIB and compares with the oracle.

    python tools/precision_ablation.py OUT.json            # the in-tree build
    MRT_LIB=scratch/libmrt_precise.so python tools/precision_ablation.py OUT.json   # tools/build_variant.sh precise ...
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import micro_raytracer_b200 as mrt
import oracle_lib
from util import load

SCENES = [("Default", (320, 180), 1.0), ("CornellBox2", (160, 160), 2.0), ("CornellBox", (320, 180), 1.0), ("dof", (320, 180), 1.0),
          ("Minecraft", (320, 180), 1.0), ("Instance", (192, 108), 1.0), ("Mesh", (320, 180), 1.0)]


def q(a, p):
    return float(np.quantile(a, p)) if a.size else 0.0


def main():
    out = {"lib": os.environ.get("MRT_LIB", "in-tree libmrt.so"), "scenes": {}}
    for jit in (0, 2):
        for name, res, ssaa in SCENES:
            r = load(name, res, ssaa)
            r.frame.cam.aprt = 0.0
            g = mrt.Sampler(device=0)
            g.set_option(2, jit)
            c = oracle_lib.OracleSampler()
            for s in (g, c):
                s.execute(r.scene, r.frame, r.rt, 2)
            hg, hc = g.trace_primary(), c.trace_primary()
            same = (hg["obj"] == hc["obj"]) & (hg["inst"] == hc["inst"]) & (hg["tri0"] == hc["tri0"])
            m = same & (hc["obj"] >= 0)
            dt = np.abs(hg["t0"][m] - hc["t0"][m]) / np.maximum(1.0, np.abs(hc["t0"][m]))
            dn = np.abs(hg["n0"][m] - hc["n0"][m]).max(axis=1)
            du = np.abs(hg["uv"][m] - hc["uv"][m]); du = np.minimum(du, 1.0 - du).max(axis=1)
            ag, ac = g.accum()[0], c.accum()[0]
            rel = np.abs(ag - ac).max(axis=2) / (1e-3 + np.abs(ac).max(axis=2))
            out["scenes"][f"{name}/{'jit' if jit else 'generic'}"] = {
                "rays": int(same.size), "ids_equal": float(same.mean()), "ids_differ": int((~same).sum()),
                "dt_rel_max": float(dt.max()) if dt.size else 0.0, "dt_rel_p9999": q(dt, 0.9999), "dt_le_1e-5": float((dt <= 1e-5).mean()) if dt.size else 1.0,
                "dn_max": float(dn.max()) if dn.size else 0.0, "dn_p9999": q(dn, 0.9999), "dn_le_1e-5": float((dn <= 1e-5).mean()) if dn.size else 1.0,
                "duv_p9999": q(du, 0.9999), "duv_le_1e-5": float((du <= 1e-5).mean()) if du.size else 1.0,
                "paths_within_1e-3": float((rel <= 1e-3).mean()), "paths_within_1e-4": float((rel <= 1e-4).mean()),
            }
    json.dump(out, open(sys.argv[1], "w"), indent=1)
    for k, v in out["scenes"].items():
        print(k, {a: (round(b, 7) if isinstance(b, float) else b) for a, b in v.items()})


if __name__ == "__main__":
    main()
