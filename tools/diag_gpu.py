"""Diagnostics (not a test): GPU vs oracle summary numbers for every parity case."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import micro_raytracer_b200 as mrt
import oracle_lib
from util import load
from test_gpu_parity import CASES

gpu, cpu = mrt.Sampler(device=0), oracle_lib.OracleSampler()
for name, res, ssaa in CASES:
    r = load(name, res, ssaa)
    for s in (gpu, cpu):
        s._bind(r.scene, r.frame, r.rt)
    hg, hc = gpu.trace_primary(), cpu.trace_primary()
    same = (hg["obj"] == hc["obj"]) & (hg["inst"] == hc["inst"]) & (hg["tri0"] == hc["tri0"])
    m = same & (hc["obj"] >= 0)
    t = hc["t0"][m]
    dt = np.abs(hg["t0"][m] - t) / np.maximum(1, np.abs(t))
    dt1 = np.abs(hg["t1"][m] - hc["t1"][m]) / np.maximum(1, np.abs(hc["t1"][m]))
    fin = m & np.isfinite(hc["n0"]).all(axis=-1)
    dn = np.abs(hg["n0"][fin] - hc["n0"][fin]).max(axis=-1)
    fin1 = m & np.isfinite(hc["n1"]).all(axis=-1)
    dn1 = np.abs(hg["n1"][fin1] - hc["n1"][fin1]).max(axis=-1)
    duv = np.abs(hg["uv"][m] - hc["uv"][m]).max(axis=-1)
    print(f"{name}: ids differ {1-same.mean():.5f} hitfrac {m.mean():.3f} dir {np.abs(hg['dir']-hc['dir']).max():.2e} "
          f"dt max {dt.max():.2e} >1e-5: {(dt>1e-5).mean():.5f} dt1 max {dt1.max():.2e} >1e-4 {(dt1>1e-4).mean():.5f} | dn>1e-4 {(dn>1e-4).mean():.5f} dn1>1e-4 {(dn1>1e-4).mean():.5f} "
          f"nan n0 cpu {1-np.isfinite(hc['n0'][m]).all(axis=-1).mean():.5f} | duv>1e-4 {(duv>1e-4).mean():.5f}")
    if (dt > 1e-5).any():
        w = np.argmax(dt)
        print("   worst dt: obj", hc["obj"][m][w], "t cpu", t[w], "t gpu", hg["t0"][m][w])
    for s in (gpu, cpu):
        s.reset()
    t0 = time.time(); gpu.execute(r.scene, r.frame, r.rt, 2); tg = time.time() - t0
    t0 = time.time(); cpu.execute(r.scene, r.frame, r.rt, 2); tc = time.time() - t0
    ag, ac = gpu.accum()[0], cpu.accum()[0]
    ok = np.abs(ag - ac).max(axis=2) <= 1e-3 + 2e-3 * np.abs(ac).max(axis=2)
    fin = np.isfinite(ac).all(axis=2)
    print(f"   shared-rng: match {ok.mean():.5f} gpu finite {np.isfinite(ag).all()} mean gpu {ag[fin].mean():.6f} cpu {ac[fin].mean():.6f} t gpu {tg:.3f} cpu {tc:.3f}")
