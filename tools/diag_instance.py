"""Diagnostic dump for the Instance.json mean discrepancy (GPU 0.85 % darker than the oracle at 64 spp, independent
seeds): per-pass radiance with SHARED random numbers, GPU and oracle, so that single paths can be compared here."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import micro_raytracer_b200 as mrt
import oracle_lib
from util import load

r = load("Instance", (256, 144), 1.0)
out = {}
for jit in (0, 2):
    g = mrt.Sampler(device=0); g.set_option(2, jit)
    passes = []
    prev = None
    for k in range(16):
        g.execute(r.scene, r.frame, r.rt, 1)
        a = g.accum()[0].astype(np.float64)
        passes.append(a if prev is None else a - prev); prev = a
    out[f"gpu_jit{jit}"] = np.stack(passes).astype(np.float32)
c = oracle_lib.OracleSampler()
passes = []; prev = None
for k in range(16):
    c.execute(r.scene, r.frame, r.rt, 1)
    a = c.accum()[0].astype(np.float64)
    passes.append(a if prev is None else a - prev); prev = a
out["cpu"] = np.stack(passes).astype(np.float32)
import os
os.environ["MRT_NO_BVH"] = "1"
g = mrt.Sampler(device=0); g.set_option(2, 0)
g.execute(r.scene, r.frame, r.rt, 16)
out["gpu_brute_sum"] = g.accum()[0]
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag_instance.npz"), **out)
for k, v in out.items():
    print(k, v.shape, float(v.mean()))
