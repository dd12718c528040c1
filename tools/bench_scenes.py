"""Throughput of every BASELINE.json config (not the headline bench): Mpaths/s per scene on one GPU,
render-only (CUDA-event seconds returned by mrt_execute), at the configs' full resolutions with a
bounded number of passes; optionally the oracle's rate on the host cores beside it.
    python tools/bench_scenes.py [--cpu] [--only Scene] [--only3] [--passes N] [--pinhole]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import micro_raytracer_b200 as mrt
from util import load

# (label, scene, res, ssaa, rt overrides, passes timed here)
CONFIGS = [
    ("1 Default 1280x720 direct light", "Default", None, None, {}, 256),
    ("2 CornellBox2 1080^2 ssaa2 (headline)", "CornellBox2", None, None, {}, 256),
    ("3 CornellBox (README 10-object) 1920x1080 bounce16", "CornellBox", (1920, 1080), 1.0, {"bounce": 16}, 256),
    ("4a Mesh 1920x1080", "Mesh", (1920, 1080), 1.0, {}, 128),
    ("4b Instance (1000 spheres) 1920x1080", "Instance", (1920, 1080), 1.0, {}, 128),
    ("5a Minecraft 3840x2160 ssaa2", "Minecraft", (3840, 2160), 2.0, {}, 64),
    ("5b dof 3840x2160", "dof", (3840, 2160), 1.0, {}, 128),
]

def main():
    cpu = "--cpu" in sys.argv
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    force_passes = int(sys.argv[sys.argv.index("--passes") + 1]) if "--passes" in sys.argv else None
    only3 = "--only3" in sys.argv  # the three scenes that are searched through a BVH
    pinhole = "--pinhole" in sys.argv  # aperture 0 instead of the scene's (default 0.001): the specialised kernel's pinhole entry point
    rows = []
    for label, name, res, ssaa, rt, passes in CONFIGS:
        if (only and name != only) or (only3 and name not in ("Mesh", "Instance", "Minecraft")):
            continue
        passes = force_passes or passes
        r = load(name, res, ssaa, **rt)
        if pinhole:
            r.frame.cam.aprt = 0.0
            label += " [aprt 0]"
        s = mrt.Sampler(device=0)
        s.execute(r.scene, r.frame, r.rt, 1)  # upload + warm-up; starts the background scene specialisation
        t_wait = time.time()
        while (os.environ.get("MRT_JIT", "1") != "0" and s.jit_status()["eligible"] and not s.jit_status()["compiled"]
               and time.time() - t_wait < 5.0):
            s.execute(r.scene, r.frame, r.rt, 1)  # polls the compile; gives up quietly when NVRTC is unavailable / disabled
            time.sleep(0.02)
        nw, nh, _ = s.film_size()
        s.reset()
        sec = s.execute(r.scene, r.frame, r.rt, passes)
        sec = min(sec, s.execute(r.scene, r.frame, r.rt, passes))
        row = {"config": label, "film": [nw, nh], "passes": passes, "gpu_mpaths_s": nw * nh * passes / sec / 1e6,
               "jit": s.jit_status()["launches"] > 0, "kernel": s.kernel_info()}
        if cpu:
            import oracle_lib
            c = oracle_lib.OracleSampler(workers=0)
            r2 = load(name, (max(1, r.frame.res[0] // 4), max(1, r.frame.res[1] // 4)), ssaa, **rt)
            t = c.execute(r2.scene, r2.frame, r2.rt, 1)
            w2, h2, _ = c.film_size()
            row["cpu_mpaths_s"] = w2 * h2 / t / 1e6
            row["cpu_cores"] = os.cpu_count()
        rows.append(row)
        print(json.dumps(row), flush=True)
    return rows

if __name__ == "__main__":
    main()
