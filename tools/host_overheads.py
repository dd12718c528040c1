"""Host-side cost of the reference-facing calls (what e2e adds to the device time): Sampler creation over 1 / all GPUs,
set_scene (pack is excluded: a PackedScene is passed), one-pass execute calls (queued), img()."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import micro_raytracer_b200 as mrt
from util import load

out = {}
r = load("CornellBox2")
packed = mrt.pack_scene(r.scene)
for label, kw in (("1 gpu", {"device": 0}), ("all gpus", {"devices": "all"})):
    t = time.perf_counter(); s = mrt.Sampler(**kw); t_new = time.perf_counter() - t
    s.set_option(2, 2)
    s._bind(packed, r.frame, r.rt)
    s.execute(packed, r.frame, r.rt, 8); s.img(r.frame)
    ts = []
    for _ in range(5):
        t = time.perf_counter(); s.set_scene(packed); ts.append(time.perf_counter() - t)
    f = s.pass_fn()
    t = time.perf_counter()
    for _ in range(1024): f()
    t_loop = time.perf_counter() - t
    t = time.perf_counter(); s.img(r.frame); t_img = time.perf_counter() - t   # renders the 1024 queued passes too
    s.sync()
    s.execute(packed, r.frame, r.rt, 0)
    t = time.perf_counter(); s.img(r.frame); t_img2 = time.perf_counter() - t  # read-out only
    out[label] = {"n_devices": s.group_info()["n_devices"], "sampler_new_s": t_new, "set_scene_ms": 1e3 * min(ts), "one_pass_call_us": 1e6 * t_loop / 1024,
                  "img_with_render_ms": 1e3 * t_img, "img_readout_only_ms": 1e3 * t_img2}
    s.close()
print(json.dumps(out, indent=1))
