#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path (contract: see the task brief).

Workload (BASELINE.json `metric`, configs[1]): CornellBox2 — 7 boxes + sphere + emissive
panel — 1080x1080, ssaa 2 (2160x2160 film), 1024 spp, bounce 8: 4 777 574 400 paths.
One "step" = one full render of that image.  At N GPUs the 1024 samples are split over the
devices (strong scaling) and the films are summed for the read-out.

  value   Mpaths/s, scene resident on the device, render + reduce only (CUDA events, max over ranks)
  e2e     Mpaths/s through the reference's OWN call sequence (cli.rs:157-174) with HOST buffers inside the
          timed region: set_scene (H2D) -> `for _ in 0..passes: execute()` (one C-ABI call per pass; the
          library queues them and renders full-length launches) -> reduce -> img() (tonemap + Lanczos3 +
          D2H of the u8 image)
  roofline  FP32: achieved = algorithmic flops per path-kernel launch / measured launch duration
  cpu_baseline  the CPU oracle (a literal port of the reference; the Rust binary cannot be built
          here) on the box's host cores, bounded sample of the same workload
  configs   the other BASELINE.json configs (1, 3, 4a, 4b, 5a, 5b) rendered with a bounded number of passes
          OUTSIDE the headline's timed region, each with the oracle's rate beside it (N = 1 only)
  cold_e2e  one-shot native `raytrace` process (start -> PNG on disk) with a cold and a warm kernel cache

Launched under torchrun (the driver's way) every rank owns one GPU and the films are summed onto rank 0 by one NCCL
reduce of the accumulator (0.2 ms at 8 GPUs: NVSwitch reduces in the switch).  --ipc-gather selects the alternative
without a reduce: every rank tone-maps one band of pixels reading all ranks' accumulators over CUDA IPC peer mappings
(distributed.FilmGather; measured 0.35 ms slower at 8 GPUs: its two device-side barriers cost more than NVLS saves).  `python bench.py --gpus N` WITHOUT torchrun renders through ONE context over N devices
(mrt_create_group): the library splits the samples and gathers the films over NVLink peer mappings.

`--impl reference` times the CPU port on all host threads (rank 0 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

SCENES = os.path.join(ROOT, "tests", "golden", "scenes")
SCENE = os.path.join(SCENES, "CornellBox2.json")
# flops/path = 60 + S*sum_inst(36 + C_kind) + H*153, CornellBox2: sum_inst = 533,
# S = 6.709 closest-hit calls and H = 6.178 hits per path (measured by the oracle, tests/test_oracle_stats.py)
FLOPS_PER_PATH = 60.0 + 6.709 * 533.0 + 6.178 * 153.0
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.45: SMs x lanes x FMA x max SM clock
DATA = "the reference's own scene file (example/CornellBox2.json, kept under tests/golden/scenes); fully specified, no dataset"
# BASELINE.json configs other than the headline: (label, scene, res, ssaa, rt overrides, passes rendered here)
OTHER_CONFIGS = [
    ("1 Default.json 1280x720 (direct light)", "Default", None, None, {}, 256),
    ("3 CornellBox.json 1920x1080 bounce 16", "CornellBox", (1920, 1080), 1.0, {"bounce": 16}, 128),
    ("4a Mesh.json 1920x1080", "Mesh", (1920, 1080), 1.0, {}, 128),
    ("4b Instance.json 1920x1080 (1000 spheres)", "Instance", (1920, 1080), 1.0, {}, 128),
    ("5a Minecraft.json 3840x2160 ssaa 2", "Minecraft", (3840, 2160), 2.0, {}, 32),
    ("5b dof.json 3840x2160", "dof", (3840, 2160), 1.0, {}, 128),
]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for t, r in self.rows if len(r) >= 9 and (t0 is None or t0 - 0.05 <= t <= t1 + 0.05)] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


def load_scene(args):
    import micro_raytracer_b200 as mrt
    r = mrt.load_render(SCENE)
    r.rt.sample = args.spp
    if args.res:
        r.frame.res = (args.res, args.res)
    return r


def run_reference(args, rank):
    """The reference's CPU implementation of the path: the oracle port, all host threads."""
    if rank != 0:
        return
    import oracle_lib
    r = load_scene(args)
    cpu = oracle_lib.OracleSampler(workers=0, mode=oracle_lib.FORWARD)
    nw, nh = r.frame.film_size()
    cores = os.cpu_count() or 1
    passes = max(1, args.ref_passes)
    cpu._bind(r.scene, r.frame, r.rt)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu.execute(r.scene, r.frame, r.rt, 1)
    times = []
    for _ in range(args.steps):
        cpu.reset()
        t0 = time.perf_counter()
        cpu.execute(r.scene, r.frame, r.rt, passes)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    val = nw * nh * passes * args.steps / total / 1e6
    sample = f"{passes} of {args.spp} passes of the full {nw}x{nh} film per step (per-pass cost is constant)"
    line = {
        "impl": "reference", "metric": "Mpaths/s", "value": val, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": DATA,
        "config": workload_config(args, nw, nh),
        "cpu_baseline": {"value": val, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "literal C++ port of rt.rs/sampler.rs, single trace per path; the Rust reference traces each path twice (rt.rs:957,961)"},
        "e2e": {"value": val, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, nw, nh):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    par = ("single GPU" if args.gpus == 1 else
           (f"sample split x{args.gpus}: one process per GPU (torchrun) + one NCCL reduce of the films" if not args.ipc_gather else
            f"sample split x{args.gpus}: one process per GPU (torchrun); films gathered band-wise by the tone-map kernel over CUDA IPC peer mappings") if world > 1 else
           f"sample split x{args.gpus}: ONE context over a device group (mrt_create_group), films gathered over NVLink peer mappings")
    return {"workload": f"CornellBox2.json {args.res or 1080}x{args.res or 1080} ssaa2 ({nw}x{nh} film) {args.spp} spp bounce 8 loss 0.15",
            "paths_per_step": nw * nh * args.spp, "parallelism": par,
            "l2": "no input reuse across steps: per step the only global traffic is the 74.6 MB accumulator (> L2 share), scene lives in the constant bank",
            "spp_per_launch": args.spp_per_launch}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--res", type=int, default=0, help="debug: square output resolution instead of 1080")
    ap.add_argument("--ref-passes", type=int, default=4, help="passes per step of the CPU reference arm")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs")
    ap.add_argument("--no-cold", action="store_true", help="skip the one-shot process timing")
    ap.add_argument("--ipc-gather", action="store_true", help="torchrun: gather the films band-wise over CUDA IPC peer mappings instead of one NCCL reduce")
    ap.add_argument("--spp-per-launch", type=int, default=int(os.environ.get("MRT_SPP_PER_LAUNCH", "1024")),
                    help="passes rendered by one kernel launch (one accumulator read-modify-write each)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import micro_raytracer_b200 as mrt
    from micro_raytracer_b200.distributed import FilmGather, passes_of_rank, reduce_accum

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE {world}"
    group = world == 1 and args.gpus > 1   # one process: ONE context over args.gpus devices
    n_gpus = args.gpus if group else world

    r = load_scene(args)
    nw, nh = r.frame.film_size()
    spp = args.spp
    my_passes = spp if group else passes_of_rank(spp, rank, world)
    packed = mrt.pack_scene(r.scene)

    s = mrt.Sampler(devices=list(range(args.gpus))) if group else mrt.Sampler(device=local_rank)
    # a real (non-legacy) stream shared by torch (events, NCCL ordering) and the C-ABI context
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    if not group:
        s.set_stream(stream.cuda_stream)
    s._bind(packed, r.frame, r.rt)
    s.set_partition(rank, world)
    s.spp_per_launch(args.spp_per_launch)
    acc = gather = None
    if not group:
        acc_dev, _ = s.accum_device()
        acc = torch.as_tensor(acc_dev, device=dev)
        if world > 1 and args.ipc_gather:
            gather = FilmGather(s)     # maps every rank's accumulator into every rank (CUDA IPC), once

    def barrier():
        if world > 1:
            dist.barrier()
        if group:
            s.sync()
        torch.cuda.synchronize(dev)

    def render_step():
        s.reset()
        s.execute_async(my_passes)
        if gather:
            gather.bands(spp)                         # every rank tone-maps its band of the summed film into rank 0's image
        elif not group:
            reduce_accum(s, spp, device_tensor=acc)   # NCCL reduce onto rank 0 (no-op at world 1) + pass count

    out_img = np.empty((r.frame.res[1], r.frame.res[0], 3), np.uint8)
    one_pass = s.pass_fn()            # mrt_execute(ctx, 1, NULL): the reference's per-pass call, cli.rs:163

    def e2e_step():
        s.set_scene(packed)           # H2D: scene description from host buffers (also starts a new film)
        s.set_partition(rank, world)
        for _ in range(my_passes):    # `for sample in 0..rt.sample { sampler.execute(..) }`, cli.rs:162
            one_pass()
        if gather:
            img = gather.img(spp)           # band-wise tonemap over IPC peer mappings; rank 0: Lanczos3 + D2H of the u8 image
            if rank == 0:
                out_img[...] = img
            return
        if not group:
            reduce_accum(s, spp, device_tensor=acc)
        if rank == 0:
            out_img[...] = s.img(r.frame)   # (group: gather over peer mappings +) tonemap + Lanczos3 + D2H of the u8 image

    # nvidia-smi takes ~0.5 s to print its first sample: start it before the warm-up, keep only the
    # samples whose arrival time falls inside the timed region
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()

    # ---- warm-up
    for _ in range(args.warmup):
        render_step()
    # one-time cost, like the warm-up steps themselves: with a cold kernel cache NVRTC needs ~0.5 s for the scene's specialised
    # kernel, and at N = 8 three warm-up steps last 0.14 s — without this wait the first timed steps would still run on the
    # generic kernel.  Bounded; a library without NVRTC stays on the generic kernel and says so in `jit` / `roofline.kernel`.
    t_wait = time.time()

    def jit_pending():  # on ANY rank: render_step holds a collective, so every rank must take the same number of extra steps
        st = s.jit_status()
        # (a compile that failed — no NVRTC on the box — reports its error and never becomes `compiled`: not pending)
        pending = (st["eligible"] and not st["compiled"] and not st["error"] and time.time() - t_wait < 10.0
                   and os.environ.get("MRT_JIT", "1") != "0")   # MRT_JIT=0: the generic-kernel A/B run, nothing to wait for
        if world > 1:
            flag = torch.tensor([1 if pending else 0], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            pending = bool(flag.item())
        return pending

    for _ in range(400):  # <= ~10 s
        if not jit_pending():
            break
        time.sleep(0.02)
        render_step()
    barrier()

    # ---- per-launch duration of the dominant kernel (rank-local, CUDA events on the launch stream)
    s.reset()
    n_launch = -(-(my_passes // n_gpus if group else my_passes) // s.spp_per_launch())
    if group:
        t_dev = s.device_seconds()
        s.execute_async(my_passes)
        s.sync()
        launch_ms = 1e3 * (s.device_seconds() - t_dev) / n_launch   # slowest device's CUDA-event time
        paths_per_launch = nw * nh * my_passes / n_gpus / n_launch
    else:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(stream)
        s.execute_async(my_passes)
        ev[1].record(stream)
        torch.cuda.synchronize(dev)
        launch_ms = ev[0].elapsed_time(ev[1]) / n_launch
        paths_per_launch = nw * nh * my_passes / n_launch

    # ---- timed region: K render steps, device timed, clocks sampled
    barrier()
    l0 = s.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    t_dev = s.device_seconds() if group else 0.0
    e0.record(stream)
    for _ in range(args.steps):
        render_step()
    e1.record(stream)
    barrier()
    tw1 = time.perf_counter()
    launches = s.launch_count() - l0
    # a group's devices run on their own streams: its device time is the library's per-launch CUDA-event
    # bookkeeping (per run of launches the slowest device counts), not a pair of events on one stream
    dev_ms = 1e3 * (s.device_seconds() - t_dev) if group else e0.elapsed_time(e1)
    ms = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    lt = torch.tensor([float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    clk = clocks.stop(tw0, tw1) if rank == 0 else None
    total_ms = float(ms.item())
    paths_per_step = nw * nh * spp
    value = paths_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # ---- e2e: host buffers in, u8 image out, wall clock between syncs
    for _ in range(1):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t1 = time.perf_counter()
    e2e_t = torch.tensor([t1 - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = paths_per_step * args.steps / float(e2e_t.item()) / 1e6

    if rank == 0:
        peak_tf, _ = s.fp32_peak()
        ach_tf = paths_per_launch * FLOPS_PER_PATH / (launch_ms * 1e-3) / 1e12
        line = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": DATA,
            "config": workload_config(args, nw, nh),
            "clocks": clk,
            "e2e": {"value": e2e_val, "unit": "Mpaths/s", "h2d_bytes_per_step": packed.nbytes(),
                    "d2h_bytes_per_step": int(out_img.nbytes), "ms_per_step": 1e3 * float(e2e_t.item()) / args.steps,
                    "calls_per_step": {"mrt_set_scene": 1, "mrt_execute(ctx, 1)": my_passes, "mrt_img": 1},
                    "sequence": "the reference's loop, cli.rs:157-174: Sampler::new once; per step set_scene, one execute per pass, img"},
            "gpu_launches": int(lt.item()),
            "roofline": {"bound": "fp32", "kernel": "path_kernel_jit" if s.jit_status()["launches"] else "path_kernel_param<0>", "achieved": ach_tf, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "traffic": None,
                         "peak_source": "FFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP32 entry)",
                         "peak_nominal": FP32_NOMINAL_TFLOPS, "frac_of_nominal": ach_tf / FP32_NOMINAL_TFLOPS,
                         "flops_per_path": FLOPS_PER_PATH, "paths_per_launch": paths_per_launch, "launch_ms": launch_ms,
                         "roofline_mpaths_per_gpu": FP32_NOMINAL_TFLOPS * 1e12 / FLOPS_PER_PATH / 1e6},
            "image_mean_u8": float(out_img.mean()),
            "jit": s.jit_status(),
        }
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                pj = json.load(open(prof))
                line["roofline"]["traffic"] = pj.get("path_kernel_bytes_per_launch")
                # the two hardware-side numbers an algorithmic flop count cannot flatter (ncu --set full of this
                # kernel, profiles/): share of issue slots used, active lanes per issued instruction
                for k in ("issue_active_pct", "active_lanes_per_inst", "lane_issue_slot_use", "ncu_capture"):
                    if k in pj:
                        line["roofline"][k] = pj[k]
            except Exception:  # noqa: BLE001
                pass
        # the other candidate bound, to show it is not the one: HBM traffic of the same launch
        # (algorithmic = one float4 read + one float4 write per supersampled pixel per launch)
        hbm_peak, hbm_src = 6534.1, "fallback: B200_PROFILING.md"
        try:
            hbm_peak, hbm_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:  # noqa: BLE001
            pass
        hbm_bytes = nw * nh * 16 * 2
        hbm_gbs = hbm_bytes / (launch_ms * 1e-3) / 1e9
        line["roofline"]["hbm"] = {"algorithmic_bytes_per_launch": hbm_bytes, "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s",
                                   "frac": hbm_gbs / hbm_peak, "peak_source": hbm_src}
    # the context is done: free its devices for what follows
    s.close()
    if rank == 0:
        if not args.no_cpu_baseline and n_gpus == 1:
            line["cpu_baseline"] = cpu_baseline(args, r, nw, nh)
        if not args.no_configs and n_gpus == 1:
            line["configs"] = other_configs(local_rank, with_cpu=not args.no_cpu_baseline)
        if not args.no_cold:
            line["cold_e2e"] = cold_e2e(args, n_gpus if (group or world > 1) else 1, local_rank)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, r, nw, nh):
    """The oracle (port of the reference) on the host cores, bounded sample: whole passes of the
    full film until ~cpu_baseline_seconds have elapsed."""
    import oracle_lib
    cpu = oracle_lib.OracleSampler(workers=0, mode=oracle_lib.FORWARD)
    cpu._bind(r.scene, r.frame, r.rt)
    n, t = 0, 0.0
    while t < args.cpu_baseline_seconds and n < args.spp:
        t += cpu.execute(r.scene, r.frame, r.rt, 1)
        n += 1
    return {"value": nw * nh * n / t / 1e6, "unit": "Mpaths/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{n} of {args.spp} passes of the full {nw}x{nh} film ({t:.1f} s)",
            "note": "literal C++ port, single trace per path; the Rust reference traces every path twice (rt.rs:957,961)"}


def other_configs(device, with_cpu=True):
    """BASELINE.json configs 1, 3, 4a, 4b, 5a, 5b on one GPU: full film, a bounded number of passes in ONE call
    (device seconds from CUDA events), after the scene's kernel is ready; the oracle renders one pass of a quarter-size
    film beside it (its per-path cost does not depend on the film size)."""
    import micro_raytracer_b200 as mrt
    from util import load
    rows = []
    for label, name, res, ssaa, rt, passes in OTHER_CONFIGS:
        r = load(name, res, ssaa, **rt)
        s = mrt.Sampler(device=device)
        s.execute(r.scene, r.frame, r.rt, 2)          # upload, start the scene specialisation, warm up
        t_wait = time.time()
        while s.jit_status()["eligible"] and not s.jit_status()["compiled"] and time.time() - t_wait < 8.0:
            s.execute(r.scene, r.frame, r.rt, 2)      # polls the background compile
            time.sleep(0.02)
        nw, nh, _ = s.film_size()
        s.reset()
        sec = min(s.execute(r.scene, r.frame, r.rt, passes) for _ in range(2))
        st = s.jit_status()
        row = {"workload": label, "film": [nw, nh], "passes": passes, "mpaths_s": nw * nh * passes / sec / 1e6,
               "launch_ms": 1e3 * sec / -(-passes // s.spp_per_launch()), "jit": st["launches"] > 0, "kernel": s.kernel_info()}
        # the same frame with aperture 0 (the scene's own is the reference's default 0.001, or what the file sets): the
        # specialised kernel's pinhole entry point — first hit cached per pixel (same cubin: no further compile)
        r.frame.cam.aprt = 0.0
        s.reset()
        sec0 = min(s.execute(r.scene, r.frame, r.rt, passes) for _ in range(2))
        row["aperture0_mpaths_s"] = nw * nh * passes / sec0 / 1e6
        s.close()
        if with_cpu:
            import oracle_lib
            c = oracle_lib.OracleSampler(workers=0)
            r2 = load(name, (max(1, r.frame.res[0] // 4), max(1, r.frame.res[1] // 4)), ssaa, **rt)
            t = c.execute(r2.scene, r2.frame, r2.rt, 1)
            w2, h2, _ = c.film_size()
            row["cpu_baseline"] = {"value": w2 * h2 / t / 1e6, "unit": "Mpaths/s", "cores": os.cpu_count() or 1, "kind": "port",
                                   "sample": f"1 pass of a {w2}x{h2} film ({t:.1f} s)"}
        rows.append(row)
    return rows


def cold_e2e(args, n_gpus, device):
    """One-shot render the way a user runs it: the native `raytrace` binary on the headline scene, first with an EMPTY
    kernel cache (NVRTC compiles in the background while the generic kernel renders), then with the cubin on disk.
    `process_s` = process start -> PNG on disk (includes CUDA driver / context start-up, which the reference does not
    have); `render_s` = what raytrace.rs:46-48 times: Sampler::new -> image saved; its parts are listed beside it."""
    import re
    exe = os.path.join(ROOT, "micro_raytracer_b200", "raytrace")
    if not os.path.exists(exe):
        return {"unavailable": "native raytrace binary not built"}
    out = {"binary": "micro_raytracer_b200/raytrace", "gpus": n_gpus, "spp": args.spp}
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, MRT_JIT_CACHE=os.path.join(td, "cache"))
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        cmd = [exe, SCENE, "-v", "--sample", str(args.spp), "--device", str(device if n_gpus == 1 else 0), "--gpus", str(n_gpus),
               "-o", os.path.join(td, "o.png")]
        if args.res:
            cmd += ["--res", str(args.res), str(args.res)]
        runs = []
        for _ in range(3):
            t0 = time.perf_counter()
            p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
            dt = time.perf_counter() - t0
            if p.returncode != 0:
                return {"unavailable": (p.stderr or p.stdout)[-300:]}

            def grab(pat):
                m = re.search(pat, p.stdout)
                return float(m.group(1)) if m else None
            create = grab(r"created in ([0-9.eE+-]+)s")
            runs.append({"process_s": dt, "sampler_new_s": create, "execute_loop_and_img_s": grab(r"image in host memory\): ([0-9.eE+-]+)s"),
                         "device_s": grab(r"cli:device: ([0-9.eE+-]+)s"), "save_png_s": grab(r"cli:save: ([0-9.eE+-]+)s"),
                         "render_s": (create or 0.0) + (grab(r"cli:done: ([0-9.eE+-]+)s") or 0.0)})
        out["png_bytes"] = os.path.getsize(os.path.join(td, "o.png"))
    out["cold_cache"] = runs[0]
    out["warm_cache"] = min(runs[1:], key=lambda r: r["process_s"])
    return out


if __name__ == "__main__":
    main()
