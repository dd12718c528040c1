"""C ABI v2 behaviour on the GPU: the reference's own call pattern (one Sampler, one execute per pass, img()) must
reach the batched throughput — pass coalescing — and span every GPU of the box — device groups; plus the hardening
items of round 1's review (empty scenes, failed uploads, content-keyed scene updates, stream ordering)."""
import ctypes as C
import os
import time

import numpy as np
import pytest

import micro_raytracer_b200 as mrt
import oracle_lib
from micro_raytracer_b200.sampler import JIT_FORCE, JIT_OFF, OPT_COALESCE, OPT_JIT, MrtError
from util import load

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return torch.cuda.device_count()


# ---------------------------------------------------------------- pass coalescing (cli.rs:162-163 / http.rs:141-142)
@pytest.mark.parametrize("jit", [JIT_OFF, JIT_FORCE], ids=["generic", "jit"])
def test_one_pass_calls_render_the_batched_launches_bit_for_bit(jit):
    """`for _ in 0..n { execute }; img()` == `execute(n); img()`: the one-pass calls are only queued, so the very same
    launches run.  Also: film_size counts queued passes, a mid-way img() (the --update case) flushes, and the
    amortised per-pass times add up to the device time."""
    r = load("CornellBox2", (96, 96), 2.0)
    n = 24
    a, b = mrt.Sampler(device=0), mrt.Sampler(device=0)
    for s in (a, b):
        s.set_option(OPT_JIT, jit)
    l0 = a.launch_count()
    reported = 0.0
    for k in range(n):
        reported += a.execute(r.scene, r.frame, r.rt)
        assert a.film_size()[2] == k + 1
    assert a.launch_count() == l0, "one-pass calls must not launch before the film is needed"
    img_a = a.img(r.frame)
    assert a.launch_count() > l0
    b.execute(r.scene, r.frame, r.rt, n)
    assert np.array_equal(a.accum()[0], b.accum()[0])
    assert np.array_equal(img_a, b.img(r.frame))
    a.sync()
    reported += a.execute(r.scene, r.frame, r.rt, 0)  # n_passes = 0: nothing to render, hands out what is unreported
    assert reported == pytest.approx(a.device_seconds(), rel=1e-6) and reported > 0.0
    # --update: an image after every pass (cli.rs:166-169) still works and still adds up
    c = mrt.Sampler(device=0)
    c.set_option(OPT_JIT, jit)
    for k in range(3):
        c.execute(r.scene, r.frame, r.rt)
        assert c.img(r.frame).shape == (96, 96, 3)
    d = mrt.Sampler(device=0)
    d.set_option(OPT_JIT, jit)
    d.set_option(OPT_COALESCE, 0)
    for k in range(3):
        d.execute(r.scene, r.frame, r.rt)
    assert d.launch_count() >= 3
    np.testing.assert_allclose(c.accum()[0], d.accum()[0], rtol=2e-5, atol=2e-6)


def test_reference_loop_reaches_the_batched_throughput_on_the_headline_render():
    """SURVEY 8(d) config 2 at full size, 1024 passes: the reference's loop of one-pass calls + img() within 3 % of
    one batched call + img() (round 1: -22 %), image identical."""
    r = load("CornellBox2")
    n = 1024
    s = mrt.Sampler(device=0)
    s.set_option(OPT_JIT, JIT_FORCE)
    s.execute(r.scene, r.frame, r.rt, 8)  # upload, compile, warm up
    s.img(r.frame)

    def batched():
        s.reset()
        t = time.perf_counter()
        s.execute(r.scene, r.frame, r.rt, n)
        im = s.img(r.frame)
        return time.perf_counter() - t, im

    def looped():
        s.reset()
        t = time.perf_counter()
        for _ in range(n):
            s.execute(r.scene, r.frame, r.rt)
        im = s.img(r.frame)
        return time.perf_counter() - t, im

    tb, ib = min((batched() for _ in range(2)), key=lambda x: x[0])
    tl, il = min((looped() for _ in range(2)), key=lambda x: x[0])
    assert np.array_equal(ib, il)
    assert tl <= 1.03 * tb + 0.002, (tl, tb)


def test_settings_changes_flush_the_queue_first():
    """Queued passes were asked for under the old RayTracer settings: they are rendered before the change."""
    r = load("CornellBox2", (64, 64), 1.0)
    a, b = mrt.Sampler(device=0), mrt.Sampler(device=0)
    for bounce, n in ((8, 2), (2, 3)):
        r.rt.bounce = bounce                      # the caller edits its RayTracer between passes
        for _ in range(n):
            a.execute(r.scene, r.frame, r.rt)     # one-pass calls: queued; set_rt flushes the first two
        b.execute(r.scene, r.frame, r.rt, n)
    np.testing.assert_allclose(a.accum()[0], b.accum()[0], rtol=2e-5, atol=2e-6)
    assert a.accum()[1] == b.accum()[1] == 5
    c = mrt.Sampler(device=0)
    r.rt.bounce = 8
    c.execute(r.scene, r.frame, r.rt, 5)
    assert not np.allclose(c.accum()[0], a.accum()[0], rtol=1e-3)   # bounce 2 really was used for three of them


# ---------------------------------------------------------------- content-keyed scene / frame updates
def test_update_scene_keeps_the_film_for_identical_content_and_restarts_it_otherwise():
    r = load("dof", (64, 36), 1.0)
    s = mrt.Sampler(device=0)
    s.execute(r.scene, r.frame, r.rt, 2)
    p1 = mrt.pack_scene(r.scene)            # a NEW packing of the same content, at other addresses
    s.update_scene(p1)
    s.update_frame(r.frame)
    assert s.film_size()[2] == 2
    r.scene.renderer[0].mat.albedo = (0.1, 0.2, 0.3)  # mutated in place: same Python object, other content
    s.update_scene(mrt.pack_scene(r.scene))
    assert s.film_size()[2] == 0
    s.execute(r.scene, r.frame, r.rt, 1)
    r.frame.cam.fov += 1.0
    s.update_frame(r.frame)
    assert s.film_size()[2] == 0


def test_empty_scene_renders_the_sky():
    """A description without renderers is valid (every key is optional, parser.rs:152-158): all rays miss, the image
    is sky.color (rt.rs:958).  NULL array pointers with zero counts must be accepted."""
    r = mrt.render_from_dict({"frame": {"res": [32, 16]}, "scene": {"sky": {"color": [0.2, 0.4, 0.8], "pwr": 0.5}}})
    g, c = mrt.Sampler(device=0), oracle_lib.OracleSampler()
    for s in (g, c):
        s.execute(r.scene, r.frame, r.rt, 2)
    ag, ac = g.accum()[0], c.accum()[0]
    np.testing.assert_allclose(ag, ac, rtol=1e-6)
    np.testing.assert_allclose(ag[0, 0] / 2, [0.2, 0.4, 0.8], rtol=1e-6)
    assert np.array_equal(g.img(r.frame), c.img(r.frame))
    # straight through the C ABI with NULL arrays
    from micro_raytracer_b200 import abi
    sc = abi.MrtScene()
    sc.sky_color[:] = [0.2, 0.4, 0.8]
    sc.sky_pwr = 0.5
    lib = mrt.sampler.load_library()
    assert lib.mrt_set_scene(g._ctx, C.byref(sc)) == 0
    sc.n_objects = 3  # a count without an array is an error, not a crash
    assert lib.mrt_set_scene(g._ctx, C.byref(sc)) == abi.MRT_ERR_INVALID


def test_rejected_scene_leaves_the_old_one_in_place():
    """Validation happens before any device state is touched: after a rejected mrt_set_scene the context still renders
    the scene it held (round 1 freed the old buffers first and kept dangling pointers)."""
    r = load("Default", (64, 36), 1.0)
    s = mrt.Sampler(device=0)
    s.execute(r.scene, r.frame, r.rt, 1)
    ref = s.accum()[0].copy()
    bad = load("Default", (64, 36), 1.0)
    bad.scene.renderer[0].mat.emit = 2.0  # gen_bool(2.0) panics in the reference: rejected here
    with pytest.raises((MrtError, ValueError)):
        s.set_scene(bad.scene)
    bad.scene.renderer[0].mat.emit = 0.0
    bad.scene.light[0].kind = "point"
    packed = mrt.pack_scene(bad.scene)
    packed.lights[0].kind = 7
    with pytest.raises(MrtError):
        s.set_scene(packed)
    s.reset()
    s._scene_key = id(r.scene)
    s.execute(r.scene, r.frame, r.rt, 1)
    assert np.array_equal(s.accum()[0], ref)


def test_device_reduce_is_ordered_without_a_shared_stream():
    """distributed.reduce_accum(device_tensor=...) on a sampler that was NOT bound to torch's stream: the helper
    synchronises both sides itself (advisor finding of round 1)."""
    import torch
    from micro_raytracer_b200.distributed import render_distributed
    r = load("CornellBox2", (128, 128), 2.0)
    s = mrt.Sampler(device=0)
    s._bind(r.scene, r.frame, r.rt)
    acc = torch.as_tensor(s.accum_device()[0], device="cuda:0")
    render_distributed(s, r.scene, r.frame, r.rt, 16, device_tensor=acc)
    got = acc.clone().cpu().numpy().reshape(256, 256, 4)[..., :3]
    t = mrt.Sampler(device=0)
    t.execute(r.scene, r.frame, r.rt, 16)
    assert np.array_equal(got, t.accum()[0])
    assert s.film_size()[2] == 16 and np.array_equal(s.img(r.frame), t.img(r.frame))


# ---------------------------------------------------------------- device groups (≥ 2 GPUs)
needs2 = pytest.mark.skipif("_n_devices() < 2", reason="needs two GPUs")


@needs2
@pytest.mark.parametrize("p2p", [True, False], ids=["peer", "staged"])
@pytest.mark.parametrize("name", ["CornellBox2", "Mesh"])
def test_group_context_renders_the_single_device_image(name, p2p, monkeypatch):
    """One Sampler over all GPUs (mrt_create_group): same paths as on one device (RNG keyed by the global sample index),
    films gathered by the tonemap kernel over peer mappings (or staged copies with MRT_NO_P2P); accum(), img_ss(),
    img() agree with the one-device render up to the f32 summation order."""
    if not p2p:
        monkeypatch.setenv("MRT_NO_P2P", "1")
    n_dev = min(_n_devices(), 8)
    r = load(name, (101, 67), 2.0)  # 202 x 134 film: not a multiple of anything
    one = mrt.Sampler(device=0)
    grp = mrt.Sampler(devices=list(range(n_dev)))
    info = grp.group_info()
    assert info["n_devices"] == n_dev and info["peer_access"] == p2p
    n = 2 * n_dev + 1
    one.execute(r.scene, r.frame, r.rt, n)
    for _ in range(n):                      # the reference's loop
        grp.execute(r.scene, r.frame, r.rt)
    a1, ag = one.accum()[0], grp.accum()
    assert ag[1] == n
    np.testing.assert_allclose(ag[0], a1, rtol=3e-5, atol=3e-6)
    s1, sg = one.img_ss().astype(int), grp.img_ss().astype(int)
    assert np.abs(s1 - sg).max() <= 1 and (s1 == sg).mean() > 0.999
    i1, ig = one.img(r.frame).astype(int), grp.img(r.frame).astype(int)
    assert np.abs(i1 - ig).max() <= 1 and (i1 == ig).mean() > 0.999
    # more passes after a read-out keep accumulating on every device
    one.execute(r.scene, r.frame, r.rt, 3)
    grp.execute(r.scene, r.frame, r.rt, 3)
    np.testing.assert_allclose(grp.accum()[0], one.accum()[0], rtol=3e-5, atol=3e-6)
    # the device view (for an outer reduce across processes) is the whole group's film on the first device
    import torch
    view = torch.as_tensor(grp.accum_device()[0], device="cuda:0")
    grp.sync()
    torch.cuda.synchronize()
    np.testing.assert_allclose(view.cpu().numpy().reshape(134, 202, 4)[..., :3], one.accum()[0], rtol=3e-5, atol=3e-6)
    np.testing.assert_allclose(grp.accum()[0], one.accum()[0], rtol=3e-5, atol=3e-6)  # nothing lost by the collapse
    assert grp.trace_primary().shape == (134, 202)


@needs2
def test_group_context_errors():
    lib = mrt.sampler.load_library()
    with pytest.raises(MrtError):
        mrt.Sampler(devices=[0, 0])
    with pytest.raises(MrtError):
        mrt.Sampler(devices=[0, 99])
    g = mrt.Sampler(devices="all")
    assert g.group_info()["n_devices"] == _n_devices()
    with pytest.raises(MrtError):
        g.img()  # no frame yet
    one = mrt.Sampler(devices=[0])  # a one-device list is a plain context
    assert one.group_info() == {"n_devices": 1, "peer_access": True}
    assert lib is not None


@needs2
def test_group_scales_the_headline_render():
    """Strong scaling through ONE context: the headline render (1024 spp, the reference's loop) on all devices vs one."""
    n_dev = min(_n_devices(), 8)
    r = load("CornellBox2")
    t = {}
    for devs in ([0], list(range(n_dev))):
        s = mrt.Sampler(devices=devs)
        s.set_option(OPT_JIT, JIT_FORCE)
        s.execute(r.scene, r.frame, r.rt, 8)
        s.img(r.frame)
        s.reset()
        t0 = time.perf_counter()
        for _ in range(1024):
            s.execute(r.scene, r.frame, r.rt)
        s.img(r.frame)
        t[len(devs)] = time.perf_counter() - t0
    assert t[n_dev] < t[1] / (0.8 * n_dev), t
