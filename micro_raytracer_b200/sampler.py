"""Host-side mirror of the reference's `Sampler` (src/sampler.rs:11-99) over the C ABI.

    sampler = Sampler(workers, n_dim)              # ≙ Sampler::new      sampler.rs:19
    dt = sampler.execute(scene, frame, rt)         # ≙ Sampler::execute  sampler.rs:28 (one pass)
    img = sampler.img(frame)                       # ≙ Sampler::img      sampler.rs:80

`Sampler(devices=[0, 1, ...])` (or `devices="all"`) is the same object over several GPUs of the box
(mrt_create_group): the passes are split over the devices inside the library and `img()` gathers their
films over NVLink peer mappings.  One-pass `execute` calls — the reference's loop — are queued by the
library and rendered in full-length launches (include/mrt.h: mrt_execute).

exactly as CLI::raytrace drives it (src/cli.rs:155-177).  All computation happens in the
CUDA library `libmrt.so` (csrc/, built by __graft_entry__.build()); there is no CPU
fallback — a missing library or device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

from . import abi
from .scene import Frame, PackedScene, RayTracer, Scene, pack_scene

_LIB_NAME = "libmrt.so"
OPT_NORMAL_SPACE, NORMAL_FORWARD_XF, NORMAL_OBJECT = 1, 0, 1  # include/mrt.h
OPT_JIT, JIT_OFF, JIT_AUTO, JIT_FORCE = 2, 0, 1, 2
OPT_COALESCE = 3
_lib = None


class MrtError(RuntimeError):
    """≙ the String of the reference's Result<_, String>."""


def lib_path() -> str:
    """In-tree libmrt.so; MRT_LIB points at another build of the same library (kernel experiments)."""
    return os.environ.get("MRT_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def declare(lib, prefix: str = "mrt_"):
    """Attach argtypes/restype for the entry points of include/mrt.h."""
    P, u32, f32, u64 = C.c_void_p, C.c_uint32, C.c_float, C.c_uint64

    def fn(name, *args, res=C.c_int):
        f = getattr(lib, prefix + name)
        f.argtypes, f.restype = list(args), res
        return f

    fn("create", C.POINTER(P), C.c_int, u32, u32) if prefix == "mrt_" else fn("create", C.POINTER(P), u32, u32)
    fn("destroy", P, res=None)
    fn("last_error", P, res=C.c_char_p)
    fn("set_scene", P, C.POINTER(abi.MrtScene))
    fn("set_frame", P, C.POINTER(abi.MrtFrame))
    fn("set_rt", P, u32, f32, u64)
    fn("set_partition", P, u32, u32)
    fn("set_option", P, u32, u32)
    fn("execute", P, u32, C.POINTER(C.c_double))
    fn("reset", P)
    fn("film_size", P, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32))
    fn("accum", P, C.POINTER(f32), C.POINTER(u32))
    fn("img", P, C.POINTER(C.c_uint8))
    fn("img_ss", P, C.POINTER(C.c_uint8))
    fn("trace_primary", P, C.c_void_p)
    if prefix == "mrt_":
        fn("abi_version")
        fn("create_group", C.POINTER(P), C.POINTER(C.c_int), C.c_int, u32, u32)
        fn("group_info", P, C.POINTER(u32), C.POINTER(u32))
        fn("device_count", C.POINTER(C.c_int))
        fn("update_scene", P, C.POINTER(abi.MrtScene))
        fn("update_frame", P, C.POINTER(abi.MrtFrame))
        fn("device_seconds", P, C.POINTER(C.c_double))
        fn("scene_info", P, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32))
        fn("ipc_export", P, C.c_char_p, C.c_char_p)
        fn("ipc_attach", P, u32, u32, C.c_char_p, C.c_char_p)
        fn("ipc_tonemap_band", P, u32)
        fn("img_gathered", P, C.POINTER(C.c_uint8))
        fn("execute_async", P, u32)
        fn("sync", P)
        fn("accum_device", P, C.POINTER(P), C.POINTER(C.c_size_t), C.POINTER(P))
        fn("set_passes", P, u32)
        fn("set_stream", P, P)
        fn("launch_count", P, C.POINTER(u64))
        fn("spp_per_launch", P, u32, C.POINTER(u32))
        fn("jit_status", P, C.POINTER(u32), C.POINTER(u32), C.POINTER(u64), C.POINTER(C.c_double))
        fn("fp32_peak", P, C.POINTER(C.c_double), C.POINTER(C.c_double))
    return lib


def load_library():
    """Load libmrt.so (in-tree).  Fails loudly: there is no other implementation."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise MrtError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        _lib = declare(C.CDLL(p))
        if _lib.mrt_abi_version() != abi.MRT_ABI_VERSION:
            raise MrtError("libmrt.so ABI version mismatch")
    return _lib


class _DeviceArray:
    """Minimal __cuda_array_interface__ holder so torch can wrap the accumulator in place."""

    def __init__(self, ptr: int, n: int, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        self._owner = owner


class Sampler:
    """≙ `Sampler` of src/sampler.rs.  `workers` / `n_dim` (--worker / --dim, cli.rs:157) are
    accepted and ignored: the CUDA grid replaces the tile pool."""

    prefix = "mrt_"

    def __init__(self, workers: int = 24, n_dim: int = 64, device: int = 0, seed: int = 0x5EED, _lib=None, devices=None):
        self._lib = _lib if _lib is not None else load_library()
        self._ctx = C.c_void_p()
        self._stream = None
        if devices is not None:
            self._create_group(workers, n_dim, devices)
        else:
            self._create(workers, n_dim, device)
        self.seed = seed
        self._scene_key = None
        self._frame_key = None
        self._rt_key = None
        self._packed: Optional[PackedScene] = None

    # -- plumbing
    def _f(self, name):
        return getattr(self._lib, self.prefix + name)

    def _create(self, workers, n_dim, device):
        rc = self._lib.mrt_create(C.byref(self._ctx), int(device), int(workers), int(n_dim))
        if rc:
            raise MrtError((self._lib.mrt_last_error(None) or b"mrt_create failed").decode())

    def _create_group(self, workers, n_dim, devices):
        """devices: a list of CUDA device indices, or "all"."""
        if isinstance(devices, str):
            if devices != "all":
                raise ValueError("devices must be a list of device indices or 'all'")
            arr, n = None, 0
        else:
            devices = [int(d) for d in devices]
            arr, n = (C.c_int * len(devices))(*devices), len(devices)
            if n == 0:
                raise ValueError("empty device list")
        rc = self._lib.mrt_create_group(C.byref(self._ctx), arr, n, int(workers), int(n_dim))
        if rc:
            raise MrtError((self._lib.mrt_last_error(None) or b"mrt_create_group failed").decode())

    def _check(self, rc):
        if rc:
            raise MrtError((self._f("last_error")(self._ctx) or b"error").decode())

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._f("destroy")(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # -- uploads (the three borrows of Sampler::execute); re-sent only when they change
    def set_scene(self, scene):
        packed = scene if isinstance(scene, PackedScene) else pack_scene(scene)
        self._check(self._f("set_scene")(self._ctx, C.byref(packed.c)))
        self._packed = packed

    def set_frame(self, frame: Frame):
        f = frame.pack()
        self._check(self._f("set_frame")(self._ctx, C.byref(f)))

    def set_rt(self, rt: RayTracer):
        self._check(self._f("set_rt")(self._ctx, int(rt.bounce), float(rt.loss), int(self.seed)))

    def set_option(self, option: int, value: int):
        """mrt_set_option; the scene is re-sent on the next execute so the option takes effect."""
        self._check(self._f("set_option")(self._ctx, int(option), int(value)))
        self._scene_key = None

    def set_partition(self, rank: int, world: int):
        self._check(self._f("set_partition")(self._ctx, int(rank), int(world)))

    @staticmethod
    def _key_of(frame: Frame):
        return (tuple(frame.res), frame.ssaa, tuple(frame.cam.pos), tuple(frame.cam.dir), frame.cam.fov,
                frame.cam.gamma, frame.cam.exp, frame.cam.aprt, frame.cam.foc)

    def _bind(self, scene, frame: Frame, rt: RayTracer):
        sk = id(scene)
        if sk != self._scene_key:
            self.set_scene(scene)
            self._scene_key = sk
            self._scene_ref = scene
        fk = self._key_of(frame)
        if fk != self._frame_key:
            self.set_frame(frame)
            self._frame_key = fk
        rk = (rt.bounce, rt.loss, self.seed)
        if rk != self._rt_key:
            self.set_rt(rt)
            self._rt_key = rk

    # -- the reference API
    def execute(self, scene, frame: Frame, rt: RayTracer, n_passes: int = 1) -> float:
        """One pass (or n_passes) = one path per supersampled pixel; returns device seconds (for a
        one-pass call: the amortised time of the launches that finished since the last call)."""
        self._bind(scene, frame, rt)
        sec = C.c_double()
        self._check(self._f("execute")(self._ctx, int(n_passes), C.byref(sec)))
        return sec.value

    def img(self, frame: Optional[Frame] = None) -> np.ndarray:
        """(res_h, res_w, 3) uint8, ≙ Sampler::img."""
        if frame is not None and self._frame_key is None:
            self.set_frame(frame)
            self._frame_key = self._key_of(frame)
        w, h = self._res
        out = np.empty((h, w, 3), np.uint8)
        self._check(self._f("img")(self._ctx, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    # -- extras used by tests / bench
    @property
    def _res(self) -> Tuple[int, int]:
        fk = self._frame_key
        if fk is None:
            raise MrtError("no frame set")
        return fk[0]

    def film_size(self) -> Tuple[int, int, int]:
        nw, nh, p = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self._f("film_size")(self._ctx, C.byref(nw), C.byref(nh), C.byref(p)))
        return nw.value, nh.value, p.value

    def reset(self):
        self._check(self._f("reset")(self._ctx))

    def accum(self) -> Tuple[np.ndarray, int]:
        nw, nh, _ = self.film_size()
        out = np.empty((nh, nw, 3), np.float32)
        p = C.c_uint32()
        self._check(self._f("accum")(self._ctx, out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(p)))
        return out, p.value

    def img_ss(self) -> np.ndarray:
        nw, nh, _ = self.film_size()
        out = np.empty((nh, nw, 3), np.uint8)
        self._check(self._f("img_ss")(self._ctx, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def trace_primary(self) -> np.ndarray:
        nw, nh, _ = self.film_size()
        out = np.zeros((nh, nw), dtype=np.dtype(abi.HIT_DTYPE))
        assert out.dtype.itemsize == C.sizeof(abi.MrtHit)
        self._check(self._f("trace_primary")(self._ctx, out.ctypes.data))
        return out

    # -- CUDA-only
    def execute_async(self, n_passes: int):
        self._check(self._lib.mrt_execute_async(self._ctx, int(n_passes)))

    def sync(self):
        self._check(self._lib.mrt_sync(self._ctx))

    def accum_device(self):
        """(object exposing __cuda_array_interface__ over nw*nh*4 floats, stream handle)."""
        p, n, s = C.c_void_p(), C.c_size_t(), C.c_void_p()
        self._check(self._lib.mrt_accum_device(self._ctx, C.byref(p), C.byref(n), C.byref(s)))
        return _DeviceArray(p.value, n.value, self), s.value

    def set_stream(self, cuda_stream: Optional[int]):
        """Queue this sampler's work on a caller-owned stream (e.g. torch's current stream)."""
        self._check(self._lib.mrt_set_stream(self._ctx, C.c_void_p(cuda_stream or 0)))
        self._stream = cuda_stream or None

    @property
    def stream(self) -> Optional[int]:
        """The caller-owned stream handle set with set_stream (None: the context's private stream)."""
        return self._stream

    def update_scene(self, packed: PackedScene):
        """mrt_update_scene: a no-op when the context already holds this very content (keeps the passes)."""
        self._check(self._lib.mrt_update_scene(self._ctx, C.byref(packed.c)))
        self._packed = packed

    def update_frame(self, frame: Frame):
        f = frame.pack()
        self._check(self._lib.mrt_update_frame(self._ctx, C.byref(f)))
        self._frame_key = self._key_of(frame)

    def pass_fn(self):
        """The reference's per-pass call with the Python overhead stripped: returns a zero-argument callable that
        is `mrt_execute(ctx, 1, NULL)` (scene, frame and rt must already be bound).  For host loops that call it
        a thousand times per image (bench.py's e2e leg)."""
        f, ctx = self._lib.mrt_execute, self._ctx

        def one_pass():
            if f(ctx, 1, None):
                self._check(1)
        return one_pass

    # -- film gather across processes (include/mrt.h: mrt_ipc_*), driven by distributed.gather_film
    def ipc_export(self) -> Tuple[bytes, bytes]:
        a, i = C.create_string_buffer(64), C.create_string_buffer(64)
        self._check(self._lib.mrt_ipc_export(self._ctx, a, i))
        return a.raw, i.raw

    def ipc_attach(self, rank: int, world: int, accum_handles, film_image_handle: bytes):
        self._check(self._lib.mrt_ipc_attach(self._ctx, int(rank), int(world), b"".join(accum_handles), film_image_handle))

    def ipc_tonemap_band(self, total_passes: int):
        self._check(self._lib.mrt_ipc_tonemap_band(self._ctx, int(total_passes)))

    def img_gathered(self) -> np.ndarray:
        w, h = self._res
        out = np.empty((h, w, 3), np.uint8)
        self._check(self._lib.mrt_img_gathered(self._ctx, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def device_seconds(self) -> float:
        t = C.c_double()
        self._check(self._lib.mrt_device_seconds(self._ctx, C.byref(t)))
        return t.value

    def kernel_info(self) -> dict:
        """How the scene is rendered: scene-level BVH or unrolled, run-time specialised kernel in use, feature mask."""
        b, p, f = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self._lib.mrt_scene_info(self._ctx, C.byref(b), C.byref(p), C.byref(f)))
        return {"scene_bvh": bool(b.value), "specialised": bool(p.value), "features": f.value}

    def group_info(self) -> dict:
        n, p = C.c_uint32(), C.c_uint32()
        self._check(self._lib.mrt_group_info(self._ctx, C.byref(n), C.byref(p)))
        return {"n_devices": n.value, "peer_access": bool(p.value)}

    def set_passes(self, passes: int):
        self._check(self._lib.mrt_set_passes(self._ctx, int(passes)))

    def spp_per_launch(self, spp: int = 0) -> int:
        """Set (spp > 0) or query the number of passes one kernel launch renders."""
        cur = C.c_uint32()
        self._check(self._lib.mrt_spp_per_launch(self._ctx, int(spp), C.byref(cur)))
        return cur.value

    def jit_status(self) -> dict:
        """State of the run-time scene-specialised kernel (include/mrt.h: MRT_OPT_JIT)."""
        e, k, n, t = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_double()
        self._check(self._lib.mrt_jit_status(self._ctx, C.byref(e), C.byref(k), C.byref(n), C.byref(t)))
        return {"eligible": bool(e.value), "compiled": bool(k.value), "launches": n.value, "compile_seconds": abs(t.value),
                "from_disk_cache": t.value < 0,
                "error": (self._lib.mrt_last_error(self._ctx) or b"").decode() if not k.value else ""}

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._check(self._lib.mrt_launch_count(self._ctx, C.byref(n)))
        return n.value

    def fp32_peak(self) -> Tuple[float, float]:
        t, s = C.c_double(), C.c_double()
        self._check(self._lib.mrt_fp32_peak(self._ctx, C.byref(t), C.byref(s)))
        return t.value, s.value
