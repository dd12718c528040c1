"""ctypes binding of the CPU parity oracle (oracle/libmrt_oracle.so) — TEST INFRASTRUCTURE.

Gives the oracle the same `Sampler` face as the product so parity tests read
`gpu.execute(...)` vs `cpu.execute(...)`.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from micro_raytracer_b200 import abi
from micro_raytracer_b200.sampler import MrtError, Sampler, declare

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libmrt_oracle.so")
LITERAL, FORWARD = 0, 1
_lib = None


class CpuStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("hits", C.c_uint64),
                ("shadow_rays", C.c_uint64), ("nan_normals", C.c_uint64), ("hit_hist", C.c_uint64 * 34)]


def build_oracle(force=False):
    if os.environ.get("MRT_ORACLE_SO"):  # another build of the oracle (tools/sanitize_oracle.sh: ASan / UBSan)
        return os.environ["MRT_ORACLE_SO"]
    src = [os.path.join(ORACLE_DIR, f) for f in ("mrt_oracle.cpp", "mrt_oracle.h")] + [os.path.join(ROOT, "include", "mrt.h")]
    if (not force and os.path.exists(ORACLE_SO)
            and all(os.path.getmtime(ORACLE_SO) >= os.path.getmtime(s) for s in src if os.path.exists(s))):
        return ORACLE_SO
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)
    return ORACLE_SO


def load_oracle():
    global _lib
    if _lib is None:
        lib = declare(C.CDLL(build_oracle()), "mrt_cpu_")
        lib.mrt_cpu_set_mode.argtypes = [C.c_void_p, C.c_int]
        lib.mrt_cpu_get_stats.argtypes = [C.c_void_p, C.POINTER(CpuStats)]
        lib.mrt_cpu_path.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
        lib.mrt_cpu_resize_lanczos3.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
        lib.mrt_cpu_tonemap.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_size_t]
        lib.mrt_cpu_tonemap.restype = None
        lib.mrt_cpu_rng_block.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_float)]
        lib.mrt_cpu_rng_block.restype = None
        _lib = lib
    return _lib


class OracleSampler(Sampler):
    prefix = "mrt_cpu_"

    def __init__(self, workers=0, n_dim=64, seed=0x5EED, mode=FORWARD):
        super().__init__(workers, n_dim, 0, seed, _lib=load_oracle())
        self.set_mode(mode)

    def _create(self, workers, n_dim, device):
        if self._lib.mrt_cpu_create(C.byref(self._ctx), int(workers), int(n_dim)):
            raise MrtError("mrt_cpu_create failed")

    def set_option(self, option, value):
        self._check(self._lib.mrt_cpu_set_option(self._ctx, int(option), int(value)))

    def set_mode(self, mode):
        self._check(self._lib.mrt_cpu_set_mode(self._ctx, int(mode)))

    def stats(self):
        s = CpuStats()
        self._check(self._lib.mrt_cpu_get_stats(self._ctx, C.byref(s)))
        return {"paths": s.paths, "segments": s.segments, "hits": s.hits, "shadow_rays": s.shadow_rays,
                "nan_normals": s.nan_normals, "hit_hist": list(s.hit_hist)}

    def path(self, x, y, sample):
        out = (C.c_float * 3)()
        self._check(self._lib.mrt_cpu_path(self._ctx, x, y, sample, out))
        return np.array(out[:], np.float32)


def resize_lanczos3(src: np.ndarray, nw: int, nh: int) -> np.ndarray:
    lib = load_oracle()
    src = np.ascontiguousarray(src, np.uint8)
    h, w, _ = src.shape
    dst = np.empty((nh, nw, 3), np.uint8)
    assert lib.mrt_cpu_resize_lanczos3(src.ctypes.data, w, h, dst.ctypes.data, nw, nh) == 0
    return dst


def tonemap(values: np.ndarray, gamma: float, exp: float) -> np.ndarray:
    lib = load_oracle()
    v = np.ascontiguousarray(values, np.float32)
    out = np.empty(v.shape, np.uint8)
    lib.mrt_cpu_tonemap(v.ctypes.data, gamma, exp, out.ctypes.data, v.size)
    return out


def rng_block(pixel, sample, block, seed=0x5EED):
    out = (C.c_float * 4)()
    load_oracle().mrt_cpu_rng_block(pixel, sample, block, seed, out)
    return np.array(out[:], np.float32)
