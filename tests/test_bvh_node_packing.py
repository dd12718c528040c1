"""Host logic of the 48-byte BVH node (csrc/mrt_scene.cu): the half extents are stored as fp16 ROUNDED UP, so a node box can
only grow (the candidate set of a traversal stays a superset and the results bit-identical to brute force).  The conversion
is compiled here from the library's own source text and checked against numpy's float16."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _half_up(tmp_path):
    src = open(os.path.join(ROOT, "micro_raytracer_b200", "csrc", "mrt_scene.cu")).read()
    a, b = src.index("uint32_t half_up(float v) {"), src.index("uint32_t half2_up")
    cpp = tmp_path / "half_up.cpp"
    cpp.write_text('#include <cstdint>\n#include <cstring>\n#include <cmath>\nextern "C" ' + src[a:b])
    so = tmp_path / "half_up.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", str(so), str(cpp)])
    lib = ctypes.CDLL(str(so))
    lib.half_up.restype = ctypes.c_uint32
    lib.half_up.argtypes = [ctypes.c_float]
    return lib.half_up


def test_half_extents_round_up_to_the_next_fp16(tmp_path):
    half_up = _half_up(tmp_path)
    rng = np.random.default_rng(7)
    vals = np.concatenate([
        np.exp(rng.uniform(-60.0, 12.0, 50000)).astype(np.float32),   # 1e-26 .. 1.6e5: subnormal halves up to overflow
        np.array([0.0, 1e-30, 2.0 ** -25, 2.0 ** -24, 5.96e-8, 6e-8, 6.1e-5, 6.2e-5, 2.0 ** -14, 0.2, 0.5, 1.0, 1.0 + 2.0 ** -11,
                  1.0 + 2.0 ** -10, 1.0005, 65504.0, 65505.0, 1e9], dtype=np.float32)])
    for v in vals:
        h = half_up(float(v))
        assert h <= 0x7c00
        f = float(np.array([h], dtype=np.uint16).view(np.float16)[0])
        if v == 0.0:
            assert h == 0
            continue
        assert f >= float(v), (v, h, f)                      # never below the value (+inf above the fp16 range)
        if h < 0x7c00:
            prev = float(np.array([h - 1], dtype=np.uint16).view(np.float16)[0])
            assert prev < float(v), (v, h, prev)             # and the smallest such fp16
    assert half_up(float("nan")) == 0 and half_up(-1.0) == 0  # not expected (extents of finite boxes): a zero extent
