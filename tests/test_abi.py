"""The C-ABI library loads and exports every symbol include/mrt.h declares; host-side logic of
the Python mirror (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

import micro_raytracer_b200 as mrt
from micro_raytracer_b200 import abi
from util import ROOT, load


def _declared(header):
    text = open(os.path.join(ROOT, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrt_(?:cpu_)?[a-z0-9_]+)\s*\(", text)))


def test_libmrt_exports_every_declared_symbol():
    lib = ctypes.CDLL(mrt.lib_path())
    names = _declared("include/mrt.h")
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libmrt.so does not export {n}"
    lib.mrt_abi_version.restype = ctypes.c_int
    assert lib.mrt_abi_version() == abi.MRT_ABI_VERSION == 2
    assert set(abi.MRT_SYMBOLS) == set(names)


def test_oracle_exports_every_declared_symbol():
    import oracle_lib
    lib = ctypes.CDLL(oracle_lib.build_oracle())
    for n in _declared("oracle/mrt_oracle.h"):
        assert hasattr(lib, n), f"libmrt_oracle.so does not export {n}"


def test_struct_sizes_match_the_header():
    # sizes the C compiler gives the structs of include/mrt.h (LP64)
    assert ctypes.sizeof(abi.MrtHit) == 80
    assert ctypes.sizeof(abi.MrtFrame) == 56
    assert ctypes.sizeof(abi.MrtScene) == 128


def test_no_device_is_a_loud_error_not_a_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(mrt.MrtError) as e:
        mrt.Sampler(device=0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_scene_loader_defaults_match_parser_rs():
    """parser.rs:188-271 defaults and the instance rules of parser.rs:838-853."""
    r = mrt.render_from_dict({"scene": {"renderer": [{"type": "sphere", "r": 0.5}]}})
    assert (r.rt.bounce, r.rt.sample, abs(r.rt.loss - 0.15) < 1e-7) == (8, 16, True)
    assert tuple(r.frame.res) == (1280, 720) and r.frame.ssaa == 1.0
    c = r.frame.cam
    assert tuple(c.pos) == (0.0, -1.0, 0.0) and tuple(c.dir) == (0.0, 0.0, 1.0, 0.0)
    assert (c.fov, c.gamma, c.exp, c.aprt, c.foc) == (70.0, 0.8, 0.2, 0.001, 100.0)
    packed = mrt.pack_scene(r.scene)
    assert packed.c.n_objects == 1 and packed.c.n_instances == 1
    r = load("Instance")
    assert mrt.pack_scene(r.scene).c.n_instances == 1000
    r = load("Minecraft")
    p = mrt.pack_scene(r.scene)
    assert p.c.n_instances == 85 and p.c.n_textures >= 9


def test_header_is_valid_c_and_links_from_c(tmp_path):
    """include/mrt.h is a C header (the Rust/cgo/ctypes side binds it as C): a C translation unit
    including it compiles with gcc and links against libmrt.so; without a GPU mrt_create must fail
    with MRT_ERR_CUDA and a message, never abort."""
    import subprocess
    src = tmp_path / "abi_c.c"
    src.write_text('''
#include <stdio.h>
#include "mrt.h"
int main(void) {
    mrt_ctx* ctx = NULL;
    mrt_scene sc; mrt_frame fr; mrt_hit h; mrt_material m;
    (void)sc; (void)fr; (void)h; (void)m;
    if (mrt_abi_version() != MRT_ABI_VERSION) return 10;
    int rc = mrt_create(&ctx, 0, 24, 64);
    if (rc == MRT_OK) { mrt_destroy(ctx); printf("created\\n"); return 0; }
    printf("rc=%d msg=%s\\n", rc, mrt_last_error(NULL));
    return rc == MRT_ERR_CUDA ? 0 : 11;
}
''')
    exe = tmp_path / "abi_c"
    libdir = os.path.dirname(mrt.lib_path())
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libmrt.so", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "created" in out.stdout or "no CUDA device" in out.stdout or "CUDA" in out.stdout
