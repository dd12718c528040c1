"""Compile the run-time-specialised kernel with NVRTC from the in-tree sources (no GPU needed) and
optionally dump its SASS: python tools/nvrtc_check.py [scene_header.h] [out.cubin]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "micro_raytracer_b200", "csrc")
DEFAULT = """
#define MRT_JIT_N_BOX 3
#define MRT_JIT_N_SPHERE 1
#define MRT_JIT_N_PLANE 1
#define MRT_JIT_BOXPAIRS(X, XS, X1, CB, CE) CB(-0x1p+2f,-0x1p+2f,-0x1p+2f,0x1p+2f,0x1p+2f,0x1p+2f) XS(1, 0x1p+0f,0x1p+1f,0x0p+0f,0x1p+0f,0x1p+0f,0x0p+0f,0x1p-1f,0x1p-2f,0x1p-1f,0x1p-1f,0x1p-1f,0x1p-3f) X1(2, 0x1p+0f,0x1p+0f,0x1p+0f,0x1p+0f,0x1p+0f,0x1p+0f,0x1p-1f,-0x1p+0f,0x1p-1f,-0x1p+0f,0x1p-1f,-0x1p+0f) CE X(0, 0x1p+0f,0x1p+0f,0x1p+0f,0x1p+0f,0x1p+0f,0x1p+0f,0x1p-1f,0x1p-1f,0x1p-1f,0x1p-1f,0x1p-1f,-0x1p+0f)
#define MRT_JIT_SPHERES(X) X(0, 0x1p+0f, 0x1p+1f, 0x1p+0f, 0x1p-2f)
#define MRT_JIT_PLANES(X) X(0, 0x0p+0f, 0x0p+0f, 0x1p+0f, -0x1p+0f)
#define MRT_JIT_BXFS(X) X(0, 0x1p-1f,0x1p-1f,0x0p+0f,0x1p+0f, -0x1p-1f,0x1p-1f,0x0p+0f,0x1p+0f, 0x0p+0f,0x0p+0f,0x1p+0f,0x1p+0f, 0x1p-2f,0x1p-2f,0x1p-2f)
#define MRT_JIT_MESHES(X) X(0, 0x0p+0f,0x0p+0f,0x0p+0f, 1u, 0u, 1.f,0.f,0.f,0.f, 0.f,1.f,0.f,0.f, 0.f,0.f,1.f,0.f)
#define MRT_JIT_FIRST_SPHERE 2
#define MRT_JIT_FIRST_PLANE 3
#define MRT_JIT_FIRST_BXF 4
#define MRT_JIT_FIRST_MESH 5
#define MRT_JIT_F %du
"""
def _main_src():
    """k_main_src of csrc/mrt_jit.cu (the translation unit NVRTC compiles), read from the source so the two cannot drift."""
    import re
    t = open(os.path.join(src, "mrt_jit.cu")).read()
    body = t[t.index("const char* k_main_src ="):]
    body = body[:body.index(";\n")]
    return "".join(bytes(m, "utf-8").decode("unicode_escape") for m in re.findall(r'"((?:[^"\\]|\\.)*)"', body)).encode()


MAIN = _main_src()

def compile_header(header: bytes, out=None, extra=()):
    n = C.CDLL("libnvrtc.so.12")
    prog = C.c_void_p()
    hs = [header, open(os.path.join(src, "mrt_path.cuh"), "rb").read(), open(os.path.join(src, "mrt_device.cuh"), "rb").read()]
    names = [b"mrt_jit_scene.h", b"mrt_path.cuh", b"mrt_device.cuh"]
    arr = (C.c_char_p * 3)(*hs); narr = (C.c_char_p * 3)(*names)
    assert n.nvrtcCreateProgram(C.byref(prog), MAIN, b"mrt_jit_main.cu", 3, arr, narr) == 0
    opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"-ftz=true", b"-prec-div=false", b"-prec-sqrt=false", b"-lineinfo", b"-DMRT_JIT=1", *[e.encode() for e in extra]]
    rc = n.nvrtcCompileProgram(prog, len(opts), (C.c_char_p * len(opts))(*opts))
    ls = C.c_size_t(); n.nvrtcGetProgramLogSize(prog, C.byref(ls))
    log = C.create_string_buffer(ls.value or 1); n.nvrtcGetProgramLog(prog, log)
    if rc != 0:
        raise RuntimeError("NVRTC failed:\n" + log.value.decode()[:6000])
    cs = C.c_size_t(); n.nvrtcGetCUBINSize(prog, C.byref(cs))
    cubin = C.create_string_buffer(cs.value); n.nvrtcGetCUBIN(prog, cubin)
    if out:
        open(out, "wb").write(cubin.raw)
    return cs.value, log.value.decode()

if __name__ == "__main__":
    import time
    if len(sys.argv) > 1 and os.path.exists(sys.argv[1]):
        t = time.time(); size, log = compile_header(open(sys.argv[1], "rb").read(), sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3:])
        print("ok", size, "bytes", round(time.time() - t, 2), "s", log[:500])
    else:
        for f in (0, 15):
            t = time.time(); size, log = compile_header((DEFAULT % f).encode(), "/tmp/jit_f%d.cubin" % f)
            print("F", f, "ok", size, "bytes", round(time.time() - t, 2), "s", log[:500])
