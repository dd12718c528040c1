"""Dump the generated scene header and the specialised cubin of every example scene (needs a GPU):
    python tools/dump_jit.py OUT_DIR
The headers feed tools/nvrtc_check.py, which recompiles them WITHOUT a GPU (SASS / register studies)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "jit")
os.makedirs(out, exist_ok=True)
code = """
import sys
sys.path[:0] = [%r, %r]
import micro_raytracer_b200 as mrt
from util import load
r = load(sys.argv[1], (96, 54), 1.0)
s = mrt.Sampler(device=0)
s.set_option(2, 2)
s.execute(r.scene, r.frame, r.rt, 2)  # a batched call: launches (and, with MRT_JIT_FORCE, compiles) at once
print(sys.argv[1], s.jit_status())
""" % (ROOT, os.path.join(ROOT, "tests"))
for name in ["Default", "CornellBox", "CornellBox2", "dof", "Mesh", "Minecraft", "Instance"]:
    env = dict(os.environ, MRT_JIT_CACHE="off", MRT_JIT_DUMP_HEADER=os.path.join(out, name + ".h"), MRT_JIT_DUMP=os.path.join(out, name + ".cubin"))
    subprocess.run([sys.executable, "-c", code, name], env=env, check=False)
