#!/usr/bin/env python
"""Writes tests/golden/vectors/<scene>.npz: what the CPU oracle computes for a small render of every
example scene — per-ray primary hits (deterministic) and the 2-pass linear accumulator with the
shared counter-based random numbers (seed 0x5EED).  Committed so that (a) a change in the oracle's
behaviour is caught without a GPU (tests/test_golden_vectors.py recomputes and compares), and (b)
the CUDA path is checked against vectors that existed before the kernel under test was built.

Regenerate only when the oracle is changed on purpose (e.g. the RNG mapping), and say so in the commit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_lib  # noqa: E402
from util import load  # noqa: E402

CASES = [("Default", (64, 36), 1.0), ("CornellBox2", (40, 40), 2.0), ("CornellBox", (64, 36), 1.0), ("dof", (64, 36), 1.0),
         ("Minecraft", (48, 27), 2.0), ("Mesh", (64, 36), 1.0), ("Instance", (48, 27), 1.0)]
PASSES = 2


def compute(name, res, ssaa):
    r = load(name, res, ssaa)
    s = oracle_lib.OracleSampler(workers=1)
    s.execute(r.scene, r.frame, r.rt, PASSES)
    h = s.trace_primary()
    acc, n = s.accum()
    assert n == PASSES
    return {"t0": h["t0"], "t1": h["t1"], "obj": h["obj"].astype(np.int16), "inst": h["inst"].astype(np.int16),
            "tri0": h["tri0"].astype(np.int16), "n0": h["n0"], "uv": h["uv"], "accum": acc}


def main():
    out = os.path.join(HERE, "vectors")
    os.makedirs(out, exist_ok=True)
    for name, res, ssaa in CASES:
        np.savez_compressed(os.path.join(out, name + ".npz"), **compute(name, res, ssaa))
        print(name, os.path.getsize(os.path.join(out, name + ".npz")))


if __name__ == "__main__":
    main()
