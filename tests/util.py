"""Shared helpers for the parity tests."""
import os

import numpy as np

import micro_raytracer_b200 as mrt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = os.path.join(ROOT, "tests", "golden", "scenes")
REF_RENDERS = os.path.join(ROOT, "tests", "golden", "ref_renders")


def load(name, res=None, ssaa=None, **rt):
    r = mrt.load_render(os.path.join(SCENES, name + ".json"))
    if res is not None:
        r.frame.res = tuple(res)
    if ssaa is not None:
        r.frame.ssaa = float(ssaa)
    for k, v in rt.items():
        setattr(r.rt, k, v)
    return r


def png(name):
    from PIL import Image
    return np.asarray(Image.open(os.path.join(REF_RENDERS, name)).convert("RGB"))


def block_mean(a, b=8):
    h, w = a.shape[0] // b * b, a.shape[1] // b * b
    a = a[:h, :w].astype(np.float64)
    return a.reshape(h // b, b, w // b, b, -1).mean(axis=(1, 3))


def tonemap_f(v, gamma, exp):
    """sampler.rs:85-94 on float values, scaled to 0..255: `as u8` saturates at 255 and truncates,
    which lowers the mean of a noisy image by 0.5."""
    g = np.power(np.maximum(v, 0.0), gamma)
    t = np.minimum(255.0 * g * (1.0 + g / (1.0 - exp) ** 2) / (1.0 + g), 255.0)
    return np.maximum(t - 0.5, 0.0)


def block_tonemapped(sampler, r, b=8):
    """Block means of the LINEAR accumulator, then tone-mapped: at a fraction of the reference's
    spp this avoids the Jensen darkening a noisy image suffers through the concave tone map, so
    it compares directly with block means of the reference's (converged) PNG."""
    acc, n = sampler.accum()
    return tonemap_f(block_mean(acc / n, b), r.frame.cam.gamma, r.frame.cam.exp)


def psnr(a, b, peak=255.0):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False
