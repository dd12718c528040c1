"""micro-raytracer per-pixel path-tracing hot path, B200-native.

Host-side mirror of the reference's operator interface for this path:
`Sampler::{new, execute, img}` (src/sampler.rs) over the C ABI of include/mrt.h, plus the
description types and JSON loader that feed it.  All pixels are computed by the CUDA
library libmrt.so (csrc/); nothing in this package falls back to a CPU implementation.
"""
from .scene import (Camera, Frame, Light, Material, RayTracer, Render, Renderer, Scene, SceneError, Sky,
                    Texture, load_render, pack_scene, render_from_dict)
from .sampler import MrtError, Sampler, lib_path, load_library

__all__ = ["Camera", "Frame", "Light", "Material", "RayTracer", "Render", "Renderer", "Scene", "SceneError",
           "Sky", "Texture", "load_render", "pack_scene", "render_from_dict", "MrtError", "Sampler",
           "lib_path", "load_library"]
