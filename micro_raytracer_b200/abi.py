"""ctypes mirror of include/mrt.h (the C-ABI boundary structs).

Shared by the product binding (sampler.py) and, in tests/, by the oracle binding; it holds
layouts only — no library is loaded here.
"""
import ctypes as C

MRT_ABI_VERSION = 2
MRT_OK, MRT_ERR_INVALID, MRT_ERR_CUDA, MRT_ERR_STATE, MRT_ERR_NOMEM = range(5)
MRT_SPHERE, MRT_PLANE, MRT_BOX, MRT_TRIANGLE, MRT_MESH = range(5)
MRT_LIGHT_POINT, MRT_LIGHT_DIR = range(2)


class MrtMaterial(C.Structure):
    _fields_ = [
        ("albedo", C.c_float * 3),
        ("rough", C.c_float), ("metal", C.c_float), ("glass", C.c_float),
        ("opacity", C.c_float), ("emit", C.c_float),
        ("tex", C.c_int32), ("rmap", C.c_int32), ("mmap", C.c_int32),
        ("gmap", C.c_int32), ("omap", C.c_int32), ("emap", C.c_int32),
    ]


class MrtObject(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("mesh", C.c_uint32),
        ("param", C.c_float * 9),
        ("first_inst", C.c_uint32), ("n_inst", C.c_uint32),
        ("mat", MrtMaterial),
    ]


class MrtInstance(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("dir", C.c_float * 4)]


class MrtTexture(C.Structure):
    _fields_ = [("w", C.c_uint32), ("h", C.c_uint32), ("first_texel", C.c_uint64),
                ("has_dat", C.c_uint32), ("_pad", C.c_uint32)]


class MrtMesh(C.Structure):
    _fields_ = [("first_tri", C.c_uint32), ("n_tri", C.c_uint32)]


class MrtLight(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("v", C.c_float * 3), ("pwr", C.c_float), ("color", C.c_float * 3)]


class MrtScene(C.Structure):
    _fields_ = [
        ("objects", C.POINTER(MrtObject)), ("n_objects", C.c_uint32),
        ("instances", C.POINTER(MrtInstance)), ("n_instances", C.c_uint32),
        ("textures", C.POINTER(MrtTexture)), ("n_textures", C.c_uint32),
        ("texels", C.POINTER(C.c_float)), ("n_texels", C.c_uint64),
        ("meshes", C.POINTER(MrtMesh)), ("n_meshes", C.c_uint32),
        ("triangles", C.POINTER(C.c_float)), ("n_triangles", C.c_uint32),
        ("lights", C.POINTER(MrtLight)), ("n_lights", C.c_uint32),
        ("sky_color", C.c_float * 3), ("sky_pwr", C.c_float),
    ]


class MrtFrame(C.Structure):
    _fields_ = [
        ("res", C.c_uint16 * 2), ("ssaa", C.c_float),
        ("cam_pos", C.c_float * 3), ("cam_dir", C.c_float * 4),
        ("fov", C.c_float), ("gamma", C.c_float), ("exp", C.c_float),
        ("aprt", C.c_float), ("foc", C.c_float),
    ]


class MrtHit(C.Structure):
    _fields_ = [
        ("t0", C.c_float), ("t1", C.c_float),
        ("obj", C.c_int32), ("inst", C.c_int32), ("tri0", C.c_int32), ("tri1", C.c_int32),
        ("n0", C.c_float * 3), ("n1", C.c_float * 3), ("uv", C.c_float * 2),
        ("orig", C.c_float * 3), ("dir", C.c_float * 3),
    ]


# numpy view of MrtHit (80 bytes, all 4-byte fields)
HIT_DTYPE = [
    ("t0", "<f4"), ("t1", "<f4"), ("obj", "<i4"), ("inst", "<i4"), ("tri0", "<i4"), ("tri1", "<i4"),
    ("n0", "<f4", (3,)), ("n1", "<f4", (3,)), ("uv", "<f4", (2,)), ("orig", "<f4", (3,)), ("dir", "<f4", (3,)),
]

# Every symbol include/mrt.h declares (tests check the built library exports all of them).
MRT_SYMBOLS = [
    "mrt_create", "mrt_create_group", "mrt_group_info", "mrt_device_count", "mrt_destroy", "mrt_last_error", "mrt_abi_version",
    "mrt_set_scene", "mrt_set_frame", "mrt_update_scene", "mrt_update_frame", "mrt_set_rt", "mrt_set_option", "mrt_set_partition",
    "mrt_device_seconds", "mrt_spp_per_launch", "mrt_jit_status", "mrt_scene_info",
    "mrt_ipc_export", "mrt_ipc_attach", "mrt_ipc_tonemap_band", "mrt_img_gathered",
    "mrt_execute", "mrt_execute_async", "mrt_sync", "mrt_reset", "mrt_film_size",
    "mrt_accum", "mrt_accum_device", "mrt_set_passes", "mrt_set_stream", "mrt_img", "mrt_img_ss",
    "mrt_trace_primary", "mrt_launch_count", "mrt_fp32_peak",
]
