//! FFI declarations of include/mrt.h (ABI version 2).  Field order and types are the header's.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const MRT_OK: c_int = 0;
pub const MRT_SPHERE: u32 = 0;
pub const MRT_PLANE: u32 = 1;
pub const MRT_BOX: u32 = 2;
pub const MRT_TRIANGLE: u32 = 3;
pub const MRT_MESH: u32 = 4;
pub const MRT_LIGHT_POINT: u32 = 0;
pub const MRT_LIGHT_DIR: u32 = 1;

#[repr(C)] #[derive(Clone, Copy)]
pub struct mrt_material {
    pub albedo: [f32; 3],
    pub rough: f32, pub metal: f32, pub glass: f32, pub opacity: f32, pub emit: f32,
    pub tex: i32, pub rmap: i32, pub mmap: i32, pub gmap: i32, pub omap: i32, pub emap: i32,
}
#[repr(C)] #[derive(Clone, Copy)]
pub struct mrt_object {
    pub kind: u32, pub mesh: u32, pub param: [f32; 9],
    pub first_inst: u32, pub n_inst: u32, pub mat: mrt_material,
}
#[repr(C)] #[derive(Clone, Copy)]
pub struct mrt_instance { pub pos: [f32; 3], pub dir: [f32; 4] }          // dir = (w, x, y, z), lin.rs:428-443
#[repr(C)] #[derive(Clone, Copy)]
pub struct mrt_texture { pub w: u32, pub h: u32, pub first_texel: u64, pub has_dat: u32, pub _pad: u32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct mrt_mesh { pub first_tri: u32, pub n_tri: u32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct mrt_light { pub kind: u32, pub v: [f32; 3], pub pwr: f32, pub color: [f32; 3] }
#[repr(C)]
pub struct mrt_scene {
    pub objects: *const mrt_object, pub n_objects: u32,
    pub instances: *const mrt_instance, pub n_instances: u32,
    pub textures: *const mrt_texture, pub n_textures: u32,
    pub texels: *const f32, pub n_texels: u64,
    pub meshes: *const mrt_mesh, pub n_meshes: u32,
    pub triangles: *const f32, pub n_triangles: u32,
    pub lights: *const mrt_light, pub n_lights: u32,
    pub sky_color: [f32; 3], pub sky_pwr: f32,
}
#[repr(C)] #[derive(Clone, Copy, PartialEq)]
pub struct mrt_frame {
    pub res: [u16; 2], pub ssaa: f32, pub cam_pos: [f32; 3], pub cam_dir: [f32; 4],
    pub fov: f32, pub gamma: f32, pub exp: f32, pub aprt: f32, pub foc: f32,
}
pub enum mrt_ctx {}

extern "C" {
    pub fn mrt_abi_version() -> c_int;
    pub fn mrt_create(out: *mut *mut mrt_ctx, device: c_int, workers: u32, n_dim: u32) -> c_int;
    pub fn mrt_create_group(out: *mut *mut mrt_ctx, devices: *const c_int, n_devices: c_int, workers: u32, n_dim: u32) -> c_int;
    pub fn mrt_update_scene(ctx: *mut mrt_ctx, scene: *const mrt_scene) -> c_int;
    pub fn mrt_update_frame(ctx: *mut mrt_ctx, frame: *const mrt_frame) -> c_int;
    pub fn mrt_sync(ctx: *mut mrt_ctx) -> c_int;
    pub fn mrt_device_seconds(ctx: *mut mrt_ctx, total: *mut f64) -> c_int;
    pub fn mrt_destroy(ctx: *mut mrt_ctx);
    pub fn mrt_last_error(ctx: *const mrt_ctx) -> *const c_char;
    pub fn mrt_set_scene(ctx: *mut mrt_ctx, scene: *const mrt_scene) -> c_int;
    pub fn mrt_set_frame(ctx: *mut mrt_ctx, frame: *const mrt_frame) -> c_int;
    pub fn mrt_set_rt(ctx: *mut mrt_ctx, bounce: u32, loss: f32, seed: u64) -> c_int;
    pub fn mrt_set_option(ctx: *mut mrt_ctx, option: u32, value: u32) -> c_int;
    pub fn mrt_execute(ctx: *mut mrt_ctx, n_passes: u32, seconds: *mut f64) -> c_int;
    pub fn mrt_reset(ctx: *mut mrt_ctx) -> c_int;
    pub fn mrt_img(ctx: *mut mrt_ctx, rgb: *mut u8) -> c_int;
    pub fn mrt_accum_device(ctx: *mut mrt_ctx, dptr: *mut *mut c_void, n_floats: *mut usize, stream: *mut *mut c_void) -> c_int;
}
