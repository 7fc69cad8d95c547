//! Raw bindings for include/psim_b200.h plus the safe wrappers that give `src/simulation` the same
//! method surface it uses today (`Quadtree::{build, build_with_domain, field, acc_pos,
//! field_at_point}`, `CellList::{rebuild, find_neighbors_within}`, `forces::*`, `iterate`).
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::c_char;

#[repr(C)]
pub struct psim_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct psim_config {
    pub theta: f32,
    pub epsilon: f32,
    pub leaf_capacity: u32,
    pub thread_capacity: u32,
    pub lj_force_max: f32,
    pub collision_passes: u32,
    pub stack_pressure_enabled: u32,
    pub stack_pressure: f32,
    pub stack_pressure_decay: f32,
    pub parity_mode: u32,
    pub node_factor: f32,
    pub strict_centres: u32,
    pub reserved: [u32; 4],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct psim_species {
    pub mass: f32,
    pub radius: f32,
    pub damping: f32,
    pub lj_epsilon: f32,
    pub lj_sigma: f32,
    pub lj_cutoff: f32,
    pub polar_offset: f32,
    pub polar_charge: f32,
    pub repulsion_strength: f32,
    pub repulsion_cutoff: f32,
    pub lj_enabled: u32,
    pub repulsion_enabled: u32,
}

/// Same field order as `quadtree::Node` (node.rs:6-14), fixed layout.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct psim_node {
    pub children: u64,
    pub next: u64,
    pub pos: [f32; 2],
    pub mass: f32,
    pub quad_center: [f32; 2],
    pub quad_size: f32,
    pub bodies_start: u64,
    pub bodies_end: u64,
    pub charge: f32,
    pub _pad: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct psim_step_params {
    pub hw: f32,
    pub hh: f32,
    pub hd: f32,
    pub dt: f32,
    pub damping_base: f32,
    pub k_e: f32,
    pub bg_x: f32,
    pub bg_y: f32,
    pub density_threshold: f32,
    pub enable_out_of_plane: u32,
    pub do_short_range: u32,
    pub do_electrons: u32,
    pub do_iterate: u32,
    pub do_polar: u32,
    pub reserved: [u32; 2],
}

pub const PSIM_BUILD_CONTAINING: i32 = 0;
pub const PSIM_BUILD_DOMAIN: i32 = 1;
pub const PSIM_SR_LJ: u32 = 1;
pub const PSIM_SR_REPULSION: u32 = 2;
pub const PSIM_SR_STACK_PRESSURE: u32 = 4;

extern "C" {
    pub fn psim_default_config(cfg: *mut psim_config);
    pub fn psim_default_species_table(rows21: *mut psim_species);
    pub fn psim_create(device: i32, max_bodies: u64, max_electrons: u64, cfg: *const psim_config, out: *mut *mut psim_ctx) -> i32;
    pub fn psim_destroy(ctx: *mut psim_ctx) -> i32;
    pub fn psim_last_error(ctx: *const psim_ctx) -> *const c_char;
    pub fn psim_upload_species_table(ctx: *mut psim_ctx, rows: *const psim_species, nrows: u32) -> i32;
    pub fn psim_upload_bodies(ctx: *mut psim_ctx, n: u64, pos_xy: *const f32, z: *const f32, vel_xy: *const f32, vz: *const f32,
                              mass: *const f32, radius: *const f32, charge: *const f32, species: *const u8) -> i32;
    pub fn psim_update_state(ctx: *mut psim_ctx, n: u64, pos_xy: *const f32, vel_xy: *const f32, charge: *const f32) -> i32;
    pub fn psim_upload_electrons(ctx: *mut psim_ctx, m: u64, body: *const u32, rel_xy: *const f32, vel_xy: *const f32) -> i32;
    pub fn psim_download_bodies(ctx: *mut psim_ctx, pos_xy: *mut f32, z: *mut f32, vel_xy: *mut f32, vz: *mut f32, acc_xy: *mut f32,
                                az: *mut f32, mass: *mut f32, radius: *mut f32, charge: *mut f32, species: *mut u8,
                                e_field_xy: *mut f32, orig_index: *mut u32) -> i32;
    pub fn psim_download_electrons(ctx: *mut psim_ctx, body: *mut u32, rel_xy: *mut f32, vel_xy: *mut f32) -> i32;
    pub fn psim_build(ctx: *mut psim_ctx, mode: i32, hw: f32, hh: f32) -> i32;
    pub fn psim_get_permutation(ctx: *mut psim_ctx, out: *mut u32) -> i32;
    pub fn psim_download_nodes(ctx: *mut psim_ctx, out: *mut psim_node, cap: u64, count: *mut u64) -> i32;
    pub fn psim_field(ctx: *mut psim_ctx, k_e: f32, bg_x: f32, bg_y: f32, write_acc: i32, out_e: *mut f32, out_acc: *mut f32) -> i32;
    pub fn psim_acc_points(ctx: *mut psim_ctx, m: u64, pts_xy: *const f32, q: *const f32, radius: *const f32, k_e: f32, out_xy: *mut f32) -> i32;
    pub fn psim_update_electrons(ctx: *mut psim_ctx, bg_x: f32, bg_y: f32, dt: f32, k_e: f32) -> i32;
    pub fn psim_cell_build(ctx: *mut psim_ctx, hw: f32, hh: f32, cell_size: f32) -> i32;
    pub fn psim_neighbors_within(ctx: *mut psim_ctx, m: u64, body_idx: *const u32, cutoff: f32, metals_only: i32, offsets: *mut u32,
                                 indices: *mut u32, indices_cap: u64, total: *mut u64) -> i32;
    pub fn psim_reset_acc(ctx: *mut psim_ctx) -> i32;
    pub fn psim_prepare_spatial_structures(ctx: *mut psim_ctx, hw: f32, hh: f32, density_threshold: f32) -> i32;
    pub fn psim_short_range(ctx: *mut psim_ctx, flags: u32) -> i32;
    pub fn psim_apply_polar_forces(ctx: *mut psim_ctx, k_e: f32, dipole_model: i32) -> i32;
    pub fn psim_iterate(ctx: *mut psim_ctx, dt: f32, damping_base: f32, hw: f32, hh: f32, hd: f32, enable_out_of_plane: i32) -> i32;
    pub fn psim_update_surrounded_flags(ctx: *mut psim_ctx, hw: f32, hh: f32, frame: u64, radius_factor: f32,
                                        neighbor_threshold: u64) -> i32;
    pub fn psim_get_surrounded(ctx: *mut psim_ctx, flags: *mut u8, last_pos_xy: *mut f32, last_frame: *mut u64) -> i32;
    pub fn psim_enforce_metal_z_boundaries(ctx: *mut psim_ctx, max_z: f32, hw: f32, hh: f32) -> i32;
    pub fn psim_shard_init(ctx: *mut psim_ctx, rank: u32, world: u32) -> i32;
    pub fn psim_shard_phase(ctx: *mut psim_ctx, phase: i32, mode: i32, hw: f32, hh: f32, out: *mut u32) -> i32;
    pub fn psim_shard_ptrs(ctx: *mut psim_ctx, out8: *mut u64) -> i32;
    pub fn psim_step(ctx: *mut psim_ctx, p: *const psim_step_params) -> i32;
    pub fn psim_step_host(ctx: *mut psim_ctx, p: *const psim_step_params, n: u64, pos_xy: *const f32,
                          vel_xy: *const f32, charge: *const f32, out_pos_xy: *mut f32, out_vel_xy: *mut f32,
                          out_e_field_xy: *mut f32, out_orig_index: *mut u32) -> i32;
    pub fn psim_sync(ctx: *mut psim_ctx) -> i32;
}

#[derive(Debug)]
pub struct PsimError(pub i32, pub String);

/// Owns one device context; `Send` but not `Sync` (one host thread per context, like the sim thread
/// that owns `Simulation`, src/app/mod.rs:36-41).
pub struct Context(*mut psim_ctx);
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32, max_bodies: usize, max_electrons: usize, theta: f32, epsilon: f32,
               leaf_capacity: usize, thread_capacity: usize) -> Result<Self, PsimError> {
        let mut cfg = unsafe { std::mem::zeroed::<psim_config>() };
        unsafe { psim_default_config(&mut cfg) };
        cfg.theta = theta;
        cfg.epsilon = epsilon;
        cfg.leaf_capacity = leaf_capacity as u32;
        cfg.thread_capacity = thread_capacity as u32;
        let mut h = std::ptr::null_mut();
        let rc = unsafe { psim_create(device, max_bodies as u64, max_electrons as u64, &cfg, &mut h) };
        if rc != 0 {
            return Err(PsimError(rc, "psim_create failed (no CUDA device / out of memory); there is no CPU fallback".into()));
        }
        Ok(Context(h))
    }
    pub fn raw(&self) -> *mut psim_ctx {
        self.0
    }
    pub fn check(&self, rc: i32) -> Result<(), PsimError> {
        if rc == 0 {
            Ok(())
        } else {
            let msg = unsafe { CStr::from_ptr(psim_last_error(self.0)) }.to_string_lossy().into_owned();
            Err(PsimError(rc, msg))
        }
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { psim_destroy(self.0) };
    }
}
