//! Bindings for include/psim_b200.h (libpsim_b200.so, the B200 implementation of ParticleSim's force hot path).
//!
//! * `ffi` — the raw `extern "C"` block, GENERATED from the header (tools/gen_rust_ffi.py), every symbol bound.
//! * `Context` — owns one device context (one per GPU, one host thread per context).
//! * `Quadtree`, `CellList`, `forces::*`, `Simulation` — safe wrappers with the method names `src/simulation` calls
//!   today (`src/quadtree/quadtree.rs`, `src/cell_list.rs`, `src/simulation/forces.rs`, `simulation.rs:1437-1486`),
//!   working on a struct-of-arrays mirror of `Vec<Body>` (`Bodies`).  The reference's methods take `&mut [Body]`; a
//!   maintainer converts at the call site or keeps `Bodies` as the source of truth.
//!
//! This crate cannot be compiled in the repository's development image (no cargo / rustc there); it is checked in
//! as the reference-side binding and kept free of dependencies so that it can be reviewed by eye.
#![allow(non_camel_case_types)]
use std::ffi::CStr;

pub mod ffi;
pub use ffi::*;

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
#[derive(Clone, Copy, Default)]
pub struct psim_stats {
    pub n_bodies: u64,
    pub n_electrons: u64,
    pub compact_nodes: u64,
    pub reference_nodes: u64,
    pub max_depth: u32,
    pub depth_cap: u32,
    pub zero_leaves: u32,
    pub cap_leaves: u32,
    pub root_center: [f32; 2],
    pub root_size: f32,
    pub grid_x: u32,
    pub grid_y: u32,
    pub traversal_warp_steps: u64,
    pub kernel_launches: u64,
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

pub const PSIM_OK: i32 = 0;
pub const PSIM_E_CUDA: i32 = -1;
pub const PSIM_E_ARG: i32 = -2;
pub const PSIM_E_OOM: i32 = -3;
pub const PSIM_E_NODE_OVERFLOW: i32 = -4;
pub const PSIM_E_STATE: i32 = -5;
pub const PSIM_E_NCCL: i32 = -6;
pub const PSIM_BUILD_CONTAINING: i32 = 0;
pub const PSIM_BUILD_DOMAIN: i32 = 1;
pub const PSIM_SR_LJ: u32 = 1;
pub const PSIM_SR_REPULSION: u32 = 2;
pub const PSIM_SR_STACK_PRESSURE: u32 = 4;

#[derive(Debug)]
pub struct PsimError(pub i32, pub String);
pub type Result<T> = std::result::Result<T, PsimError>;

/// Owns one device context; `Send` but not `Sync` (one host thread per context, like the sim thread that owns
/// `Simulation`, src/app/mod.rs:36-41).
pub struct Context(*mut psim_ctx);
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32, max_bodies: usize, max_electrons: usize, theta: f32, epsilon: f32, leaf_capacity: usize,
               thread_capacity: usize) -> Result<Self> {
        let mut cfg = unsafe { std::mem::zeroed::<psim_config>() };
        unsafe { psim_default_config(&mut cfg) };
        cfg.theta = theta;
        cfg.epsilon = epsilon;
        cfg.leaf_capacity = leaf_capacity as u32;
        cfg.thread_capacity = thread_capacity as u32;
        Self::with_config(device, max_bodies, max_electrons, &cfg)
    }
    pub fn with_config(device: i32, max_bodies: usize, max_electrons: usize, cfg: &psim_config) -> Result<Self> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { psim_create(device, max_bodies as u64, max_electrons as u64, cfg, &mut h) };
        if rc != 0 {
            return Err(PsimError(rc, "psim_create failed (no CUDA device / out of memory); there is no CPU fallback".into()));
        }
        Ok(Context(h))
    }
    pub fn raw(&self) -> *mut psim_ctx {
        self.0
    }
    pub fn check(&self, rc: i32) -> Result<()> {
        if rc == 0 {
            Ok(())
        } else {
            let msg = unsafe { CStr::from_ptr(psim_last_error(self.0)) }.to_string_lossy().into_owned();
            Err(PsimError(rc, msg))
        }
    }
    pub fn stats(&self) -> Result<psim_stats> {
        let mut st = psim_stats::default();
        self.check(unsafe { psim_stats_get(self.0, &mut st) })?;
        Ok(st)
    }
    /// how the last build made the node charges: `[integer prefix path (1) or level sweeps (0), charged bodies,
    /// sum |q| of the integer charges, non-integer charges]` (`psim_build_info`)
    pub fn build_info(&self) -> Result<[u64; 4]> {
        let mut out = [0u64; 4];
        self.check(unsafe { psim_build_info(self.0, out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// NCCL communicator for the sharded hot path (`id` from `unique_id()` on rank 0, distributed by the host).
    pub fn comm_init(&self, id: &[u8; 128], rank: u32, nranks: u32) -> Result<()> {
        self.check(unsafe { psim_comm_init(self.0, id.as_ptr(), rank, nranks) })
    }
    pub fn unique_id() -> Result<[u8; 128]> {
        let mut id = [0u8; 128];
        match unsafe { psim_comm_unique_id(id.as_mut_ptr()) } {
            0 => Ok(id),
            rc => Err(PsimError(rc, "NCCL not available".into())),
        }
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { psim_destroy(self.0) };
    }
}

fn opt(v: &[f32]) -> *const f32 {
    if v.is_empty() { std::ptr::null() } else { v.as_ptr() }
}

/// `Vec<Body>` as a struct of arrays: the hot fields of body/types.rs:38-62 plus the flattened electrons
/// (body/electron.rs:9-13).  Row order is the device's body order (a build permutes it, like the reference).
#[derive(Default, Clone)]
pub struct Bodies {
    pub pos: Vec<[f32; 2]>,
    pub z: Vec<f32>,
    pub vel: Vec<[f32; 2]>,
    pub vz: Vec<f32>,
    pub acc: Vec<[f32; 2]>,
    pub az: Vec<f32>,
    pub mass: Vec<f32>,
    pub radius: Vec<f32>,
    pub charge: Vec<f32>,
    pub species: Vec<u8>,
    pub e_field: Vec<[f32; 2]>,
    /// index the body had when it was uploaded (`Body::id` stand-in)
    pub id: Vec<u32>,
    pub electron_body: Vec<u32>,
    pub electron_rel_pos: Vec<[f32; 2]>,
    pub electron_vel: Vec<[f32; 2]>,
}

impl Bodies {
    pub fn len(&self) -> usize {
        self.pos.len()
    }
    pub fn is_empty(&self) -> bool {
        self.pos.is_empty()
    }
    fn flat(v: &[[f32; 2]]) -> *const f32 {
        if v.is_empty() { std::ptr::null() } else { v.as_ptr() as *const f32 }
    }
    fn flat_mut(v: &mut Vec<[f32; 2]>, n: usize) -> *mut f32 {
        v.resize(n, [0.0; 2]);
        v.as_mut_ptr() as *mut f32
    }
    fn permute(&mut self, perm: &[u32]) {
        fn apply<T: Clone>(v: &mut Vec<T>, perm: &[u32]) {
            if v.len() == perm.len() {
                *v = perm.iter().map(|&k| v[k as usize].clone()).collect();
            }
        }
        apply(&mut self.pos, perm);
        apply(&mut self.z, perm);
        apply(&mut self.vel, perm);
        apply(&mut self.vz, perm);
        apply(&mut self.acc, perm);
        apply(&mut self.az, perm);
        apply(&mut self.mass, perm);
        apply(&mut self.radius, perm);
        apply(&mut self.charge, perm);
        apply(&mut self.species, perm);
        apply(&mut self.e_field, perm);
        apply(&mut self.id, perm);
    }
}

/// `Quadtree` of src/quadtree/quadtree.rs:11-34 with the methods `Simulation` calls.
pub struct Quadtree<'c> {
    ctx: &'c Context,
}

impl<'c> Quadtree<'c> {
    pub const ROOT: usize = 0;
    /// `Quadtree::build` (quadtree.rs:153-170): permutes `bodies` like the reference's in-place partition.
    pub fn build(&mut self, bodies: &mut Bodies) -> Result<()> {
        self.build_mode(bodies, PSIM_BUILD_CONTAINING, 0.0, 0.0)
    }
    /// `Quadtree::build_with_domain` (quadtree.rs:173-195).
    pub fn build_with_domain(&mut self, bodies: &mut Bodies, domain_width: f32, domain_height: f32) -> Result<()> {
        self.build_mode(bodies, PSIM_BUILD_DOMAIN, domain_width, domain_height)
    }
    fn build_mode(&mut self, bodies: &mut Bodies, mode: i32, hw: f32, hh: f32) -> Result<()> {
        let c = self.ctx;
        c.check(unsafe { psim_build(c.raw(), mode, hw, hh) })?;
        let mut perm = vec![0u32; bodies.len()];
        c.check(unsafe { psim_get_permutation(c.raw(), perm.as_mut_ptr()) })?;
        bodies.permute(&perm);
        // electrons follow their bodies on the device; refresh the host's flattened copy
        let m = bodies.electron_body.len();
        if m > 0 {
            let rel = Bodies::flat_mut(&mut bodies.electron_rel_pos, m);
            let vel = Bodies::flat_mut(&mut bodies.electron_vel, m);
            c.check(unsafe { psim_download_electrons(c.raw(), bodies.electron_body.as_mut_ptr(), rel, vel) })?;
        }
        Ok(())
    }
    /// `Quadtree::field` (quadtree.rs:418-427): `e_field[i] = acc_pos(pos_i, 1, radius_i)`.
    pub fn field(&self, bodies: &mut Bodies, k_e: f32) -> Result<()> {
        let n = bodies.len();
        let out = Bodies::flat_mut(&mut bodies.e_field, n);
        self.ctx.check(unsafe { psim_field(self.ctx.raw(), k_e, 0.0, 0.0, 0, out, std::ptr::null_mut()) })
    }
    /// `Quadtree::acc_pos` (quadtree.rs:350-407) for a batch of points; `q` / `radius` empty => 1 / 0.
    pub fn acc_pos(&self, points: &[[f32; 2]], q: &[f32], radius: &[f32], k_e: f32) -> Result<Vec<[f32; 2]>> {
        let mut out = vec![[0.0f32; 2]; points.len()];
        self.ctx.check(unsafe {
            psim_acc_points(self.ctx.raw(), points.len() as u64, Bodies::flat(points), opt(q), opt(radius), k_e,
                            out.as_mut_ptr() as *mut f32)
        })?;
        Ok(out)
    }
    /// `Quadtree::field_at_point` (quadtree.rs:504-507), batched.
    pub fn field_at_point(&self, points: &[[f32; 2]], k_e: f32) -> Result<Vec<[f32; 2]>> {
        self.acc_pos(points, &[], &[], k_e)
    }
    /// `quadtree.nodes` in the reference's layout (renderer / diagnostics consumers).
    pub fn nodes(&self) -> Result<Vec<psim_node>> {
        let mut count = 0u64;
        self.ctx.check(unsafe { psim_download_nodes(self.ctx.raw(), std::ptr::null_mut(), 0, &mut count) })?;
        let mut v = vec![unsafe { std::mem::zeroed::<psim_node>() }; count as usize];
        self.ctx.check(unsafe { psim_download_nodes(self.ctx.raw(), v.as_mut_ptr(), count, &mut count) })?;
        Ok(v)
    }
}

/// `CellList` of src/cell_list.rs.
pub struct CellList<'c> {
    ctx: &'c Context,
    pub domain_width: f32,
    pub domain_height: f32,
    pub cell_size: f32,
}

impl<'c> CellList<'c> {
    pub fn update_domain_size(&mut self, domain_width: f32, domain_height: f32) {
        self.domain_width = domain_width;
        self.domain_height = domain_height;
    }
    /// `CellList::rebuild` (cell_list.rs:27-39)
    pub fn rebuild(&mut self, _bodies: &Bodies) -> Result<()> {
        self.ctx.check(unsafe { psim_cell_build(self.ctx.raw(), self.domain_width, self.domain_height, self.cell_size) })
    }
    fn query(&self, idx: &[u32], cutoff: f32, metals_only: bool) -> Result<(Vec<u32>, Vec<u32>)> {
        let c = self.ctx;
        let mut offsets = vec![0u32; idx.len() + 1];
        let mut total = 0u64;
        c.check(unsafe {
            psim_neighbors_within(c.raw(), idx.len() as u64, idx.as_ptr(), cutoff, metals_only as i32, offsets.as_mut_ptr(),
                                  std::ptr::null_mut(), 0, &mut total)
        })?;
        let mut indices = vec![0u32; total as usize];
        if total > 0 {
            c.check(unsafe {
                psim_neighbors_within(c.raw(), idx.len() as u64, idx.as_ptr(), cutoff, metals_only as i32,
                                      offsets.as_mut_ptr(), indices.as_mut_ptr(), total, &mut total)
            })?;
        }
        Ok((offsets, indices))
    }
    /// `CellList::find_neighbors_within` (cell_list.rs:57-85), same order as the reference.
    pub fn find_neighbors_within(&self, _bodies: &Bodies, i: usize, cutoff: f32) -> Result<Vec<usize>> {
        let (_, ind) = self.query(&[i as u32], cutoff, false)?;
        Ok(ind.into_iter().map(|k| k as usize).collect())
    }
    /// the same for many bodies at once: CSR (offsets, indices)
    pub fn find_neighbors_within_batch(&self, idx: &[u32], cutoff: f32) -> Result<(Vec<u32>, Vec<u32>)> {
        self.query(idx, cutoff, false)
    }
    /// `CellList::metal_neighbor_count` (cell_list.rs:92-127)
    pub fn metal_neighbor_count(&self, _bodies: &Bodies, i: usize, cutoff: f32) -> Result<usize> {
        Ok(self.query(&[i as u32], cutoff, true)?.1.len())
    }
}

/// The slice of `Simulation` the hot path touches (simulation.rs:83-139).
pub struct Simulation {
    pub ctx: Context,
    pub bodies: Bodies,
    pub domain_width: f32,
    pub domain_height: f32,
    pub domain_depth: f32,
    pub dt: f32,
    pub coulomb_constant: f32,
    pub damping_base: f32,
    pub cell_list_density_threshold: f32,
    pub enable_out_of_plane: bool,
    pub background_e_field: [f32; 2],
    pub frame: u64,
}

impl Simulation {
    pub fn new(bodies: Bodies, domain_width: f32, domain_height: f32, theta: f32, epsilon: f32, leaf_capacity: usize,
               thread_capacity: usize) -> Result<Self> {
        let ctx = Context::new(0, bodies.len().max(1), bodies.electron_body.len().max(1), theta, epsilon, leaf_capacity,
                               thread_capacity)?;
        let mut sim = Simulation { ctx, bodies, domain_width, domain_height, domain_depth: 1.0, dt: 5.0,
                                   coulomb_constant: 0.138935, damping_base: 1.0, cell_list_density_threshold: 0.001,
                                   enable_out_of_plane: false, background_e_field: [0.0; 2], frame: 0 };
        sim.upload()?;
        Ok(sim)
    }
    /// Vec<Body> -> device (all hot fields + electrons)
    pub fn upload(&mut self) -> Result<()> {
        let b = &mut self.bodies;
        let n = b.len();
        b.id = (0..n as u32).collect();
        let c = &self.ctx;
        c.check(unsafe {
            psim_upload_bodies(c.raw(), n as u64, Bodies::flat(&b.pos), opt(&b.z), Bodies::flat(&b.vel), opt(&b.vz),
                               opt(&b.mass), opt(&b.radius), opt(&b.charge),
                               if b.species.is_empty() { std::ptr::null() } else { b.species.as_ptr() })
        })?;
        if !b.electron_body.is_empty() {
            c.check(unsafe {
                psim_upload_electrons(c.raw(), b.electron_body.len() as u64, b.electron_body.as_ptr(),
                                      Bodies::flat(&b.electron_rel_pos), Bodies::flat(&b.electron_vel))
            })?;
        }
        Ok(())
    }
    /// what the rest of `Simulation::step` changed on the host since the last call (collisions, foil logic ...)
    pub fn update_state(&mut self) -> Result<()> {
        let b = &self.bodies;
        self.ctx.check(unsafe {
            psim_update_state(self.ctx.raw(), b.len() as u64, Bodies::flat(&b.pos), Bodies::flat(&b.vel), opt(&b.charge))
        })
    }
    /// device -> Vec<Body> (positions, velocities, accelerations, fields, z / vz)
    pub fn download(&mut self) -> Result<()> {
        let n = self.bodies.len();
        let b = &mut self.bodies;
        b.z.resize(n, 0.0);
        b.vz.resize(n, 0.0);
        b.az.resize(n, 0.0);
        b.id.resize(n, 0);
        let (pos, vel) = (Bodies::flat_mut(&mut b.pos, n), Bodies::flat_mut(&mut b.vel, n));
        let (acc, ef) = (Bodies::flat_mut(&mut b.acc, n), Bodies::flat_mut(&mut b.e_field, n));
        self.ctx.check(unsafe {
            psim_download_bodies(self.ctx.raw(), pos, b.z.as_mut_ptr(), vel, b.vz.as_mut_ptr(), acc, b.az.as_mut_ptr(),
                                 std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(),
                                 ef, b.id.as_mut_ptr())
        })
    }
    pub fn quadtree(&self) -> Quadtree<'_> {
        Quadtree { ctx: &self.ctx }
    }
    pub fn cell_list(&self, cell_size: f32) -> CellList<'_> {
        CellList { ctx: &self.ctx, domain_width: self.domain_width, domain_height: self.domain_height, cell_size }
    }
    /// `Simulation::use_cell_list` (simulation.rs:1798-1802)
    pub fn use_cell_list(&self) -> bool {
        unsafe { psim_use_cell_list(self.ctx.raw(), self.domain_width, self.domain_height, self.cell_list_density_threshold) > 0 }
    }
    /// `Simulation::iterate` (simulation.rs:1437-1486)
    pub fn iterate(&mut self) -> Result<()> {
        self.ctx.check(unsafe {
            psim_iterate(self.ctx.raw(), self.dt, self.damping_base, self.domain_width, self.domain_height,
                         self.domain_depth, self.enable_out_of_plane as i32)
        })
    }
    /// the loop over `Body::update_electrons` (simulation.rs:1186-1196, body/electron.rs:19-46)
    pub fn update_electrons(&mut self) -> Result<()> {
        let bg = self.background_e_field;
        self.ctx.check(unsafe { psim_update_electrons(self.ctx.raw(), bg[0], bg[1], self.dt, self.coulomb_constant) })
    }
    /// `Simulation::update_surrounded_flags` (simulation.rs:1893-1918); returns `surrounded_by_metal` per body
    pub fn update_surrounded_flags(&mut self, radius_factor: f32, neighbor_threshold: u64) -> Result<Vec<u8>> {
        let c = &self.ctx;
        c.check(unsafe {
            psim_update_surrounded_flags(c.raw(), self.domain_width, self.domain_height, self.frame, radius_factor,
                                         neighbor_threshold)
        })?;
        let mut flags = vec![0u8; self.bodies.len()];
        c.check(unsafe { psim_get_surrounded(c.raw(), flags.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut()) })?;
        Ok(flags)
    }
    /// `out_of_plane::enforce_metal_z_boundaries` (simulation/out_of_plane.rs:140-254)
    pub fn enforce_metal_z_boundaries(&mut self, max_z: f32) -> Result<()> {
        self.ctx.check(unsafe { psim_enforce_metal_z_boundaries(self.ctx.raw(), max_z, self.domain_width, self.domain_height) })
    }
    /// the field part of the hopping candidate predicate (simulation/electron_hopping.rs:283-329) for a batch:
    /// `candidates[i]` are the acceptors of donor `src[i]`; returns (local_field per donor, alignment per pair)
    pub fn hop_alignment(&self, src: &[u32], candidates: &[Vec<u32>], alignment_bias: f32)
                         -> Result<(Vec<[f32; 2]>, Vec<f32>)> {
        let mut off = vec![0u32; src.len() + 1];
        for (i, c) in candidates.iter().enumerate() {
            off[i + 1] = off[i] + c.len() as u32;
        }
        let dst: Vec<u32> = candidates.iter().flatten().copied().collect();
        let mut field = vec![[0.0f32; 2]; src.len()];
        let mut al = vec![0.0f32; dst.len()];
        let bg = self.background_e_field;
        self.ctx.check(unsafe {
            psim_hop_alignment(self.ctx.raw(), src.len() as u64, src.as_ptr(), off.as_ptr(), dst.as_ptr(),
                               self.coulomb_constant, bg[0], bg[1], alignment_bias, field.as_mut_ptr() as *mut f32,
                               al.as_mut_ptr())
        })?;
        Ok((field, al))
    }
    pub fn step_params(&self) -> psim_step_params {
        psim_step_params { hw: self.domain_width, hh: self.domain_height, hd: self.domain_depth, dt: self.dt,
                           damping_base: self.damping_base, k_e: self.coulomb_constant, bg_x: self.background_e_field[0],
                           bg_y: self.background_e_field[1], density_threshold: self.cell_list_density_threshold,
                           enable_out_of_plane: self.enable_out_of_plane as u32, do_short_range: 1, do_electrons: 1,
                           do_iterate: 1, do_polar: 1, reserved: [0; 2] }
    }
    /// the whole hot path of `Simulation::step` (simulation.rs:1000-1196) without a host round trip
    pub fn step_device(&mut self) -> Result<()> {
        let p = self.step_params();
        self.ctx.check(unsafe { psim_step(self.ctx.raw(), &p) })?;
        self.frame += 1;
        Ok(())
    }
    /// the same across the communicator of `Context::comm_init` (one process / thread per GPU)
    pub fn step_sharded(&mut self) -> Result<()> {
        let p = self.step_params();
        self.ctx.check(unsafe { psim_step_sharded(self.ctx.raw(), &p) })?;
        self.frame += 1;
        Ok(())
    }
}

/// src/simulation/forces.rs — free functions taking the simulation, like the reference.
pub mod forces {
    use super::*;

    /// forces.rs:14-25: `quadtree.build` + (when `use_cell_list`) the grid at max(3 lj, repulsion, lj)
    pub fn prepare_spatial_structures(sim: &mut Simulation) -> Result<()> {
        let c = &sim.ctx;
        c.check(unsafe {
            psim_prepare_spatial_structures(c.raw(), sim.domain_width, sim.domain_height, sim.cell_list_density_threshold)
        })?;
        let mut perm = vec![0u32; sim.bodies.len()];
        c.check(unsafe { psim_get_permutation(c.raw(), perm.as_mut_ptr()) })?;
        sim.bodies.permute(&perm);
        Ok(())
    }
    /// forces.rs:33-44: field + background, `acc = q E / m`
    pub fn attract(sim: &mut Simulation) -> Result<()> {
        let n = sim.bodies.len();
        let (ef, acc) = (Bodies::flat_mut(&mut sim.bodies.e_field, n), Bodies::flat_mut(&mut sim.bodies.acc, n));
        let bg = sim.background_e_field;
        sim.ctx.check(unsafe { psim_field(sim.ctx.raw(), sim.coulomb_constant, bg[0], bg[1], 1, ef, acc) })
    }
    /// forces.rs:52-175 (dipole_model 1 = ConjugatePair, the default)
    pub fn apply_polar_forces(sim: &mut Simulation, dipole_model: i32) -> Result<()> {
        sim.ctx.check(unsafe { psim_apply_polar_forces(sim.ctx.raw(), sim.coulomb_constant, dipole_model) })
    }
    /// forces.rs:182-231
    pub fn apply_lj_forces(sim: &mut Simulation) -> Result<()> {
        sim.ctx.check(unsafe { psim_short_range(sim.ctx.raw(), PSIM_SR_LJ) })
    }
    /// forces.rs:250-289
    pub fn apply_repulsive_forces(sim: &mut Simulation) -> Result<()> {
        sim.ctx.check(unsafe { psim_short_range(sim.ctx.raw(), PSIM_SR_REPULSION) })
    }
    /// forces.rs:294-321
    pub fn apply_stack_pressure(sim: &mut Simulation) -> Result<()> {
        sim.ctx.check(unsafe { psim_short_range(sim.ctx.raw(), PSIM_SR_STACK_PRESSURE) })
    }
}
