/*
 * psim_b200.h — C ABI of libpsim_b200.so, the B200 (sm_100a) implementation of ParticleSim's
 * force hot path.  Plain pointers and sizes only; no torch, no C++ types.
 *
 * The reference (PMantix/ParticleSim, Rust) has no FFI layer for this path: the boundary is the
 * method surface of `Quadtree` and `CellList` that `src/simulation` calls.  Each entry point below
 * names the reference interface it replaces (file:line relative to the reference root).  The Rust
 * binding a maintainer would add is shown in INTEGRATION.md and checked in under rust/.
 *
 * Threading: one host thread per context; contexts are independent (one per GPU).
 * Ownership: the caller owns every host buffer it passes; the library owns all device memory.
 * Errors: every function returns 0 (PSIM_OK) or a negative PSIM_E_* code and never aborts or
 * throws across the ABI; psim_last_error() gives the message of the last failure on a context.
 * There is no CPU fallback: without a CUDA device psim_create fails with PSIM_E_CUDA.
 *
 * Body order: like the reference's in-place partition (quadtree.rs:56-63), psim_build PERMUTES
 * the bodies held by the context into tree order.  Every index-based call afterwards refers to the
 * post-build order; psim_get_permutation tells the host how to reorder its own Vec<Body>.
 */
#ifndef PSIM_B200_H
#define PSIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSIM_OK 0
#define PSIM_E_CUDA (-1)
#define PSIM_E_ARG (-2)
#define PSIM_E_OOM (-3)
#define PSIM_E_NODE_OVERFLOW (-4) /* more tree nodes than the arena holds (raise node_factor) */
#define PSIM_E_STATE (-5)         /* call order: e.g. field before build */
#define PSIM_E_NCCL (-6)

#define PSIM_BUILD_CONTAINING 0 /* Quadtree::build, quadtree.rs:153-170 (root = tight AABB square) */
#define PSIM_BUILD_DOMAIN 1     /* Quadtree::build_with_domain, quadtree.rs:173-195 */

#define PSIM_SR_LJ 1u             /* forces.rs:182-231 */
#define PSIM_SR_REPULSION 2u      /* forces.rs:250-289 */
#define PSIM_SR_STACK_PRESSURE 4u /* forces.rs:294-321 */

typedef struct psim_ctx psim_ctx;

/* Quadtree::new(theta, epsilon, leaf_capacity, thread_capacity) (quadtree.rs:24-34) plus the
 * SimConfig fields the path reads (config.rs:126,204,213-216; units.rs:32-34). */
typedef struct {
  float theta;
  float epsilon;
  uint32_t leaf_capacity;
  uint32_t thread_capacity; /* only its effect on the tree shape is kept */
  float lj_force_max;       /* LJ_FORCE_MAX = 200 */
  uint32_t collision_passes;/* COLLISION_PASSES = 7 (LJ clamp = passes * lj_force_max) */
  uint32_t stack_pressure_enabled;
  float stack_pressure;
  float stack_pressure_decay;
  uint32_t parity_mode;     /* 1: IEEE sqrt/div, reference operation order (default); 0: fast math */
  float node_factor;        /* node arena = node_factor * max_bodies + 1024 compact nodes (default 4) */
  uint32_t strict_centres;  /* 1 (default): every internal node's centre is the reference's serial f32 running
                               sum over its body range (quadtree.rs:114-139), evaluated in parallel but bit for
                               bit (csrc/strict.cuh) - node centres equal the reference's for leaf_capacity 1
                               and fields meet the 1e-5 bar against it; 0: f64 sums carried up the tree
                               (~10 % faster builds, fields ~1e-4 from the reference: its own summation noise).
                               Both modes work in sharded builds (each rank sums the centres of its own piece). */
  uint32_t reserved[4];
} psim_config;

/* one row per Species (body/types.rs:12-36), the SpeciesProps columns the path reads (species.rs:7-24) */
typedef struct {
  float mass, radius, damping;
  float lj_epsilon, lj_sigma, lj_cutoff;
  float polar_offset, polar_charge;
  float repulsion_strength, repulsion_cutoff;
  uint32_t lj_enabled, repulsion_enabled;
} psim_species;

/* Node in the reference's field order (quadtree/node.rs:6-14), fixed C layout, 64 bytes. */
typedef struct {
  uint64_t children; /* first of 4 contiguous children, 0 = leaf */
  uint64_t next;     /* skip pointer, 0 = end */
  float pos[2];
  float mass;
  float quad_center[2];
  float quad_size;
  uint64_t bodies_start, bodies_end;
  float charge;
  uint32_t _pad;
} psim_node;

typedef struct {
  uint64_t n_bodies;
  uint64_t n_electrons;
  uint64_t compact_nodes;   /* non-empty cells held on the device */
  uint64_t reference_nodes; /* 4 * internal + 1: what Vec<Node> holds in the reference */
  uint32_t max_depth;
  uint32_t depth_cap;       /* first depth refused by the 1e-6 size rule / 32-level key */
  uint32_t zero_leaves;     /* refused or thread-capacity leaves with zeroed aggregates (Q2) */
  uint32_t cap_leaves;      /* multi-body leaves that stopped at depth_cap */
  float root_center[2];
  float root_size;
  uint32_t grid_x, grid_y;
  uint64_t traversal_warp_steps; /* nodes visited by warps since psim_reset_counters */
  uint64_t kernel_launches;      /* kernels launched by this context since psim_reset_counters */
} psim_stats;

/* ---- lifecycle ---- */
void psim_default_config(psim_config *cfg);
void psim_default_species_table(psim_species *rows21); /* species.rs:26-408 */
int32_t psim_create(int32_t device, uint64_t max_bodies, uint64_t max_electrons,
                    const psim_config *cfg, psim_ctx **out);
int32_t psim_destroy(psim_ctx *ctx);
const char *psim_last_error(const psim_ctx *ctx);
int32_t psim_set_config(psim_ctx *ctx, const psim_config *cfg);
/* run on this CUDA stream (cudaStream_t as an integer); 0 = the legacy default stream */
int32_t psim_set_stream(psim_ctx *ctx, uint64_t cuda_stream);
int32_t psim_sync(psim_ctx *ctx);
int32_t psim_stats_get(psim_ctx *ctx, psim_stats *out);
int32_t psim_reset_counters(psim_ctx *ctx);
/* Diagnostics for the roofline (not on the hot path):
 * interaction counters of Quadtree::field on the current tree, summed over all bodies:
 * out[0] = internal nodes opened, out[1] = monopoles accepted (non-empty nodes), out[2] = direct
 * body terms, out[3] = warp steps.  The reference's node visits are n + 4 * out[0]. */
int32_t psim_field_counters(psim_ctx *ctx, uint64_t *out4);
/* how the last psim_build (strict_centres on, single GPU) made the node charges:
 * out[0] = 1 if every body charge is an integer (what Body::update_charge_from_electrons produces, body/electron.rs)
 * with sum |q| < 2^24 and no leaf is left unaggregated - node charges are then exact differences of an integer prefix
 * over the charged bodies, equal to the reference's nested f32 sums (quadtree.rs:142-149) bit for bit; 0 = bottom-up
 * level sweeps in the reference's child order.  out[1] = charged bodies, out[2] = sum |q| of the integer charges,
 * out[3] = bodies whose charge is not an integer below 2^20. */
int32_t psim_build_info(psim_ctx *ctx, uint64_t *out4);
/* measured FP32 FMA throughput of this device in TFLOP/s (2 flops per FMA), and the SM count */
int32_t psim_fp32_peak(psim_ctx *ctx, float *tflops, int32_t *sm_count);

/* ---- data in / out (host pointers; arrays marked opt may be NULL) ---- */
/* species.rs get_species_props table; nrows <= 32 */
int32_t psim_upload_species_table(psim_ctx *ctx, const psim_species *rows, uint32_t nrows);
/* Vec<Body> hot fields (body/types.rs:38-62).  opt NULL => zeros (mass: 1). */
int32_t psim_upload_bodies(psim_ctx *ctx, uint64_t n, const float *pos_xy, const float *z_opt,
                           const float *vel_xy_opt, const float *vz_opt, const float *mass_opt,
                           const float *radius_opt, const float *charge_opt,
                           const uint8_t *species_opt);
/* the per-step host -> device refresh of a resident body set: positions, velocities and charges
 * (what the rest of Simulation::step changes on the host); electrons and all other fields stay */
int32_t psim_update_state(psim_ctx *ctx, uint64_t n, const float *pos_xy, const float *vel_xy_opt,
                          const float *charge_opt);
/* only positions / charges changed on the host */
int32_t psim_update_positions(psim_ctx *ctx, uint64_t n, const float *pos_xy);
int32_t psim_update_charges(psim_ctx *ctx, uint64_t n, const float *charge);
/* Body::electrons flattened: body index (current order), rel_pos, vel (body/electron.rs:9-13) */
int32_t psim_upload_electrons(psim_ctx *ctx, uint64_t m, const uint32_t *body, const float *rel_xy,
                              const float *vel_xy_opt);
/* every pointer opt; orig_index[i] = index the body had at psim_upload_bodies time */
int32_t psim_download_bodies(psim_ctx *ctx, float *pos_xy, float *z, float *vel_xy, float *vz,
                             float *acc_xy, float *az, float *mass, float *radius, float *charge,
                             uint8_t *species, float *e_field_xy, uint32_t *orig_index);
int32_t psim_download_electrons(psim_ctx *ctx, uint32_t *body, float *rel_xy, float *vel_xy);

/* ---- src/quadtree ---- */
/* Quadtree::build (mode 0) / build_with_domain(hw, hh) (mode 1): quadtree.rs:153-195.
 * Morton keys -> onesweep sort -> tree construction -> bottom-up aggregation.  Permutes bodies. */
int32_t psim_build(psim_ctx *ctx, int32_t mode, float hw, float hh);
/* same launches without the final host synchronisation / arena-overflow check (psim_stats_get or
 * psim_sync + psim_build_status report it later); for callers that chain phases on the stream */
int32_t psim_build_async(psim_ctx *ctx, int32_t mode, float hw, float hh);
int32_t psim_build_status(psim_ctx *ctx);
/* out[i] = position, before the last psim_build, of the body now at i */
int32_t psim_get_permutation(psim_ctx *ctx, uint32_t *out);
/* 32-level quadrant keys of the bodies in current (sorted) order */
int32_t psim_get_keys(psim_ctx *ctx, uint64_t *out);
/* Vec<Node> for the renderer / diagnostics consumers (renderer/draw/mod.rs:62-63,851-900) */
int32_t psim_download_nodes(psim_ctx *ctx, psim_node *out, uint64_t cap, uint64_t *count);
/* Quadtree::field (quadtree.rs:418-427) + the attract epilogue (forces.rs:37-43):
 * e_field = acc_pos(pos, 1, radius) + bg; if write_acc: acc = charge * e_field / mass.
 * out pointers opt (host, current order). */
int32_t psim_field(psim_ctx *ctx, float k_e, float bg_x, float bg_y, int32_t write_acc,
                   float *out_e_field_xy, float *out_acc_xy);
/* Quadtree::acc_pos (quadtree.rs:350-407) for m points; q NULL => 1, radius NULL => 0, which is
 * Quadtree::field_at_point (quadtree.rs:504-507) */
int32_t psim_acc_points(psim_ctx *ctx, uint64_t m, const float *pts_xy, const float *q_opt,
                        const float *radius_opt, float k_e, float *out_xy);
/* collision::collide (src/simulation/collision.rs:62-372), `passes` passes (Simulation::step runs COLLISION_PASSES of
 * them, simulation.rs:1025-1028), each with correction scale 1 / num_passes.  Broad phase on the cell list (cell =
 * the largest diameter present), then resolve() for every pair of intersecting bounding squares from the state at
 * the start of the pass, the changes of a body's pairs summed (the reference resolves the pairs in whatever order its
 * thread pool reaches them, on shared mutable state - there is no order to reproduce; for a body in one pair the
 * result is resolve()'s, see csrc/collide.cuh).  li_collision_softness / soft_collision_* are SimConfig's
 * (config.rs:235,358-362,483-485).  Moves bodies: tree and grid are stale afterwards.  *touching_pairs (opt)
 * receives the number of touching pairs found by the LAST pass. */
int32_t psim_collide(psim_ctx *ctx, float hw, float hh, float domain_depth, uint32_t passes, uint32_t num_passes,
                     float li_collision_softness, int32_t soft_collision_lithium_ion, int32_t soft_collision_anion,
                     uint64_t *touching_pairs);
/* Electron hopping, the field part of the candidate predicate (simulation/electron_hopping.rs:283-329), batched:
 * m_src donors (indices in the current body order), their candidate acceptors in CSR form (pair_offsets[m_src + 1]
 * into dst_idx).  For every donor: local_field = (bg_x, bg_y) + Quadtree::field_at_point(bodies[src].pos)
 * (:290-295, one Barnes-Hut walk per donor instead of one per candidate); for every pair:
 * alignment = max(0, -hop_dir . field_dir) (1 if the field vanishes) * max(alignment_bias, 0), raised to 0.5 on a
 * metal / electrode conduction path (:301-328).  out_local_field_xy may be NULL.  Needs a tree (psim_build). */
int32_t psim_hop_alignment(psim_ctx *ctx, uint64_t m_src, const uint32_t *src_idx, const uint32_t *pair_offsets,
                           const uint32_t *dst_idx, float k_e, float bg_x, float bg_y, float alignment_bias,
                           float *out_local_field_xy, float *out_alignment);
/* the loop at simulation.rs:1186-1196 over Body::update_electrons (body/electron.rs:19-46) */
int32_t psim_update_electrons(psim_ctx *ctx, float bg_x, float bg_y, float dt, float k_e);

/* ---- src/cell_list.rs ---- */
/* CellList::update_domain_size + cell_size + rebuild (cell_list.rs:27-45) */
int32_t psim_cell_build(psim_ctx *ctx, float hw, float hh, float cell_size);
/* CSR dump: offsets[gx*gy + 1], indices[n] (per-cell body lists in index order) */
int32_t psim_cell_download(psim_ctx *ctx, uint64_t *gx, uint64_t *gy, uint32_t *offsets_opt,
                           uint32_t *indices_opt);
/* CellList::find_neighbors_within (cell_list.rs:57-85) for m bodies, CSR result in the reference's
 * order; metals_only = metal_neighbor_count's filter (cell_list.rs:92-127).  indices_opt may be
 * NULL to get offsets (counts) only; *total receives offsets[m]. */
int32_t psim_neighbors_within(psim_ctx *ctx, uint64_t m, const uint32_t *body_idx, float cutoff,
                              int32_t metals_only, uint32_t *offsets, uint32_t *indices_opt,
                              uint64_t indices_cap, uint64_t *total);

/* ---- src/simulation/forces.rs + Simulation::iterate ---- */
/* simulation.rs:1000-1003 */
int32_t psim_reset_acc(psim_ctx *ctx);
/* simulation.rs:1798-1802 */
int32_t psim_use_cell_list(const psim_ctx *ctx, float hw, float hh, float density_threshold);
/* forces.rs:14-25: build(CONTAINING) and, when use_cell_list, the grid at
 * cell_size = max(3 * max_lj_cutoff, max_repulsion_cutoff, max_lj_cutoff) */
int32_t psim_prepare_spatial_structures(psim_ctx *ctx, float hw, float hh, float density_threshold);
/* apply_lj_forces / apply_repulsive_forces / apply_stack_pressure accumulated into acc.
 * Needs a cell grid whose cell_size >= the largest cutoff in use. */
int32_t psim_short_range(psim_ctx *ctx, uint32_t flags);
/* forces::apply_polar_forces (forces.rs:52-175), accumulated into acc: EC / DMC bodies with a bound
 * electron against their neighbours within 3 * radius.  dipole_model 0 = SingleOffset, 1 = ConjugatePair
 * (the default, config.rs:278-281).  Needs a cell grid (any cell size) built after the last psim_build. */
int32_t psim_apply_polar_forces(psim_ctx *ctx, float k_e, int32_t dipole_model);
/* Simulation::iterate (simulation.rs:1437-1486); base damping = damping_base ^ (dt / 0.01) */
int32_t psim_iterate(psim_ctx *ctx, float dt, float damping_base, float hw, float hh, float hd,
                     int32_t enable_out_of_plane);

/* One pass of the hot path exactly in Simulation::step's order (simulation.rs:1000-1196), with no
 * host round trip in between: reset acc -> prepare_spatial_structures -> field + attract ->
 * LJ / repulsion / stack pressure -> iterate -> build_with_domain -> update_electrons. */
typedef struct {
  float hw, hh, hd;
  float dt, damping_base;
  float k_e;
  float bg_x, bg_y;
  float density_threshold;
  uint32_t enable_out_of_plane;
  uint32_t do_short_range;  /* 0: Coulomb only (config 2) */
  uint32_t do_electrons;    /* second build + electron field sampling + drift */
  uint32_t do_iterate;
  uint32_t do_polar;        /* 1: apply_polar_forces (ConjugatePair) between attract and the LJ pass, on a
                               grid of the reference's size max(3 * lj, repulsion, lj) (forces.rs:17-22) */
  uint32_t reserved[2];
} psim_step_params;
int32_t psim_step(psim_ctx *ctx, const psim_step_params *p);
/* The same step for a caller whose Vec<Body> lives in host memory (the reference's Simulation owns its
 * bodies on the host, simulation.rs:40-60): psim_update_state + psim_step + the read-back of the
 * results in one call, with the transfers pipelined against the device work.
 *   in : positions are copied first; charges and velocities follow on a copy stream while the keys are
 *        generated and sorted, and are applied where the step first needs them;
 *   out: each result leaves on a second copy stream as soon as it is final (original indices after the
 *        first build, fields after the traversal, positions and velocities after the integrator),
 *        underneath the second build and the electron field sampling.
 * Row order: the outputs are in the body order of the step's first build (the order quadtree.build
 * leaves in Vec<Body>, simulation.rs:1004), out_orig_index[i] = the upload index of row i.  The inputs
 * of the next psim_step_host call are in the order of these outputs (the first call: the order the
 * device holds, i.e. the upload order or the last downloaded order); the context keeps the extra
 * permutation of the electron pass to itself.  Any other call that reads or writes bodies sees the
 * device order: follow it with psim_download_bodies before mixing the two styles.
 * Host arrays should be page-locked (cudaHostAlloc / cudaHostRegister) or the copies serialise.
 * Optional arrays may be NULL.  Synchronises, and returns the build status (PSIM_E_NODE_OVERFLOW ...). */
int32_t psim_step_host(psim_ctx *ctx, const psim_step_params *p, uint64_t n, const float *pos_xy,
                       const float *vel_xy_opt, const float *charge_opt, float *out_pos_xy_opt,
                       float *out_vel_xy_opt, float *out_e_field_xy_opt, uint32_t *out_orig_index_opt);
/* Device time of each phase of the last psim_step, in ms (CUDA events on the context's stream;
 * synchronises).  Names follow the reference's profile scopes (src/profiler.rs users):
 * [0] quadtree_build  [1] cell_list_rebuild  [2] quadtree_field (+attract)  [3] forces_polar + forces_lj/repulsion
 * [4] iterate  [5] quadtree_build_domain  [6] electron_updates  [7] whole step */
#define PSIM_NUM_PHASES 8
int32_t psim_phase_times(psim_ctx *ctx, float *ms8);

/* ---- multi-GPU (one context per rank) ----------------------------------------------------------
 * Every rank holds the whole body set and builds the same tree (deterministic sort => identical
 * order on every rank); rank r OWNS a contiguous range of the sorted (Morton) order and computes
 * field / short-range / integrator only for it, and a contiguous range of the electrons.  After the
 * integrator and after the electron update the owned slices are exchanged with an NCCL all-gather
 * issued by the caller (particlesim_b200/parallel.py) directly on the device arrays below. */
int32_t psim_set_target_range(psim_ctx *ctx, uint64_t first, uint64_t count);
int32_t psim_set_electron_range(psim_ctx *ctx, uint64_t first, uint64_t count);
/* out[0..5] = device addresses of the CURRENT body / electron arrays (they swap at every psim_build):
 * pos_charge_radius (float4 per body: x, y, charge, radius), vel_z_vz (float4), acc_mass (float4),
 * e_field (float2), electron rel_pos (float2), electron vel (float2); out[6] = body capacity,
 * out[7] = electron capacity */
int32_t psim_device_ptrs(psim_ctx *ctx, uint64_t *out8);
/* Simulation::update_surrounded_flags (src/simulation/simulation.rs:1893-1918): rebins the cell list at
 * max_lj_cutoff and, for every body that moved more than SURROUND_MOVE_THRESHOLD * radius or was last
 * checked SURROUND_CHECK_INTERVAL frames ago (body/types.rs:243-286, config.rs:186-188), recounts its
 * LithiumMetal / FoilMetal neighbours within radius * radius_factor (CellList::metal_neighbor_count,
 * cell_list.rs:92-127) and sets surrounded_by_metal = count >= neighbor_threshold.  radius_factor and
 * neighbor_threshold are the reference's runtime-mutable values (renderer/state.rs:30-36; defaults 4.0, 8).
 * The per-body state (flag, last position, last frame) lives on the device, keyed by original index, and is
 * initialised by psim_upload_bodies like Body::new (last position = position, frame 0). */
int32_t psim_update_surrounded_flags(psim_ctx *ctx, float hw, float hh, uint64_t frame, float radius_factor,
                                     uint64_t neighbor_threshold);
/* the state in the CURRENT body order; any pointer may be NULL.  Synchronises. */
int32_t psim_get_surrounded(psim_ctx *ctx, uint8_t *flags, float *last_pos_xy, uint64_t *last_frame);
/* enforce_metal_z_boundaries (src/simulation/out_of_plane.rs:140-254): rebins at 4 x the metal radius and
 * clamps z / zeroes vz of every non-metal body against its first five metal neighbours in the reference's
 * cell order, then Body::clamp_z(max_z).  (The reference walks the tree instead when the density is below
 * the cell-list threshold; the neighbour sets agree, the order of the first five may not.) */
int32_t psim_enforce_metal_z_boundaries(psim_ctx *ctx, float max_z, float hw, float hh);

/* Sharded build (replaces the single work queue of Quadtree::build_internal,
 * src/quadtree/quadtree.rs:197-345, across ranks).  Every rank holds all bodies; rank r sorts the keys of
 * a contiguous range of 65 536 top-level cells and emits the tree nodes that start in it.  The pieces
 * concatenate into exactly the single-GPU tree.  psim_shard_phase runs phase 0..6 in order; between
 * phases the caller performs the exchange on the device buffers psim_shard_ptrs names:
 *   after 0: nothing (out = first sorted body of each rank, world + 1 entries; synchronises)
 *   after 1: all-gather of the sorted-order segments  [out[r], out[r+1]) of ptrs[0] (uint32 per body)
 *   after 2, after 4: all-reduce (sum) of ptrs[1], ptrs[5] uint64 words
 *   after 3: all-reduce (sum) of ptrs[2], ptrs[6] uint64 words
 *   after 5: all-gather of the traversal segments [out[r], out[r+1]) of ptrs[3] and ptrs[4] (16 bytes
 *            per node each; out = first traversal node of each rank; synchronises), then phase 6
 *            (children links over the gathered array; the tree is usable after it)
 * ptrs[7] = node capacity.  The tree export (psim_download_nodes) and parity_mode 2 need the whole node
 * array and are single-GPU only. */
int32_t psim_shard_init(psim_ctx *ctx, uint32_t rank, uint32_t world);
int32_t psim_shard_phase(psim_ctx *ctx, int32_t phase, int32_t mode, float hw, float hh, uint32_t *out);
int32_t psim_shard_ptrs(psim_ctx *ctx, uint64_t *out8);
/* ---- the same, orchestrated by the library over NCCL (NVLink / NVSwitch) ----------------------------------------
 * One context per rank, created with max_bodies / max_electrons >= psim_shard_capacity(n, nranks) (slices are padded
 * to a multiple of 64 bodies so that the 32-target groups of the walk are the single-GPU ones).  Every rank uploads
 * the same bodies (psim_upload_bodies / psim_upload_electrons), rank 0 makes a 128-byte id with psim_comm_unique_id
 * and the host distributes it (MPI, a file, torch.distributed ...).  libnccl.so.2 is loaded at run time (the copy
 * already in the process, else $PSIM_NCCL_LIB, else the system's); failures return PSIM_E_NCCL.
 *   psim_build_sharded = Quadtree::build / build_with_domain (quadtree.rs:153-195) with each rank building the part
 *                        of the tree that starts in its key range (the seven phases above + their exchanges);
 *   psim_step_sharded  = psim_step with both builds sharded, each rank computing field / polar / short-range /
 *                        integrator for its slice of the Morton order and an equal slice of the electrons, the slices
 *                        all-gathered in place (positions at once, velocities behind the next build's first phases).
 * Traversal exchange: by default each rank sends every other rank only the traversal records that rank's own targets
 * (its slice of the bodies, the sample points of its slice of the electrons) can reach - the locally essential tree,
 * a geometric superset test on the parent cell against the destination's bins (csrc/let.cuh); the records keep their
 * global indices, so the walks and their results are unchanged.  Consequence: after such a build a rank answers
 * field queries for ITS targets only; psim_acc_points / psim_hop_alignment return PSIM_E_STATE.  PSIM_LET=0 in the
 * environment at psim_comm_init time restores the full all-gather (every rank can then query any point).
 * psim_comm_stats: out[0] = records this rank sent in the last exchange, out[1] = what the full all-gather would have
 * sent, out[2] = LET enabled, out[3] = the current tree is a LET.
 * Results are bit-identical to psim_step on one GPU (same psim_config; strict_centres included: every rank sums the
 * reference's serial f32 centres for the nodes of its own piece).  psim_phase_times works after psim_step_sharded (the exchanges are inside the phases they follow). */
uint64_t psim_shard_capacity(uint64_t n, uint32_t nranks);
int32_t psim_comm_unique_id(uint8_t *out128);
int32_t psim_comm_init(psim_ctx *ctx, const uint8_t *unique_id128, uint32_t rank, uint32_t nranks);
int32_t psim_comm_destroy(psim_ctx *ctx);
int32_t psim_comm_stats(psim_ctx *ctx, uint64_t *out4);
int32_t psim_build_sharded(psim_ctx *ctx, int32_t mode, float hw, float hh);
int32_t psim_step_sharded(psim_ctx *ctx, const psim_step_params *p);
/* the caller wrote positions into the device arrays (e.g. an all-gather): tree and grid are stale */
int32_t psim_mark_positions_changed(psim_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
