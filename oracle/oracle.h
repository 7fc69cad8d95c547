/*
 * oracle.h — C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the force hot path of
 * PMantix/ParticleSim (Rust).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (libpsim_b200.so) never links, loads or calls anything in this directory.
 *
 * Parity status: the reference cannot be compiled in this image (no cargo/rustc, and its
 * `quarkstrom` dependency is an un-vendored sibling path), so this restatement is pinned
 * only by the reference's own known-answer tests for the path:
 *   - src/quadtree/tests.rs:9-78    single charge, radial / equal-magnitude field
 *   - src/quadtree/tests.rs:80-138  overlapping pair gives a finite field
 *   - src/body/tests/anion.rs:48-49 degenerate (leaf=1, thread=1) build does not fail
 *   - tests/physics_invariants/baselines/quadtree_force_error.json (statistical band)
 * Everything else (topology, permutation, LJ, integrator) is "parity unpinned" by the
 * reference: it has no golden vectors with inputs for them.
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef PSIM_ORACLE_H
#define PSIM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Node in the reference's field order (src/quadtree/node.rs:6-14), fixed C layout. */
typedef struct {
  uint64_t children; /* index of first of 4 contiguous children, 0 = leaf */
  uint64_t next;     /* skip pointer, 0 = end of traversal */
  float pos[2];      /* charge centre */
  float mass;
  float quad_center[2];
  float quad_size;
  uint64_t bodies_start;
  uint64_t bodies_end;
  float charge;
  uint32_t _pad;
} OrcNode;

/* One entry of the canonical DFS pre-order listing (SURVEY.md §8a "canonical forms"). */
typedef struct {
  uint64_t path_hi; /* quadrant digits of levels 33.. (2 bits each, MSB first) */
  uint64_t path_lo; /* quadrant digits of levels 1..32 (2 bits each, level 1 in bits 63:62) */
  uint32_t depth;
  uint32_t is_leaf;
  uint64_t start;
  uint64_t end;
  float pos[2];
  float mass;
  float charge;
  float quad_center[2];
  float quad_size;
  uint32_t _pad;
} OrcCanon;

typedef struct {
  uint64_t visits;   /* V: nodes whose MAC was evaluated */
  uint64_t accepts;  /* A: monopole terms */
  uint64_t pairs;    /* P: direct body terms actually summed (after the self skip) */
} OrcCounters;

/* Species property row, the columns the hot path reads (src/species.rs:7-24). */
typedef struct {
  float mass, radius, damping;
  float lj_epsilon, lj_sigma, lj_cutoff;
  float polar_offset, polar_charge;
  float repulsion_strength, repulsion_cutoff;
  uint32_t lj_enabled, repulsion_enabled;
} OrcSpecies;

typedef struct OrcSim OrcSim;

/* ---- lifecycle: an OrcSim owns a Vec<Body>-like AoS array and one Quadtree + CellList ---- */
OrcSim *orc_create(float theta, float epsilon, uint64_t leaf_capacity, uint64_t thread_capacity);
void orc_destroy(OrcSim *);
void orc_set_species_table(OrcSim *, const OrcSpecies *rows, uint32_t nrows);
/* default table copied from src/species.rs:26-408 + src/config.rs */
void orc_default_species_table(OrcSpecies *rows21);

/* bodies; arrays may be NULL (zeros / species 0).  ids are 0..n-1 in upload order. */
void orc_set_bodies(OrcSim *, uint64_t n, const float *pos_xy, const float *z, const float *vel_xy,
                    const float *vz, const float *mass, const float *radius, const float *charge,
                    const uint8_t *species);
/* electrons, flattened: body index (in CURRENT order), rel_pos, vel */
void orc_set_positions(OrcSim *, const float *pos_xy);
void orc_set_electrons(OrcSim *, uint64_t m, const uint32_t *body, const float *rel_xy,
                       const float *vel_xy);
uint64_t orc_num_bodies(const OrcSim *);
uint64_t orc_num_electrons(const OrcSim *);
/* any pointer may be NULL */
void orc_get_bodies(const OrcSim *, uint64_t *id, float *pos_xy, float *z, float *vel_xy, float *vz,
                    float *acc_xy, float *az, float *mass, float *radius, float *charge,
                    uint8_t *species, float *e_field_xy);
void orc_get_electrons(const OrcSim *, uint32_t *body, float *rel_xy, float *vel_xy);

/* ---- src/quadtree ---- */
/* mode 0 = Quadtree::build (tight AABB square), 1 = build_with_domain(hw, hh).
 * threads <= 1: single-worker replay (deterministic raw indices); >1: OpenMP workers that
 * mirror the rayon::broadcast worker pool (raw indices schedule dependent, shape is not). */
void orc_build(OrcSim *, int mode, float hw, float hh, int threads);
uint64_t orc_num_nodes(const OrcSim *);   /* highest live node index + 1 */
void orc_get_nodes(const OrcSim *, OrcNode *out);
uint64_t orc_canonical(const OrcSim *, OrcCanon *out, uint64_t cap); /* returns count */
/* test hook: overwrite node centres, canonical DFS pre-order */
void orc_set_canonical_pos(OrcSim *, const float *pos_xy, uint64_t count);
uint32_t orc_max_depth(const OrcSim *);
uint32_t orc_flags(const OrcSim *); /* bit0: some subdivision went deeper than 32 levels;
                                       bit1: a degenerate (refused) leaf exists;
                                       bit2: chain coincidence test differs from "all bit-identical";
                                       bit3: more than 4N+1024 nodes (the reference's fixed claim
                                             array, quadtree.rs:207, would be indexed out of bounds) */
/* Quadtree::field: e_field[i] = acc_pos(pos_i, 1, radius_i).  threads as above (rayon par_iter). */
void orc_field(OrcSim *, float k_e, int threads, OrcCounters *ctr);
/* acc_pos for arbitrary points (field_at_point is q=1, radius=0) */
void orc_acc_points(const OrcSim *, uint64_t m, const float *pts_xy, const float *q,
                    const float *radius, float k_e, float *out_xy, int threads, OrcCounters *ctr);
/* Quadtree::find_neighbors_within / CellList::find_neighbors_within for body i; returns count,
 * writes up to cap indices in the reference's order */
uint64_t orc_tree_neighbors(const OrcSim *, uint64_t i, float cutoff, uint64_t *out, uint64_t cap);

/* ---- src/cell_list.rs ---- */
void orc_cell_set_domain(OrcSim *, float hw, float hh);
void orc_cell_rebuild(OrcSim *, float cell_size);
void orc_cell_dims(const OrcSim *, uint64_t *gx, uint64_t *gy);
uint64_t orc_cell_contents(const OrcSim *, uint64_t cell, uint64_t *out, uint64_t cap);
uint64_t orc_cell_neighbors(const OrcSim *, uint64_t i, float cutoff, uint64_t *out, uint64_t cap);
uint64_t orc_cell_metal_neighbor_count(const OrcSim *, uint64_t i, float cutoff);

/* ---- src/simulation/forces.rs + Simulation::iterate + Body::update_electrons ---- */
int orc_use_cell_list(const OrcSim *, float hw, float hh, float density_threshold);
void orc_reset_acc(OrcSim *);
void orc_prepare_spatial_structures(OrcSim *, float hw, float hh, float density_threshold,
                                    int threads);
void orc_attract(OrcSim *, float k_e, float bg_x, float bg_y, int threads);
/* forces.rs:52-175; dipole_model 0 = SingleOffset, 1 = ConjugatePair (default) */
void orc_apply_polar_forces(OrcSim *, int use_cell, float k_e, int dipole_model);
void orc_apply_lj_forces(OrcSim *, int use_cell, float lj_force_max, uint32_t collision_passes);
void orc_apply_repulsive_forces(OrcSim *, int use_cell);
void orc_apply_stack_pressure(OrcSim *, int enabled, float pressure, float decay, float hw);
void orc_iterate(OrcSim *, float dt, float damping_base, float hw, float hh, float hd,
                 int enable_out_of_plane, int threads);
/* the serial loop at simulation.rs:1186-1196 (threads>1 = the "all-parallel" variant) */
void orc_update_electrons(OrcSim *, float bg_x, float bg_y, float dt, float k_e, int threads);
/* simulation.rs:1893-1918 + body/types.rs:243-286; out_of_plane.rs:140-254 */
void orc_update_surrounded_flags(OrcSim *, float hw, float hh, float density_threshold, uint64_t frame,
                                 float radius_factor, uint64_t neighbor_threshold);
void orc_get_surrounded(const OrcSim *, uint8_t *flags, float *last_pos_xy, uint64_t *last_frame);
void orc_enforce_metal_z_boundaries(OrcSim *, float max_z, float hw, float hh, float density_threshold);

/* ---- FP64 direct O(N^2) field with the same softening (physics_invariants.rs:2430-2450) ----
 * target_radius NULL => 0.  Sources are the current bodies. */
void orc_direct_f64(const OrcSim *, uint64_t m, const float *pts_xy, const float *target_radius,
                    double k_e, double epsilon, double *out_xy, int threads);

/* collision::collide (simulation/collision.rs:62-372), one pass: every pair of intersecting bounding squares is
 * handed to resolve() - here in index order (i ascending, j ascending), one of the orders the reference's thread pool
 * may produce.  Returns the number of pairs that touched. */
uint64_t orc_collide(OrcSim *, float domain_depth, uint32_t num_passes, float li_collision_softness,
                     int soft_collision_lithium_ion, int soft_collision_anion);
/* simulation/electron_hopping.rs:283-329, the field part of the candidate predicate for a batch of (donor, acceptor)
 * pairs in CSR form; local_field_xy may be NULL */
void orc_hop_alignment(const OrcSim *, uint64_t m_src, const uint32_t *src_idx, const uint32_t *pair_offsets,
                       const uint32_t *dst_idx, float k_e, float bg_x, float bg_y, float alignment_bias,
                       float *local_field_xy, float *alignment);
int orc_max_threads(void);
/* 1 if Vec2::mag_sq/dot were compiled as mul_add (ORC_UV_FMA), else 0 */
int orc_uv_fma(void);

#ifdef __cplusplus
}
#endif
#endif
