// psim_b200.hpp — C++ host-side mirror of the reference's interface for the force hot path, over the
// C ABI of psim_b200.h.  The reference is compiled Rust; no Rust toolchain exists in this image, so
// this header plays the role of the FFI crate's safe wrapper (rust/psim-b200-sys shows the Rust one):
// same names, argument meaning and side effects as
//   Quadtree      src/quadtree/quadtree.rs   new / build / build_with_domain / field / acc_pos /
//                                            field_at_point / find_neighbors_within / nodes
//   CellList      src/cell_list.rs           new / rebuild / update_domain_size / find_neighbors_within /
//                                            metal_neighbor_count
//   forces::*     src/simulation/forces.rs   prepare_spatial_structures / attract / apply_lj_forces /
//                                            apply_repulsive_forces / apply_stack_pressure
//   Simulation    src/simulation/simulation.rs  iterate (:1437-1486), use_cell_list (:1798-1802),
//                                            the update_electrons loop (:1186-1196), step hot path
// Header only; link with -lpsim_b200.  All arithmetic runs in the sm_100a kernels; a failing call
// throws psim::Error (the reference's convention is "never fail"; a missing GPU must not be silent).
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "psim_b200.h"

namespace psim {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct Vec2 {
  float x = 0, y = 0;
};

// body/electron.rs:9-13
struct Electron {
  Vec2 rel_pos, vel;
};

// body/types.rs:12-36 (enum order = row index of the species table)
enum class Species : uint8_t {
  LithiumIon, LithiumMetal, FoilMetal, ElectrolyteAnion, EC, DMC, VC, FEC, EMC, LLZO, LLZT, S40B, SEI,
  Graphite, HardCarbon, SiliconOxide, LTO, LFP, LMFP, NMC, NCA
};

// body/types.rs:38-62, the fields the path reads or writes
struct Body {
  Vec2 pos;
  float z = 0;
  Vec2 vel;
  float vz = 0;
  Vec2 acc;
  float az = 0;
  float mass = 1, radius = 0, charge = 0;
  uint64_t id = 0;
  Species species = Species::LithiumIon;
  std::vector<Electron> electrons;
  Vec2 e_field;
};

using Node = psim_node;  // quadtree/node.rs:6-14 field order

// config.rs:337-339,409-417 (the SimConfig fields the path reads)
struct SimConfig {
  float coulomb_constant = 0.138935f;  // units.rs:32-34
  float damping_base = 1.0f;
  float cell_list_density_threshold = 0.001f;
  bool stack_pressure_enabled = false;
  float stack_pressure = 0.0f, stack_pressure_decay = 1.0f;
  bool enable_out_of_plane = false;
};

class Simulation;

// quadtree.rs:11-34
class Quadtree {
 public:
  static constexpr size_t ROOT = 0;
  float t_sq, e_sq;
  size_t leaf_capacity, thread_capacity;
  Quadtree(float theta, float epsilon, size_t leaf_capacity_, size_t thread_capacity_)
      : t_sq(theta * theta), e_sq(epsilon * epsilon), leaf_capacity(leaf_capacity_),
        thread_capacity(thread_capacity_), theta_(theta), epsilon_(epsilon) {}
  void build(std::vector<Body>& bodies);                                                  // :153-170
  void build_with_domain(std::vector<Body>& bodies, float domain_width, float domain_height);  // :173-195
  void field(std::vector<Body>& bodies, float k_e);                                       // :418-427
  Vec2 acc_pos(Vec2 pos, float q, float radius, const std::vector<Body>& bodies, float k_e) const;  // :350-407
  Vec2 field_at_point(const std::vector<Body>& bodies, Vec2 pos, float k_e) const;        // :504-507
  std::vector<Vec2> field_at_points(const std::vector<Vec2>& pts, float k_e) const;       // batch form
  std::vector<size_t> find_neighbors_within(const std::vector<Body>& bodies, size_t i, float cutoff) const;  // :430-501
  std::vector<Node> nodes() const;  // the Vec<Node> the renderer snapshots

 private:
  friend class Simulation;
  float theta_, epsilon_;
  Simulation* sim_ = nullptr;
};

// cell_list.rs:4-25
class CellList {
 public:
  float domain_width, domain_height, cell_size;
  CellList(float w, float h, float cs) : domain_width(w), domain_height(h), cell_size(cs) {}
  void update_domain_size(float w, float h) { domain_width = w, domain_height = h; }
  void rebuild(const std::vector<Body>& bodies);
  std::vector<size_t> find_neighbors_within(const std::vector<Body>& bodies, size_t i, float cutoff) const;
  size_t metal_neighbor_count(const std::vector<Body>& bodies, size_t i, float cutoff) const;

 private:
  friend class Simulation;
  Simulation* sim_ = nullptr;
};

// The slice of `Simulation` the hot path touches (simulation.rs:83-139): owns Vec<Body>, the
// quadtree, the cell list and — here — the device context.
class Simulation {
 public:
  std::vector<Body> bodies;
  Quadtree quadtree;
  CellList cell_list;
  SimConfig config;
  float domain_width, domain_height, domain_depth = 1.0f, dt = 5.0f;
  Vec2 background_e_field;

  Simulation(float domain_w, float domain_h, float theta = 1.0f, float epsilon = 2.0f, size_t leaf_capacity = 1,
             size_t thread_capacity = 1024, size_t max_bodies = 1 << 20, size_t max_electrons = 1 << 21, int device = 0)
      : quadtree(theta, epsilon, leaf_capacity, thread_capacity), cell_list(domain_w, domain_h, 1.0f),
        domain_width(domain_w), domain_height(domain_h) {
    psim_config cfg;
    psim_default_config(&cfg);
    cfg.theta = theta, cfg.epsilon = epsilon;
    cfg.leaf_capacity = (uint32_t)leaf_capacity, cfg.thread_capacity = (uint32_t)thread_capacity;
    const int rc = psim_create(device, max_bodies, max_electrons, &cfg, &ctx_);
    if (rc != PSIM_OK) throw Error(rc, "psim_create failed: no CUDA device or out of memory (there is no CPU fallback)");
    quadtree.sim_ = this;
    cell_list.sim_ = this;
  }
  ~Simulation() { psim_destroy(ctx_); }
  Simulation(const Simulation&) = delete;
  Simulation& operator=(const Simulation&) = delete;

  psim_ctx* raw() const { return ctx_; }
  void check(int rc) const {
    if (rc != PSIM_OK) throw Error(rc, psim_last_error(ctx_));
  }

  // host Vec<Body> -> device (all hot fields + electrons)
  void upload() {
    const size_t n = bodies.size();
    std::vector<float> pos(2 * n), z(n), vel(2 * n), vz(n), mass(n), radius(n), charge(n);
    std::vector<uint8_t> sp(n);
    std::vector<uint32_t> eb;
    std::vector<float> er, ev;
    for (size_t i = 0; i < n; ++i) {
      const Body& b = bodies[i];
      pos[2 * i] = b.pos.x, pos[2 * i + 1] = b.pos.y, z[i] = b.z;
      vel[2 * i] = b.vel.x, vel[2 * i + 1] = b.vel.y, vz[i] = b.vz;
      mass[i] = b.mass, radius[i] = b.radius, charge[i] = b.charge, sp[i] = (uint8_t)b.species;
      for (const Electron& e : b.electrons) {
        eb.push_back((uint32_t)i);
        er.push_back(e.rel_pos.x), er.push_back(e.rel_pos.y);
        ev.push_back(e.vel.x), ev.push_back(e.vel.y);
      }
    }
    check(psim_upload_bodies(ctx_, n, pos.data(), z.data(), vel.data(), vz.data(), mass.data(), radius.data(),
                             charge.data(), sp.data()));
    if (!eb.empty()) check(psim_upload_electrons(ctx_, eb.size(), eb.data(), er.data(), ev.data()));
    uploaded_ = true;
  }
  void ensure_uploaded() {
    if (!uploaded_) upload();
  }
  void invalidate() { uploaded_ = false; }  // call after editing `bodies` on the host

  // reorder the host Vec<Body> the way the reference's in-place partition does (quadtree.rs:56-63)
  void apply_permutation() {
    const size_t n = bodies.size();
    if (!n) return;
    std::vector<uint32_t> perm(n);
    check(psim_get_permutation(ctx_, perm.data()));
    std::vector<Body> out(n);
    for (size_t i = 0; i < n; ++i) out[i] = std::move(bodies[perm[i]]);
    bodies.swap(out);
  }
  void download_fields(bool pos_vel, bool acc, bool e_field) {
    const size_t n = bodies.size();
    if (!n) return;
    std::vector<float> p(2 * n), v(2 * n), zz(n), vzz(n), a(2 * n), e(2 * n);
    check(psim_download_bodies(ctx_, pos_vel ? p.data() : nullptr, pos_vel ? zz.data() : nullptr,
                               pos_vel ? v.data() : nullptr, pos_vel ? vzz.data() : nullptr,
                               acc ? a.data() : nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                               e_field ? e.data() : nullptr, nullptr));
    for (size_t i = 0; i < n; ++i) {
      Body& b = bodies[i];
      if (pos_vel) b.pos = {p[2 * i], p[2 * i + 1]}, b.vel = {v[2 * i], v[2 * i + 1]}, b.z = zz[i], b.vz = vzz[i];
      if (acc) b.acc = {a[2 * i], a[2 * i + 1]};
      if (e_field) b.e_field = {e[2 * i], e[2 * i + 1]};
    }
  }
  void download_electrons() {
    size_t m = 0;
    for (const Body& b : bodies) m += b.electrons.size();
    if (!m) return;
    std::vector<uint32_t> eb(m);
    std::vector<float> er(2 * m), ev(2 * m);
    check(psim_download_electrons(ctx_, eb.data(), er.data(), ev.data()));
    size_t k = 0;
    for (Body& b : bodies)
      for (Electron& e : b.electrons) {
        e.rel_pos = {er[2 * k], er[2 * k + 1]};
        e.vel = {ev[2 * k], ev[2 * k + 1]};
        ++k;
      }
  }

  bool use_cell_list() const {  // simulation.rs:1798-1802
    const float area = (2.0f * domain_width) * (2.0f * domain_height);
    return (float)bodies.size() / area > config.cell_list_density_threshold;
  }
  void reset_acc() {  // simulation.rs:1000-1003
    ensure_uploaded();
    check(psim_reset_acc(ctx_));
    for (Body& b : bodies) b.acc = {0, 0}, b.az = 0;
  }
  void iterate() {  // simulation.rs:1437-1486
    ensure_uploaded();
    check(psim_iterate(ctx_, dt, config.damping_base, domain_width, domain_height, domain_depth,
                       config.enable_out_of_plane ? 1 : 0));
    download_fields(true, false, false);
  }
  // simulation.rs:1893-1918 (+ body/types.rs:243-286, cell_list.rs:92-127); flags in the current body order
  std::vector<uint8_t> update_surrounded_flags(uint64_t frame, float radius_factor = 4.0f,
                                               uint64_t neighbor_threshold = 8) {
    ensure_uploaded();
    check(psim_update_surrounded_flags(ctx_, domain_width, domain_height, frame, radius_factor, neighbor_threshold));
    std::vector<uint8_t> flags(bodies.size());
    check(psim_get_surrounded(ctx_, flags.data(), nullptr, nullptr));
    return flags;
  }
  void enforce_metal_z_boundaries(float max_z) {  // simulation/out_of_plane.rs:140-254
    ensure_uploaded();
    check(psim_enforce_metal_z_boundaries(ctx_, max_z, domain_width, domain_height));
    download_fields(true, false, false);
  }
  void update_electrons() {  // simulation.rs:1186-1196 + body/electron.rs:19-46
    ensure_uploaded();
    check(psim_update_electrons(ctx_, background_e_field.x, background_e_field.y, dt, config.coulomb_constant));
    download_electrons();
  }
  // collision::collide, COLLISION_PASSES passes (simulation.rs:1025-1028, simulation/collision.rs:62-372); returns the
  // touching pairs of the last pass
  uint64_t collide(uint32_t num_passes = 7, float li_collision_softness = 0.8f, bool soft_collision_lithium_ion = true,
                   bool soft_collision_anion = false) {
    ensure_uploaded();
    uint64_t touching = 0;
    check(psim_collide(ctx_, domain_width, domain_height, domain_depth, num_passes, num_passes, li_collision_softness,
                       soft_collision_lithium_ion, soft_collision_anion, &touching));
    download_fields(true, false, false);
    return touching;
  }
  // the field part of the hopping predicate for a batch of (donor, acceptor) candidates
  // (simulation/electron_hopping.rs:283-329): alignment per candidate, local_field per donor
  std::vector<float> hop_alignment(const std::vector<uint32_t>& src, const std::vector<std::vector<uint32_t>>& candidates,
                                   float alignment_bias, std::vector<Vec2>* local_field = nullptr) {
    ensure_uploaded();
    std::vector<uint32_t> off(src.size() + 1, 0), dst;
    for (size_t i = 0; i < src.size(); ++i) {
      off[i + 1] = off[i] + (uint32_t)candidates[i].size();
      dst.insert(dst.end(), candidates[i].begin(), candidates[i].end());
    }
    std::vector<float> lf(2 * src.size() + 2), al(dst.size() + 1);
    check(psim_hop_alignment(ctx_, src.size(), src.data(), off.data(), dst.data(), config.coulomb_constant,
                             background_e_field.x, background_e_field.y, alignment_bias, lf.data(), al.data()));
    if (local_field) {
      local_field->resize(src.size());
      for (size_t i = 0; i < src.size(); ++i) (*local_field)[i] = Vec2{lf[2 * i], lf[2 * i + 1]};
    }
    al.resize(dst.size());
    return al;
  }
  psim_step_params step_params() const {
    psim_step_params p{};
    p.hw = domain_width, p.hh = domain_height, p.hd = domain_depth, p.dt = dt, p.damping_base = config.damping_base;
    p.k_e = config.coulomb_constant, p.bg_x = background_e_field.x, p.bg_y = background_e_field.y;
    p.density_threshold = config.cell_list_density_threshold;
    p.enable_out_of_plane = config.enable_out_of_plane, p.do_short_range = 1, p.do_electrons = 1, p.do_iterate = 1;
    p.do_polar = 1;  // forces::apply_polar_forces runs every step (simulation.rs:1007)
    return p;
  }
  // the hot path of Simulation::step (simulation.rs:1000-1196) without host round trips
  void step_hot_path() {
    ensure_uploaded();
    const psim_step_params p = step_params();
    check(psim_step(ctx_, &p));
  }
  // multi-GPU (one Simulation per rank, created with psim_shard_capacity(n, nranks) bodies): NCCL communicator from a
  // 128-byte id made by psim_comm_unique_id on rank 0, then the same step across the ranks
  void comm_init(const uint8_t id[128], uint32_t rank, uint32_t nranks) { check(psim_comm_init(ctx_, id, rank, nranks)); }
  void step_hot_path_sharded() {
    ensure_uploaded();
    const psim_step_params p = step_params();
    check(psim_step_sharded(ctx_, &p));
  }

 private:
  psim_ctx* ctx_ = nullptr;
  bool uploaded_ = false;
};

// ---- Quadtree / CellList method bodies ---------------------------------------------------------
inline void Quadtree::build(std::vector<Body>& bodies) {
  (void)bodies;
  sim_->ensure_uploaded();
  sim_->check(psim_build(sim_->raw(), PSIM_BUILD_CONTAINING, 0.0f, 0.0f));
  sim_->apply_permutation();
}
inline void Quadtree::build_with_domain(std::vector<Body>& bodies, float w, float h) {
  (void)bodies;
  sim_->ensure_uploaded();
  sim_->check(psim_build(sim_->raw(), PSIM_BUILD_DOMAIN, w, h));
  sim_->apply_permutation();
}
inline void Quadtree::field(std::vector<Body>& bodies, float k_e) {
  (void)bodies;
  sim_->check(psim_field(sim_->raw(), k_e, 0.0f, 0.0f, 0, nullptr, nullptr));
  sim_->download_fields(false, false, true);
}
inline Vec2 Quadtree::acc_pos(Vec2 pos, float q, float radius, const std::vector<Body>&, float k_e) const {
  float p[2] = {pos.x, pos.y}, out[2] = {0, 0};
  sim_->check(psim_acc_points(sim_->raw(), 1, p, &q, &radius, k_e, out));
  return {out[0], out[1]};
}
inline Vec2 Quadtree::field_at_point(const std::vector<Body>& bodies, Vec2 pos, float k_e) const {
  return acc_pos(pos, 1.0f, 0.0f, bodies, k_e);
}
inline std::vector<Vec2> Quadtree::field_at_points(const std::vector<Vec2>& pts, float k_e) const {
  std::vector<Vec2> out(pts.size());
  if (!pts.empty())
    sim_->check(psim_acc_points(sim_->raw(), pts.size(), &pts[0].x, nullptr, nullptr, k_e, &out[0].x));
  return out;
}
inline std::vector<Node> Quadtree::nodes() const {
  uint64_t count = 0;
  sim_->check(psim_download_nodes(sim_->raw(), nullptr, 0, &count));
  std::vector<Node> out(count);
  if (count) sim_->check(psim_download_nodes(sim_->raw(), out.data(), count, &count));
  return out;
}
namespace detail {
inline std::vector<size_t> neighbors(const Simulation& s, size_t i, float cutoff, bool metals) {
  uint32_t idx = (uint32_t)i, off[2] = {0, 0};
  uint64_t total = 0;
  s.check(psim_neighbors_within(s.raw(), 1, &idx, cutoff, metals, off, nullptr, 0, &total));
  std::vector<uint32_t> ind(total ? total : 1);
  if (total) s.check(psim_neighbors_within(s.raw(), 1, &idx, cutoff, metals, off, ind.data(), total, &total));
  return std::vector<size_t>(ind.begin(), ind.begin() + total);
}
}  // namespace detail
inline std::vector<size_t> Quadtree::find_neighbors_within(const std::vector<Body>&, size_t i, float cutoff) const {
  return detail::neighbors(*sim_, i, cutoff, false);
}
inline void CellList::rebuild(const std::vector<Body>&) {
  sim_->ensure_uploaded();
  sim_->check(psim_cell_build(sim_->raw(), domain_width, domain_height, cell_size));
}
inline std::vector<size_t> CellList::find_neighbors_within(const std::vector<Body>&, size_t i, float cutoff) const {
  return detail::neighbors(*sim_, i, cutoff, false);
}
inline size_t CellList::metal_neighbor_count(const std::vector<Body>&, size_t i, float cutoff) const {
  return detail::neighbors(*sim_, i, cutoff, true).size();
}

// ---- src/simulation/forces.rs --------------------------------------------------------------------
namespace forces {
inline void prepare_spatial_structures(Simulation& sim) {  // forces.rs:14-25
  sim.ensure_uploaded();
  sim.check(psim_prepare_spatial_structures(sim.raw(), sim.domain_width, sim.domain_height,
                                            sim.config.cell_list_density_threshold));
  sim.apply_permutation();
}
inline void attract(Simulation& sim) {  // forces.rs:33-44
  sim.check(psim_field(sim.raw(), sim.config.coulomb_constant, sim.background_e_field.x, sim.background_e_field.y,
                       1, nullptr, nullptr));
  sim.download_fields(false, true, true);
}
inline void apply_polar_forces(Simulation& sim, int dipole_model = 1) {  // forces.rs:52-175
  sim.check(psim_apply_polar_forces(sim.raw(), sim.config.coulomb_constant, dipole_model));
  sim.download_fields(false, true, false);
}
inline void apply_lj_forces(Simulation& sim) {  // forces.rs:182-231
  sim.check(psim_short_range(sim.raw(), PSIM_SR_LJ));
  sim.download_fields(false, true, false);
}
inline void apply_repulsive_forces(Simulation& sim) {  // forces.rs:250-289
  sim.check(psim_short_range(sim.raw(), PSIM_SR_REPULSION));
  sim.download_fields(false, true, false);
}
inline void apply_stack_pressure(Simulation& sim) {  // forces.rs:294-321
  sim.check(psim_short_range(sim.raw(), PSIM_SR_STACK_PRESSURE));
  sim.download_fields(false, true, false);
}
}  // namespace forces

}  // namespace psim
