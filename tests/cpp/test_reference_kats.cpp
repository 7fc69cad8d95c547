// The reference's own tests for the path, written against the C++ host mirror (include/psim_b200.hpp)
// so they read like the originals:
//   src/quadtree/tests.rs:9-78     test_quadtree_field_centered_on_body
//   src/quadtree/tests.rs:80-138   overlapping_particles_produce_finite_force
//   src/body/tests/anion.rs:32-56  Quadtree::new(0.5, 0.01, 1, 1) builds on a single body
// plus one pass of the force phase in Simulation::step's order.  Needs a CUDA device.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "psim_b200.hpp"

using namespace psim;
static const float COULOMB_CONSTANT = 0.138935f;  // units.rs:32-34

#define ASSERT(cond, ...)                         \
  do {                                            \
    if (!(cond)) {                                \
      std::fprintf(stderr, "FAILED %s:%d: ", __FILE__, __LINE__); \
      std::fprintf(stderr, __VA_ARGS__);          \
      std::fprintf(stderr, "\n");                 \
      std::exit(1);                               \
    }                                             \
  } while (0)

static float mag(Vec2 v) { return std::sqrt(v.x * v.x + v.y * v.y); }

static void test_quadtree_field_centered_on_body() {
  Simulation sim(10.0f, 10.0f, /*theta*/ 0.5f, /*epsilon*/ 1e-6f, /*leaf_capacity*/ 8, /*thread_capacity*/ 32, 16, 16);
  Body body;
  body.mass = 1.0f, body.radius = 1.0f, body.charge = 1.0f, body.species = Species::LithiumIon;
  sim.bodies = {body};
  sim.quadtree.build(sim.bodies);
  const Vec2 test_positions[4] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
  float magnitudes[4];
  for (int k = 0; k < 4; ++k) {
    const Vec2 pos = test_positions[k];
    const Vec2 field = sim.quadtree.acc_pos(pos, 1.0f, 0.0f, sim.bodies, COULOMB_CONSTANT);
    const float dot = (field.x * pos.x + field.y * pos.y) / (mag(field) * mag(pos));
    ASSERT(std::fabs(dot - 1.0f) < 1e-5f, "Field at (%g, %g) not pointing radially out: dot=%g", pos.x, pos.y, dot);
    magnitudes[k] = mag(field);
  }
  const float avg = (magnitudes[0] + magnitudes[1] + magnitudes[2] + magnitudes[3]) / 4.0f;
  for (int k = 0; k < 4; ++k)
    ASSERT(std::fabs(magnitudes[k] - avg) < 1e-5f, "Field magnitude at direction %d differs: %g vs avg %g", k, magnitudes[k], avg);
}

static void overlapping_particles_produce_finite_force() {
  Simulation sim(10.0f, 10.0f, 1.0f, 2.0f, 8, 32, 16, 16);
  Body a, b;
  a.mass = b.mass = 1.0f, a.radius = b.radius = 1.0f;
  a.charge = 1.0f, b.charge = -1.0f, b.pos = {0.5f, 0.0f}, b.id = 1;
  sim.bodies = {a, b};
  sim.quadtree.build(sim.bodies);
  const Vec2 field = sim.quadtree.acc_pos(sim.bodies[0].pos, sim.bodies[0].charge, sim.bodies[0].radius, sim.bodies, COULOMB_CONSTANT);
  ASSERT(std::isfinite(field.x) && std::isfinite(field.y), "Field should be finite for overlapping bodies");
}

static void degenerate_parameters_build() {
  Simulation sim(10.0f, 10.0f, 0.5f, 0.01f, 1, 1, 16, 16);
  Body anion;
  anion.pos = {3, 4}, anion.mass = 145.0f, anion.radius = 2.0f, anion.charge = -1.0f, anion.species = Species::ElectrolyteAnion;
  sim.bodies = {anion};
  sim.quadtree.build(sim.bodies);
  const std::vector<Node> nodes = sim.quadtree.nodes();
  ASSERT(nodes.size() == 1 && nodes[0].children == 0, "single body: one leaf");
}

static void force_phase_in_step_order() {
  const int n = 4000;
  const float hw = 126.5f, hh = 126.5f;
  Simulation sim(hw, hh, 1.0f, 2.0f, 1, 1024, n, 2 * n);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f; };
  for (int i = 0; i < n; ++i) {
    Body b;
    b.pos = {(rnd() - 0.5f) * 2 * hw, (rnd() - 0.5f) * 2 * hh};
    b.id = (uint64_t)i;
    const int kind = i % 8;
    if (kind == 0) b.species = Species::LithiumIon, b.charge = 1, b.radius = 0.76f, b.mass = 6.94f;
    else if (kind == 1) b.species = Species::ElectrolyteAnion, b.charge = -1, b.radius = 2.0f, b.mass = 145.0f, b.electrons = {{{0.1f, 0.0f}, {0, 0}}};
    else if (kind == 2) b.species = Species::LithiumMetal, b.charge = 0, b.radius = 1.52f, b.mass = 6.94f;
    else b.species = Species::EC, b.charge = 0, b.radius = 2.5f, b.mass = 88.06f, b.electrons = {{{0.0f, 0.2f}, {0, 0}}};
    sim.bodies.push_back(b);
  }
  sim.reset_acc();
  forces::prepare_spatial_structures(sim);
  forces::attract(sim);
  forces::apply_polar_forces(sim);
  forces::apply_lj_forces(sim);
  forces::apply_repulsive_forces(sim);
  forces::apply_stack_pressure(sim);
  sim.iterate();
  sim.quadtree.build_with_domain(sim.bodies, hw, hh);
  sim.update_electrons();
  double e2 = 0;
  for (const Body& b : sim.bodies) {
    ASSERT(std::isfinite(b.pos.x) && std::isfinite(b.acc.x) && std::isfinite(b.e_field.y), "non-finite state");
    ASSERT(std::fabs(b.pos.x) <= hw && std::fabs(b.pos.y) <= hh, "body left the domain");
    e2 += (double)b.e_field.x * b.e_field.x + (double)b.e_field.y * b.e_field.y;
  }
  ASSERT(e2 > 0, "field is identically zero");
  sim.cell_list.cell_size = 11.88f;
  sim.cell_list.rebuild(sim.bodies);  // the build above re-ordered the bodies: indices are stale until a rebuild
  const auto nb = sim.cell_list.find_neighbors_within(sim.bodies, 0, 11.88f);
  for (size_t j : nb) {
    const float dx = sim.bodies[j].pos.x - sim.bodies[0].pos.x, dy = sim.bodies[j].pos.y - sim.bodies[0].pos.y;
    ASSERT(j != 0 && dx * dx + dy * dy < 11.88f * 11.88f, "neighbour outside the cutoff");
  }
  sim.invalidate();
  sim.step_hot_path();
}

int main() {
  try {
    test_quadtree_field_centered_on_body();
    overlapping_particles_produce_finite_force();
    degenerate_parameters_build();
    force_phase_in_step_order();
  } catch (const Error& e) {
    std::fprintf(stderr, "psim error %d: %s\n", e.code, e.what());
    return 2;
  }
  std::puts("cpp reference KATs: all passed");
  return 0;
}
