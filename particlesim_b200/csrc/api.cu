// api.cu — the C ABI of libpsim_b200.so (include/psim_b200.h): context, device arenas and the
// launch sequences.  All compute is in the sm_100a kernels of sort.cuh / tree.cuh / traverse.cuh /
// cells.cuh; nothing here computes on the host and there is no CPU fallback.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <algorithm>
#include <vector>

#include "../../include/psim_b200.h"
#include "cells.cuh"
#include "collide.cuh"
#include "comm.h"
#include "hopping.cuh"
#include "let.cuh"
#include "polar.cuh"
#include "sort.cuh"
#include "traverse.cuh"
#include "neighbors.cuh"
#include "shard.cuh"
#include "strict.cuh"
#include "tree.cuh"

using namespace psim;

static_assert(sizeof(psim_species) == sizeof(SpeciesRow), "species row layout");
static_assert(sizeof(psim_node) == sizeof(PsimNodeOut) && sizeof(psim_node) == 64, "node layout");

namespace {

constexpr int kTreePasses = 4;  // radix passes over the upper key word (sort.cuh, two-tier key sort)

struct BodyArrays {
  float4* pqr;      // {x, y, charge, radius}
  float4* velz;     // {vx, vy, z, vz}
  float4* accm;     // {ax, ay, az, mass}
  float2* efield;
  uint8_t* species;
  uint32_t* orig;
  uint8_t* ecount;  // electrons per body
};

}  // namespace

struct psim_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = 0;
  psim_config cfg;
  std::string err;
  uint64_t cap_bodies = 0, cap_elec = 0;
  uint32_t n = 0, m = 0;
  uint32_t node_cap = 0;

  BodyArrays b[2] = {};
  int cur = 0;
  // electrons grouped by body (body order), double buffered for the regroup after a build
  uint32_t* ebody[2] = {nullptr, nullptr};
  float2* erel[2] = {nullptr, nullptr};
  float2* evel[2] = {nullptr, nullptr};
  int ecur = 0;
  uint32_t* eoff[2] = {nullptr, nullptr};  // exclusive offsets per body, n entries
  float2* epts = nullptr;
  float2* efld = nullptr;

  // sort
  uint64_t* keys[2] = {nullptr, nullptr};  // [0] keys in body order (scratch after the sort), [1] sorted
  uint32_t* khi[2] = {nullptr, nullptr};   // upper key words, ping-pong buffers of the radix passes
  uint32_t* idx[2] = {nullptr, nullptr};
  uint32_t* long_runs = nullptr;           // [0] count, [1..] heads of long equal-upper-word runs
  uint32_t* keys_idx_all = nullptr;        // sharded build: the all-gathered sorted order
  uint32_t* ckeys[2] = {nullptr, nullptr};
  uint32_t* cidx[2] = {nullptr, nullptr};
  SortScratch sc = {};
  SortPlan* tree_plan = nullptr;
  SortPlan* cell_plan = nullptr;

  // tree
  TreeMeta* meta = nullptr;
  TreeMeta meta_h = {};
  uint16_t* le = nullptr;
  uint32_t* nodebase = nullptr;
  uint32_t* scan_partials = nullptr;
  uint32_t* irank = nullptr;
  float4* bounds_partial = nullptr;
  TreeArrays t = {};
  float4* travA = nullptr;   // traversal arrays: charged nodes only, pre-order
  uint4* travB = nullptr;
  uint32_t* trav_rank = nullptr;
  uint32_t* trav_count = nullptr;
  uint32_t* perm = nullptr;
  uint32_t* inv = nullptr;
  bool tree_valid = false;
  bool perm_valid = false;
  StrictArrays strict = {};  // psim_config.strict_centres scratch, allocated by the first strict build
  bool strict_ready = false;
  SurroundState surround = {nullptr, nullptr, nullptr};  // by original body id

  // sharded build (shard.cuh): allocated by psim_shard_init
  struct Shard {
    bool on = false;
    bool tree_is_sharded = false;  // the node array holds this rank's piece only
    bool tree_is_let = false;      // ... and the traversal arrays only what this rank's own targets can reach
    uint32_t rank = 0, world = 1;
    int phase = 0, mode = 0;
    uint32_t halo = 0, n_local = 0, hl = 0, L = 0, s_lo = 0;
    uint32_t* binhist = nullptr;
    uint16_t* bins = nullptr;
    uint32_t* binprefix = nullptr;
    uint32_t* nb_bin = nullptr;
    uint32_t* trav_bin = nullptr;
    uint64_t* lkeys = nullptr;
    unsigned long long* xbuf = nullptr;  // kBins + kMaxRanks words (tables, all-reduced)
    TopRec* heap = nullptr;              // kTopSlots records (all-reduced)
    ShardPlan* plan = nullptr;
    ShardMeta* meta = nullptr;
    ShardPlan plan_h = {};
    ShardMeta meta_h = {};
  } sh;

  // multi-GPU communicator (psim_comm_init): NCCL on the context's stream, plus a side stream for the exchanges
  // that run beside independent device work
  struct Comm {
    bool on = false;
    ncclComm_t comm = nullptr;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_main = nullptr, ev_side = nullptr, ev_vel = nullptr;
    void* stage = nullptr;  // padded all-gather staging
    size_t stage_bytes = 0;
    bool vel_pending = false;
    // locally essential traversal tree (let.cuh)
    bool let = true, let_poison = false;
    LetRegions* regions = nullptr;
    LetRec *send = nullptr, *recv = nullptr;
    uint32_t *cnt = nullptr, *cnt_all = nullptr;
    uint32_t cap_per_rank = 0;
    uint64_t let_sent = 0, let_full = 0;  // records sent by the last exchange / what the full all-gather would send
  } comm;

  // cells
  uint32_t *cell_start = nullptr, *cell_end = nullptr, *order = nullptr, *body_cell = nullptr;
  float4* cpos = nullptr;
  uint32_t* cell_off = nullptr;      // monotone per-cell offsets (ncells + 1), built on demand
  bool cell_off_valid = false;
  uint32_t species_present = 0;      // bit s: some uploaded body has species s
  float4* polarB = nullptr;          // cell-ordered {species | flags, index, electron rel_pos}
  uint32_t* polar_cutoff = nullptr;  // max 3 * radius over polar bodies with an electron (float bits)
  uint64_t cell_cap = 0;
  GridDims grid = {0, 0, 1.0f, 0.0f, 0.0f};
  int cell_passes = 0;
  bool grid_valid = false;

  // species
  SpeciesRow* table_d = nullptr;
  SpeciesRow table_h[kMaxSpecies];
  uint32_t nspecies = 21;

  // staging for host <-> device packing and point queries (grow only)
  void* stage = nullptr;
  size_t stage_bytes = 0;
  void* qstage = nullptr;
  size_t qstage_bytes = 0;

  unsigned long long* step_counter = nullptr;
  unsigned int* grid_barrier = nullptr;  // aggregate_levels_kernel
  uint64_t launches = 0;
  // multi-GPU: the slice of bodies / electrons this rank computes (default: everything)
  bool tgt_set = false, etgt_set = false;
  uint32_t tgt_first = 0, tgt_count = 0, e_first = 0, e_count = 0;
  cudaEvent_t ev[9] = {};
  bool ev_ok = false, ev_recorded = false;
  // psim_step_host: copy streams, their events, a private staging area and the plan of the call in flight
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  // psim_step: work that the tree build does not wait for (electron regroup, cell list) runs beside it on a side stream
  struct Overlap {
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    uint32_t* partials = nullptr;  // scan scratch of the side stream
    bool active = false;           // set by step_async around its builds
    bool forked = false;           // this build has put work on the side stream
    bool pending = false;          // the main stream has not waited for the side stream yet
    float cell_hw = 0.f, cell_hh = 0.f, cell_size = 0.f;  // > 0: rebuild the cell list beside the build
  } ov;
  cudaEvent_t ev_start = nullptr, ev_q = nullptr, ev_vel = nullptr, ev_mid = nullptr, ev_out = nullptr;
  void* hstage = nullptr;
  size_t hstage_bytes = 0;
  // row of the caller's arrays (the order of the previous call's outputs) that device row i holds;
  // valid from a call whose electron pass re-sorted the bodies until the next build
  uint32_t* host_map = nullptr;
  bool host_map_valid = false;
  struct HostedStep {
    bool active = false;
    const float* late_q = nullptr;      // staged charges, applied before the first gather
    const float2* late_vel = nullptr;   // staged velocities, applied before the integrator
    const float* vel_src = nullptr;     // host velocities whose copy has not been queued yet
    const uint32_t* map = nullptr;      // host_map while the inputs are applied (null: identity)
    float2 *s_pos = nullptr, *s_vel = nullptr;  // device staging of the unpacked outputs
    float *out_pos = nullptr, *out_vel = nullptr, *out_ef = nullptr;
    uint32_t* out_orig = nullptr;
  } hosted;
};

namespace {

int32_t fail(psim_ctx* c, int32_t code, const char* what, cudaError_t e = cudaSuccess) {
  if (c) {
    c->err = what;
    if (e != cudaSuccess) {
      c->err += ": ";
      c->err += cudaGetErrorString(e);
    }
  }
  return code;
}

#define CK(call)                                                        \
  do {                                                                  \
    cudaError_t _e = (call);                                            \
    if (_e != cudaSuccess) return fail(ctx, PSIM_E_CUDA, #call, _e);    \
  } while (0)

#define LAUNCHED(ctx) ((ctx)->launches++)

// Every entry point runs on its context's device, whatever device the calling thread had current (a host
// thread may own one context per GPU), and leaves the caller's current device as it found it.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) cudaSetDevice(device); else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

template <typename T>
cudaError_t dalloc(T** p, size_t count) {
  if (count == 0) count = 1;
  return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
}

int grid_for(const psim_ctx* c, uint64_t work, int threads, int per_sm = 8) {
  uint64_t need = (work + threads - 1) / threads;
  uint64_t cap = (uint64_t)c->sm_count * per_sm;  // multiple of the SM count
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- pack / unpack between the C ABI's plain arrays and the device SoA -------------------------
struct RawBodies {  // device staging pointers (nullptr => default)
  const float2* pos;
  const float* z;
  const float2* vel;
  const float* vz;
  const float* mass;
  const float* radius;
  const float* charge;
  const uint8_t* species;
};

__global__ void __launch_bounds__(256) pack_bodies_kernel(RawBodies r, uint32_t n, BodyArrays b) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float2 p = r.pos[i];
    const float2 v = r.vel ? r.vel[i] : make_float2(0.f, 0.f);
    b.pqr[i] = make_float4(p.x, p.y, r.charge ? r.charge[i] : 0.f, r.radius ? r.radius[i] : 0.f);
    b.velz[i] = make_float4(v.x, v.y, r.z ? r.z[i] : 0.f, r.vz ? r.vz[i] : 0.f);
    b.accm[i] = make_float4(0.f, 0.f, 0.f, r.mass ? r.mass[i] : 1.f);
    b.efield[i] = make_float2(0.f, 0.f);
    b.species[i] = r.species ? r.species[i] : (uint8_t)0;
    b.orig[i] = i;
    b.ecount[i] = 0;
  }
}

__global__ void __launch_bounds__(256) set_positions_kernel(const float2* pos, uint32_t n, float4* pqr) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 p = pqr[i];
    p.x = pos[i].x, p.y = pos[i].y;
    pqr[i] = p;
  }
}
__global__ void __launch_bounds__(256) set_charges_kernel(const float* q, uint32_t n, float4* pqr) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 p = pqr[i];
    p.z = q[i];
    pqr[i] = p;
  }
}

__global__ void __launch_bounds__(256)
    update_state_kernel(const float2* pos, const float2* vel, const float* q, uint32_t n, float4* pqr, float4* velz) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 p = pqr[i];
    p.x = pos[i].x, p.y = pos[i].y;
    if (q) p.z = q[i];
    pqr[i] = p;
    if (vel) {
      float4 v = velz[i];
      v.x = vel[i].x, v.y = vel[i].y;
      velz[i] = v;
    }
  }
}

// psim_step_host: the pieces of a pipelined state refresh.  `map` (may be null) is the caller's row of
// each device row; velocities arrive after the first gather, so they also go through its permutation.
__global__ void __launch_bounds__(256)
    hosted_positions_kernel(const float2* pos, const uint32_t* map, uint32_t n, float4* pqr) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float2 p = pos[map ? map[i] : i];
    pqr[i].x = p.x, pqr[i].y = p.y;
  }
}
__global__ void __launch_bounds__(256)
    hosted_charges_kernel(const float* q, const uint32_t* map, uint32_t n, float4* pqr) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) pqr[i].z = q[map ? map[i] : i];
}
__global__ void __launch_bounds__(256)
    hosted_velocities_kernel(const float2* vel, const uint32_t* perm, const uint32_t* map, uint32_t n, float4* velz) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t s = perm[i];
    const float2 v = vel[map ? map[s] : s];
    velz[i].x = v.x, velz[i].y = v.y;
  }
}
__global__ void __launch_bounds__(256)
    hosted_unpack_kernel(const float4* pqr, const float4* velz, uint32_t n, float2* pos, float2* vel) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (pos) pos[i] = make_float2(pqr[i].x, pqr[i].y);
    if (vel) vel[i] = make_float2(velz[i].x, velz[i].y);
  }
}

struct RawOut {
  float2* pos;
  float* z;
  float2* vel;
  float* vz;
  float2* acc;
  float* az;
  float* mass;
  float* radius;
  float* charge;
};
__global__ void __launch_bounds__(256) unpack_bodies_kernel(BodyArrays b, uint32_t n, RawOut o) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float4 p = b.pqr[i], v = b.velz[i], a = b.accm[i];
    if (o.pos) o.pos[i] = make_float2(p.x, p.y);
    if (o.z) o.z[i] = v.z;
    if (o.vel) o.vel[i] = make_float2(v.x, v.y);
    if (o.vz) o.vz[i] = v.w;
    if (o.acc) o.acc[i] = make_float2(a.x, a.y);
    if (o.az) o.az[i] = a.z;
    if (o.mass) o.mass[i] = a.w;
    if (o.radius) o.radius[i] = p.w;
    if (o.charge) o.charge[i] = p.z;
  }
}

// bodies into tree order (the reference's in-place partition side effect, quadtree.rs:56-63)
__global__ void __launch_bounds__(256)
    gather_bodies_kernel(const uint32_t* __restrict__ idx0, const uint32_t* __restrict__ idx1,
                         const SortPlan* __restrict__ plan, int npass, uint32_t n, BodyArrays in,
                         BodyArrays out, uint32_t* __restrict__ perm, uint32_t* __restrict__ inv) {
  const uint32_t* __restrict__ idx = plan->src[npass] ? idx1 : idx0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t s = idx[i];
    out.pqr[i] = in.pqr[s];
    out.velz[i] = in.velz[s];
    out.accm[i] = in.accm[s];
    out.efield[i] = in.efield[s];
    out.species[i] = in.species[s];
    out.orig[i] = in.orig[s];
    out.ecount[i] = in.ecount[s];
    perm[i] = s;
    inv[s] = i;
  }
}

struct EcountFn {
  const uint8_t* ecount;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return ecount[i]; }
};

__global__ void __launch_bounds__(256)
    regroup_electrons_kernel(const uint32_t* __restrict__ perm, const uint8_t* __restrict__ ecount_new,
                             const uint32_t* __restrict__ eoff_old, const uint32_t* __restrict__ eoff_new,
                             uint32_t n, const float2* __restrict__ rel_in,
                             const float2* __restrict__ vel_in, uint32_t* __restrict__ body_out,
                             float2* __restrict__ rel_out, float2* __restrict__ vel_out) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t cnt = ecount_new[i];
    if (!cnt) continue;
    const uint32_t src = eoff_old[perm[i]], dst = eoff_new[i];
    for (uint32_t s = 0; s < cnt; ++s) {
      body_out[dst + s] = i;
      rel_out[dst + s] = rel_in[src + s];
      vel_out[dst + s] = vel_in[src + s];
    }
  }
}

__global__ void __launch_bounds__(256) reset_acc_kernel(float4* accm, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 a = accm[i];
    a.x = 0.f, a.y = 0.f, a.z = 0.f;
    accm[i] = a;
  }
}

struct U32Fn {
  const uint32_t* v;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return v[i]; }
};

int32_t ensure_stage(psim_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->stage_bytes) return PSIM_OK;
  if (ctx->stage) cudaFree(ctx->stage);
  ctx->stage = nullptr;
  ctx->stage_bytes = 0;
  cudaError_t e = cudaMalloc(&ctx->stage, bytes);
  if (e != cudaSuccess) return fail(ctx, PSIM_E_OOM, "staging buffer", e);
  ctx->stage_bytes = bytes;
  return PSIM_OK;
}
int32_t ensure_qstage(psim_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->qstage_bytes) return PSIM_OK;
  if (ctx->qstage) cudaFree(ctx->qstage);
  ctx->qstage = nullptr;
  ctx->qstage_bytes = 0;
  cudaError_t e = cudaMalloc(&ctx->qstage, bytes);
  if (e != cudaSuccess) return fail(ctx, PSIM_E_OOM, "query staging buffer", e);
  ctx->qstage_bytes = bytes;
  return PSIM_OK;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

void body_range(const psim_ctx* ctx, uint32_t& first, uint32_t& count) {
  first = 0, count = ctx->n;
  if (ctx->tgt_set) {
    first = ctx->tgt_first < ctx->n ? ctx->tgt_first : ctx->n;
    count = ctx->tgt_count < ctx->n - first ? ctx->tgt_count : ctx->n - first;
  }
}
void electron_range(const psim_ctx* ctx, uint32_t& first, uint32_t& count) {
  first = 0, count = ctx->m;
  if (ctx->etgt_set) {
    first = ctx->e_first < ctx->m ? ctx->e_first : ctx->m;
    count = ctx->e_count < ctx->m - first ? ctx->e_count : ctx->m - first;
  }
}

FieldParams field_params(const psim_ctx* ctx, float k_e, float bg_x, float bg_y) {
  FieldParams P;
  P.t_sq = ctx->cfg.theta * ctx->cfg.theta;      // Quadtree::new, quadtree.rs:26-27
  P.e_sq = ctx->cfg.epsilon * ctx->cfg.epsilon;
  P.k_e = k_e;
  P.bg_x = bg_x;
  P.bg_y = bg_y;
  P.inv_theta = 1.0f / ctx->cfg.theta;
  const char* off = getenv("PSIM_ONE_MUFU");
  P.one_mufu = ctx->cfg.parity_mode == 0 && ctx->cfg.theta <= 1.0f && ctx->cfg.theta > 0.0f && ctx->cfg.epsilon >= 1e-2f &&
               ctx->cfg.epsilon <= 1e3f && !(off && off[0] == '0');
  return P;
}

// species.rs:412-479
float max_lj_cutoff(const psim_ctx* ctx) {
  float m = 0.0f;
  for (uint32_t i = 0; i < ctx->nspecies; ++i)
    if (ctx->table_h[i].lj_enabled) m = fmaxf(m, ctx->table_h[i].lj_cutoff * ctx->table_h[i].lj_sigma);
  return m;
}
float max_repulsion_cutoff(const psim_ctx* ctx) {
  float m = 0.0f;
  for (uint32_t i = 0; i < ctx->nspecies; ++i)
    if (ctx->table_h[i].repulsion_enabled) m = fmaxf(m, ctx->table_h[i].repulsion_cutoff);
  return m;
}

// Cell size the fused steps bin at.  The reference sizes its grid at max(3 x LJ cutoff, repulsion, LJ) = 11.88 A
// (forces.rs:17-22); the PAIR SETS of the polar / LJ / repulsion passes do not depend on the cell size, only the
// number of candidates scanned does, so the steps bin at the largest cutoff the passes actually use: the LJ and
// repulsion cutoffs and, with the polar pass, 3 x the radius of the polar species (EC / DMC, forces.rs:74).
float step_cell_size(const psim_ctx* ctx, bool do_polar) {
  float cell = fmaxf(max_repulsion_cutoff(ctx), max_lj_cutoff(ctx));
  if (do_polar) {
    float polar = 0.0f;
    for (uint32_t s = 4; s <= 5 && s < ctx->nspecies; ++s) polar = fmaxf(polar, 3.0f * ctx->table_h[s].radius);
    cell = fmaxf(cell, polar);
  }
  return cell;
}

void default_species(SpeciesRow* t) {
  // species.rs:26-408 with config.rs:40-48,106-126 and units.rs
  const double EV_TO_SIM =
      1.602176634e-19 / (1.66053906660e-27 * 1.0e-10 * 1.0e-10 / (1.0e-15 * 1.0e-15));
  const float LJ_EPS = (float)((double)0.0103f * EV_TO_SIM);
  struct Row {
    float mass, radius, damping;
    int lj;
    float eps, polar_offset, polar_charge, rep_k, rep_cut;
  };
  static const Row rows[21] = {
      {6.94f, 0.76f, 1.0f, 0, 0.0f, 0.0f, 1.0f, 5.0f, 2.0f},      // LithiumIon
      {6.94f, 1.52f, 0.01f, 1, 0.1f, 1.0f, 1.0f, 5.0f, 2.0f},     // LithiumMetal
      {1.0e6f, 1.52f, 0.1f, 1, 10.0f, 1.0f, 1.0f, 5.0f, 2.0f},    // FoilMetal
      {145.0f, 2.0f, 1.0f, 0, 0.0f, 0.3f, 1.0f, 5.0f, 2.0f},      // ElectrolyteAnion
      {88.06f, 2.5f, 1.0f, 0, 0.0f, 0.85f, 0.80f, 5.0f, 5.0f},    // EC
      {90.08f, 2.5f, 1.0f, 0, 0.0f, 0.60f, 0.20f, 5.0f, 5.0f},    // DMC
      {86.0f, 2.4f, 1.0f, 0, 0.0f, 0.85f, 0.80f, 5.0f, 5.0f},     // VC
      {107.0f, 2.5f, 0.8f, 0, 0.0f, 0.85f, 0.80f, 6.0f, 5.0f},    // FEC
      {104.0f, 2.6f, 1.0f, 0, 0.0f, 0.60f, 0.20f, 4.5f, 5.5f},    // EMC
      {840.0f, 4.5f, 0.2f, 1, -1.f, 0.20f, 0.05f, 5.0f, 2.0f},    // LLZO
      {865.0f, 4.7f, 0.2f, 1, -1.f, 0.20f, 0.06f, 5.0f, 2.0f},    // LLZT
      {340.0f, 4.2f, 0.25f, 1, -1.f, 0.22f, 0.04f, 5.0f, 2.0f},   // S40B
      {100.0f, 2.0f, 0.01f, 1, -1.f, 0.0f, 0.0f, 5.0f, 2.0f},     // SEI
      {72.0f, 1.7f, 0.01f, 1, -1.f, 0.0f, 0.0f, 5.0f, 2.0f},      // Graphite
      {72.0f, 1.8f, 0.01f, 1, -1.f, 0.0f, 0.0f, 5.0f, 2.0f},      // HardCarbon
      {60.0f, 2.0f, 0.01f, 1, -1.f, 0.0f, 0.0f, 5.0f, 2.0f},      // SiliconOxide
      {460.0f, 2.5f, 0.01f, 1, -1.f, 0.0f, 0.0f, 5.0f, 2.0f},     // LTO
      {158.0f, 2.2f, 0.01f, 1, -1.f, 0.0f, 0.0f, 5.0f, 2.0f},     // LFP
      {158.0f, 2.2f, 0.01f, 1, -1.f, 0.0f, 0.0f, 5.0f, 2.0f},     // LMFP
      {97.0f, 2.0f, 0.01f, 1, -1.f, 0.0f, 0.0f, 5.0f, 2.0f},      // NMC
      {97.0f, 2.0f, 0.01f, 1, -1.f, 0.0f, 0.0f, 5.0f, 2.0f},      // NCA
  };
  for (int i = 0; i < 21; ++i) {
    t[i].mass = rows[i].mass;
    t[i].radius = rows[i].radius;
    t[i].damping = rows[i].damping;
    t[i].lj_enabled = (uint32_t)rows[i].lj;
    t[i].lj_epsilon = rows[i].eps < 0.f ? LJ_EPS : rows[i].eps;
    t[i].lj_sigma = 1.80f;
    t[i].lj_cutoff = 2.2f;
    t[i].polar_offset = rows[i].polar_offset;
    t[i].polar_charge = rows[i].polar_charge;
    t[i].repulsion_enabled = 0;
    t[i].repulsion_strength = rows[i].rep_k;
    t[i].repulsion_cutoff = rows[i].rep_cut;
  }
}

int32_t fetch_meta(psim_ctx* ctx) {
  CK(cudaMemcpyAsync(&ctx->meta_h, ctx->meta, sizeof(TreeMeta), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PSIM_OK;
}

// ---- launch sequences (no host synchronisation inside) ------------------------------------------
// bodies (and their electrons) into the order given by the sorted payload; swaps the body buffers
int32_t gather_stage(psim_ctx* ctx, const uint32_t* idx0, const uint32_t* idx1, const SortPlan* plan, int npass) {
  const uint32_t n = ctx->n;
  cudaStream_t st = ctx->stream;
  BodyArrays& in = ctx->b[ctx->cur];
  BodyArrays& out = ctx->b[ctx->cur ^ 1];
  if (ctx->hosted.active && ctx->hosted.late_q) {  // psim_step_host: the charges arrived during the sort
    CK(cudaStreamWaitEvent(st, ctx->ev_q, 0));
    hosted_charges_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(ctx->hosted.late_q, ctx->hosted.map, n, in.pqr);
    LAUNCHED(ctx);
    ctx->hosted.late_q = nullptr;
  }
  if (ctx->hosted.active && ctx->hosted.vel_src) {
    // the velocities are not needed before the integrator: their copy starts here, so that it neither
    // shares the link with the positions and charges nor delays their completion event
    CK(cudaEventRecord(ctx->ev_start, st));
    CK(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_start, 0));
    CK(cudaMemcpyAsync(const_cast<float2*>(ctx->hosted.late_vel), ctx->hosted.vel_src, 8 * (size_t)n,
                       cudaMemcpyHostToDevice, ctx->copy_in));
    CK(cudaEventRecord(ctx->ev_vel, ctx->copy_in));
    ctx->hosted.vel_src = nullptr;
  }
  gather_bodies_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(
      idx0, idx1, plan, npass, n, in, out, ctx->perm, ctx->inv);
  LAUNCHED(ctx);
  ctx->cur ^= 1;
  BodyArrays& b = ctx->b[ctx->cur];
  cudaStream_t es = st;
  uint32_t* e_partials = ctx->scan_partials;
  ctx->ov.forked = false;
  if (ctx->ov.active) {  // from here on the side stream may read the permuted bodies
    CK(cudaEventRecord(ctx->ov.ev_fork, st));
    CK(cudaStreamWaitEvent(ctx->ov.side, ctx->ov.ev_fork, 0));
    es = ctx->ov.side, e_partials = ctx->ov.partials;
    ctx->ov.forked = true;
  }
  if (ctx->m > 0) {
    // electrons follow their bodies: new offsets from the permuted counts, then a grouped copy
    const int eo = ctx->ecur, en = ctx->ecur ^ 1;
    CK(exclusive_scan(EcountFn{b.ecount}, n, ctx->eoff[en], e_partials, nullptr, es));
    ctx->launches += 3;
    regroup_electrons_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, es>>>(
        ctx->perm, b.ecount, ctx->eoff[eo], ctx->eoff[en], n, ctx->erel[eo], ctx->evel[eo],
        ctx->ebody[en], ctx->erel[en], ctx->evel[en]);
    LAUNCHED(ctx);
    ctx->ecur = en;
  }
  return PSIM_OK;
}

// scratch of strict.cuh, sized for the context's body capacity (grow-only, first strict build)
int32_t ensure_strict(psim_ctx* ctx) {
  if (ctx->strict_ready) return PSIM_OK;
  StrictArrays& S = ctx->strict;
  const size_t nb = ctx->cap_bodies ? ctx->cap_bodies : 1;
  S.long_cap = (uint32_t)(nb / 32 + 1024);
  S.item_cap = (uint32_t)(nb / 16 + nb / 32 + 1024);
  S.blk_cap = (uint32_t)(nb / kStrictBlock + 2);
  bool ok = true;
  auto A = [&](auto** p, size_t cnt) {
    if (ok && dalloc(p, cnt) != cudaSuccess) ok = false;
  };
  S.cand_cap = (uint32_t)(nb / 16 + 4096);  // chains with a node of more than kStrictDirect bodies: <= 32 levels * nb / 64
  A(&S.qstat, 3), A(&S.qc, nb + 2);
  A(&S.cidx, nb + 1), A(&S.cw, nb), A(&S.chains, S.cand_cap), A(&S.cand, S.cand_cap), A(&S.hist, 96), A(&S.longs, S.long_cap), A(&S.counters, 8);
  A(&S.item_first, (size_t)S.long_cap + 1), A(&S.pblk, 3 * ((size_t)S.blk_cap + 1)), A(&S.fns, 3 * (size_t)S.item_cap);
  if (!ok) {
    cudaGetLastError();
    return fail(ctx, PSIM_E_OOM, "strict_centres scratch");
  }
  CK(cudaFuncSetAttribute(strict_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStreamSmem));
  CK(cudaFuncSetAttribute(strict_blockfn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStreamSmem));
  ctx->strict_ready = true;
  return PSIM_OK;
}

// psim_config.strict_centres: overwrite the centres of all charged internal nodes with the reference's
// serial f32 sums (strict.cuh); runs after the tree is complete and before the traversal compaction
// charged-body index and addends of the sorted bodies (needed from the emit kernel on); resets the stage's counters
int32_t strict_prepare(psim_ctx* ctx) {
  const int32_t rc = ensure_strict(ctx);
  if (rc) return rc;
  const uint32_t n = ctx->n;
  cudaStream_t st = ctx->stream;
  BodyArrays& b = ctx->b[ctx->cur];
  StrictArrays& S = ctx->strict;
  CK(exclusive_scan(ChargedBodyFn{b.pqr}, n, S.cidx, ctx->scan_partials, S.cidx + n, st));
  CK(cudaMemsetAsync(S.qstat, 0, 3 * sizeof(unsigned long long), st));
  strict_addends_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(b.pqr, n, S.cidx, S.cw, S.hist, S.counters, S.qstat);
  CK(cudaMemsetAsync(S.counters + 4, 0, sizeof(uint32_t), st));
  // integer prefix of the charges over the charged bodies (tree.cuh StrictEmit::qc); its length only exists on the device
  CK(exclusive_scan_dyn(ChargeIntFn{S.cw, S.cidx + n}, reinterpret_cast<const uint32_t*>(S.qstat + 2), n + 1, S.qc,
                        ctx->scan_partials, nullptr, st));
  ctx->launches += 7;
  return PSIM_OK;
}

int32_t strict_stage(psim_ctx* ctx, bool sharded = false) {
  const int32_t rc = ensure_strict(ctx);
  if (rc) return rc;
  const uint32_t n = ctx->n;
  cudaStream_t st = ctx->stream;
  BodyArrays& b = ctx->b[ctx->cur];
  StrictArrays& S = ctx->strict;
  if (sharded) {
    // the rank's own piece of the tree: its chains with a node of more than kStrictDirect bodies (the single-GPU emit
    // kernel reports them itself).  Body indices are global (the sorted bodies are replicated), node indices local.
    strict_candidates_shard_kernel<<<grid_for(ctx, ctx->sh.n_local, 256, 16), 256, 0, st>>>(
        ctx->sh.meta, ctx->le, ctx->nodebase, ctx->t, kStrictDirect, S.cand, S.counters + 4, S.cand_cap);
    LAUNCHED(ctx);
  }
  strict_chain_count_kernel<<<grid_for(ctx, n / 16 + 1, 256, 8), 256, 0, st>>>(S.cand, S.counters + 4, S.cand_cap, S.cidx,
                                                                               S.hist, S.counters);
  strict_chain_scatter_kernel<<<grid_for(ctx, n / 16 + 1, 1024, 8), 256, 0, st>>>(S.cand, S.counters + 4, S.cand_cap, 1024,
                                                                                  S.cidx, S.hist, S.chains);
  strict_chain_kernel<<<grid_for(ctx, n / 16 + 1, kStreamThreads, 3), kStreamThreads, kStreamSmem, st>>>(ctx->meta, ctx->t, S);
  strict_blocksum_kernel<<<grid_for(ctx, (uint64_t)n / 16 + 1, 256, 8), 256, 0, st>>>(S.cw, S.cidx + n, S.blk_cap, S.pblk);
  strict_long_setup_kernel<<<1, 1024, 0, st>>>(S.cidx + n, S);
  strict_blockfn_kernel<<<grid_for(ctx, (uint64_t)n / 16 + 1, kStreamThreads, 3), kStreamThreads, kStreamSmem, st>>>(S);
  strict_compose_kernel<<<ctx->sm_count * 4, 96, 0, st>>>(ctx->meta, ctx->t, S);
  strict_slow_kernel<<<grid_for(ctx, (uint64_t)n * 2, 128, 16), 128, 0, st>>>(ctx->meta, b.pqr, b.accm, ctx->t, S.counters);
  ctx->launches += 8;
  return PSIM_OK;
}

int32_t cell_build_async(psim_ctx* ctx, float hw, float hh, float cell_size);

// the main stream waits for what the last build left on the side stream (no-op otherwise)
int32_t overlap_join(psim_ctx* ctx) {
  if (ctx->ov.pending) {
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ov.ev_join, 0));
    ctx->ov.pending = false;
  }
  return PSIM_OK;
}
int32_t overlap_ready(psim_ctx* ctx) {
  if (ctx->ov.side) return PSIM_OK;
  bool ok = cudaStreamCreateWithFlags(&ctx->ov.side, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->ov.ev_fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->ov.ev_join, cudaEventDisableTiming) == cudaSuccess &&
            cudaMalloc(&ctx->ov.partials, ((size_t)scan_num_tiles((uint32_t)ctx->cap_bodies) + 2) * sizeof(uint32_t)) == cudaSuccess;
  if (!ok) return fail(ctx, PSIM_E_CUDA, "psim_step: side stream", cudaGetLastError());
  return PSIM_OK;
}

int32_t build_async(psim_ctx* ctx, int mode, float hw, float hh) {
  const uint32_t n = ctx->n;
  cudaStream_t st = ctx->stream;
  ctx->sh.tree_is_sharded = false;
  ctx->tree_valid = false;
  ctx->perm_valid = false;
  ctx->host_map_valid = false;
  ctx->grid_valid = false;  // the cell list indexes bodies by position in the array
  if (n == 0) {
    CK(cudaMemsetAsync(ctx->meta, 0, sizeof(TreeMeta), st));
    CK(cudaMemsetAsync(ctx->trav_count, 0, sizeof(uint32_t), st));
    ctx->tree_valid = true;
    return PSIM_OK;
  }
  BodyArrays& in = ctx->b[ctx->cur];
  int nb = 1;
  if (mode == PSIM_BUILD_CONTAINING) {
    nb = grid_for(ctx, n, 256, 4);
    bounds_partial_kernel<<<nb, 256, 0, st>>>(in.pqr, n, ctx->bounds_partial);
    LAUNCHED(ctx);
  }
  root_quad_kernel<<<1, 32, 0, st>>>(ctx->bounds_partial, nb, mode, hw, hh, n, ctx->meta);
  LAUNCHED(ctx);
  keygen_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(in.pqr, n, ctx->meta, ctx->keys[0], ctx->khi[0],
                                                           ctx->idx[0]);
  LAUNCHED(ctx);
  ctx->sc.plan = ctx->tree_plan;
  CK(onesweep_sort<uint32_t>(ctx->khi[0], ctx->khi[1], ctx->idx[0], ctx->idx[1], n, 0, kTreePasses,
                             ctx->sc, ctx->sm_count, st));
  ctx->launches += 2 + kTreePasses;
  gather_keys_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(ctx->keys[0], ctx->idx[0], ctx->idx[1], ctx->tree_plan,
                                                                kTreePasses, n, ctx->keys[1], ctx->long_runs);
  fix_runs_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(ctx->keys[1], ctx->idx[0], ctx->idx[1], ctx->tree_plan,
                                                             kTreePasses, n, ctx->long_runs, ctx->long_runs + 1);
  sort_long_runs_kernel<<<ctx->sm_count, 256, 0, st>>>(ctx->keys[1], ctx->idx[0], ctx->idx[1], ctx->tree_plan,
                                                       kTreePasses, n, ctx->long_runs, ctx->long_runs + 1,
                                                       ctx->keys[0]);
  ctx->launches += 3;
  {
    const int32_t rc = gather_stage(ctx, ctx->idx[0], ctx->idx[1], ctx->tree_plan, kTreePasses);
    if (rc) return rc;
  }
  if (ctx->ov.forked) {
    if (ctx->ov.cell_size > 0.0f) {  // the cell list only needs the sorted bodies
      ctx->stream = ctx->ov.side;
      const int32_t rc = cell_build_async(ctx, ctx->ov.cell_hw, ctx->ov.cell_hh, ctx->ov.cell_size);
      ctx->stream = st;
      if (rc) return rc;
    }
    CK(cudaEventRecord(ctx->ov.ev_join, ctx->ov.side));
    ctx->ov.pending = true;
  }
  BodyArrays& b = ctx->b[ctx->cur];
  const uint32_t c_eff = effective_capacity(ctx->cfg.leaf_capacity, ctx->cfg.thread_capacity);
  // the emit kernel's slabs: one CTA per `per_block` consecutive bodies
  // (small slabs: what the CTAs in flight write and then sum must still be in L2 - measured at 16 M bodies:
  // 256 -> 4.96 ms per build, 512 -> 4.76, 1024 -> 4.84, 2048 -> 5.20, 6757 (one slab per resident CTA) -> 5.6)
  const uint32_t per_block = 512u;
  const int emit_grid = (int)((n + per_block - 1) / per_block);
  tree_count_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, st>>>(
      ctx->keys[1], ctx->keys[1], ctx->tree_plan, kTreePasses, b.pqr, n, c_eff, per_block, ctx->meta, ctx->le,
      ctx->cfg.leaf_capacity, ctx->cfg.thread_capacity);
  LAUNCHED(ctx);
  CK(exclusive_scan(LeCountFn{ctx->le}, n, ctx->nodebase, ctx->scan_partials, &ctx->meta->num_nodes, st));
  ctx->launches += 3;
  level_scan_kernel<<<1, 32, 0, st>>>(ctx->meta, ctx->node_cap);
  LAUNCHED(ctx);
  StrictEmit se = {0, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr};
  if (ctx->cfg.strict_centres) {
    const int32_t rc = strict_prepare(ctx);
    if (rc) return rc;
    static const bool no_intq = getenv("PSIM_INTEGER_CHARGES") && getenv("PSIM_INTEGER_CHARGES")[0] == '0';
    se = StrictEmit{kStrictDirect, ctx->strict.cidx, ctx->strict.cw, ctx->strict.cand, ctx->strict.counters + 4,
                    ctx->strict.cand_cap, no_intq ? nullptr : ctx->strict.qstat, ctx->strict.qc};
  }
  {
    static const int minb = getenv("PSIM_EMIT_MINB") ? atoi(getenv("PSIM_EMIT_MINB")) : 12;
    auto kern = minb >= 12 ? tree_emit_kernel<12> : minb >= 10 ? tree_emit_kernel<10> : minb >= 8 ? tree_emit_kernel<8> : tree_emit_kernel<5>;
    kern<<<emit_grid, 128, 0, st>>>(ctx->keys[1], ctx->keys[1], ctx->tree_plan, kTreePasses, n, per_block, ctx->le,
                                    ctx->nodebase, b.pqr, b.accm, ctx->cfg.leaf_capacity, ctx->cfg.thread_capacity,
                                    ctx->meta, ctx->t, se);
  }
  LAUNCHED(ctx);
  // bottom-up sweeps over the cells that straddle the emit slabs (a few thousand per level at most),
  // deepest level first; a level's node count is only known on the device
  // one CTA per SM, all resident (128 threads each): the kernel's grid barrier needs every CTA running
  CK(cudaMemsetAsync(ctx->grid_barrier, 0, sizeof(unsigned int), st));
  aggregate_levels_kernel<<<ctx->sm_count, 128, 0, st>>>(ctx->meta, b.pqr, b.accm, ctx->t,
                                                        StrictDirect{se.direct, se.cidx, se.cw}, ctx->grid_barrier, se.qstat);
  LAUNCHED(ctx);
  if (ctx->cfg.strict_centres) {
    const int32_t rc = strict_stage(ctx);
    if (rc) return rc;
  }
  // traversal arrays (charged nodes only)
  CK(exclusive_scan_dyn(ChargedFlagFn{ctx->t.ndepth}, &ctx->meta->num_nodes, ctx->node_cap, ctx->trav_rank,
                        ctx->scan_partials, ctx->trav_count, st));
  ctx->launches += 3;
  compact_traversal_kernel<<<grid_for(ctx, (uint64_t)n * 2, 256, 16), 256, 0, st>>>(
      ctx->meta, ctx->t.nodeA, ctx->t.nodeB, ctx->trav_rank, ctx->trav_count, ctx->node_cap, ctx->travA,
      ctx->travB);
  LAUNCHED(ctx);
  link_children_kernel<<<grid_for(ctx, (uint64_t)n, 256, 16), 256, 0, st>>>(ctx->trav_count, ctx->travA, ctx->travB);
  LAUNCHED(ctx);
  ctx->tree_valid = true;
  ctx->perm_valid = true;
  return PSIM_OK;
}

// ---- sharded build (shard.cuh).  The phases alternate with the exchanges the host performs on the
// buffers psim_shard_ptrs names (particlesim_b200/parallel.py):
//   0  root, keys + bin histogram of all bodies, splitters        -> sync, out = body_lo[world + 1]
//   1  owned keys: select, radix passes, run fix-up, index segment -> all-gather perm (uneven segments)
//   2  gather bodies, halo keys, levels, node scan, table 1        -> all-reduce xbuf
//   3  node offsets, emit, local sweeps, top heap                  -> all-reduce heap
//   4  top sweep, centres, charged scan, table 2                   -> all-reduce xbuf
//   5  traversal offsets, compaction into the rank's segment       -> sync, out = trav_lo[world + 1];
//                                                                    all-gather travA / travB segments
//   6  children links over the gathered traversal array
int32_t shard_phase(psim_ctx* ctx, int phase, int mode, float hw, float hh, uint32_t* out) {
  auto& S = ctx->sh;
  if (!S.on) return fail(ctx, PSIM_E_STATE, "psim_shard_phase: call psim_shard_init first");
  if (ctx->cfg.parity_mode == 2)
    return fail(ctx, PSIM_E_ARG, "sharded build: parity_mode 2 walks the whole node array, which only a single-GPU build has");
  if (phase != 0 && phase != S.phase) return fail(ctx, PSIM_E_STATE, "psim_shard_phase: phases must run in order");
  const uint32_t n = ctx->n;
  cudaStream_t st = ctx->stream;
  const uint32_t c_eff = effective_capacity(ctx->cfg.leaf_capacity, ctx->cfg.thread_capacity);
  if (c_eff + 1 > kHaloMax) return fail(ctx, PSIM_E_ARG, "sharded build: leaf capacity above 2048");
  if (n == 0) return fail(ctx, PSIM_E_ARG, "sharded build: no bodies");
  const uint32_t nl = S.n_local;
  switch (phase) {
    case 0: {
      ctx->tree_valid = ctx->perm_valid = ctx->host_map_valid = ctx->grid_valid = false;
      S.tree_is_sharded = true;
      S.tree_is_let = false;
      S.mode = mode;
      BodyArrays& in = ctx->b[ctx->cur];
      int nb = 1;
      if (mode == PSIM_BUILD_CONTAINING) {
        nb = grid_for(ctx, n, 256, 4);
        bounds_partial_kernel<<<nb, 256, 0, st>>>(in.pqr, n, ctx->bounds_partial);
        LAUNCHED(ctx);
      }
      root_quad_kernel<<<1, 32, 0, st>>>(ctx->bounds_partial, nb, mode, hw, hh, n, ctx->meta);
      CK(cudaMemsetAsync(S.binhist, 0, kBins * sizeof(uint32_t), st));
      bins_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(in.pqr, n, ctx->meta, S.bins, S.binhist);
      bin_split_kernel<<<1, 1024, 0, st>>>(S.binhist, n, S.world, S.binprefix, S.plan);
      ctx->launches += 3;
      CK(cudaMemcpyAsync(&S.plan_h, S.plan, sizeof(ShardPlan), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      S.s_lo = S.plan_h.body_lo[S.rank];
      S.n_local = S.plan_h.body_lo[S.rank + 1] - S.s_lo;
      S.halo = c_eff + 1;
      S.hl = S.s_lo < S.halo ? S.s_lo : S.halo;
      const uint32_t after = n - S.plan_h.body_lo[S.rank + 1];
      S.L = S.hl + S.n_local + (after < S.halo ? after : S.halo);
      ShardMeta& m = S.meta_h;
      memset(&m, 0, sizeof m);
      m.rank = S.rank, m.world = S.world, m.n_local = S.n_local, m.hl = S.hl, m.L = S.L;
      m.body_base = S.s_lo - S.hl;
      CK(cudaMemcpyAsync(S.meta, &m, sizeof m, cudaMemcpyHostToDevice, st));
      if (out) memcpy(out, S.plan_h.body_lo, (S.world + 1) * sizeof(uint32_t));
      break;
    }
    case 1: {
      const uint32_t b0 = S.plan_h.bin_lo[S.rank], b1 = S.plan_h.bin_lo[S.rank + 1];
      // keys[0]: the owned bodies' keys by compact slot; nodebase (free until phase 2): their body indices
      CK(exclusive_scan(InRangeFn{S.bins, b0, b1}, n, ctx->trav_rank, ctx->scan_partials, nullptr, st));
      select_owned_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(ctx->b[ctx->cur].pqr, S.bins, ctx->trav_rank, n,
                                                                     b0, b1, ctx->meta, ctx->keys[0], ctx->khi[0],
                                                                     ctx->idx[0], ctx->nodebase);
      ctx->launches += 4;
      ctx->sc.plan = ctx->tree_plan;
      CK(onesweep_sort<uint32_t>(ctx->khi[0], ctx->khi[1], ctx->idx[0], ctx->idx[1], nl, 0, kTreePasses, ctx->sc,
                                 ctx->sm_count, st));
      ctx->launches += 2 + kTreePasses;
      uint64_t* lk = S.lkeys + S.hl;
      gather_keys_kernel<<<grid_for(ctx, nl, 256, 16), 256, 0, st>>>(ctx->keys[0], ctx->idx[0], ctx->idx[1],
                                                                    ctx->tree_plan, kTreePasses, nl, lk, ctx->long_runs);
      fix_runs_kernel<<<grid_for(ctx, nl, 256, 16), 256, 0, st>>>(lk, ctx->idx[0], ctx->idx[1], ctx->tree_plan,
                                                                 kTreePasses, nl, ctx->long_runs, ctx->long_runs + 1);
      sort_long_runs_kernel<<<ctx->sm_count, 256, 0, st>>>(lk, ctx->idx[0], ctx->idx[1], ctx->tree_plan, kTreePasses,
                                                           nl, ctx->long_runs, ctx->long_runs + 1, ctx->keys[0]);
      copy_sorted_idx_kernel<<<grid_for(ctx, nl, 256, 16), 256, 0, st>>>(ctx->idx[0], ctx->idx[1], ctx->tree_plan,
                                                                        kTreePasses, nl, ctx->nodebase,
                                                                        ctx->keys_idx_all + S.s_lo);
      ctx->launches += 4;
      break;
    }
    case 2: {
      // keys_idx_all now holds the global sorted order (all-gathered)
      const int32_t rc = gather_stage(ctx, ctx->keys_idx_all, ctx->keys_idx_all, ctx->tree_plan, kTreePasses);
      if (rc) return rc;
      BodyArrays& b = ctx->b[ctx->cur];
      halo_kernel<<<grid_for(ctx, 2 * S.halo, 256, 1), 256, 0, st>>>(S.binprefix, S.meta, S.lkeys, ctx->le);
      tree_count_range_kernel<<<grid_for(ctx, nl, 256, 8), 256, 0, st>>>(
          S.lkeys, b.pqr + (S.s_lo - S.hl), S.L, S.hl, nl, c_eff, kShardDepth, ctx->meta, ctx->le);
      CK(exclusive_scan(LeCountFn{ctx->le}, S.L, ctx->nodebase, ctx->scan_partials, &S.meta->M_local, st));
      CK(cudaMemsetAsync(S.xbuf, 0, (kBins + kMaxRanks) * sizeof(unsigned long long), st));
      table_nodes_kernel<<<grid_for(ctx, kBins, 256, 2), 256, 0, st>>>(S.binhist, S.binprefix, S.plan, S.meta,
                                                                      ctx->nodebase, &S.meta->M_local, S.xbuf);
      ctx->launches += 6;
      break;
    }
    case 3: {
      BodyArrays& b = ctx->b[ctx->cur];
      resolve_table_kernel<<<grid_for(ctx, kBins + 1, 256, 2), 256, 0, st>>>(S.binhist, S.binprefix, S.plan, S.meta,
                                                                            S.xbuf, n, 0, S.nb_bin);
      globalize_nodebase_kernel<<<grid_for(ctx, S.L, 256, 16), 256, 0, st>>>(S.meta, S.lkeys, S.nb_bin, ctx->nodebase,
                                                                            ctx->t.ndepth);
      level_scan_shard_kernel<<<1, 32, 0, st>>>(ctx->meta, S.meta, ctx->node_cap);
      CK(cudaMemsetAsync(S.heap, 0, kTopSlots * sizeof(TopRec), st));
      tree_emit_range_kernel<<<grid_for(ctx, nl, 128, 16), 128, 0, st>>>(
          S.lkeys, S.meta, ctx->le, ctx->nodebase, b.pqr + (S.s_lo - S.hl), ctx->cfg.leaf_capacity,
          ctx->cfg.thread_capacity, ctx->meta, ctx->t, S.heap);
      for (int level = kMaxLevels - 1; level >= kShardDepth; --level)
        aggregate_level_shard_kernel<<<ctx->sm_count * 16, 128, 0, st>>>(level, ctx->meta, S.meta, ctx->t);
      heap_bins_kernel<<<grid_for(ctx, kBins, 256, 2), 256, 0, st>>>(S.plan, S.meta, ctx->t, S.heap);
      ctx->launches += 5 + (kMaxLevels - kShardDepth);
      break;
    }
    case 4: {
      BodyArrays& b = ctx->b[ctx->cur];
      for (int d = kShardDepth - 1; d >= 0; --d)
        heap_sweep_kernel<<<grid_for(ctx, 1u << (2 * d), 256, 1), 256, 0, st>>>(S.heap, d);
      heap_writeback_kernel<<<grid_for(ctx, kTopSlots, 256, 2), 256, 0, st>>>(S.meta, S.heap, ctx->t);
      StrictDirect sd;
      if (ctx->cfg.strict_centres) {
        const int32_t rc = strict_prepare(ctx);
        if (rc) return rc;
        sd = StrictDirect{kStrictDirect, ctx->strict.cidx, ctx->strict.cw};
      }
      finalize_nodes_shard_kernel<<<grid_for(ctx, (uint64_t)nl * 2 + 1, 256, 16), 256, 0, st>>>(
          ctx->meta, S.meta, b.pqr, b.accm, S.lkeys, S.binprefix, ctx->t, sd);
      if (ctx->cfg.strict_centres) {
        // node centres by the reference's serial f32 sums (strict.cuh) for this rank's piece, the cells above the
        // bins included: each belongs to the piece of the rank that owns its first body
        const int32_t rc = strict_stage(ctx, true);
        if (rc) return rc;
      }
      CK(exclusive_scan_dyn(ChargedFlagFn{ctx->t.ndepth}, &S.meta->M_local, ctx->node_cap, ctx->trav_rank,
                            ctx->scan_partials, &S.meta->T_local, st));
      CK(cudaMemsetAsync(S.xbuf, 0, (kBins + kMaxRanks) * sizeof(unsigned long long), st));
      table_trav_kernel<<<grid_for(ctx, kBins, 256, 2), 256, 0, st>>>(S.binhist, S.plan, S.meta, S.nb_bin, ctx->trav_rank,
                                                                     &S.meta->T_local, S.xbuf);
      ctx->launches += 6 + kShardDepth;
      break;
    }
    case 5: {
      resolve_table_kernel<<<grid_for(ctx, kBins + 1, 256, 2), 256, 0, st>>>(S.binhist, S.binprefix, S.plan, S.meta,
                                                                            S.xbuf, n, 1, S.trav_bin);
      compact_traversal_shard_kernel<<<grid_for(ctx, (uint64_t)nl * 2 + 1, 256, 16), 256, 0, st>>>(
          S.meta, ctx->t.nodeA, ctx->t.nodeB, ctx->trav_rank, S.nb_bin, S.trav_bin, ctx->node_cap, ctx->travA,
          ctx->travB, ctx->trav_count);
      ctx->launches += 2;
      CK(cudaMemcpyAsync(&S.meta_h, S.meta, sizeof(ShardMeta), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (S.meta_h.M_local > ctx->node_cap || S.meta_h.T_total > ctx->node_cap)
        return fail(ctx, PSIM_E_NODE_OVERFLOW, "sharded build: node arena overflow (raise node_factor)");
      if (out) memcpy(out, S.meta_h.trav_lo, (S.world + 1) * sizeof(uint32_t));
      break;
    }
    case 6: {
      // the traversal segments of all ranks are in place: children links over the whole array
      link_children_kernel<<<grid_for(ctx, (uint64_t)n, 256, 16), 256, 0, st>>>(ctx->trav_count, ctx->travA, ctx->travB);
      LAUNCHED(ctx);
      ctx->tree_valid = true;
      ctx->perm_valid = true;
      break;
    }
    default:
      return fail(ctx, PSIM_E_ARG, "psim_shard_phase: phase 0..6");
  }
  S.phase = phase == 6 ? 0 : phase + 1;
  const cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) return fail(ctx, PSIM_E_CUDA, "sharded build launch", le);
  return PSIM_OK;
}

int32_t cell_build_async(psim_ctx* ctx, float hw, float hh, float cell_size) {
  cudaStream_t st = ctx->stream;
  ctx->grid_valid = false;
  ctx->cell_off_valid = false;
  // cell_list.rs:28-29
  const float fx = ceilf((2.0f * hw) / cell_size), fy = ceilf((2.0f * hh) / cell_size);
  if (!(fx >= 0.0f) || !(fy >= 0.0f) || !(cell_size > 0.0f)) return fail(ctx, PSIM_E_ARG, "cell grid: bad domain or cell size");
  const double gxd = (double)fx + 1.0, gyd = (double)fy + 1.0;
  if (gxd * gyd > (double)ctx->cell_cap || gxd * gyd >= 4294967295.0) {
    // grow the cell arrays
    if (gxd * gyd >= 4.0e9) return fail(ctx, PSIM_E_ARG, "cell grid: more than 4e9 cells");
    const uint64_t need = (uint64_t)(gxd * gyd);
    if (ctx->cell_start) cudaFree(ctx->cell_start);
    if (ctx->cell_end) cudaFree(ctx->cell_end);
    if (ctx->cell_off) cudaFree(ctx->cell_off);
    ctx->cell_start = ctx->cell_end = ctx->cell_off = nullptr;
    ctx->cell_cap = 0;
    cudaError_t e1 = dalloc(&ctx->cell_start, need + 1), e2 = dalloc(&ctx->cell_end, need + 1);
    if (e1 == cudaSuccess) e1 = dalloc(&ctx->cell_off, need + 2);
    if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(ctx, PSIM_E_OOM, "cell arrays", e1 != cudaSuccess ? e1 : e2);
    ctx->cell_cap = need;
  }
  GridDims g;
  g.gx = (uint32_t)gxd, g.gy = (uint32_t)gyd, g.cell_size = cell_size, g.hw = hw, g.hh = hh;
  ctx->grid = g;
  const uint64_t ncells = (uint64_t)g.gx * g.gy;
  const uint32_t n = ctx->n;
  CK(cudaMemsetAsync(ctx->cell_start, 0, ncells * sizeof(uint32_t), st));
  CK(cudaMemsetAsync(ctx->cell_end, 0, ncells * sizeof(uint32_t), st));
  if (n == 0) {
    ctx->grid_valid = true;
    return PSIM_OK;
  }
  BodyArrays& b = ctx->b[ctx->cur];
  cell_id_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(b.pqr, n, g, ctx->ckeys[0], ctx->cidx[0]);
  LAUNCHED(ctx);
  int bits = 1;
  while (bits < 32 && (1ull << bits) < ncells) ++bits;
  const int npass = (bits + 7) / 8;
  ctx->cell_passes = npass;
  ctx->sc.plan = ctx->cell_plan;
  CK(onesweep_sort<uint32_t>(ctx->ckeys[0], ctx->ckeys[1], ctx->cidx[0], ctx->cidx[1], n, 0, npass,
                             ctx->sc, ctx->sm_count, st));
  ctx->launches += 2 + npass;
  cell_ranges_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(
      ctx->ckeys[0], ctx->ckeys[1], ctx->cidx[0], ctx->cidx[1], ctx->cell_plan, npass, n,
      ctx->cell_start, ctx->cell_end, ctx->order, ctx->body_cell, b.pqr, b.species, ctx->cpos);
  LAUNCHED(ctx);
  ctx->grid_valid = true;
  return PSIM_OK;
}

int32_t field_async(psim_ctx* ctx, float k_e, float bg_x, float bg_y, int write_acc) {
  if (!ctx->tree_valid) return fail(ctx, PSIM_E_STATE, "psim_field: no tree (call psim_build first)");
  if (ctx->cfg.parity_mode == 2 && ctx->sh.tree_is_sharded)
    return fail(ctx, PSIM_E_STATE, "psim_field: parity_mode 2 needs the whole node array (the last build was sharded)");
  const uint32_t n = ctx->n;
  if (n == 0) return PSIM_OK;
  BodyArrays& b = ctx->b[ctx->cur];
  const FieldParams P = field_params(ctx, k_e, bg_x, bg_y);
  uint32_t first, count;
  body_range(ctx, first, count);
  if (count == 0) return PSIM_OK;
  const uint32_t groups = ((ctx->cfg.parity_mode == 2 ? n : count) + 31) / 32;
  if (ctx->cfg.parity_mode == 2) {
    // reference-order walk: bit-identical additions, one node per warp step, on the full tree
    const int blocks = (int)((groups + 3) / 4);
    bh_field_bodies_kernel<true><<<blocks, 128, 0, ctx->stream>>>(
        ctx->meta, ctx->t.nodeA, ctx->t.nodeB, b.pqr, b.accm, n, P, b.efield, b.accm, write_acc,
        ctx->step_counter);
  } else {
    const int blocks = (int)((groups + 3) / 4);  // one group per warp, launched in Morton order
    if (ctx->cfg.parity_mode)
      bh_group_bodies_kernel<true><<<blocks, 128, 0, ctx->stream>>>(
          ctx->travA, ctx->travB, ctx->trav_count, b.pqr, b.accm, first, count, P, b.efield, b.accm, write_acc,
          ctx->step_counter, ctx->meta);
    else
      bh_group_bodies_kernel<false><<<blocks, 128, 0, ctx->stream>>>(
          ctx->travA, ctx->travB, ctx->trav_count, b.pqr, b.accm, first, count, P, b.efield, b.accm, write_acc,
          ctx->step_counter, ctx->meta);
  }
  LAUNCHED(ctx);
  return PSIM_OK;
}

int32_t points_async(psim_ctx* ctx, const float2* pts, const float* q, const float* radius, uint32_t m,
                     float k_e, float2* out, uint32_t first = 0) {
  if (!ctx->tree_valid) return fail(ctx, PSIM_E_STATE, "acc_pos: no tree (call psim_build first)");
  if (ctx->cfg.parity_mode == 2 && ctx->sh.tree_is_sharded)
    return fail(ctx, PSIM_E_STATE, "acc_pos: parity_mode 2 needs the whole node array (the last build was sharded)");
  if (m == 0) return PSIM_OK;
  BodyArrays& b = ctx->b[ctx->cur];
  const FieldParams P = field_params(ctx, k_e, 0.f, 0.f);
  const uint32_t groups = (m + 31) / 32;
  const int blocks = (int)((groups + 3) / 4);
  if (ctx->cfg.parity_mode == 2)
    bh_field_points_kernel<true><<<blocks, 128, 0, ctx->stream>>>(
        ctx->meta, ctx->t.nodeA, ctx->t.nodeB, b.pqr, pts + first, q ? q + first : nullptr,
        radius ? radius + first : nullptr, m, P, out + first, ctx->step_counter);
  else if (ctx->cfg.parity_mode)
    bh_group_points_kernel<true><<<blocks, 128, 0, ctx->stream>>>(
        ctx->travA, ctx->travB, ctx->trav_count, b.pqr, pts, q, radius, first, m, P, out, ctx->step_counter, ctx->meta);
  else
    bh_group_points_kernel<false><<<blocks, 128, 0, ctx->stream>>>(
        ctx->travA, ctx->travB, ctx->trav_count, b.pqr, pts, q, radius, first, m, P, out, ctx->step_counter, ctx->meta);
  LAUNCHED(ctx);
  return PSIM_OK;
}

int32_t electrons_async(psim_ctx* ctx, float bg_x, float bg_y, float dt, float k_e) {
  if (ctx->m == 0) return PSIM_OK;
  if (!ctx->tree_valid) return fail(ctx, PSIM_E_STATE, "psim_update_electrons: no tree");
  uint32_t first, m;
  electron_range(ctx, first, m);
  if (m == 0) return PSIM_OK;
  BodyArrays& b = ctx->b[ctx->cur];
  const int e = ctx->ecur;
  electron_points_kernel<<<grid_for(ctx, m, 256, 16), 256, 0, ctx->stream>>>(
      b.pqr, ctx->ebody[e] + first, ctx->erel[e] + first, m, ctx->epts + first);
  LAUNCHED(ctx);
  int32_t rc = points_async(ctx, ctx->epts, nullptr, nullptr, m, k_e, ctx->efld, first);
  if (rc) return rc;
  // config.rs:6-9,27-35: electron_spring_k() is 5.0 for every species; config.rs:49
  electron_drift_kernel<<<grid_for(ctx, m, 256, 16), 256, 0, ctx->stream>>>(
      b.pqr, b.species, ctx->table_d, ctx->ebody[e] + first, ctx->erel[e] + first, ctx->evel[e] + first,
      ctx->efld + first, m, bg_x, bg_y, dt, 5.0f, 10.2f);
  LAUNCHED(ctx);
  return PSIM_OK;
}

// exclusive prefix of the cell populations: the bodies of cells [c0, c1] are the slots [cell_off[c0], cell_off[c1 + 1])
// of the cell order (a row of neighbouring cells is ONE contiguous run)
int32_t ensure_cell_off(psim_ctx* ctx) {
  if (ctx->cell_off_valid) return PSIM_OK;
  const uint64_t ncells = (uint64_t)ctx->grid.gx * ctx->grid.gy;
  int32_t rc = ensure_qstage(ctx, ((size_t)scan_num_tiles((uint32_t)ncells) + 2) * sizeof(uint32_t));
  if (rc) return rc;
  CK(exclusive_scan(CellCountFn{ctx->cell_start, ctx->cell_end}, (uint32_t)ncells, ctx->cell_off,
                    static_cast<uint32_t*>(ctx->qstage), ctx->cell_off + ncells, ctx->stream));
  ctx->launches += 3;
  ctx->cell_off_valid = true;
  return PSIM_OK;
}

int32_t short_range_async(psim_ctx* ctx, uint32_t flags) {
  const uint32_t n = ctx->n;
  if (n == 0) return PSIM_OK;
  ShortRangeParams P;
  memset(&P, 0, sizeof(P));
  // cutoffs over the species table (species.rs:412-479); a pass is a no-op when no uploaded body
  // belongs to a species it applies to (apply_lj_forces / apply_repulsive_forces skip such bodies)
  const float lj_cut = max_lj_cutoff(ctx), rep_cut = max_repulsion_cutoff(ctx);
  bool lj_present = false, rep_present = false;
  for (uint32_t sp = 0; sp < ctx->nspecies; ++sp) {
    if (!(ctx->species_present >> sp & 1u)) continue;
    lj_present |= ctx->table_h[sp].lj_enabled != 0;
    rep_present |= ctx->table_h[sp].repulsion_enabled != 0;
  }
  P.do_lj = (flags & PSIM_SR_LJ) && lj_cut > 0.0f && lj_present;
  P.do_rep = (flags & PSIM_SR_REPULSION) && rep_cut > 0.0f && rep_present;  // forces.rs:252-255
  P.do_stack = (flags & PSIM_SR_STACK_PRESSURE) && ctx->cfg.stack_pressure_enabled && ctx->cfg.stack_pressure > 0.0f;
  if (!P.do_lj && !P.do_rep && !P.do_stack) return PSIM_OK;
  if ((P.do_lj || P.do_rep) && !ctx->grid_valid)
    return fail(ctx, PSIM_E_STATE, "psim_short_range: no cell grid (call psim_cell_build after the last psim_build)");
  P.g = ctx->grid;
  if (!ctx->grid_valid) P.g.hw = 0.f, P.g.gx = P.g.gy = 1;
  P.max_lj_cutoff = lj_cut;
  P.max_rep_cutoff = P.do_rep ? rep_cut : 0.0f;
  P.max_lj_force = (float)ctx->cfg.collision_passes * ctx->cfg.lj_force_max;
  P.stack_pressure = ctx->cfg.stack_pressure;
  P.stack_decay = ctx->cfg.stack_pressure_decay;
  if (!ctx->grid_valid && P.do_stack) P.g.hw = ctx->grid.hw;
  const float reach = fmaxf(P.do_lj ? lj_cut : 0.0f, P.do_rep ? rep_cut : 0.0f);
  P.range = (P.do_lj || P.do_rep) ? (int)ceilf(reach / ctx->grid.cell_size) : 0;
  if (P.range < 0) P.range = 0;
  BodyArrays& b = ctx->b[ctx->cur];
  if (P.do_lj || P.do_rep) {
    const int32_t rc = ensure_cell_off(ctx);
    if (rc) return rc;
  }
  uint32_t first, count;
  body_range(ctx, first, count);
  if (count == 0) return PSIM_OK;
  short_range_kernel<<<(count + 127) / 128, 128, 0, ctx->stream>>>(
      b.pqr, b.species, ctx->table_d, first, first + count, ctx->cell_off, ctx->cpos, ctx->body_cell, P, b.accm);
  LAUNCHED(ctx);
  return PSIM_OK;
}

int32_t polar_async(psim_ctx* ctx, float k_e, int dipole_model) {
  const uint32_t n = ctx->n;
  if (n == 0) return PSIM_OK;
  if (!ctx->grid_valid)
    return fail(ctx, PSIM_E_STATE, "psim_apply_polar_forces: no cell grid (call psim_cell_build after the last psim_build)");
  cudaStream_t st = ctx->stream;
  BodyArrays& b = ctx->b[ctx->cur];
  const int e = ctx->ecur;
  CK(cudaMemsetAsync(ctx->polar_cutoff, 0, sizeof(uint32_t), st));
  {
    const int32_t rc = ensure_cell_off(ctx);
    if (rc) return rc;
  }
  // cell-ordered records; record A reuses the staging area of the short-range pass layout (x, y, charge, radius)
  int32_t rc = ensure_stage(ctx, (size_t)n * sizeof(float4));
  if (rc) return rc;
  float4* recA = static_cast<float4*>(ctx->stage);
  polar_records_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(ctx->order, n, b.pqr, b.species, b.ecount,
                                                                ctx->eoff[e], ctx->erel[e], recA, ctx->polarB,
                                                                ctx->polar_cutoff);
  LAUNCHED(ctx);
  PolarParams P;
  P.g = ctx->grid;
  P.k_e = k_e;
  P.epsilon_sq = 2.0f * 2.0f;  // config::QUADTREE_EPSILON squared (forces.rs:60), not the tree's epsilon
  P.dipole_model = dipole_model;
  uint32_t first, count;
  body_range(ctx, first, count);
  if (count == 0) return PSIM_OK;
  if (ctx->cfg.parity_mode)
    polar_forces_kernel<true><<<(count + kPolarThreads - 1) / kPolarThreads, kPolarThreads, 0, st>>>(
        b.pqr, b.species, b.ecount, ctx->eoff[e], ctx->erel[e], ctx->table_d, first, first + count, ctx->cell_off, recA,
        ctx->polarB, ctx->body_cell, ctx->polar_cutoff, P, b.accm);
  else
    polar_forces_kernel<false><<<(count + kPolarThreads - 1) / kPolarThreads, kPolarThreads, 0, st>>>(
        b.pqr, b.species, b.ecount, ctx->eoff[e], ctx->erel[e], ctx->table_d, first, first + count, ctx->cell_off, recA,
        ctx->polarB, ctx->body_cell, ctx->polar_cutoff, P, b.accm);
  LAUNCHED(ctx);
  return PSIM_OK;
}

int32_t iterate_async(psim_ctx* ctx, float dt, float damping_base, float hw, float hh, float hd, int enable_z) {
  const uint32_t n = ctx->n;
  if (n == 0) return PSIM_OK;
  IterateParams P;
  P.dt = dt;
  P.base_damping = powf(damping_base, dt / 0.01f);  // simulation.rs:1441
  P.hw = hw, P.hh = hh, P.hd = hd, P.enable_z = enable_z;
  BodyArrays& b = ctx->b[ctx->cur];
  uint32_t first, count;
  body_range(ctx, first, count);
  if (count)
    iterate_kernel<<<grid_for(ctx, count, 256, 16), 256, 0, ctx->stream>>>(
        b.pqr + first, b.velz + first, b.accm + first, b.species + first, ctx->table_d, count, P);
  LAUNCHED(ctx);
  // positions moved: tree and grid no longer describe them
  ctx->tree_valid = false;
  ctx->grid_valid = false;
  return PSIM_OK;
}

int32_t check_build(psim_ctx* ctx) {
  int32_t rc = fetch_meta(ctx);
  if (rc) return rc;
  if (ctx->meta_h.err & 1u) {
    ctx->tree_valid = false;
    char msg[160];
    snprintf(msg, sizeof msg, "tree needs %u nodes but the arena holds %u (raise psim_config.node_factor)",
             ctx->meta_h.num_nodes, ctx->node_cap);
    return fail(ctx, PSIM_E_NODE_OVERFLOW, msg);
  }
  if (ctx->meta_h.err & 2u) {
    ctx->tree_valid = false;
    return fail(ctx, PSIM_E_NODE_OVERFLOW, "strict_centres: long-chain scratch overflow");
  }
  return PSIM_OK;
}

void free_all(psim_ctx* c) {
  auto F = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  for (int k = 0; k < 2; ++k) {
    F(c->b[k].pqr), F(c->b[k].velz), F(c->b[k].accm), F(c->b[k].efield), F(c->b[k].species), F(c->b[k].orig), F(c->b[k].ecount);
    F(c->ebody[k]), F(c->erel[k]), F(c->evel[k]), F(c->eoff[k]);
    F(c->keys[k]), F(c->khi[k]), F(c->idx[k]), F(c->ckeys[k]), F(c->cidx[k]);
  }
  F(c->epts), F(c->efld);
  F(c->sc.hist), F(c->sc.status), F(c->sc.ticket), F(c->tree_plan), F(c->cell_plan), F(c->long_runs), F(c->keys_idx_all);
  F(c->sh.binhist), F(c->sh.bins), F(c->sh.binprefix), F(c->sh.nb_bin), F(c->sh.trav_bin), F(c->sh.lkeys), F(c->sh.xbuf), F(c->sh.heap);
  F(c->sh.plan), F(c->sh.meta);
  F(c->meta), F(c->le), F(c->nodebase), F(c->scan_partials), F(c->irank), F(c->bounds_partial);
  F(c->t.nodeA), F(c->t.nodeB), F(c->t.node_mass), F(c->t.parent), F(c->t.sums), F(c->t.level_nodes), F(c->t.local_nodes);
  F(c->t.rec), F(c->t.ndepth);
  F(c->travA), F(c->travB), F(c->trav_rank), F(c->trav_count);
  F(c->perm), F(c->inv);
  F(c->surround.last_pos), F(c->surround.last_frame), F(c->surround.flag);
  F(c->cell_start), F(c->cell_end), F(c->order), F(c->body_cell), F(c->cpos), F(c->polarB), F(c->polar_cutoff), F(c->cell_off);
  F(c->table_d), F(c->stage), F(c->qstage), F(c->step_counter), F(c->grid_barrier);
  F(c->strict.qstat), F(c->strict.qc), F(c->strict.cidx), F(c->strict.cw), F(c->strict.chains), F(c->strict.hist), F(c->strict.longs), F(c->strict.counters);
  F(c->strict.item_first), F(c->strict.pblk), F(c->strict.fns), F(c->strict.cand);
}

}  // namespace

// ================================================================================================
extern "C" {

void psim_default_config(psim_config* cfg) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->theta = 1.0f;            // config.rs:213
  cfg->epsilon = 2.0f;          // config.rs:214
  cfg->leaf_capacity = 1;       // config.rs:215
  cfg->thread_capacity = 1024;  // config.rs:216
  cfg->lj_force_max = 200.0f;   // config.rs:126
  cfg->collision_passes = 7;    // config.rs:204
  cfg->stack_pressure_enabled = 0;
  cfg->stack_pressure = 0.0f;
  cfg->stack_pressure_decay = 1.0f;
  cfg->parity_mode = 1;
  cfg->node_factor = 4.0f;
  cfg->strict_centres = 1;      // node centres by the reference's own serial f32 sums (quadtree.rs:114-139)
}

void psim_default_species_table(psim_species* rows21) {
  default_species(reinterpret_cast<SpeciesRow*>(rows21));
}

const char* psim_last_error(const psim_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int32_t psim_create(int32_t device, uint64_t max_bodies, uint64_t max_electrons, const psim_config* cfg,
                    psim_ctx** out) {
  if (!out) return PSIM_E_ARG;
  *out = nullptr;
  if (max_bodies >= (1ull << 30) || max_electrons >= (1ull << 30)) return PSIM_E_ARG;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return PSIM_E_CUDA;  // no device: fail loudly, there is no CPU path
  }
  psim_ctx* ctx = new (std::nothrow) psim_ctx();
  if (!ctx) return PSIM_E_OOM;
  ctx->device = device;
  if (cfg) ctx->cfg = *cfg; else psim_default_config(&ctx->cfg);
  if (!(ctx->cfg.node_factor >= 1.0f)) ctx->cfg.node_factor = 4.0f;
  if (cudaSetDevice(device) != cudaSuccess) {
    delete ctx;
    return PSIM_E_CUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  const uint64_t nb = max_bodies ? max_bodies : 1, ne = max_electrons ? max_electrons : 1;
  ctx->cap_bodies = max_bodies;
  ctx->cap_elec = max_electrons;
  ctx->node_cap = (uint32_t)fmin((double)ctx->cfg.node_factor * (double)nb + 1024.0, 4.0e9);
  bool ok = true;
  auto A = [&](auto** p, size_t cnt) {
    if (ok && dalloc(p, cnt) != cudaSuccess) ok = false;
  };
  for (int k = 0; k < 2; ++k) {
    A(&ctx->b[k].pqr, nb), A(&ctx->b[k].velz, nb), A(&ctx->b[k].accm, nb), A(&ctx->b[k].efield, nb);
    A(&ctx->b[k].species, nb), A(&ctx->b[k].orig, nb), A(&ctx->b[k].ecount, nb);
    A(&ctx->ebody[k], ne), A(&ctx->erel[k], ne), A(&ctx->evel[k], ne), A(&ctx->eoff[k], nb + 1);
    A(&ctx->keys[k], nb), A(&ctx->khi[k], nb), A(&ctx->idx[k], nb), A(&ctx->ckeys[k], nb), A(&ctx->cidx[k], nb);
  }
  A(&ctx->epts, ne), A(&ctx->efld, ne);
  A(&ctx->sc.hist, 8 * kRadix), A(&ctx->sc.ticket, 8);
  ctx->sc.status_words = sort_status_words((uint32_t)nb, kTreePasses);
  A(&ctx->sc.status, ctx->sc.status_words);
  A(&ctx->tree_plan, 1), A(&ctx->cell_plan, 1), A(&ctx->long_runs, nb / kFixInsertion + 2);
  A(&ctx->meta, 1), A(&ctx->le, nb + 2 * kHaloMax + 2), A(&ctx->nodebase, nb + 2 * kHaloMax + 2);
  A(&ctx->scan_partials, (size_t)scan_num_tiles(ctx->node_cap > nb ? ctx->node_cap : (uint32_t)nb) + 1);
  A(&ctx->irank, ctx->node_cap), A(&ctx->bounds_partial, (size_t)ctx->sm_count * 4 + 1);
  A(&ctx->t.nodeA, ctx->node_cap), A(&ctx->t.nodeB, ctx->node_cap), A(&ctx->t.node_mass, ctx->node_cap);
  A(&ctx->t.parent, ctx->node_cap), A(&ctx->t.sums, ctx->node_cap), A(&ctx->t.level_nodes, ctx->node_cap), A(&ctx->t.local_nodes, ctx->node_cap);
  A(&ctx->t.rec, ctx->node_cap), A(&ctx->t.ndepth, (size_t)ctx->node_cap + 1);
  A(&ctx->travA, ctx->node_cap), A(&ctx->travB, ctx->node_cap), A(&ctx->trav_rank, ctx->node_cap), A(&ctx->trav_count, 1);
  ctx->t.node_cap = ctx->node_cap;
  A(&ctx->perm, nb), A(&ctx->inv, nb);
  A(&ctx->surround.last_pos, nb), A(&ctx->surround.last_frame, nb), A(&ctx->surround.flag, nb);
  A(&ctx->order, nb), A(&ctx->body_cell, nb), A(&ctx->cpos, nb);
  A(&ctx->polarB, nb), A(&ctx->polar_cutoff, 1);
  A(&ctx->table_d, kMaxSpecies), A(&ctx->step_counter, 1), A(&ctx->grid_barrier, 1);
  if (!ok) {
    cudaGetLastError();
    free_all(ctx);
    delete ctx;
    return PSIM_E_OOM;
  }
  memset(ctx->table_h, 0, sizeof(ctx->table_h));
  default_species(ctx->table_h);
  ctx->nspecies = 21;
  cudaMemcpy(ctx->table_d, ctx->table_h, sizeof(ctx->table_h), cudaMemcpyHostToDevice);
  cudaMemset(ctx->meta, 0, sizeof(TreeMeta));
  cudaMemset(ctx->step_counter, 0, sizeof(unsigned long long));
  ctx->ev_ok = true;
  for (int k = 0; k < 9; ++k)
    if (cudaEventCreate(&ctx->ev[k]) != cudaSuccess) ctx->ev_ok = false;
  *out = ctx;
  return PSIM_OK;
}

int32_t psim_destroy(psim_ctx* ctx) {
  if (!ctx) return PSIM_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  psim_comm_destroy(ctx);
  if (ctx->ev_ok)
    for (int k = 0; k < 9; ++k) cudaEventDestroy(ctx->ev[k]);
  if (ctx->ov.side) cudaStreamSynchronize(ctx->ov.side), cudaStreamDestroy(ctx->ov.side);
  if (ctx->ov.ev_fork) cudaEventDestroy(ctx->ov.ev_fork);
  if (ctx->ov.ev_join) cudaEventDestroy(ctx->ov.ev_join);
  if (ctx->ov.partials) cudaFree(ctx->ov.partials);
  if (ctx->copy_in) cudaStreamSynchronize(ctx->copy_in), cudaStreamDestroy(ctx->copy_in);
  if (ctx->copy_out) cudaStreamSynchronize(ctx->copy_out), cudaStreamDestroy(ctx->copy_out);
  for (cudaEvent_t e : {ctx->ev_start, ctx->ev_q, ctx->ev_vel, ctx->ev_mid, ctx->ev_out})
    if (e) cudaEventDestroy(e);
  if (ctx->hstage) cudaFree(ctx->hstage);
  if (ctx->host_map) cudaFree(ctx->host_map);
  free_all(ctx);
  delete ctx;
  return PSIM_OK;
}

int32_t psim_set_config(psim_ctx* ctx, const psim_config* cfg) {
  if (!ctx || !cfg) return PSIM_E_ARG;
  const float nf = ctx->cfg.node_factor;
  ctx->cfg = *cfg;
  ctx->cfg.node_factor = nf;  // the arena is sized at create time
  ctx->tree_valid = false;
  return PSIM_OK;
}

int32_t psim_set_stream(psim_ctx* ctx, uint64_t cuda_stream) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  return PSIM_OK;
}

int32_t psim_sync(psim_ctx* ctx) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  CK(cudaStreamSynchronize(ctx->stream));
  return PSIM_OK;
}

int32_t psim_reset_counters(psim_ctx* ctx) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  ctx->launches = 0;
  CK(cudaMemsetAsync(ctx->step_counter, 0, sizeof(unsigned long long), ctx->stream));
  return PSIM_OK;
}

int32_t psim_field_counters(psim_ctx* ctx, uint64_t* out4) {
  if (!ctx || !out4) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (!ctx->tree_valid) return fail(ctx, PSIM_E_STATE, "psim_field_counters: no tree");
  out4[0] = out4[1] = out4[2] = out4[3] = 0;
  if (ctx->n == 0) return PSIM_OK;
  unsigned long long* d = nullptr;
  CK(cudaMalloc(&d, 4 * sizeof(unsigned long long)));
  cudaMemsetAsync(d, 0, 4 * sizeof(unsigned long long), ctx->stream);
  BodyArrays& b = ctx->b[ctx->cur];
  const uint32_t groups = (ctx->n + 31) / 32;
  bh_count_bodies_kernel<<<grid_for(ctx, (uint64_t)groups * 32, 128, 64), 128, 0, ctx->stream>>>(
      ctx->meta, ctx->t.nodeA, ctx->t.nodeB, b.pqr, ctx->n, field_params(ctx, 1.0f, 0.f, 0.f), d);
  unsigned long long h[4] = {0, 0, 0, 0};
  cudaError_t e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(ctx, PSIM_E_CUDA, "psim_field_counters", e);
  for (int k = 0; k < 4; ++k) out4[k] = h[k];
  return PSIM_OK;
}

int32_t psim_build_info(psim_ctx* ctx, uint64_t* out4) {
  if (!ctx || !out4) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (!ctx->tree_valid) return fail(ctx, PSIM_E_STATE, "psim_build_info: no tree");
  out4[0] = out4[1] = out4[2] = out4[3] = 0;
  if (ctx->n == 0 || !ctx->cfg.strict_centres || !ctx->strict_ready || ctx->sh.tree_is_sharded) return PSIM_OK;
  const int32_t rc = fetch_meta(ctx);
  if (rc) return rc;
  unsigned long long q[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(q, ctx->strict.qstat, sizeof q, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const bool disabled = getenv("PSIM_INTEGER_CHARGES") && getenv("PSIM_INTEGER_CHARGES")[0] == '0';
  out4[0] = (!disabled && q[1] == 0 && q[0] < (1ull << 24) && ctx->meta_h.zero_agg_hint == 0) ? 1 : 0;
  out4[1] = (uint32_t)q[2] ? (uint32_t)q[2] - 1u : 0u;
  out4[2] = q[0];
  out4[3] = q[1];
  return PSIM_OK;
}

int32_t psim_fp32_peak(psim_ctx* ctx, float* tflops, int32_t* sm_count) {
  if (!ctx || !tflops) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  float* d = nullptr;
  CK(cudaMalloc(&d, sizeof(float)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int blocks = ctx->sm_count * 8, iters = 4096;
  float best = 0.0f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, ctx->stream);
    fp32_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(d, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * iters * 256.0 * blocks;
    const float tf = (float)(flops / (ms * 1e-3) / 1e12);
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  cudaFree(d);
  CK(cudaGetLastError());
  *tflops = best;
  if (sm_count) *sm_count = ctx->sm_count;
  return PSIM_OK;
}

int32_t psim_stats_get(psim_ctx* ctx, psim_stats* out) {
  if (!ctx || !out) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  memset(out, 0, sizeof(*out));
  int32_t rc = fetch_meta(ctx);
  if (rc) return rc;
  unsigned long long steps = 0;
  CK(cudaMemcpy(&steps, ctx->step_counter, sizeof steps, cudaMemcpyDeviceToHost));
  out->n_bodies = ctx->n;
  out->n_electrons = ctx->m;
  out->compact_nodes = ctx->meta_h.num_nodes;
  out->reference_nodes = ctx->n ? 4ull * ctx->meta_h.num_internal + 1ull : 0ull;
  out->max_depth = ctx->meta_h.max_depth;
  out->depth_cap = ctx->meta_h.dcap;
  out->zero_leaves = ctx->meta_h.num_zero_leaves;
  out->cap_leaves = ctx->meta_h.num_cap_leaves;
  out->root_center[0] = ctx->meta_h.root.cx;
  out->root_center[1] = ctx->meta_h.root.cy;
  out->root_size = ctx->meta_h.root.size;
  out->grid_x = ctx->grid.gx;
  out->grid_y = ctx->grid.gy;
  out->traversal_warp_steps = steps;
  out->kernel_launches = ctx->launches;
  return PSIM_OK;
}

int32_t psim_upload_species_table(psim_ctx* ctx, const psim_species* rows, uint32_t nrows) {
  if (!ctx || !rows || nrows == 0 || nrows > kMaxSpecies) return fail(ctx, PSIM_E_ARG, "species table: 1..32 rows");
  DeviceGuard guard(ctx->device);
  memset(ctx->table_h, 0, sizeof(ctx->table_h));
  memcpy(ctx->table_h, rows, nrows * sizeof(SpeciesRow));
  ctx->nspecies = nrows;
  CK(cudaMemcpyAsync(ctx->table_d, ctx->table_h, sizeof(ctx->table_h), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return PSIM_OK;
}

int32_t psim_upload_bodies(psim_ctx* ctx, uint64_t n, const float* pos_xy, const float* z, const float* vel_xy,
                           const float* vz, const float* mass, const float* radius, const float* charge,
                           const uint8_t* species) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (n > ctx->cap_bodies) return fail(ctx, PSIM_E_ARG, "psim_upload_bodies: n exceeds max_bodies");
  if (n && !pos_xy) return fail(ctx, PSIM_E_ARG, "psim_upload_bodies: pos_xy is null");
  cudaStream_t st = ctx->stream;
  ctx->n = (uint32_t)n;
  ctx->m = 0;
  ctx->tgt_set = ctx->etgt_set = false;
  ctx->tree_valid = ctx->grid_valid = ctx->perm_valid = false;
  ctx->host_map_valid = false;  // the hosted-step row map (psim_step_host) belongs to the previous body set
  if (n == 0) return PSIM_OK;
  // layout of the staging buffer: pos | vel | z | vz | mass | radius | charge | species
  const size_t o_pos = 0, o_vel = o_pos + align256(8 * n), o_z = o_vel + align256(8 * n),
               o_vz = o_z + align256(4 * n), o_m = o_vz + align256(4 * n), o_r = o_m + align256(4 * n),
               o_q = o_r + align256(4 * n), o_s = o_q + align256(4 * n), total = o_s + align256(n);
  int32_t rc = ensure_stage(ctx, total);
  if (rc) return rc;
  char* sb = static_cast<char*>(ctx->stage);
  RawBodies r;
  memset(&r, 0, sizeof r);
  auto up = [&](const void* src, size_t off, size_t bytes) -> const void* {
    if (!src) return nullptr;
    if (cudaMemcpyAsync(sb + off, src, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return nullptr;
    return sb + off;
  };
  r.pos = static_cast<const float2*>(up(pos_xy, o_pos, 8 * n));
  r.vel = static_cast<const float2*>(up(vel_xy, o_vel, 8 * n));
  r.z = static_cast<const float*>(up(z, o_z, 4 * n));
  r.vz = static_cast<const float*>(up(vz, o_vz, 4 * n));
  r.mass = static_cast<const float*>(up(mass, o_m, 4 * n));
  r.radius = static_cast<const float*>(up(radius, o_r, 4 * n));
  r.charge = static_cast<const float*>(up(charge, o_q, 4 * n));
  r.species = static_cast<const uint8_t*>(up(species, o_s, n));
  ctx->species_present = 0;
  if (species) {
    for (uint64_t i = 0; i < n; ++i) ctx->species_present |= 1u << (species[i] < kMaxSpecies ? species[i] : 0);
  } else {
    ctx->species_present = 1u;
  }
  CK(cudaGetLastError());
  if (!r.pos) return fail(ctx, PSIM_E_CUDA, "psim_upload_bodies: host to device copy failed");
  pack_bodies_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(r, (uint32_t)n, ctx->b[ctx->cur]);
  LAUNCHED(ctx);
  surround_init_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(ctx->b[ctx->cur].pqr, ctx->b[ctx->cur].orig,
                                                                 (uint32_t)n, ctx->surround);
  LAUNCHED(ctx);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));  // the caller may reuse its buffers
  return PSIM_OK;
}

int32_t psim_update_positions(psim_ctx* ctx, uint64_t n, const float* pos_xy) {
  if (!ctx || n != ctx->n || (n && !pos_xy)) return fail(ctx, PSIM_E_ARG, "psim_update_positions: size mismatch");
  DeviceGuard guard(ctx->device);
  ctx->host_map_valid = false;
  if (n == 0) return PSIM_OK;
  int32_t rc = ensure_stage(ctx, 8 * n);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->stage, pos_xy, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
  set_positions_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(static_cast<const float2*>(ctx->stage), (uint32_t)n, ctx->b[ctx->cur].pqr);
  LAUNCHED(ctx);
  ctx->tree_valid = ctx->grid_valid = false;
  CK(cudaStreamSynchronize(ctx->stream));
  return PSIM_OK;
}

int32_t psim_update_charges(psim_ctx* ctx, uint64_t n, const float* charge) {
  if (!ctx || n != ctx->n || (n && !charge)) return fail(ctx, PSIM_E_ARG, "psim_update_charges: size mismatch");
  DeviceGuard guard(ctx->device);
  ctx->host_map_valid = false;
  if (n == 0) return PSIM_OK;
  int32_t rc = ensure_stage(ctx, 4 * n);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->stage, charge, 4 * n, cudaMemcpyHostToDevice, ctx->stream));
  set_charges_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(static_cast<const float*>(ctx->stage), (uint32_t)n, ctx->b[ctx->cur].pqr);
  LAUNCHED(ctx);
  ctx->tree_valid = false;
  CK(cudaStreamSynchronize(ctx->stream));
  return PSIM_OK;
}

int32_t psim_upload_electrons(psim_ctx* ctx, uint64_t m, const uint32_t* body, const float* rel_xy, const float* vel_xy) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  CK(cudaStreamSynchronize(ctx->stream));  // the copies below are synchronous and must not overtake queued work
  if (m > ctx->cap_elec) return fail(ctx, PSIM_E_ARG, "psim_upload_electrons: m exceeds max_electrons");
  if (m && (!body || !rel_xy)) return fail(ctx, PSIM_E_ARG, "psim_upload_electrons: null array");
  const uint32_t n = ctx->n;
  // group by body on the host (stable): this is data layout, not path arithmetic
  std::vector<uint8_t> cnt(n ? n : 1, 0);
  std::vector<uint32_t> off((size_t)n + 1, 0);
  for (uint64_t k = 0; k < m; ++k) {
    if (body[k] >= n) return fail(ctx, PSIM_E_ARG, "psim_upload_electrons: body index out of range");
    if (cnt[body[k]] == 255) return fail(ctx, PSIM_E_ARG, "psim_upload_electrons: more than 255 electrons on a body");
    cnt[body[k]]++;
  }
  for (uint32_t i = 0; i < n; ++i) off[i + 1] = off[i] + cnt[i];
  std::vector<uint32_t> cursor(off.begin(), off.end() - 1), gbody(m ? m : 1);
  std::vector<float> grel(2 * (m ? m : 1)), gvel(2 * (m ? m : 1), 0.0f);
  for (uint64_t k = 0; k < m; ++k) {
    const uint32_t d = cursor[body[k]]++;
    gbody[d] = body[k];
    grel[2 * d] = rel_xy[2 * k], grel[2 * d + 1] = rel_xy[2 * k + 1];
    if (vel_xy) gvel[2 * d] = vel_xy[2 * k], gvel[2 * d + 1] = vel_xy[2 * k + 1];
  }
  const int e = ctx->ecur;
  if (n) {
    CK(cudaMemcpy(ctx->b[ctx->cur].ecount, cnt.data(), n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->eoff[e], off.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  }
  if (m) {
    CK(cudaMemcpy(ctx->ebody[e], gbody.data(), m * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->erel[e], grel.data(), m * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->evel[e], gvel.data(), m * 8, cudaMemcpyHostToDevice));
  }
  ctx->m = (uint32_t)m;
  return PSIM_OK;
}

int32_t psim_download_bodies(psim_ctx* ctx, float* pos_xy, float* z, float* vel_xy, float* vz, float* acc_xy,
                             float* az, float* mass, float* radius, float* charge, uint8_t* species,
                             float* e_field_xy, uint32_t* orig_index) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  const uint64_t n = ctx->n;
  if (n == 0) return PSIM_OK;
  ctx->host_map_valid = false;  // the caller now holds the device's row order
  cudaStream_t st = ctx->stream;
  const size_t o_pos = 0, o_vel = o_pos + align256(8 * n), o_acc = o_vel + align256(8 * n),
               o_z = o_acc + align256(8 * n), o_vz = o_z + align256(4 * n), o_az = o_vz + align256(4 * n),
               o_m = o_az + align256(4 * n), o_r = o_m + align256(4 * n), o_q = o_r + align256(4 * n),
               total = o_q + align256(4 * n);
  int32_t rc = ensure_stage(ctx, total);
  if (rc) return rc;
  char* sb = static_cast<char*>(ctx->stage);
  RawOut o;
  o.pos = pos_xy ? reinterpret_cast<float2*>(sb + o_pos) : nullptr;
  o.vel = vel_xy ? reinterpret_cast<float2*>(sb + o_vel) : nullptr;
  o.acc = acc_xy ? reinterpret_cast<float2*>(sb + o_acc) : nullptr;
  o.z = z ? reinterpret_cast<float*>(sb + o_z) : nullptr;
  o.vz = vz ? reinterpret_cast<float*>(sb + o_vz) : nullptr;
  o.az = az ? reinterpret_cast<float*>(sb + o_az) : nullptr;
  o.mass = mass ? reinterpret_cast<float*>(sb + o_m) : nullptr;
  o.radius = radius ? reinterpret_cast<float*>(sb + o_r) : nullptr;
  o.charge = charge ? reinterpret_cast<float*>(sb + o_q) : nullptr;
  BodyArrays& b = ctx->b[ctx->cur];
  unpack_bodies_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(b, (uint32_t)n, o);
  LAUNCHED(ctx);
  auto down = [&](void* dst, const void* src, size_t bytes) {
    if (dst) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
  };
  down(pos_xy, o.pos, 8 * n), down(vel_xy, o.vel, 8 * n), down(acc_xy, o.acc, 8 * n);
  down(z, o.z, 4 * n), down(vz, o.vz, 4 * n), down(az, o.az, 4 * n), down(mass, o.mass, 4 * n);
  down(radius, o.radius, 4 * n), down(charge, o.charge, 4 * n);
  down(species, b.species, n), down(e_field_xy, b.efield, 8 * n), down(orig_index, b.orig, 4 * n);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  return PSIM_OK;
}

int32_t psim_download_electrons(psim_ctx* ctx, uint32_t* body, float* rel_xy, float* vel_xy) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  const uint64_t m = ctx->m;
  if (m == 0) return PSIM_OK;
  const int e = ctx->ecur;
  CK(cudaStreamSynchronize(ctx->stream));
  if (body) CK(cudaMemcpy(body, ctx->ebody[e], m * 4, cudaMemcpyDeviceToHost));
  if (rel_xy) CK(cudaMemcpy(rel_xy, ctx->erel[e], m * 8, cudaMemcpyDeviceToHost));
  if (vel_xy) CK(cudaMemcpy(vel_xy, ctx->evel[e], m * 8, cudaMemcpyDeviceToHost));
  return PSIM_OK;
}

int32_t psim_build(psim_ctx* ctx, int32_t mode, float hw, float hh) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (mode != PSIM_BUILD_CONTAINING && mode != PSIM_BUILD_DOMAIN) return fail(ctx, PSIM_E_ARG, "psim_build: mode");
  int32_t rc = build_async(ctx, mode, hw, hh);
  if (rc) return rc;
  CK(cudaGetLastError());
  return check_build(ctx);
}

int32_t psim_build_async(psim_ctx* ctx, int32_t mode, float hw, float hh) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (mode != PSIM_BUILD_CONTAINING && mode != PSIM_BUILD_DOMAIN) return fail(ctx, PSIM_E_ARG, "psim_build: mode");
  int32_t rc = build_async(ctx, mode, hw, hh);
  if (rc) return rc;
  CK(cudaGetLastError());
  return PSIM_OK;
}

int32_t psim_build_status(psim_ctx* ctx) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  return check_build(ctx);
}

int32_t psim_get_permutation(psim_ctx* ctx, uint32_t* out) {
  if (!ctx || !out) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (!ctx->perm_valid && ctx->n) return fail(ctx, PSIM_E_STATE, "psim_get_permutation: no build since the last upload");
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->n) CK(cudaMemcpy(out, ctx->perm, (size_t)ctx->n * 4, cudaMemcpyDeviceToHost));
  return PSIM_OK;
}

int32_t psim_get_keys(psim_ctx* ctx, uint64_t* out) {
  if (!ctx || !out) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (!ctx->perm_valid && ctx->n) return fail(ctx, PSIM_E_STATE, "psim_get_keys: no build since the last upload");
  if (ctx->sh.tree_is_sharded) return fail(ctx, PSIM_E_STATE, "psim_get_keys: the last build was sharded (each rank holds its own keys only)");
  if (!ctx->n) return PSIM_OK;
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(out, ctx->keys[1], (size_t)ctx->n * 8, cudaMemcpyDeviceToHost));
  return PSIM_OK;
}

int32_t psim_download_nodes(psim_ctx* ctx, psim_node* out, uint64_t cap, uint64_t* count) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (!ctx->perm_valid && ctx->n) return fail(ctx, PSIM_E_STATE, "psim_download_nodes: no tree");
  if (ctx->sh.tree_is_sharded)
    return fail(ctx, PSIM_E_STATE, "psim_download_nodes: the last build was sharded (each rank holds its piece of the node array only)");
  if (count) *count = 0;
  if (ctx->n == 0) return PSIM_OK;
  int32_t rc = fetch_meta(ctx);
  if (rc) return rc;
  const uint32_t M = ctx->meta_h.num_nodes;
  const uint64_t total = 4ull * ctx->meta_h.num_internal + 1ull;
  if (count) *count = total;
  if (!out || cap == 0) return PSIM_OK;
  cudaStream_t st = ctx->stream;
  {
    // export sweep: parent links, node masses, body counts, centres of chargeless nodes
    BodyArrays& b = ctx->b[ctx->cur];
    export_root_leaf_kernel<<<1, 32, 0, st>>>(ctx->meta, b.accm, ctx->t);
    export_reset_levels_kernel<<<1, 32, 0, st>>>(ctx->meta, 0, ctx->node_cap);
    export_count_levels_kernel<<<grid_for(ctx, M, 256, 8), 256, 0, st>>>(ctx->meta, ctx->t.nodeB);
    export_reset_levels_kernel<<<1, 32, 0, st>>>(ctx->meta, 1, ctx->node_cap);
    export_fill_levels_kernel<<<grid_for(ctx, M, 256, 8), 256, 0, st>>>(ctx->meta, ctx->t);
    ctx->launches += 5;
    if (ctx->cfg.strict_centres) {  // chargeless nodes: the reference's mass / centroid sums (strict.cuh)
      strict_chargeless_kernel<<<grid_for(ctx, M, 128, 16), 128, 0, st>>>(ctx->meta, b.pqr, b.accm, ctx->t);
      LAUNCHED(ctx);
    }
    for (int level = kMaxLevels - 1; level >= 0; --level) {
      export_level_kernel<<<ctx->sm_count * 8, 128, 0, st>>>(level, ctx->meta, b.pqr, b.accm, ctx->t,
                                                              !ctx->cfg.strict_centres);
      LAUNCHED(ctx);
    }
  }
  CK(exclusive_scan(InternalFlagFn{ctx->t.nodeB}, M, ctx->irank, ctx->scan_partials, nullptr, st));
  ctx->launches += 3;
  const uint64_t ncopy = total < cap ? total : cap;
  rc = ensure_qstage(ctx, ncopy * sizeof(PsimNodeOut));
  if (rc) return rc;
  CK(cudaMemsetAsync(ctx->qstage, 0, ncopy * sizeof(PsimNodeOut), st));
  export_nodes_kernel<<<grid_for(ctx, M, 128, 16), 128, 0, st>>>(
      ctx->keys[1], ctx->keys[1], ctx->tree_plan, kTreePasses, ctx->meta, ctx->t, ctx->irank,
      static_cast<PsimNodeOut*>(ctx->qstage), ncopy);
  LAUNCHED(ctx);
  CK(cudaMemcpyAsync(out, ctx->qstage, ncopy * sizeof(PsimNodeOut), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return PSIM_OK;
}

int32_t psim_field(psim_ctx* ctx, float k_e, float bg_x, float bg_y, int32_t write_acc, float* out_e, float* out_acc) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  int32_t rc = field_async(ctx, k_e, bg_x, bg_y, write_acc);
  if (rc) return rc;
  CK(cudaGetLastError());
  if (out_e || out_acc)
    return psim_download_bodies(ctx, nullptr, nullptr, nullptr, nullptr, out_acc, nullptr, nullptr, nullptr, nullptr,
                                nullptr, out_e, nullptr);
  return PSIM_OK;
}

int32_t psim_acc_points(psim_ctx* ctx, uint64_t m, const float* pts_xy, const float* q, const float* radius,
                        float k_e, float* out_xy) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (m == 0) return PSIM_OK;
  if (ctx->sh.tree_is_sharded && ctx->sh.tree_is_let)
    return fail(ctx, PSIM_E_STATE, "psim_acc_points: the last sharded build kept only the records this rank's own targets "
                                   "can reach (locally essential tree); set PSIM_LET=0 to query arbitrary points");
  if (!pts_xy || !out_xy || m >= (1ull << 30)) return fail(ctx, PSIM_E_ARG, "psim_acc_points: null array or m too large");
  const size_t o_p = 0, o_q = o_p + align256(8 * m), o_r = o_q + align256(4 * m), o_o = o_r + align256(4 * m),
               total = o_o + align256(8 * m);
  int32_t rc = ensure_qstage(ctx, total);
  if (rc) return rc;
  char* sb = static_cast<char*>(ctx->qstage);
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(sb + o_p, pts_xy, 8 * m, cudaMemcpyHostToDevice, st));
  if (q) CK(cudaMemcpyAsync(sb + o_q, q, 4 * m, cudaMemcpyHostToDevice, st));
  if (radius) CK(cudaMemcpyAsync(sb + o_r, radius, 4 * m, cudaMemcpyHostToDevice, st));
  rc = points_async(ctx, reinterpret_cast<const float2*>(sb + o_p), q ? reinterpret_cast<const float*>(sb + o_q) : nullptr,
                    radius ? reinterpret_cast<const float*>(sb + o_r) : nullptr, (uint32_t)m, k_e,
                    reinterpret_cast<float2*>(sb + o_o));
  if (rc) return rc;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_xy, sb + o_o, 8 * m, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return PSIM_OK;
}

int32_t psim_collide(psim_ctx* ctx, float hw, float hh, float domain_depth, uint32_t passes, uint32_t num_passes,
                     float li_collision_softness, int32_t soft_li, int32_t soft_an, uint64_t* touching_pairs) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (touching_pairs) *touching_pairs = 0;
  const uint32_t n = ctx->n;
  if (n == 0 || passes == 0) return PSIM_OK;
  if (num_passes == 0) return fail(ctx, PSIM_E_ARG, "psim_collide: num_passes must be positive");
  // broad-phase cell: the largest diameter among the species present (bodies carry their species' radius)
  float rmax = 0.0f;
  for (uint32_t s = 0; s < ctx->nspecies; ++s)
    if (ctx->species_present & (1u << s)) rmax = fmaxf(rmax, ctx->table_h[s].radius);
  if (!(rmax > 0.0f)) rmax = 1.0f;
  cudaStream_t st = ctx->stream;
  int32_t rc = ensure_stage(ctx, 2 * (size_t)n * sizeof(float4) + 256);
  if (rc) return rc;
  float4* recA = static_cast<float4*>(ctx->stage);
  float4* recB = recA + n;
  unsigned long long* counter = reinterpret_cast<unsigned long long*>(recB + n);
  CollideParams P;
  P.correction_scale = 1.0f / (float)num_passes;
  P.softness = fminf(fmaxf(li_collision_softness, 0.0f), 1.0f);
  P.soft_li = soft_li ? 1u : 0u, P.soft_an = soft_an ? 1u : 0u;
  P.domain_depth = domain_depth;
  uint32_t first, count;
  body_range(ctx, first, count);
  for (uint32_t pass = 0; pass < passes; ++pass) {
    if ((rc = cell_build_async(ctx, hw, hh, 2.0f * rmax))) return rc;
    BodyArrays& b = ctx->b[ctx->cur];
    P.g = ctx->grid;
    collide_records_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(ctx->order, n, b.pqr, b.velz, b.accm, recA, recB);
    CK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    if (count)
      collide_kernel<<<(count + 127) / 128, 128, 0, st>>>(first, first + count, ctx->cell_start, ctx->cell_end,
                                                          ctx->body_cell, ctx->cpos, recA, recB, b.species, b.accm, P,
                                                          b.pqr, b.velz, counter);
    ctx->launches += 2;
    ctx->tree_valid = ctx->grid_valid = false;
  }
  CK(cudaGetLastError());
  if (touching_pairs) {
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, counter, sizeof h, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *touching_pairs = h / 2;
  }
  return PSIM_OK;
}

int32_t psim_hop_alignment(psim_ctx* ctx, uint64_t m_src, const uint32_t* src_idx, const uint32_t* pair_offsets,
                           const uint32_t* dst_idx, float k_e, float bg_x, float bg_y, float alignment_bias,
                           float* out_local_field_xy, float* out_alignment) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (m_src == 0) return PSIM_OK;
  if (!src_idx || !pair_offsets || m_src >= (1ull << 30)) return fail(ctx, PSIM_E_ARG, "psim_hop_alignment: null array or too many donors");
  const uint64_t np = pair_offsets[m_src];
  if (np >= (1ull << 31) || (np && (!dst_idx || !out_alignment))) return fail(ctx, PSIM_E_ARG, "psim_hop_alignment: pair arrays");
  if (!ctx->tree_valid) return fail(ctx, PSIM_E_STATE, "psim_hop_alignment: no tree (call psim_build first)");
  if (ctx->sh.tree_is_sharded && ctx->sh.tree_is_let)
    return fail(ctx, PSIM_E_STATE, "psim_hop_alignment: locally essential tree (see psim_acc_points)");
  const uint32_t m = (uint32_t)m_src;
  const size_t o_src = 0, o_off = o_src + align256(4 * m_src), o_dst = o_off + align256(4 * (m_src + 1)),
               o_pts = o_dst + align256(4 * np), o_fld = o_pts + align256(8 * m_src), o_loc = o_fld + align256(8 * m_src),
               o_al = o_loc + align256(8 * m_src), total = o_al + align256(4 * np);
  int32_t rc = ensure_qstage(ctx, total);
  if (rc) return rc;
  char* sb = static_cast<char*>(ctx->qstage);
  cudaStream_t st = ctx->stream;
  BodyArrays& b = ctx->b[ctx->cur];
  CK(cudaMemcpyAsync(sb + o_src, src_idx, 4 * m_src, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(sb + o_off, pair_offsets, 4 * (m_src + 1), cudaMemcpyHostToDevice, st));
  if (np) CK(cudaMemcpyAsync(sb + o_dst, dst_idx, 4 * np, cudaMemcpyHostToDevice, st));
  hop_points_kernel<<<grid_for(ctx, m, 256, 8), 256, 0, st>>>(b.pqr, reinterpret_cast<const uint32_t*>(sb + o_src), m, ctx->n,
                                                              reinterpret_cast<float2*>(sb + o_pts));
  LAUNCHED(ctx);
  rc = points_async(ctx, reinterpret_cast<const float2*>(sb + o_pts), nullptr, nullptr, m, k_e,
                    reinterpret_cast<float2*>(sb + o_fld));
  if (rc) return rc;
  hop_alignment_kernel<<<grid_for(ctx, m, 128, 8), 128, 0, st>>>(
      b.pqr, b.species, reinterpret_cast<const uint32_t*>(sb + o_src), reinterpret_cast<const uint32_t*>(sb + o_off),
      reinterpret_cast<const uint32_t*>(sb + o_dst), m, ctx->n, reinterpret_cast<const float2*>(sb + o_fld), bg_x, bg_y,
      alignment_bias, reinterpret_cast<float2*>(sb + o_loc), reinterpret_cast<float*>(sb + o_al));
  LAUNCHED(ctx);
  CK(cudaGetLastError());
  if (out_local_field_xy) CK(cudaMemcpyAsync(out_local_field_xy, sb + o_loc, 8 * m_src, cudaMemcpyDeviceToHost, st));
  if (np) CK(cudaMemcpyAsync(out_alignment, sb + o_al, 4 * np, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return PSIM_OK;
}

int32_t psim_update_electrons(psim_ctx* ctx, float bg_x, float bg_y, float dt, float k_e) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  int32_t rc = electrons_async(ctx, bg_x, bg_y, dt, k_e);
  if (rc) return rc;
  CK(cudaGetLastError());
  return PSIM_OK;
}

int32_t psim_cell_build(psim_ctx* ctx, float hw, float hh, float cell_size) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  int32_t rc = cell_build_async(ctx, hw, hh, cell_size);
  if (rc) return rc;
  CK(cudaGetLastError());
  return PSIM_OK;
}

int32_t psim_cell_download(psim_ctx* ctx, uint64_t* gx, uint64_t* gy, uint32_t* offsets, uint32_t* indices) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (!ctx->grid_valid) return fail(ctx, PSIM_E_STATE, "psim_cell_download: no cell grid");
  if (gx) *gx = ctx->grid.gx;
  if (gy) *gy = ctx->grid.gy;
  const uint64_t ncells = (uint64_t)ctx->grid.gx * ctx->grid.gy;
  CK(cudaStreamSynchronize(ctx->stream));
  if (offsets) {
    // cells are laid out in cell-id order in `order`, so offsets are a running sum of the counts
    std::vector<uint32_t> s(ncells), e(ncells);
    CK(cudaMemcpy(s.data(), ctx->cell_start, ncells * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(e.data(), ctx->cell_end, ncells * 4, cudaMemcpyDeviceToHost));
    uint32_t run = 0;
    for (uint64_t c = 0; c < ncells; ++c) {
      offsets[c] = run;
      run += e[c] - s[c];
    }
    offsets[ncells] = run;
  }
  if (indices && ctx->n) CK(cudaMemcpy(indices, ctx->order, (size_t)ctx->n * 4, cudaMemcpyDeviceToHost));
  return PSIM_OK;
}

int32_t psim_neighbors_within(psim_ctx* ctx, uint64_t m, const uint32_t* body_idx, float cutoff, int32_t metals_only,
                              uint32_t* offsets, uint32_t* indices, uint64_t indices_cap, uint64_t* total) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (!ctx->grid_valid) return fail(ctx, PSIM_E_STATE, "psim_neighbors_within: no cell grid");
  if (total) *total = 0;
  if (m == 0) {
    if (offsets) offsets[0] = 0;
    return PSIM_OK;
  }
  if (!body_idx || !offsets || m >= (1ull << 30)) return fail(ctx, PSIM_E_ARG, "psim_neighbors_within: null array");
  cudaStream_t st = ctx->stream;
  BodyArrays& b = ctx->b[ctx->cur];
  // qstage: query | counts | offsets(m+1) | partials | [indices]
  const size_t o_qry = 0, o_cnt = o_qry + align256(4 * m), o_off = o_cnt + align256(4 * m),
               o_par = o_off + align256(4 * (m + 1)), o_end = o_par + align256(4 * ((size_t)scan_num_tiles((uint32_t)m) + 1));
  int32_t rc = ensure_qstage(ctx, o_end);
  if (rc) return rc;
  char* sb = static_cast<char*>(ctx->qstage);
  uint32_t* d_q = reinterpret_cast<uint32_t*>(sb + o_qry);
  uint32_t* d_cnt = reinterpret_cast<uint32_t*>(sb + o_cnt);
  uint32_t* d_off = reinterpret_cast<uint32_t*>(sb + o_off);
  uint32_t* d_par = reinterpret_cast<uint32_t*>(sb + o_par);
  CK(cudaMemcpyAsync(d_q, body_idx, 4 * m, cudaMemcpyHostToDevice, st));
  const int blocks = (int)((m + 127) / 128);
  cell_neighbors_kernel<<<blocks, 128, 0, st>>>(b.pqr, b.species, ctx->n, ctx->cell_start, ctx->cell_end, ctx->order,
                                                ctx->grid, d_q, (uint32_t)m, cutoff, metals_only, d_cnt, nullptr, nullptr);
  LAUNCHED(ctx);
  CK(exclusive_scan(U32Fn{d_cnt}, (uint32_t)m, d_off, d_par, d_off + m, st));
  ctx->launches += 3;
  CK(cudaMemcpyAsync(offsets, d_off, 4 * (m + 1), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const uint64_t tot = offsets[m];
  if (total) *total = tot;
  if (!indices || tot == 0) return PSIM_OK;
  if (tot > indices_cap) return fail(ctx, PSIM_E_ARG, "psim_neighbors_within: indices_cap too small (see *total)");
  uint32_t* d_idx = nullptr;
  CK(cudaMalloc(&d_idx, tot * 4));
  cell_neighbors_kernel<<<blocks, 128, 0, st>>>(b.pqr, b.species, ctx->n, ctx->cell_start, ctx->cell_end, ctx->order,
                                                ctx->grid, d_q, (uint32_t)m, cutoff, metals_only, d_cnt, d_off, d_idx);
  LAUNCHED(ctx);
  cudaError_t e = cudaMemcpyAsync(indices, d_idx, tot * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_idx);
  if (e != cudaSuccess) return fail(ctx, PSIM_E_CUDA, "psim_neighbors_within", e);
  return PSIM_OK;
}

int32_t psim_reset_acc(psim_ctx* ctx) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (ctx->n == 0) return PSIM_OK;
  reset_acc_kernel<<<grid_for(ctx, ctx->n, 256, 16), 256, 0, ctx->stream>>>(ctx->b[ctx->cur].accm, ctx->n);
  LAUNCHED(ctx);
  CK(cudaGetLastError());
  return PSIM_OK;
}

int32_t psim_use_cell_list(const psim_ctx* ctx, float hw, float hh, float density_threshold) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  const float area = (2.0f * hw) * (2.0f * hh);  // simulation.rs:1798-1802
  const float density = (float)ctx->n / area;
  return density > density_threshold ? 1 : 0;
}

int32_t psim_prepare_spatial_structures(psim_ctx* ctx, float hw, float hh, float density_threshold) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  int32_t rc = build_async(ctx, PSIM_BUILD_CONTAINING, 0.f, 0.f);
  if (rc) return rc;
  // forces.rs:17-24.  Below the density threshold the reference answers neighbour queries from the
  // tree; the result sets are the same, so the grid is built in both cases.
  (void)density_threshold;
  const float lj_cutoff = max_lj_cutoff(ctx), repulsion_cutoff = max_repulsion_cutoff(ctx);
  const float polar_cutoff = 3.0f * lj_cutoff;
  const float max_cutoff = fmaxf(fmaxf(polar_cutoff, repulsion_cutoff), lj_cutoff);
  if (max_cutoff > 0.0f) {
    rc = cell_build_async(ctx, hw, hh, max_cutoff);
    if (rc) return rc;
  }
  CK(cudaGetLastError());
  return check_build(ctx);
}

int32_t psim_short_range(psim_ctx* ctx, uint32_t flags) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  int32_t rc = short_range_async(ctx, flags);
  if (rc) return rc;
  CK(cudaGetLastError());
  return PSIM_OK;
}

int32_t psim_apply_polar_forces(psim_ctx* ctx, float k_e, int32_t dipole_model) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (dipole_model != 0 && dipole_model != 1) return fail(ctx, PSIM_E_ARG, "psim_apply_polar_forces: dipole_model");
  int32_t rc = polar_async(ctx, k_e, dipole_model);
  if (rc) return rc;
  CK(cudaGetLastError());
  return PSIM_OK;
}

int32_t psim_iterate(psim_ctx* ctx, float dt, float damping_base, float hw, float hh, float hd, int32_t enable_z) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  int32_t rc = iterate_async(ctx, dt, damping_base, hw, hh, hd, enable_z);
  if (rc) return rc;
  CK(cudaGetLastError());
  return PSIM_OK;
}

// ---- multi-GPU orchestration (one context per rank; replaces the Python loop of particlesim_b200/parallel.py) -----
#define NCK(call)                                                                          \
  do {                                                                                     \
    ncclResult_t _r = (call);                                                              \
    if (_r != ncclSuccess) {                                                               \
      ctx->err = std::string(#call) + ": " + nccl_api().GetErrorString(_r);               \
      return PSIM_E_NCCL;                                                                  \
    }                                                                                      \
  } while (0)

static uint32_t shard_width(uint64_t n, uint32_t world) {  // bodies per rank, a multiple of 64 (two walk groups)
  return n ? (uint32_t)(((n + world - 1) / world + 63) / 64 * 64) : 0u;
}

// In-place all-gather of UNEVEN row segments: rank r contributes rows [lo[r], lo[r+1]) of `full` (row_bytes each,
// `rows_cap` rows allocated).  One ncclAllGather of max-segment blocks into a staging area, then the valid rows of
// the other ranks are copied into place.
static int32_t all_gatherv(psim_ctx* ctx, void* full, size_t row_bytes, size_t rows_cap, const uint32_t* lo,
                           cudaStream_t st) {
  auto& K = ctx->comm;
  const uint32_t world = ctx->sh.world, rank = ctx->sh.rank;
  size_t seg = 0;
  for (uint32_t r = 0; r < world; ++r) seg = std::max<size_t>(seg, lo[r + 1] - lo[r]);
  if (seg == 0) return PSIM_OK;
  seg = (seg + 1023) / 1024 * 1024;
  const size_t need = (size_t)(world + 1) * seg * row_bytes;
  if (need > K.stage_bytes) {
    if (K.stage) {
      CK(cudaStreamSynchronize(ctx->stream));
      CK(cudaStreamSynchronize(K.side));
      cudaFree(K.stage);
    }
    K.stage = nullptr, K.stage_bytes = 0;
    const size_t want = need + need / 4;
    if (cudaMalloc(&K.stage, want) != cudaSuccess) return fail(ctx, PSIM_E_OOM, "all-gather staging");
    K.stage_bytes = want;
  }
  char* stage = static_cast<char*>(K.stage);
  char* base = static_cast<char*>(full);
  const size_t a = lo[rank], cnt = lo[rank + 1] - lo[rank];
  const char* mine = base + a * row_bytes;
  if (a + seg > rows_cap) {  // the padded block would run past the array: go through the spare block
    char* spare = stage + (size_t)world * seg * row_bytes;
    if (cnt) CK(cudaMemcpyAsync(spare, mine, cnt * row_bytes, cudaMemcpyDeviceToDevice, st));
    mine = spare;
  }
  NCK(nccl_api().AllGather(mine, stage, seg * row_bytes, ncclChar, K.comm, st));
  for (uint32_t r = 0; r < world; ++r) {
    const size_t c = lo[r + 1] - lo[r];
    if (r != rank && c)
      CK(cudaMemcpyAsync(base + (size_t)lo[r] * row_bytes, stage + (size_t)r * seg * row_bytes, c * row_bytes,
                         cudaMemcpyDeviceToDevice, st));
  }
  return PSIM_OK;
}

// equal slices of `width` rows, in place (rank r owns rows [r * width, (r + 1) * width))
static int32_t all_gather_slices(psim_ctx* ctx, void* full, size_t row_bytes, uint32_t width, cudaStream_t st) {
  if (width == 0) return PSIM_OK;
  char* base = static_cast<char*>(full);
  NCK(nccl_api().AllGather(base + (size_t)ctx->sh.rank * width * row_bytes, base, (size_t)width * row_bytes, ncclChar,
                           ctx->comm.comm, st));
  return PSIM_OK;
}

// Quadtree::build / build_with_domain across the communicator: the seven phases of shard_phase with the exchanges
// between them.  `cell_size` > 0: the cell list is rebuilt while the traversal pieces travel (it only needs the
// sorted bodies).
static int32_t build_sharded(psim_ctx* ctx, int mode, float hw, float hh, const psim_step_params* p, float cell_size) {
  auto& K = ctx->comm;
  auto& S = ctx->sh;
  cudaStream_t st = ctx->stream;
  int32_t rc;
  uint32_t lo[kMaxRanks + 1], tl[kMaxRanks + 1];
  if ((rc = shard_phase(ctx, 0, mode, hw, hh, lo))) return rc;
  if ((rc = shard_phase(ctx, 1, mode, hw, hh, nullptr))) return rc;
  if ((rc = all_gatherv(ctx, ctx->keys_idx_all, sizeof(uint32_t), ctx->cap_bodies, lo, st))) return rc;
  if (K.vel_pending) {  // the velocities all-gathered behind the first two phases must be in place before the gather
    CK(cudaStreamWaitEvent(st, K.ev_vel, 0));
    K.vel_pending = false;
  }
  if ((rc = shard_phase(ctx, 2, mode, hw, hh, nullptr))) return rc;
  const uint32_t world = S.world;
  if (K.let) {  // who can reach what: the bins of every rank's targets (bodies are in sorted order from here on)
    const uint32_t wb = shard_width(ctx->n, world), we = shard_width(ctx->m, world);
    let_regions_kernel<<<1, kMaxRanks, 0, st>>>(S.binprefix, ctx->n, ctx->m, wb, we, ctx->m ? ctx->ebody[ctx->ecur] : nullptr,
                                                 world, K.regions);
    LAUNCHED(ctx);
  }
  NCK(nccl_api().AllReduce(S.xbuf, S.xbuf, kBins + kMaxRanks, ncclUint64, ncclSum, K.comm, st));
  if ((rc = shard_phase(ctx, 3, mode, hw, hh, nullptr))) return rc;
  NCK(nccl_api().AllReduce(S.heap, S.heap, (size_t)kTopSlots * (sizeof(TopRec) / 8), ncclUint64, ncclSum, K.comm, st));
  if ((rc = shard_phase(ctx, 4, mode, hw, hh, nullptr))) return rc;
  NCK(nccl_api().AllReduce(S.xbuf, S.xbuf, kBins + kMaxRanks, ncclUint64, ncclSum, K.comm, st));
  if ((rc = shard_phase(ctx, 5, mode, hw, hh, tl))) return rc;
  // traversal pieces on the side stream, the cell list on the main one
  CK(cudaEventRecord(K.ev_main, st));
  CK(cudaStreamWaitEvent(K.side, K.ev_main, 0));
  bool full = !K.let;
  uint32_t h_cnt[kMaxRanks * kMaxRanks];
  if (K.let) {
    // each rank selects, per destination, the records that destination's walks can reach (let.cuh)
    CK(cudaMemsetAsync(K.cnt, 0, kMaxRanks * sizeof(uint32_t), K.side));
    const float margin_abs = 8.0f;  // largest body radius (4.7 A) + largest electron offset (2.2 A), rounded up
    let_select_kernel<<<grid_for(ctx, S.meta_h.T_local + 1, 256, 8), 256, 0, K.side>>>(
        S.meta, ctx->travA, ctx->travB, S.lkeys, K.regions, ctx->meta, ctx->cfg.theta, margin_abs, K.cap_per_rank, K.send,
        K.cnt);
    LAUNCHED(ctx);
    NCK(nccl_api().AllGather(K.cnt, K.cnt_all, world, ncclUint32, K.comm, K.side));
    CK(cudaMemcpyAsync(h_cnt, K.cnt_all, (size_t)world * world * sizeof(uint32_t), cudaMemcpyDeviceToHost, K.side));
  }
  if (cell_size > 0.0f && p && (rc = cell_build_async(ctx, p->hw, p->hh, cell_size))) return rc;
  if (K.let) {
    CK(cudaStreamSynchronize(K.side));
    K.let_sent = 0, K.let_full = 0;
    for (uint32_t a = 0; a < world; ++a)
      for (uint32_t b = 0; b < world; ++b)
        if (a != b) {
          if (h_cnt[a * world + b] > K.cap_per_rank) full = true;  // a send area overflowed somewhere: everyone falls back
          if (a == S.rank) K.let_sent += h_cnt[a * world + b];
        }
    K.let_full = (uint64_t)(tl[S.rank + 1] - tl[S.rank]) * (world - 1);
  }
  S.tree_is_let = !full;
  if (full) {
    if ((rc = all_gatherv(ctx, ctx->travA, sizeof(float4), ctx->node_cap, tl, K.side))) return rc;
    if ((rc = all_gatherv(ctx, ctx->travB, sizeof(uint4), ctx->node_cap, tl, K.side))) return rc;
  } else {
    if (K.let_poison) {
      let_poison_kernel<<<grid_for(ctx, ctx->node_cap, 256, 8), 256, 0, K.side>>>(S.meta, ctx->travA, ctx->travB);
      LAUNCHED(ctx);
    }
    NCK(nccl_api().GroupStart());
    size_t roff = 0;
    for (uint32_t peer = 0; peer < world; ++peer) {
      if (peer == S.rank) continue;
      const uint32_t ns = h_cnt[S.rank * world + peer], nr = h_cnt[peer * world + S.rank];
      if (ns) NCK(nccl_api().Send(K.send + (size_t)peer * K.cap_per_rank, (size_t)ns * sizeof(LetRec), ncclChar, (int)peer, K.comm, K.side));
      if (nr) NCK(nccl_api().Recv(K.recv + roff, (size_t)nr * sizeof(LetRec), ncclChar, (int)peer, K.comm, K.side));
      roff += nr;
    }
    NCK(nccl_api().GroupEnd());
    if (roff) {
      let_scatter_kernel<<<grid_for(ctx, roff, 256, 8), 256, 0, K.side>>>(K.recv, (uint32_t)roff, ctx->node_cap, ctx->travA,
                                                                         ctx->travB);
      LAUNCHED(ctx);
    }
  }
  CK(cudaEventRecord(K.ev_side, K.side));
  CK(cudaStreamWaitEvent(st, K.ev_side, 0));
  return shard_phase(ctx, 6, mode, hw, hh, nullptr);
}

static int32_t step_sharded(psim_ctx* ctx, const psim_step_params* p) {
  auto& K = ctx->comm;
  auto& S = ctx->sh;
  cudaStream_t st = ctx->stream;
  const uint32_t world = S.world, rank = S.rank, n = ctx->n, m = ctx->m;
  const uint32_t wb = shard_width(n, world), we = shard_width(m, world);
  if ((uint64_t)wb * world > ctx->cap_bodies || (m && (uint64_t)we * world > ctx->cap_elec))
    return fail(ctx, PSIM_E_ARG, "psim_step_sharded: create the context with psim_shard_capacity() bodies / electrons");
  {  // this rank's targets: a contiguous slice of the Morton order, and an equal slice of the electrons
    const uint32_t f = std::min(rank * wb, n), c = std::min(wb, n - f);
    ctx->tgt_set = true, ctx->tgt_first = f, ctx->tgt_count = c;
    const uint32_t ef = std::min(rank * we, m), ec = std::min(we, m - ef);
    ctx->etgt_set = true, ctx->e_first = ef, ctx->e_count = ec;
  }
  int32_t rc;
  auto mark = [&](int k) {
    if (ctx->ev_ok) cudaEventRecord(ctx->ev[k], st);
  };
  mark(0);
  if ((rc = psim_reset_acc(ctx))) return rc;
  float cell = 0.0f;
  if (p->do_short_range) {
    cell = step_cell_size(ctx, p->do_polar != 0);
  }
  if ((rc = build_sharded(ctx, PSIM_BUILD_CONTAINING, 0.f, 0.f, p, cell))) return rc;
  mark(1);
  mark(2);
  if ((rc = field_async(ctx, p->k_e, p->bg_x, p->bg_y, 1))) return rc;
  mark(3);
  if (p->do_polar && p->do_short_range && cell > 0.0f && (rc = polar_async(ctx, p->k_e, 1))) return rc;
  if (p->do_short_range && (rc = short_range_async(ctx, PSIM_SR_LJ | PSIM_SR_REPULSION | PSIM_SR_STACK_PRESSURE))) return rc;
  mark(4);
  if (p->do_iterate) {
    if ((rc = iterate_async(ctx, p->dt, p->damping_base, p->hw, p->hh, p->hd, (int)p->enable_out_of_plane))) return rc;
    BodyArrays& b = ctx->b[ctx->cur];
    if ((rc = all_gather_slices(ctx, b.pqr, sizeof(float4), wb, st))) return rc;
    // the velocities are not needed before the next build permutes the bodies: their all-gather runs on the side
    // stream behind that build's first two phases, which only read positions
    CK(cudaEventRecord(K.ev_main, st));
    CK(cudaStreamWaitEvent(K.side, K.ev_main, 0));
    if ((rc = all_gather_slices(ctx, b.velz, sizeof(float4), wb, K.side))) return rc;
    CK(cudaEventRecord(K.ev_vel, K.side));
    K.vel_pending = true;
    ctx->tree_valid = ctx->grid_valid = false;
  }
  mark(5);
  if (p->do_electrons) {
    if ((rc = build_sharded(ctx, PSIM_BUILD_DOMAIN, p->hw, p->hh, p, 0.0f))) return rc;
    mark(6);
    if ((rc = electrons_async(ctx, p->bg_x, p->bg_y, p->dt, p->k_e))) return rc;
    if (we) {
      if ((rc = all_gather_slices(ctx, ctx->erel[ctx->ecur], sizeof(float2), we, st))) return rc;
      if ((rc = all_gather_slices(ctx, ctx->evel[ctx->ecur], sizeof(float2), we, st))) return rc;
    }
  } else {
    mark(6);
  }
  if (K.vel_pending) {
    CK(cudaStreamWaitEvent(st, K.ev_vel, 0));
    K.vel_pending = false;
  }
  mark(7);
  ctx->ev_recorded = ctx->ev_ok;
  CK(cudaGetLastError());
  return PSIM_OK;
}

static int32_t step_async(psim_ctx* ctx, const psim_step_params* p) {
  int32_t rc;
  auto mark = [&](int k) {
    if (ctx->ev_ok) cudaEventRecord(ctx->ev[k], ctx->stream);
  };
  // psim_step_host: results leave on the copy stream as soon as the main stream has produced them
  psim_ctx::HostedStep& H = ctx->hosted;
  const uint64_t n = ctx->n;
  auto results_from_here = [&]() {
    cudaEventRecord(ctx->ev_mid, ctx->stream);
    cudaStreamWaitEvent(ctx->copy_out, ctx->ev_mid, 0);
  };
  mark(0);
  if ((rc = psim_reset_acc(ctx))) return rc;
  static const bool no_overlap = getenv("PSIM_OVERLAP") && getenv("PSIM_OVERLAP")[0] == '0';
  const float cell = p->do_short_range ? step_cell_size(ctx, p->do_polar != 0) : 0.0f;
  if (!no_overlap) {
    if ((rc = overlap_ready(ctx))) return rc;
    ctx->ov.active = true;
    ctx->ov.cell_hw = p->hw, ctx->ov.cell_hh = p->hh, ctx->ov.cell_size = cell;
  }
  rc = build_async(ctx, PSIM_BUILD_CONTAINING, 0.f, 0.f);
  ctx->ov.active = false, ctx->ov.cell_size = 0.0f;
  if (rc) return rc;
  if (H.active && H.out_orig) {
    results_from_here();
    cudaMemcpyAsync(H.out_orig, ctx->b[ctx->cur].orig, 4 * n, cudaMemcpyDeviceToHost, ctx->copy_out);
  }
  mark(1);
  if (p->do_short_range) {
    // The reference sizes its grid for the polar pass too (3 x the LJ cutoff, forces.rs:17-22).  The
    // pair sets of the LJ / repulsion passes do not depend on the cell size, so the fused step bins at
    // the largest cutoff those passes use: 9x fewer candidates per body than at 3 x cutoff.
    if (cell > 0.0f && no_overlap && (rc = cell_build_async(ctx, p->hw, p->hh, cell))) return rc;
  }
  mark(2);
  if ((rc = field_async(ctx, p->k_e, p->bg_x, p->bg_y, 1))) return rc;
  if (H.active && H.out_ef) {
    results_from_here();
    cudaMemcpyAsync(H.out_ef, ctx->b[ctx->cur].efield, 8 * n, cudaMemcpyDeviceToHost, ctx->copy_out);
  }
  mark(3);
  if ((rc = overlap_join(ctx))) return rc;  // cell list and regrouped electrons
  if (p->do_polar && p->do_short_range && (rc = polar_async(ctx, p->k_e, 1))) return rc;
  if (p->do_short_range && (rc = short_range_async(ctx, PSIM_SR_LJ | PSIM_SR_REPULSION | PSIM_SR_STACK_PRESSURE))) return rc;
  mark(4);
  if (H.active && H.late_vel) {  // the velocities have had the whole force phase to arrive
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_vel, 0));
    hosted_velocities_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(H.late_vel, ctx->perm, H.map, (uint32_t)n,
                                                                               ctx->b[ctx->cur].velz);
    LAUNCHED(ctx);
    H.late_vel = nullptr;
  }
  if (p->do_iterate && (rc = iterate_async(ctx, p->dt, p->damping_base, p->hw, p->hh, p->hd, (int)p->enable_out_of_plane))) return rc;
  if (H.active && (H.out_pos || H.out_vel)) {
    results_from_here();
    BodyArrays& b = ctx->b[ctx->cur];
    hosted_unpack_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->copy_out>>>(b.pqr, b.velz, (uint32_t)n,
                                                                             H.out_pos ? H.s_pos : nullptr,
                                                                             H.out_vel ? H.s_vel : nullptr);
    LAUNCHED(ctx);
    if (H.out_pos) cudaMemcpyAsync(H.out_pos, H.s_pos, 8 * n, cudaMemcpyDeviceToHost, ctx->copy_out);
    if (H.out_vel) cudaMemcpyAsync(H.out_vel, H.s_vel, 8 * n, cudaMemcpyDeviceToHost, ctx->copy_out);
  }
  mark(5);
  if (p->do_electrons) {
    ctx->ov.active = !no_overlap;
    rc = build_async(ctx, PSIM_BUILD_DOMAIN, p->hw, p->hh);
    ctx->ov.active = false;
    if (rc) return rc;
    if ((rc = overlap_join(ctx))) return rc;  // regrouped electrons
    mark(6);
    if ((rc = electrons_async(ctx, p->bg_x, p->bg_y, p->dt, p->k_e))) return rc;
  } else {
    mark(6);
  }
  mark(7);
  ctx->ev_recorded = ctx->ev_ok;
  CK(cudaGetLastError());
  return PSIM_OK;
}

int32_t psim_step(psim_ctx* ctx, const psim_step_params* p) {
  if (!ctx || !p) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  const int32_t rc = step_async(ctx, p);
  overlap_join(ctx);  // an early error return must not leave side-stream work unordered
  return rc;
}

// One hot-path step for a caller whose bodies live in host memory: the state refresh, psim_step and
// the read-back of the results, pipelined against the device work.
//   in : positions first, on the main stream (the build needs them at once); charges and velocities
//        follow on a copy stream and are applied where they are first needed — the charges before the
//        bodies are gathered into tree order, the velocities before the integrator;
//   out: every result leaves on a second copy stream as soon as it is final — the original indices
//        after the first build, the fields after the traversal, positions and velocities after the
//        integrator — underneath the second build and the electron field sampling.
// The outputs are therefore in the body order of the step's FIRST build; the electron pass re-sorts the
// device arrays afterwards, and its permutation is kept (host_map) so that the next call, whose inputs
// are in the order of these outputs, lands every row on the right body.
int32_t psim_step_host(psim_ctx* ctx, const psim_step_params* p, uint64_t n, const float* pos_xy,
                       const float* vel_xy, const float* charge, float* out_pos_xy, float* out_vel_xy,
                       float* out_e_field_xy, uint32_t* out_orig_index) {
  if (!ctx || !p) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (n != ctx->n || (n && !pos_xy)) return fail(ctx, PSIM_E_ARG, "psim_step_host: size mismatch");
  if (ctx->tgt_set || ctx->etgt_set) return fail(ctx, PSIM_E_STATE, "psim_step_host: not available on a sharded context");
  if (n == 0) return step_async(ctx, p);  // nothing is launched
  if (!ctx->copy_in) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    bool ok = cudaStreamCreateWithPriority(&ctx->copy_in, cudaStreamNonBlocking, hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&ctx->copy_out, cudaStreamNonBlocking, hi) == cudaSuccess;
    for (cudaEvent_t* e : {&ctx->ev_start, &ctx->ev_q, &ctx->ev_vel, &ctx->ev_mid, &ctx->ev_out})
      ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->host_map, (size_t)ctx->cap_bodies * sizeof(uint32_t)) == cudaSuccess;
    if (!ok) return fail(ctx, PSIM_E_CUDA, "psim_step_host: copy streams", cudaGetLastError());
  }
  const size_t o_pos = 0, o_vel = align256(8 * n), o_q = o_vel + align256(8 * n), o_opos = o_q + align256(4 * n),
               o_ovel = o_opos + align256(8 * n), total = o_ovel + align256(8 * n);
  if (total > ctx->hstage_bytes) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->hstage) cudaFree(ctx->hstage);
    ctx->hstage = nullptr, ctx->hstage_bytes = 0;
    cudaError_t e = cudaMalloc(&ctx->hstage, total);
    if (e != cudaSuccess) return fail(ctx, PSIM_E_OOM, "psim_step_host: staging", e);
    ctx->hstage_bytes = total;
  }
  char* sb = static_cast<char*>(ctx->hstage);
  cudaStream_t st = ctx->stream;
  psim_ctx::HostedStep& H = ctx->hosted;
  H = psim_ctx::HostedStep();
  H.map = ctx->host_map_valid ? ctx->host_map : nullptr;
  H.s_pos = reinterpret_cast<float2*>(sb + o_opos), H.s_vel = reinterpret_cast<float2*>(sb + o_ovel);
  H.out_pos = out_pos_xy, H.out_vel = out_vel_xy, H.out_ef = out_e_field_xy, H.out_orig = out_orig_index;
  // the late inputs queue up behind the positions so that they do not share the link with them
  CK(cudaMemcpyAsync(sb + o_pos, pos_xy, 8 * n, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(ctx->ev_start, st));
  CK(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_start, 0));
  hosted_positions_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(reinterpret_cast<const float2*>(sb + o_pos), H.map,
                                                                     (uint32_t)n, ctx->b[ctx->cur].pqr);
  LAUNCHED(ctx);
  ctx->tree_valid = ctx->grid_valid = false;
  if (charge) {
    CK(cudaMemcpyAsync(sb + o_q, charge, 4 * n, cudaMemcpyHostToDevice, ctx->copy_in));
    CK(cudaEventRecord(ctx->ev_q, ctx->copy_in));
    H.late_q = reinterpret_cast<const float*>(sb + o_q);
  }
  if (vel_xy) {  // queued by the first build, see build_async
    H.vel_src = vel_xy;
    H.late_vel = reinterpret_cast<const float2*>(sb + o_vel);
  }
  H.active = true;
  int32_t rc = step_async(ctx, p);
  overlap_join(ctx);
  H.active = false;
  if (rc) {
    cudaStreamSynchronize(ctx->copy_in), cudaStreamSynchronize(ctx->copy_out), cudaStreamSynchronize(st);
    return rc;
  }
  if (p->do_electrons) {  // the device rows moved once more: remember where the caller's rows went
    CK(cudaMemcpyAsync(ctx->host_map, ctx->perm, 4 * n, cudaMemcpyDeviceToDevice, st));
    ctx->host_map_valid = true;
  }
  CK(cudaEventRecord(ctx->ev_out, ctx->copy_out));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  CK(cudaEventSynchronize(ctx->ev_out));
  return check_build(ctx);
}

int32_t psim_set_target_range(psim_ctx* ctx, uint64_t first, uint64_t count) {
  if (!ctx || first > ctx->n || count > ctx->n - first) return fail(ctx, PSIM_E_ARG, "psim_set_target_range: out of range");
  DeviceGuard guard(ctx->device);
  ctx->tgt_set = true, ctx->tgt_first = (uint32_t)first, ctx->tgt_count = (uint32_t)count;
  return PSIM_OK;
}
int32_t psim_set_electron_range(psim_ctx* ctx, uint64_t first, uint64_t count) {
  if (!ctx || first > ctx->m || count > ctx->m - first) return fail(ctx, PSIM_E_ARG, "psim_set_electron_range: out of range");
  ctx->etgt_set = true, ctx->e_first = (uint32_t)first, ctx->e_count = (uint32_t)count;
  return PSIM_OK;
}
int32_t psim_device_ptrs(psim_ctx* ctx, uint64_t* out8) {
  if (!ctx || !out8) return PSIM_E_ARG;
  BodyArrays& b = ctx->b[ctx->cur];
  out8[0] = (uint64_t)(uintptr_t)b.pqr, out8[1] = (uint64_t)(uintptr_t)b.velz, out8[2] = (uint64_t)(uintptr_t)b.accm;
  out8[3] = (uint64_t)(uintptr_t)b.efield;
  out8[4] = (uint64_t)(uintptr_t)ctx->erel[ctx->ecur], out8[5] = (uint64_t)(uintptr_t)ctx->evel[ctx->ecur];
  out8[6] = ctx->cap_bodies, out8[7] = ctx->cap_elec;
  return PSIM_OK;
}
// ---- neighbour-count consumers (neighbors.cuh) --------------------------------
int32_t psim_update_surrounded_flags(psim_ctx* ctx, float hw, float hh, uint64_t frame, float radius_factor,
                                     uint64_t neighbor_threshold) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (ctx->n == 0) return PSIM_OK;  // simulation.rs:1894-1896
  const float neighbor_radius = max_lj_cutoff(ctx);
  if (!(neighbor_radius > 0.0f)) return fail(ctx, PSIM_E_ARG, "psim_update_surrounded_flags: no LJ species (cell size 0)");
  // the reference bins at max_lj_cutoff when the density is above the cell-list threshold and walks the
  // tree otherwise (simulation.rs:1898-1909); the neighbour SETS are the same, so the grid serves both
  int32_t rc = cell_build_async(ctx, hw, hh, neighbor_radius);
  if (rc) return rc;
  BodyArrays& b = ctx->b[ctx->cur];
  SurroundParams P;
  P.frame = frame, P.interval = 10ull, P.neighbor_threshold = neighbor_threshold;  // config.rs:186-188
  P.radius_factor = radius_factor, P.move_threshold = 0.5f;
  surrounded_kernel<<<(int)((ctx->n + 127) / 128), 128, 0, ctx->stream>>>(
      b.pqr, b.species, b.orig, ctx->n, ctx->cell_start, ctx->cell_end, ctx->order, ctx->grid, P, ctx->surround);
  LAUNCHED(ctx);
  CK(cudaGetLastError());
  return PSIM_OK;
}

int32_t psim_get_surrounded(psim_ctx* ctx, uint8_t* flags, float* last_pos_xy, uint64_t* last_frame) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  const uint64_t n = ctx->n;
  if (n == 0) return PSIM_OK;
  const size_t o_f = 0, o_p = o_f + align256(n), o_l = o_p + align256(8 * n), total = o_l + align256(8 * n);
  int32_t rc = ensure_stage(ctx, total);
  if (rc) return rc;
  char* sb = static_cast<char*>(ctx->stage);
  cudaStream_t st = ctx->stream;
  surround_export_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(
      ctx->b[ctx->cur].orig, (uint32_t)n, ctx->surround, reinterpret_cast<uint8_t*>(sb + o_f),
      reinterpret_cast<float2*>(sb + o_p), reinterpret_cast<unsigned long long*>(sb + o_l));
  LAUNCHED(ctx);
  if (flags) CK(cudaMemcpyAsync(flags, sb + o_f, n, cudaMemcpyDeviceToHost, st));
  if (last_pos_xy) CK(cudaMemcpyAsync(last_pos_xy, sb + o_p, 8 * n, cudaMemcpyDeviceToHost, st));
  if (last_frame) CK(cudaMemcpyAsync(last_frame, sb + o_l, 8 * n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return PSIM_OK;
}

int32_t psim_enforce_metal_z_boundaries(psim_ctx* ctx, float max_z, float hw, float hh) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (!isfinite(max_z) || max_z <= 0.0f) return PSIM_OK;            // out_of_plane.rs:142-145
  if (ctx->n == 0 || !(ctx->species_present & 0x6u)) return PSIM_OK;  // no metals: :147-154
  const float metal_max_r = fmaxf(ctx->table_h[1].radius, ctx->table_h[2].radius);
  int32_t rc = cell_build_async(ctx, hw, hh, 4.0f * metal_max_r);  // :160-163
  if (rc) return rc;
  BodyArrays& b = ctx->b[ctx->cur];
  metal_z_kernel<<<(int)((ctx->n + 127) / 128), 128, 0, ctx->stream>>>(
      b.pqr, b.velz, b.species, ctx->n, ctx->cell_start, ctx->cell_end, ctx->order, ctx->grid, metal_max_r, max_z);
  LAUNCHED(ctx);
  CK(cudaGetLastError());
  return PSIM_OK;
}

int32_t psim_shard_init(psim_ctx* ctx, uint32_t rank, uint32_t world) {
  if (!ctx || world < 1 || world > (uint32_t)kMaxRanks || rank >= world)
    return fail(ctx, PSIM_E_ARG, "psim_shard_init: rank / world (at most 64 ranks)");
  DeviceGuard guard(ctx->device);
  auto& S = ctx->sh;
  if (!S.binhist) {
    bool ok = true;
    auto A = [&](auto** p, size_t cnt) {
      if (ok && dalloc(p, cnt) != cudaSuccess) ok = false;
    };
    A(&S.binhist, kBins), A(&S.bins, ctx->cap_bodies), A(&S.binprefix, kBins + 1), A(&S.nb_bin, kBins + 1), A(&S.trav_bin, kBins + 1);
    A(&S.lkeys, ctx->cap_bodies + 2 * kHaloMax + 2), A(&S.xbuf, kBins + kMaxRanks), A(&S.heap, kTopSlots);
    A(&S.plan, 1), A(&S.meta, 1), A(&ctx->keys_idx_all, ctx->cap_bodies);
    if (!ok) {
      cudaGetLastError();
      return fail(ctx, PSIM_E_OOM, "psim_shard_init: device allocation");
    }
  }
  S.on = true, S.rank = rank, S.world = world, S.phase = 0;
  return PSIM_OK;
}
int32_t psim_shard_phase(psim_ctx* ctx, int32_t phase, int32_t mode, float hw, float hh, uint32_t* out) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (mode != PSIM_BUILD_CONTAINING && mode != PSIM_BUILD_DOMAIN) return fail(ctx, PSIM_E_ARG, "psim_shard_phase: mode");
  return shard_phase(ctx, phase, mode, hw, hh, out);
}
int32_t psim_shard_ptrs(psim_ctx* ctx, uint64_t* out8) {
  if (!ctx || !out8) return PSIM_E_ARG;
  if (!ctx->sh.on) return fail(ctx, PSIM_E_STATE, "psim_shard_ptrs: call psim_shard_init first");
  out8[0] = (uint64_t)(uintptr_t)ctx->keys_idx_all, out8[1] = (uint64_t)(uintptr_t)ctx->sh.xbuf;
  out8[2] = (uint64_t)(uintptr_t)ctx->sh.heap, out8[3] = (uint64_t)(uintptr_t)ctx->travA;
  out8[4] = (uint64_t)(uintptr_t)ctx->travB;
  out8[5] = kBins + kMaxRanks, out8[6] = (uint64_t)kTopSlots * (sizeof(TopRec) / 8), out8[7] = ctx->node_cap;
  return PSIM_OK;
}
uint64_t psim_shard_capacity(uint64_t n, uint32_t world) { return world ? (uint64_t)shard_width(n, world) * world : n; }

int32_t psim_comm_unique_id(uint8_t* out128) {
  if (!out128) return PSIM_E_ARG;
  NcclApi& api = nccl_api();
  if (!api.ok) return PSIM_E_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (api.GetUniqueId(&id) != ncclSuccess) return PSIM_E_NCCL;
  memcpy(out128, &id, 128);
  return PSIM_OK;
}

int32_t psim_comm_init(psim_ctx* ctx, const uint8_t* unique_id128, uint32_t rank, uint32_t nranks) {
  if (!ctx || !unique_id128) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  NcclApi& api = nccl_api();
  if (!api.ok) return fail(ctx, PSIM_E_NCCL, "psim_comm_init: libnccl.so.2 not found (set PSIM_NCCL_LIB)");
  auto& K = ctx->comm;
  if (K.on) return fail(ctx, PSIM_E_STATE, "psim_comm_init: the context already has a communicator");
  int32_t rc = psim_shard_init(ctx, rank, nranks);
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, unique_id128, 128);
  NCK(api.CommInitRank(&K.comm, (int)nranks, id, (int)rank));
  CK(cudaStreamCreateWithFlags(&K.side, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&K.ev_main, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&K.ev_side, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&K.ev_vel, cudaEventDisableTiming));
  {
    const char* e = getenv("PSIM_LET");
    K.let = !(e && e[0] == '0') && nranks > 1;
    e = getenv("PSIM_LET_POISON");
    K.let_poison = e && e[0] == '1';
    if (K.let) {
      const uint64_t nb = ctx->cap_bodies ? ctx->cap_bodies : 1;
      K.cap_per_rank = (uint32_t)std::max<uint64_t>(65536, nb / nranks / 2);
      e = getenv("PSIM_LET_CAP");  // test aid: a small send area forces the fall-back to the full all-gather
      if (e && atoi(e) > 0) K.cap_per_rank = (uint32_t)atoi(e);
      bool ok = dalloc(&K.regions, 1) == cudaSuccess && dalloc(&K.cnt, kMaxRanks) == cudaSuccess &&
                dalloc(&K.cnt_all, (size_t)kMaxRanks * kMaxRanks) == cudaSuccess &&
                dalloc(&K.send, (size_t)nranks * K.cap_per_rank) == cudaSuccess &&
                dalloc(&K.recv, (size_t)nranks * K.cap_per_rank) == cudaSuccess;
      if (!ok) {
        cudaGetLastError();
        return fail(ctx, PSIM_E_OOM, "psim_comm_init: locally-essential-tree buffers");
      }
    }
  }
  K.on = true;
  return PSIM_OK;
}

int32_t psim_comm_destroy(psim_ctx* ctx) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  auto& K = ctx->comm;
  if (!K.on) return PSIM_OK;
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(K.side);
  nccl_api().CommDestroy(K.comm);
  cudaStreamDestroy(K.side);
  cudaEventDestroy(K.ev_main), cudaEventDestroy(K.ev_side), cudaEventDestroy(K.ev_vel);
  if (K.stage) cudaFree(K.stage);
  if (K.regions) cudaFree(K.regions);
  if (K.cnt) cudaFree(K.cnt);
  if (K.cnt_all) cudaFree(K.cnt_all);
  if (K.send) cudaFree(K.send);
  if (K.recv) cudaFree(K.recv);
  K = psim_ctx::Comm();
  return PSIM_OK;
}

int32_t psim_comm_stats(psim_ctx* ctx, uint64_t* out4) {
  if (!ctx || !out4) return PSIM_E_ARG;
  out4[0] = ctx->comm.let_sent, out4[1] = ctx->comm.let_full;
  out4[2] = ctx->comm.on && ctx->comm.let ? 1 : 0, out4[3] = ctx->sh.tree_is_let ? 1 : 0;
  return PSIM_OK;
}

int32_t psim_build_sharded(psim_ctx* ctx, int32_t mode, float hw, float hh) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (!ctx->comm.on) return fail(ctx, PSIM_E_STATE, "psim_build_sharded: call psim_comm_init first");
  if (mode != PSIM_BUILD_CONTAINING && mode != PSIM_BUILD_DOMAIN) return fail(ctx, PSIM_E_ARG, "psim_build_sharded: mode");
  return build_sharded(ctx, mode, hw, hh, nullptr, 0.0f);
}

int32_t psim_step_sharded(psim_ctx* ctx, const psim_step_params* p) {
  if (!ctx || !p) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  if (!ctx->comm.on) return fail(ctx, PSIM_E_STATE, "psim_step_sharded: call psim_comm_init first");
  return step_sharded(ctx, p);
}

int32_t psim_mark_positions_changed(psim_ctx* ctx) {
  if (!ctx) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  ctx->tree_valid = ctx->grid_valid = false;
  return PSIM_OK;
}

int32_t psim_phase_times(psim_ctx* ctx, float* ms8) {
  if (!ctx || !ms8) return PSIM_E_ARG;
  DeviceGuard guard(ctx->device);
  for (int k = 0; k < PSIM_NUM_PHASES; ++k) ms8[k] = 0.0f;
  if (!ctx->ev_recorded) return fail(ctx, PSIM_E_STATE, "psim_phase_times: no psim_step recorded");
  CK(cudaEventSynchronize(ctx->ev[7]));
  for (int k = 0; k < 7; ++k) CK(cudaEventElapsedTime(&ms8[k], ctx->ev[k], ctx->ev[k + 1]));
  CK(cudaEventElapsedTime(&ms8[7], ctx->ev[0], ctx->ev[7]));
  return PSIM_OK;
}

int32_t psim_update_state(psim_ctx* ctx, uint64_t n, const float* pos_xy, const float* vel_xy, const float* charge) {
  if (!ctx || n != ctx->n || (n && !pos_xy)) return fail(ctx, PSIM_E_ARG, "psim_update_state: size mismatch");
  DeviceGuard guard(ctx->device);
  ctx->host_map_valid = false;  // rows are in the device's order from here on
  if (n == 0) return PSIM_OK;
  const size_t o_pos = 0, o_vel = align256(8 * n), o_q = o_vel + align256(8 * n), total = o_q + align256(4 * n);
  int32_t rc = ensure_stage(ctx, total);
  if (rc) return rc;
  char* sb = static_cast<char*>(ctx->stage);
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(sb + o_pos, pos_xy, 8 * n, cudaMemcpyHostToDevice, st));
  if (vel_xy) CK(cudaMemcpyAsync(sb + o_vel, vel_xy, 8 * n, cudaMemcpyHostToDevice, st));
  if (charge) CK(cudaMemcpyAsync(sb + o_q, charge, 4 * n, cudaMemcpyHostToDevice, st));
  BodyArrays& b = ctx->b[ctx->cur];
  update_state_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, st>>>(
      reinterpret_cast<const float2*>(sb + o_pos), vel_xy ? reinterpret_cast<const float2*>(sb + o_vel) : nullptr,
      charge ? reinterpret_cast<const float*>(sb + o_q) : nullptr, (uint32_t)n, b.pqr, b.velz);
  LAUNCHED(ctx);
  ctx->tree_valid = ctx->grid_valid = false;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  return PSIM_OK;
}

}  // extern "C"
