/* Plain C99 client of include/psim_b200.h: create / upload / build / field / acc_points / step / download /
 * destroy.  Compiled with `gcc -std=c99 -pedantic -Wall -Werror` to prove that the header is C (no C++ types,
 * no torch) and linked with -lpsim_b200; tests/test_abi.py builds it on the CPU box and runs it on the GPU box.
 * Output: one line per body "i x y ex ey" after Quadtree::build + Quadtree::field (quadtree.rs:153-170,418-427). */
#include <stdio.h>
#include <stdlib.h>

#include "psim_b200.h"

#define CHECK(call)                                                                   \
  do {                                                                                \
    int32_t rc_ = (call);                                                             \
    if (rc_ != PSIM_OK) {                                                             \
      fprintf(stderr, "%s -> %d: %s\n", #call, (int)rc_, ctx ? psim_last_error(ctx) : "no context"); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char **argv) {
  const uint64_t n = argc > 1 ? (uint64_t)strtoull(argv[1], NULL, 10) : 64;
  psim_ctx *ctx = NULL;
  psim_config cfg;
  psim_step_params p;
  psim_stats st;
  float *pos = (float *)malloc(sizeof(float) * 2 * n), *q = (float *)malloc(sizeof(float) * n);
  float *radius = (float *)malloc(sizeof(float) * n), *mass = (float *)malloc(sizeof(float) * n);
  float *ef = (float *)malloc(sizeof(float) * 2 * n), *pts = (float *)malloc(sizeof(float) * 4);
  float out2[4];
  uint32_t *perm = (uint32_t *)malloc(sizeof(uint32_t) * n);
  uint64_t i, seed = 0xC0FFEEull;
  for (i = 0; i < n; ++i) { /* splitmix64 -> positions in [-50, 50)^2, alternating charges */
    int k;
    for (k = 0; k < 2; ++k) {
      uint64_t z;
      seed += 0x9E3779B97F4A7C15ull;
      z = seed;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      pos[2 * i + k] = (float)((double)(z >> 11) / 9007199254740992.0 * 100.0 - 50.0);
    }
    q[i] = (i & 1) ? -1.0f : 1.0f;
    radius[i] = (i & 1) ? 2.0f : 0.76f;
    mass[i] = (i & 1) ? 145.0f : 6.94f;
  }
  psim_default_config(&cfg);
  cfg.theta = 0.5f;
  CHECK(psim_create(0, n, 1, &cfg, &ctx));
  CHECK(psim_upload_bodies(ctx, n, pos, NULL, NULL, NULL, mass, radius, q, NULL));
  CHECK(psim_build(ctx, PSIM_BUILD_CONTAINING, 0.0f, 0.0f));
  CHECK(psim_get_permutation(ctx, perm));
  CHECK(psim_field(ctx, 0.138935f, 0.0f, 0.0f, 1, ef, NULL));
  CHECK(psim_download_bodies(ctx, pos, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL));
  CHECK(psim_stats_get(ctx, &st));
  pts[0] = 0.0f, pts[1] = 0.0f, pts[2] = 75.0f, pts[3] = -75.0f;
  CHECK(psim_acc_points(ctx, 2, pts, NULL, NULL, 0.138935f, out2));
  printf("nodes %llu depth %u\n", (unsigned long long)st.reference_nodes, (unsigned)st.max_depth);
  for (i = 0; i < n; ++i) printf("%u %.9g %.9g %.9g %.9g\n", (unsigned)perm[i], pos[2 * i], pos[2 * i + 1], ef[2 * i], ef[2 * i + 1]);
  printf("points %.9g %.9g %.9g %.9g\n", out2[0], out2[1], out2[2], out2[3]);
  /* one fused hot-path step (simulation.rs:1000-1196) */
  p.hw = 50.0f, p.hh = 50.0f, p.hd = 1.0f, p.dt = 5.0f, p.damping_base = 1.0f, p.k_e = 0.138935f;
  p.bg_x = 0.0f, p.bg_y = 0.0f, p.density_threshold = 0.001f, p.enable_out_of_plane = 0;
  p.do_short_range = 1, p.do_electrons = 0, p.do_iterate = 1, p.do_polar = 1, p.reserved[0] = p.reserved[1] = 0;
  CHECK(psim_step(ctx, &p));
  CHECK(psim_sync(ctx));
  CHECK(psim_build_status(ctx));
  CHECK(psim_destroy(ctx));
  ctx = NULL;
  free(pos), free(q), free(radius), free(mass), free(ef), free(pts), free(perm);
  printf("ok\n");
  return 0;
}
