/*
 * gsm_oracle.c — CPU oracle for the batched GS-MARL env hot path.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product library
 * (gs_marl_b200/csrc) never links or calls it.
 *
 * PARITY UNPINNED.  The reference sources this would follow (core.py, environment.py,
 * scenarios/*.py — GSMARL.egg-info/SOURCES.txt:14,15,21-25) are withheld
 * (reference readme.md:1), and the reference ships no tests or golden vectors.  This
 * file restates /SPEC.md, the declared model.  The linear-assignment routine alone is
 * pinned: to scipy.optimize.linear_sum_assignment of this image (scipy 1.18.1), with
 * golden vectors under tests/golden/.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off -pthread -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "orc_types.h"

/* Philox4x32-10 (Salmon et al., SC'11), counter (c0..c3), key (k0,k1). SPEC §8. */
static void orc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                              uint32_t k1, uint32_t out[4]) {
  for (int round = 0; round < 10; round++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double orc_u53(uint32_t hi, uint32_t lo) {
  return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) / 9007199254740992.0;
}

void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                uint32_t* out) {
  orc_philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}

/* Minimal fork-join over contiguous env ranges (the image has no libgomp). */
static int orc_n_threads = 1;
int orc_set_threads(int n) {
  if (n > 0) orc_n_threads = n > 256 ? 256 : n;
  return orc_n_threads;
}
typedef void (*orc_range_fn)(int64_t lo, int64_t hi, void* ctx);
typedef struct { orc_range_fn fn; void* ctx; int64_t lo, hi; } orc_task;
static void* orc_thread_main(void* p) {
  orc_task* t = (orc_task*)p;
  t->fn(t->lo, t->hi, t->ctx);
  return NULL;
}
static void orc_parallel_for(int64_t n, orc_range_fn fn, void* ctx) {
  int nt = orc_n_threads;
  if (nt > n) nt = (int)(n > 0 ? n : 1);
  if (nt <= 1) { fn(0, n, ctx); return; }
  pthread_t th[256];
  orc_task tk[256];
  const int64_t chunk = (n + nt - 1) / nt;
  for (int k = 0; k < nt; k++) {
    tk[k].fn = fn; tk[k].ctx = ctx;
    tk[k].lo = k * chunk; tk[k].hi = (k + 1) * chunk < n ? (k + 1) * chunk : n;
    if (tk[k].lo > n) tk[k].lo = n;
    pthread_create(&th[k], NULL, orc_thread_main, &tk[k]);
  }
  for (int k = 0; k < nt; k++) pthread_join(th[k], NULL);
}

#define REAL double
#define ORC(name) orc_##name##_f64
#define R_EXP exp
#define R_LOG1P log1p
#define R_SQRT sqrt
#include "gsm_oracle_impl.h"
#undef REAL
#undef ORC
#undef R_EXP
#undef R_LOG1P
#undef R_SQRT

#define REAL float
#define ORC(name) orc_##name##_f32
#define R_EXP expf
#define R_LOG1P log1pf
#define R_SQRT sqrtf
#include "gsm_oracle_impl.h"
