/*
 * gsm_oracle_impl.h — body of the CPU oracle, included once per precision.
 *
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: this restates /SPEC.md (the declared
 * model), not the GS-MARL sources, which are withheld (reference readme.md:1;
 * the files it would follow are core.py / environment.py / scenarios/*.py,
 * GSMARL.egg-info/SOURCES.txt:14,15,21-25).  The one pinned piece is ORC(lsa):
 * scipy's rectangular_lsap (scipy==1.7.3 pinned in requirements.txt:101), checked
 * against this image's scipy 1.18.1 in tests/test_oracle_lsa.py.
 *
 * Expects: REAL (float|double), ORC(name) name-mangling macro, R_EXP, R_LOG1P, R_SQRT.
 * Plain sequential loops, one rounding per operation (build with -ffp-contract=off).
 */

/* SPEC §3: numpy logaddexp(0, x). */
static REAL ORC(softplus)(REAL x) {
  if (x > (REAL)0) return x + R_LOG1P(R_EXP(-x));
  return R_LOG1P(R_EXP(x));
}

/*
 * SPEC §5.  Shortest augmenting path LSAP, square n x n, as published in
 * D. F. Crouse, "On implementing 2D rectangular assignment algorithms", IEEE TAES
 * 52(4), 2016, in the form scipy ships (scan over `remaining` filled in reverse,
 * swap-with-last removal, ties prefer an unassigned column).
 * Returns 0, or -1 if infeasible (cannot happen for finite costs).
 */
int ORC(lsa)(const REAL* cost, int n, int32_t* col4row_out) {
  REAL u[ORC_MAX_LSA_N], v[ORC_MAX_LSA_N], spc[ORC_MAX_LSA_N];
  int path[ORC_MAX_LSA_N], col4row[ORC_MAX_LSA_N], row4col[ORC_MAX_LSA_N];
  int remaining[ORC_MAX_LSA_N];
  unsigned char SR[ORC_MAX_LSA_N], SC[ORC_MAX_LSA_N];
  if (n < 1 || n > ORC_MAX_LSA_N) return -1;
  for (int a = 0; a < n; a++) { u[a] = 0; v[a] = 0; path[a] = -1; col4row[a] = -1; row4col[a] = -1; }
  for (int cur = 0; cur < n; cur++) {
    REAL minval = 0;
    int i = cur, nrem = n, sink = -1;
    for (int it = 0; it < n; it++) { remaining[it] = n - it - 1; SR[it] = 0; SC[it] = 0; spc[it] = (REAL)INFINITY; }
    while (sink == -1) {
      int index = -1;
      REAL lowest = (REAL)INFINITY;
      SR[i] = 1;
      for (int it = 0; it < nrem; it++) {
        int j = remaining[it];
        REAL r = minval + cost[i * n + j] - u[i] - v[j];
        if (r < spc[j]) { path[j] = i; spc[j] = r; }
        if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
      }
      minval = lowest;
      if (minval == (REAL)INFINITY) return -1;
      int j = remaining[index];
      if (row4col[j] == -1) sink = j; else i = row4col[j];
      SC[j] = 1;
      remaining[index] = remaining[--nrem];
    }
    u[cur] += minval;
    for (int a = 0; a < n; a++) if (SR[a] && a != cur) u[a] += minval - spc[col4row[a]];
    for (int b = 0; b < n; b++) if (SC[b]) v[b] -= minval - spc[b];
    int j = sink;
    for (;;) {
      int a = path[j];
      row4col[j] = a;
      int t = col4row[a]; col4row[a] = j; j = t;
      if (a == cur) break;
    }
  }
  for (int a = 0; a < n; a++) col4row_out[a] = col4row[a];
  return 0;
}

/* SPEC §5: slots + cost matrix + solve for one env; targets[i] = slot of agent i. */
static void ORC(targets)(const orc_config* c, const REAL* ag, const REAL* lm, int32_t* assign,
                         REAL* tx, REAL* ty) {
  const int N = c->n_agents;
  if (c->scenario == ORC_SCN_NAVIGATION) {
    for (int i = 0; i < N; i++) { assign[i] = i; tx[i] = lm[2 * i]; ty[i] = lm[2 * i + 1]; }
    return;
  }
  REAL sx[ORC_MAX_LSA_N], sy[ORC_MAX_LSA_N];
  REAL cost[ORC_MAX_LSA_N * ORC_MAX_LSA_N];
  for (int k = 0; k < N; k++) {
    if (c->scenario == ORC_SCN_POLYGON) {
      REAL R = (REAL)c->polygon_radius;
      sx[k] = lm[0] + R * (REAL)c->slot_table[2 * k];
      sy[k] = lm[1] + R * (REAL)c->slot_table[2 * k + 1];
    } else {
      REAL f = (REAL)c->slot_table[2 * k];
      sx[k] = lm[0] + f * (lm[2] - lm[0]);
      sy[k] = lm[1] + f * (lm[3] - lm[1]);
    }
  }
  for (int i = 0; i < N; i++)
    for (int k = 0; k < N; k++) {
      REAL dx = sx[k] - ag[4 * i], dy = sy[k] - ag[4 * i + 1];
      cost[i * N + k] = R_SQRT(dx * dx + dy * dy);
    }
  /* non-finite costs (coincident entities -> NaN forces) are outside SPEC.md: identity, no crash */
  if (ORC(lsa)(cost, N, assign) != 0)
    for (int i = 0; i < N; i++) assign[i] = i;
  for (int i = 0; i < N; i++) { tx[i] = sx[assign[i]]; ty[i] = sy[assign[i]]; }
}

/* SPEC §5-7 on the current state of one env.  with_rcd: also reward/cost/done. */
static void ORC(observe_env)(const orc_config* c, const REAL* ag, const REAL* lm, int32_t t,
                             int64_t env, const orc_step_io* io, int with_rcd) {
  const int N = c->n_agents, L = c->n_landmarks, E = N + L, K = c->max_nbrs;
  const int W = (E + 31) / 32;
  REAL* tx = (REAL*)malloc(sizeof(REAL) * 2 * (size_t)N);
  REAL* ty = tx + N;
  REAL* rew = (REAL*)malloc(sizeof(REAL) * (size_t)N);
  int32_t* asg = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
  const REAL Rs = (REAL)c->sensing_radius;
  ORC(targets)(c, ag, lm, asg, tx, ty);
  for (int i = 0; i < N; i++) {
    const REAL px = ag[4 * i], py = ag[4 * i + 1], vx = ag[4 * i + 2], vy = ag[4 * i + 3];
    const int64_t row = env * N + i;
    if (io->obs) {
      REAL* o = (REAL*)io->obs + row * ORC_OBS_DIM;
      o[0] = vx; o[1] = vy; o[2] = px; o[3] = py; o[4] = tx[i] - px; o[5] = ty[i] - py;
    }
    if (io->assign) io->assign[row] = asg[i];
    int cnt = 0, ncol = 0;
    uint32_t words[64];
    for (int w = 0; w < W; w++) words[w] = 0;
    for (int e = 0; e < E; e++) {
      if (e == i) continue;
      REAL ex, ey, evx, evy;
      if (e < N) { ex = ag[4 * e]; ey = ag[4 * e + 1]; evx = ag[4 * e + 2]; evy = ag[4 * e + 3]; }
      else { ex = lm[2 * (e - N)]; ey = lm[2 * (e - N) + 1]; evx = 0; evy = 0; }
      const REAL dx = ex - px, dy = ey - py;
      const REAL dist = R_SQRT(dx * dx + dy * dy);
      int nb = dist < Rs;
      if (c->own_goal_always && c->scenario == ORC_SCN_NAVIGATION && e == N + i) nb = 1;
      if (nb) {
        words[e >> 5] |= 1u << (e & 31);
        if (cnt < K) {
          if (io->nbr_idx) io->nbr_idx[row * K + cnt] = e;
          if (io->nbr_feat) {
            REAL* f = (REAL*)io->nbr_feat + (row * K + cnt) * ORC_NBR_FEAT_DIM;
            f[0] = dx; f[1] = dy; f[2] = evx - vx; f[3] = evy - vy; f[4] = dist; f[5] = (REAL)c->type[e];
          }
          cnt++;
        }
      }
      const REAL dmin = (REAL)c->size[i] + (REAL)c->size[e];
      if (dist < dmin) {
        if (e < N) ncol++;
        else if (c->cost_obstacles && c->type[e] == ORC_ENT_OBSTACLE) ncol++;
      }
    }
    for (int k = cnt; k < K; k++) {
      if (io->nbr_idx) io->nbr_idx[row * K + k] = -1;
      if (io->nbr_feat) {
        REAL* f = (REAL*)io->nbr_feat + (row * K + k) * ORC_NBR_FEAT_DIM;
        for (int q = 0; q < ORC_NBR_FEAT_DIM; q++) f[q] = 0;
      }
    }
    if (io->nbr_cnt) io->nbr_cnt[row] = cnt;
    if (io->adj) for (int w = 0; w < W; w++) io->adj[row * W + w] = words[w];
    if (with_rcd) {
      const REAL gx = tx[i] - px, gy = ty[i] - py;
      const REAL d = R_SQRT(gx * gx + gy * gy);
      rew[i] = ((REAL)0 - (REAL)c->w_dist * d) + (d < (REAL)c->goal_tol ? (REAL)c->w_goal : (REAL)0);
      if (io->cost) ((REAL*)io->cost)[row] = (REAL)ncol;
      if (io->done) io->done[row] = (uint8_t)(t >= c->episode_length);
    }
  }
  if (with_rcd && io->reward) {
    REAL* r = (REAL*)io->reward + env * N;
    if (c->share_reward) {
      REAL s = rew[0];
      for (int i = 1; i < N; i++) s = s + rew[i];
      s = s / (REAL)N;
      for (int i = 0; i < N; i++) r[i] = s;
    } else {
      for (int i = 0; i < N; i++) r[i] = rew[i];
    }
  }
  free(asg);
  free(rew);
  free(tx);
}

/* SPEC §2-4 for one env, in place. */
static void ORC(physics_env)(const orc_config* c, REAL* ag, const REAL* lm, const void* actions,
                             int64_t env) {
  const int N = c->n_agents, L = c->n_landmarks, E = N + L;
  REAL* nv = (REAL*)malloc(sizeof(REAL) * 4 * (size_t)N);
  const REAL cf = (REAL)c->contact_force, km = (REAL)c->contact_margin;
  const REAL dt = (REAL)c->dt, damp = (REAL)c->damping;
  for (int i = 0; i < N; i++) {
    REAL ux = 0, uy = 0;
    if (c->action_mode == ORC_ACT_DISCRETE) {
      int a = ((const int32_t*)actions)[env * N + i];
      if (a >= 0 && a < c->n_discrete_actions) { ux = (REAL)c->discrete_u[2 * a]; uy = (REAL)c->discrete_u[2 * a + 1]; }
    } else {
      ux = ((const REAL*)actions)[(env * N + i) * 2];
      uy = ((const REAL*)actions)[(env * N + i) * 2 + 1];
    }
    const REAL px = ag[4 * i], py = ag[4 * i + 1];
    REAL fx = (REAL)c->accel[i] * ux, fy = (REAL)c->accel[i] * uy;
    if (c->collide[i]) {
      for (int j = 0; j < E; j++) {
        if (j == i || !c->collide[j]) continue;
        REAL qx, qy;
        if (j < N) { qx = ag[4 * j]; qy = ag[4 * j + 1]; } else { qx = lm[2 * (j - N)]; qy = lm[2 * (j - N) + 1]; }
        const REAL dx = px - qx, dy = py - qy;
        const REAL dist = R_SQRT(dx * dx + dy * dy);
        const REAL dmin = (REAL)c->size[i] + (REAL)c->size[j];
        const REAL x = -(dist - dmin) / km;
        const REAL pen = ORC(softplus)(x) * km;
        fx = fx + cf * dx / dist * pen;
        fy = fy + cf * dy / dist * pen;
      }
    }
    REAL vx = ag[4 * i + 2] * ((REAL)1 - damp), vy = ag[4 * i + 3] * ((REAL)1 - damp);
    const REAL m = (REAL)c->mass[i];
    vx = vx + (fx / m) * dt;
    vy = vy + (fy / m) * dt;
    const REAL ms = (REAL)c->max_speed[i];
    if (ms > (REAL)0) {
      const REAL sp = R_SQRT(vx * vx + vy * vy);
      if (sp > ms) { vx = vx / sp * ms; vy = vy / sp * ms; }
    }
    nv[4 * i] = px + vx * dt; nv[4 * i + 1] = py + vy * dt; nv[4 * i + 2] = vx; nv[4 * i + 3] = vy;
  }
  memcpy(ag, nv, sizeof(REAL) * 4 * (size_t)N);
  free(nv);
}

typedef struct ORC(job) {
  const orc_config* c; REAL* ag; const REAL* lm; int32_t* t; const orc_step_io* io; int physics;
} ORC(job);

static void ORC(range)(int64_t lo, int64_t hi, void* ctx) {
  const ORC(job)* jb = (const ORC(job)*)ctx;
  const int N = jb->c->n_agents, L = jb->c->n_landmarks;
  for (int64_t env = lo; env < hi; env++) {
    REAL* ag = jb->ag + env * N * 4;
    const REAL* lm = jb->lm + env * L * 2;
    if (jb->physics == 1) {
      ORC(physics_env)(jb->c, ag, lm, jb->io->actions, env);
      jb->t[env] += 1;
    }
    ORC(observe_env)(jb->c, ag, lm, jb->t[env], env, jb->io, jb->physics != 0);
  }
}

/* SPEC §5-7 (observation, graph, assignment, reward, cost, done) of a GIVEN state, no physics:
 * what tools/unblock.py uses to check the scenario callbacks on the reference's own post-step
 * states, independently of whether the physics sections agree. */
int ORC(evaluate)(const orc_config* c, int64_t n_envs, const REAL* agent_state, const REAL* lm_pos,
                  const int32_t* step_count, const orc_step_io* io) {
  ORC(job) jb = {c, (REAL*)agent_state, lm_pos, (int32_t*)step_count, io, 2};
  orc_parallel_for(n_envs, ORC(range), &jb);
  return 0;
}

int ORC(step)(const orc_config* c, int64_t n_envs, REAL* agent_state, const REAL* lm_pos,
              int32_t* step_count, const orc_step_io* io) {
  ORC(job) jb = {c, agent_state, lm_pos, step_count, io, 1};
  orc_parallel_for(n_envs, ORC(range), &jb);
  return 0;
}

int ORC(observe)(const orc_config* c, int64_t n_envs, const REAL* agent_state, const REAL* lm_pos,
                 const int32_t* step_count, const orc_step_io* io) {
  ORC(job) jb = {c, (REAL*)agent_state, lm_pos, (int32_t*)step_count, io, 0};
  orc_parallel_for(n_envs, ORC(range), &jb);
  return 0;
}

/* SPEC §8. */
int ORC(reset)(const orc_config* c, int64_t n_envs, int64_t env_offset, uint64_t seed,
               const uint8_t* mask, int64_t mask_stride, REAL* agent_state, REAL* lm_pos,
               int32_t* step_count, int32_t* episode) {
  const int N = c->n_agents, L = c->n_landmarks, E = N + L;
  for (int64_t env = 0; env < n_envs; env++) {
    if (mask && !mask[env * mask_stride]) continue;
    const uint64_t g = (uint64_t)(env_offset + env);
    for (int e = 0; e < E; e++) {
      uint32_t r[4];
      orc_philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)episode[env], (uint32_t)e,
                        (uint32_t)seed, (uint32_t)(seed >> 32), r);
      const double ux = orc_u53(r[0], r[1]), uy = orc_u53(r[2], r[3]);
      const double ext = c->spawn_extent[c->type[e]];
      const double x = -ext + (2.0 * ext) * ux, y = -ext + (2.0 * ext) * uy;
      if (e < N) {
        REAL* a = agent_state + (env * N + e) * 4;
        a[0] = (REAL)x; a[1] = (REAL)y; a[2] = 0; a[3] = 0;
      } else {
        REAL* l = lm_pos + (env * L + (e - N)) * 2;
        l[0] = (REAL)x; l[1] = (REAL)y;
      }
    }
    step_count[env] = 0;
    episode[env] += 1;
  }
  return 0;
}
