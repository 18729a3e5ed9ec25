// gsm_kernels_team.cuh — polygon / line kernel with the per-step linear assignment
// (SPEC.md §2-7; reference scenarios/simple_formation.py, simple_line.py — SOURCES.txt:24-25,
// readme.md:89-90; scipy.optimize.linear_sum_assignment, requirements.txt:101).
//
// The assignment dominates these scenarios (profiles/README.md: 74 % of the instructions).  An env is served
// by a GROUP of G lanes (of a physical group of GP lanes, GP a power of two) and every lane owns A = N / G
// agents, which are also its A assignment columns and rows — so a warp solves 32 / GP problems in lockstep
// (8 for N = 12 with G = 4; 4 for N = 6 with one agent per lane, G = 6 of 8 lanes).  The solver
// (lsa_augment / lsa_cold below) keeps the column duals, path costs, positions and tie keys of a lane's columns
// in registers and everything that is indexed at run time in the env's shared-memory block; group-wide
// minimum / tie-rule maximum are log2(GP) xor-shuffles.  Physics and the neighbour graph use the same layout
// (a lane sweeps the other agents for each of its A agents from a shared-memory position table), rows go to
// HBM over a warp-cooperative coalesced clear like in env_lane_kernel, and `n_steps` steps are fused.
#pragma once
#include "gsm_kernels_spec.cuh"
#include "gsm_kernels_lane.cuh"   // row_store

namespace gsm {

// ---- the group-parallel solver ------------------------------------------------------------------------------
// Round 1's version (per-lane register arrays for everything; removed) spent 165 instructions per Dijkstra
// iteration (profiles/ncu_r2_team_polygon12_lines.txt): select
// chains that index per-lane register arrays with a run-time index (u[il], r4c[jl], path[jl]: 16), the relax
// step behind divergent branches (38), tie keys rebuilt from scratch (22), boolean arrays (22), index splits by
// A (11).  Here everything that is addressed by a run-time index — u, row4col, col4row, path and, after the
// search, the shortest-path costs — lives in the env's shared-memory block and is read with one LDS (group-
// uniform address: a broadcast); the sets SR / SC are N-bit masks that every lane of the group carries
// redundantly (i and j are group-uniform); the relax step is branch-free (a finished group runs with
// minval = +inf, which can update nothing); the tie key of a column (unassigned: 64 + pos, assigned: 31 - pos,
// column in the low bits) is kept in a register and touched only when its position or status changes.
// Scan order and tie rule are scipy's (rectangular_lsap, Crouse 2016): `remaining` filled in reverse with
// swap-with-last removal is tracked as a per-column position, "last unassigned minimum in scan order, else the
// first minimum" is one maximum over packed (key, column) words.
struct TeamLsaSmem { int u, spc, r4c, c4r, path, words; };      // word offsets inside the env's solver block
__host__ __device__ constexpr TeamLsaSmem team_lsa_smem(int rb, int N) {
  TeamLsaSmem t{};
  int o = 0;
  t.u = o; o += N * rb / 4;
  t.spc = o; o += N * rb / 4;
  t.r4c = o; o += N;
  t.c4r = o; o += N;
  t.path = o; o += N;
  t.words = o;
  return t;
}

// Minimum / maximum over the GP lanes of a group: log2(GP) xor-shuffle + compare stages.  The alternative — ONE
// redux.sync over the group's lane mask (order-preserving integer key for the float) — is 2.3x SLOWER end to end
// (polygon-12 137 vs 60 us per step, measured in round 2): a redux.sync over a sub-warp mask is executed once per
// distinct mask, i.e. 8 times per warp here.
template <int GP, typename T>
__device__ __forceinline__ T group_min(T v, unsigned) {
#pragma unroll
  for (int m = GP / 2; m >= 1; m >>= 1) { const T o = __shfl_xor_sync(0xffffffffu, v, m); v = o < v ? o : v; }
  return v;
}
template <int GP>
__device__ __forceinline__ int group_max(int v, unsigned) {
#pragma unroll
  for (int m = GP / 2; m >= 1; m >>= 1) { const int o = __shfl_xor_sync(0xffffffffu, v, m); v = o > v ? o : v; }
  return v;
}

// One shortest-augmenting-path step for row `cur` of every ACTIVE group, in lockstep: Dijkstra search from
// the current duals and matching, dual update, augmentation.  v[] (my column duals) lives in registers, the
// rest of the solver state in the env's shared-memory block.  Inactive groups idle.
template <typename T, int N, int G, int GP>
__device__ __forceinline__ void lsa_augment(const T* __restrict__ C, uint32_t* __restrict__ ws, int g, bool live,
                                            int cur, bool active, T (&v)[N / G]) {
  constexpr int A = N / G;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr TeamLsaSmem L = team_lsa_smem((int)sizeof(T), N);
  const T INF = r_inf<T>();
  T* const su = (T*)(ws + L.u);
  T* const sspc = (T*)(ws + L.spc);
  int* const sr4c = (int*)(ws + L.r4c);
  int* const sc4r = (int*)(ws + L.c4r);
  int* const spath = (int*)(ws + L.path);
  const int col0 = (g < G ? g : 0) * A;                   // my first column / row
  const T* Cg = C + col0;
  const unsigned gmask = low_mask(GP) << ((threadIdx.x & 31) / GP * GP);   // the lanes of my group
  T spc[A];
  int path[A], pos[A], keyp[A];
  bool asg[A];                                            // my column is assigned (row4col != -1)
  T minval = 0;
  int i = active ? cur : 0, nrem = N, sink = active ? -1 : 0;
  unsigned SR = 0, SC = 0;                                // group-uniform
  unsigned inr = live ? low_mask(A) : 0u;                 // my columns still in scipy's `remaining`
#pragma unroll
  for (int a = 0; a < A; a++) {
    asg[a] = live && sr4c[col0 + a] != -1;
    spc[a] = INF; path[a] = -1; pos[a] = N - 1 - (col0 + a);
    keyp[a] = ((asg[a] ? 31 - pos[a] : 64 + pos[a]) << 6) | (col0 + a);
  }
  for (int iter = 0; iter < N && __any_sync(FULL, sink == -1); iter++) {
    const bool run = sink == -1;
    if (run) SR |= 1u << i;
    const T u_i = su[i];
    const T mv = run ? minval : INF;                      // a finished group relaxes nothing
    const T* Ci = Cg + i * N;
    T lo = INF;
#pragma unroll
    for (int a = 0; a < A; a++) {
      const T r = mv + Ci[a] - u_i - v[a];
      const bool in_a = (inr >> a) & 1u;
      const bool upd = in_a && r < spc[a];
      spc[a] = upd ? r : spc[a];
      path[a] = upd ? i : path[a];
      const T cand = in_a ? spc[a] : INF;
      lo = cand < lo ? cand : lo;
    }
    lo = group_min<GP>(lo, gmask);
    // tie rule: an unassigned minimum with the largest position wins, else the minimum with the smallest
    // position; positions are unique, the column index rides in the low bits of the key
    int best = -1;
#pragma unroll
    for (int a = 0; a < A; a++) {
      const int packed = (((inr >> a) & 1u) && spc[a] == lo) ? keyp[a] : -1;
      best = packed > best ? packed : best;
    }
    best = group_max<GP>(best, gmask);
    const int key = best >> 6, j = best < 0 ? 0 : (best & 63);
    const int selpos = key >= 64 ? key - 64 : 31 - key;
    const int r4c_j = sr4c[j];
    if (run && best >= 0) {
      minval = lo;
      if (r4c_j == -1) sink = j; else i = r4c_j;
      nrem--;
      SC |= 1u << j;
#pragma unroll
      for (int a = 0; a < A; a++) {
        if (col0 + a == j) inr &= ~(1u << a);
        if (((inr >> a) & 1u) && pos[a] == nrem) {
          pos[a] = selpos;
          keyp[a] = ((asg[a] ? 31 - selpos : 64 + selpos) << 6) | (col0 + a);
        }
      }
    }
  }
  // dual update (col4row as it was before this augmentation); the shortest-path costs and predecessors
  // of my columns go to shared memory for the indexed reads
#pragma unroll
  for (int a = 0; a < A; a++) if (live) { sspc[col0 + a] = spc[a]; spath[col0 + a] = path[a]; }
  __syncwarp();
  const bool found = active && sink >= 0;
#pragma unroll
  for (int a = 0; a < A; a++) {
    const int row = col0 + a;
    if (found && live && ((SR >> row) & 1u)) {
      const int c = sc4r[row];
      su[row] += row == cur ? minval : minval - sspc[c < 0 ? 0 : c];
    }
    if (found && ((SC >> (col0 + a)) & 1u)) v[a] -= minval - spc[a];
  }
  __syncwarp();
  // augment along the path (every lane of the group walks it redundantly; lane 0 writes)
  int j = sink < 0 ? 0 : sink;
  bool going = found;
  for (int iter = 0; iter < N && __any_sync(FULL, going); iter++) {
    int arow = spath[j];
    arow = arow < 0 ? 0 : arow;
    const int tprev = sc4r[arow];
    __syncwarp();                                         // all lanes of the group have read before lane 0 writes
    if (going) {
      if (g == 0) { sr4c[j] = arow; sc4r[arow] = j; }
      j = tprev < 0 ? 0 : tprev;
      if (arow == cur) going = false;
    }
    __syncwarp();
  }
}

// Cold solve = scipy's rectangular_lsap: zero duals, empty matching, rows 0..N-1 in order.
template <typename T, int N, int G, int GP>
__device__ __forceinline__ void lsa_cold(const T* __restrict__ C, uint32_t* __restrict__ ws, int g, bool live,
                                         bool active, T (&v)[N / G]) {
  constexpr int A = N / G;
  constexpr TeamLsaSmem L = team_lsa_smem((int)sizeof(T), N);
  const int col0 = (g < G ? g : 0) * A;
  if (active) {
#pragma unroll
    for (int a = 0; a < A; a++) {
      v[a] = 0;
      if (live) { ((T*)(ws + L.u))[col0 + a] = 0; ((int*)(ws + L.r4c))[col0 + a] = -1; ((int*)(ws + L.c4r))[col0 + a] = -1; }
    }
  }
  __syncwarp();
  for (int cur = 0; cur < N; cur++) lsa_augment<T, N, G, GP>(C, ws, g, live, cur, active, v);
}

template <typename T, int N, int G, int GP>
__device__ __forceinline__ void lsa_group2(const T* __restrict__ C, uint32_t* __restrict__ ws, int g, int base,
                                           bool grp_live, int (&c4r_out)[N / G]) {
  constexpr int A = N / G;
  constexpr TeamLsaSmem L = team_lsa_smem((int)sizeof(T), N);
  const bool live = grp_live && g < G;
  T v[A];
  lsa_cold<T, N, G, GP>(C, ws, g, live, grp_live, v);
  const int col0 = (g < G ? g : 0) * A;
#pragma unroll
  for (int a = 0; a < A; a++) c4r_out[a] = live ? ((const int*)(ws + L.c4r))[col0 + a] : -1;
}

// ---- N = 6, fp32: exhaustive solve with a uniqueness certificate ---------------------------------------------
// 6! = 720 assignments; working lane g of the group takes the 120 with row 0 -> column g: its 5 x 5 sub-matrix
// (rows 1..5, the columns other than g in ascending order) sits in registers and the enumeration is straight-line
// code with compile-time indices — rows 1..3 as a prefix tree (5 + 20 + 60 partial sums), rows 4, 5 as the cheaper
// of the two completions of the remaining column pair; best total, second best total and the best prefix number
// are tracked with min / max / select, no dependent chain longer than a sum, no shared-memory state.
// The result is taken only if the best total beats EVERY other one (the other completion of the winning prefix
// included) by GSM_TEAM_ENUM6_TOL = 2e-5 of the largest matrix entry: the enumeration's own rounding is below 2e-6
// of it (five additions), scipy's shortest-augmenting-path arithmetic in fp32 moves its path costs by ~1e-6 of it
// (worst-case bound of the same order as the margin, SPEC.md §9 deviation 7), so scipy's answer is this one;
// otherwise — ties, near-ties — the warp runs scipy's own procedure (lsa_group2).  Measured (profiles/README.md):
// the certificate fails for 0.2 % (polygon) / 1.4 % (line) of the env-steps; a fallback is expensive (a lone
// latency-bound solve), which is why the margin is not wider.  fp64 (verification) always runs lsa_group2.
#ifndef GSM_TEAM_ENUM6
#define GSM_TEAM_ENUM6 1       // 0: off.  Diagnostic builds: 2 no fallback (timing only), 3 = 2 + failure flag in bit 1 of `done`
#endif
#ifndef GSM_TEAM_ENUM6_TOL
#define GSM_TEAM_ENUM6_TOL 2e-5f
#endif
template <int GP>
__device__ __forceinline__ bool lsa_enum6(const float* __restrict__ C, int g, int base, int& col_out) {
  constexpr unsigned FULL = 0xffffffffu;
  const bool has = g < 6;
  const int gg = has ? g : 0;
  float M[5][5];
  const float c0 = C[gg];
  float cmax = c0;
#pragma unroll
  for (int r = 0; r < 5; r++)
#pragma unroll
    for (int j = 0; j < 5; j++) {
      M[r][j] = C[(r + 1) * 6 + j + (j >= gg ? 1 : 0)];
      cmax = fmaxf(cmax, M[r][j]);
    }
  float Smin[5][5];                                        // completions of rows 4, 5 for every column pair p < q
#pragma unroll
  for (int pp = 0; pp < 5; pp++)
#pragma unroll
    for (int q = pp + 1; q < 5; q++) Smin[pp][q] = fminf(M[3][pp] + M[4][q], M[3][q] + M[4][pp]);
  const float inf = __int_as_float(0x7f800000);
  float best = inf, second = inf;
  int id = 0;
#pragma unroll
  for (int j0 = 0; j0 < 5; j0++) {
    const float a = c0 + M[0][j0];
#pragma unroll
    for (int j1 = 0; j1 < 5; j1++) {
      if (j1 == j0) continue;
      const float b = a + M[1][j1];
#pragma unroll
      for (int j2 = 0; j2 < 5; j2++) {
        if (j2 == j0 || j2 == j1) continue;
        int pp = -1, q = -1;
#pragma unroll
        for (int c = 0; c < 5; c++) if (c != j0 && c != j1 && c != j2) { if (pp < 0) pp = c; else q = c; }
        const int r1 = j1 - (j1 > j0 ? 1 : 0);                              // rank of j1 among the 4 left
        const int r2 = j2 - (j2 > j0 ? 1 : 0) - (j2 > j1 ? 1 : 0);          // rank of j2 among the 3 left
        const float t = (b + M[2][j2]) + Smin[pp][q];
        id = t < best ? j0 * 12 + r1 * 3 + r2 : id;
        second = fminf(second, fmaxf(best, t));
        best = fminf(best, t);
      }
    }
  }
  if (!has) { best = inf; second = inf; }
  const float mine = best;
#pragma unroll
  for (int m = GP / 2; m >= 1; m >>= 1) {                  // best / second best / largest entry over the group
    const float ob = __shfl_xor_sync(FULL, best, m), os = __shfl_xor_sync(FULL, second, m);
    second = fminf(fminf(second, os), fmaxf(best, ob));
    best = fminf(best, ob);
    cmax = fmaxf(cmax, __shfl_xor_sync(FULL, cmax, m));
  }
  const unsigned winners = (__ballot_sync(FULL, has && mine == best) >> base) & ((1u << GP) - 1u);
  const int gw = winners ? __ffs(winners) - 1 : 0;                            // row 0 -> column gw
  const float margin = GSM_TEAM_ENUM6_TOL * cmax;
  bool ok = winners != 0 && second - best > margin;
  // the winner decodes its prefix, settles the completion on the two sums and checks the other completion
  unsigned packed = 0;
  if (g == gw) {
    const int j0 = id / 12, r1 = (id - j0 * 12) / 3, r2 = id - j0 * 12 - r1 * 3;
    const int j1 = r1 + (r1 >= j0 ? 1 : 0);
    int j2 = r2;                                                              // r2-th of the columns other than j0, j1
    { const int lo = j0 < j1 ? j0 : j1, hi = j0 < j1 ? j1 : j0; if (j2 >= lo) j2++; if (j2 >= hi) j2++; }
    unsigned left = 31u & ~(1u << j0) & ~(1u << j1) & ~(1u << j2);
    const int pp = __ffs(left) - 1; left &= left - 1;
    const int q = __ffs(left) - 1;
    const int cj0 = j0 + (j0 >= g ? 1 : 0), cj1 = j1 + (j1 >= g ? 1 : 0), cj2 = j2 + (j2 >= g ? 1 : 0);
    const int cp = pp + (pp >= g ? 1 : 0), cq = q + (q >= g ? 1 : 0);
    const float s1 = C[4 * 6 + cp] + C[5 * 6 + cq], s2 = C[4 * 6 + cq] + C[5 * 6 + cp];
    const bool sw = s2 < s1;
    if (!(fabsf(s1 - s2) > margin)) ok = false;
    packed = (unsigned)g | ((unsigned)cj0 << 3) | ((unsigned)cj1 << 6) | ((unsigned)cj2 << 9) |
             ((unsigned)(sw ? cq : cp) << 12) | ((unsigned)(sw ? cp : cq) << 15) | (ok ? 0u : 0x80000000u);
  }
  packed = __shfl_sync(FULL, packed, base + gw);
  col_out = has ? (int)((packed >> (3 * g)) & 7u) : -1;
  return ok && !(packed >> 31);
}

// A warm-started variant (Bellman-Ford re-centred duals, free rows, uniqueness certificate, cold fallback) was built
// on lsa_augment, is parity-green and not faster: profiles/rejected/team_warm_start_lsa_r2.diff, profiles/README.md.

constexpr int kTeamThreads = 128;

__host__ __device__ inline size_t team_env_bytes(int rb, int N) {
  // per env: agent positions + velocities, N x N costs, shared-reward scratch, solver block (lsa_group2)
  return ((size_t)(4 * N + N * N + N) * rb + (size_t)team_lsa_smem(rb, N).words * 4 + 15) / 16 * 16;
}

#ifndef GSM_TEAM_ROW16         // A/B: fp32 rows as a 16-byte + an 8-byte store (lane kernel, profiles/README.md)
#define GSM_TEAM_ROW16 1
#endif
#ifndef GSM_TEAM_BLOCKS        // A/B: resident CTAs per SM the fp32 instances are compiled for (4 -> 128 registers)
#define GSM_TEAM_BLOCKS 4
#endif
#ifndef GSM_TEAM_BLOCKS_A1     // ... of the one-agent-per-lane instances (N = 6, polygon, one launch / 4 streams: 4 CTAs 22.4 /
#define GSM_TEAM_BLOCKS_A1 5   //     16.1 us, 5 CTAs 17.9 / 16.4, 6 CTAs 19.6 / 17.9, 7 CTAs 23.0 / 21.2; with lsa_enum6: 4 CTAs 13.4 / 10.8,
                               //     5 CTAs (96 registers) 12.9 / 10.2, 6 CTAs (78) 15.2 / 11.8)
#endif
template <typename T, int A> struct TeamMinBlocks {
  static constexpr int value = sizeof(T) != 4 ? 1 : (A == 1 ? GSM_TEAM_BLOCKS_A1 : GSM_TEAM_BLOCKS);
};
template <typename T, int SCN, int N, int G, int GP>
__global__ void __launch_bounds__(kTeamThreads, TeamMinBlocks<T, N / G>::value)
env_team_kernel(const __grid_constant__ KParams<T> p, const int n_steps,
                const __grid_constant__ StepStrides ss) {
  constexpr int A = N / G, EPW = 32 / GP, L = SCN == GSM_SCN_POLYGON ? 1 : 2, E = N + L;
  static_assert(N % G == 0 && (GP & (GP - 1)) == 0 && G <= GP && GP <= 32, "G of GP lanes per env, A agents per lane");
  static_assert(E <= 32, "adjacency is one 32-bit word");
  typedef Arith<T> AR;
  extern __shared__ __align__(128) unsigned char sm[];
  const int K = p.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane % GP, grp = lane / GP, base = grp * GP;
  const bool has = g < G;                                 // this lane owns A agents / columns
  const int64_t env_w0 = ((int64_t)blockIdx.x * (kTeamThreads / 32) + warp) * EPW;   // first env of the warp
  const int64_t env_raw = env_w0 + grp;
  const bool env_ok = env_raw < p.n_envs;
  const bool live = env_ok && has;                        // lanes that store
  const int64_t env = env_ok ? env_raw : 0;               // shadow groups compute, never store
  const int ga = has ? g : 0;                             // idle lanes shadow lane 0's agents (no stores, no table writes)

  const size_t eb = team_env_bytes((int)sizeof(T), N);
  T* e_sm = (T*)(sm + (size_t)(warp * EPW + grp) * eb);
  T* apos = e_sm;                                         // [N][2]
  T* avel = e_sm + 2 * N;                                 // [N][2]
  T* cmat = e_sm + 4 * N;                                 // [N][N]
  T* rsm = e_sm + 4 * N + N * N;                          // [N]

  // ---- my A agents ----------------------------------------------------------------------------
  T px[A], py[A], vx[A], vy[A], size_a[A], mass_a[A], accel_a[A], maxsp_a[A];
  bool coll_a[A];
#pragma unroll
  for (int a = 0; a < A; a++) {
    const int i = ga * A + a;
    const T* s = p.agent_state + (env * N + i) * 4;
    px[a] = s[0]; py[a] = s[1]; vx[a] = s[2]; vy[a] = s[3];
    size_a[a] = p.size[i]; coll_a[a] = p.eflag[i] & 1;
    mass_a[a] = p.mass[i]; accel_a[a] = p.accel[i]; maxsp_a[a] = p.max_speed[i];
    if (has) { apos[2 * i] = px[a]; apos[2 * i + 1] = py[a]; avel[2 * i] = vx[a]; avel[2 * i + 1] = vy[a]; }
  }
  T mx[L], my[L], msize[L];
  int mflag[L];
#pragma unroll
  for (int l = 0; l < L; l++) {
    const T* m = p.lm_pos + (env * L + l) * 2;
    mx[l] = m[0]; my[l] = m[1]; msize[l] = p.size[N + l]; mflag[l] = p.eflag[N + l];
  }
  int t_now = p.t[env];
  const bool auto_reset = p.auto_reset != 0;
  int ep = auto_reset ? p.episode[env] : 0;
  const int ep0 = ep;
  const uint64_t genv = (uint64_t)(p.env_offset + env);
  const T cut = (T)kFarCut * p.km;
  __syncwarp();

  const unsigned char* c_act = (const unsigned char*)p.actions;
  unsigned char* c_obs = (unsigned char*)p.obs;
  unsigned char* c_idx = (unsigned char*)p.nbr_idx;
  unsigned char* c_feat = (unsigned char*)p.nbr_feat;
  unsigned char* c_cnt = (unsigned char*)p.nbr_cnt;
  unsigned char* c_adj = (unsigned char*)p.adj;
  unsigned char* c_rew = (unsigned char*)p.reward;
  unsigned char* c_cost = (unsigned char*)p.cost;
  unsigned char* c_done = (unsigned char*)p.done;
  unsigned char* c_asg = (unsigned char*)p.assign;

  for (int step = 0; step < n_steps; step++) {
    // ---- SPEC §2-4 for my A agents (position table = state at the start of the step) ----------
#pragma unroll
    for (int a = 0; a < A; a++) {
      const int i = ga * A + a;
      const int64_t row = env * N + i;
      T ux = 0, uy = 0;
      if (p.action_mode == GSM_ACT_DISCRETE) {
        const int ac = ((const int32_t*)c_act)[row];
        if (ac >= 0 && ac < p.n_actions) { ux = p.discrete_u[ac][0]; uy = p.discrete_u[ac][1]; }
      } else { ux = ((const T*)c_act)[row * 2]; uy = ((const T*)c_act)[row * 2 + 1]; }
      T fx = accel_a[a] * ux, fy = accel_a[a] * uy;
      if (coll_a[a]) {
        for (int j = 0; j < E; j++) {
          if (j == i) continue;
          T qx, qy, qs;
          bool qc;
          if (j < N) { qx = apos[2 * j]; qy = apos[2 * j + 1]; qs = p.size[j]; qc = p.eflag[j] & 1; }
          else { qx = mx[j - N < L ? j - N : 0]; qy = my[j - N < L ? j - N : 0]; qs = msize[j - N < L ? j - N : 0]; qc = mflag[j - N < L ? j - N : 0] & 1; }
          if (!qc) continue;
          const T dx = px[a] - qx, dy = py[a] - qy;
          const T d2 = dx * dx + dy * dy;
          const T dmin = size_a[a] + qs;
          if (Prec<T>::kCut) { const T far = dmin + cut; if (d2 > far * far) continue; }
          const T dist = AR::sqrt(d2);
          const T x = AR::div_const(-(dist - dmin), p.km, p.km_inv);
          if (Prec<T>::kCut && x < (T)(-kFarCut)) continue;
          const T pen = softplus(x) * p.km;
          fx = fx + AR::div(p.cf * dx, dist) * pen;
          fy = fy + AR::div(p.cf * dy, dist) * pen;
        }
      }
      T nvx = vx[a] * p.one_minus_damp, nvy = vy[a] * p.one_minus_damp;
      nvx = nvx + AR::div(fx, mass_a[a]) * p.dt;
      nvy = nvy + AR::div(fy, mass_a[a]) * p.dt;
      if (maxsp_a[a] > (T)0) {
        const T sp = AR::sqrt(nvx * nvx + nvy * nvy);
        if (sp > maxsp_a[a]) { nvx = AR::div(nvx, sp) * maxsp_a[a]; nvy = AR::div(nvy, sp) * maxsp_a[a]; }
      }
      vx[a] = nvx; vy[a] = nvy;
      px[a] = px[a] + nvx * p.dt; py[a] = py[a] + nvy * p.dt;
    }
    t_now += 1;
    __syncwarp();                                          // every lane has read the old table
#pragma unroll
    for (int a = 0; a < A; a++) {
      const int i = ga * A + a;
      if (has) { apos[2 * i] = px[a]; apos[2 * i + 1] = py[a]; avel[2 * i] = vx[a]; avel[2 * i + 1] = vy[a]; }
    }
    __syncwarp();

    // ---- SPEC §5: slots of my A columns, cost columns, group-parallel assignment --------------
    T sx[A], sy[A];
#pragma unroll
    for (int a = 0; a < A; a++) {
      const int k = ga * A + a;
      if (SCN == GSM_SCN_POLYGON) {
        sx[a] = mx[0] + p.poly_r * p.slot_table[2 * k];
        sy[a] = my[0] + p.poly_r * p.slot_table[2 * k + 1];
      } else {
        const T f = p.slot_table[2 * k];
        sx[a] = mx[0] + f * (mx[L - 1] - mx[0]);
        sy[a] = my[0] + f * (my[L - 1] - my[0]);
      }
      for (int r = 0; r < N && has; r++) {
        const T dx = sx[a] - apos[2 * r], dy = sy[a] - apos[2 * r + 1];
        cmat[r * N + k] = r_sqrt(dx * dx + dy * dy);
      }
    }
    __syncwarp();
    int c4r[A];
    bool enum_fail = false;                                // (diagnostic builds: GSM_TEAM_ENUM6 = 3 reports it in bit 1 of `done`)
#ifdef GSM_TEAM_NO_LSA   // diagnostic build only (profiles/README.md): what the step costs without the solve
#pragma unroll
    for (int a = 0; a < A; a++) c4r[a] = ga * A + a;
#else
    // (polygon only: with the collinear slots of `line` 1.2 % of the env-steps fail the certificate, 4.9 % of the warps fall
    //  back, and one launch is slower than without the enumeration — 21.5 vs 19.7 us per step)
    if constexpr (GSM_TEAM_ENUM6 != 0 && SCN == GSM_SCN_POLYGON && N == 6 && G == 6 && sizeof(T) == 4) {
      const bool certified = lsa_enum6<GP>((const float*)cmat, g, base, c4r[0]);
      enum_fail = !certified;
      if (GSM_TEAM_ENUM6 < 2 && __any_sync(0xffffffffu, !certified))   // a tie or near-tie somewhere in the warp: scipy's own procedure (2: timing-only build without it)
        lsa_group2<T, N, G, GP>(cmat, (uint32_t*)(rsm + N), g, base, true, c4r);
    } else {
      lsa_group2<T, N, G, GP>(cmat, (uint32_t*)(rsm + N), g, base, true, c4r);
    }
#endif
    __syncwarp();

    // ---- padding first: the rows of this warp's envs are one contiguous region -------------------
    {
      int64_t ne = p.n_envs - env_w0;
      ne = ne > EPW ? EPW : (ne < 0 ? 0 : ne);
      const int64_t row0 = env_w0 * N, nrows = ne * N;
      unsigned char* zi = c_idx + row0 * K * 4;
      const int64_t bi = nrows * K * 4;
      if ((((uintptr_t)zi | (uintptr_t)bi) & 15) == 0) {
        for (int64_t q = (int64_t)lane * 16; q < bi; q += 512) *reinterpret_cast<int4*>(zi + q) = make_int4(-1, -1, -1, -1);
      } else {
        for (int64_t q = (int64_t)lane * 4; q < bi; q += 128) *reinterpret_cast<int32_t*>(zi + q) = -1;
      }
      unsigned char* zf = c_feat + row0 * K * GSM_NBR_FEAT_DIM * (int64_t)sizeof(T);
      const int64_t bf = nrows * K * GSM_NBR_FEAT_DIM * (int64_t)sizeof(T);
      if ((((uintptr_t)zf | (uintptr_t)bf) & 15) == 0) {
        for (int64_t q = (int64_t)lane * 16; q < bf; q += 512) *reinterpret_cast<int4*>(zf + q) = make_int4(0, 0, 0, 0);
      } else {
        for (int64_t q = (int64_t)lane * sizeof(T); q < bf; q += 32 * sizeof(T)) *reinterpret_cast<T*>(zf + q) = (T)0;
      }
      __syncwarp();
    }

    // ---- SPEC §6-7 for my A agents ---------------------------------------------------------------
    T rew[A];
#pragma unroll
    for (int a = 0; a < A; a++) {
      const int i = ga * A + a;
      const int64_t row = env * N + i;
      int cnt = 0, ncol = 0;
      uint32_t word = 0;
      int32_t* g_idx = (int32_t*)c_idx + row * K;
      T* g_feat = (T*)c_feat + row * K * GSM_NBR_FEAT_DIM;
      const bool feat16 = GSM_TEAM_ROW16 && A == 1 && sizeof(T) == 4 && (((uintptr_t)g_feat) & 15) == 0;   // A = 3 (N = 12): measured slower (polygon-12 68.9 vs 63.0 us)
      for (int e = 0; e < E; e++) {
        if (e == i) continue;
        T ex, ey, evx = 0, evy = 0, es;
        int fl;
        if (e < N) { ex = apos[2 * e]; ey = apos[2 * e + 1]; evx = avel[2 * e]; evy = avel[2 * e + 1]; es = p.size[e]; fl = p.eflag[e]; }
        else { const int l = e - N < L ? e - N : 0; ex = mx[l]; ey = my[l]; es = msize[l]; fl = mflag[l]; }
        const T dx = ex - px[a], dy = ey - py[a];
        const T dist = AR::sqrt(dx * dx + dy * dy);
        if (dist < size_a[a] + es && (e < N || (p.cost_obstacles && (fl >> 1) == GSM_ENT_OBSTACLE))) ncol++;
        if (dist < p.Rs) {
          word |= 1u << e;
          if (cnt < K && live) {
            g_idx[cnt] = e;
            T* f = g_feat + cnt * GSM_NBR_FEAT_DIM;
            row_store<T>(f, feat16, cnt & 1, dx, dy, evx - vx[a], evy - vy[a], dist, (T)(fl >> 1));   // fp32: 16 + 8 bytes
          }
          cnt++;
        }
      }
      if (cnt > K) cnt = K;
      const int k = c4r[a] < 0 ? 0 : c4r[a];
      T tx, ty;
      if (SCN == GSM_SCN_POLYGON) {
        tx = mx[0] + p.poly_r * p.slot_table[2 * k];
        ty = my[0] + p.poly_r * p.slot_table[2 * k + 1];
      } else {
        const T f = p.slot_table[2 * k];
        tx = mx[0] + f * (mx[L - 1] - mx[0]);
        ty = my[0] + f * (my[L - 1] - my[0]);
      }
      const T gx = tx - px[a], gy = ty - py[a];
      const T d = AR::sqrt(gx * gx + gy * gy);
      rew[a] = ((T)0 - p.w_dist * d) + (d < p.goal_tol ? p.w_goal : (T)0);
      if (live) {
        T* o = (T*)c_obs + row * GSM_OBS_DIM;
        st2<T>(o, vx[a], vy[a]); st2<T>(o + 2, px[a], py[a]); st2<T>(o + 4, gx, gy);
        ((int32_t*)c_cnt)[row] = cnt;
        ((uint32_t*)c_adj)[row] = word;
        ((T*)c_cost)[row] = (T)ncol;
        c_done[row] = (uint8_t)(t_now >= p.episode_length) | (uint8_t)(GSM_TEAM_ENUM6 == 3 && enum_fail ? 2 : 0);
        ((int32_t*)c_asg)[row] = c4r[a];
        if (!p.share_reward) ((T*)c_rew)[row] = rew[a];
      }
    }
    if (p.share_reward) {                                  // mean over the env's agents, ascending order
#pragma unroll
      for (int a = 0; a < A; a++) if (has) rsm[ga * A + a] = rew[a];
      __syncwarp();
      T s = rsm[0];
      for (int k = 1; k < N; k++) s = s + rsm[k];
      s = s / (T)N;
      if (live) {
#pragma unroll
        for (int a = 0; a < A; a++) ((T*)c_rew)[env * N + ga * A + a] = s;
      }
      __syncwarp();
    }

    // ---- episode end inside a fused rollout (SPEC §8 draws, cold) ----------------------------------
    if (auto_reset && t_now >= p.episode_length) {
#pragma unroll
      for (int a = 0; a < A; a++) {
        spawn_draw<T>(genv, ep, ga * A + a, p.seed, p.ext[GSM_ENT_AGENT], px[a], py[a]);
        vx[a] = 0; vy[a] = 0;
      }
#pragma unroll
      for (int l = 0; l < L; l++) spawn_draw<T>(genv, ep, N + l, p.seed, p.ext[mflag[l] >> 1], mx[l], my[l]);
      t_now = 0;
      ep += 1;
    }
    if (auto_reset) {                                      // the table must show the re-drawn agents
      __syncwarp();
#pragma unroll
      for (int a = 0; a < A; a++) {
        const int i = ga * A + a;
        if (has) { apos[2 * i] = px[a]; apos[2 * i + 1] = py[a]; avel[2 * i] = vx[a]; avel[2 * i + 1] = vy[a]; }
      }
      __syncwarp();
    }
    c_act += ss.actions; c_obs += ss.obs; c_idx += ss.nbr_idx; c_feat += ss.nbr_feat; c_cnt += ss.nbr_cnt;
    c_adj += ss.adj; c_rew += ss.reward; c_cost += ss.cost; c_done += ss.done; c_asg += ss.assign;
  }

  if (live) {
#pragma unroll
    for (int a = 0; a < A; a++) {
      T* s = p.agent_state + (env * N + ga * A + a) * 4;
      st2<T>(s, px[a], py[a]); st2<T>(s + 2, vx[a], vy[a]);
    }
    if (g == 0) {
      p.t[env] = t_now;
      if (auto_reset && ep != ep0) {
#pragma unroll
        for (int l = 0; l < L; l++) { T* m = p.lm_pos + (env * L + l) * 2; m[0] = mx[l]; m[1] = my[l]; }
        p.episode[env] = ep;
      }
    }
  }
}

}  // namespace gsm
