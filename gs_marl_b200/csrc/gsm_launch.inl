// gsm_launch.inl — launch glue, included by each precision's translation unit after
// defining GSM_REAL and GSM_SFX(name).
#include <cstdlib>
#include <cstring>

namespace gsm {

static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

static bool has_spec(const HostParams& hp);
static bool has_lane(const HostParams& hp);
static bool has_team(const HostParams& hp);
static int next_pow2(int v) { int p = 1; while (p < v) p *= 2; return p; }

int GSM_SFX(plan)(const HostParams& hp, LaunchPlan* plan) {
  const int N = hp.N, E = hp.N + hp.L;
  const bool lsa = hp.scenario != GSM_SCN_NAVIGATION;
  // One lane per (agent, other entity) pair where a warp can hold an env: P = min(32/N,
  // next_pow2(E-1)).  Larger teams get one CTA per env with a warp per agent.
  int cta_env, P;
  if (N <= 12) {
    cta_env = 0; P = 1;
    while (N * P * 2 <= 32 && P * 2 <= next_pow2(E - 1 > 1 ? E - 1 : 1)) P *= 2;
  } else {
    cta_env = 1; P = 32;
  }
  const int fm = env_int("GSM_FORCE_CTA_ENV", -1), fp = env_int("GSM_FORCE_P", 0);
  if (fm >= 0) cta_env = fm;
  if (fp > 0) P = fp;
  if (P != 1 && P != 2 && P != 4 && P != 8 && P != 16 && P != 32) return (int)cudaErrorInvalidValue;
  if (!cta_env && N * P > 32) return (int)cudaErrorInvalidValue;
  if (lsa && N > GSM_MAX_LSA_N) return (int)cudaErrorInvalidValue;
  plan->cta_env = cta_env;
  plan->P = P;
  plan->envs_per_warp = cta_env ? 0 : 32 / (N * P);
  plan->envs_per_cta = cta_env ? 1 : plan->envs_per_warp * (kThreads / 32);
  const SmemLayout lay = make_layout((int)sizeof(GSM_REAL), N, hp.L, E, hp.K, plan->envs_per_cta,
                                     kThreads / P, lsa ? 1 : 0);
  plan->smem = lay.total;
  plan->grid = (hp.n_envs + plan->envs_per_cta - 1) / plan->envs_per_cta;
  if (plan->smem > 227 * 1024) return (int)cudaErrorInvalidValue;
  plan->spec = has_spec(hp) ? 1 : 0;
  plan->team = has_team(hp) ? 1 : 0;   // polygon / line steps; observe stays on the specialised kernel
  // preference: lane-per-agent (N >= GSM_LANE_MIN_N) > specialised (wide instance first) > generic
  plan->lane = has_lane(hp) ? 1 : 0;
  if (plan->lane) plan->spec = 0;
  return 0;
}

template <int P, bool CTA_ENV, bool PHYS>
static int launch_one(const KParams<GSM_REAL>& kp, const LaunchPlan& plan, cudaStream_t st) {
  auto k = env_kernel<GSM_REAL, P, CTA_ENV, PHYS>;
  if (plan.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (e != cudaSuccess) return (int)e;
  }
  k<<<(unsigned)plan.grid, kThreads, plan.smem, st>>>(kp);
  return (int)cudaGetLastError();
}

template <int P, bool CTA_ENV>
static int launch_phys(const KParams<GSM_REAL>& kp, const LaunchPlan& plan, int physics, cudaStream_t st) {
  return physics ? launch_one<P, CTA_ENV, true>(kp, plan, st) : launch_one<P, CTA_ENV, false>(kp, plan, st);
}

static void fill_kparams(KParams<GSM_REAL>& kp, const HostParams& hp, const gsm_step_io& io,
                         int envs_per_warp, const uint8_t* mask, int64_t mask_stride) {
  typedef GSM_REAL T;
  std::memset(&kp, 0, sizeof(kp));
  kp.n_envs = hp.n_envs; kp.env_offset = hp.env_offset;
  kp.N = hp.N; kp.L = hp.L; kp.E = hp.N + hp.L; kp.K = hp.K; kp.W = (kp.E + 31) / 32;
  kp.scenario = hp.scenario; kp.action_mode = hp.action_mode; kp.n_actions = hp.n_actions;
  kp.episode_length = hp.episode_length;
  kp.share_reward = hp.share_reward; kp.cost_obstacles = hp.cost_obstacles;
  kp.own_goal_always = hp.own_goal_always;
  kp.envs_per_warp = envs_per_warp;
  kp.dt = (T)hp.dt; kp.one_minus_damp = (T)1 - (T)hp.damping;
  kp.cf = (T)hp.cf; kp.km = (T)hp.km; kp.km_inv = (T)1 / (T)hp.km; kp.Rs = (T)hp.Rs;
  kp.w_dist = (T)hp.w_dist; kp.w_goal = (T)hp.w_goal; kp.goal_tol = (T)hp.goal_tol;
  kp.poly_r = (T)hp.poly_r;
  for (int a = 0; a < GSM_MAX_DISCRETE; a++) {
    kp.discrete_u[a][0] = (T)hp.discrete_u[a][0];
    kp.discrete_u[a][1] = (T)hp.discrete_u[a][1];
  }
  kp.size = (const T*)hp.size; kp.eflag = hp.eflag; kp.mass = (const T*)hp.mass;
  kp.accel = (const T*)hp.accel; kp.max_speed = (const T*)hp.max_speed;
  kp.slot_table = (const T*)hp.slot_table;
  kp.agent_state = (T*)hp.agent_state; kp.lm_pos = (T*)hp.lm_pos; kp.t = hp.t;
  kp.mask = mask; kp.mask_stride = mask_stride;
  kp.auto_reset = hp.auto_reset; kp.seed = hp.seed; kp.episode = hp.episode;
  for (int k = 0; k < 4; k++) kp.ext[k] = hp.ext[k];
  kp.actions = io.actions;
  kp.obs = (T*)io.obs; kp.nbr_idx = io.nbr_idx; kp.nbr_feat = (T*)io.nbr_feat;
  kp.nbr_cnt = io.nbr_cnt; kp.adj = io.adj; kp.reward = (T*)io.reward; kp.cost = (T*)io.cost;
  kp.done = io.done; kp.assign = io.assign;
}

int GSM_SFX(launch_env)(const HostParams& hp, const LaunchPlan& plan, const gsm_step_io& io,
                        int physics, const uint8_t* mask, int64_t mask_stride, cudaStream_t st) {
  if (hp.n_envs == 0) return 0;
  KParams<GSM_REAL> kp;
  fill_kparams(kp, hp, io, plan.envs_per_warp, mask, mask_stride);
  if (plan.cta_env) {
    switch (plan.P) {
      case 1: return launch_phys<1, true>(kp, plan, physics, st);
      case 2: return launch_phys<2, true>(kp, plan, physics, st);
      case 4: return launch_phys<4, true>(kp, plan, physics, st);
      case 8: return launch_phys<8, true>(kp, plan, physics, st);
      case 16: return launch_phys<16, true>(kp, plan, physics, st);
      default: return launch_phys<32, true>(kp, plan, physics, st);
    }
  }
  switch (plan.P) {
    case 1: return launch_phys<1, false>(kp, plan, physics, st);
    case 2: return launch_phys<2, false>(kp, plan, physics, st);
    case 4: return launch_phys<4, false>(kp, plan, physics, st);
    case 8: return launch_phys<8, false>(kp, plan, physics, st);
    case 16: return launch_phys<16, false>(kp, plan, physics, st);
    default: return launch_phys<32, false>(kp, plan, physics, st);
  }
}

// ---- size-specialised instances (gsm_kernels_spec.cuh) ------------------------------------
// (scenario, N, L, P): P = lanes per agent, chosen so that N*P <= 32 and P covers as many of
// the E-1 "other" entities per pass as a warp allows.
#define GSM_SPEC_TABLE(X)                                                              \
  X(GSM_SCN_NAVIGATION, 3, 6, 4) X(GSM_SCN_NAVIGATION, 3, 6, 8) X(GSM_SCN_NAVIGATION, 3, 6, 2) \
  X(GSM_SCN_NAVIGATION, 3, 6, 1)                                                       \
  X(GSM_SCN_NAVIGATION, 6, 12, 4) X(GSM_SCN_NAVIGATION, 6, 12, 1)                      \
  X(GSM_SCN_NAVIGATION, 12, 24, 2)                                                     \
  X(GSM_SCN_POLYGON, 3, 1, 1) X(GSM_SCN_POLYGON, 4, 1, 1) X(GSM_SCN_POLYGON, 5, 1, 1)  \
  X(GSM_SCN_POLYGON, 6, 1, 1) X(GSM_SCN_POLYGON, 6, 1, 4) X(GSM_SCN_POLYGON, 12, 1, 1) \
  X(GSM_SCN_POLYGON, 12, 1, 2)                                                         \
  X(GSM_SCN_LINE, 3, 2, 1) X(GSM_SCN_LINE, 4, 2, 1) X(GSM_SCN_LINE, 5, 2, 1)           \
  X(GSM_SCN_LINE, 6, 2, 1) X(GSM_SCN_LINE, 6, 2, 4) X(GSM_SCN_LINE, 12, 2, 1)          \
  X(GSM_SCN_LINE, 12, 2, 2)

static bool spec_enabled() {
  return env_int("GSM_NO_SPEC", 0) == 0 && env_int("GSM_FORCE_P", 0) == 0 &&
         env_int("GSM_FORCE_CTA_ENV", -1) < 0;
}

// Lanes per agent of the specialised kernel for this handle (0: no instance).  The first
// table entry of a (scenario, N, L) is the default; GSM_SPEC_P picks another compiled one.
static int spec_P(const HostParams& hp) {
  if (!spec_enabled()) return 0;
  const int want = env_int("GSM_SPEC_P", 0);
  int first = 0;
#define X(S, n, l, pp)                                        \
  if (hp.scenario == S && hp.N == n && hp.L == l) {           \
    if (!first) first = pp;                                   \
    if (want == pp) return pp;                                \
  }
  GSM_SPEC_TABLE(X)
#undef X
  return first;
}

static bool has_wide(const HostParams& hp);
static bool has_spec(const HostParams& hp) { return spec_P(hp) != 0 || has_wide(hp); }

template <int SCN, int N, int L, int P>
static int launch_spec_one(const KParams<GSM_REAL>& kp, int n_steps, const StepStrides& ss,
                           bool observe, cudaStream_t st) {
  constexpr int EPW = 32 / (N * P), WPC = kSpecThreads / 32;
  const int64_t grid = (kp.n_envs + EPW * WPC - 1) / (EPW * WPC);
  const size_t smem = SCN == GSM_SCN_NAVIGATION ? 0 : (size_t)WPC * EPW * N * N * sizeof(GSM_REAL);
  if (observe)
    env_steps_kernel<GSM_REAL, SCN, N, L, P, 1><<<(unsigned)grid, kSpecThreads, smem, st>>>(kp, 1, ss);
  else if (kp.auto_reset)
    env_steps_kernel<GSM_REAL, SCN, N, L, P, 2><<<(unsigned)grid, kSpecThreads, smem, st>>>(kp, n_steps, ss);
  else
    env_steps_kernel<GSM_REAL, SCN, N, L, P, 0><<<(unsigned)grid, kSpecThreads, smem, st>>>(kp, n_steps, ss);
  return (int)cudaGetLastError();
}

// ---- one-lane-per-other navigation kernel (gsm_kernels_wide.cuh) ---------------------------------
// (N, L) instances; preferred over the (N*P lanes per env) specialised kernel where one exists.
// K: max_nbrs the instance is compiled for (0: any, read at run time)
#define GSM_WIDE_TABLE(X) X(3, 6, 8) X(3, 6, 0) X(4, 8, 0) X(5, 10, 0)

static bool has_wide(const HostParams& hp) {
  if (!spec_enabled() || env_int("GSM_NO_WIDE", 0) != 0 || env_int("GSM_SPEC_P", 0) != 0) return false;
  if (hp.scenario != GSM_SCN_NAVIGATION || !hp.h_consts) return false;
  for (int i = 1; i < hp.N; i++)          // the agents must be alike (size, collide): chunk-independent lane constants
    if (hp.h_size[i] != hp.h_size[0] || (hp.h_eflag[i] & 1) != (hp.h_eflag[0] & 1)) return false;
  // 32-bit lane offsets into one slot of the largest output
  if (hp.n_envs * hp.N * hp.K * (int64_t)(GSM_NBR_FEAT_DIM * sizeof(GSM_REAL)) >= (1ll << 31)) return false;
#define X(n, l, k) if (hp.N == n && hp.L == l) return true;
  GSM_WIDE_TABLE(X)
#undef X
  return false;
}

template <int N, int L, int KT>
static int launch_wide_one(const HostParams& hp, const KParams<GSM_REAL>& kp, int n_steps, const StepStrides& ss,
                           bool observe, cudaStream_t st) {
  typedef GSM_REAL T;
  constexpr int E = N + L, EPW = 32 / wide_pow2(E - 1), WPC = kWideThreads / 32;
  WideConsts<T, N, E> wc;
  for (int e = 0; e < E; e++) { wc.size[e] = (T)hp.h_size[e]; wc.eflag[e] = hp.h_eflag[e]; }
  for (int i = 0; i < N; i++) {
    wc.mass[i] = (T)hp.h_mass[i]; wc.mass_inv[i] = (T)1 / (T)hp.h_mass[i];
    wc.accel[i] = (T)hp.h_accel[i]; wc.maxsp[i] = (T)hp.h_maxsp[i];
  }
  const int64_t grid = (kp.n_envs + EPW * WPC - 1) / (EPW * WPC);
  const size_t smem = wide_smem_bytes((int)sizeof(T), N, E, kp.K, EPW);
  auto k = observe ? env_wide_kernel<T, N, L, 1, KT> : (kp.auto_reset ? env_wide_kernel<T, N, L, 2, KT> : env_wide_kernel<T, N, L, 0, KT>);
  // function attributes are per device: one bit per (instance, device), set on first use there
  static unsigned long long attr_done[3] = {0, 0, 0};    // per instance (this function is one per (T, N, L, K))
  const int which = observe ? 1 : (kp.auto_reset ? 2 : 0);
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !((attr_done[which] >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem > 48 * 1024 ? (int)smem : 48 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) attr_done[which] |= 1ull << dev;
  }
  k<<<(unsigned)grid, kWideThreads, smem, st>>>(kp, observe ? 1 : n_steps, ss, wc, wide_smem_layout((int)sizeof(T), N, E, kp.K, EPW));
  return (int)cudaGetLastError();
}

int GSM_SFX(launch_spec)(const HostParams& hp, const gsm_step_io& io, int n_steps,
                         const RolloutStrides& rs, int observe, const uint8_t* mask,
                         int64_t mask_stride, cudaStream_t st) {
  const int P = spec_P(hp);
  if (!P && !has_wide(hp)) return -1;
  if (hp.n_envs == 0) return 0;
  KParams<GSM_REAL> kp;
  fill_kparams(kp, hp, io, 0, mask, mask_stride);
  StepStrides ss;
  ss.actions = rs.actions; ss.obs = rs.obs; ss.nbr_idx = rs.nbr_idx; ss.nbr_feat = rs.nbr_feat;
  ss.nbr_cnt = rs.nbr_cnt; ss.adj = rs.adj; ss.reward = rs.reward; ss.cost = rs.cost;
  ss.done = rs.done; ss.assign = rs.assign;
  const bool one_stride = sizeof(GSM_REAL) != 4 || (ss.nbr_cnt == ss.adj && ss.adj == ss.reward && ss.reward == ss.cost &&
                                                    ss.cost == ss.assign);
  if (has_wide(hp) && one_stride) {
#define X(n, l, k) if (hp.N == n && hp.L == l && (k == 0 || hp.K == k)) return launch_wide_one<n, l, k>(hp, kp, n_steps, ss, observe != 0, st);
    GSM_WIDE_TABLE(X)
#undef X
  }
#define X(S, n, l, pp) \
  if (hp.scenario == S && hp.N == n && hp.L == l && P == pp) return launch_spec_one<S, n, l, pp>(kp, n_steps, ss, observe != 0, st);
  GSM_SPEC_TABLE(X)
#undef X
  return -1;
}

// ---- lane-per-agent kernel (gsm_kernels_lane.cuh) ----------------------------------------------
static bool has_lane(const HostParams& hp) {
  if (env_int("GSM_NO_LANE", 0) != 0 || env_int("GSM_FORCE_P", 0) != 0 || env_int("GSM_FORCE_CTA_ENV", -1) >= 0)
    return false;
  if (env_int("GSM_SPEC_P", 0) != 0 && spec_P(hp) != 0) return false;   // an explicit spec variant was asked for
  return hp.scenario == GSM_SCN_NAVIGATION && hp.N >= env_int("GSM_LANE_MIN_N", 6) && hp.N <= 128;
}

int GSM_SFX(launch_lane)(const HostParams& hp, const gsm_step_io& io, int n_steps,
                         const RolloutStrides& rs, cudaStream_t st) {
  if (!has_lane(hp)) return -1;
  if (hp.n_envs == 0) return 0;
  KParams<GSM_REAL> kp;
  fill_kparams(kp, hp, io, 0, nullptr, 0);
  StepStrides ss;
  ss.actions = rs.actions; ss.obs = rs.obs; ss.nbr_idx = rs.nbr_idx; ss.nbr_feat = rs.nbr_feat;
  ss.nbr_cnt = rs.nbr_cnt; ss.adj = rs.adj; ss.reward = rs.reward; ss.cost = rs.cost;
  ss.done = rs.done; ss.assign = rs.assign;
  const LaneGeom g = lane_geom(hp.N);
  const size_t smem = lane_smem((int)sizeof(LaneEnt<GSM_REAL>), (int)sizeof(GSM_REAL), hp.N, hp.N + hp.L, hp.K,
                                g.envs_per_cta, g.warps_per_cta);
  const bool carry = sizeof(GSM_REAL) == 4 && hp.N >= env_int("GSM_LANE_CARRY_MIN_N", 48);
  auto k = carry ? (hp.auto_reset ? env_lane_kernel<GSM_REAL, true, true> : env_lane_kernel<GSM_REAL, false, true>)
                 : (hp.auto_reset ? env_lane_kernel<GSM_REAL, true, false> : env_lane_kernel<GSM_REAL, false, false>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int64_t grid = (hp.n_envs + g.envs_per_cta - 1) / g.envs_per_cta;
  k<<<(unsigned)grid, g.warps_per_cta * 32, smem, st>>>(kp, n_steps, ss);
  return (int)cudaGetLastError();
}

// ---- polygon / line kernel with the group-parallel assignment (gsm_kernels_team.cuh) -----------
// (scenario, N, G, GP): G working lanes in a physical group of GP lanes per env, N / G agents =
// assignment rows/columns per lane.  The first entry of an (scenario, N) is the default,
// GSM_TEAM_G picks another compiled one.
#define GSM_TEAM_TABLE(X)                                                                     \
  X(GSM_SCN_POLYGON, 3, 1, 1) X(GSM_SCN_POLYGON, 4, 2, 2) X(GSM_SCN_POLYGON, 5, 1, 1)           \
  X(GSM_SCN_POLYGON, 6, 6, 8) X(GSM_SCN_POLYGON, 6, 3, 4) X(GSM_SCN_POLYGON, 6, 2, 2)           \
  X(GSM_SCN_POLYGON, 12, 4, 4) X(GSM_SCN_POLYGON, 12, 6, 8)                                     \
  X(GSM_SCN_LINE, 3, 1, 1) X(GSM_SCN_LINE, 4, 2, 2) X(GSM_SCN_LINE, 5, 1, 1)                    \
  X(GSM_SCN_LINE, 6, 6, 8) X(GSM_SCN_LINE, 6, 3, 4) X(GSM_SCN_LINE, 6, 2, 2)                    \
  X(GSM_SCN_LINE, 12, 4, 4) X(GSM_SCN_LINE, 12, 6, 8)
// measured on B200 (profiles/README.md): round 1 (lsa_group): N = 12: G=4 70 us, G=6 of 8 lanes 117 us, G=12 of
// 16 lanes 94 us per step; N = 6: G=3 of 4 lanes 19.6 us, G=2 20.6 us, G=6 of 8 lanes 23.8 us.  Round 2
// (lsa_group2, one launch / 4 sub-shard streams): N = 12: G=4 60.2 / 59.0 us, G=6 82.1 / 69.4; N = 6: G=3
// 18.9 / 19.0, G=6 (one agent per lane, GSM_TEAM_G=6) 22.4 / 16.1 — faster only when several launches overlap
// (27.7 warps per SM do not fit one wave at 122 registers); compiled for 5 CTAs per SM (68 registers) G=6 is
// 17.9 / 16.4 (line-6: 20.9 / 19.0 vs 20.0 / 20.0 with G=3) and is the N = 6 default since.  N = 12 with G=12 of 16 lanes (lsa_group2): 83 / 70 us.

static bool has_team(const HostParams& hp) {
  if (env_int("GSM_NO_TEAM", 0) != 0 || env_int("GSM_NO_SPEC", 0) != 0 || env_int("GSM_FORCE_P", 0) != 0 ||
      env_int("GSM_FORCE_CTA_ENV", -1) >= 0 || env_int("GSM_SPEC_P", 0) != 0)
    return false;
#define X(S, n, gg, gp) if (hp.scenario == S && hp.N == n) return true;
  GSM_TEAM_TABLE(X)
#undef X
  return false;
}

template <int SCN, int N, int G, int GP>
static int launch_team_one(const KParams<GSM_REAL>& kp, int n_steps, const StepStrides& ss, cudaStream_t st) {
  constexpr int EPW = 32 / GP, WPC = kTeamThreads / 32;
  const int64_t grid = (kp.n_envs + EPW * WPC - 1) / (EPW * WPC);
  const size_t smem = (size_t)WPC * EPW * team_env_bytes((int)sizeof(GSM_REAL), N);
  auto k = env_team_kernel<GSM_REAL, SCN, N, G, GP>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  k<<<(unsigned)grid, kTeamThreads, smem, st>>>(kp, n_steps, ss);
  return (int)cudaGetLastError();
}

int GSM_SFX(launch_team)(const HostParams& hp, const gsm_step_io& io, int n_steps,
                         const RolloutStrides& rs, cudaStream_t st) {
  if (!has_team(hp)) return -1;
  if (hp.n_envs == 0) return 0;
  KParams<GSM_REAL> kp;
  fill_kparams(kp, hp, io, 0, nullptr, 0);
  StepStrides ss;
  ss.actions = rs.actions; ss.obs = rs.obs; ss.nbr_idx = rs.nbr_idx; ss.nbr_feat = rs.nbr_feat;
  ss.nbr_cnt = rs.nbr_cnt; ss.adj = rs.adj; ss.reward = rs.reward; ss.cost = rs.cost;
  ss.done = rs.done; ss.assign = rs.assign;
  const int want = env_int("GSM_TEAM_G", 0);
#define X(S, n, gg, gp) if (hp.scenario == S && hp.N == n && want == gg) return launch_team_one<S, n, gg, gp>(kp, n_steps, ss, st);
  GSM_TEAM_TABLE(X)
#undef X
#define X(S, n, gg, gp) if (hp.scenario == S && hp.N == n) return launch_team_one<S, n, gg, gp>(kp, n_steps, ss, st);
  GSM_TEAM_TABLE(X)
#undef X
  return -1;
}

int GSM_SFX(launch_reset)(const HostParams& hp, uint64_t seed, const uint8_t* mask,
                          int64_t mask_stride, cudaStream_t st) {
  if (hp.n_envs == 0) return 0;
  ResetParams rp;
  std::memset(&rp, 0, sizeof(rp));
  rp.n_envs = hp.n_envs; rp.env_offset = hp.env_offset; rp.N = hp.N; rp.L = hp.L; rp.seed = seed;
  for (int k = 0; k < 4; k++) rp.ext[k] = hp.ext[k];
  rp.eflag = hp.eflag; rp.mask = mask; rp.mask_stride = mask_stride;
  rp.agent_state = hp.agent_state; rp.lm_pos = hp.lm_pos; rp.t = hp.t; rp.episode = hp.episode;
  const int64_t total = hp.n_envs * (hp.N + hp.L);
  reset_kernel<GSM_REAL><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(rp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  reset_counters_kernel<GSM_REAL><<<(unsigned)((hp.n_envs + 255) / 256), 256, 0, st>>>(
      hp.n_envs, mask, mask_stride, hp.t, hp.episode);
  return (int)cudaGetLastError();
}

int GSM_SFX(launch_lsa)(const void* cost, int32_t* col4row, int64_t n_problems, int n,
                        cudaStream_t st) {
  if (n_problems == 0) return 0;
  int G = 4;
  while (G < n) G *= 2;
  const int threads = 128, per_cta = threads / G;
  const size_t smem = (size_t)per_cta * n * n * sizeof(GSM_REAL);
  lsa_kernel<GSM_REAL><<<(unsigned)((n_problems + per_cta - 1) / per_cta), threads, smem, st>>>(
      (const GSM_REAL*)cost, col4row, n_problems, n, G);
  return (int)cudaGetLastError();
}

}  // namespace gsm
