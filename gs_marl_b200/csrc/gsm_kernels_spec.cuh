// gsm_kernels_spec.cuh — register-resident, size-specialised env kernel (SPEC.md §2-7).
//
// Same arithmetic as env_kernel in gsm_kernels.cuh, for team sizes known at compile time
// (N agents, L landmarks, P lanes per agent, N*P <= 32).  One lane owns one (agent, other
// entity) pair per chunk; an env lives in N*P consecutive lanes of one warp and its whole
// state stays in REGISTERS across `n_steps` consecutive steps:
//   * other agents' positions/velocities are fetched with warp shuffles,
//   * the pair force is reduced with xor-shuffles, the integration is done redundantly by
//     the P lanes of the agent (no broadcast, no divergence),
//   * neighbour rows are compacted with ballot+popc and written straight to HBM, the
//     adjacency word is assembled from the ballot bits,
//   * no shared memory and no block barrier on the navigation path (polygon/line keep the
//     N x N assignment matrix in shared memory for the in-warp LSA).
// With n_steps > 1 this is the fused rollout kernel: per step it only reads the actions
// and streams the outputs to slot s of the [T][...] rollout buffers.
#pragma once
#include "gsm_kernels.cuh"

namespace gsm {

constexpr int kSpecThreads = 128;

template <typename T> struct Vec2;
template <> struct Vec2<float> { typedef float2 type; };
template <> struct Vec2<double> { typedef double2 type; };

template <typename T>
__device__ __forceinline__ void st2(T* p, T a, T b) {
  typename Vec2<T>::type v; v.x = a; v.y = b;
  *reinterpret_cast<typename Vec2<T>::type*>(p) = v;
}

// One 24-byte fp32 feature row as ONE 16-byte + ONE 8-byte store instead of three 8-byte stores.
// Row `pos` starts at pos * 24 bytes behind a 16-byte aligned agent block: even rows are 16-byte
// aligned at their start, odd rows at start + 8, so every lane issues the same two instructions
// (st.v4 then st.v2) with parity-selected addresses and operands — no divergence, a third fewer
// store instructions and L1TEX line passes for the largest output.
__device__ __forceinline__ void st_row6_f32(float* f, int pos, float a0, float a1, float a2, float a3,
                                            float a4, float a5) {
  const bool odd = pos & 1;
  float4 q;
  q.x = odd ? a2 : a0; q.y = odd ? a3 : a1; q.z = odd ? a4 : a2; q.w = odd ? a5 : a3;
  *reinterpret_cast<float4*>(f + (odd ? 2 : 0)) = q;
  float2 d;
  d.x = odd ? a0 : a4; d.y = odd ? a1 : a5;
  *reinterpret_cast<float2*>(f + (odd ? 0 : 4)) = d;
}

// Arithmetic policy.  fp64 (verification): exactly the operations SPEC.md writes.  fp32
// (production, 1e-4 relative): reciprocal-multiply for the two constant divisors and the
// approximate SFU sqrt / divide (<= 2 ulp) — no IEEE slow paths in the step loop.
template <typename T> struct Arith;
template <> struct Arith<double> {
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double div(double a, double b) { return a / b; }
  static __device__ __forceinline__ double div_const(double a, double b, double /*b_inv*/) { return a / b; }
};
template <> struct Arith<float> {
  static __device__ __forceinline__ float sqrt(float x) {
    float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
  }
  static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
  static __device__ __forceinline__ float div_const(float a, float /*b*/, float b_inv) { return a * b_inv; }
};

// Per-step byte strides of the rollout buffers (0 for single-step use).
struct StepStrides {
  int64_t actions, obs, nbr_idx, nbr_feat, nbr_cnt, adj, reward, cost, done, assign;
};

template <int E> struct AdjBits { typedef uint64_t type; };
template <> struct AdjBits<0> { typedef uint32_t type; };

// Every output pointer must be non-NULL (the API layer falls back to the generic kernel
// otherwise).  All 32 lanes stay alive and use FULL-mask warp primitives: lanes without a
// real env compute on a clamped env index and only their stores are predicated off.
// OBS = true: observe only (reset path) — no physics, no reward/cost/done, one "step",
// optional per-env mask (p.mask); the state is not written back.
// Resident CTAs per SM the compiler must allow for.  The fp32 navigation-3 instances get 7
// (<= 72 registers -> 28 warps/SM): at the bench batch of 16384 envs (55.4 warps per SM) that
// is two full rounds of warps instead of 2.3 rounds at 24 warps/SM (ncu_r1_v6: 80 registers).
#ifndef GSM_SPEC_P2_BLOCKS      // A/B: resident-CTA target of the 2-lanes-per-agent nav-3 instance
#define GSM_SPEC_P2_BLOCKS 1
#endif
template <typename T, int N, int P> struct SpecMinBlocks {
  static constexpr int value = (sizeof(T) == 4 && N == 3 && (P == 4 || P == 8)) ? 7
                               : (sizeof(T) == 4 && N == 3 && P == 2) ? GSM_SPEC_P2_BLOCKS : 1;
};

// MODE 0: step(s).  MODE 1 (OBS): observe only.  MODE 2: steps with compiled-in auto-reset
// (fused rollouts after gsm_set_auto_reset) — kept out of MODE 0 so the common loop carries
// none of it.
template <typename T, int SCN, int N, int L, int P, int MODE>
__global__ void __launch_bounds__(kSpecThreads, SpecMinBlocks<T, N, P>::value)
env_steps_kernel(const __grid_constant__ KParams<T> p, const int n_steps,
                 const __grid_constant__ StepStrides ss) {
  constexpr int E = N + L, M = E - 1, LPE = N * P, EPW = 32 / LPE;
  constexpr int CH = (M + P - 1) / P;              // chunks of "others" per lane
  constexpr int W = (E + 31) / 32;
  constexpr bool LSA = SCN != GSM_SCN_NAVIGATION;
  constexpr bool OBS = MODE == 1;
  constexpr unsigned FULL = 0xffffffffu;
  static_assert(LPE <= 32 && EPW >= 1, "an env must fit in one warp");
  static_assert(E <= 64, "adjacency is assembled in one 64-bit register");
  typedef typename AdjBits<(E <= 32 ? 0 : E)>::type adj_t;
  typedef Arith<T> A;
  const int K = p.K;
  // 16-byte stores into the feature rows need every agent block (K * 24 bytes) and every slot 16-byte aligned
  const bool row16 = sizeof(T) == 4 && (K & 1) == 0 && (((uintptr_t)p.nbr_feat | (uintptr_t)ss.nbr_feat) & 15) == 0;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int eiw = lane / LPE;                       // compile-time divisor
  const int off = lane - eiw * LPE;
  const int i = off / P, sub = off % P;             // my agent, my lane in its group
  const int64_t env_raw = ((int64_t)blockIdx.x * (kSpecThreads / 32) + warp) * EPW + eiw;
  bool active = eiw < EPW && env_raw < p.n_envs;
  if (OBS && p.mask != nullptr && active) active = p.mask[env_raw * p.mask_stride] != 0;
  const int64_t env = active ? env_raw : 0;         // idle lanes shadow env 0, never store
  const int env_base = eiw * LPE;
  const int grp_base = env_base + i * P;
  const unsigned envmask = low_mask(LPE) << (env_base & 31);

  extern __shared__ __align__(16) unsigned char smem_spec[];
  T* s_cm = nullptr;                                // [EPW per warp][N*N] assignment costs
  if (LSA) s_cm = (T*)smem_spec + ((size_t)warp * EPW + (eiw < EPW ? eiw : 0)) * N * N;

  // ---- per-lane constants of my pairs ---------------------------------------------------
  const T size_i = p.size[i];
  const bool coll_i = p.eflag[i] & 1;
  const T mass_i = p.mass[i], accel_i = p.accel[i], maxsp_i = p.max_speed[i];
  const T mass_inv = (T)1 / mass_i;
  int e_c[CH];                                       // entity of my pair in chunk c (-1: none)
  int src_c[CH];                                     // lane holding that entity if it is an agent
  T dmin_c[CH], type_c[CH];
  bool cpair_c[CH], colc_c[CH], valid_c[CH], goal_c[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) {
    const int o = c * P + sub;
    const bool valid = o < M;
    const int e = o + (o >= i ? 1 : 0);
    const int fl = valid ? (int)p.eflag[e] : 0;
    e_c[c] = valid ? e : -1;
    src_c[c] = env_base + ((valid && e < N) ? e : 0) * P;
    dmin_c[c] = valid ? size_i + p.size[e] : (T)0;
    type_c[c] = (T)(fl >> 1);
    cpair_c[c] = valid && coll_i && (fl & 1);
    colc_c[c] = valid && (e < N || (p.cost_obstacles && (fl >> 1) == GSM_ENT_OBSTACLE));
    valid_c[c] = valid;
    goal_c[c] = valid && !LSA && p.own_goal_always && e == N + i;   // own goal: always a neighbour
  }
  // lane-role masks for the per-agent scalar stores
  const uint32_t rm0 = sub == 0 ? ~0u : 0u, rm1 = sub == 1 ? ~0u : 0u, rm2 = sub == 2 ? ~0u : 0u,
                 rm3 = sub == 3 ? ~0u : 0u, rm4 = sub == 4 ? ~0u : 0u;

  // ---- state into registers -----------------------------------------------------------------
  T px, py, vx, vy;
  {
    const T* a = p.agent_state + (env * N + i) * 4;
    px = a[0]; py = a[1]; vx = a[2]; vy = a[3];
  }
  T lmx[CH], lmy[CH];                                // landmark positions of my landmark pairs
#pragma unroll
  for (int c = 0; c < CH; c++) {
    lmx[c] = 0; lmy[c] = 0;
    if (e_c[c] >= N) {
      const T* l = p.lm_pos + (env * L + (e_c[c] - N)) * 2;
      lmx[c] = l[0]; lmy[c] = l[1];
    }
  }
  T goalx = 0, goaly = 0;                            // navigation: own goal
  if (!LSA) { const T* l = p.lm_pos + (env * L + i) * 2; goalx = l[0]; goaly = l[1]; }
  T slotx = 0, sloty = 0;                            // polygon / line: slot `off` (lanes off < N)
  if (LSA && off < N) {
    const T* l = p.lm_pos + env * L * 2;
    if (SCN == GSM_SCN_POLYGON) {
      slotx = l[0] + p.poly_r * p.slot_table[2 * off];
      sloty = l[1] + p.poly_r * p.slot_table[2 * off + 1];
    } else {
      const T f = p.slot_table[2 * off];
      slotx = l[0] + f * (l[2] - l[0]);
      sloty = l[1] + f * (l[3] - l[1]);
    }
  }
  int t_now = p.t[env];
  // in-kernel auto-reset (fused rollouts only): SPEC §8 draws with the counters gsm_reset uses
  const bool auto_reset = MODE == 2 && p.auto_reset != 0;
  int ep = auto_reset ? p.episode[env] : 0;
  const int ep0 = ep;
  const uint64_t genv = (uint64_t)(p.env_offset + env);

  // ---- per-lane output cursors (advanced by the slot strides every step) -----------------------
  const int64_t row = env * N + i;
  const unsigned char* c_act = (const unsigned char*)p.actions +
      (p.action_mode == GSM_ACT_DISCRETE ? row * 4 : row * 2 * (int64_t)sizeof(T));
  unsigned char* c_idx = (unsigned char*)(p.nbr_idx + row * K);
  unsigned char* c_feat = (unsigned char*)(p.nbr_feat + row * K * GSM_NBR_FEAT_DIM);
  // obs: lane sub (and sub + P, ...) writes pair `part` of (vx,vy | px,py | gx,gy)
  unsigned char* c_obs = (unsigned char*)(p.obs + row * GSM_OBS_DIM + 2 * (sub % 3));
  constexpr int OBS_PASSES = (3 + P - 1) / P;
  // per-agent scalars by "lane roles": lane sub of the agent's group owns one output
  // (0 cnt, 1 adj word 0, 2 reward, 3 cost; with P >= 8 also 4 assign, 5 done), so a step issues
  // one 32-bit (fp64: plus one 64-bit) store for them instead of one store per output on lane 0.
  // With P == 4 lane 0 additionally walks the assign and done cursors; P < 4 walks all six.
  // (measured per variant on nav-3, P = 4: the plain MODE 0 loop is 3 % faster with lane 0 walking
  // six cursors, the auto-reset MODE 2 loop 10 % faster with 4 roles — register pressure differs)
  constexpr bool ROLES = (P >= 8 || (P >= 4 && MODE == 2)) && W == 1;
  constexpr int NROLES = P >= 8 ? 6 : 4;
  unsigned char* c_role = nullptr;
  int64_t role_stride = 0;
  unsigned char *c_cnt = nullptr, *c_adj = nullptr, *c_rew = nullptr, *c_cost = nullptr,
                *c_asg = nullptr, *c_done = nullptr;
  if (ROLES) {
    switch (sub) {
      case 0: c_role = (unsigned char*)(p.nbr_cnt + row); role_stride = ss.nbr_cnt; break;
      case 1: c_role = (unsigned char*)(p.adj + row); role_stride = ss.adj; break;
      case 2: c_role = (unsigned char*)(p.reward + row); role_stride = ss.reward; break;
      case 3: c_role = (unsigned char*)(p.cost + row); role_stride = ss.cost; break;
      case 4: if (NROLES > 4) { c_role = (unsigned char*)(p.assign + row); role_stride = ss.assign; } break;
      case 5: if (NROLES > 4) { c_role = (unsigned char*)(p.done + row); role_stride = ss.done; } break;
      default: break;
    }
    if (NROLES == 4) { c_asg = (unsigned char*)(p.assign + row); c_done = (unsigned char*)(p.done + row); }
  } else {
    c_cnt = (unsigned char*)(p.nbr_cnt + row); c_adj = (unsigned char*)(p.adj + row * W);
    c_rew = (unsigned char*)(p.reward + row); c_cost = (unsigned char*)(p.cost + row);
    c_asg = (unsigned char*)(p.assign + row); c_done = (unsigned char*)(p.done + row);
  }

  int act_next = 0;
  T actx_next = 0, acty_next = 0;
  if (!OBS) {
    if (p.action_mode == GSM_ACT_DISCRETE) act_next = *(const int32_t*)c_act;
    else { actx_next = ((const T*)c_act)[0]; acty_next = ((const T*)c_act)[1]; }
  }

  for (int step = 0; step < n_steps; step++) {
    // ---- SPEC §2: action force; the next step's action is prefetched -------------------------
    T fx = 0, fy = 0;
    if (!OBS) {
      T ux = actx_next, uy = acty_next;
      if (p.action_mode == GSM_ACT_DISCRETE) {
        const int a = act_next;
        ux = 0; uy = 0;
        if (a >= 0 && a < p.n_actions) { ux = p.discrete_u[a][0]; uy = p.discrete_u[a][1]; }
      }
      if (sub == 0) { fx = accel_i * ux; fy = accel_i * uy; }
      // Unconditional load (the last step re-reads its own, valid, address): a load under
      // `if (step + 1 < n_steps)` becomes LDG + predicated MOV in the same iteration and the
      // warp then waits out the whole memory latency here instead of one step later.
      if (step + 1 < n_steps) c_act += ss.actions;
      if (p.action_mode == GSM_ACT_DISCRETE) act_next = *(const int32_t*)c_act;
      else { actx_next = ((const T*)c_act)[0]; acty_next = ((const T*)c_act)[1]; }
    }
    // ---- SPEC §3: pair forces ---------------------------------------------------------------
#pragma unroll
    for (int c = 0; c < (OBS ? 0 : CH); c++) {
      T qx = lmx[c], qy = lmy[c];
      if (c * P < N - 1) {                           // this chunk can hold agent pairs
        const T ax = shfl(FULL, px, src_c[c]), ay = shfl(FULL, py, src_c[c]);
        if (e_c[c] >= 0 && e_c[c] < N) { qx = ax; qy = ay; }
      }
      if (cpair_c[c]) {
        const T dx = px - qx, dy = py - qy;
        const T dist = A::sqrt(dx * dx + dy * dy);
        const T x = A::div_const(-(dist - dmin_c[c]), p.km, p.km_inv);
        if (!(Prec<T>::kCut && x < (T)(-kFarCut))) {
          const T pen = softplus(x) * p.km;
          fx = fx + A::div(p.cf * dx, dist) * pen;
          fy = fy + A::div(p.cf * dy, dist) * pen;
        }
      }
    }
    if (P > 1 && !OBS) {
#pragma unroll
      for (int m = P / 2; m >= 1; m >>= 1) {
        fx += __shfl_xor_sync(FULL, fx, m);
        fy += __shfl_xor_sync(FULL, fy, m);
      }
    }
    // ---- SPEC §4: integration, redundantly on the P lanes of the agent --------------------
    if (!OBS) {
    vx = vx * p.one_minus_damp; vy = vy * p.one_minus_damp;
    vx = vx + A::div_const(fx, mass_i, mass_inv) * p.dt;
    vy = vy + A::div_const(fy, mass_i, mass_inv) * p.dt;
    if (maxsp_i > (T)0) {
      const T sp = A::sqrt(vx * vx + vy * vy);
      if (sp > maxsp_i) { vx = A::div(vx, sp) * maxsp_i; vy = A::div(vy, sp) * maxsp_i; }
    }
    px = px + vx * p.dt; py = py + vy * p.dt;
    t_now += 1;
    }

    // ---- SPEC §5: assignment (polygon / line) -----------------------------------------------
    T tx = goalx, ty = goaly;
    int asg = i;
    if (LSA) {
      // lane off < N is column `off`: cost of every agent row to my slot
#pragma unroll
      for (int r = 0; r < N; r++) {
        const T rx = shfl(FULL, px, env_base + r * P), ry = shfl(FULL, py, env_base + r * P);
        if (off < N && eiw < EPW) {                  // shadow lanes must not touch the matrix
          const T dx = slotx - rx, dy = sloty - ry;
          s_cm[r * N + off] = r_sqrt(dx * dx + dy * dy);
        }
      }
      __syncwarp();
      // all envs of the warp are solved in lockstep; lanes off < N are the columns / rows
      const int a_row = lsa_seg<T, N>(s_cm, off, env_base, eiw < EPW);
      __syncwarp();
      asg = shfl(FULL, a_row, env_base + i);
      tx = shfl(FULL, slotx, env_base + asg);
      ty = shfl(FULL, sloty, env_base + asg);
    }

    // ---- SPEC §6: neighbour graph -------------------------------------------------------------
    // Every (lane, chunk) pair owns exactly one of the K output rows: neighbours take rows
    // [0, cnt) in entity order (ballot + popc), non-neighbours the zero rows behind them —
    // no separate padding pass.  TWO_PASS keeps the pair features in registers between the
    // ballot sweep and the row writes (small CH); otherwise rows are written as found and
    // the padding is a short loop.
    constexpr bool TWO_PASS = CH <= 4;
    int ncol = 0;
    adj_t obits = 0;                                  // neighbour bits over "others" index o
    T g_dx[TWO_PASS ? CH : 1], g_dy[TWO_PASS ? CH : 1], g_dvx[TWO_PASS ? CH : 1],
      g_dvy[TWO_PASS ? CH : 1], g_d[TWO_PASS ? CH : 1];
    int cnt_run = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) {
      T ex = lmx[c], ey = lmy[c], evx = 0, evy = 0;
      if (c * P < N - 1) {
        const T ax = shfl(FULL, px, src_c[c]), ay = shfl(FULL, py, src_c[c]);
        const T bx = shfl(FULL, vx, src_c[c]), by = shfl(FULL, vy, src_c[c]);
        if (e_c[c] >= 0 && e_c[c] < N) { ex = ax; ey = ay; evx = bx; evy = by; }
      }
      const T dx = ex - px, dy = ey - py;
      const T dist = A::sqrt(dx * dx + dy * dy);
      const bool nb = (valid_c[c] && dist < p.Rs) || goal_c[c];
      const bool col = colc_c[c] && dist < dmin_c[c];
      unsigned bits, cbits;
      if (P == 1) { bits = nb ? 1u : 0u; cbits = col ? 1u : 0u; }
      else {
        bits = (__ballot_sync(FULL, nb) >> grp_base) & low_mask(P);
        cbits = (__ballot_sync(FULL, col) >> grp_base) & low_mask(P);
      }
      ncol += __popc(cbits);
      obits |= (adj_t)bits << (c * P);
      if (TWO_PASS) {
        g_dx[c] = dx; g_dy[c] = dy; g_dvx[c] = evx - vx; g_dvy[c] = evy - vy; g_d[c] = dist;
      } else {
        const int pos = cnt_run + __popc(bits & low_mask(sub));
        if (nb && pos < K && active) {
          ((int32_t*)c_idx)[pos] = e_c[c];
          T* f = (T*)c_feat + pos * GSM_NBR_FEAT_DIM;
          st2<T>(f, dx, dy); st2<T>(f + 2, evx - vx, evy - vy); st2<T>(f + 4, dist, type_c[c]);
        }
        cnt_run += __popc(bits);
      }
    }
    int cnt = sizeof(adj_t) == 4 ? __popc((uint32_t)obits) : __popcll((uint64_t)obits);
    if (TWO_PASS) {
#pragma unroll
      for (int c = 0; c < CH; c++) {
        const int o = c * P + sub;
        const adj_t below = obits & (((adj_t)1 << o) - 1);
        const int rank = sizeof(adj_t) == 4 ? __popc((uint32_t)below) : __popcll((uint64_t)below);
        const bool nb = (obits >> o) & 1;
        const int pos = nb ? rank : cnt + (o - rank);
        if (valid_c[c] && pos < K && active) {
          ((int32_t*)c_idx)[pos] = nb ? e_c[c] : -1;
          T* f = (T*)c_feat + pos * GSM_NBR_FEAT_DIM;
          const T z = (T)0;
          if (sizeof(T) == 4 && row16) {
            st_row6_f32((float*)f, pos, nb ? (float)g_dx[c] : 0.f, nb ? (float)g_dy[c] : 0.f,
                        nb ? (float)g_dvx[c] : 0.f, nb ? (float)g_dvy[c] : 0.f, nb ? (float)g_d[c] : 0.f,
                        nb ? (float)type_c[c] : 0.f);
          } else {
            st2<T>(f, nb ? g_dx[c] : z, nb ? g_dy[c] : z);
            st2<T>(f + 2, nb ? g_dvx[c] : z, nb ? g_dvy[c] : z);
            st2<T>(f + 4, nb ? g_d[c] : z, nb ? type_c[c] : z);
          }
        }
      }
      if (cnt > K) cnt = K;
    } else {
      if (cnt > K) cnt = K;
      if (active) {
        for (int k = cnt + sub; k < K; k += P) {      // padding rows
          ((int32_t*)c_idx)[k] = -1;
          T* f = (T*)c_feat + k * GSM_NBR_FEAT_DIM;
          st2<T>(f, (T)0, (T)0); st2<T>(f + 2, (T)0, (T)0); st2<T>(f + 4, (T)0, (T)0);
        }
      }
    }
    // entity-indexed adjacency: open a zero bit at my own index i
    const adj_t lo_m = ((adj_t)1 << i) - 1;
    const adj_t ebits = (obits & lo_m) | ((obits & ~lo_m) << 1);

    // ---- SPEC §6-7: per-agent outputs, one role per lane ----------------------------------------
    const T gx = tx - px, gy = ty - py;
    const T d = A::sqrt(gx * gx + gy * gy);
    T r = ((T)0 - p.w_dist * d) + (d < p.goal_tol ? p.w_goal : (T)0);
    if (p.share_reward) {
      T s = shfl(FULL, r, env_base);
#pragma unroll
      for (int k = 1; k < N; k++) s = s + shfl(FULL, r, env_base + k * P);
      r = s / (T)N;
    }
    if (active) {
      if (P >= 3) {
        if (sub < 3) st2<T>((T*)c_obs, sub == 0 ? vx : (sub == 1 ? px : gx), sub == 0 ? vy : (sub == 1 ? py : gy));
      } else {
#pragma unroll
        for (int q = 0; q < OBS_PASSES; q++) {
          const int part = sub + q * P;               // 0: (vx,vy) 1: (px,py) 2: (gx,gy)
          if (part < 3) {
            const T a0 = part == 0 ? vx : (part == 1 ? px : gx);
            const T a1 = part == 0 ? vy : (part == 1 ? py : gy);
            st2<T>((T*)(c_obs) + 2 * q * P, a0, a1);
          }
        }
      }
      const uint8_t dn = (uint8_t)(t_now >= p.episode_length);
      if (ROLES) {
        if (sizeof(T) == 4) {
          const uint32_t v = ((uint32_t)cnt & rm0) | ((uint32_t)ebits & rm1) | (__float_as_uint((float)r) & rm2) |
                             (__float_as_uint((float)ncol) & rm3) | ((uint32_t)asg & rm4);
          if (OBS ? (sub == 0 || sub == 1 || (NROLES > 4 && sub == 4)) : (sub < (NROLES > 4 ? 5 : 4))) *(uint32_t*)c_role = v;
        } else {
          const uint32_t v = ((uint32_t)cnt & rm0) | ((uint32_t)ebits & rm1) | ((uint32_t)asg & rm4);
          if (sub == 0 || sub == 1 || (NROLES > 4 && sub == 4)) *(uint32_t*)c_role = v;
          if (!OBS && (sub == 2 || sub == 3)) *(T*)c_role = sub == 2 ? r : (T)ncol;
        }
        if (NROLES > 4) { if (!OBS && sub == 5) *c_role = dn; }
        else if (sub == 0) { *(int32_t*)c_asg = asg; if (!OBS) *c_done = dn; }
      } else if (sub == 0) {
        *(int32_t*)c_cnt = cnt;
        *(uint32_t*)c_adj = (uint32_t)ebits;
        if (W > 1) ((uint32_t*)c_adj)[1] = (uint32_t)((uint64_t)ebits >> 32);
        if (!OBS) { *(T*)c_rew = r; *(T*)c_cost = (T)ncol; *c_done = dn; }
        *(int32_t*)c_asg = asg;
      }
    }
    // ---- episode end inside a fused rollout: re-draw this env (terminal outputs stay in slot s) ----
    if (auto_reset && t_now >= p.episode_length) {
      // cold path: everything is recomputed from (i, sub, off) and global memory here, so the
      // per-chunk constants of the hot loop do not have to stay live for it
      spawn_draw<T>(genv, ep, i, p.seed, p.ext[GSM_ENT_AGENT], px, py);
      vx = 0; vy = 0;
#pragma unroll
      for (int c = 0; c < CH; c++) {
        const int o = c * P + sub;
        const int e = o + (o >= i ? 1 : 0);
        if (o < M && e >= N) spawn_draw<T>(genv, ep, e, p.seed, p.ext[p.eflag[e] >> 1], lmx[c], lmy[c]);
      }
      if (!LSA) spawn_draw<T>(genv, ep, N + i, p.seed, p.ext[p.eflag[N + i] >> 1], goalx, goaly);
      if (LSA && off < N) {
        T ax, ay, bx = 0, by = 0;
        spawn_draw<T>(genv, ep, N, p.seed, p.ext[p.eflag[N] >> 1], ax, ay);
        if (L > 1) spawn_draw<T>(genv, ep, N + 1, p.seed, p.ext[p.eflag[N + 1] >> 1], bx, by);
        if (SCN == GSM_SCN_POLYGON) {
          slotx = ax + p.poly_r * p.slot_table[2 * off];
          sloty = ay + p.poly_r * p.slot_table[2 * off + 1];
        } else {
          const T f = p.slot_table[2 * off];
          slotx = ax + f * (bx - ax);
          sloty = ay + f * (by - ay);
        }
      }
      t_now = 0;
      ep += 1;
    }
    // advance the cursors to the next slot of the rollout buffers
    c_idx += ss.nbr_idx; c_feat += ss.nbr_feat; c_obs += ss.obs;
    if (ROLES) {
      c_role += role_stride;
      if (NROLES == 4) { c_asg += ss.assign; c_done += ss.done; }
    } else {
      c_cnt += ss.nbr_cnt; c_adj += ss.adj; c_rew += ss.reward; c_cost += ss.cost;
      c_asg += ss.assign; c_done += ss.done;
    }
  }

  // ---- state back to HBM ------------------------------------------------------------------------
  if (!OBS && sub == 0 && active) {
    T* a = p.agent_state + row * 4;
    st2<T>(a, px, py); st2<T>(a + 2, vx, vy);
    if (i == 0) p.t[env] = t_now;
  }
  if (auto_reset && ep != ep0 && active) {             // the landmarks of the last re-draw + episode count
    for (int l = off; l < L; l += LPE) {
      T x, y;
      spawn_draw<T>(genv, ep - 1, N + l, p.seed, p.ext[p.eflag[N + l] >> 1], x, y);
      T* lp = p.lm_pos + (env * L + l) * 2;
      lp[0] = x; lp[1] = y;
    }
    if (off == 0) p.episode[env] = ep;
  }
}

}  // namespace gsm
