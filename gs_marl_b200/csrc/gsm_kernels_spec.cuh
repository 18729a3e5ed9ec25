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
//     adjacency word comes from redux.or,
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
template <typename T> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<double> { typedef double4 type; };

template <typename T>
__device__ __forceinline__ void st2(T* p, T a, T b) {
  typename Vec2<T>::type v; v.x = a; v.y = b;
  *reinterpret_cast<typename Vec2<T>::type*>(p) = v;
}

// Per-step byte strides of the rollout buffers (0 for single-step use).
struct StepStrides {
  int64_t actions, obs, nbr_idx, nbr_feat, nbr_cnt, adj, reward, cost, done, assign;
};

template <typename T, int SCN, int N, int L, int P>
__global__ void __launch_bounds__(kSpecThreads)
env_steps_kernel(const __grid_constant__ KParams<T> p, const int n_steps,
                 const __grid_constant__ StepStrides ss) {
  constexpr int E = N + L, M = E - 1, LPE = N * P, EPW = 32 / LPE;
  constexpr int CH = (M + P - 1) / P;              // chunks of "others" per lane
  constexpr int W = (E + 31) / 32;
  constexpr bool LSA = SCN != GSM_SCN_NAVIGATION;
  static_assert(LPE <= 32 && EPW >= 1, "an env must fit in one warp");
  static_assert(W == 1 || P == 1 || true, "");
  const int K = p.K;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int eiw = lane / LPE;                       // compile-time divisor
  const int off = lane - eiw * LPE;
  const int i = off / P, sub = off % P;             // my agent, my lane in its group
  const int64_t env = ((int64_t)blockIdx.x * (kSpecThreads / 32) + warp) * EPW + eiw;
  const bool active = eiw < EPW && env < p.n_envs;
  const int env_base = eiw * LPE;
  const unsigned envmask = low_mask(LPE) << (env_base & 31);
  const unsigned grpmask = low_mask(P) << ((lane - sub) & 31);

  extern __shared__ __align__(16) unsigned char smem_spec[];
  T* s_cm = nullptr;                                // [EPW per warp][N*N] assignment costs
  if (LSA) s_cm = (T*)smem_spec + ((size_t)warp * EPW + (eiw < EPW ? eiw : 0)) * N * N;

  if (!active) return;                              // whole env groups leave together

  // ---- per-lane constants of my pairs ---------------------------------------------------
  const T size_i = p.size[i];
  const bool coll_i = p.eflag[i] & 1;
  const T mass_i = p.mass[i], accel_i = p.accel[i], maxsp_i = p.max_speed[i];
  int e_c[CH];
  T dmin_c[CH];
  int fl_c[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) {
    const int o = c * P + sub;
    const int e = o + (o >= i ? 1 : 0);
    e_c[c] = e;
    const bool valid = o < M;
    fl_c[c] = valid ? (int)p.eflag[e] : 0;
    dmin_c[c] = valid ? size_i + p.size[e] : (T)0;
    if (!valid) e_c[c] = -1;
  }

  // ---- state into registers -----------------------------------------------------------------
  typedef typename Vec4<T>::type V4;
  typedef typename Vec2<T>::type V2;
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
  // navigation target = own goal; polygon/line markers for the slots
  T m0x = 0, m0y = 0, m1x = 0, m1y = 0;
  if (L > 0) { const T* l = p.lm_pos + env * L * 2; m0x = l[0]; m0y = l[1]; if (L > 1) { m1x = l[2]; m1y = l[3]; } }
  T goalx = 0, goaly = 0;
  if (!LSA) { const T* l = p.lm_pos + (env * L + i) * 2; goalx = l[0]; goaly = l[1]; }
  T slotx = 0, sloty = 0;                            // LSA: slot `off` (lanes off < N)
  if (LSA && off < N) {
    if (SCN == GSM_SCN_POLYGON) {
      slotx = m0x + p.poly_r * p.slot_table[2 * off];
      sloty = m0y + p.poly_r * p.slot_table[2 * off + 1];
    } else {
      const T f = p.slot_table[2 * off];
      slotx = m0x + f * (m1x - m0x);
      sloty = m0y + f * (m1y - m0y);
    }
  }
  int t_now = p.t[env];

  const unsigned char* act_ptr = (const unsigned char*)p.actions;
  unsigned char* o_obs = (unsigned char*)p.obs;
  unsigned char* o_idx = (unsigned char*)p.nbr_idx;
  unsigned char* o_feat = (unsigned char*)p.nbr_feat;
  unsigned char* o_cnt = (unsigned char*)p.nbr_cnt;
  unsigned char* o_adj = (unsigned char*)p.adj;
  unsigned char* o_rew = (unsigned char*)p.reward;
  unsigned char* o_cost = (unsigned char*)p.cost;
  unsigned char* o_done = (unsigned char*)p.done;
  unsigned char* o_asg = (unsigned char*)p.assign;
  const int64_t row = env * N + i;

  for (int step = 0; step < n_steps; step++) {
    // ---- SPEC §2: action force (all P lanes of the agent, same address -> one request) ----
    T fx = 0, fy = 0;
    {
      T ux = 0, uy = 0;
      if (p.action_mode == GSM_ACT_DISCRETE) {
        const int a = ((const int32_t*)act_ptr)[row];
        if (a >= 0 && a < p.n_actions) { ux = p.discrete_u[a][0]; uy = p.discrete_u[a][1]; }
      } else {
        const T* ap = (const T*)act_ptr + row * 2;
        ux = ap[0]; uy = ap[1];
      }
      if (sub == 0) { fx = accel_i * ux; fy = accel_i * uy; }
    }
    // ---- SPEC §3: pair forces ---------------------------------------------------------------
#pragma unroll
    for (int c = 0; c < CH; c++) {
      T qx = lmx[c], qy = lmy[c];
      if (c * P < N - 1) {                           // this chunk can hold agent pairs
        const int src = env_base + (e_c[c] >= 0 && e_c[c] < N ? e_c[c] : 0) * P;
        const T ax = shfl(envmask, px, src), ay = shfl(envmask, py, src);
        if (e_c[c] >= 0 && e_c[c] < N) { qx = ax; qy = ay; }
      }
      if (coll_i && (fl_c[c] & 1)) {
        const T dx = px - qx, dy = py - qy;
        const T dist = r_sqrt(dx * dx + dy * dy);
        const T x = -(dist - dmin_c[c]) / p.km;
        if (!(Prec<T>::kCut && x < (T)(-kFarCut))) {
          const T pen = softplus(x) * p.km;
          fx = fx + p.cf * dx / dist * pen;
          fy = fy + p.cf * dy / dist * pen;
        }
      }
    }
    if (P > 1) {
#pragma unroll
      for (int m = P / 2; m >= 1; m >>= 1) {
        fx += __shfl_xor_sync(grpmask, fx, m);
        fy += __shfl_xor_sync(grpmask, fy, m);
      }
    }
    // ---- SPEC §4: integration, redundantly on the P lanes of the agent --------------------
    vx = vx * p.one_minus_damp; vy = vy * p.one_minus_damp;
    vx = vx + (fx / mass_i) * p.dt;
    vy = vy + (fy / mass_i) * p.dt;
    if (maxsp_i > (T)0) {
      const T sp = r_sqrt(vx * vx + vy * vy);
      if (sp > maxsp_i) { vx = vx / sp * maxsp_i; vy = vy / sp * maxsp_i; }
    }
    px = px + vx * p.dt; py = py + vy * p.dt;
    t_now += 1;

    // ---- SPEC §5: assignment (polygon / line) -----------------------------------------------
    T tx = goalx, ty = goaly;
    int asg = i;
    if (LSA) {
      // lane off < N is column `off`: cost of every agent row to my slot
#pragma unroll
      for (int r = 0; r < N; r++) {
        const T rx = shfl(envmask, px, env_base + r * P), ry = shfl(envmask, py, env_base + r * P);
        if (off < N) {
          const T dx = slotx - rx, dy = sloty - ry;
          s_cm[r * N + off] = r_sqrt(dx * dx + dy * dy);
        }
      }
      __syncwarp(envmask);
      const int a_row = lsa_lanes<T>(s_cm, N, off, envmask, env_base);   // col4row for row == off
      __syncwarp(envmask);
      asg = shfl(envmask, a_row, env_base + i);
      tx = shfl(envmask, slotx, env_base + asg);
      ty = shfl(envmask, sloty, env_base + asg);
    }

    // ---- SPEC §6: neighbour graph -------------------------------------------------------------
    int cnt = 0, ncol = 0;
    uint32_t words[W];
#pragma unroll
    for (int w = 0; w < W; w++) words[w] = 0;
    int32_t* g_idx = (int32_t*)o_idx + row * K;
    T* g_feat = (T*)o_feat + row * K * GSM_NBR_FEAT_DIM;
#pragma unroll
    for (int c = 0; c < CH; c++) {
      T ex = lmx[c], ey = lmy[c], evx = 0, evy = 0;
      if (c * P < N - 1) {
        const int src = env_base + (e_c[c] >= 0 && e_c[c] < N ? e_c[c] : 0) * P;
        const T ax = shfl(envmask, px, src), ay = shfl(envmask, py, src);
        const T bx = shfl(envmask, vx, src), by = shfl(envmask, vy, src);
        if (e_c[c] >= 0 && e_c[c] < N) { ex = ax; ey = ay; evx = bx; evy = by; }
      }
      const int e = e_c[c];
      const T dx = ex - px, dy = ey - py;
      const T dist = r_sqrt(dx * dx + dy * dy);
      bool nb = e >= 0 && dist < p.Rs;
      if (!LSA && p.own_goal_always && e == N + i) nb = true;
      const bool col = e >= 0 && dist < dmin_c[c] &&
                       (e < N || (p.cost_obstacles && (fl_c[c] >> 1) == GSM_ENT_OBSTACLE));
      unsigned bits, cbits;
      if (P == 1) { bits = nb ? 1u : 0u; cbits = col ? 1u : 0u; }
      else {
        const int sh = (lane - sub) & 31;
        bits = (__ballot_sync(grpmask, nb) >> sh) & low_mask(P);
        cbits = (__ballot_sync(grpmask, col) >> sh) & low_mask(P);
      }
      const int pos = cnt + __popc(bits & low_mask(sub));
      if (nb && pos < K) {
        if (o_idx) g_idx[pos] = e;
        if (o_feat) {
          T* f = g_feat + pos * GSM_NBR_FEAT_DIM;
          st2<T>(f, dx, dy); st2<T>(f + 2, evx - vx, evy - vy); st2<T>(f + 4, dist, (T)(fl_c[c] >> 1));
        }
      }
      cnt += __popc(bits);
      ncol += __popc(cbits);
#pragma unroll
      for (int w = 0; w < W; w++) {
        uint32_t cw = (nb && (e >> 5) == w) ? (1u << (e & 31)) : 0u;
        // only words this chunk can touch are reduced (compile-time range)
        if (w * 32 <= c * P + P && (w + 1) * 32 > c * P) {
          if (P > 1) cw = __reduce_or_sync(grpmask, cw);
          words[w] |= cw;
        }
      }
    }
    if (cnt > K) cnt = K;
    for (int k = cnt + sub; k < K; k += P) {          // padding rows
      if (o_idx) g_idx[k] = -1;
      if (o_feat) {
        T* f = g_feat + k * GSM_NBR_FEAT_DIM;
        st2<T>(f, (T)0, (T)0); st2<T>(f + 2, (T)0, (T)0); st2<T>(f + 4, (T)0, (T)0);
      }
    }

    // ---- SPEC §6-7: per-agent scalars ---------------------------------------------------------
    const T gx = tx - px, gy = ty - py;
    const T d = r_sqrt(gx * gx + gy * gy);
    T r = ((T)0 - p.w_dist * d) + (d < p.goal_tol ? p.w_goal : (T)0);
    if (p.share_reward) {
      T s = shfl(envmask, r, env_base);
#pragma unroll
      for (int k = 1; k < N; k++) s = s + shfl(envmask, r, env_base + k * P);
      r = s / (T)N;
    }
    if (sub == 0) {
      if (o_obs) {
        T* o = (T*)o_obs + row * GSM_OBS_DIM;
        st2<T>(o, vx, vy); st2<T>(o + 2, px, py); st2<T>(o + 4, gx, gy);
      }
      if (o_cnt) ((int32_t*)o_cnt)[row] = cnt;
      if (o_adj) {
#pragma unroll
        for (int w = 0; w < W; w++) ((uint32_t*)o_adj)[row * W + w] = words[w];
      }
      if (o_rew) ((T*)o_rew)[row] = r;
      if (o_cost) ((T*)o_cost)[row] = (T)ncol;
      if (o_done) o_done[row] = (uint8_t)(t_now >= p.episode_length);
      if (o_asg) ((int32_t*)o_asg)[row] = asg;
    }

    // next slot of the rollout buffers
    act_ptr += ss.actions;
    if (o_obs) o_obs += ss.obs;
    if (o_idx) o_idx += ss.nbr_idx;
    if (o_feat) o_feat += ss.nbr_feat;
    if (o_cnt) o_cnt += ss.nbr_cnt;
    if (o_adj) o_adj += ss.adj;
    if (o_rew) o_rew += ss.reward;
    if (o_cost) o_cost += ss.cost;
    if (o_done) o_done += ss.done;
    if (o_asg) o_asg += ss.assign;
  }

  // ---- state back to HBM ------------------------------------------------------------------------
  if (sub == 0) {
    T* a = p.agent_state + row * 4;
    st2<T>(a, px, py); st2<T>(a + 2, vx, vy);
    if (i == 0) p.t[env] = t_now;
  }
}

}  // namespace gsm
