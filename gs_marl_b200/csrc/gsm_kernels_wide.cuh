// gsm_kernels_wide.cuh — navigation kernel for small teams, "one lane per OTHER entity"
// (SPEC.md §2-8; reference environment.py / core.py / scenarios, SOURCES.txt:14,15,21-22 — withheld).
//
// env_steps_kernel (gsm_kernels_spec.cuh) gives an env N*P lanes: P lanes per agent, every lane
// owning one (agent, other) pair per chunk.  At navigation-3 (E = 9 entities, M = 8 others per
// agent) that is 12 lanes per env, 2 envs and 24 of 32 lanes per warp, and the per-warp overhead
// (cursors, action fetch, shuffles of the other agents' state, role stores) is paid per 2 envs.
// This kernel turns the layout by 90 degrees:
//   * an env owns GW = pow2(M) lanes; lane j is "other number j" of EVERY agent, the agents are the
//     (unrolled) chunks.  M = 8 fills a warp with 4 envs and all 32 lanes;
//   * every lane keeps the state of all N agents in registers and integrates all of them
//     redundantly: no shuffle ever fetches a position, lanes j >= N-1 see the same landmark in every
//     chunk and keep just that one;
//   * one ballot per agent covers its whole neighbourhood (all 4 envs at once), so the row of a
//     pair is known right after its own chunk — no second pass, no feature registers held;
//   * the contact force of step s+1 is a function of exactly the pair geometry the graph pass of
//     step s computes (same positions, same rounded distance: bit-identical in fp64 too), so the
//     separate force sweep disappears from the steady state: it runs on the first step of a launch
//     and after an in-kernel re-draw only;
//   * lane N-1+i holds agent i's own goal as its landmark, so its chunk-i geometry IS the goal
//     vector and the goal distance: reward and the goal half of obs cost no extra arithmetic;
//   * per-agent constants (size, mass, accel, max_speed, flags) travel in the kernel parameter
//     space and are compile-time indexed: constant-bank operands, no registers.
// Everything else — arithmetic policy, output layout, auto-reset draws — is env_steps_kernel's.
#pragma once
#include "gsm_kernels_spec.cuh"
#include "gsm_kernels_big.cuh"   // bulk-store helpers

namespace gsm {

constexpr int kWideThreads = 128;
__host__ __device__ constexpr int wide_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

template <typename T, int N, int E>
struct WideConsts {
  T size[E];
  T mass[N], mass_inv[N], accel[N], maxsp[N];
  int32_t eflag[E];
};

#ifndef GSM_WIDE_BLOCKS        // A/B: resident CTAs per SM the fp32 instance is compiled for
#define GSM_WIDE_BLOCKS 7
#endif
template <typename T> struct WideMinBlocks { static constexpr int value = sizeof(T) == 4 ? GSM_WIDE_BLOCKS : 1; };

// SPEC §3 softplus as ONE code path: max(x, 0) + log1p(exp(-|x|)).  Bit-identical to the two-branch
// form of gsm::softplus (x > 0: x + log1p(exp(-x)); else 0 + log1p(exp(x)) = log1p(exp(x)) exactly).
__device__ __forceinline__ float softplus1(float x) { return fmaxf(x, 0.f) + r_log1p(r_exp(-fabsf(x))); }
__device__ __forceinline__ double softplus1(double x) { return fmax(x, 0.0) + r_log1p(r_exp(-fabs(x))); }

template <typename T>
__device__ __forceinline__ T shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// Shared-memory staging of one warp-step's rows: EPW envs' nbr_feat, nbr_idx and obs blocks are three
// contiguous, 16-byte aligned regions of the slot (env-major layout), so the rows are written with STS
// and leave as three bulk copies (cp.async.bulk.global.shared::cta, UBLKCP) issued by one lane — no
// per-lane 24-byte global stores, no L1TEX sector inflation.  Two buffers per warp: the copies of step s
// drain while step s+1 is computed.
__host__ __device__ constexpr size_t wide_stage_bytes(int rb, int N, int K, int EPW) {
  return (size_t)EPW * N * K * (GSM_NBR_FEAT_DIM * rb + 4) + (size_t)EPW * N * GSM_OBS_DIM * rb;
}
__host__ __device__ constexpr size_t wide_smem_bytes(int rb, int N, int K, int EPW) {
  return 2 * (kWideThreads / 32) * wide_stage_bytes(rb, N, K, EPW);
}

// MODE 0: step(s).  MODE 1: observe only (reset path, optional per-env mask).  MODE 2: steps with
// in-kernel auto-reset.  Needs every output pointer non-NULL and n_envs * N * K * 6 * sizeof(T) < 2^31
// (32-bit lane offsets against warp-uniform 64-bit slot bases); the launcher checks both.
template <typename T, int N, int L, int MODE>
__global__ void __launch_bounds__(kWideThreads, WideMinBlocks<T>::value)
env_wide_kernel(const __grid_constant__ KParams<T> p, const int n_steps,
                const __grid_constant__ StepStrides ss,
                const __grid_constant__ WideConsts<T, N, N + L> wc) {
  constexpr int E = N + L, M = E - 1, GW = wide_pow2(M), EPW = 32 / GW;
  constexpr bool OBS = MODE == 1;
  constexpr unsigned FULL = 0xffffffffu;
  static_assert(GW <= 32 && E <= 32, "an env must fit in one warp, adjacency in one word");
  static_assert(L >= N && 2 * N <= GW, "navigation: goal i is landmark i; obs roles need 2N lanes");
  static_assert(N <= 8, "per-chunk flag bits");
  typedef Arith<T> A;
  extern __shared__ __align__(128) unsigned char wsm[];
  const int K = p.K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / GW, j = lane % GW, shift = grp * GW;
  const int64_t env_w0 = ((int64_t)blockIdx.x * (kWideThreads / 32) + warp) * EPW;
  const int64_t env_raw = env_w0 + grp;
  bool active = env_raw < p.n_envs;
  if (OBS && p.mask != nullptr && active) active = p.mask[env_raw * p.mask_stride] != 0;
  const int env = active ? (int)env_raw : 0;           // idle groups shadow env 0, never store
  const bool has = M == GW ? true : j < M;
  const bool is_lm = j >= N - 1;                        // my other is landmark j + 1 - N in every chunk
  const unsigned below = low_mask(j);
  const unsigned act_bits = __ballot_sync(FULL, active);
  // staged bulk path: the warp's EPW envs all exist (and are unmasked) and every slot is 16-byte aligned
  const bool bulk = act_bits == FULL &&
      ((((uintptr_t)p.obs | (uintptr_t)p.nbr_idx | (uintptr_t)p.nbr_feat | (uintptr_t)ss.obs |
         (uintptr_t)ss.nbr_idx | (uintptr_t)ss.nbr_feat) & 15) == 0);

  // ---- lane constants: one flag word (bit i: contact pair in chunk i, 8+i: counts as collision,
  //      16+i: agent i's own goal and own_goal_always), the size and type of my landmark ----------------
  unsigned cfl = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const int e = has ? j + (j >= i ? 1 : 0) : 0;
    const int fl = wc.eflag[e];
    if (has && (wc.eflag[i] & 1) && (fl & 1)) cfl |= 1u << i;
    if (has && (e < N || (p.cost_obstacles && (fl >> 1) == GSM_ENT_OBSTACLE))) cfl |= 1u << (8 + i);
    if (has && p.own_goal_always && e == N + i) cfl |= 1u << (16 + i);
  }
  const T size_lm = wc.size[has && is_lm ? j + 1 : 0];
  const T type_lm = (T)(wc.eflag[has && is_lm ? j + 1 : 0] >> 1);

  // ---- state: all N agents in every lane, one landmark per lane ---------------------------------
  T px[N], py[N], vx[N], vy[N];
  {
    const T* a = p.agent_state + (int64_t)env * N * 4;
#pragma unroll
    for (int i = 0; i < N; i++) { px[i] = a[4 * i]; py[i] = a[4 * i + 1]; vx[i] = a[4 * i + 2]; vy[i] = a[4 * i + 3]; }
  }
  T lmx = 0, lmy = 0;
  if (is_lm && has) {
    const T* l = p.lm_pos + ((int64_t)env * L + (j + 1 - N)) * 2;
    lmx = l[0]; lmy = l[1];
  }
  int t_now = p.t[env];
  const bool auto_reset = MODE == 2 && p.auto_reset != 0;
  int ep = auto_reset ? p.episode[env] : 0;
  const int ep0 = ep;

  // butterfly all-reduce over the env's GW lanes; IEEE addition commutes, so every lane of the
  // group ends with the same bits and the redundant integrations stay identical
  auto reduce = [&](T (&fx)[N], T (&fy)[N]) {
#pragma unroll
    for (int m = GW / 2; m >= 1; m >>= 1) {
#pragma unroll
      for (int i = 0; i < N; i++) { fx[i] = fx[i] + shfl_xor(fx[i], m); fy[i] = fy[i] + shfl_xor(fy[i], m); }
    }
  };

  // ---- warp-uniform slot bases, 32-bit lane offsets ----------------------------------------------
  const unsigned char* b_act = (const unsigned char*)p.actions;
  unsigned char* b_obs = (unsigned char*)p.obs;
  unsigned char* b_idx = (unsigned char*)p.nbr_idx;
  unsigned char* b_feat = (unsigned char*)p.nbr_feat;
  unsigned char* b_cnt = (unsigned char*)p.nbr_cnt;
  unsigned char* b_adj = (unsigned char*)p.adj;
  unsigned char* b_rew = (unsigned char*)p.reward;
  unsigned char* b_cost = (unsigned char*)p.cost;
  unsigned char* b_done = (unsigned char*)p.done;
  unsigned char* b_asg = (unsigned char*)p.assign;
  const unsigned row0 = (unsigned)env * N;              // first agent row of my env
  const int ja = j % N, jp = j / N;                     // roles behind the chunk loop: agent, part
  // staging: [feat EPW*N*K rows][idx EPW*N*K][obs EPW*N rows], two buffers per warp
  const unsigned sz_feat = (unsigned)(EPW * N * K * GSM_NBR_FEAT_DIM * sizeof(T)), sz_idx = (unsigned)(EPW * N * K * 4),
                 sz_obs = (unsigned)(EPW * N * GSM_OBS_DIM * sizeof(T)), sz_buf = sz_feat + sz_idx + sz_obs;
  unsigned char* const stage0 = wsm + (size_t)warp * 2 * sz_buf;
  const unsigned lrow0 = (unsigned)grp * N * K;         // my env's first neighbour row inside the warp's block
  const unsigned w_feat = (unsigned)env_w0 * (unsigned)(N * K * GSM_NBR_FEAT_DIM * sizeof(T)),   // warp offsets in a slot
                 w_idx = (unsigned)env_w0 * (unsigned)(N * K * 4),
                 w_obs = (unsigned)env_w0 * (unsigned)(N * GSM_OBS_DIM * sizeof(T));

  // action of agent j on lane j < N, prefetched one step ahead; that lane also turns it into the
  // control force accel[j] * u so that the N agents' forces cost one lookup per lane, not N
  const bool discrete = p.action_mode == GSM_ACT_DISCRETE;
  const unsigned act_off = discrete ? (row0 + (j < N ? j : 0)) * 4u : (row0 + (j < N ? j : 0)) * 2u * (unsigned)sizeof(T);
  const T accel_m = wc.accel[j < N ? j : 0];
  int act_next = 0;
  T actx_next = 0, acty_next = 0;
  if (!OBS) {
    if (discrete) act_next = *(const int32_t*)(b_act + act_off);
    else { actx_next = ((const T*)(b_act + act_off))[0]; acty_next = ((const T*)(b_act + act_off))[1]; }
  }

  T fcx[N], fcy[N];                                      // contact force on agent i for the NEXT integration
#pragma unroll
  for (int i = 0; i < N; i++) { fcx[i] = 0; fcy[i] = 0; }
  // `pro` (warp-uniform): a prologue pass — geometry and contact forces of the CURRENT positions only,
  // no integration, no outputs, no step consumed.  The first pass of a launch and the pass after an
  // in-kernel re-draw are prologues; every other pass is one env step.  One copy of the code serves both.
  bool pro = !OBS;
  int step = 0;
  while (step < n_steps) {
    if (!OBS && !pro) {
      // ---- SPEC §2 + §4 for all N agents, redundantly on every lane ---------------------------------
      T ux = actx_next, uy = acty_next;
      if (discrete) {
        const int a = act_next;
        ux = 0; uy = 0;
        if (a >= 0 && a < p.n_actions) { ux = p.discrete_u[a][0]; uy = p.discrete_u[a][1]; }
      }
      const T cfx = accel_m * ux, cfy = accel_m * uy;
      if (step + 1 < n_steps) b_act += ss.actions;      // the last step re-reads its own, valid, address
      if (discrete) act_next = *(const int32_t*)(b_act + act_off);
      else { actx_next = ((const T*)(b_act + act_off))[0]; acty_next = ((const T*)(b_act + act_off))[1]; }
#pragma unroll
      for (int i = 0; i < N; i++) {
        const T fx = shfl(FULL, cfx, shift + i) + fcx[i], fy = shfl(FULL, cfy, shift + i) + fcy[i];
        T nvx = vx[i] * p.one_minus_damp, nvy = vy[i] * p.one_minus_damp;
        nvx = nvx + A::div_const(fx, wc.mass[i], wc.mass_inv[i]) * p.dt;
        nvy = nvy + A::div_const(fy, wc.mass[i], wc.mass_inv[i]) * p.dt;
        if (wc.maxsp[i] > (T)0) {
          const T sp = A::sqrt(nvx * nvx + nvy * nvy);
          if (sp > wc.maxsp[i]) { nvx = A::div(nvx, sp) * wc.maxsp[i]; nvy = A::div(nvy, sp) * wc.maxsp[i]; }
        }
        vx[i] = nvx; vy[i] = nvy;
        px[i] = px[i] + nvx * p.dt; py[i] = py[i] + nvy * p.dt;
      }
      t_now += 1;
    }
    const bool st_on = !pro;                             // warp-uniform: this pass produces outputs
    unsigned char* const sb = stage0 + (step & 1) * sz_buf;
    // the bulk copies that read this buffer two steps ago must have finished reading it; the first
    // ballot below orders every lane's STS behind lane 0's wait
    if (st_on && bulk && lane == 0) bulk_wait_read<1>();

    // ---- SPEC §6-7: one chunk per agent; the pair's row is known after the chunk's own ballot -----
    T gxs = 0, gys = 0, rs = 0;                          // lanes N-1 .. 2N-2: goal vector / reward of agent j-(N-1)
    int cnt_m = 0, ncol_m = 0;                           // lane roles behind the loop: values of agent ja
    unsigned adj_m = 0;
    bool anyc = false;
#pragma unroll
    for (int i = 0; i < N; i++) {
      // the other entity of chunk i: an agent out of my own registers (lanes j < N-1) or my landmark
      T ex = lmx, ey = lmy, evx = 0, evy = 0, size_o = size_lm, type_o = type_lm;
#pragma unroll
      for (int a = 0; a < N - 1; a++) {
        const int oa = a + (a >= i ? 1 : 0);
        if (j == a) { ex = px[oa]; ey = py[oa]; evx = vx[oa]; evy = vy[oa]; size_o = wc.size[oa]; type_o = (T)(wc.eflag[oa] >> 1); }
      }
      const int e = j + (j >= i ? 1 : 0);
      const T dmin = wc.size[i] + size_o;
      const T dx = ex - px[i], dy = ey - py[i];
      const T dist = A::sqrt(dx * dx + dy * dy);
      const bool nb = (has && dist < p.Rs) || ((cfl >> (16 + i)) & 1u);
      const bool col = ((cfl >> (8 + i)) & 1u) && dist < dmin;
      const unsigned bits = (__ballot_sync(FULL, nb) >> shift) & low_mask(GW);
      const unsigned cbits = (__ballot_sync(FULL, col) >> shift) & low_mask(GW);
      int cnt = __popc(bits);
      const int rank = __popc(bits & below);
      const int pos = nb ? rank : cnt + (j - rank);
      if (has && pos < K && st_on) {
        const unsigned lr = lrow0 + (unsigned)(i * K + pos);
        *(int32_t*)(sb + sz_feat + lr * 4u) = nb ? e : -1;
        T* f = (T*)(sb + lr * (unsigned)(GSM_NBR_FEAT_DIM * sizeof(T)));
        const T z = (T)0;
        st2<T>(f, nb ? dx : z, nb ? dy : z);
        st2<T>(f + 2, nb ? evx - vx[i] : z, nb ? evy - vy[i] : z);
        st2<T>(f + 4, nb ? dist : z, nb ? type_o : z);
      }
      if (cnt > K) cnt = K;
      // entity-indexed adjacency: open a zero bit at the agent's own index i
      const unsigned ebits = (bits & low_mask(i)) | ((bits & ~low_mask(i)) << 1);
      if (ja == i) { cnt_m = cnt; adj_m = ebits; ncol_m = __popc(cbits); }
      if (j == N - 1 + i) {                               // my landmark is this agent's goal
        gxs = dx; gys = dy;
        rs = ((T)0 - p.w_dist * dist) + (dist < p.goal_tol ? p.w_goal : (T)0);
      }
      if (!OBS) {                                         // SPEC §3 for the next step, on the same geometry:
        fcx[i] = 0; fcy[i] = 0;                           // the force uses p_i - p_e = -d, exactly
        if ((cfl >> i) & 1u) {
          const T x = A::div_const(-(dist - dmin), p.km, p.km_inv);
          if (!(Prec<T>::kCut && x < (T)(-kFarCut))) {
            const T pen = softplus1(x) * p.km;
            fcx[i] = A::div(p.cf * (-dx), dist) * pen;
            fcy[i] = A::div(p.cf * (-dy), dist) * pen;
            anyc = true;
          }
        }
      }
    }
    // all-zero partial forces sum to exactly zero: the butterfly is skipped when no lane of the warp has a term
    if (!OBS && (pro || step + 1 < n_steps) && __any_sync(FULL, anyc)) reduce(fcx, fcy);

    if (st_on) {
      // ---- per-agent outputs by lane roles ---------------------------------------------------------
      if (p.share_reward && !OBS) {
        T s = shfl(FULL, rs, shift + N - 1);
#pragma unroll
        for (int k = 1; k < N; k++) s = s + shfl(FULL, rs, shift + N - 1 + k);
        rs = s / (T)N;
      }
      // obs: lanes 0..N-1 stage (vx, vy) of agent j, lanes N..2N-1 (px, py) of agent j-N, lanes N-1..2N-2 the goal part
      T* const so = (T*)(sb + sz_feat + sz_idx) + (unsigned)grp * (N * GSM_OBS_DIM);
      if (j < 2 * N) {
        T a0 = 0, a1 = 0;
#pragma unroll
        for (int k = 0; k < N; k++)
          if (ja == k) { a0 = jp == 0 ? vx[k] : px[k]; a1 = jp == 0 ? vy[k] : py[k]; }
        st2<T>(so + ja * GSM_OBS_DIM + 2 * jp, a0, a1);
      }
      if (j >= N - 1 && j < 2 * N - 1) {
        st2<T>(so + (j - (N - 1)) * GSM_OBS_DIM + 4, gxs, gys);
        if (!OBS && active) *(T*)(b_rew + (row0 + (j - (N - 1))) * (unsigned)sizeof(T)) = rs;
      }
      // lanes 0..N-1: cnt, cost, done of agent j; lanes N..2N-1: adj, assign of agent j-N
      if (j < 2 * N && active) {
        const unsigned r = row0 + ja;
        if (jp == 0) {
          *(int32_t*)(b_cnt + r * 4u) = cnt_m;
          if (!OBS) {
            *(T*)(b_cost + r * (unsigned)sizeof(T)) = (T)ncol_m;
            b_done[r] = (uint8_t)(t_now >= p.episode_length);
          }
        } else {
          *(uint32_t*)(b_adj + r * 4u) = adj_m;
          *(int32_t*)(b_asg + r * 4u) = ja;
        }
      }
      // ---- the staged rows leave: three bulk copies, or (ragged / masked / unaligned warps) plain stores ----
      if (bulk) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          bulk_store(b_feat + w_feat, sb, sz_feat);
          bulk_store(b_idx + w_idx, sb + sz_feat, sz_idx);
          bulk_store(b_obs + w_obs, sb + sz_feat + sz_idx, sz_obs);
          bulk_commit();
        }
      } else {
        __syncwarp();
        const unsigned wf = sz_feat / (4 * EPW), wi = sz_idx / (4 * EPW), wo = sz_obs / (4 * EPW);   // words per env
        for (unsigned q = lane; q < sz_feat / 4; q += 32)
          if ((act_bits >> ((q / wf) * GW)) & 1u) ((uint32_t*)(b_feat + w_feat))[q] = ((const uint32_t*)sb)[q];
        for (unsigned q = lane; q < sz_idx / 4; q += 32)
          if ((act_bits >> ((q / wi) * GW)) & 1u) ((uint32_t*)(b_idx + w_idx))[q] = ((const uint32_t*)(sb + sz_feat))[q];
        for (unsigned q = lane; q < sz_obs / 4; q += 32)
          if ((act_bits >> ((q / wo) * GW)) & 1u) ((uint32_t*)(b_obs + w_obs))[q] = ((const uint32_t*)(sb + sz_feat + sz_idx))[q];
        __syncwarp();
      }
    }
    if (OBS) break;
    if (pro) { pro = false; continue; }                  // the prologue consumed no step

    // ---- episode end inside a fused rollout: re-draw (terminal outputs stay in slot s) -----------------
    if (MODE == 2 && auto_reset) {
      const bool fin = t_now >= p.episode_length;
      if (__any_sync(FULL, fin)) {
        if (fin) {
          const uint64_t genv = (uint64_t)(p.env_offset + env);
#pragma unroll
          for (int i = 0; i < N; i++) {
            spawn_draw<T>(genv, ep, i, p.seed, p.ext[GSM_ENT_AGENT], px[i], py[i]);
            vx[i] = 0; vy[i] = 0;
          }
          if (is_lm && has) spawn_draw<T>(genv, ep, j + 1, p.seed, p.ext[wc.eflag[j + 1] >> 1], lmx, lmy);
          t_now = 0;
          ep += 1;
        }
        pro = step + 1 < n_steps;                         // the new positions need their forces
      }
    }
    b_obs += ss.obs; b_idx += ss.nbr_idx; b_feat += ss.nbr_feat; b_cnt += ss.nbr_cnt; b_adj += ss.adj;
    b_rew += ss.reward; b_cost += ss.cost; b_done += ss.done; b_asg += ss.assign;
    step++;
  }
  if (bulk && lane == 0) bulk_wait_read<0>();            // shared memory must outlive the copies that read it

  // ---- state back to HBM ------------------------------------------------------------------------------
  if (!OBS && active) {
    if (j == 0) {
      T* a = p.agent_state + (int64_t)env * N * 4;
#pragma unroll
      for (int i = 0; i < N; i++) { st2<T>(a + 4 * i, px[i], py[i]); st2<T>(a + 4 * i + 2, vx[i], vy[i]); }
      p.t[env] = t_now;
      if (auto_reset && ep != ep0) p.episode[env] = ep;
    }
    if (auto_reset && ep != ep0 && is_lm && has) {
      T* l = p.lm_pos + ((int64_t)env * L + (j + 1 - N)) * 2;
      l[0] = lmx; l[1] = lmy;
    }
  }
}

}  // namespace gsm
