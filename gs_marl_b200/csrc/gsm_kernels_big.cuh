// gsm_kernels_big.cuh — large-team navigation kernel (N > 12: the 24/48/96-agent zero-shot
// configurations), SPEC.md §2-4, §6-7.
//
// One CTA per env.  The env's entities live in shared memory for `n_steps` consecutive steps
// (fused rollout, like env_steps_kernel).  A warp owns one agent at a time and sweeps the
// other entities 32 per pass:
//   * force sweep only over the compacted list of colliding entities (goals never collide),
//     fp32 pre-check on the squared distance so that the sqrt / softplus path runs only for
//     pairs in or near contact (fp64 verification mode evaluates every pair exactly);
//   * neighbour sweep with a squared-distance pre-check, ballot + popc compaction of the
//     rows into a per-warp shared-memory staging block, zero padding behind them, and ONE
//     TMA bulk store (cp.async.bulk.global.shared::cta) per agent for the K x 6 feature block
//     and one for the K index block — full-line writes, no per-lane 8-byte row pieces;
//     two staging buffers per warp so the store of agent a overlaps the sweep of agent a+1;
//   * adjacency words straight from the ballots.
// Requires navigation, K % 4 == 0 (16-byte bulk-copy granularity) and every output present;
// the API layer falls back to env_kernel otherwise.
#pragma once
#include "gsm_kernels_spec.cuh"

namespace gsm {

constexpr int kBigThreads = 256;

template <typename T> struct Ent { T x, y, size; int flag; int pad_; };
template <> struct Ent<float> { float x, y, size; int flag; };

struct BigLayout {
  size_t off_ent, off_vel, off_nxt, off_agc, off_rew, off_clist, off_stage, stage_idx_bytes,
      stage_bytes, total;
};
__host__ __device__ inline BigLayout make_big_layout(int rb, int ent_bytes, int N, int E, int K, int warps) {
  BigLayout b;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += ((bytes + 127) / 128) * 128; return r; };
  b.off_ent = take((size_t)((E + 31) / 32 * 32) * ent_bytes);   // padded with far-away dummies
  b.off_vel = take((size_t)N * 2 * rb);
  b.off_nxt = take((size_t)N * 4 * rb);
  b.off_agc = take((size_t)N * 3 * rb);
  b.off_rew = take((size_t)N * rb);
  b.off_clist = take((size_t)E * 4);
  b.stage_idx_bytes = ((size_t)K * 4 + 15) / 16 * 16;
  b.stage_bytes = b.stage_idx_bytes + ((size_t)K * GSM_NBR_FEAT_DIM * rb + 15) / 16 * 16;
  b.off_stage = take((size_t)warps * 2 * b.stage_bytes);
  b.total = o;
  return b;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T>
__global__ void __launch_bounds__(kBigThreads)
env_big_kernel(const __grid_constant__ KParams<T> p, const int n_steps,
               const __grid_constant__ StepStrides ss) {
  extern __shared__ __align__(128) unsigned char sm[];
  typedef Arith<T> A;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NW = kBigThreads / 32;
  const int N = p.N, L = p.L, E = p.E, K = p.K, W = p.W;
  const BigLayout lay = make_big_layout((int)sizeof(T), (int)sizeof(Ent<T>), N, E, K, NW);
  Ent<T>* ent = (Ent<T>*)(sm + lay.off_ent);
  T* vel = (T*)(sm + lay.off_vel);
  T* nxt = (T*)(sm + lay.off_nxt);
  T* agc = (T*)(sm + lay.off_agc);
  T* rew = (T*)(sm + lay.off_rew);
  int* clist = (int*)(sm + lay.off_clist);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t env = blockIdx.x;
  unsigned char* stage0 = sm + lay.off_stage + (size_t)warp * 2 * lay.stage_bytes;

  // ---- stage the env --------------------------------------------------------------------------
  {
    const T* g_ag = p.agent_state + env * N * 4;
    const T* g_lm = p.lm_pos + env * L * 2;
    for (int e = tid; e < E; e += kBigThreads) {
      Ent<T> q;
      if (e < N) { q.x = g_ag[4 * e]; q.y = g_ag[4 * e + 1]; vel[2 * e] = g_ag[4 * e + 2]; vel[2 * e + 1] = g_ag[4 * e + 3]; }
      else { q.x = g_lm[2 * (e - N)]; q.y = g_lm[2 * (e - N) + 1]; }
      q.size = p.size[e]; q.flag = p.eflag[e];
      ent[e] = q;
    }
    for (int e = E + tid; e < (E + 31) / 32 * 32; e += kBigThreads) {
      Ent<T> q;                                        // never a neighbour, never in contact
      q.x = (T)1e30; q.y = (T)1e30; q.size = 0; q.flag = 0;
      ent[e] = q;
    }
    for (int i = tid; i < N; i += kBigThreads) {
      agc[3 * i] = p.mass[i]; agc[3 * i + 1] = p.accel[i]; agc[3 * i + 2] = p.max_speed[i];
    }
    {                                                  // staging starts as all-padding rows
      int32_t* z = (int32_t*)(sm + lay.off_stage);
      const int words = (int)(NW * 2 * lay.stage_bytes / 4), iw = (int)(lay.stage_idx_bytes / 4),
                sw = (int)(lay.stage_bytes / 4);
      for (int k = tid; k < words; k += kBigThreads) z[k] = (k % sw) < iw ? -1 : 0;
    }
    if (warp == 0) {                                  // compact list of colliding entities, ascending
      int n = 0;
      for (int e0 = 0; e0 < E; e0 += 32) {
        const int e = e0 + lane;
        const bool c = e < E && (p.eflag[e] & 1);
        const unsigned b = __ballot_sync(FULL, c);
        if (c) clist[n + __popc(b & low_mask(lane))] = e;
        n += __popc(b);
      }
    }
  }
  __syncthreads();
  int nc = 0;                                          // number of colliding entities
  T max_size = 0;
  for (int e = 0; e < E; e++) {                        // small, uniform; avoids another barrier
    nc += (ent[e].flag & 1);
    max_size = ent[e].size > max_size ? ent[e].size : max_size;
  }
  // contact distances below the sensing radius: a colliding pair is always a neighbour pair
  // too, so the fp32 sweep needs only ONE squared-distance pre-check per pair
  const bool col_in_nb = Prec<T>::kCut && ((T)2 * max_size) * (T)1.0001 < p.Rs;
  int dirty0 = 0, dirty1 = 0;                          // rows of each staging buffer holding real data
  int t_now = p.t[env];

  // pre-check thresholds are a few ulp inclusive; the decision is made on the rounded dist
  const T Rs2 = p.Rs * p.Rs * (T)1.000001;
  const T cut = (T)kFarCut * p.km;
  int64_t so_act = 0, so_obs = 0, so_idx = 0, so_feat = 0, so_cnt = 0, so_adj = 0, so_rew = 0,
          so_cost = 0, so_done = 0, so_asg = 0;       // byte offsets of the current rollout slot
  unsigned it = 0;                                     // per-warp agent-iteration counter (staging parity)

  for (int step = 0; step < n_steps; step++) {
    // ---- A. SPEC §2-4: force + integration (positions of this step are read-only in `ent`) ----
    for (int i = warp; i < N; i += NW) {
      const Ent<T> me = ent[i];
      T fx = 0, fy = 0;
      if (lane == 0) {
        T ux = 0, uy = 0;
        if (p.action_mode == GSM_ACT_DISCRETE) {
          const int a = *(const int32_t*)((const unsigned char*)p.actions + so_act + (env * N + i) * 4);
          if (a >= 0 && a < p.n_actions) { ux = p.discrete_u[a][0]; uy = p.discrete_u[a][1]; }
        } else {
          const T* ap = (const T*)((const unsigned char*)p.actions + so_act) + (env * N + i) * 2;
          ux = ap[0]; uy = ap[1];
        }
        fx = agc[3 * i + 1] * ux; fy = agc[3 * i + 1] * uy;
      }
      if (me.flag & 1) {
        for (int c0 = 0; c0 < nc; c0 += 32) {
          const int c = c0 + lane;
          if (c >= nc) continue;
          const int j = clist[c];
          if (j == i) continue;
          const Ent<T> q = ent[j];
          const T dx = me.x - q.x, dy = me.y - q.y;
          const T d2 = dx * dx + dy * dy;
          const T dmin = me.size + q.size;
          if (Prec<T>::kCut) {                         // fp32: x < -kFarCut  <=>  dist > dmin + cut
            const T far = dmin + cut;
            if (d2 > far * far) continue;
          }
          const T dist = A::sqrt(d2);
          const T x = A::div_const(-(dist - dmin), p.km, p.km_inv);
          if (Prec<T>::kCut && x < (T)(-kFarCut)) continue;
          const T pen = softplus(x) * p.km;
          fx = fx + A::div(p.cf * dx, dist) * pen;
          fy = fy + A::div(p.cf * dy, dist) * pen;
        }
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
        fx += __shfl_xor_sync(FULL, fx, m);
        fy += __shfl_xor_sync(FULL, fy, m);
      }
      if (lane == 0) {
        T vx = vel[2 * i] * p.one_minus_damp, vy = vel[2 * i + 1] * p.one_minus_damp;
        const T m = agc[3 * i];
        vx = vx + A::div(fx, m) * p.dt;
        vy = vy + A::div(fy, m) * p.dt;
        const T ms = agc[3 * i + 2];
        if (ms > (T)0) {
          const T sp = A::sqrt(vx * vx + vy * vy);
          if (sp > ms) { vx = A::div(vx, sp) * ms; vy = A::div(vy, sp) * ms; }
        }
        nxt[4 * i] = me.x + vx * p.dt; nxt[4 * i + 1] = me.y + vy * p.dt;
        nxt[4 * i + 2] = vx; nxt[4 * i + 3] = vy;
      }
    }
    __syncthreads();
    for (int k = tid; k < N; k += kBigThreads) {
      ent[k].x = nxt[4 * k]; ent[k].y = nxt[4 * k + 1];
      vel[2 * k] = nxt[4 * k + 2]; vel[2 * k + 1] = nxt[4 * k + 3];
    }
    t_now += 1;
    __syncthreads();

    // ---- C. SPEC §6-7: neighbour graph, obs, reward, cost, done --------------------------------
    for (int i = warp; i < N; i += NW, it++) {
      unsigned char* stage = stage0 + (size_t)(it & 1) * lay.stage_bytes;
      int32_t* st_idx = (int32_t*)stage;
      T* st_feat = (T*)(stage + lay.stage_idx_bytes);
      if (lane == 0) bulk_wait_read<1>();              // the store that last read this buffer is done
      __syncwarp();
      const Ent<T> me = ent[i];
      const T vx = vel[2 * i], vy = vel[2 * i + 1];
      const int64_t row = env * N + i;
      int cnt = 0, ncol = 0;
      uint32_t myword = 0;                             // lane w keeps adjacency word w
      for (int e0 = 0; e0 < E; e0 += 32) {             // sweep by ENTITY index: ballot == adjacency word
        const int e = e0 + lane;
        const Ent<T> q = ent[e];                       // padded: no bounds check
        const T dx = q.x - me.x, dy = q.y - me.y;
        const T d2 = dx * dx + dy * dy;
        const T dmin = me.size + q.size;
        const bool goal = p.own_goal_always && e == N + i;
        bool nb, col;
        T dist = 0;
        if (Prec<T>::kCut) {
          nb = (d2 < Rs2 && e != i) || goal;
          col = col_in_nb ? nb : (d2 < dmin * dmin * (T)1.000001 && e != i);
          if (__ballot_sync(FULL, nb || col) == 0u) continue;      // warp-uniform: nothing near
          if (nb || col) {                             // decide on the rounded distance, like every kernel
            dist = A::sqrt(d2);
            nb = (dist < p.Rs && e != i) || goal;
            col = dist < dmin && e != i;
          }
        } else {
          dist = A::sqrt(d2);
          nb = (dist < p.Rs && e != i && e < E) || goal;
          col = dist < dmin && e != i && e < E;
        }
        col = col && (e < N || (p.cost_obstacles && (q.flag >> 1) == GSM_ENT_OBSTACLE));
        const unsigned bits = __ballot_sync(FULL, nb);
        const unsigned cbits = __ballot_sync(FULL, col);
        if (nb) {
          const int pos = cnt + __popc(bits & low_mask(lane));
          if (pos < K) {
            st_idx[pos] = e;
            T evx = 0, evy = 0;
            if (e < N) { evx = vel[2 * e]; evy = vel[2 * e + 1]; }
            T* f = st_feat + pos * GSM_NBR_FEAT_DIM;
            f[0] = dx; f[1] = dy; f[2] = evx - vx; f[3] = evy - vy; f[4] = dist; f[5] = (T)(q.flag >> 1);
          }
        }
        cnt += __popc(bits);
        ncol += __popc(cbits);
        if (lane == (e0 >> 5)) myword = bits;
      }
      if (lane < W) ((uint32_t*)((unsigned char*)p.adj + so_adj))[row * W + lane] = myword;
      if (cnt > K) cnt = K;
      // rows behind the neighbours must read as padding: only rows a previous use of this buffer
      // filled (dirty) and this agent did not overwrite need clearing
      {
        const int hi = (it & 1) ? dirty1 : dirty0;
        for (int k = cnt + lane; k < hi; k += 32) st_idx[k] = -1;
        for (int q = cnt * GSM_NBR_FEAT_DIM + lane; q < hi * GSM_NBR_FEAT_DIM; q += 32) st_feat[q] = (T)0;
        if (it & 1) dirty1 = cnt; else dirty0 = cnt;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        bulk_store((unsigned char*)p.nbr_idx + so_idx + row * K * 4, st_idx, (uint32_t)(K * 4));
        bulk_store((unsigned char*)p.nbr_feat + so_feat + row * K * GSM_NBR_FEAT_DIM * (int64_t)sizeof(T), st_feat,
                   (uint32_t)(K * GSM_NBR_FEAT_DIM * sizeof(T)));
        bulk_commit();
      }
      // per-agent scalars
      const Ent<T> goal = ent[N + i];
      const T gx = goal.x - me.x, gy = goal.y - me.y;
      const T d = A::sqrt(gx * gx + gy * gy);
      const T r = ((T)0 - p.w_dist * d) + (d < p.goal_tol ? p.w_goal : (T)0);
      if (lane < 3) {
        T* o = (T*)((unsigned char*)p.obs + so_obs) + row * GSM_OBS_DIM + 2 * lane;
        st2<T>(o, lane == 0 ? vx : (lane == 1 ? me.x : gx), lane == 0 ? vy : (lane == 1 ? me.y : gy));
      } else if (lane == 3) {
        *((int32_t*)((unsigned char*)p.nbr_cnt + so_cnt) + row) = cnt;
      } else if (lane == 4) {
        if (p.share_reward) rew[i] = r;
        else *((T*)((unsigned char*)p.reward + so_rew) + row) = r;
      } else if (lane == 5) {
        *((T*)((unsigned char*)p.cost + so_cost) + row) = (T)ncol;
      } else if (lane == 6) {
        *((unsigned char*)p.done + so_done + row) = (uint8_t)(t_now >= p.episode_length);
      } else if (lane == 7) {
        *((int32_t*)((unsigned char*)p.assign + so_asg) + row) = i;
      }
    }
    if (p.share_reward) {
      __syncthreads();
      T s = rew[0];
      for (int k = 1; k < N; k++) s = s + rew[k];
      s = s / (T)N;
      for (int k = tid; k < N; k += kBigThreads)
        *((T*)((unsigned char*)p.reward + so_rew) + env * N + k) = s;
      __syncthreads();
    }
    so_act += ss.actions; so_obs += ss.obs; so_idx += ss.nbr_idx; so_feat += ss.nbr_feat;
    so_cnt += ss.nbr_cnt; so_adj += ss.adj; so_rew += ss.reward; so_cost += ss.cost;
    so_done += ss.done; so_asg += ss.assign;
  }

  // ---- state back to HBM; staging must outlive the bulk reads ----------------------------------
  if (lane == 0) bulk_wait_read<0>();
  {
    T* g_ag = p.agent_state + env * N * 4;
    for (int k = tid; k < N; k += kBigThreads) {
      g_ag[4 * k] = ent[k].x; g_ag[4 * k + 1] = ent[k].y;
      g_ag[4 * k + 2] = vel[2 * k]; g_ag[4 * k + 3] = vel[2 * k + 1];
    }
    if (tid == 0) p.t[env] = t_now;
  }
}

}  // namespace gsm
