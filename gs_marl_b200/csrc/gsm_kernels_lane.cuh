// gsm_kernels_lane.cuh — lane-per-agent navigation kernel for medium and large teams
// (N >= 12 when no better-fitting instance exists), SPEC.md §2-4, §6-7.
//
// Issue-efficient mapping for teams whose pair count is too large for one (agent, other)
// pair per lane: ONE LANE OWNS ONE AGENT and sweeps the other entities sequentially from a
// shared-memory entity table {x, y, size, flags} (16-byte broadcast loads).  Every issued
// instruction works for up to 32 agents, there is no per-chunk ballot/popc/branch overhead,
// and the pair loops carry no dependency except two accumulators, so they pipeline.
//   N <= 32 : a warp holds floor(32/N) whole envs, phases are separated by __syncwarp;
//   N  > 32 : one CTA per env, ceil(N/32) warps, phases separated by __syncthreads.
// The env state stays in shared memory / registers for `n_steps` fused steps.  fp32 uses a
// squared-distance pre-check so sqrt/softplus run only for pairs in or near contact or within
// sensing range; fp64 evaluates every pair exactly as SPEC.md writes it.  A found neighbour
// row goes straight to HBM over a warp-cooperative, fully coalesced clear of the feature blocks
// (fp32: one 16-byte and one 8-byte store per row); the rows' entity indices are collected in a
// shared-memory list per agent and nbr_idx leaves as one coalesced store per agent.
// Round 2 (profiles/README.md, diagnostic builds GSM_LANE_DIAG): the scattered row stores — one L1
// pass per lane and store instruction — were 25 % of the step, instruction count was not the limit
// (a packed-arithmetic candidate sweep over a pair table, 11 % fewer instructions, changed nothing
// in time; it stays because it is no slower and frees issue slots).
#pragma once
#include "gsm_kernels_spec.cuh"

#ifndef GSM_LANE_ROW16
#define GSM_LANE_ROW16 1     // A/B: fp32 rows as a 16-byte + an 8-byte store (0: three 8-byte stores)
#endif
#ifndef GSM_LANE_DIAG
#define GSM_LANE_DIAG 0      // diagnostic builds only: 1 no clear, 2 no row stores, 4 no adj / obs stores
#endif

namespace gsm {

template <typename T> struct LaneEnt { T x, y, size; int flag; int pad_; };
template <> struct LaneEnt<float> { float x, y, size; int flag; };

struct LaneGeom {        // host-chosen launch geometry
  int envs_per_warp;     // N <= 32: floor(32 / N); else 0
  int warps_per_env;     // N > 32: ceil(N / 32); else 0
  int warps_per_cta;
  int envs_per_cta;
};
__host__ __device__ inline LaneGeom lane_geom(int N) {
  LaneGeom g;
  if (N <= 32) { g.envs_per_warp = 32 / N; g.warps_per_env = 0; g.warps_per_cta = 4; g.envs_per_cta = g.envs_per_warp * 4; }
  else { g.envs_per_warp = 0; g.warps_per_env = (N + 31) / 32; g.warps_per_cta = g.warps_per_env; g.envs_per_cta = 1; }
  return g;
}
// Entity tables are padded to a multiple of 8 plus 8 far-away dummies (index lane_epad(E)):
// the pair sweeps run in fully unrolled groups of 8 without bounds checks.
__host__ __device__ inline int lane_epad(int E) { return (E + 7) / 8 * 8; }
// Pair table (fp32 sweeps): entity positions again, two entities per 16 bytes {x0, x1, y0, y1}, so that the
// candidate sweep loads two entities per LDS.128 and tests them with packed fp32 arithmetic (FADD2 / FMUL2 / FFMA2).
__host__ __device__ inline size_t lane_pair_bytes(int rb, int E) { return rb == 4 ? (size_t)(lane_epad(E) + 8) * 8 : 0; }
__host__ __device__ inline size_t lane_env_bytes(int ent_bytes, int rb, int N, int E) {
  return ((size_t)(lane_epad(E) + 8) * ent_bytes + 15) / 16 * 16 + lane_pair_bytes(rb, E) + ((size_t)N * 2 * rb + 15) / 16 * 16 +
         ((size_t)N * rb + 15) / 16 * 16;
}
// Neighbour list of a lane's agent (entity indices in row order, 16 bits each): the rows' nbr_idx block leaves
// from it as one coalesced store per agent.  Odd word stride: lanes writing entry r of their lists hit 32 banks.
__host__ __device__ inline int lane_list_words(int K) { return ((K + 1) / 2) | 1; }
__host__ __device__ inline size_t lane_smem(int ent_bytes, int rb, int N, int E, int K, int envs_per_cta, int warps_per_cta) {
  // per env: entity table, agent velocities, shared-reward scratch; per CTA: collider list
  // (padded to a multiple of 8) and the per-word masks of entities that count for the cost
  // ... and 16 bytes of launch constants found while staging (collider count, largest size)
  return lane_env_bytes(ent_bytes, rb, N, E) * envs_per_cta + ((size_t)(lane_epad(E) + 8) * 4 + 15) / 16 * 16 +
         ((size_t)((E + 31) / 32) * 4 + 15) / 16 * 16 + 16 + (size_t)warps_per_cta * 32 * lane_list_words(K) * 4;
}

// Candidate mask of 8 consecutive entities (4 entries of the pair table) on the squared distance.
template <typename T> struct PairSweep {
  static constexpr bool kOn = false;
  static __device__ __forceinline__ uint32_t mask8(const float4*, T, T, T) { return 0; }
  static __device__ __forceinline__ void put(void*, int, T, T) {}
};
template <> struct PairSweep<float> {
  static constexpr bool kOn = true;
  static __device__ __forceinline__ uint32_t mask8(const float4* pp, float npx, float npy, float Rs2) {
    const float2 nx = make_float2(npx, npx), ny = make_float2(npy, npy);
    uint32_t m8 = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float4 q = pp[k];
      const float2 dx = __fadd2_rn(make_float2(q.x, q.y), nx), dy = __fadd2_rn(make_float2(q.z, q.w), ny);
      const float2 d2 = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
      if (d2.x < Rs2) m8 |= 1u << (2 * k);
      if (d2.y < Rs2) m8 |= 2u << (2 * k);
    }
    return m8;
  }
  static __device__ __forceinline__ void put(void* pp, int e, float x, float y) {
    float* f = (float*)pp + 4 * (e >> 1) + (e & 1);
    f[0] = x; f[2] = y;
  }
};

// One neighbour row.  Scattered stores cost the L1 one pass per lane whatever their size (diagnostic builds,
// profiles/README.md: the rows were 25 % of the step), so fp32 rows go out in two stores instead of three.
template <typename T>
__device__ __forceinline__ void row_store(T* f, bool al16, int odd, T a, T b, T c, T d, T e, T g) {
  st2<T>(f, a, b); st2<T>(f + 2, c, d); st2<T>(f + 4, e, g);
}
template <>
__device__ __forceinline__ void row_store<float>(float* f, bool al16, int odd, float a, float b, float c, float d, float e, float g) {
  if (!al16) { st2<float>(f, a, b); st2<float>(f + 2, c, d); st2<float>(f + 4, e, g); }
  else if (odd) { st2<float>(f, a, b); *reinterpret_cast<float4*>(f + 2) = make_float4(c, d, e, g); }
  else { *reinterpret_cast<float4*>(f) = make_float4(a, b, c, d); st2<float>(f + 4, e, g); }
}
template <typename T> __device__ __forceinline__ void st_zero16(void* p) {
  *reinterpret_cast<int4*>(p) = make_int4(0, 0, 0, 0);
}

// fp32: at most 72 registers (7 CTAs of 4 warps per SM), the allocation the step loops need;
// the cold auto-reset path must not inflate it.
// AUTO: compiled-in auto-reset (fused rollouts with gsm_set_auto_reset); the plain variant
// carries none of that code.
#ifndef GSM_LANE_BLOCKS        // A/B: resident CTAs per SM the fp32 instances are compiled for
#define GSM_LANE_BLOCKS 7
#endif
#ifndef GSM_LANE_BLOCKS_CARRY  // ... of the CARRY instances (large teams; 6 -> 80 registers)
#define GSM_LANE_BLOCKS_CARRY 6
#endif
template <typename T, bool CARRY> struct LaneMinBlocks {
  static constexpr int value = sizeof(T) != 4 ? 1 : (CARRY ? GSM_LANE_BLOCKS_CARRY : GSM_LANE_BLOCKS);
};
// CARRY (fp32, chosen by the launcher for N >= GSM_LANE_CARRY_MIN_N = 48): the graph phase of step s also sums the
// contact force of step s+1 from the same table and rounded distances, so the collider sweep runs only on the
// first step of a launch and after a re-draw.  Round 1 rejected it for every N under the 72-register cap
// (profiles/rejected/lane_contact_force_carry.diff); as its own instance with 80 registers it pays where the
// sweeps dominate: nav-48 51.8 -> 49.6 us/step, nav-96 131.9 -> 119.2 (4 streams; one launch 57.3 -> 56.4,
// 143.2 -> 129.5), and costs small teams (nav-24 one launch 26.1 -> 33.3, nav-12 24.9 -> 32.8): not used there.
template <typename T, bool AUTO, bool CARRY>
__global__ void __launch_bounds__(128, LaneMinBlocks<T, CARRY>::value)
env_lane_kernel(const __grid_constant__ KParams<T> p, const int n_steps,
                const __grid_constant__ StepStrides ss) {
  extern __shared__ __align__(128) unsigned char sm[];
  typedef Arith<T> A;
  typedef LaneEnt<T> EntT;
  const int N = p.N, L = p.L, E = p.E, K = p.K, W = p.W;
  const LaneGeom g = lane_geom(N);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool multi = N > 32;                           // env spans the whole CTA

  // ---- which env / agent am I ---------------------------------------------------------------
  int env_l, i;                                        // env slot in the CTA, my agent
  if (multi) { env_l = 0; i = tid; }
  else { const int eiw = lane / N; env_l = warp * g.envs_per_warp + eiw; i = lane - eiw * N; }
  const int64_t env_raw = (int64_t)blockIdx.x * g.envs_per_cta + env_l;
  const bool has_env = env_raw < p.n_envs && (multi || lane < g.envs_per_warp * N);
  const bool active = has_env && i < N;                // lanes without an agent idle (never store)
  const int64_t env = has_env ? env_raw : 0;

  const int EP = lane_epad(E);                         // padded entity count; dummy index = EP
  typedef PairSweep<T> PS;
  const size_t ent_b = ((size_t)(EP + 8) * sizeof(EntT) + 15) / 16 * 16, pair_b = lane_pair_bytes((int)sizeof(T), E),
               vel_b = ((size_t)N * 2 * sizeof(T) + 15) / 16 * 16, rew_b = ((size_t)N * sizeof(T) + 15) / 16 * 16;
  unsigned char* base = sm + (size_t)(has_env || multi ? env_l : 0) * (ent_b + pair_b + vel_b + rew_b);
  EntT* ent = (EntT*)base;
  float4* pairs = (float4*)(base + ent_b);             // fp32 only (pair_b = 0 otherwise)
  T* vel = (T*)(base + ent_b + pair_b);
  T* rew = (T*)(base + ent_b + pair_b + vel_b);
  int* clist = (int*)(sm + (size_t)g.envs_per_cta * (ent_b + pair_b + vel_b + rew_b));
  uint32_t* costmask = (uint32_t*)((unsigned char*)clist + ((size_t)(EP + 8) * 4 + 15) / 16 * 16);
  uint32_t* meta = (uint32_t*)((unsigned char*)costmask + ((size_t)((E + 31) / 32) * 4 + 15) / 16 * 16);   // [0] colliders, [2..] largest size
  const int lw = lane_list_words(K);
  uint16_t* wlist = (uint16_t*)(meta + 4) + (size_t)warp * 32 * lw * 2;   // this warp's 32 neighbour lists
  uint16_t* elist = wlist + (size_t)lane * lw * 2;                        // mine
  const unsigned list_sa = (unsigned)__cvta_generic_to_shared(wlist);

  // ---- stage: entity tables (each env by its own lanes / CTA), collider list ------------------
  {
    const T* g_ag = p.agent_state + env * N * 4;
    const T* g_lm = p.lm_pos + env * L * 2;
    const int nl = multi ? blockDim.x : N, me = multi ? tid : i;     // lanes cooperating on this env
    if (has_env || multi) {
      for (int e = me; e < EP + 8; e += nl) {
        EntT q;
        if (e < N) { q.x = g_ag[4 * e]; q.y = g_ag[4 * e + 1]; vel[2 * e] = g_ag[4 * e + 2]; vel[2 * e + 1] = g_ag[4 * e + 3]; }
        else if (e < E) { q.x = g_lm[2 * (e - N)]; q.y = g_lm[2 * (e - N) + 1]; }
        else { q.x = (T)1e18; q.y = (T)1e18; }         // dummy: never near anything
        q.size = e < E ? p.size[e] : (T)0; q.flag = e < E ? (int)p.eflag[e] : 0;
        ent[e] = q;
        PS::put(pairs, e, q.x, q.y);
      }
    }
    if (warp == 0) {
      int n = 0;
      T ms = 0;
      for (int e0 = 0; e0 < E; e0 += 32) {
        const int e = e0 + lane;
        const int fl = e < E ? (int)p.eflag[e] : 0;
        if (e < E) { const T sz = p.size[e]; ms = sz > ms ? sz : ms; }
        const bool c = fl & 1;
        const unsigned b = __ballot_sync(0xffffffffu, c);
        if (c) clist[n + __popc(b & low_mask(lane))] = e;
        n += __popc(b);
        // entities whose overlap counts into the cost: agents, and obstacles if configured
        const bool cc = e < E && (e < N || (p.cost_obstacles && (fl >> 1) == GSM_ENT_OBSTACLE));
        const unsigned cb = __ballot_sync(0xffffffffu, cc);
        if (lane == 0) costmask[e0 >> 5] = cb;
      }
      for (int k = n + lane; k < (n + 7) / 8 * 8; k += 32) clist[k] = EP;   // pad with the dummy
#pragma unroll
      for (int o = 16; o; o >>= 1) { const T v = __shfl_xor_sync(0xffffffffu, ms, o); ms = v > ms ? v : ms; }
      if (lane == 0) { meta[0] = (uint32_t)n; *(T*)(meta + 2) = ms; }
    }
  }
  __syncthreads();
  const int nc = (int)meta[0];
  const T max_size = *(const T*)(meta + 2);
  // every contact distance below the sensing radius: a colliding pair is a neighbour candidate
  // anyway, so the candidate sweep needs one squared-distance test per pair instead of two
  const bool col_in_nb = ((T)2 * max_size) * (T)1.0001 < p.Rs;
  // fp32, and every pair that can contribute to the contact force (dist <= size_i + size_j + cut)
  // lies inside the sensing radius: the graph phase of step s already takes the rounded distance of
  // exactly those pairs on the table that step s+1's forces are computed from, so it also sums the
  // contact force of step s+1 ("carry") and the collider sweep (nc of the E + nc pair tests per
  // step) runs only on the first step of a launch and after a re-draw.  Summation order changes
  // (SPEC §9 production deviation 4); fp64 keeps the exact two-sweep sequence.
  const bool carry_ok = CARRY && Prec<T>::kCut && ((T)2 * max_size + (T)kFarCut * p.km) * (T)1.0001 < p.Rs;
  T cfx = 0, cfy = 0;
  bool have_carry = false;

  const int ii = active ? i : 0;
  T px = ent[ii].x, py = ent[ii].y, vx = vel[2 * ii], vy = vel[2 * ii + 1];
  const T size_i = ent[ii].size;
  const bool coll_i = ent[ii].flag & 1;
  const T mass_i = p.mass[ii], accel_i = p.accel[ii], maxsp_i = p.max_speed[ii];
  int t_now = p.t[env];
  const bool auto_reset = AUTO && p.auto_reset != 0;   // fused rollouts only (SPEC §8 draws)
  int ep = auto_reset ? p.episode[env] : 0;
  const int ep0 = ep;
  const uint64_t genv = (uint64_t)(p.env_offset + env);
  T gxl = ent[N + ii].x, gyl = ent[N + ii].y;          // own goal (static within an episode)
  const T Rs2 = p.Rs * p.Rs * (T)1.000001;             // pre-checks a few ulp inclusive
  const T cut = (T)kFarCut * p.km;
  const int64_t row = env * N + ii;

  // output cursors of my agent
  const unsigned char* c_act = (const unsigned char*)p.actions +
      (p.action_mode == GSM_ACT_DISCRETE ? row * 4 : row * 2 * (int64_t)sizeof(T));
  unsigned char* c_feat = (unsigned char*)(p.nbr_feat + row * K * GSM_NBR_FEAT_DIM);
  unsigned char* c_obs = (unsigned char*)(p.obs + row * GSM_OBS_DIM);
  unsigned char* c_cnt = (unsigned char*)(p.nbr_cnt + row);
  unsigned char* c_adj = (unsigned char*)(p.adj + row * W);
  unsigned char* c_rew = (unsigned char*)(p.reward + row);
  unsigned char* c_cost = (unsigned char*)(p.cost + row);
  unsigned char* c_done = (unsigned char*)(p.done + row);
  unsigned char* c_asg = (unsigned char*)(p.assign + row);
  // fp32 rows (24 bytes) leave as a 16-byte and an 8-byte store (which first depends on the row's parity) when
  // my agent's block is 16-byte aligned in every slot.  Measured (one launch, us per step, two stores / three):
  // nav-6 15.3 / 18.1, nav-12 21.8 / 24.4, nav-24 24.1 / 24.5, but nav-48 61.8 / 56.7 and nav-96 132.3 / 122.9 —
  // envs that span several warps keep the three 8-byte stores.
  const bool feat16 = GSM_LANE_ROW16 && !CARRY && !multi && sizeof(T) == 4 &&   // (!CARRY: compiled out of the large-team instances)
      ((((uintptr_t)c_feat) | (uintptr_t)ss.nbr_feat) & 15) == 0;
  const int idx_a0 = lane / K, idx_k0 = lane - idx_a0 * K, idx_da = 32 / K, idx_dk = 32 - idx_da * K;
  unsigned char* c_idx_base = (unsigned char*)p.nbr_idx;   // slot bases for the warp-wide clears
  unsigned char* c_feat_base = (unsigned char*)p.nbr_feat;

  for (int step = 0; step < n_steps; step++) {
    // ---- SPEC §2-3: forces (entity table = state at the start of the step) --------------------
    T fx = 0, fy = 0;
    if (active) {
      T ux = 0, uy = 0;
      if (p.action_mode == GSM_ACT_DISCRETE) {
        const int a = *(const int32_t*)c_act;
        if (a >= 0 && a < p.n_actions) { ux = p.discrete_u[a][0]; uy = p.discrete_u[a][1]; }
      } else { ux = ((const T*)c_act)[0]; uy = ((const T*)c_act)[1]; }
      fx = accel_i * ux; fy = accel_i * uy;
      if (coll_i && have_carry) { fx = fx + cfx; fy = fy + cfy; }
      else if (coll_i) {
        // per 32-collider block: a branch-free, sqrt-free sweep (unrolled groups of 8 over the padded
        // list) marks the pairs that can be in contact; only those, in ascending order like SPEC §3,
        // go through sqrt / softplus.  fp64 marks every pair.
        for (int c0 = 0; c0 < nc; c0 += 32) {
          uint32_t cm = 0;
          for (int g0 = 0; g0 < 32 && c0 + g0 < nc; g0 += 8) {
            uint32_t m8 = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
              const int j = clist[c0 + g0 + k];
              const EntT q = ent[j];
              const T dx = px - q.x, dy = py - q.y;
              const T d2 = dx * dx + dy * dy;
              const T far = size_i + q.size + cut;
              if (Prec<T>::kCut ? (d2 <= far * far) : (j < E)) m8 |= 1u << k;
            }
            cm |= m8 << g0;
          }
          while (cm) {
            const int k = __ffs(cm) - 1;
            cm &= cm - 1;
            const int j = clist[c0 + k];
            if (j == i) continue;
            const EntT q = ent[j];
            const T dx = px - q.x, dy = py - q.y;
            const T dist = A::sqrt(dx * dx + dy * dy);
            const T dmin = size_i + q.size;
            const T x = A::div_const(-(dist - dmin), p.km, p.km_inv);
            if (Prec<T>::kCut && x < (T)(-kFarCut)) continue;
            const T pen = softplus(x) * p.km;
            fx = fx + A::div(p.cf * dx, dist) * pen;
            fy = fy + A::div(p.cf * dy, dist) * pen;
          }
        }
      }
      // ---- SPEC §4 ------------------------------------------------------------------------------
      vx = vx * p.one_minus_damp; vy = vy * p.one_minus_damp;
      vx = vx + A::div(fx, mass_i) * p.dt;
      vy = vy + A::div(fy, mass_i) * p.dt;
      if (maxsp_i > (T)0) {
        const T sp = A::sqrt(vx * vx + vy * vy);
        if (sp > maxsp_i) { vx = A::div(vx, sp) * maxsp_i; vy = A::div(vy, sp) * maxsp_i; }
      }
      px = px + vx * p.dt; py = py + vy * p.dt;
    }
    t_now += 1;
    if (multi) __syncthreads(); else __syncwarp();     // everyone has read the old table
    if (active) { ent[i].x = px; ent[i].y = py; vel[2 * i] = vx; vel[2 * i + 1] = vy; PS::put(pairs, i, px, py); }
    if (multi) __syncthreads(); else __syncwarp();

    // ---- padding first: the rows of this warp's agents are one contiguous region of nbr_idx /
    // nbr_feat, so the warp clears it with fully coalesced 16-byte stores (-1 / zeros); the few
    // real neighbour rows are written over it after the __syncwarp.  (Per-lane padding loops cost
    // L1TEX one line per lane per instruction: ncu_r1_lane24.)
    int64_t w_row0;
    int w_nrows;
    {
      int64_t row0, nrows;
      if (multi) { row0 = env * N + warp * 32; nrows = N - warp * 32; nrows = nrows > 32 ? 32 : (nrows < 0 ? 0 : nrows); }
      else {
        const int64_t envw = (int64_t)blockIdx.x * g.envs_per_cta + warp * g.envs_per_warp;
        int64_t ne = p.n_envs - envw;
        ne = ne > g.envs_per_warp ? g.envs_per_warp : (ne < 0 ? 0 : ne);
        row0 = envw * N; nrows = ne * N;
      }
      if (GSM_LANE_DIAG & 1) nrows = 0;
      w_row0 = row0; w_nrows = (int)nrows;
      unsigned char* zf = c_feat_base + row0 * K * GSM_NBR_FEAT_DIM * (int64_t)sizeof(T);
      const int64_t bf = nrows * K * GSM_NBR_FEAT_DIM * (int64_t)sizeof(T);
      if ((((uintptr_t)zf | (uintptr_t)bf) & 15) == 0) {
        unsigned char* z = zf + lane * 16;                 // (a warp's block is < 2^31 bytes)
        for (int q = lane * 16, n = (int)bf; q < n; q += 512, z += 512) st_zero16<T>(z);
      } else {
        for (int64_t q = (int64_t)lane * sizeof(T); q < bf; q += 32 * sizeof(T)) *reinterpret_cast<T*>(zf + q) = (T)0;
      }
      __syncwarp();
    }

    // ---- SPEC §6-7: neighbour graph on the new table --------------------------------------------
    int cnt = 0, ncol = 0;
    T r = 0;
    const bool want_carry = carry_ok && coll_i && step + 1 < n_steps;
    cfx = 0; cfy = 0;
    if (active) {
      // per 32-entity block: a branch-free, sqrt-free sweep (unrolled groups of 8 over the padded
      // table) marks candidates on the squared distance (a few ulp inclusive); the set bits, in
      // ascending entity order, then decide on the rounded distance like every other kernel,
      // count collisions and write their rows.  fp64 marks every pair.
      const int goal_e = p.own_goal_always ? N + i : -1;
      for (int e0 = 0; e0 < E; e0 += 32) {
        uint32_t cand = 0;
        if (PS::kOn && col_in_nb) {                    // fp32: two entities per load, packed arithmetic
          for (int g0 = 0; g0 < 32 && e0 + g0 < E; g0 += 8)
            cand |= PS::mask8(pairs + ((e0 + g0) >> 1), -px, -py, Rs2) << g0;
        } else
        for (int g0 = 0; g0 < 32 && e0 + g0 < E; g0 += 8) {
          uint32_t m8 = 0;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const EntT q = ent[e0 + g0 + k];
            const T dx = q.x - px, dy = q.y - py;
            const T d2 = dx * dx + dy * dy;
            bool c;
            if (!Prec<T>::kCut) c = e0 + g0 + k < E;
            else if (col_in_nb) c = d2 < Rs2;
            else { const T dm = size_i + q.size; c = d2 < Rs2 || d2 < dm * dm * (T)1.000001; }
            if (c) m8 |= 1u << k;
          }
          cand |= m8 << g0;
        }
        if ((unsigned)(goal_e - e0) < 32u) cand |= 1u << (goal_e - e0);
        if ((unsigned)(i - e0) < 32u) cand &= ~(1u << (i - e0));
        const uint32_t cmask = costmask[e0 >> 5];
        uint32_t word = 0;
        while (cand) {
          const int k = __ffs(cand) - 1;
          cand &= cand - 1;
          const int e = e0 + k;
          const EntT q = ent[e];
          const T dx = q.x - px, dy = q.y - py;
          const T dist = A::sqrt(dx * dx + dy * dy);
          if (dist < size_i + q.size && ((cmask >> k) & 1u)) ncol++;
          if (want_carry && (q.flag & 1)) {            // SPEC §3 term of the NEXT step (dx is other - self here)
            const T x = A::div_const(-(dist - (size_i + q.size)), p.km, p.km_inv);
            if (!(x < (T)(-kFarCut))) {
              const T pen = softplus(x) * p.km;
              cfx = cfx - A::div(p.cf * dx, dist) * pen;
              cfy = cfy - A::div(p.cf * dy, dist) * pen;
            }
          }
          if (dist < p.Rs || e == goal_e) {
            word |= 1u << k;
            if (cnt < K && !((GSM_LANE_DIAG & 2) && cnt > 0)) {
              T evx = 0, evy = 0;
              if (e < N) { evx = vel[2 * e]; evy = vel[2 * e + 1]; }
              elist[cnt] = (uint16_t)e;
              T* f = (T*)c_feat + cnt * GSM_NBR_FEAT_DIM;
              row_store<T>(f, feat16, cnt & 1, dx, dy, evx - vx, evy - vy, dist, (T)(q.flag >> 1));
            }
            cnt++;
          }
        }
        if (!(GSM_LANE_DIAG & 4) || e0 == 0) ((uint32_t*)c_adj)[e0 >> 5] = word;
      }
      if (cnt > K) cnt = K;
      const T gx = gxl - px, gy = gyl - py;
      const T d = A::sqrt(gx * gx + gy * gy);
      r = ((T)0 - p.w_dist * d) + (d < p.goal_tol ? p.w_goal : (T)0);
      T* o = (T*)c_obs;
      st2<T>(o, vx, vy); st2<T>(o + 2, px, py); st2<T>(o + 4, gx, gy);
      *(int32_t*)c_cnt = cnt;
      *(T*)c_cost = (T)ncol;
      *c_done = (uint8_t)(t_now >= p.episode_length);
      *(int32_t*)c_asg = i;
      if (!p.share_reward) *(T*)c_rew = r;
    }
    // ---- nbr_idx: the warp's rows are one contiguous block of w_nrows * K words; pass q writes words
    //      [32 q, 32 q + 32) from the agents' neighbour lists (-1 from cnt on), fully coalesced ---------------
    __syncwarp();
    {
      const int npass = (w_nrows * K + 31) >> 5, last = w_nrows * K - lane;   // my word exists while 32 q < last
      int32_t* gi = (int32_t*)c_idx_base + w_row0 * K + lane;
      unsigned sl = list_sa + (unsigned)(idx_a0 * lw * 2 + idx_k0) * 2u;       // shared address of my list entry
      int a = idx_a0;
      if (idx_dk == 0) {                                   // K divides 32 (8, 16, 32): my row number never changes
        const unsigned dsl = (unsigned)(idx_da * lw * 4);
        for (int q = 0; q < npass; q++) {
          const int cnt_a = __shfl_sync(0xffffffffu, cnt, a);
          unsigned v;
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(sl) : "memory");
          if ((q << 5) < last) *gi = idx_k0 < cnt_a ? (int32_t)v : -1;
          a += idx_da; sl += dsl; gi += 32;
        }
      } else {
        int k = idx_k0;
        for (int q = 0; q < npass; q++) {
          const int cnt_a = __shfl_sync(0xffffffffu, cnt, a);
          unsigned v;
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(sl) : "memory");
          if ((q << 5) < last) *gi = k < cnt_a ? (int32_t)v : -1;
          a += idx_da; k += idx_dk;
          const bool wrap = k >= K;
          k -= wrap ? K : 0; a += wrap ? 1 : 0;
          sl += (unsigned)(idx_da * lw * 4 + idx_dk * 2) + (wrap ? (unsigned)(lw * 4 - K * 2) : 0u);
          gi += 32;
        }
      }
    }
    __syncwarp();
    if (p.share_reward) {                              // mean over the env's agents, ascending order
      if (active) rew[i] = r;
      if (multi) __syncthreads(); else __syncwarp();
      if (active) {
        T s = rew[0];
        for (int k = 1; k < N; k++) s = s + rew[k];
        *(T*)c_rew = s / (T)N;
      }
      if (multi) __syncthreads(); else __syncwarp();
    }
    // ---- episode end inside a fused rollout: re-draw the env's table (terminal outputs stay) ------
    if (auto_reset) {
      const bool rs = active && t_now >= p.episode_length;
      if (multi) __syncthreads(); else __syncwarp();   // everyone is done reading the table
      if (rs) {
        spawn_draw<T>(genv, ep, i, p.seed, p.ext[GSM_ENT_AGENT], px, py);
        vx = 0; vy = 0;
        ent[i].x = px; ent[i].y = py; vel[2 * i] = 0; vel[2 * i + 1] = 0;
        PS::put(pairs, i, px, py);
        for (int l = i; l < L; l += N) {
          T x, y;
          spawn_draw<T>(genv, ep, N + l, p.seed, p.ext[ent[N + l].flag >> 1], x, y);
          ent[N + l].x = x; ent[N + l].y = y;
          PS::put(pairs, N + l, x, y);
        }
      }
      if (multi) __syncthreads(); else __syncwarp();
      if (rs) { gxl = ent[N + i].x; gyl = ent[N + i].y; t_now = 0; ep += 1; }
      have_carry = want_carry && !rs;                  // a re-drawn table invalidates the carried force
    } else {
      have_carry = want_carry;
    }
    c_act += ss.actions; c_feat += ss.nbr_feat; c_obs += ss.obs;
    c_cnt += ss.nbr_cnt; c_adj += ss.adj; c_rew += ss.reward; c_cost += ss.cost;
    c_done += ss.done; c_asg += ss.assign;
    c_idx_base += ss.nbr_idx; c_feat_base += ss.nbr_feat;
  }

  if (active) {
    T* a = p.agent_state + row * 4;
    st2<T>(a, px, py); st2<T>(a + 2, vx, vy);
    if (i == 0) p.t[env] = t_now;
    if (auto_reset && ep != ep0) {
      for (int l = i; l < L; l += N) {
        T* lp = p.lm_pos + (env * L + l) * 2;
        lp[0] = ent[N + l].x; lp[1] = ent[N + l].y;
      }
      if (i == 0) p.episode[env] = ep;
    }
  }
}

}  // namespace gsm
