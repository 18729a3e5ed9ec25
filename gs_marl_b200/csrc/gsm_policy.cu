// gsm_policy.cu — SURVEY.md §8 row f3: the graph-attention actor's forward pass over the padded
// neighbour rows the env kernels write, plus action sampling, as ONE kernel, and the collect
// loop (row f1) as a C-level launch sequence of {actor, env step} pairs.
//
// Reference side (withheld): the GNN encoder of gsmarl/algorithms/* over torch-geometric
// (requirements.txt:119) called from runner/mpe_runner.py's collect (SOURCES.txt:28).  The
// architecture is therefore DECLARED (SPEC.md §10), not GS-MARL's: shared parameters over agents,
//   e   = relu(W_e obs + b_e)                               [H]
//   m_r = relu(W_n feat_r + b_n),  r < cnt                  [H] per valid neighbour row
//   a_r = softmax_r(w_a . m_r + b_a)                        masked to the cnt valid rows
//   z   = W_h [e ; sum_r a_r m_r] + b_h                     [n_actions] logits
//   action = argmax_k (z_k + Gumbel_k),  Gumbel from Philox4x32-10 keyed like the env resets.
//
// Design for sm_100a (details in front of graph_actor_kernel below): a block owns 128 agents; the
// ego branch runs as packed FFMA2 over agent pairs, the block's VALID neighbour rows are flattened
// into one list of work items processed two per thread as packed FFMA2 (no divergence on the
// per-agent row count), and a per-agent online softmax folds the parked row results.  The weights
// ride in the kernel PARAMETER space (9.5 KB, __grid_constant__): the 64-wide hidden loops are fully
// unrolled, every weight has an immediate constant-bank offset, ptxas fetches four at a time with
// LDCU.128 into uniform registers and the FFMA2s take them as the broadcast uniform-scalar operand —
// no weight ever goes through a load/store unit.  The head is linear, so W_h's neighbour half is
// applied to every row's m_r on the fly (hr = W_h[:, H:] m_r): nothing of size H is ever stored.
// Only the cnt valid rows are visited; the padded rows the env kernel zero-fills are never read.
// Bound: fp32 issue / FMA pipe (54 % of the measured 72 TFLOP/s FFMA peak at 786k rows), not HBM
// (220 B per agent in, 8 B out).  History of the rejected forms (thread per agent, hidden-unit pair
// packing, shared-memory weights with rolled loops): profiles/README.md, profiles/rejected/.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/gsmarl_b200.h"

namespace gsm {

constexpr int H = GSM_POLICY_HIDDEN;

struct PolicyParams {                 // kernel-parameter image of gsm_policy_weights
  float ego_w[H][GSM_OBS_DIM];
  float ego_b[H];
  float nbr_w[H][GSM_NBR_FEAT_DIM];
  float nbr_b[H];
  float att_w[H];
  // rows 0..NA-1: action logits; rows NA, NA+1: the two critics (reward value, cost value)
  float head_w[GSM_POLICY_MAX_ACTIONS + GSM_POLICY_VALUE_HEADS][2 * H];
  float head_b[GSM_POLICY_MAX_ACTIONS + GSM_POLICY_VALUE_HEADS];
  float att_b;
};

struct PolicyIO {
  const float* obs; const float* nbr_feat; const int32_t* nbr_cnt;
  int32_t* actions; float* logp; float* logits; float* values;
  int64_t n_rows; uint64_t row_offset, seed, step;
  int K, greedy;
};

__device__ __forceinline__ void philox_p(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                         uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// NV = 0: actor only; NV = GSM_POLICY_VALUE_HEADS: the critics ride along as NV more rows of the head
// (+2 FFMA per hidden unit and row), for collect loops that store value predictions.
//
// Work decomposition ("flattened rows").  A block owns 128 consecutive agents.  Phase 1: the ego
// branch, agents (t, t + 64) packed into the two halves of float2 registers on warps 0-1.  Phase 2:
// the block's valid neighbour rows — sum(cnt) of them, found through an exclusive prefix of cnt —
// are one flat list of independent work items; a thread takes items q and q + 128 (agent by binary
// search in the prefix, 7 steps), again packed in float2 halves, computes their attention scores and
// head contributions and parks 1 + NZ floats per row in shared memory.  Phase 3: thread a folds its
// own agent's rows, in row order, into the online softmax.  Every lane of every warp carries a real
// row in phase 2 whatever the per-agent counts are (30 of 32 lanes active; the thread-per-agent
// version had 20.8, and 28.1 after a per-block counting sort by cnt).  Row items are processed in
// chunks of CH so that shared memory does not depend on K.
//
// Arithmetic: Blackwell's packed fp32 FMA (FFMA2, `fma.rn.f32x2`) with the weight as the BROADCAST
// uniform scalar operand (`FFMA2 R, R.F32x2, UR.F32, R.F32x2`) — profiles/fma_peak.cu measures that
// form at the full 127 FMA/clk/SM for half the issue slots of scalar FFMA, while a constant PAIR
// operand (packing two hidden units instead of two rows) runs at 32.  Feature loads are scalar on
// purpose: a 64-bit load pins (x, y) of one row to an aligned register pair and ptxas then
// re-assembles every (row 0, row 1) operand pair with two MOVs per FFMA2.
__device__ __forceinline__ float2 bc(float w) { return make_float2(w, w); }

template <int NA, int NV>
__global__ void __launch_bounds__(128)
graph_actor_kernel(const __grid_constant__ PolicyParams w, const __grid_constant__ PolicyIO io) {
  constexpr int NZ = NA + NV, AG = 128, RS = NZ + 1, CH = RS <= 10 ? 1024 : 512;   // static smem <= 48 KB
  __shared__ int s_pre[AG + 1];
  __shared__ int s_wtot[4];
  __shared__ float s_row[CH * RS];
  __shared__ float s_z[AG * NZ];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * AG;
  const int64_t i = base + tid;
  const bool valid = i < io.n_rows;
  int cnt = 0;
  if (valid) { cnt = io.nbr_cnt[i]; cnt = cnt < 0 ? 0 : (cnt > io.K ? io.K : cnt); }
  {  // exclusive prefix of cnt over the block
    int x = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(~0u, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_wtot[warp] = x;
    __syncthreads();
    int off = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) off += k < warp ? s_wtot[k] : 0;
    s_pre[tid] = off + x - cnt;
    if (tid == AG - 1) s_pre[AG] = off + x;
  }

  // phase 1: ego branch, agents (t, t + 64) packed into one thread of warps 0-1 (FFMA2, as below);
  // warps 2-3 go straight to the barrier and leave their issue slots to other blocks
  if (tid < AG / 2 && valid) {
    const int64_t iB = i + AG / 2 < io.n_rows ? i + AG / 2 : i;
    const float* oA = io.obs + i * GSM_OBS_DIM;
    const float* oB = io.obs + iB * GSM_OBS_DIM;
    const float2 d0 = make_float2(oA[0], oB[0]), d1 = make_float2(oA[1], oB[1]), d2 = make_float2(oA[2], oB[2]),
                 d3 = make_float2(oA[3], oB[3]), d4 = make_float2(oA[4], oB[4]), d5 = make_float2(oA[5], oB[5]);
    float2 z2[NZ];
#pragma unroll
    for (int a = 0; a < NZ; a++) z2[a] = bc(w.head_b[a]);
#pragma unroll
    for (int j = 0; j < H; j++) {
      float2 e = bc(w.ego_b[j]);
      e = __ffma2_rn(bc(w.ego_w[j][0]), d0, e); e = __ffma2_rn(bc(w.ego_w[j][1]), d1, e);
      e = __ffma2_rn(bc(w.ego_w[j][2]), d2, e); e = __ffma2_rn(bc(w.ego_w[j][3]), d3, e);
      e = __ffma2_rn(bc(w.ego_w[j][4]), d4, e); e = __ffma2_rn(bc(w.ego_w[j][5]), d5, e);
      e.x = fmaxf(e.x, 0.f); e.y = fmaxf(e.y, 0.f);
#pragma unroll
      for (int a = 0; a < NZ; a++) z2[a] = __ffma2_rn(bc(w.head_w[a][j]), e, z2[a]);
    }
#pragma unroll
    for (int a = 0; a < NZ; a++) { s_z[tid * NZ + a] = z2[a].x; s_z[(tid + AG / 2) * NZ + a] = z2[a].y; }
  }
  __syncthreads();                      // s_pre complete
  const int total = s_pre[AG];
  const int my0 = s_pre[tid], my1 = my0 + cnt;
  float z[NZ];
#pragma unroll
  for (int a = 0; a < NZ; a++) z[a] = s_z[tid * NZ + a];       // garbage for !valid threads, never used

  // online softmax state of MY agent over its rows, in row order
  float mx = -__int_as_float(0x7f800000), s = 0.f, acc[NZ];
#pragma unroll
  for (int a = 0; a < NZ; a++) acc[a] = 0.f;
  const float* ffeat = io.nbr_feat + base * (int64_t)io.K * GSM_NBR_FEAT_DIM;

  for (int c0 = 0; c0 < total; c0 += CH) {          // block-uniform trip count
    const int cend = total < c0 + CH ? total : c0 + CH;
    // phase 2: TWO neighbour rows per thread (items q and q + AG), packed into the halves of float2
    // registers for Blackwell's packed fp32 FMA (FFMA2, `fma.rn.f32x2`) with the weight as the
    // broadcast uniform scalar operand — full FMA rate at half the issue slots of scalar FFMA
    // (profiles/fma_peak.cu).  Items are homogeneous, so the pairing costs no divergence; an odd
    // tail duplicates its row into the second half and stores it once.
    for (int q = c0 + tid; q < cend; q += 2 * AG) {
      const int q1 = q + AG < cend ? q + AG : q;
      int lo0 = 0, hi0 = AG, lo1 = 0, hi1 = AG;     // largest a with s_pre[a] <= q
#pragma unroll
      for (int it = 0; it < 7; it++) {
        const int m0 = (lo0 + hi0) >> 1, m1 = (lo1 + hi1) >> 1;
        if (s_pre[m0] <= q) lo0 = m0; else hi0 = m0;
        if (s_pre[m1] <= q1) lo1 = m1; else hi1 = m1;
      }
      // scalar loads on purpose: a 64-bit load pins (x, y) of ONE row to an aligned register pair and
      // ptxas then re-assembles every (row 0, row 1) operand pair with two MOVs per FFMA2
      const float* g0 = ffeat + ((int64_t)lo0 * io.K + (q - s_pre[lo0])) * GSM_NBR_FEAT_DIM;
      const float* g1 = ffeat + ((int64_t)lo1 * io.K + (q1 - s_pre[lo1])) * GSM_NBR_FEAT_DIM;
      const float2 d0 = make_float2(g0[0], g1[0]), d1 = make_float2(g0[1], g1[1]), d2 = make_float2(g0[2], g1[2]),
                   d3 = make_float2(g0[3], g1[3]), d4 = make_float2(g0[4], g1[4]), d5 = make_float2(g0[5], g1[5]);
      float2 t2 = bc(w.att_b), hr[NZ];
#pragma unroll
      for (int a = 0; a < NZ; a++) hr[a] = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < H; j++) {
        float2 m = bc(w.nbr_b[j]);
        m = __ffma2_rn(bc(w.nbr_w[j][0]), d0, m); m = __ffma2_rn(bc(w.nbr_w[j][1]), d1, m);
        m = __ffma2_rn(bc(w.nbr_w[j][2]), d2, m); m = __ffma2_rn(bc(w.nbr_w[j][3]), d3, m);
        m = __ffma2_rn(bc(w.nbr_w[j][4]), d4, m); m = __ffma2_rn(bc(w.nbr_w[j][5]), d5, m);
        m.x = fmaxf(m.x, 0.f); m.y = fmaxf(m.y, 0.f);
        t2 = __ffma2_rn(bc(w.att_w[j]), m, t2);
#pragma unroll
        for (int a = 0; a < NZ; a++) hr[a] = __ffma2_rn(bc(w.head_w[a][H + j]), m, hr[a]);
      }
      float* o0 = s_row + (q - c0) * RS;
      o0[0] = t2.x;
#pragma unroll
      for (int a = 0; a < NZ; a++) o0[1 + a] = hr[a].x;
      if (q1 != q) {
        float* o1 = s_row + (q1 - c0) * RS;
        o1[0] = t2.y;
#pragma unroll
        for (int a = 0; a < NZ; a++) o1[1 + a] = hr[a].y;
      }
    }
    __syncthreads();
    {  // phase 3: fold my agent's rows of this chunk
      const int lo = my0 > c0 ? my0 : c0, hi = my1 < cend ? my1 : cend;
      for (int q = lo; q < hi; q++) {
        const float* o = s_row + (q - c0) * RS;
        const float t = o[0];
        const float nm = fmaxf(mx, t);
        const float sc = expf(mx - nm), pw = expf(t - nm);   // first row: exp(-inf) = 0
        s = fmaf(s, sc, pw);
#pragma unroll
        for (int a = 0; a < NZ; a++) acc[a] = fmaf(acc[a], sc, pw * o[1 + a]);
        mx = nm;
      }
    }
    __syncthreads();                    // before the next chunk overwrites s_row
  }
  if (!valid) return;
  if (cnt > 0) {
    const float inv = 1.f / s;
#pragma unroll
    for (int a = 0; a < NZ; a++) z[a] = fmaf(acc[a], inv, z[a]);
  }
  if (NV > 0) {
#pragma unroll
    for (int v = 0; v < NV; v++) io.values[i * NV + v] = z[NA + v];
  }

  if (io.logits) {
#pragma unroll
    for (int a = 0; a < NA; a++) io.logits[i * NA + a] = z[a];
  }

  // sample: Gumbel-max, one Philox block per 4 actions; counter (row lo, row hi, step, block)
  int best = 0;
  float zb = z[0];                      // logit of the chosen action
  if (io.greedy) {
#pragma unroll
    for (int a = 1; a < NA; a++) if (z[a] > zb) { zb = z[a]; best = a; }
  } else {
    const uint64_t g = io.row_offset + (uint64_t)i;
    float bv = 0.f;
#pragma unroll
    for (int b = 0; b < (NA + 3) / 4; b++) {
      uint32_t r[4];
      philox_p((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)io.step, 0x80000000u | (uint32_t)b,
               (uint32_t)io.seed, (uint32_t)(io.seed >> 32), r);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int a = 4 * b + q;
        if (a < NA) {
          // 23 random bits + half a step: every value k * 2^-23 + 2^-24 is exact in fp32 and lies in (0, 1)
          const float u = (float)(r[q] >> 9) * 1.1920928955078125e-7f + 5.9604644775390625e-8f;
          const float v = z[a] - logf(-logf(u));
          if (a == 0 || v > bv) { bv = v; best = a; zb = z[a]; }
        }
      }
    }
  }
  io.actions[i] = best;
  if (io.logp) {
    float zm = z[0];
#pragma unroll
    for (int a = 1; a < NA; a++) zm = fmaxf(zm, z[a]);
    float se = 0.f;
#pragma unroll
    for (int a = 0; a < NA; a++) se += expf(z[a] - zm);
    io.logp[i] = zb - zm - logf(se);
  }
}

// ---- buffer.compute_returns / compute_cost_returns (utils/graph_separated_buffer.py, SOURCES.txt:33;
// withheld — the on-policy lineage's GAE recursion, [DECL] SPEC.md §11): one thread per (row, head),
// backward scan over the T slots; consecutive threads touch consecutive addresses in every slot.
struct GaeParams {
  const float* reward; const float* cost; const float* values; const uint8_t* done;
  float* returns; float* adv;
  int64_t n_rows, slot_rows;
  int T;
  float gamma, lam;
};
__global__ void __launch_bounds__(256) gae_kernel(const __grid_constant__ GaeParams p) {
  constexpr int V = GSM_POLICY_VALUE_HEADS;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= p.n_rows * V) return;
  const int64_t row = g / V;
  const int head = (int)(g - row * V);
  const float* rw = head == 0 ? p.reward : p.cost;
  float gae = 0.f;
  float v_next = p.values[((int64_t)p.T * p.slot_rows + row) * V + head];
  for (int t = p.T - 1; t >= 0; t--) {
    const int64_t o = (int64_t)t * p.slot_rows + row;
    const float mask = p.done[o] ? 0.f : 1.f;      // done at step t: slot t+1 starts a new episode
    const float v = p.values[o * V + head];
    const float delta = rw[o] + p.gamma * v_next * mask - v;
    gae = delta + p.gamma * p.lam * mask * gae;
    if (p.adv) p.adv[o * V + head] = gae;
    p.returns[o * V + head] = gae + v;
    v_next = v;
  }
}

static int launch_actor(const PolicyParams& w, const PolicyIO& io, int n_actions, cudaStream_t st) {
  if (io.n_rows == 0) return 0;
  const int block = 128;            // = AG, the agents a block owns
  const int64_t grid = (io.n_rows + block - 1) / block;
  constexpr int V = GSM_POLICY_VALUE_HEADS;
  const unsigned gd = (unsigned)grid;
  if (n_actions == 5 && !io.values) graph_actor_kernel<5, 0><<<gd, block, 0, st>>>(w, io);
  else if (n_actions == 5) graph_actor_kernel<5, V><<<gd, block, 0, st>>>(w, io);
  else if (n_actions == 9 && !io.values) graph_actor_kernel<9, 0><<<gd, block, 0, st>>>(w, io);
  else if (n_actions == 9) graph_actor_kernel<9, V><<<gd, block, 0, st>>>(w, io);
  else return -1;
  return (int)cudaGetLastError();
}

}  // namespace gsm

// ---- C ABI ------------------------------------------------------------------------------------
namespace {
thread_local char g_policy_err[256] = "";
int pfail(int status, const char* msg) {
  std::strncpy(g_policy_err, msg, sizeof(g_policy_err) - 1);
  return status;
}
int pack(const gsm_policy_weights* w, gsm::PolicyParams* p) {
  if (!w) return pfail(GSM_ERR_INVALID_ARG, "gsm_policy: weights is NULL");
  if (w->struct_size != sizeof(gsm_policy_weights)) return pfail(GSM_ERR_ABI, "gsm_policy: weights.struct_size mismatch");
  if (w->n_actions != 5 && w->n_actions != 9)
    return pfail(GSM_ERR_UNSUPPORTED, "gsm_policy: compiled actor instances exist for n_actions 5 and 9 only");
  static_assert(sizeof(p->ego_w) == sizeof(w->ego_w), "layout");
  std::memcpy(p->ego_w, w->ego_w, sizeof(p->ego_w));   std::memcpy(p->ego_b, w->ego_b, sizeof(p->ego_b));
  std::memcpy(p->nbr_w, w->nbr_w, sizeof(p->nbr_w));   std::memcpy(p->nbr_b, w->nbr_b, sizeof(p->nbr_b));
  std::memcpy(p->att_w, w->att_w, sizeof(p->att_w));   p->att_b = w->att_b;
  std::memset(p->head_w, 0, sizeof(p->head_w)); std::memset(p->head_b, 0, sizeof(p->head_b));
  std::memcpy(p->head_w, w->head_w, sizeof(float) * 2 * GSM_POLICY_HIDDEN * w->n_actions);
  std::memcpy(p->head_b, w->head_b, sizeof(float) * w->n_actions);
  std::memcpy(p->head_w[w->n_actions], w->value_w, sizeof(w->value_w));   // critics right behind the logits' rows
  std::memcpy(&p->head_b[w->n_actions], w->value_b, sizeof(w->value_b));
  return GSM_OK;
}
struct DevGuard {
  int prev = -1;
  bool ok = true;
  explicit DevGuard(int d) {
    cudaGetDevice(&prev);
    if (d != prev) ok = cudaSetDevice(d) == cudaSuccess;
  }
  ~DevGuard() { int cur; cudaGetDevice(&cur); if (cur != prev && prev >= 0) cudaSetDevice(prev); }
};
// 0 <= device < device count, as gsm_create checks; GSM_OK or the error status to return.
int check_device(int device, const char* who) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return pfail(GSM_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  }
  if (device < 0 || device >= ndev) {
    static thread_local char msg[96];
    snprintf(msg, sizeof(msg), "%s: device index out of range", who);
    return pfail(GSM_ERR_INVALID_ARG, msg);
  }
  return GSM_OK;
}
}  // namespace

extern "C" {

const char* gsm_policy_last_error(void) { return g_policy_err; }

int gsm_policy_act(const gsm_policy_weights* w, const gsm_policy_io* io, int device, void* stream) {
  gsm::PolicyParams p;
  int st = pack(w, &p);
  if (st) return st;
  if (!io || !io->obs || !io->nbr_feat || !io->nbr_cnt || !io->actions)
    return pfail(GSM_ERR_INVALID_ARG, "gsm_policy_act: obs, nbr_feat, nbr_cnt and actions are required");
  if (io->n_rows < 0 || io->max_nbrs < 1) return pfail(GSM_ERR_INVALID_ARG, "gsm_policy_act: bad n_rows / max_nbrs");
  if (io->n_rows > ((int64_t)1 << 31) * 128 - 128) return pfail(GSM_ERR_INVALID_ARG, "gsm_policy_act: n_rows too large");
  if (const int ds = check_device(device, "gsm_policy_act")) return ds;
  DevGuard guard(device);
  if (!guard.ok) return pfail(GSM_ERR_CUDA, "gsm_policy_act: cudaSetDevice failed");
  gsm::PolicyIO k;
  k.obs = io->obs; k.nbr_feat = io->nbr_feat; k.nbr_cnt = io->nbr_cnt;
  k.actions = io->actions; k.logp = io->logp; k.logits = io->logits; k.values = io->values;
  k.n_rows = io->n_rows; k.row_offset = io->row_offset; k.seed = io->seed; k.step = io->step;
  k.K = io->max_nbrs; k.greedy = io->greedy;
  const int e = gsm::launch_actor(p, k, w->n_actions, (cudaStream_t)stream);
  if (e) return pfail(GSM_ERR_CUDA, cudaGetErrorString((cudaError_t)e));
  return GSM_OK;
}

int gsm_gae(const float* reward, const float* cost, const float* values, const uint8_t* done, int32_t n_steps,
            int64_t n_rows, int64_t slot_rows, float gamma, float lam, float* returns, float* advantages,
            int device, void* stream) {
  if (!reward || !cost || !values || !done || !returns)
    return pfail(GSM_ERR_INVALID_ARG, "gsm_gae: reward, cost, values, done and returns are required");
  if (n_steps < 1 || n_rows < 0 || slot_rows < n_rows) return pfail(GSM_ERR_INVALID_ARG, "gsm_gae: bad sizes");
  if (const int ds = check_device(device, "gsm_gae")) return ds;
  if (n_rows == 0) return GSM_OK;
  DevGuard guard(device);
  if (!guard.ok) return pfail(GSM_ERR_CUDA, "gsm_gae: cudaSetDevice failed");
  gsm::GaeParams p{reward, cost, values, done, returns, advantages, n_rows, slot_rows, n_steps, gamma, lam};
  const int64_t threads = n_rows * GSM_POLICY_VALUE_HEADS;
  gsm::gae_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return pfail(GSM_ERR_CUDA, cudaGetErrorString(e));
  return GSM_OK;
}

}  // extern "C"
