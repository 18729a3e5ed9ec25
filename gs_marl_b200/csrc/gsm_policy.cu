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
// Design for sm_100a: one thread per agent row, no shuffles, shared memory only for the 128-entry
// sort below.  The weights ride in the kernel PARAMETER space (9.5 KB, __grid_constant__): the
// 64-wide hidden loops are fully unrolled, so every weight has an immediate constant-bank offset,
// ptxas fetches four at a time with LDCU.128 into uniform registers and the FFMAs take them as
// uniform-register operands — no weight ever goes through a load/store unit.  (Partially unrolled
// loops with a uniform-register INDEX into the constant bank measured 4x slower; packed FFMA2
// variants did not pay either: profiles/README.md, profiles/fma_peak.cu.)  The head is linear, so
// W_h's neighbour half is applied to every row's m_r on the fly (hr = W_h[:, H:] m_r) and the
// softmax is the online (running max / running sum) form over 1 + n_actions accumulators: a row
// costs 64 x 13 FFMA-class instructions and nothing of size H is ever stored.  Only the cnt valid
// rows are visited; the padded rows the env kernel zero-fills are never read.  Bound: fp32 issue
// slots (17.25 per 13 FFMA; 88.7 % issue-active at 786k rows), not HBM (220 B per agent in, 8 B out).
#include <cuda_runtime.h>

#include <cstdint>
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
template <int NA, int NV>
__global__ void __launch_bounds__(128)
graph_actor_kernel(const __grid_constant__ PolicyParams w, const __grid_constant__ PolicyIO io) {
  // The row loop runs cnt times and cnt differs from agent to agent (0..K): a warp would pay for
  // its busiest lane (measured: 20.8 of 32 lanes active per instruction).  So the block first
  // counting-sorts its 128 agents by cnt, busiest first, and thread t serves the t-th agent of
  // that order: lanes of one warp then loop (almost) equally long.  Outputs are per agent, so the
  // order inside a bin does not matter.
  __shared__ int s_bin[34];
  __shared__ int s_perm[128];
  const int tid = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * 128;
  if (tid < 34) s_bin[tid] = 0;
  __syncthreads();
  int my_cnt = -1, my_bin = 33, my_rank = 0;
  if (base + tid < io.n_rows) {
    my_cnt = io.nbr_cnt[base + tid];
    my_cnt = my_cnt < 0 ? 0 : (my_cnt > io.K ? io.K : my_cnt);
    my_bin = 32 - (my_cnt > 32 ? 32 : my_cnt);
    my_rank = atomicAdd(&s_bin[my_bin], 1);
  }
  s_perm[tid] = -1;
  __syncthreads();
  if (tid < 32) {                       // exclusive scan of the 33 bins by warp 0
    const int v = s_bin[tid];
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(~0u, x, d); if (tid >= d) x += y; }
    const int tot = __shfl_sync(~0u, x, 31);
    __syncwarp();
    s_bin[tid] = x - v;
    if (tid == 0) s_bin[32] = tot;      // bin 32 (cnt == 0) starts after bins 0..31
  }
  __syncthreads();
  if (my_cnt >= 0) s_perm[s_bin[my_bin] + my_rank] = tid | (my_cnt << 8);
  __syncthreads();
  const int slot = s_perm[tid];
  if (slot < 0) return;
  const int64_t i = base + (slot & 0xff);
  const int cnt = slot >> 8;

  constexpr int NZ = NA + NV;
  float z[NZ];
#pragma unroll
  for (int a = 0; a < NZ; a++) z[a] = w.head_b[a];

  {  // ego branch
    const float2* o2 = reinterpret_cast<const float2*>(io.obs + i * GSM_OBS_DIM);
    const float2 o01 = o2[0], o23 = o2[1], o45 = o2[2];
#pragma unroll
    for (int j = 0; j < H; j++) {
      float e = w.ego_b[j];
      e = fmaf(w.ego_w[j][0], o01.x, e); e = fmaf(w.ego_w[j][1], o01.y, e);
      e = fmaf(w.ego_w[j][2], o23.x, e); e = fmaf(w.ego_w[j][3], o23.y, e);
      e = fmaf(w.ego_w[j][4], o45.x, e); e = fmaf(w.ego_w[j][5], o45.y, e);
      e = fmaxf(e, 0.f);
#pragma unroll
      for (int a = 0; a < NZ; a++) z[a] = fmaf(w.head_w[a][j], e, z[a]);
    }
  }

  // neighbour rows: online softmax over the attention score, W_h's second half applied per row
  float mx = -__int_as_float(0x7f800000), s = 0.f, acc[NZ];
#pragma unroll
  for (int a = 0; a < NZ; a++) acc[a] = 0.f;
  const float2* f2 = reinterpret_cast<const float2*>(io.nbr_feat + i * (int64_t)io.K * GSM_NBR_FEAT_DIM);
  float2 n01, n23, n45;                 // next row, loaded one iteration ahead
  if (cnt > 0) { n01 = f2[0]; n23 = f2[1]; n45 = f2[2]; }
  for (int r = 0; r < cnt; r++) {
    const float2 f01 = n01, f23 = n23, f45 = n45;
    if (r + 1 < cnt) { n01 = f2[3 * r + 3]; n23 = f2[3 * r + 4]; n45 = f2[3 * r + 5]; }
    float t = w.att_b, hr[NZ];
#pragma unroll
    for (int a = 0; a < NZ; a++) hr[a] = 0.f;
#pragma unroll
    for (int j = 0; j < H; j++) {
      float m = w.nbr_b[j];
      m = fmaf(w.nbr_w[j][0], f01.x, m); m = fmaf(w.nbr_w[j][1], f01.y, m);
      m = fmaf(w.nbr_w[j][2], f23.x, m); m = fmaf(w.nbr_w[j][3], f23.y, m);
      m = fmaf(w.nbr_w[j][4], f45.x, m); m = fmaf(w.nbr_w[j][5], f45.y, m);
      m = fmaxf(m, 0.f);
      t = fmaf(w.att_w[j], m, t);
#pragma unroll
      for (int a = 0; a < NZ; a++) hr[a] = fmaf(w.head_w[a][H + j], m, hr[a]);
    }
    const float nm = fmaxf(mx, t);
    const float sc = expf(mx - nm), p = expf(t - nm);   // first row: exp(-inf) = 0
    s = fmaf(s, sc, p);
#pragma unroll
    for (int a = 0; a < NZ; a++) acc[a] = fmaf(acc[a], sc, p * hr[a]);
    mx = nm;
  }
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
          const float u = ((float)(r[q] >> 8) + 0.5f) * 5.9604644775390625e-8f;   // (0, 1)
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
  const int block = 128;            // the kernel's counting sort assumes exactly 128
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
  explicit DevGuard(int d) { cudaGetDevice(&prev); if (d != prev) cudaSetDevice(d); }
  ~DevGuard() { int cur; cudaGetDevice(&cur); if (cur != prev && prev >= 0) cudaSetDevice(prev); }
};
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
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return pfail(GSM_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  }
  DevGuard guard(device);
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
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return pfail(GSM_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  }
  if (n_rows == 0) return GSM_OK;
  DevGuard guard(device);
  gsm::GaeParams p{reward, cost, values, done, returns, advantages, n_rows, slot_rows, n_steps, gamma, lam};
  const int64_t threads = n_rows * GSM_POLICY_VALUE_HEADS;
  gsm::gae_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return pfail(GSM_ERR_CUDA, cudaGetErrorString(e));
  return GSM_OK;
}

}  // extern "C"
