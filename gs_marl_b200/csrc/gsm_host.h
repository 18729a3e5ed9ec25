// gsm_host.h — precision-agnostic launch interface between gsm_api.cu and the two
// kernel translation units (gsm_kernels_f32.cu / gsm_kernels_f64.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gsmarl_b200.h"

namespace gsm {

struct HostParams {
  int64_t n_envs, env_offset;
  int N, L, K;
  int scenario, action_mode, n_actions, episode_length;
  int share_reward, cost_obstacles, own_goal_always;
  double dt, damping, cf, km, Rs, w_dist, w_goal, goal_tol, poly_r;
  double discrete_u[GSM_MAX_DISCRETE][2];
  double ext[4];
  // device arrays (typed by the handle's dtype)
  const void* size; const uint8_t* eflag; const void* mass; const void* accel;
  const void* max_speed; const void* slot_table;
  void* agent_state; void* lm_pos; int32_t* t; int32_t* episode;
  int auto_reset;        // set per launch by the API layer (only gsm_rollout turns it on)
  uint64_t seed;         // seed of the handle's last reset
  // host copies of the per-entity constants for kernels that take them in the parameter space
  // (gsm_kernels_wide.cuh); h_consts = 1 when E <= kHostConstE and N <= kHostConstN
  int h_consts;
  double h_size[32], h_mass[8], h_accel[8], h_maxsp[8];
  int32_t h_eflag[32];
};
constexpr int kHostConstE = 32, kHostConstN = 8;

struct RolloutStrides {  // bytes between consecutive steps of each rollout buffer
  int64_t actions, obs, nbr_idx, nbr_feat, nbr_cnt, adj, reward, cost, done, assign;
};

struct LaunchPlan {     // chosen once per handle
  int cta_env;          // 1: one env per CTA; 0: packed envs per warp
  int P;                // lanes per agent
  int envs_per_warp;    // packed
  int envs_per_cta;
  size_t smem;
  int64_t grid;
  int spec;             // 1: a size-specialised register-resident kernel exists (gsm_kernels_spec.cuh)
  int lane;             // 1: the lane-per-agent kernel applies (gsm_kernels_lane.cuh)
  int team;             // 1: a polygon/line group-LSA instance exists (gsm_kernels_team.cuh)
};

// Each returns a cudaError_t (as int).  physics: 1 = full step, 0 = observe only.
int plan_f32(const HostParams& hp, LaunchPlan* plan);
int plan_f64(const HostParams& hp, LaunchPlan* plan);
int launch_env_f32(const HostParams& hp, const LaunchPlan& plan, const gsm_step_io& io, int physics,
                   const uint8_t* mask, int64_t mask_stride, cudaStream_t st);
int launch_env_f64(const HostParams& hp, const LaunchPlan& plan, const gsm_step_io& io, int physics,
                   const uint8_t* mask, int64_t mask_stride, cudaStream_t st);
// n_steps consecutive steps in ONE launch of the specialised kernel; returns -1 if the
// handle's (scenario, N, L) has no compiled instance.
// observe != 0: observe-only variant (reset path) with an optional per-env mask.
int launch_spec_f32(const HostParams& hp, const gsm_step_io& io, int n_steps,
                    const RolloutStrides& rs, int observe, const uint8_t* mask,
                    int64_t mask_stride, cudaStream_t st);
int launch_spec_f64(const HostParams& hp, const gsm_step_io& io, int n_steps,
                    const RolloutStrides& rs, int observe, const uint8_t* mask,
                    int64_t mask_stride, cudaStream_t st);
// Lane-per-agent navigation kernel (n_steps fused); returns -1 if it does not apply.
int launch_lane_f32(const HostParams& hp, const gsm_step_io& io, int n_steps,
                    const RolloutStrides& rs, cudaStream_t st);
int launch_lane_f64(const HostParams& hp, const gsm_step_io& io, int n_steps,
                    const RolloutStrides& rs, cudaStream_t st);
// Polygon / line kernel with the group-parallel assignment; returns -1 if no instance.
int launch_team_f32(const HostParams& hp, const gsm_step_io& io, int n_steps,
                    const RolloutStrides& rs, cudaStream_t st);
int launch_team_f64(const HostParams& hp, const gsm_step_io& io, int n_steps,
                    const RolloutStrides& rs, cudaStream_t st);
int launch_reset_f32(const HostParams& hp, uint64_t seed, const uint8_t* mask, int64_t mask_stride,
                     cudaStream_t st);
int launch_reset_f64(const HostParams& hp, uint64_t seed, const uint8_t* mask, int64_t mask_stride,
                     cudaStream_t st);
int launch_lsa_f32(const void* cost, int32_t* col4row, int64_t n_problems, int n, cudaStream_t st);
int launch_lsa_f64(const void* cost, int32_t* col4row, int64_t n_problems, int n, cudaStream_t st);

}  // namespace gsm
