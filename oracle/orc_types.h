/*
 * orc_types.h — the oracle's OWN data types (SPEC.md §1, §6-7).
 *
 * TEST INFRASTRUCTURE ONLY.  Deliberately independent of include/gsmarl_b200.h: the oracle must
 * not inherit a mistake of the product's header (VERDICT r1 "break the oracle's dependence on
 * product code").  The two layouts are meant to agree — tests/test_abi.py compares sizeof /
 * offsetof of both with a C compiler — but each is declared from SPEC.md on its own.
 */
#ifndef ORC_TYPES_H
#define ORC_TYPES_H
#include <stdint.h>

#define ORC_OBS_DIM 6        /* SPEC §6: (v.x, v.y, p.x, p.y, target - p)            */
#define ORC_NBR_FEAT_DIM 6   /* SPEC §6: (d.x, d.y, dv.x, dv.y, dist, type)          */
#define ORC_MAX_LSA_N 32

enum { ORC_SCN_NAVIGATION = 0, ORC_SCN_POLYGON = 1, ORC_SCN_LINE = 2 };
enum { ORC_ACT_DISCRETE = 0, ORC_ACT_CONTINUOUS = 1 };
enum { ORC_ENT_AGENT = 0, ORC_ENT_GOAL = 1, ORC_ENT_OBSTACLE = 2, ORC_ENT_MARKER = 3 };

typedef struct orc_config {
  uint32_t struct_size, abi_version;
  int32_t dtype, scenario, action_mode;
  int32_t n_agents, n_landmarks, max_nbrs, episode_length, n_discrete_actions;
  int32_t share_reward, cost_obstacles, own_goal_always, reserved0;
  double dt, damping, contact_force, contact_margin, sensing_radius;
  double w_dist, w_goal, goal_tol, polygon_radius;
  double spawn_extent[4];          /* per entity type: reset draws U(-e, e)^2 (SPEC §8)          */
  const double* discrete_u;        /* [A][2]                                                     */
  const double* size;              /* [N+L]                                                      */
  const uint8_t* collide;          /* [N+L]                                                      */
  const int32_t* type;             /* [N+L]                                                      */
  const double* mass;              /* [N]                                                        */
  const double* accel;             /* [N]                                                        */
  const double* max_speed;         /* [N]                                                        */
  const double* slot_table;        /* POLYGON [N][2] unit offsets, LINE [N][2] ([k][0] fraction) */
} orc_config;

typedef struct orc_step_io {
  const void* actions;
  void* obs;
  int32_t* nbr_idx;
  void* nbr_feat;
  int32_t* nbr_cnt;
  uint32_t* adj;
  void* reward;
  void* cost;
  uint8_t* done;
  int32_t* assign;
} orc_step_io;

#endif
