/*
 * earl_tt3_b200.h -- C ABI of the three-object tabletop task inside libearl_b200.so.
 *
 * Replaces `earl_benchmark/envs/tabletop_manipulation_3obj.py` (class TabletopManipulation, the variant with
 * three draggable objects and closest-object attach) under a PersistentStateWrapper
 * (`earl_benchmark/wrappers/persistent_state_wrapper.py:17-45`).  The reference has no FFI on this path;
 * each entry point cites the Python method it replaces.  Conventions are those of earl_b200.h (plain C,
 * 0 = ok / <0 = earl_status, earl_last_error(), device pointers unless the name ends in _host, work is
 * only ENQUEUED on `stream`, no CPU fallback).
 *
 * Per-env state: qpos[0:8] as fp64 (fist x,y, object A x,y, object B x,y, object C x,y) like the
 * reference keeps it in MuJoCo's qpos, so open-loop rollouts of any length are bit-exact; one u32 of flags
 * (bits 0-1 attached object: 0 none, 1..3 = object_dict entries (0,0) / (0.5,0.5) / (1,1),
 * tabletop_manipulation_3obj.py:31-35; bits 8-15 goal row), u32 steps_since_reset, i64 num_interventions.
 * Observation [N,20] f32 = qpos[0:8], attached marker x2, goal[0:10]  (:49-54).
 */
#ifndef EARL_TT3_B200_H_
#define EARL_TT3_B200_H_

#include "earl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct earl_tt3_handle earl_tt3_handle;

#define EARL_TT3_MAX_GOALS 16

typedef struct {
  int32_t num_envs;          /* N >= 1 */
  int32_t device;            /* CUDA device ordinal */
  uint32_t flags;            /* EARL_FLAG_DENSE_REWARD | EARL_FLAG_AUTO_RESET (earl_b200.h) */
  int32_t num_goals;         /* rows used in goal_table, 1..EARL_TT3_MAX_GOALS (reference: 1, :12-18) */
  int64_t episode_horizon;   /* PersistentStateWrapper(episode_horizon) */
  double threshold;          /* 0.4  attach radius, strict <            (:38, :107) */
  double move_distance;      /* 0.2  action scale                       (:39, :88-89) */
  double clip;               /* 2.8  workspace clip                     (:113, :119) */
  double success_radius;     /* 0.4  sparse success, <=                 (:165) */
  double initial_state[10];  /* initial_states[0]                       (:11) */
  double goal_table[EARL_TT3_MAX_GOALS][10]; /* goal_states             (:12-18) */
} earl_tt3_config;

/* construct env + wrapper (tabletop_manipulation_3obj.py:25-47, persistent_state_wrapper.py:10-15) */
EARL_API int earl_tt3_create(const earl_tt3_config* cfg, size_t cfg_nbytes, earl_tt3_handle** out);
EARL_API int earl_tt3_destroy(earl_tt3_handle* h);

/* PersistentStateWrapper.reset + env.reset (:17-20; tabletop_manipulation_3obj.py:67-84) for envs with
 * mask[i] != 0 (NULL = all): num_interventions += 1, steps_since_reset = 0, attached cleared,
 * goal row <- goal_idx[i] (NULL = row 0), qpos[0:8] <- init_qpos[i] when given (the host computes
 * goal[:8] + uniform(-0.3, 0.3) for reset_at_goal, :72-76) else initial_state[0:8].
 * Writes the post-reset observation rows of the masked envs to obs_out when non-NULL. */
EARL_API int earl_tt3_reset(earl_tt3_handle* h, const uint8_t* mask_dev, const int32_t* goal_idx_dev,
                            const double* init_qpos_dev /*[N,8]*/, float* obs_out_dev /*[N,20]*/, void* stream);

/* env.reset_goal(goal) (:61-64): goal row per env (NULL = row 0 for every masked env) */
EARL_API int earl_tt3_set_goal(earl_tt3_handle* h, const uint8_t* mask_dev, const int32_t* goal_idx_dev, void* stream);

/* The hot path: PersistentStateWrapper.step (:22-31) over env.step + move + _get_obs + compute_reward
 * (tabletop_manipulation_3obj.py:86-165).  actions [N,3] f32 in; obs [N,20] f32, reward [N] f32,
 * done [N] u8, success [N] u8 (nullable) out. */
EARL_API int earl_tt3_step(earl_tt3_handle* h, const float* actions_dev, float* obs_dev, float* reward_dev,
                           uint8_t* done_dev, uint8_t* success_dev, void* stream);

/* `num_steps` steps back to back: step t reads actions[t % action_ring], writes output slot t % out_ring */
EARL_API int earl_tt3_rollout(earl_tt3_handle* h, const float* actions_dev, int32_t action_ring, int32_t num_steps,
                              float* obs_dev, float* reward_dev, uint8_t* done_dev, uint8_t* success_dev,
                              int32_t out_ring, void* stream);

/* same step with HOST buffers (copies inside the call, synchronous): the call a CPU-side RL loop makes */
EARL_API int earl_tt3_step_host(earl_tt3_handle* h, const float* actions_host, float* obs_host, float* reward_host,
                                uint8_t* done_host, uint8_t* success_host);

/* env._get_obs() (:49-54) */
EARL_API int earl_tt3_get_obs(earl_tt3_handle* h, float* obs_dev, void* stream);

/* env.compute_reward(obs) / env.is_successful(obs) (:146-165) on caller observations [M,20] */
EARL_API int earl_tt3_compute_reward(earl_tt3_handle* h, const float* obs_dev, int64_t num_obs, float* reward_dev,
                                     uint8_t* success_dev, void* stream);

/* total_steps / num_interventions / steps_since_reset (persistent_state_wrapper.py:39-45) */
EARL_API int earl_tt3_counters(earl_tt3_handle* h, int64_t* total_steps_host, int64_t* num_interventions_dev,
                               uint32_t* steps_since_reset_dev, void* stream);

/* env.sim.data.qpos[:8] and the attached object (0 none, 1..3); either pointer may be NULL */
EARL_API int earl_tt3_get_state(earl_tt3_handle* h, double* qpos_dev /*[N,8]*/, int32_t* attached_dev, void* stream);
EARL_API int earl_tt3_set_state(earl_tt3_handle* h, const double* qpos_dev /*[N,8]*/, const int32_t* attached_dev,
                                void* stream);

EARL_API int64_t earl_tt3_launch_count(const earl_tt3_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* EARL_TT3_B200_H_ */
