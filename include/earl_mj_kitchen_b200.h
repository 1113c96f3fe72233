/*
 * earl_mj_kitchen_b200.h -- ENGINE-LEVEL entry points of the kitchen capacity set of libearl_b200.so (round 1).
 *
 * The reference's kitchen step is KitchenV0.step -> Robot.step -> MujocoEnv.do_simulation =
 * 40 x sim.step() (kitchen_assets/adept_envs/adept_envs/franka/kitchen_multitask_v0.py:91-125,
 * adept_envs/mujoco_env.py:148-153) on franka_kitchen_jntpos_act_ab.xml: 23 dofs, 118 colliding geoms, friction loss on
 * every dof, 5 joint equalities, pyramidal friction cones.  What is declared here replaces the `n x sim.step()` part on
 * caller-held state arrays, one warp per environment instance, so that the device engine of this capacity set can be
 * checked against the fp64 checker and measured.  The TASK-LEVEL kitchen ABI (control from the noisy observation,
 * observation noise, reward) is not built yet: EARLEnvs('kitchen') stays unavailable until it is.
 * Same conventions as include/earl_b200.h; all arrays are DEVICE pointers; no CPU fallback.
 */
#ifndef EARL_MJ_KITCHEN_B200_H_
#define EARL_MJ_KITCHEN_B200_H_

#include "earl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct earl_mjk_engine earl_mjk_engine;

/* model_blob: earl_benchmark_b200/models/kitchen.npz serialized by Model.to_blob() (blob version 2).
 * Replaces mujoco_py.load_model_from_path + MjSim (adept_envs/simulation/sim_robot.py:67-69). */
EARL_API int earl_mjk_engine_create(const void* model_blob, size_t model_nbytes, int32_t device, earl_mjk_engine** out);
EARL_API int earl_mjk_engine_destroy(earl_mjk_engine* e);
EARL_API int earl_mjk_engine_nv(const earl_mjk_engine* e);

/* `nsub` x mj_step for every environment (sim.step(), adept_envs/mujoco_env.py:152-153).
 * In / out: qpos, qvel, qacc_warmstart f32 [N, nv].  In: mocap_pos f64 [N,3], mocap_quat f32 [4] (one for all: the task
 * never moves it), ctrl f32 [N,2] (the engine clamps to ctrlrange).  Out: info i32 [N,4] = { rows of the last substep,
 * contacts of the last substep, Newton iterations summed over the substeps, flags (bit 0 non-positive pivot, bits 1-3
 * capacity overflow: candidate pairs / contacts / rows) }. */
EARL_API int earl_mjk_engine_substeps(earl_mjk_engine* e, int32_t num_envs, int32_t nsub, float* qpos_dev, float* qvel_dev,
                                      float* warm_dev, const double* mocap_pos_dev, const float* mocap_quat_host,
                                      const float* ctrl_dev, int32_t* info_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EARL_MJ_KITCHEN_B200_H_ */
