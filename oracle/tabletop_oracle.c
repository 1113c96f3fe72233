/*
 * oracle/tabletop_oracle.c -- CPU restatement of the reference's tabletop hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the CHECKER and the timed CPU baseline; the product
 * package (earl_benchmark_b200/) never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against outputs
 * of the unmodified reference run behind a fake gym backend (oracle/gen_golden.py ->
 * tests/golden/tabletop_ref_*.npz) and against the 2,534 shipped demonstration transitions.
 *
 * Reference code restated (paths relative to /root/reference):
 *   earl_benchmark/envs/tabletop_manipulation.py:128-138   step   (clip + rescale action)
 *   earl_benchmark/envs/tabletop_manipulation.py:140-174   move   (attach / drag / clip)
 *   earl_benchmark/envs/tabletop_manipulation.py:55-60     _get_obs
 *   earl_benchmark/envs/tabletop_manipulation.py:176-204   compute_reward / is_successful
 *   earl_benchmark/wrappers/persistent_state_wrapper.py:17-31   reset / step counters + horizon
 *   earl_benchmark/wrappers/lifelong_wrapper.py:30-44      lifelong return + goal swap cadence
 *
 * Floating point: the reference keeps qpos in fp64, does the step arithmetic in fp64 (np.clip of an
 * fp32 action against fp64 bounds promotes to fp64), rounds ONCE to fp32 for the observation and
 * evaluates the sparse success norm in fp32 on that observation, comparing with 0.2 in fp64
 * (numpy 1.22.2 scalar promotion, the reference's pin).  np.linalg.norm(x) = sqrt(dot(x,x)); the fp64
 * 2-element dot of the attach test is evaluated as fma(dy,dy,dx*dx), which is what the OpenBLAS behind
 * the live reference run in the build container does (bit-verified there); it can matter only within
 * one fp64 ulp of the 0.4 attach radius.  Build with -ffp-contract=off so nothing else is fused.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define TT_THRESHOLD 0.4     /* self.threshold, tabletop_manipulation.py:42 */
#define TT_CLIP 2.8          /* np.clip(..., -2.8, 2.8), :157,:163 */
#define TT_SUCCESS 0.2       /* is_successful, :202,:204 */

static inline double clipd(double x, double lo, double hi) {
  /* np.clip == minimum(maximum(x, lo), hi); NaN propagates */
  if (x != x) return x;
  return x < lo ? lo : (x > hi ? hi : x);
}

/* np.linalg.norm on an fp32 vector: fp32 products, accumulated in index order in fp64 (OpenBLAS sdot keeps a
 * double accumulator: checked against numpy in tests/test_oracle_golden.py), rounded to fp32, fp32 sqrt */
static inline float norm_f32(const float *d, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) {
    float p = d[i] * d[i];
    s += (double)p;
  }
  return sqrtf((float)s);
}

/* test hook: the norm restatement on caller data */
float earl_oracle_norm_f32(const float *d, int n) { return norm_f32(d, n); }

/* tabletop_manipulation.py:55-60 */
static inline void tt_obs(const double *qpos, int attached, const double *goal, float *obs) {
  for (int i = 0; i < 4; ++i) obs[i] = (float)qpos[i];
  obs[4] = obs[5] = attached ? 0.0f : -1.0f;   /* np.asarray(self.attached_object) */
  for (int i = 0; i < 6; ++i) obs[6 + i] = (float)goal[i];
}

/* tabletop_manipulation.py:197-204 (obs[6:-2] is goal[0:4], obs[8:-2] is goal[2:4]) */
static inline int tt_success(const float *obs, int wide) {
  float d[4];
  float nrm;
  if (wide) {
    d[0] = obs[2] - obs[8];
    d[1] = obs[3] - obs[9];
    nrm = norm_f32(d, 2);
  } else {
    for (int i = 0; i < 4; ++i) d[i] = obs[i] - obs[6 + i];
    nrm = norm_f32(d, 4);
  }
  return (double)nrm <= TT_SUCCESS;
}

/* tabletop_manipulation.py:176-191 */
static inline double tt_reward(const float *obs, int dense, int wide) {
  if (!dense) return (double)tt_success(obs, wide);
  float d[2] = {obs[2] - obs[8], obs[3] - obs[9]};
  float n1 = norm_f32(d, 2);
  double reward = (double)(-n1);
  reward += 2.0 * exp(-((double)n1 * (double)n1) / 0.01);
  float e[2] = {obs[0] - obs[2], obs[1] - obs[3]};
  double grip_to_object = 0.5 * (double)norm_f32(e, 2);
  reward += -grip_to_object;
  reward += 0.5 * exp(-(grip_to_object * grip_to_object) / 0.01);
  return reward;
}

/* One env, one step.  qpos/attached are updated in place. */
static inline void tt_step_one(double *qpos, int32_t *attached, const double *goal, const float *action,
                               int dense, int wide, int state_f32, float *obs, double *reward,
                               uint8_t *success) {
  /* step(): clip to [-1,1], then lb + (a + 1) * 0.5 * (ub - lb) with lb=-0.2, ub=0.2 (:130-132) */
  double a[3];
  const double span = 0.2 - (-0.2);
  for (int i = 0; i < 3; ++i) {
    double c = clipd((double)action[i], -1.0, 1.0);
    a[i] = -0.2 + (c + 1.0) * 0.5 * span;
  }
  /* move() (:140-174) */
  const double fx = qpos[0], fy = qpos[1];
  if (a[2] > 0) {
    if (!*attached) {
      double dx = fx - qpos[2], dy = fy - qpos[3];
      double dist = sqrt(fma(dy, dy, dx * dx));
      if (dist < TT_THRESHOLD) *attached = 1;   /* single object: "closest" is that object */
    }
  } else {
    *attached = 0;
  }
  double nfx = clipd(fx + a[0], -TT_CLIP, TT_CLIP);
  double nfy = clipd(fy + a[1], -TT_CLIP, TT_CLIP);
  if (*attached) {
    qpos[2] = clipd(qpos[2] + (nfx - fx), -TT_CLIP, TT_CLIP);
    qpos[3] = clipd(qpos[3] + (nfy - fy), -TT_CLIP, TT_CLIP);
  }
  qpos[0] = nfx;
  qpos[1] = nfy;
  tt_obs(qpos, *attached, goal, obs);
  if (state_f32) /* model of a device state kept in fp32: the stored state IS the observation */
    for (int i = 0; i < 4; ++i) qpos[i] = (double)obs[i];
  *reward = tt_reward(obs, dense, wide);
  if (success) *success = (uint8_t)tt_success(obs, wide);
}

/* ------------------------------------------------------------------ exported batch entry points */

/* TabletopManipulation.step over n independent envs. */
void earl_oracle_tt_step(int64_t n, double *qpos /*[n,4]*/, int32_t *attached /*[n]*/,
                         const double *goal /*[n,6]*/, const float *action /*[n,3]*/, int dense, int wide,
                         int state_f32, float *obs /*[n,12]*/, double *reward /*[n]*/,
                         uint8_t *success /*[n]*/) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    tt_step_one(qpos + 4 * i, attached + i, goal + 6 * i, action + 3 * i, dense, wide, state_f32,
                obs + 12 * i, reward + i, success ? success + i : 0);
}

/* compute_reward / is_successful on given observations. */
void earl_oracle_tt_reward(int64_t n, const float *obs /*[n,12]*/, int dense, int wide, double *reward,
                           uint8_t *success) {
  for (int64_t i = 0; i < n; ++i) {
    reward[i] = tt_reward(obs + 12 * i, dense, wide);
    success[i] = (uint8_t)tt_success(obs + 12 * i, wide);
  }
}

/* PersistentStateWrapper.step bookkeeping (persistent_state_wrapper.py:22-31). env_done may be NULL. */
void earl_oracle_psw_step(int64_t n, int64_t *total_steps, int64_t *steps_since_reset, int64_t horizon,
                          const uint8_t *env_done, uint8_t *done) {
  for (int64_t i = 0; i < n; ++i) {
    int d = env_done ? env_done[i] : 0;
    total_steps[i] += 1;
    steps_since_reset[i] += 1;
    if (!d && steps_since_reset[i] >= horizon) d = 1;
    done[i] = (uint8_t)d;
  }
}

/* PersistentStateWrapper.reset bookkeeping (:17-20) for envs with mask != 0 (mask NULL = all). */
void earl_oracle_psw_reset(int64_t n, const uint8_t *mask, int64_t *steps_since_reset,
                           int64_t *num_interventions) {
  for (int64_t i = 0; i < n; ++i)
    if (!mask || mask[i]) {
      num_interventions[i] += 1;
      steps_since_reset[i] = 0;
    }
}

/* LifelongWrapper.step bookkeeping (lifelong_wrapper.py:30-44): returns swap[i]=1 where the caller
 * must now draw a new goal and re-read the observation (the reward stays the pre-swap one). */
void earl_oracle_lifelong_step(int64_t n, const double *reward, double *lifelong_return,
                               int64_t *steps_since_goal_change, int64_t goal_change_frequency,
                               uint8_t *swap) {
  for (int64_t i = 0; i < n; ++i) {
    steps_since_goal_change[i] += 1;
    lifelong_return[i] += reward[i];
    swap[i] = 0;
    if (steps_since_goal_change[i] >= goal_change_frequency) {
      steps_since_goal_change[i] = 0;
      swap[i] = 1;
    }
  }
}

/* Whole hot path (env step + PersistentStateWrapper) for `steps` consecutive steps over n envs;
 * actions cycle through a ring of `ring` batches [ring,n,3].  Used as the timed CPU baseline.
 * Outputs of the LAST step are left in obs/reward/done.  Threads: OpenMP over envs. */
void earl_oracle_tt_rollout(int64_t n, int64_t steps, int64_t ring, double *qpos, int32_t *attached,
                            const double *goal, const float *actions, int dense, int wide, int state_f32,
                            int64_t *total_steps, int64_t *steps_since_reset, int64_t horizon, float *obs,
                            double *reward, uint8_t *done) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    for (int64_t t = 0; t < steps; ++t) {
      const float *a = actions + ((t % ring) * n + i) * 3;
      tt_step_one(qpos + 4 * i, attached + i, goal + 6 * i, a, dense, wide, state_f32, obs + 12 * i,
                  reward + i, 0);
      total_steps[i] += 1;
      steps_since_reset[i] += 1;
      done[i] = (uint8_t)(steps_since_reset[i] >= horizon);
    }
  }
}

/* Same work, but step-major (all envs advance one step before the next), i.e. the memory access
 * pattern of a batched vector env; the fairer CPU counterpart of the GPU path. */
void earl_oracle_tt_rollout_stepmajor(int64_t n, int64_t steps, int64_t ring, double *qpos,
                                      int32_t *attached, const double *goal, const float *actions, int dense,
                                      int wide, int state_f32, int64_t *total_steps,
                                      int64_t *steps_since_reset, int64_t horizon, float *obs, double *reward,
                                      uint8_t *done) {
  for (int64_t t = 0; t < steps; ++t) {
    const float *abase = actions + (t % ring) * n * 3;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      tt_step_one(qpos + 4 * i, attached + i, goal + 6 * i, abase + 3 * i, dense, wide, state_f32,
                  obs + 12 * i, reward + i, 0);
      total_steps[i] += 1;
      steps_since_reset[i] += 1;
      done[i] = (uint8_t)(steps_since_reset[i] >= horizon);
    }
  }
}

int earl_oracle_abi_version(void) { return 1; }
