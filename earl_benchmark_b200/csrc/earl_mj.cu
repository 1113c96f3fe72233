// earl_mj.cu -- C-ABI implementation (include/earl_mj_b200.h) of the batched Sawyer-task step: one warp per
// environment instance, all frame_skip substeps of an env step inside ONE launch so the 256-byte state record makes
// one HBM round trip per env step.  Device code: mj_engine.cuh / mj_collide.cuh / mj_step.cuh.  No CPU fallback.
#include "../../include/earl_mj_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "mj_model_host.hpp"
#include "mj_step.cuh"

namespace earl {
int set_error(int code, const char* msg);  // earl_b200.cu: thread-local message behind earl_last_error()
}

namespace {

int failf(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return earl::set_error(code, buf);
}

#define CU(call)                                                                                                 \
  do {                                                                                                           \
    cudaError_t e_ = (call);                                                                                     \
    if (e_ != cudaSuccess)                                                                                       \
      return failf(EARL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);    \
  } while (0)

using namespace earl::mj;

constexpr int kWPB = 16;  // warps (= environments in flight) per block: 16 x 13.2 KB workspaces + the model fill one SM
constexpr int kObs = 14, kAct = 4, kGoal = 7, kMaxGoals = 32;
constexpr int kBuckets = 32;
constexpr size_t kModelBytes = (sizeof(Model) + 15) & ~size_t(15);
constexpr size_t kSmemBytes = kModelBytes + kWPB * ((sizeof(Work) + 15) & ~size_t(15));
constexpr size_t kWorkStride = (sizeof(Work) + 15) & ~size_t(15);

struct StepArgs {
  const Model* model;
  const float* hull;
  float* state;              // [N][REC_FLOATS]
  const float* goals;        // [kMaxGoals][8]
  long long* interventions;  // [N]
  double* ep_return;         // [N] or null
  double* ll_return;         // [N] or null (lifelong handles)
  unsigned* ll_steps;        // [N] steps_since_goal_change
  unsigned goal_freq;
  unsigned long long* work;  // 6 counters
  // cost-sorted scheduling: environments are visited in the order of `order_cur` in chunks of kWPB grabbed from an
  // atomic counter; every env files itself under one of kBuckets cost buckets (estimated from the work it just did)
  // and a tiny second kernel turns (bucket, rank) into next step's order, most expensive first
  const int* order_cur;
  int* order_next;
  unsigned* sched;           // [0] next chunk, [1 .. kBuckets] bucket counters
  unsigned char* env_bucket; // [N]
  unsigned* env_rank;        // [N]
  int bucket_width;          // estimated warp instructions per cost bucket
  int n;
  unsigned horizon;
  unsigned flags;
  // step
  const float* actions;
  float* obs;
  float* reward;
  uint8_t* done;
  uint8_t* success;
  // reset
  const float* tmpl;         // [REC_FLOATS]
  const uint8_t* mask;
  const double* obj_qpos;
  const int* goal_idx;
  int obj_qadr, obj_dadr, obj_nq_set, obj_nv;
  // settle
  double hand_init[3];
  float ctrl[2];
  int steps;
};

__device__ __forceinline__ void load_model(Model* sm, const Model* gm) {
  const uint32_t* src = reinterpret_cast<const uint32_t*>(gm);
  uint32_t* dst = reinterpret_cast<uint32_t*>(sm);
  for (unsigned i = threadIdx.x; i < sizeof(Model) / 4; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

__device__ __forceinline__ void scatter_rec(Work& w, int idx, float v) {
  if (idx < REC_QVEL) w.qpos[idx] = v;
  else if (idx < REC_WARM) w.qvel[idx - REC_QVEL] = v;
  else if (idx < REC_MOCAP) w.warm[idx - REC_WARM] = v;
  else if (idx < REC_STEPS) reinterpret_cast<float*>(w.mocap_pos)[idx - REC_MOCAP] = v;
  else if (idx == REC_STEPS) w.steps = __float_as_uint(v);
  else if (idx == REC_FLAGS) w.flags = __float_as_uint(v);
  else if (idx == REC_GOALROW) w.goalrow = __float_as_uint(v);
}
__device__ __forceinline__ float gather_rec(const Work& w, int idx) {
  if (idx < REC_QVEL) return w.qpos[idx];
  if (idx < REC_WARM) return w.qvel[idx - REC_QVEL];
  if (idx < REC_MOCAP) return w.warm[idx - REC_WARM];
  if (idx < REC_STEPS) return reinterpret_cast<const float*>(w.mocap_pos)[idx - REC_MOCAP];
  if (idx == REC_STEPS) return __uint_as_float(w.steps);
  if (idx == REC_FLAGS) return __uint_as_float(w.flags);
  if (idx == REC_GOALROW) return __uint_as_float(w.goalrow);
  return 0.0f;
}
__device__ __forceinline__ void load_env(Work& w, const float* rec, int lane) {
  scatter_rec(w, lane, rec[lane]);
  scatter_rec(w, lane + 32, rec[lane + 32]);
  if (lane == 0) { w.bad = 0; w.acc_iter = w.acc_rows = w.acc_con = w.acc_mpr = w.acc_sup = 0; }
#ifdef MJ_PHASE_TIMING
  if (lane < 8) w.phase[lane] = 0;
#endif
  __syncwarp();
}
__device__ __forceinline__ void store_env(const Work& w, float* rec, int lane) {
  __syncwarp();
  rec[lane] = gather_rec(w, lane);
  rec[lane + 32] = gather_rec(w, lane + 32);
}

// sparse success of the current observation against the current goal (sawyer_door.py:173-177, sawyer_peg.py:301-305)
__device__ __forceinline__ bool obs_success(const Model& m, const Work& w, const float* goals) {
  const float* g = goals + 8 * w.goalrow;
  const float dx = w.obs7[4] - g[4], dy = w.obs7[5] - g[5], dz = w.obs7[6] - g[6];
  return sqrtf(dx * dx + dy * dy + dz * dz) <= m.success_radius;
}
// observation row [hand(3), gripper(1), object(3), goal(7)] (sawyer_door.py:86-94)
__device__ __forceinline__ void write_obs_row(const Work& w, const float* goals, float* obs_row, int lane) {
  const float* g = goals + 8 * w.goalrow;
  if (obs_row && lane < kObs) obs_row[lane] = lane < 7 ? w.obs7[lane] : g[lane - 7];
}
__device__ __forceinline__ bool write_obs(const Model& m, Work& w, const float* goals, float* obs_row, int lane) {
  if (lane == 0) observe(m, w, w.obs7);
  __syncwarp();
  write_obs_row(w, goals, obs_row, lane);
  return obs_success(m, w, goals);
}

__global__ void __launch_bounds__(kWPB * 32, 1) mj_step_kernel(const StepArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  Model* sm = reinterpret_cast<Model*>(smem);
  load_model(sm, a.model);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Work& w = *reinterpret_cast<Work*>(smem + kModelBytes + warp * kWorkStride);
  unsigned long long it = 0, rows = 0, cons = 0, bad = 0, over = 0, envs = 0, ov_hit = 0, ov_con = 0, ov_row = 0;
  // Every warp of a block makes the same trips (the substep has block-wide phase barriers); a warp without an
  // environment in the last chunk re-runs another environment and discards the result.  Chunks of kWPB consecutive
  // entries of the cost-sorted order are handed out dynamically, so blocks that draw expensive chunks (gripper on
  // the handle: MPR solves, more Newton iterations) take fewer of them and cheap environments are not held back at
  // the phase barriers by an expensive neighbour.
  __shared__ int s_chunk;
  const int nchunks = (a.n + kWPB - 1) / kWPB;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_chunk = (int)atomicAdd(&a.sched[0], 1u);
    __syncthreads();
    const int base = s_chunk * kWPB;
    if (s_chunk >= nchunks) break;
    const bool live = base + warp < a.n;
    const int env = a.order_cur[live ? base + warp : base];  // idle warps shadow the chunk's first env (same barriers, no store)
    float* rec = a.state + (size_t)env * REC_FLOATS;
    load_env(w, rec, lane);
    if (lane < kAct) w.action[lane] = a.actions[(size_t)env * kAct + lane];
    __syncwarp();
    env_step<32>(*sm, a.hull, w, w.action, lane);
    if (!live) continue;
    if (lane == 0) observe(*sm, w, w.obs7);
    __syncwarp();
    const bool ok = obs_success(*sm, w, a.goals);  // reward / success against the goal the step was taken with
    float r = ok ? 1.0f : 0.0f;
    if (a.flags & EARL_FLAG_DENSE_REWARD) r = door_dense_reward(*sm, w.obs7, a.goals + 8 * w.goalrow + 4);
    if (lane == 0 && a.ll_return) {
      // LifelongWrapper.step: lifetime return, periodic reset_goal() (single-goal tasks: goal_states[0] = row 0)
      a.ll_return[env] += (double)r;
      unsigned s = a.ll_steps[env] + 1;
      if (s >= a.goal_freq) { s = 0; w.goalrow = 0; }
      a.ll_steps[env] = s;
    }
    __syncwarp();
    write_obs_row(w, a.goals, a.obs + (size_t)env * kObs, lane);
    if (lane == 0) {
      // PersistentStateWrapper.step: counters, horizon `done` (persistent_state_wrapper.py:22-31)
      const unsigned steps = w.steps == 0xffffffffu ? w.steps : w.steps + 1;
      w.steps = steps;
      w.flags = (w.flags & ~2u) | (ok ? 3u : 0u) | ((w.bad & 1) ? 4u : 0u) | ((w.bad & 14) ? 8u : 0u);
      a.reward[env] = r;
      a.done[env] = steps >= a.horizon ? 1 : 0;
      if (a.success) a.success[env] = ok ? 1 : 0;
      if (a.ep_return) a.ep_return[env] += (double)r;
      it += w.acc_iter; rows += w.acc_rows; cons += w.acc_con; bad += (w.bad & 1) ? 1 : 0; over += (w.bad & 14) ? 1 : 0; envs += 1;
      ov_hit += (w.bad & 2) ? 1 : 0; ov_con += (w.bad & 4) ? 1 : 0; ov_row += (w.bad & 8) ? 1 : 0;
      // next step's visiting order: estimated warp instructions of this env step above the contact-free baseline
      // (one extra Newton iteration ~2.5k, a contact ~0.3k, a support-function call ~0.3k), in 32 buckets of 3k
      const int est = 2500 * (w.acc_iter - sm->frame_skip) + 300 * (w.acc_con - 4 * sm->frame_skip) + 300 * w.acc_sup;
      int b = est <= 0 ? 0 : 1 + est / a.bucket_width;
      b = b > kBuckets - 1 ? kBuckets - 1 : b;
      a.env_bucket[env] = (unsigned char)b;
      a.env_rank[env] = atomicAdd(&a.sched[1 + b], 1u);
#ifdef MJ_PHASE_TIMING
      for (int k = 0; k < 8; ++k) atomicAdd(&a.work[8 + k], (unsigned long long)w.phase[k]);
#endif
    }
    store_env(w, rec, lane);
    __syncwarp();
  }
  if (lane == 0 && envs) {
    atomicAdd(&a.work[0], envs);
    atomicAdd(&a.work[1], envs * (unsigned long long)sm->frame_skip);
    atomicAdd(&a.work[2], it);
    atomicAdd(&a.work[3], rows);
    atomicAdd(&a.work[4], cons);
    atomicAdd(&a.work[5], bad);
    atomicAdd(&a.work[6], over);
    atomicAdd(&a.work[7], ov_hit);
    atomicAdd(&a.work[16], ov_con);
    atomicAdd(&a.work[17], ov_row);
  }
}

// reset (mode 0) / get_obs (mode 1): fresh kinematics of the (new) state, observation out
__global__ void __launch_bounds__(kWPB * 32, 1) mj_reset_kernel(const StepArgs a, const int mode) {
  extern __shared__ __align__(16) unsigned char smem[];
  Model* sm = reinterpret_cast<Model*>(smem);
  load_model(sm, a.model);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Work& w = *reinterpret_cast<Work*>(smem + kModelBytes + warp * kWorkStride);
  for (int env = blockIdx.x * kWPB + warp; env < a.n; env += gridDim.x * kWPB) {
    if (mode == 0 && a.mask && !a.mask[env]) continue;
    float* rec = a.state + (size_t)env * REC_FLOATS;
    load_env(w, mode == 0 ? a.tmpl : rec, lane);
    if (mode == 0 && lane == 0) {
      if (a.obj_qpos) {  // _set_obj_xyz: leading qpos entries of the object joint, zero velocity on all its dofs
        for (int k = 0; k < a.obj_nq_set; ++k) w.qpos[a.obj_qadr + k] = (float)a.obj_qpos[(size_t)env * a.obj_nq_set + k];
        for (int k = 0; k < a.obj_nv; ++k) w.qvel[a.obj_dadr + k] = 0.0f;
      }
      w.goalrow = a.goal_idx ? (unsigned)a.goal_idx[env] : 0u;
      w.steps = 0;
      w.flags = 0;
      a.interventions[env] += 1;  // PersistentStateWrapper.reset (persistent_state_wrapper.py:17-20)
      if (a.ep_return) a.ep_return[env] = 0.0;
      if (a.ll_steps) a.ll_steps[env] = 0;  // LifelongWrapper.reset (lifelong_wrapper.py:25-28); the return is kept
    }
    __syncwarp();
    kinematics<32>(*sm, w, lane);
    write_obs(*sm, w, a.goals, a.obs ? a.obs + (size_t)env * kObs : nullptr, lane);
    if (mode == 0) store_env(w, rec, lane);
    __syncwarp();
  }
}

// sim.reset() + _reset_hand(steps) for ONE environment -> reset template record
__global__ void __launch_bounds__(32) mj_settle_kernel(const StepArgs a, float* tmpl_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  Model* sm = reinterpret_cast<Model*>(smem);
  load_model(sm, a.model);
  const int lane = threadIdx.x & 31;
  Work& w = *reinterpret_cast<Work*>(smem + kModelBytes);
  if (lane == 0) {
    for (int k = 0; k < MAXQ; ++k) w.qpos[k] = k < sm->nq ? sm->qpos0[k] : 0.0f;
    for (int k = 0; k < MAXV; ++k) { w.qvel[k] = 0; w.warm[k] = 0; }
    w.steps = 0; w.flags = 0; w.goalrow = 0; w.bad = 0; w.acc_iter = w.acc_rows = w.acc_con = 0;
  }
  __syncwarp();
  for (int s = 0; s < a.steps; ++s) {
    if (lane == 0) {
      for (int k = 0; k < 3; ++k) w.mocap_pos[k] = a.hand_init[k];
      w.mocap_quat[0] = 1; w.mocap_quat[1] = 0; w.mocap_quat[2] = 1; w.mocap_quat[3] = 0;
      w.ctrl[0] = a.ctrl[0]; w.ctrl[1] = a.ctrl[1];
    }
    __syncwarp();
    for (int k = 0; k < sm->frame_skip; ++k) substep<32>(*sm, a.hull, w, lane);
  }
  store_env(w, tmpl_out, lane);
}

// (bucket, rank within bucket) -> position in next step's visiting order, most expensive bucket first
__global__ void mj_order_kernel(const StepArgs a) {
  __shared__ unsigned base[kBuckets];
  if (threadIdx.x == 0) {
    unsigned acc = 0;
    for (int b = kBuckets - 1; b >= 0; --b) { base[b] = acc; acc += a.sched[1 + b]; }
  }
  __syncthreads();
  for (int env = blockIdx.x * blockDim.x + threadIdx.x; env < a.n; env += gridDim.x * blockDim.x)
    a.order_next[base[a.env_bucket[env]] + a.env_rank[env]] = env;
}

__global__ void mj_eval_stats_kernel(const float* state, const double* ep_return, int n, double* out4) {
  double ret = 0, last = 0, any = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned fl = __float_as_uint(state[(size_t)i * REC_FLOATS + REC_FLAGS]);
    ret += ep_return[i];
    last += (fl & 2u) ? 1.0 : 0.0;
    any += (fl & 1u) ? 1.0 : 0.0;
  }
  for (int o = 16; o > 0; o >>= 1) {
    ret += __shfl_xor_sync(0xffffffffu, ret, o);
    last += __shfl_xor_sync(0xffffffffu, last, o);
    any += __shfl_xor_sync(0xffffffffu, any, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out4[0], ret);
    atomicAdd(&out4[1], last);
    atomicAdd(&out4[2], any);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&out4[3], (double)n);
}

}  // namespace

struct earl_mj_handle {
  earl_mj_config cfg{};
  HostModel hm;
  int device = 0, sm_count = 0, grid = 0;
  int obj_qadr = -1, obj_dadr = -1, obj_nv = 0;
  bool have_template = false;
  int64_t total_steps = 0, launches = 0;
  std::vector<void*> owned;
  StepArgs a{};
  float* d_tmpl = nullptr;
  float* d_goals = nullptr;
  int* d_order[2] = {nullptr, nullptr};
  unsigned* d_sched = nullptr;
  unsigned char* d_env_bucket = nullptr;
  unsigned* d_env_rank = nullptr;
  int order_sel = 0;
  int bucket_width = 3000;
  // host-path staging
  float* d_act = nullptr;
  float* d_obs = nullptr;
  float* d_rew = nullptr;
  uint8_t* d_done = nullptr;
  uint8_t* d_succ = nullptr;
  cudaStream_t host_stream = nullptr;

  template <typename T>
  int alloc(T** ptr, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 16);
    if (e != cudaSuccess) return failf(EARL_ERR_NOMEM, "cudaMalloc(%zu B) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    e = cudaMemset(q, 0, count * sizeof(T));
    if (e != cudaSuccess) return failf(EARL_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    owned.push_back(q);
    *ptr = static_cast<T*>(q);
    return 0;
  }
};

namespace {
int check_handle(const earl_mj_handle* h) {
  if (!h) return failf(EARL_ERR_INVALID, "null handle");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return failf(EARL_ERR_CUDA, "cudaSetDevice(%d) failed: %s", h->device, cudaGetErrorString(e));
  return 0;
}
int grid_for(const earl_mj_handle* h, int n) {
  const int blocks = (n + kWPB - 1) / kWPB;
  return blocks < h->grid ? blocks : h->grid;
}
}  // namespace

extern "C" {

int earl_mj_create(const earl_mj_config* cfg, const void* model_blob, size_t model_nbytes, const earl_mj_task* task,
                   earl_mj_handle** out) {
  if (!cfg || !out || !task || !model_blob) return failf(EARL_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->env_kind != EARL_ENV_SAWYER_DOOR && cfg->env_kind != EARL_ENV_SAWYER_PEG)
    return failf(EARL_ERR_UNSUPPORTED, "env_kind %d is not built on the articulated-body engine (sawyer_door, sawyer_peg)", cfg->env_kind);
  if (cfg->num_envs < 1) return failf(EARL_ERR_INVALID, "num_envs must be >= 1");
  if (cfg->episode_horizon < 1) return failf(EARL_ERR_INVALID, "episode_horizon must be >= 1");
  if (cfg->flags & ~(uint32_t)(EARL_FLAG_EVAL_STATS | EARL_FLAG_LIFELONG | EARL_FLAG_DENSE_REWARD))
    return failf(EARL_ERR_UNSUPPORTED, "unsupported flags %#x", cfg->flags);
  if ((cfg->flags & EARL_FLAG_DENSE_REWARD) && cfg->env_kind != EARL_ENV_SAWYER_DOOR)
    return failf(EARL_ERR_UNSUPPORTED, "the dense reward is only built for sawyer_door");
  if ((cfg->flags & EARL_FLAG_LIFELONG) && cfg->goal_change_frequency < 1)
    return failf(EARL_ERR_INVALID, "lifelong handles need goal_change_frequency >= 1");
  static_assert(sizeof(earl_mj_task) == sizeof(TaskSpec), "earl_mj_task must mirror earl::mj::TaskSpec");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return failf(EARL_ERR_INVALID, "device %d out of range (%d visible)", cfg->device, ndev);
  CU(cudaSetDevice(cfg->device));
  earl_mj_handle* h = new (std::nothrow) earl_mj_handle();
  if (!h) return failf(EARL_ERR_NOMEM, "host allocation failed");
  h->cfg = *cfg;
  h->device = cfg->device;
  TaskSpec ts;
  memcpy(&ts, task, sizeof(ts));
  std::string err;
  if (!build_model(model_blob, model_nbytes, ts, &h->hm, &err)) {
    delete h;
    return failf(EARL_ERR_INVALID, "%s", err.c_str());
  }
  const Model& m = h->hm.m;
  // the observed object: the joint of the body that carries it (door hinge)
  {
    const int b = ts.obj_geom >= 0 ? m.geom_body[ts.obj_geom] : m.site_body[ts.obj_site];
    if (b <= 0) { delete h; return failf(EARL_ERR_INVALID, "observed object is attached to the world"); }
    const int j = m.body_jnt[b];
    h->obj_qadr = m.jnt_qposadr[j];
    h->obj_dadr = m.jnt_dofadr[j];
    h->obj_nv = m.jnt_type[j] == 0 ? 6 : 1;
    const int nq_j = m.jnt_type[j] == 0 ? 7 : 1;
    if (ts.obj_qpos_count < 1 || ts.obj_qpos_count > nq_j) { delete h; return failf(EARL_ERR_INVALID, "obj_qpos_count %d does not fit the object joint", ts.obj_qpos_count); }
  }
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, cfg->device);
  if (e != cudaSuccess) { delete h; return failf(EARL_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
  h->sm_count = prop.multiProcessorCount;
  const size_t n = (size_t)cfg->num_envs;
  StepArgs& a = h->a;
  Model* d_model = nullptr;
  float* d_hull = nullptr;
  int rc = h->alloc(&d_model, 1);
  if (!rc) rc = h->alloc(&d_hull, h->hm.hull_vert.size() + 4);
  if (!rc) rc = h->alloc(&a.state, n * REC_FLOATS);
  if (!rc) rc = h->alloc(&h->d_goals, kMaxGoals * 8);
  if (!rc) rc = h->alloc(&a.interventions, n);
  if (!rc && (cfg->flags & EARL_FLAG_EVAL_STATS)) rc = h->alloc(&a.ep_return, n);
  if (!rc && (cfg->flags & EARL_FLAG_LIFELONG)) { rc = h->alloc(&a.ll_return, n); if (!rc) rc = h->alloc(&a.ll_steps, n); }
  if (!rc) rc = h->alloc(&a.work, 20);
  if (!rc) rc = h->alloc(&h->d_tmpl, REC_FLOATS);
  if (!rc) rc = h->alloc(&h->d_order[0], n);
  if (!rc) rc = h->alloc(&h->d_order[1], n);
  if (!rc) rc = h->alloc(&h->d_sched, 2 + kBuckets);
  if (!rc) rc = h->alloc(&h->d_env_bucket, n);
  if (!rc) rc = h->alloc(&h->d_env_rank, n);
  if (rc) { earl_mj_destroy(h); return rc; }
  {
    std::vector<int> ident(n);
    for (size_t k = 0; k < n; ++k) ident[k] = (int)k;
    e = cudaMemcpy(h->d_order[0], ident.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { earl_mj_destroy(h); return failf(EARL_ERR_CUDA, "order upload: %s", cudaGetErrorString(e)); }
  }
  e = cudaMemcpy(d_model, &m, sizeof(Model), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !h->hm.hull_vert.empty())
    e = cudaMemcpy(d_hull, h->hm.hull_vert.data(), h->hm.hull_vert.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mj_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mj_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mj_settle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  int per_sm = 0;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mj_step_kernel, kWPB * 32, kSmemBytes);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { earl_mj_destroy(h); return failf(EARL_ERR_CUDA, "engine setup: %s", cudaGetErrorString(e)); }
  if (per_sm < 1) { earl_mj_destroy(h); return failf(EARL_ERR_CUDA, "step kernel does not fit on an SM (%zu B shared memory)", kSmemBytes); }
  h->grid = h->sm_count * per_sm;
  if (const char* v = getenv("EARL_MJ_BUCKET_WIDTH")) { const int bw = atoi(v); if (bw >= 100) h->bucket_width = bw; }
  a.model = d_model;
  a.hull = d_hull;
  a.goals = h->d_goals;
  a.n = cfg->num_envs;
  a.horizon = cfg->episode_horizon > 0xffffffffLL ? 0xffffffffu : (unsigned)cfg->episode_horizon;
  a.flags = cfg->flags;
  a.goal_freq = cfg->goal_change_frequency > 0xffffffffLL ? 0xffffffffu : (unsigned)(cfg->goal_change_frequency > 0 ? cfg->goal_change_frequency : 1);
  a.obj_qadr = h->obj_qadr;
  a.obj_dadr = h->obj_dadr;
  a.obj_nq_set = ts.obj_qpos_count;
  a.obj_nv = h->obj_nv;
  *out = h;
  return 0;
}

int earl_mj_destroy(earl_mj_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  for (void* q : h->owned) cudaFree(q);
  if (h->host_stream) cudaStreamDestroy(h->host_stream);
  delete h;
  return 0;
}

int earl_mj_obs_dim(const earl_mj_handle* h) { return h ? kObs : 0; }
int earl_mj_action_dim(const earl_mj_handle* h) { return h ? kAct : 0; }
int earl_mj_nq(const earl_mj_handle* h) { return h ? h->hm.m.nq : 0; }
int earl_mj_nv(const earl_mj_handle* h) { return h ? h->hm.m.nv : 0; }
int64_t earl_mj_launch_count(const earl_mj_handle* h) { return h ? h->launches : 0; }

int earl_mj_set_goal_table(earl_mj_handle* h, const double* rows_host, int32_t count) {
  if (int rc = check_handle(h)) return rc;
  if (!rows_host || count < 1 || count > kMaxGoals) return failf(EARL_ERR_INVALID, "goal table must have 1..%d rows", kMaxGoals);
  std::vector<float> g(kMaxGoals * 8, 0.f);
  for (int r = 0; r < count; ++r)
    for (int c = 0; c < kGoal; ++c) g[r * 8 + c] = (float)rows_host[r * kGoal + c];
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(h->d_goals, g.data(), g.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

int earl_mj_build_reset_template(earl_mj_handle* h, const double* hand_init_pos_host, const float* ctrl_host, int32_t steps) {
  if (int rc = check_handle(h)) return rc;
  if (!hand_init_pos_host || !ctrl_host || steps < 0) return failf(EARL_ERR_INVALID, "bad reset-template arguments");
  StepArgs a = h->a;
  for (int k = 0; k < 3; ++k) a.hand_init[k] = hand_init_pos_host[k];
  a.ctrl[0] = ctrl_host[0];
  a.ctrl[1] = ctrl_host[1];
  a.steps = steps;
  mj_settle_kernel<<<1, 32, kSmemBytes>>>(a, h->d_tmpl);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  h->launches += 1;
  h->have_template = true;
  return 0;
}

int earl_mj_reset(earl_mj_handle* h, const uint8_t* mask_dev, const double* obj_qpos_dev, const int32_t* goal_idx_dev,
                  float* obs_out_dev, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!h->have_template) return failf(EARL_ERR_INVALID, "earl_mj_build_reset_template must run before the first reset");
  StepArgs a = h->a;
  a.tmpl = h->d_tmpl;
  a.mask = mask_dev;
  a.obj_qpos = obj_qpos_dev;
  a.goal_idx = goal_idx_dev;
  a.obs = obs_out_dev;
  mj_reset_kernel<<<grid_for(h, a.n), kWPB * 32, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(a, 0);
  CU(cudaGetLastError());
  h->launches += 1;
  return 0;
}

int earl_mj_get_obs(earl_mj_handle* h, float* obs_dev, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!obs_dev) return failf(EARL_ERR_INVALID, "null obs");
  StepArgs a = h->a;
  a.obs = obs_dev;
  mj_reset_kernel<<<grid_for(h, a.n), kWPB * 32, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(a, 1);
  CU(cudaGetLastError());
  h->launches += 1;
  return 0;
}

int earl_mj_step(earl_mj_handle* h, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                 uint8_t* success_dev, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!actions_dev || !obs_dev || !reward_dev || !done_dev) return failf(EARL_ERR_INVALID, "actions, obs, reward and done must be non-null");
  StepArgs a = h->a;
  a.actions = actions_dev;
  a.obs = obs_dev;
  a.reward = reward_dev;
  a.done = done_dev;
  a.success = success_dev;
  a.order_cur = h->d_order[h->order_sel];
  a.order_next = h->d_order[h->order_sel ^ 1];
  a.sched = h->d_sched;
  a.env_bucket = h->d_env_bucket;
  a.env_rank = h->d_env_rank;
  a.bucket_width = h->bucket_width;
  h->order_sel ^= 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CU(cudaMemsetAsync(h->d_sched, 0, (2 + kBuckets) * sizeof(unsigned), s));
  mj_step_kernel<<<grid_for(h, a.n), kWPB * 32, kSmemBytes, s>>>(a);
  CU(cudaGetLastError());
  int og = (a.n + 255) / 256;
  mj_order_kernel<<<og < 4 * h->sm_count ? og : 4 * h->sm_count, 256, 0, s>>>(a);
  CU(cudaGetLastError());
  h->launches += 2;
  h->total_steps += 1;
  return 0;
}

int earl_mj_step_host(earl_mj_handle* h, const float* actions_host, float* obs_host, float* reward_host, uint8_t* done_host,
                      uint8_t* success_host) {
  if (int rc = check_handle(h)) return rc;
  if (!actions_host || !obs_host || !reward_host || !done_host) return failf(EARL_ERR_INVALID, "null host buffer");
  const size_t n = (size_t)h->a.n;
  if (!h->d_act) {
    int rc = h->alloc(&h->d_act, n * kAct);
    if (!rc) rc = h->alloc(&h->d_obs, n * kObs);
    if (!rc) rc = h->alloc(&h->d_rew, n);
    if (!rc) rc = h->alloc(&h->d_done, n);
    if (!rc) rc = h->alloc(&h->d_succ, n);
    if (rc) return rc;
  }
  cudaStream_t s = h->host_stream;
  CU(cudaMemcpyAsync(h->d_act, actions_host, n * kAct * sizeof(float), cudaMemcpyHostToDevice, s));
  if (int rc = earl_mj_step(h, h->d_act, h->d_obs, h->d_rew, h->d_done, success_host ? h->d_succ : nullptr, s)) return rc;
  CU(cudaMemcpyAsync(obs_host, h->d_obs, n * kObs * sizeof(float), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(reward_host, h->d_rew, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(done_host, h->d_done, n, cudaMemcpyDeviceToHost, s));
  if (success_host) CU(cudaMemcpyAsync(success_host, h->d_succ, n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return 0;
}

int earl_mj_get_state(earl_mj_handle* h, double* qpos_host, double* qvel_host, double* warm_host, double* mocap_host) {
  if (int rc = check_handle(h)) return rc;
  const size_t n = (size_t)h->a.n;
  const int nq = h->hm.m.nq, nv = h->hm.m.nv;
  std::vector<float> rec(n * REC_FLOATS);
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(rec.data(), h->a.state, rec.size() * sizeof(float), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) {
    const float* r = &rec[i * REC_FLOATS];
    if (qpos_host) for (int k = 0; k < nq; ++k) qpos_host[i * nq + k] = r[REC_QPOS + k];
    if (qvel_host) for (int k = 0; k < nv; ++k) qvel_host[i * nv + k] = r[REC_QVEL + k];
    if (warm_host) for (int k = 0; k < nv; ++k) warm_host[i * nv + k] = r[REC_WARM + k];
    if (mocap_host) memcpy(&mocap_host[i * 3], &r[REC_MOCAP], 3 * sizeof(double));
  }
  return 0;
}

int earl_mj_set_state(earl_mj_handle* h, const double* qpos_host, const double* qvel_host, const double* warm_host,
                      const double* mocap_host) {
  if (int rc = check_handle(h)) return rc;
  const size_t n = (size_t)h->a.n;
  const int nq = h->hm.m.nq, nv = h->hm.m.nv;
  std::vector<float> rec(n * REC_FLOATS);
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(rec.data(), h->a.state, rec.size() * sizeof(float), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) {
    float* r = &rec[i * REC_FLOATS];
    if (qpos_host) for (int k = 0; k < nq; ++k) r[REC_QPOS + k] = (float)qpos_host[i * nq + k];
    if (qvel_host) for (int k = 0; k < nv; ++k) r[REC_QVEL + k] = (float)qvel_host[i * nv + k];
    if (warm_host) for (int k = 0; k < nv; ++k) r[REC_WARM + k] = (float)warm_host[i * nv + k];
    if (mocap_host) memcpy(&r[REC_MOCAP], &mocap_host[i * 3], 3 * sizeof(double));
  }
  CU(cudaMemcpy(h->a.state, rec.data(), rec.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

int earl_mj_counters(earl_mj_handle* h, int64_t* total_steps_host, int64_t* num_interventions_dev, uint32_t* steps_since_reset_dev,
                     double* lifelong_return_dev, void* stream) {
  if (int rc = check_handle(h)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = (size_t)h->a.n;
  if (total_steps_host) *total_steps_host = h->total_steps;
  if (num_interventions_dev)
    CU(cudaMemcpyAsync(num_interventions_dev, h->a.interventions, n * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  if (steps_since_reset_dev)
    CU(cudaMemcpy2DAsync(steps_since_reset_dev, sizeof(uint32_t), h->a.state + REC_STEPS, REC_FLOATS * sizeof(float),
                         sizeof(uint32_t), n, cudaMemcpyDeviceToDevice, s));
  if (lifelong_return_dev) {
    if (!h->a.ll_return) return failf(EARL_ERR_INVALID, "lifelong_return needs EARL_FLAG_LIFELONG");
    CU(cudaMemcpyAsync(lifelong_return_dev, h->a.ll_return, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

int earl_mj_eval_stats(earl_mj_handle* h, double* out4_dev, void* stream) {
  if (int rc = check_handle(h)) return rc;
  if (!out4_dev) return failf(EARL_ERR_INVALID, "null out4");
  if (!h->a.ep_return) return failf(EARL_ERR_INVALID, "eval stats need EARL_FLAG_EVAL_STATS");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CU(cudaMemsetAsync(out4_dev, 0, 4 * sizeof(double), s));
  int grid = (h->a.n + 255) / 256;
  if (grid > h->sm_count * 4) grid = h->sm_count * 4;
  mj_eval_stats_kernel<<<grid, 256, 0, s>>>(h->a.state, h->a.ep_return, h->a.n, out4_dev);
  CU(cudaGetLastError());
  h->launches += 1;
  return 0;
}

int earl_mj_work_counters(earl_mj_handle* h, uint64_t* out7_host) {
  if (int rc = check_handle(h)) return rc;
  if (!out7_host) return failf(EARL_ERR_INVALID, "null out");
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(out7_host, h->a.work, 7 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  if (getenv("EARL_MJ_OVERFLOW_DETAIL")) {
    uint64_t d[20];
    CU(cudaMemcpy(d, h->a.work, sizeof(d), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[mj overflow] env-steps with dropped candidate pairs %llu, dropped contacts %llu, dropped rows %llu\n",
            (unsigned long long)d[7], (unsigned long long)d[16], (unsigned long long)d[17]);
  }
#ifdef MJ_PHASE_TIMING
  uint64_t ph[8];
  CU(cudaMemcpy(ph, h->a.work + 8, sizeof(ph), cudaMemcpyDeviceToHost));
  static const char* names[8] = {"kinematics", "mass_matrix", "collide", "constraint_rows", "bias", "smooth", "solve", "euler"};
  uint64_t tot = 0;
  for (int k = 0; k < 8; ++k) tot += ph[k];
  for (int k = 0; k < 8; ++k) fprintf(stderr, "[mj phase] %-16s %6.2f %%  %10.0f cycles/env-step\n", names[k], 100.0 * ph[k] / (tot ? tot : 1), (double)ph[k] / (out7_host[0] ? out7_host[0] : 1));
#endif
  return 0;
}

}  // extern "C"
