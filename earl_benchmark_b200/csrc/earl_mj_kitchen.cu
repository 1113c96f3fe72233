// earl_mj_kitchen.cu -- the kitchen capacity set of the articulated-body engine (24 dofs, 128 geoms, 192 rows, 24
// contacts; joint-equality / friction-loss / pyramidal rows, capsule geoms) and its ENGINE-LEVEL entry points
// (include/earl_mj_kitchen_b200.h).  One warp per environment; the 40.7 KB workspace of an environment lives in shared
// memory (5 environments in flight per SM), the 62 KB model stays in global memory (L1 / L2 resident: every block reads
// the same tables).  No CPU fallback.
#define MJ_CAPSET_KITCHEN 1
#define mj mjk  // engine namespace of this translation unit: no symbol is shared with the door / peg capacity sets
#include "../../include/earl_mj_kitchen_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include <cuda_runtime.h>

#include "mj_model_host.hpp"
#include "mj_step.cuh"

namespace earl {
int set_error(int code, const char* msg);  // earl_b200.cu
}

namespace {

using namespace earl::mjk;

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

constexpr int kWPB = 5;  // warps (= environments in flight) per block
constexpr size_t kWorkStride = (sizeof(Work) + 15) & ~size_t(15);
constexpr size_t kSmemBytes = kWPB * kWorkStride;
static_assert(kSmemBytes <= 227 * 1024, "workspaces exceed the 227 KB of shared memory per block");

// All warps of a block walk the same number of environments and substeps (the engine's phase barriers are block-wide);
// a warp without an environment of its own shadows the last one and stores nothing.
__global__ void __launch_bounds__(kWPB * 32, 1)
mjk_substeps_kernel(const Model* __restrict__ gm, const real* __restrict__ hull, int n, int nsub, float* qpos, float* qvel, float* warm,
                    const double* mocap_pos, float4 mocap_quat, const float* ctrl, int* info) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Work& w = *reinterpret_cast<Work*>(smem + warp * kWorkStride);
  const Model& m = *gm;
  const int nq = m.nq, nv = m.nv;
  for (int base = blockIdx.x * kWPB; base < n; base += gridDim.x * kWPB) {
    const bool own = base + warp < n;
    const int env = own ? base + warp : n - 1;
    for (int k = lane; k < nq; k += 32) w.qpos[k] = qpos[(size_t)env * nq + k];
    for (int k = lane; k < nv; k += 32) { w.qvel[k] = qvel[(size_t)env * nv + k]; w.warm[k] = warm[(size_t)env * nv + k]; }
    if (lane == 0) {
      for (int k = 0; k < 3; ++k) w.mocap_pos[k] = mocap_pos[(size_t)env * 3 + k];
      w.mocap_quat[0] = mocap_quat.x; w.mocap_quat[1] = mocap_quat.y; w.mocap_quat[2] = mocap_quat.z; w.mocap_quat[3] = mocap_quat.w;
      for (int k = 0; k < m.nu; ++k) w.ctrl[k] = ctrl[(size_t)env * m.nu + k];
      w.bad = 0; w.acc_iter = w.acc_rows = w.acc_con = w.acc_mpr = w.acc_sup = 0;
    }
    __syncwarp();
    for (int s = 0; s < nsub; ++s) substep<32>(m, hull, w, lane);
    __syncwarp();
    if (own) {
      for (int k = lane; k < nq; k += 32) qpos[(size_t)env * nq + k] = w.qpos[k];
      for (int k = lane; k < nv; k += 32) { qvel[(size_t)env * nv + k] = w.qvel[k]; warm[(size_t)env * nv + k] = w.warm[k]; }
      if (lane == 0) {
        info[4 * env] = w.nefc; info[4 * env + 1] = w.ncon; info[4 * env + 2] = w.acc_iter; info[4 * env + 3] = w.bad;
      }
    }
    __syncthreads();
  }
}

}  // namespace

struct earl_mjk_engine {
  HostModel hm;
  int device = 0, sm_count = 0;
  Model* d_model = nullptr;
  real* d_hull = nullptr;
};

extern "C" {

int earl_mjk_engine_create(const void* model_blob, size_t model_nbytes, int32_t device, earl_mjk_engine** out) {
  if (!model_blob || !out) return failf(EARL_ERR_INVALID, "null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return failf(EARL_ERR_CUDA, "no CUDA device: the kitchen engine has no CPU path");
  if (device < 0 || device >= ndev) return failf(EARL_ERR_INVALID, "device %d out of range", device);
  earl_mjk_engine* e = new (std::nothrow) earl_mjk_engine();
  if (!e) return failf(EARL_ERR_NOMEM, "out of host memory");
  TaskSpec t;
  memset(&t, 0, sizeof(t));
  t.frame_skip = 40;
  t.obj_geom = -1;
  std::string msg;
  if (!build_model(model_blob, model_nbytes, t, &e->hm, &msg)) {
    delete e;
    return failf(EARL_ERR_INVALID, "%s", msg.c_str());
  }
  e->device = device;
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  e->sm_count = prop.multiProcessorCount;
  CU(cudaMalloc(&e->d_model, sizeof(Model)));
  CU(cudaMemcpy(e->d_model, &e->hm.m, sizeof(Model), cudaMemcpyHostToDevice));
  const size_t hb = (e->hm.hull_vert.size() + 4) * sizeof(real);
  CU(cudaMalloc(&e->d_hull, hb));
  CU(cudaMemcpy(e->d_hull, e->hm.hull_vert.data(), e->hm.hull_vert.size() * sizeof(real), cudaMemcpyHostToDevice));
  CU(cudaFuncSetAttribute(mjk_substeps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  *out = e;
  return 0;
}

int earl_mjk_engine_destroy(earl_mjk_engine* e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  cudaFree(e->d_model);
  cudaFree(e->d_hull);
  delete e;
  return 0;
}

int earl_mjk_engine_nv(const earl_mjk_engine* e) { return e ? e->hm.m.nv : 0; }

int earl_mjk_engine_substeps(earl_mjk_engine* e, int32_t num_envs, int32_t nsub, float* qpos_dev, float* qvel_dev, float* warm_dev,
                             const double* mocap_pos_dev, const float* mocap_quat_host, const float* ctrl_dev, int32_t* info_dev,
                             void* stream) {
  if (!e) return failf(EARL_ERR_INVALID, "null engine");
  if (num_envs <= 0 || nsub < 0 || !qpos_dev || !qvel_dev || !warm_dev || !mocap_pos_dev || !mocap_quat_host || !ctrl_dev || !info_dev)
    return failf(EARL_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(e->device));
  const int blocks = (num_envs + kWPB - 1) / kWPB;
  const int grid = blocks < e->sm_count ? blocks : e->sm_count;
  const float4 mq = make_float4(mocap_quat_host[0], mocap_quat_host[1], mocap_quat_host[2], mocap_quat_host[3]);
  mjk_substeps_kernel<<<grid, kWPB * 32, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(
      e->d_model, e->d_hull, num_envs, nsub, qpos_dev, qvel_dev, warm_dev, mocap_pos_dev, mq, ctrl_dev, info_dev);
  CU(cudaGetLastError());
  return 0;
}

}  // extern "C"
