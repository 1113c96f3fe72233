// tt3_env.cuh -- per-environment arithmetic of the three-object tabletop step, written once for the device kernels
// (csrc/earl_tt3.cu) and for the host build of the same source that the CPU test suite checks against the reference
// fixtures (tests/host_emulation/emul_tt3.cpp).  Follows earl_benchmark/envs/tabletop_manipulation_3obj.py (3OBJ:line).
//
// Rounding contract: every operation rounds exactly where numpy rounds.  On the device that is spelled with the
// round-to-nearest intrinsics (the translation unit is also compiled with --fmad=false); on the host the same
// expressions are plain operators and the translation unit MUST be compiled with -ffp-contract=off.  The one fused
// operation numpy's BLAS performs (the 2-element fp64 dot) is an explicit fma in both builds.
#pragma once

#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define T3_HD __host__ __device__ __forceinline__
#else
#define T3_HD inline
#endif

namespace earl {
namespace tt3 {

#ifdef __CUDA_ARCH__
T3_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
T3_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
T3_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
T3_HD double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
T3_HD double dsqrt(double a) { return __dsqrt_rn(a); }
T3_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
T3_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
T3_HD float fsqrt(float a) { return __fsqrt_rn(a); }
#else
T3_HD double dadd(double a, double b) { return a + b; }
T3_HD double dsub(double a, double b) { return a - b; }
T3_HD double dmul(double a, double b) { return a * b; }
T3_HD double dfma(double a, double b, double c) { return std::fma(a, b, c); }
T3_HD double dsqrt(double a) { return std::sqrt(a); }
T3_HD float fsub(float a, float b) { return a - b; }
T3_HD float fmul(float a, float b) { return a * b; }
T3_HD float fsqrt(float a) { return std::sqrt(a); }
#endif

struct EnvConst {
  double act_lo, act_span;  // -move_distance, move_distance - (-move_distance)      3OBJ:89
  double threshold;         // 0.4, strict <                                         3OBJ:38,107
  double clip;              // 2.8                                                   3OBJ:113,119
  double success_radius;    // 0.4, <=                                               3OBJ:165
};

struct EnvState {
  double fx, fy;            // fist                         qpos[0:2]
  double ox[3], oy[3];      // objects in object_dict order qpos[2:8]                3OBJ:31-35
  uint32_t att;             // 0 none, 1..3 = (0,0) / (0.5,0.5) / (1,1)
};

T3_HD double clipd(double x, double lo, double hi) {  // np.clip: NaN propagates
  return x < lo ? lo : (x > hi ? hi : x);
}

// 3OBJ:88-90: clip to [-1,1] (fp32 action against fp64 bounds promotes), then lb + (a + 1) * 0.5 * (ub - lb)
T3_HD double rescale(const EnvConst& c, float a) {
  return dadd(c.act_lo, dmul(dmul(dadd(clipd((double)a, -1.0, 1.0), 1.0), 0.5), c.act_span));
}

// env.step's state update (3OBJ:86-144): grasp the closest object strictly inside the threshold (dict order breaks
// exact ties), move the fist by the clipped action, drag the held object by the fist's clipped displacement
T3_HD void move(const EnvConst& c, EnvState& s, float a0f, float a1f, float a2f) {
  const double a0 = rescale(c, a0f), a1 = rescale(c, a1f), a2 = rescale(c, a2f);
  if (a2 > 0.0) {  // 3OBJ:98-108
    if (s.att == 0) {
      double held = HUGE_VAL;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double dx = dsub(s.fx, s.ox[k]), dy = dsub(s.fy, s.oy[k]);
        const double dist = dsqrt(dfma(dy, dy, dmul(dx, dx)));  // np.linalg.norm of the fp64 2-vector
        if (dist < c.threshold && dist < held) {
          s.att = (uint32_t)(k + 1);
          held = dist;
        }
      }
    }
  } else {
    s.att = 0;  // 3OBJ:109-110
  }
  const double nfx = clipd(dadd(s.fx, a0), -c.clip, c.clip);  // 3OBJ:112-113
  const double nfy = clipd(dadd(s.fy, a1), -c.clip, c.clip);
  if (s.att) {  // 3OBJ:114-119
    const double ddx = dsub(nfx, s.fx), ddy = dsub(nfy, s.fy);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (s.att == (uint32_t)(k + 1)) {
        s.ox[k] = clipd(dadd(s.ox[k], ddx), -c.clip, c.clip);
        s.oy[k] = clipd(dadd(s.oy[k], ddy), -c.clip, c.clip);
      }
  }
  s.fx = nfx;
  s.fy = nfy;
}

// obs[0:8] (3OBJ:49-54): the fp32 casts of qpos[0:8]
T3_HD void observe8(const EnvState& s, float (&o)[8]) {
  o[0] = (float)s.fx;
  o[1] = (float)s.fy;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    o[2 + 2 * k] = (float)s.ox[k];
    o[3 + 2 * k] = (float)s.oy[k];
  }
}

T3_HD float marker(uint32_t att) {  // attached_object tuples, 3OBJ:31-37
  return att == 0 ? -1.0f : 0.5f * (float)(att - 1);
}

// np.linalg.norm of an fp32 vector as numpy's BLAS evaluates it: fp32 products, accumulated in index order in fp64,
// rounded to fp32, fp32 sqrt
template <int K>
T3_HD float norm_f32(const float (&d)[K]) {
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < K; ++k) s = dadd(s, (double)fmul(d[k], d[k]));
  return fsqrt((float)s);
}

// is_successful (3OBJ:161-165): ||obs[0:8] - obs[10:18]|| <= 0.4, fp32 norm against the fp64 constant
T3_HD bool success(const float (&o)[8], const float* g, double radius) {
  float d[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) d[k] = fsub(o[k], g[k]);
  return (double)norm_f32(d) <= radius;
}

// dense reward (3OBJ:150-157): fp32 norms and squares, fp64 from the division by 0.01 on (numpy 1.22 promotion)
T3_HD double dense(const float (&o)[8], const float* g) {
  float d6[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) d6[k] = fsub(o[2 + k], g[2 + k]);
  double r = (double)(-norm_f32(d6));
#pragma unroll
  for (int j = 1; j < 4; ++j) {
    float d2[2] = {fsub(o[2 * j], g[2 * j]), fsub(o[2 * j + 1], g[2 * j + 1])};
    const float nk = norm_f32(d2);
    r += 2.0 * exp((double)(-fmul(nk, nk)) / 0.01);
  }
  return r;
}

}  // namespace tt3
}  // namespace earl
