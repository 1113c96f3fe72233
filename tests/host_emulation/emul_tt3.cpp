// Host build of the per-env arithmetic of the three-object tabletop kernels (earl_benchmark_b200/csrc/tt3_env.cuh).
// TEST INFRASTRUCTURE: lets the CPU suite check the exact kernel source against the reference fixtures without a
// GPU.  Must be compiled with -ffp-contract=off (see the header's rounding contract).  The product library never
// links or loads this file.
#include "../../earl_benchmark_b200/csrc/tt3_env.cuh"

using namespace earl::tt3;

extern "C" {

// One env.step of n envs, laid out like the C ABI's buffers: qpos [n,8] and att [n] in/out, actions [n,3], goal [n,10]
// (fp32 rows as _get_obs emits them) in; obs [n,20], reward [n], success [n] out.
void emu_tt3_step(int n, double* qpos, int* att, const float* actions, const float* goal, int dense_reward, double threshold,
                  double move_distance, double clip, double success_radius, float* obs, float* reward, unsigned char* succ) {
  const EnvConst c{-move_distance, move_distance - (-move_distance), threshold, clip, success_radius};
  for (int i = 0; i < n; ++i) {
    double* q = qpos + 8 * i;
    EnvState s{q[0], q[1], {q[2], q[4], q[6]}, {q[3], q[5], q[7]}, (uint32_t)att[i]};
    move(c, s, actions[3 * i], actions[3 * i + 1], actions[3 * i + 2]);
    q[0] = s.fx;
    q[1] = s.fy;
    for (int k = 0; k < 3; ++k) {
      q[2 + 2 * k] = s.ox[k];
      q[3 + 2 * k] = s.oy[k];
    }
    att[i] = (int)s.att;
    float o[8];
    observe8(s, o);
    const float* g = goal + 10 * i;
    float* out = obs + 20 * i;
    for (int k = 0; k < 8; ++k) out[k] = o[k];
    out[8] = out[9] = marker(s.att);
    for (int k = 0; k < 10; ++k) out[10 + k] = g[k];
    const bool ok = success(o, g, success_radius);
    succ[i] = ok ? 1 : 0;
    reward[i] = dense_reward ? (float)dense(o, g) : (ok ? 1.0f : 0.0f);
  }
}

}  // extern "C"
