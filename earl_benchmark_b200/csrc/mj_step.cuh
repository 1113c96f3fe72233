// mj_step.cuh -- one engine substep (mj_forward + mj_Euler) and the Sawyer task layer on top of it.
// See mj_engine.cuh for the execution model.
#pragma once

#include "mj_collide.cuh"
#include "mj_engine.cuh"

namespace earl {
namespace mj {

// ------------------------------------------------------------------------------------------------ forward + Euler
template <int NL>
MJ_FN void substep(const Model& m, const real* hull, Work& w, int lane) {
  const int nv = m.nv;
  bsync<NL>();
  MJ_PHASE_BEGIN(w);
  kinematics<NL>(m, w, lane);
  MJ_PHASE_END(w, 0);
  mass_matrix<NL>(m, w, lane);
  MJ_PHASE_END(w, 1);
  collide<NL>(m, hull, w, lane);
  bsync<NL>();
#if defined(MJ_TRACE_DEVICE) && defined(__CUDA_ARCH__)
  if (threadIdx.x == 0)
    for (int c = 0; c < w.ncon; ++c)
      printf("  [dev] con %d g %d %d dist %.9g pos %.9g %.9g %.9g n %.9g %.9g %.9g\n", c, w.con_g1[c], w.con_g2[c], (double)w.con_dist[c],
             (double)w.con_pos[c][0], (double)w.con_pos[c][1], (double)w.con_pos[c][2], (double)w.con_frame[c][0],
             (double)w.con_frame[c][1], (double)w.con_frame[c][2]);
#endif
  MJ_PHASE_END(w, 2);
  make_constraints<NL>(m, w, lane);
  contact_rows<NL>(m, w, lane);
  MJ_PHASE_END(w, 3);
  bias_forces<NL>(m, w, lane);
  MJ_PHASE_END(w, 4);
  // passive (damping, springs) - bias + actuation
  for (int i = lane; i < nv; i += NL) w.smooth[i] = -m.dof_damping[i] * w.qvel[i] - w.bias[i];
  wsync<NL>();
  if (lane == 0) {
    for (int j = 0; j < m.njnt; ++j)
      if (m.jnt_stiffness[j] != 0 && m.jnt_type[j] >= 2)
        w.smooth[m.jnt_dofadr[j]] -= m.jnt_stiffness[j] * (w.qpos[m.jnt_qposadr[j]] - m.jnt_springref[j]);
    for (int u = 0; u < m.nu; ++u) {
      real c = w.ctrl[u];
      if (m.act_ctrllimited[u]) c = clampr(c, m.act_ctrlrange[u][0], m.act_ctrlrange[u][1]);
      real f = m.act_kp[u] * c - m.act_kp[u] * w.qpos[m.act_qposadr[u]];
      if (m.act_forcelimited[u]) f = clampr(f, m.act_forcerange[u][0], m.act_forcerange[u][1]);
      w.smooth[m.act_dof[u]] += f;
    }
  }
  wsync<NL>();
  bsync<NL>();
  MJ_PHASE_END(w, 5);
  solve<NL>(m, w, lane);
  bsync<NL>();
  MJ_PHASE_END(w, 6);
  if (lane == 0) { w.acc_iter += w.solver_iter; w.acc_rows += w.nefc; w.acc_con += w.ncon; }
  // mj_Euler: implicit in joint damping
  const real h = m.timestep;
  for (int i = lane; i < nv; i += NL) {
    for (int j = 0; j < nv; ++j) w.H[i][j] = w.M[i][j];
    w.H[i][i] += h * m.dof_damping[i];
    w.tmp[i] = w.smooth[i] + w.fcon[i];
    w.warm[i] = w.acc[i];
  }
  wsync<NL>();
  if (!spd_solve<NL>(w.H, nv, w.tmp, lane)) w.bad |= 1;
  for (int i = lane; i < nv; i += NL) w.qvel[i] += h * w.tmp[i];
  wsync<NL>();
  for (int j = lane; j < m.njnt; j += NL) {
    const int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
    if (m.jnt_type[j] == 0) {
      for (int k = 0; k < 3; ++k) w.qpos[qa + k] += h * w.qvel[da + k];
      const real* wl = &w.qvel[da + 3];
      const real ang = sqrtf(dot3(wl, wl)) * h;
      if (ang >= MINVAL) {
        const real s = sinf(0.5f * ang) * h / ang, c = cosf(0.5f * ang);
        real dq[4] = {c, wl[0] * s, wl[1] * s, wl[2] * s}, q[4];
        mulquat(q, &w.qpos[qa + 3], dq);
        normquat(q);
        for (int k = 0; k < 4; ++k) w.qpos[qa + 3 + k] = q[k];
      }
    } else {
      w.qpos[qa] += h * w.qvel[da];
    }
  }
  wsync<NL>();
  MJ_PHASE_END(w, 7);
}

// ------------------------------------------------------------------------------------------------ task layer
// metaworld SawyerXYZEnv.step (set_xyz_action + do_simulation) followed by the EARL observation / sparse reward
// (reference earl_benchmark/envs/sawyer_door.py:86-94,168-177).  `action` has 4 entries.
template <int NL>
MJ_FN void env_step(const Model& m, const real* hull, Work& w, const real* action, int lane) {
  if (lane == 0) {
    for (int k = 0; k < 3; ++k) {
      const double a = (double)clampr(action[k], -1.0f, 1.0f);
      double p = w.mocap_pos[k] + a * (double)m.action_scale;
      p = p < (double)m.mocap_low[k] ? (double)m.mocap_low[k] : (p > (double)m.mocap_high[k] ? (double)m.mocap_high[k] : p);
      w.mocap_pos[k] = p;
    }
    w.mocap_quat[0] = 1; w.mocap_quat[1] = 0; w.mocap_quat[2] = 1; w.mocap_quat[3] = 0;
    const real g = clampr(action[3], -1.0f, 1.0f);
    w.ctrl[0] = g; w.ctrl[1] = -g;
  }
  wsync<NL>();
  for (int s = 0; s < m.frame_skip; ++s) substep<NL>(m, hull, w, lane);
}

// observation from the kinematics of the LAST forward pass (one substep stale, as in the reference: mj_step
// integrates after its forward pass and the env reads body / site / geom poses right after sim.step()).
MJ_HD void site_xpos(const Model& m, const Work& w, int s, real* out) {
  real t[3];
  const int b = m.site_body[s];
  mulmatvec3(t, w.xmat[b], m.site_pos[s]);
  for (int k = 0; k < 3; ++k) out[k] = w.xpos[b][k] + t[k];
}
MJ_HD void geom_xpos(const Model& m, const Work& w, int g, real* out) {
  real t[3];
  const int b = m.geom_body[g];
  mulmatvec3(t, w.xmat[b], m.geom_pos[g]);
  for (int k = 0; k < 3; ++k) out[k] = w.xpos[b][k] + t[k];
}
MJ_FN void observe(const Model& m, const Work& w, real* obs7) {
  real r[3], l[3];
  site_xpos(m, w, m.obs_hand_site, obs7);
  site_xpos(m, w, m.obs_ree_site, r);
  site_xpos(m, w, m.obs_lee_site, l);
  real d[3] = {r[0] - l[0], r[1] - l[1], r[2] - l[2]};
  obs7[3] = clampr(sqrtf(dot3(d, d)) / 0.1f, 0.0f, 1.0f);
  if (m.obs_obj_geom >= 0) geom_xpos(m, w, m.obs_obj_geom, obs7 + 4);
  else site_xpos(m, w, m.obs_obj_site, obs7 + 4);
}

// metaworld reward_utils.tolerance(x, bounds=(0, upper), margin, sigmoid='gaussian', value_at_margin=0.1)
// [metaworld@master, not under /root/reference; formula as in SURVEY.md Appendix C]
MJ_HD real tolerance_gaussian(real x, real upper, real margin) {
  if (x >= 0 && x <= upper) return 1.0f;
  if (margin == 0) return 0.0f;
  const real d = (x < 0 ? -x : x - upper) / margin;
  const real scale2 = 4.605170185988092f;  // -2 ln(0.1)
  return expf(-0.5f * d * d * scale2);
}

// SawyerDoorV2.compute_reward, dense branch (reference earl_benchmark/envs/sawyer_door.py:141-171) on an observation
// [hand(3), gripper, handle(3)] and the goal's target position
MJ_HD real door_dense_reward(const Model& m, const real* obs7, const real* target) {
  const real TARGET_RADIUS = 0.05f;
  real d_ot = 0, d_to = 0, m_in = 0, m_hand = 0;
  for (int k = 0; k < 3; ++k) {
    const real a = obs7[4 + k] - target[k], b = obs7[k] - obs7[4 + k], c = m.obj_init_pos[k] - target[k],
               e = m.hand_init_pos[k] - obs7[4 + k];
    d_ot += a * a; d_to += b * b; m_in += c * c; m_hand += e * e;
  }
  d_ot = sqrtf(d_ot); d_to = sqrtf(d_to); m_in = sqrtf(m_in); m_hand = sqrtf(m_hand) + 0.1f;
  if (d_ot < TARGET_RADIUS) return 10.0f;
  return 3.0f * tolerance_gaussian(d_to, 0.25f * TARGET_RADIUS, m_hand) + 6.0f * tolerance_gaussian(d_ot, TARGET_RADIUS, m_in);
}

}  // namespace mj
}  // namespace earl
