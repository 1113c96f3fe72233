// mj_step.cuh -- one engine substep (mj_forward + mj_Euler) and the Sawyer task layer on top of it.
// See mj_engine.cuh for the execution model.
#pragma once

#include "mj_collide.cuh"
#include "mj_engine.cuh"

// which of the four block-wide phase barriers of a substep are kept (bit 0: top, 1: after the collision phase, 2: before
// the solve, 3: after it); all four by default (measured: profiles/r01/README.md)
#ifndef MJ_BSYNC_MASK
#define MJ_BSYNC_MASK 15
#endif

namespace earl {
namespace mj {

// ------------------------------------------------------------------------------------------------ forward + Euler
template <int NL>
MJ_FN void substep(const Model& m, const real* hull, Work& w, int lane) {
  const int nv = m.nv;
  if (MJ_BSYNC_MASK & 1) bsync<NL>();
  MJ_PHASE_BEGIN(w);
  kinematics<NL>(m, w, lane);
  MJ_PHASE_END(w, 0);
  mass_matrix<NL>(m, w, lane);
  MJ_PHASE_END(w, 1);
  collide<NL>(m, hull, w, lane);
  if (MJ_BSYNC_MASK & 2) bsync<NL>();
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
  if (MJ_BSYNC_MASK & 4) bsync<NL>();
  MJ_PHASE_END(w, 5);
  solve<NL>(m, w, lane);
  if (MJ_BSYNC_MASK & 8) bsync<NL>();
  MJ_PHASE_END(w, 6);
  if (lane == 0) {
    w.acc_iter += w.solver_iter; w.acc_rows += w.nefc; w.acc_con += w.ncon;
    if (w.nefc > w.peak_efc) w.peak_efc = w.nefc;
    if (w.ncon > w.peak_con) w.peak_con = w.ncon;
    if (w.nhit > w.peak_hit) w.peak_hit = w.nhit;
  }
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
// A warp that has nothing (more) to compute in this substep keeps its block's phase barriers company: exactly the
// barriers of substep(), nothing else.
template <int NL>
MJ_HD void substep_idle() {
  if (MJ_BSYNC_MASK & 1) bsync<NL>();
  if (MJ_BSYNC_MASK & 2) bsync<NL>();
  if (MJ_BSYNC_MASK & 4) bsync<NL>();
  if (MJ_BSYNC_MASK & 8) bsync<NL>();
}

// mocap target and gripper control of one env step (SawyerXYZEnv.set_xyz_action + the gripper effort)
template <int NL>
MJ_FN void env_set_action(const Model& m, Work& w, const real* action, int lane) {
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
}

template <int NL>
MJ_FN void env_step(const Model& m, const real* hull, Work& w, const real* action, int lane) {
  env_set_action<NL>(m, w, action, lane);
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


// metaworld reward_utils [metaworld@master, not under /root/reference; formulas as in SURVEY.md Appendix C]:
// tolerance(x, bounds=(lower, upper), margin, sigmoid='long_tail', value_at_margin=0.1)
MJ_HD real tolerance_long_tail(real x, real lower, real upper, real margin) {
  if (x >= lower && x <= upper) return 1.0f;
  if (margin == 0) return 0.0f;
  const real d = (x < lower ? lower - x : x - upper) / margin;
  return 1.0f / (d * d * 9.0f + 1.0f);  // scale^2 = 1 / 0.1 - 1
}
// hamacher_product(a, b) = ab / (a + b - ab), 0 when the denominator is 0
MJ_HD real hamacher(real a, real b) {
  const real den = a + b - a * b;
  return den > 0 ? a * b / den : 0.0f;
}
// rect_prism_tolerance(curr, zero, one): inside the prism spanned by the corners `zero` and `one` the product of the
// normalised coordinates (0 at `zero`, 1 at `one`), outside 1
MJ_HD real rect_prism_tolerance(const real* curr, const real* zero, const real* one) {
  real prod = 1;
  for (int k = 0; k < 3; ++k) {
    const bool in = one[k] >= zero[k] ? (zero[k] <= curr[k] && curr[k] <= one[k]) : (one[k] <= curr[k] && curr[k] <= zero[k]);
    if (!in) return 1.0f;
    prod *= (curr[k] - zero[k]) / (one[k] - zero[k]);
  }
  return prod;
}

// SawyerPegV2.compute_reward, dense branch (reference earl_benchmark/envs/sawyer_peg.py:231-299) with metaworld's
// SawyerXYZEnv._gripper_caging_reward(high_density=True).  obs7 = [hand(3), gripper, pegHead(3)]; obj_init = the peg
// position the last reset wrote (self.obj_init_pos), head_init = self.peg_head_pos_init, init_tcp = self.init_tcp.
MJ_FN real peg_dense_reward(const Model& m, const Work& w, const real* obs7, const real* target, real grip_action, const real* obj_init,
                            const real* head_init, const real* init_tcp) {
  real head[3], grasp[3], obj[3], zero[3], one[3], lp[3], rp[3], ree[3], lee[3];
  site_xpos(m, w, m.obs_obj_site, head);
  site_xpos(m, w, m.grasp_site, grasp);
  real tto = 0, ott = 0, ipm = 0;
  const real scale[3] = {1.0f, 2.0f, 2.0f};
  for (int k = 0; k < 3; ++k) {
    obj[k] = obs7[4 + k] - head[k] + grasp[k];
    const real a = obj[k] - obs7[k], b = (obs7[4 + k] - target[k]) * scale[k], c = (head_init[k] - target[k]) * scale[k];
    tto += a * a; ott += b * b; ipm += c * c;
  }
  tto = sqrtf(tto); ott = sqrtf(ott); ipm = sqrtf(ipm);
  real in_place = tolerance_long_tail(ott, 0.0f, m.success_radius, ipm);
  real cb[2];
  for (int q = 0; q < 2; ++q) {
    site_xpos(m, w, m.corner_site[2 * q], zero);      // bottom right corner: reward 0
    site_xpos(m, w, m.corner_site[2 * q + 1], one);   // top left corner: reward 1
    cb[q] = rect_prism_tolerance(obs7 + 4, zero, one);
  }
  in_place = hamacher(in_place, hamacher(cb[1], cb[0]));
  const bool lifted = tto < 0.08f && obs7[3] > 0 && obj[2] - 0.01f > obj_init[2];
  real grasped;
  if (lifted) {
    grasped = 1.0f;
  } else {  // _gripper_caging_reward(action, obj, object_reach_radius=0.01, obj_radius=0.0075, pad_success_thresh=0.03, xz_thresh=0.005)
    const real obj_radius = 0.0075f, pad_thresh = 0.03f, xz_thresh = 0.005f;
    site_xpos(m, w, m.lpad_site, lp);
    site_xpos(m, w, m.rpad_site, rp);
    const real pad_y[2] = {lp[1], rp[1]};
    real cy[2];
    for (int q = 0; q < 2; ++q) {
      const real to_obj = fabsf(pad_y[q] - obj[1]), to_init = fabsf(pad_y[q] - obj_init[1]);
      cy[q] = tolerance_long_tail(to_obj, obj_radius, pad_thresh, fabsf(to_init - pad_thresh));
    }
    const real caging_y = hamacher(cy[0], cy[1]);
    site_xpos(m, w, m.obs_ree_site, ree);
    site_xpos(m, w, m.obs_lee_site, lee);
    const real tx = 0.5f * (ree[0] + lee[0]), tz = 0.5f * (ree[2] + lee[2]);
    const real mx = obj_init[0] - init_tcp[0], mz = obj_init[2] - init_tcp[2];
    real xz_margin = sqrtf(mx * mx + mz * mz) - xz_thresh;
    if (xz_margin < 0) xz_margin = 0;  // the reference raises ValueError for a negative margin; never reached from a reset pose
    const real dx = tx - obj[0], dz = tz - obj[2];
    const real caging_xz = tolerance_long_tail(sqrtf(dx * dx + dz * dz), 0.0f, xz_thresh, xz_margin);
    const real closed = fminf(fmaxf(0.0f, grip_action), 1.0f);
    const real caging = hamacher(caging_y, caging_xz);
    const real gripping = caging > 0.97f ? closed : 0.0f;
    grasped = 0.5f * (hamacher(caging, gripping) + caging);  // high_density
  }
  real reward = hamacher(grasped, in_place);
  if (lifted) reward += 1.0f + 5.0f * in_place;
  if (ott <= m.success_radius) reward = 10.0f;
  return reward;
}

}  // namespace mj
}  // namespace earl
