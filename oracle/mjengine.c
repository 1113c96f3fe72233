/*
 * oracle/mjengine.c -- fp64 CPU restatement of the MuJoCo 2.1.0 `mj_step` subset that the reference's Sawyer
 * tasks exercise through mujoco-py (reference call sites: metaworld SawyerXYZEnv.do_simulation -> sim.step(),
 * invoked by earl_benchmark/envs/sawyer_door.py and sawyer_peg.py, which do not override step()).
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs; never by the
 * product package.
 *
 * The engine is a third-party dependency that is NOT under /root/reference (MuJoCo 2.1.0 via mujoco-py
 * 2.1.2.14, setup.py:10); this file restates its published algorithm (MuJoCo documentation, "Computation"
 * chapter; open-sourced engine_forward.c / engine_core_constraint.c / engine_core_smooth.c):
 *   mj_step = mj_forward (kinematics, CoM, CRB mass matrix, collision, constraint rows, passive forces, RNE bias,
 *   actuation, smooth acceleration, convex constraint solve) + mj_Euler (implicit joint damping).
 * Parity status: PARTIAL -- pinned only by the reference's golden constants and demonstrations (hand rest pose
 * of sawyer_door.py:13-16, free-space hand trajectories of the shipped demos); see tests/test_engine_oracle.py.
 *
 * The model is the FUSED structure-of-arrays model produced by earl_benchmark_b200/mjcf/compile.py
 * (one joint per body, static bodies merged into their moving ancestor).
 */
#include "mjengine.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MJMINVAL 1e-15
#define MJMINIMP 0.0001
#define MJMAXIMP 0.9999

/* ------------------------------------------------------------------------------------------ small math */
static inline void cross3(double *r, const double *a, const double *b) {
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void mulmatvec3(double *r, const double *m, const double *v) {
  r[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  r[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  r[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}
static inline void mulquat(double *r, const double *a, const double *b) {
  double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                 a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  memcpy(r, t, sizeof t);
}
static inline void normquat(double *q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MJMINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  for (int i = 0; i < 4; ++i) q[i] /= n;
}
static inline void quat2mat(double *m, const double *q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = 1 - 2 * (x * x + y * y);
}
static inline void axisangle2quat(double *q, const double *axis, double ang) {
  double s = sin(0.5 * ang);
  q[0] = cos(0.5 * ang); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}

/* ------------------------------------------------------------------------------------------ model blob */
typedef struct { const unsigned char *p, *end; } Rd;
static const void *rd_field(Rd *r, int elem, long long *count) {
  int ndim = *(const int *)r->p;
  long long n = 1;
  for (int i = 0; i < ndim; ++i) n *= ((const int *)r->p)[1 + i];
  size_t hb = 4 * (size_t)(1 + ndim);
  if (hb % 8) hb += 4;
  r->p += hb;
  const void *data = r->p;
  size_t db = (size_t)n * (size_t)elem;
  if (db % 8) db += 8 - db % 8;
  r->p += db;
  if (count) *count = n;
  return data;
}
#define RI(f) m->f = *(const int *)rd_field(&r, 4, 0)
#define RD(f) m->f = *(const double *)rd_field(&r, 8, 0)
#define PI(f) m->f = (const int *)rd_field(&r, 4, 0)
#define PD(f) m->f = (const double *)rd_field(&r, 8, 0)

mjModelF *mje_load(const void *blob, long long nbytes) {
  mjModelF *m = (mjModelF *)calloc(1, sizeof(mjModelF));
  m->blob = malloc((size_t)nbytes);
  memcpy(m->blob, blob, (size_t)nbytes);
  Rd r = {(const unsigned char *)m->blob, (const unsigned char *)m->blob + nbytes};
  m->version = ((const int *)r.p)[1];
  if (((const int *)r.p)[0] != 0x4C444D45 || (m->version != 1 && m->version != 2)) { free(m->blob); free(m); return 0; }
  r.p += 8;
  RI(nbody); RI(nq); RI(nv); RI(ngeom); RI(nsite); RI(nu); RI(nweld); RI(nhullvert); RI(iterations); RI(cone_elliptic);
  RD(timestep); RD(tolerance); RD(impratio); PD(gravity);
  PI(body_parent); PD(body_pos); PD(body_quat); PD(body_mass); PD(body_ipos); PD(body_inertia); PI(body_jnt);
  PI(jnt_type); PI(jnt_body); PI(jnt_qposadr); PI(jnt_dofadr); PD(jnt_pos); PD(jnt_axis); PI(jnt_limited); PD(jnt_range);
  PD(jnt_margin); PD(jnt_solref); PD(jnt_solimp); PD(jnt_stiffness); PD(jnt_springref);
  PI(dof_body); PD(dof_damping); PD(dof_armature); PD(dof_frictionloss); PD(dof_invweight0); PD(qpos0);
  PI(geom_body); PI(geom_type); PD(geom_size); PD(geom_pos); PD(geom_quat); PI(geom_contype); PI(geom_conaffinity);
  PI(geom_condim); PI(geom_priority); PD(geom_friction); PD(geom_margin); PD(geom_gap); PD(geom_solref); PD(geom_solimp);
  PD(geom_solmix); PD(geom_invweight0); PD(geom_rbound); PI(geom_hulladr); PI(geom_hullnum); PI(geom_srcbody);
  PI(geom_srcparent); PD(hull_vert);
  PI(site_body); PD(site_pos); PD(site_quat);
  PI(act_dof); PI(act_qposadr); PD(act_kp); PD(act_ctrlrange); PI(act_ctrllimited); PD(act_forcerange); PI(act_forcelimited);
  PI(weld_body); PD(weld_pos); PD(weld_quat); PD(weld_relpose); PD(weld_solref); PD(weld_solimp); PD(weld_invweight);
  PD(mocap_pos0); PD(mocap_quat0);
  if (m->version == 2) {
    RI(neq); PI(eq_qposadr); PI(eq_dofadr); PD(eq_polycoef); PD(eq_solref); PD(eq_solimp); PD(eq_invweight);
    PD(dof_solref_friction); PD(dof_solimp_friction);
  }
  if (r.p != r.end || m->nbody > MJ_MAXB || m->nv > MJ_MAXV || m->nq > MJ_MAXQ || m->ngeom > MJ_MAXG || m->nsite > MJ_MAXS) {
    free(m->blob); free(m); return 0;
  }
  return m;
}
void mje_free(mjModelF *m) { if (m) { free(m->blob); free(m); } }

mjDataF *mje_make_data(void) { return (mjDataF *)calloc(1, sizeof(mjDataF)); }
void mje_free_data(mjDataF *d) { free(d); }

/* sim.reset(): qpos0, zero velocities / warm start / ctrl, mocap pose from the model */
void mje_reset(const mjModelF *m, mjDataF *d) {
  memset(d, 0, sizeof *d);
  memcpy(d->qpos, m->qpos0, sizeof(double) * m->nq);
  memcpy(d->mocap_pos, m->mocap_pos0, 3 * sizeof(double));
  memcpy(d->mocap_quat, m->mocap_quat0, 4 * sizeof(double));
}

/* ------------------------------------------------------------------------------------------ mj_kinematics + mj_comPos */
void mje_kinematics(const mjModelF *m, mjDataF *d) {
  static const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memset(d->xpos[0], 0, 3 * sizeof(double));
  memcpy(d->xmat[0], eye, sizeof eye);
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  for (int b = 1; b < m->nbody; ++b) {
    int p = m->body_parent[b], j = m->body_jnt[b];
    double pos[3], quat[4], R[9], t[3];
    mulmatvec3(t, d->xmat[p], m->body_pos + 3 * b);
    for (int k = 0; k < 3; ++k) pos[k] = d->xpos[p][k] + t[k];
    mulquat(quat, d->xquat[p], m->body_quat + 4 * b);
    quat2mat(R, quat);
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == 3) { /* hinge */
      double anchor[3], axis[3], qj[4];
      mulmatvec3(t, R, m->jnt_pos + 3 * j);
      for (int k = 0; k < 3; ++k) anchor[k] = pos[k] + t[k];
      mulmatvec3(axis, R, m->jnt_axis + 3 * j);
      axisangle2quat(qj, m->jnt_axis + 3 * j, d->qpos[qa] - m->qpos0[qa]);
      mulquat(quat, quat, qj);
      normquat(quat);
      quat2mat(R, quat);
      mulmatvec3(t, R, m->jnt_pos + 3 * j);
      for (int k = 0; k < 3; ++k) { pos[k] = anchor[k] - t[k]; d->dof_axis[da][k] = axis[k]; d->dof_anchor[da][k] = anchor[k]; }
      d->dof_rot[da] = 1;
    } else if (m->jnt_type[j] == 2) { /* slide */
      double axis[3];
      mulmatvec3(axis, R, m->jnt_axis + 3 * j);
      for (int k = 0; k < 3; ++k) { pos[k] += axis[k] * (d->qpos[qa] - m->qpos0[qa]); d->dof_axis[da][k] = axis[k]; d->dof_anchor[da][k] = pos[k]; }
      d->dof_rot[da] = 0;
    } else if (m->jnt_type[j] == 0) { /* free: 3 world-axis translations, then 3 body-axis rotations */
      for (int k = 0; k < 3; ++k) pos[k] = d->qpos[qa + k];
      for (int k = 0; k < 4; ++k) quat[k] = d->qpos[qa + 3 + k];
      normquat(quat);
      for (int k = 0; k < 4; ++k) d->qpos[qa + 3 + k] = quat[k]; /* mj_kinematics normalises quaternions in place */
      quat2mat(R, quat);
      for (int a = 0; a < 3; ++a) {
        for (int k = 0; k < 3; ++k) {
          d->dof_axis[da + a][k] = (a == k);
          d->dof_anchor[da + a][k] = pos[k];
          d->dof_axis[da + 3 + a][k] = R[3 * k + a];
          d->dof_anchor[da + 3 + a][k] = pos[k];
        }
        d->dof_rot[da + a] = 0;
        d->dof_rot[da + 3 + a] = 1;
      }
    }
    memcpy(d->xpos[b], pos, sizeof pos);
    memcpy(d->xquat[b], quat, sizeof quat);
    memcpy(d->xmat[b], R, sizeof R);
    mulmatvec3(t, R, m->body_ipos + 3 * b);
    for (int k = 0; k < 3; ++k) d->xipos[b][k] = pos[k] + t[k];
  }
  for (int g = 0; g < m->ngeom; ++g) {
    int b = m->geom_body[g];
    double t[3], q[4];
    mulmatvec3(t, d->xmat[b], m->geom_pos + 3 * g);
    for (int k = 0; k < 3; ++k) d->geom_xpos[g][k] = d->xpos[b][k] + t[k];
    mulquat(q, d->xquat[b], m->geom_quat + 4 * g);
    quat2mat(d->geom_xmat[g], q);
  }
  d->flops += (long long)m->nbody * 130 + (long long)m->ngeom * 40 + (long long)m->nsite * 40;
  for (int s = 0; s < m->nsite; ++s) {
    int b = m->site_body[s];
    double t[3], q[4];
    mulmatvec3(t, d->xmat[b], m->site_pos + 3 * s);
    for (int k = 0; k < 3; ++k) d->site_xpos[s][k] = d->xpos[b][k] + t[k];
    mulquat(q, d->xquat[b], m->site_quat + 4 * s);
    quat2mat(d->site_xmat[s], q);
  }
}

/* world-frame rotational inertia of body b about its CoM */
static void body_inertia_world(const mjModelF *m, const mjDataF *d, int b, double *Iw) {
  const double *I6 = m->body_inertia + 6 * b, *R = d->xmat[b];
  double Il[9] = {I6[0], I6[3], I6[4], I6[3], I6[1], I6[5], I6[4], I6[5], I6[2]}, T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[3 * i + j] = R[3 * i] * Il[j] + R[3 * i + 1] * Il[3 + j] + R[3 * i + 2] * Il[6 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Iw[3 * i + j] = T[3 * i] * R[3 * j] + T[3 * i + 1] * R[3 * j + 1] + T[3 * i + 2] * R[3 * j + 2];
}

/* Jacobian columns of body b at world point `pt`: jp/jr [3][nv] (zero for dofs that do not move b) */
void mje_body_jac(const mjModelF *m, const mjDataF *d, int b, const double *pt, double jp[3][MJ_MAXV], double jr[3][MJ_MAXV]) {
  for (int k = 0; k < 3; ++k) { memset(jp[k], 0, sizeof(double) * m->nv); memset(jr[k], 0, sizeof(double) * m->nv); }
  for (int c = b; c > 0; c = m->body_parent[c]) {
    int j = m->body_jnt[c], da = m->jnt_dofadr[j], n = m->jnt_type[j] == 0 ? 6 : 1;
    for (int q = da; q < da + n; ++q) {
      if (d->dof_rot[q]) {
        double r[3] = {pt[0] - d->dof_anchor[q][0], pt[1] - d->dof_anchor[q][1], pt[2] - d->dof_anchor[q][2]}, c3[3];
        cross3(c3, d->dof_axis[q], r);
        for (int k = 0; k < 3; ++k) { jp[k][q] = c3[k]; jr[k][q] = d->dof_axis[q][k]; }
      } else {
        for (int k = 0; k < 3; ++k) jp[k][q] = d->dof_axis[q][k];
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------ mj_crb: joint-space inertia */
void mje_mass_matrix(const mjModelF *m, mjDataF *d) {
  int nv = m->nv;
  for (int i = 0; i < nv; ++i) memset(d->M[i], 0, sizeof(double) * nv);
  static double jp[3][MJ_MAXV], jr[3][MJ_MAXV];
  for (int b = 1; b < m->nbody; ++b) {
    double mass = m->body_mass[b];
    if (mass <= 0) continue;
    double Iw[9];
    body_inertia_world(m, d, b, Iw);
    mje_body_jac(m, d, b, d->xipos[b], jp, jr);
    int dofs[MJ_MAXV], nd = 0;
    for (int c = b; c > 0; c = m->body_parent[c]) {
      int j = m->body_jnt[c], da = m->jnt_dofadr[j], n = m->jnt_type[j] == 0 ? 6 : 1;
      for (int q = da; q < da + n; ++q) dofs[nd++] = q;
    }
    for (int a = 0; a < nd; ++a)
      for (int c = 0; c < nd; ++c) {
        int i = dofs[a], j = dofs[c];
        double Ij[3] = {Iw[0] * jr[0][j] + Iw[1] * jr[1][j] + Iw[2] * jr[2][j], Iw[3] * jr[0][j] + Iw[4] * jr[1][j] + Iw[5] * jr[2][j],
                        Iw[6] * jr[0][j] + Iw[7] * jr[1][j] + Iw[8] * jr[2][j]};
        d->M[i][j] += mass * (jp[0][i] * jp[0][j] + jp[1][i] * jp[1][j] + jp[2][i] * jp[2][j]) + jr[0][i] * Ij[0] + jr[1][i] * Ij[1] +
                      jr[2][i] * Ij[2];
      }
    d->flops += (long long)nd * nd * 18;
  }
  for (int i = 0; i < nv; ++i) d->M[i][i] += m->dof_armature[i];
}

/* ------------------------------------------------------------------------------------------ mj_rne (bias: Coriolis, centrifugal, gravity)
 * Spatial vectors about the WORLD ORIGIN in world axes: motion [w; vO], force [nO; f]. */
void mje_bias(const mjModelF *m, mjDataF *d) {
  int nb = m->nbody;
  double V[MJ_MAXB][6], A[MJ_MAXB][6], F[MJ_MAXB][6];
  memset(V[0], 0, sizeof V[0]);
  for (int k = 0; k < 3; ++k) { A[0][k] = 0; A[0][3 + k] = -m->gravity[k]; }
  for (int b = 1; b < nb; ++b) {
    int p = m->body_parent[b], j = m->body_jnt[b], da = m->jnt_dofadr[j], n = m->jnt_type[j] == 0 ? 6 : 1;
    memcpy(V[b], V[p], sizeof V[b]);
    memcpy(A[b], A[p], sizeof A[b]);
    for (int q = da; q < da + n; ++q) {
      double S[6], qd = d->qvel[q];
      if (d->dof_rot[q]) {
        for (int k = 0; k < 3; ++k) S[k] = d->dof_axis[q][k];
        cross3(S + 3, d->dof_anchor[q], d->dof_axis[q]);
      } else {
        S[0] = S[1] = S[2] = 0;
        for (int k = 0; k < 3; ++k) S[3 + k] = d->dof_axis[q][k];
      }
      /* Sdot = V x S for axes that move with the body; world-fixed axes (free-joint translations) have Sdot = 0 */
      int world_fixed = (m->jnt_type[j] == 0 && q < da + 3);
      if (!world_fixed) {
        /* for the free joint's rotational dofs V must already contain the translational part and the preceding rotations */
        double c1[3], c2[3], c3[3];
        cross3(c1, V[b], S);          /* w x s_w */
        cross3(c2, V[b], S + 3);      /* w x s_v */
        cross3(c3, V[b] + 3, S);      /* v x s_w */
        for (int k = 0; k < 3; ++k) { A[b][k] += c1[k] * qd; A[b][3 + k] += (c2[k] + c3[k]) * qd; }
      }
      for (int k = 0; k < 6; ++k) V[b][k] += S[k] * qd;
    }
  }
  for (int b = 1; b < nb; ++b) {
    double mass = m->body_mass[b], *c = d->xipos[b];
    if (mass <= 0) { memset(F[b], 0, sizeof F[b]); continue; }
    double Iw[9], w[3] = {V[b][0], V[b][1], V[b][2]}, vO[3] = {V[b][3], V[b][4], V[b][5]};
    double al[3] = {A[b][0], A[b][1], A[b][2]}, aO[3] = {A[b][3], A[b][4], A[b][5]};
    body_inertia_world(m, d, b, Iw);
    double t[3], p[3], L[3], f[3], n[3], Iw_w[3], Iw_al[3];
    cross3(t, w, c);
    for (int k = 0; k < 3; ++k) p[k] = mass * (vO[k] + t[k]);
    mulmatvec3(Iw_w, Iw, w);
    cross3(t, c, p);
    for (int k = 0; k < 3; ++k) L[k] = Iw_w[k] + t[k];
    cross3(t, al, c);
    for (int k = 0; k < 3; ++k) f[k] = mass * (aO[k] + t[k]);
    mulmatvec3(Iw_al, Iw, al);
    cross3(t, c, f);
    for (int k = 0; k < 3; ++k) n[k] = Iw_al[k] + t[k];
    double x1[3], x2[3], x3[3];
    cross3(x1, w, L);
    cross3(x2, vO, p);
    cross3(x3, w, p);
    for (int k = 0; k < 3; ++k) { F[b][k] = n[k] + x1[k] + x2[k]; F[b][3 + k] = f[k] + x3[k]; }
    d->flops += 120;
  }
  d->flops += (long long)nb * 90 + (long long)m->nv * 20;
  for (int b = nb - 1; b >= 1; --b) {
    int p = m->body_parent[b], j = m->body_jnt[b], da = m->jnt_dofadr[j], n = m->jnt_type[j] == 0 ? 6 : 1;
    for (int q = da; q < da + n; ++q) {
      if (d->dof_rot[q]) {
        double sv[3];
        cross3(sv, d->dof_anchor[q], d->dof_axis[q]);
        d->qfrc_bias[q] = dot3(d->dof_axis[q], F[b]) + dot3(sv, F[b] + 3);
      } else {
        d->qfrc_bias[q] = dot3(d->dof_axis[q], F[b] + 3);
      }
    }
    if (p > 0) for (int k = 0; k < 6; ++k) F[p][k] += F[b][k];
  }
}

/* ------------------------------------------------------------------------------------------ passive + actuation */
void mje_passive(const mjModelF *m, mjDataF *d) {
  for (int i = 0; i < m->nv; ++i) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  for (int b = 1; b < m->nbody; ++b) { /* joint springs (hinge / slide) */
    int j = m->body_jnt[b];
    if (m->jnt_stiffness[j] != 0 && m->jnt_type[j] >= 2)
      d->qfrc_passive[m->jnt_dofadr[j]] -= m->jnt_stiffness[j] * (d->qpos[m->jnt_qposadr[j]] - m->jnt_springref[j]);
  }
}

void mje_actuation(const mjModelF *m, mjDataF *d) {
  memset(d->qfrc_actuator, 0, sizeof(double) * m->nv);
  for (int u = 0; u < m->nu; ++u) { /* <position kp>: force = kp * (clamp(ctrl) - q), then forcerange */
    double c = d->ctrl[u];
    if (m->act_ctrllimited[u]) c = c < m->act_ctrlrange[2 * u] ? m->act_ctrlrange[2 * u] : (c > m->act_ctrlrange[2 * u + 1] ? m->act_ctrlrange[2 * u + 1] : c);
    double f = m->act_kp[u] * c - m->act_kp[u] * d->qpos[m->act_qposadr[u]];
    if (m->act_forcelimited[u]) f = f < m->act_forcerange[2 * u] ? m->act_forcerange[2 * u] : (f > m->act_forcerange[2 * u + 1] ? m->act_forcerange[2 * u + 1] : f);
    d->qfrc_actuator[m->act_dof[u]] += f;
  }
}

/* dense Cholesky solve of (nv x nv) SPD system A x = b; A is overwritten by its factor */
static int chol_factor(double A[MJ_MAXV][MJ_MAXV], int n) {
  for (int j = 0; j < n; ++j) {
    double s = A[j][j];
    for (int k = 0; k < j; ++k) s -= A[j][k] * A[j][k];
    if (s <= MJMINVAL) return -1;
    A[j][j] = sqrt(s);
    for (int i = j + 1; i < n; ++i) {
      double t = A[i][j];
      for (int k = 0; k < j; ++k) t -= A[i][k] * A[j][k];
      A[i][j] = t / A[j][j];
    }
  }
  return 0;
}
static void chol_solve(double A[MJ_MAXV][MJ_MAXV], int n, double *x) {
  for (int i = 0; i < n; ++i) {
    double t = x[i];
    for (int k = 0; k < i; ++k) t -= A[i][k] * x[k];
    x[i] = t / A[i][i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double t = x[i];
    for (int k = i + 1; k < n; ++k) t -= A[k][i] * x[k];
    x[i] = t / A[i][i];
  }
}

/* ------------------------------------------------------------------------------------------ constraint rows */
/* impedance d(r) of solimp = (dmin, dmax, width, midpoint, power); engine_core_constraint.c: getimpedance */
static double impedance(const double *solimp, double pos, double margin) {
  double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  dmin = dmin < MJMINIMP ? MJMINIMP : (dmin > MJMAXIMP ? MJMAXIMP : dmin);
  dmax = dmax < MJMINIMP ? MJMINIMP : (dmax > MJMAXIMP ? MJMAXIMP : dmax);
  if (width < MJMINVAL) width = MJMINVAL;
  mid = mid < MJMINIMP ? MJMINIMP : (mid > MJMAXIMP ? MJMAXIMP : mid);
  if (power < 1) power = 1;
  if (dmin == dmax || width <= MJMINVAL) return 0.5 * (dmin + dmax);
  double x = fabs(pos - margin) / width, y;
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  return dmin + y * (dmax - dmin);
}

/* experiment switches (diagnostics only; all zero = the documented behaviour this file restates) */
double mje_opt[16];
void mje_set_opt(int i, double v) { if (i >= 0 && i < 16) mje_opt[i] = v; }

int mje_row_is_contact = 0;
static double mje_imp_pos = -1; /* >= 0: violation used for the impedance instead of the row's own (vector residuals) */
/* fill aref / R / D of row i from (solref, solimp, pos, margin, vel, diagApprox); mj_makeImpedance */
void mje_finish_row(const mjModelF *m, mjDataF *d, int i, const double *solref, const double *solimp, double margin, double diag) {
  double vel = 0;
  for (int k = 0; k < m->nv; ++k) vel += d->efc_J[i][k] * d->qvel[k];
  if (mje_row_is_contact && mje_opt[7] > 0 && (mje_opt[10] == 0 || mje_row_is_contact == 2)) d->efc_pos[i] *= mje_opt[7]; /* experiment: contact distance scale */
  double imp = mje_imp_pos >= 0 ? impedance(solimp, mje_imp_pos, margin) : impedance(solimp, d->efc_pos[i], margin);
  double dmax = solimp[1] < MJMINIMP ? MJMINIMP : (solimp[1] > MJMAXIMP ? MJMAXIMP : solimp[1]);
  double k, b;
  if (solref[0] > 0) {
    double tc = solref[0] < 2 * m->timestep ? 2 * m->timestep : solref[0]; /* refsafe */
    double dr = solref[1];
    k = 1 / (dmax * dmax * tc * tc * dr * dr);
    b = 2 / (dmax * tc);
  } else { /* direct (-stiffness, -damping) */
    k = -solref[0] / (dmax * dmax);
    b = -solref[1] / dmax;
  }
  d->flops += 2 * m->nv + 30;
  double R = (1 - imp) * diag / imp;
  if (R < MJMINVAL) R = MJMINVAL;
  d->efc_R[i] = R;
  d->efc_D[i] = 1 / R;
  d->efc_aref[i] = -b * vel - k * imp * (d->efc_pos[i] - margin);
  if (mje_row_is_contact && (mje_opt[10] == 0 || mje_row_is_contact == 2)) { /* experiment knobs for contact rows: [4] stiffness scale, [5] damping scale, [6] R scale; [10] = 1: only contacts with the table */
    const double ks = mje_opt[4] > 0 ? mje_opt[4] : 1.0, bs = mje_opt[5] > 0 ? mje_opt[5] : 1.0, rs = mje_opt[6] > 0 ? mje_opt[6] : 1.0;
    d->efc_aref[i] = -bs * b * vel - ks * k * imp * (d->efc_pos[i] - margin);
    d->efc_R[i] = R * rs;
    d->efc_D[i] = 1 / d->efc_R[i];
  }
}

void mje_collision(const mjModelF *m, mjDataF *d);
int mje_contact_rows(const mjModelF *m, mjDataF *d, int row);

void mje_make_constraints(const mjModelF *m, mjDataF *d) {
  int nv = m->nv, r = 0;
  static double jp[3][MJ_MAXV], jr[3][MJ_MAXV];
  /* --- equality: weld(mocap frame, frame on body), engine_core_constraint.c mj_instantiateEquality, mjEQ_WELD */
  for (int w = 0; w < m->nweld; ++w) {
    int b = m->weld_body[w];
    const double *rel = m->weld_relpose + 7 * w;
    double p1[3], q1[4], t[3], p0[3], cpos[6];
    mulmatvec3(t, d->xmat[b], m->weld_pos + 3 * w);
    for (int k = 0; k < 3; ++k) p1[k] = d->xpos[b][k] + t[k];
    mulquat(q1, d->xquat[b], m->weld_quat + 4 * w);
    /* body1 = mocap: anchor = mocap_pos + R(mocap_quat) * relpose_pos */
    double mq[4] = {d->mocap_quat[0], d->mocap_quat[1], d->mocap_quat[2], d->mocap_quat[3]}, mR[9];
    normquat(mq);
    quat2mat(mR, mq);
    mulmatvec3(t, mR, rel);
    for (int k = 0; k < 3; ++k) { p0[k] = d->mocap_pos[k] + t[k]; cpos[k] = p0[k] - p1[k]; }
    mje_body_jac(m, d, b, p1, jp, jr);
    double quat[4], quat1[4] = {q1[0], -q1[1], -q1[2], -q1[3]}, quat2[4];
    mulquat(quat, mq, rel + 3);
    mulquat(quat2, quat1, quat);
    for (int k = 0; k < 3; ++k) cpos[3 + k] = quat2[1 + k];
    for (int k = 0; k < 3; ++k)
      for (int c = 0; c < nv; ++c) d->efc_J[r + k][c] = -jp[k][c]; /* jac(body1) - jac(body2), body1 static */
    for (int c = 0; c < nv; ++c) {
      double ax[4] = {0, -jr[0][c], -jr[1][c], -jr[2][c]}, q2[4], q3[4];
      mulquat(q2, quat1, ax);
      mulquat(q3, q2, quat);
      for (int k = 0; k < 3; ++k) d->efc_J[r + 3 + k][c] = (mje_opt[2] > 0 ? mje_opt[2] : 0.5) * q3[1 + k];
      if (mje_opt[3] == 1) for (int k = 0; k < 3; ++k) d->efc_J[r + 3 + k][c] = -(mje_opt[2] > 0 ? mje_opt[2] : 1.0) * jr[k][c]; /* experiment: uncorrected world-frame angular velocity difference */
    }
    d->flops += (long long)nv * 60 + 120;
    /* vector residual: all six rows share the impedance of its Euclidean norm (getposdim) */
    double cnorm = sqrt(cpos[0]*cpos[0]+cpos[1]*cpos[1]+cpos[2]*cpos[2]+cpos[3]*cpos[3]+cpos[4]*cpos[4]+cpos[5]*cpos[5]);
    for (int k = 0; k < 6; ++k) {
      mje_imp_pos = mje_opt[0] == 1 ? -1 : cnorm;
      d->efc_pos[r + k] = cpos[k];
      d->efc_type[r + k] = 0;
      mje_finish_row(m, d, r + k, m->weld_solref + 2 * w, m->weld_solimp + 5 * w, 0.0, m->weld_invweight[2 * w + (k >= 3)]);
    }
    mje_imp_pos = -1;
    r += 6;
  }
  /* --- equality: joint1 - ref1 = poly(joint2 - ref2), mj_instantiateEquality, mjEQ_JOINT */
  for (int e = 0; e < m->neq; ++e) {
    int q1 = m->eq_qposadr[2 * e], q2 = m->eq_qposadr[2 * e + 1], d1 = m->eq_dofadr[2 * e], d2 = m->eq_dofadr[2 * e + 1];
    const double *c = m->eq_polycoef + 5 * e;
    double dif = d->qpos[q2] - m->qpos0[q2], pw = 1, cpos = d->qpos[q1] - m->qpos0[q1], deriv = 0;
    for (int k = 0; k < 5; ++k) {
      cpos -= c[k] * pw;
      if (k < 4) deriv += (k + 1) * c[k + 1] * pw;
      pw *= dif;
    }
    memset(d->efc_J[r], 0, sizeof(double) * nv);
    d->efc_J[r][d1] = 1;
    d->efc_J[r][d2] = -deriv;
    d->efc_pos[r] = cpos;
    d->efc_type[r] = 0;
    mje_finish_row(m, d, r, m->eq_solref + 2 * e, m->eq_solimp + 5 * e, 0.0, m->eq_invweight[e]);
    ++r;
  }
  /* --- dof friction loss, mj_instantiateFriction: one row per dof with frictionloss > 0, residual 0 */
  if (m->version == 2)
    for (int i = 0; i < nv; ++i) {
      if (!(m->dof_frictionloss[i] > 0)) continue;
      memset(d->efc_J[r], 0, sizeof(double) * nv);
      d->efc_J[r][i] = 1;
      d->efc_pos[r] = 0;
      d->efc_type[r] = 4;
      d->efc_floss[r] = m->dof_frictionloss[i];
      mje_finish_row(m, d, r, m->dof_solref_friction + 2 * i, m->dof_solimp_friction + 5 * i, 0.0, m->dof_invweight0[i]);
      ++r;
    }
  /* --- joint limits (hinge / slide), mj_instantiateLimit */
  for (int b = 1; b < m->nbody; ++b) {
    int j = m->body_jnt[b];
    if (!m->jnt_limited[j] || m->jnt_type[j] < 2) continue;
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    for (int side = 0; side < 2; ++side) {
      double dist = side == 0 ? d->qpos[qa] - m->jnt_range[2 * j] : m->jnt_range[2 * j + 1] - d->qpos[qa];
      if (dist < m->jnt_margin[j]) {
        memset(d->efc_J[r], 0, sizeof(double) * nv);
        d->efc_J[r][da] = side == 0 ? 1 : -1;
        d->efc_pos[r] = dist;
        d->efc_type[r] = 1;
        mje_finish_row(m, d, r, m->jnt_solref + 2 * j, m->jnt_solimp + 5 * j, m->jnt_margin[j], m->dof_invweight0[da]);
        ++r;
      }
    }
  }
  /* --- contacts */
  r = mje_contact_rows(m, d, r);
  d->nefc = r;
}

/* ------------------------------------------------------------------------------------------ convex solve (Newton)
 * minimise 0.5 (a - a_s)' M (a - a_s) + sum_i s_i(J_i a - aref_i)   (MuJoCo "Computation": primal problem) */
static double row_cost_update(const mjModelF *m, mjDataF *d, const double *jar, double *force, int *active, double hcone[][36]);

static double total_cost(const mjModelF *m, mjDataF *d, const double *a, double *jar, double *Ma_out) {
  int nv = m->nv;
  double cost = 0;
  for (int i = 0; i < nv; ++i) {
    double s = 0;
    for (int j = 0; j < nv; ++j) s += d->M[i][j] * (a[j] - d->qacc_smooth[j]);
    if (Ma_out) Ma_out[i] = s;
    cost += 0.5 * s * (a[i] - d->qacc_smooth[i]);
  }
  for (int r = 0; r < d->nefc; ++r) {
    double s = -d->efc_aref[r];
    for (int j = 0; j < nv; ++j) s += d->efc_J[r][j] * a[j];
    jar[r] = s;
  }
  return cost;
}

void mje_solve(const mjModelF *m, mjDataF *d) {
  int nv = m->nv, ne = d->nefc;
  static double H[MJ_MAXV][MJ_MAXV], hcone[MJ_MAXCON][36];
  static double jar[MJ_MAXEFC], jv[MJ_MAXEFC], force[MJ_MAXEFC], jar_t[MJ_MAXEFC], force_t[MJ_MAXEFC];
  static int active[MJ_MAXEFC], active_t[MJ_MAXEFC];
  double a[MJ_MAXV], Ma[MJ_MAXV], grad[MJ_MAXV], dir[MJ_MAXV], at[MJ_MAXV];
  /* qacc_smooth = M^-1 qfrc_smooth */
  for (int i = 0; i < nv; ++i) memcpy(H[i], d->M[i], sizeof(double) * nv);
  chol_factor(H, nv);
  memcpy(d->qacc_smooth, d->qfrc_smooth, sizeof(double) * nv);
  chol_solve(H, nv, d->qacc_smooth);
  d->flops += (long long)nv * nv * nv / 3 + 2LL * nv * nv + 2LL * (2LL * nv * nv + 2LL * ne * nv + 6LL * ne);
  if (ne == 0) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
    memset(d->qfrc_constraint, 0, sizeof(double) * nv);
    d->solver_iter = 0;
    return;
  }
  /* warm start: previous qacc if it has the lower cost, else qacc_smooth (mj_fwdConstraint) */
  double c_ws = total_cost(m, d, d->qacc_warmstart, jar, 0) + row_cost_update(m, d, jar, force, active, 0);
  double c_sm = total_cost(m, d, d->qacc_smooth, jar, 0) + row_cost_update(m, d, jar, force, active, 0);
  memcpy(a, c_ws < c_sm ? d->qacc_warmstart : d->qacc_smooth, sizeof(double) * nv);
  double meaninertia = 0;
  for (int i = 0; i < nv; ++i) meaninertia += d->M[i][i];
  meaninertia /= nv;
  double scale = 1.0 / (meaninertia * (nv > 1 ? nv : 1));
  int it;
  for (it = 0; it < 200; ++it) {
    double cost = total_cost(m, d, a, jar, Ma) + row_cost_update(m, d, jar, force, active, hcone);
    /* gradient = M (a - a_s) - J' f */
    double gn = 0;
    for (int i = 0; i < nv; ++i) {
      double s = Ma[i];
      for (int r = 0; r < ne; ++r) s -= d->efc_J[r][i] * force[r];
      grad[i] = s;
      gn += s * s;
    }
    if (sqrt(gn) * scale < 1e-14) break;
    /* Hessian = M + J' D_active J (+ cone blocks) */
    for (int i = 0; i < nv; ++i) memcpy(H[i], d->M[i], sizeof(double) * nv);
    int con = 0;
    for (int r = 0; r < ne; ++r) {
      if (d->efc_type[r] == 2) {
        int dim = d->efc_dim[r];
        if (active[r] == 2) { /* cone (middle) zone: dense dim x dim block */
          for (int p = 0; p < dim; ++p)
            for (int q = 0; q < dim; ++q) {
              double hpq = hcone[con][6 * p + q];
              if (hpq == 0) continue;
              for (int i = 0; i < nv; ++i) {
                double ji = d->efc_J[r + p][i] * hpq;
                if (ji == 0) continue;
                for (int j = 0; j < nv; ++j) H[i][j] += ji * d->efc_J[r + q][j];
              }
            }
        } else if (active[r] == 1) { /* bottom zone: plain quadratic on all rows */
          for (int p = 0; p < dim; ++p)
            for (int i = 0; i < nv; ++i) {
              double ji = d->efc_J[r + p][i] * d->efc_D[r + p];
              if (ji == 0) continue;
              for (int j = 0; j < nv; ++j) H[i][j] += ji * d->efc_J[r + p][j];
            }
        }
        ++con;
        r += dim - 1;
        continue;
      }
      if (!active[r]) continue;
      for (int i = 0; i < nv; ++i) {
        double ji = d->efc_J[r][i] * d->efc_D[r];
        if (ji == 0) continue;
        for (int j = 0; j < nv; ++j) H[i][j] += ji * d->efc_J[r][j];
      }
    }
    d->flops += (long long)ne * nv * nv + (long long)nv * nv * nv / 3 + 4LL * nv * nv + 6LL * ne * nv + 8LL * 10 * ne;
    if (chol_factor(H, nv)) break;
    for (int i = 0; i < nv; ++i) dir[i] = -grad[i];
    chol_solve(H, nv, dir);
    /* line search along dir: the cost is convex along the ray; 1-D Newton with bracketing on its derivative */
    for (int r = 0; r < ne; ++r) {
      double s = 0;
      for (int j = 0; j < nv; ++j) s += d->efc_J[r][j] * dir[j];
      jv[r] = s;
    }
    double lo = 0, hi = -1, alpha = 1, best_alpha = 0, best_cost = cost;
    for (int ls = 0; ls < 60; ++ls) {
      for (int i = 0; i < nv; ++i) at[i] = a[i] + alpha * dir[i];
      double c = total_cost(m, d, at, jar_t, Ma) + row_cost_update(m, d, jar_t, force_t, active_t, 0);
      if (c < best_cost) { best_cost = c; best_alpha = alpha; }
      /* derivative along dir: dir' (M (at - a_s)) - sum f_r jv_r */
      double der = 0;
      for (int i = 0; i < nv; ++i) der += dir[i] * Ma[i];
      for (int r = 0; r < ne; ++r) der -= force_t[r] * jv[r];
      if (fabs(der) * scale < 1e-16) break;
      if (der > 0) hi = alpha; else lo = alpha;
      if (hi < 0) alpha *= 2; else alpha = 0.5 * (lo + hi);
      if (hi > 0 && hi - lo < 1e-14 * (hi > 1 ? hi : 1)) break;
    }
    if (best_alpha == 0) break;
    double step = 0;
    for (int i = 0; i < nv; ++i) { a[i] += best_alpha * dir[i]; step += dir[i] * dir[i]; }
    if ((cost - best_cost) * scale < 1e-17 && best_alpha * sqrt(step) * scale < 1e-15) { ++it; break; }
  }
  d->solver_iter = it;
  total_cost(m, d, a, jar, 0);
  row_cost_update(m, d, jar, force, active, 0);
  memcpy(d->qacc, a, sizeof(double) * nv);
  memcpy(d->efc_force, force, sizeof(double) * ne);
  for (int i = 0; i < nv; ++i) {
    double s = 0;
    for (int r = 0; r < ne; ++r) s += d->efc_J[r][i] * force[r];
    d->qfrc_constraint[i] = s;
  }
}

/* per-row cost s_i(jar) and force f_i = -ds/djar; for elliptic contacts the three-zone cone cost.
 * active: 0 inactive, 1 quadratic, 2 cone (middle) zone.  hcone (optional): per-contact Hessian blocks wrt jar. */
static double row_cost_update(const mjModelF *m, mjDataF *d, const double *jar, double *force, int *active, double hcone[][36]) {
  (void)m;
  double cost = 0;
  int con = 0;
  for (int r = 0; r < d->nefc; ++r) {
    if (d->efc_type[r] == 0) {
      force[r] = -d->efc_D[r] * jar[r];
      active[r] = 1;
      cost += 0.5 * d->efc_D[r] * jar[r] * jar[r];
    } else if (d->efc_type[r] == 1) {
      if (jar[r] < 0) { force[r] = -d->efc_D[r] * jar[r]; active[r] = 1; cost += 0.5 * d->efc_D[r] * jar[r] * jar[r]; }
      else { force[r] = 0; active[r] = 0; }
    } else if (d->efc_type[r] == 4) {
      /* friction loss (PrimalUpdateConstraint, mjCNSTR_FRICTION_*): quadratic inside |jar| < R f, linear outside */
      double f = d->efc_floss[r], rf = d->efc_R[r] * f;
      if (jar[r] <= -rf) { force[r] = f; active[r] = 0; cost += -0.5 * rf * f - f * jar[r]; }
      else if (jar[r] >= rf) { force[r] = -f; active[r] = 0; cost += -0.5 * rf * f + f * jar[r]; }
      else { force[r] = -d->efc_D[r] * jar[r]; active[r] = 1; cost += 0.5 * d->efc_D[r] * jar[r] * jar[r]; }
    } else {
      /* elliptic cone (MuJoCo engine_solver.c: PrimalUpdateConstraint / HessianCone).  Scaled variables:
       * U0 = jar0 * mu, Uj = jar_j * fri_j;  N = U0, T = |U_1..| */
      int dim = d->efc_dim[r];
      double mu = d->efc_mu[r], *fri = d->efc_fri[r];
      double U[6], N, T = 0;
      U[0] = jar[r] * mu;
      for (int j = 1; j < dim; ++j) { U[j] = jar[r + j] * fri[j - 1]; T += U[j] * U[j]; }
      T = sqrt(T);
      N = U[0];
      double Dm = d->efc_D[r] / (mu * mu * (1 + mu * mu));
      if (hcone) memset(hcone[con], 0, sizeof(double) * 36);
      if (N >= mu * T || (T <= 0 && N >= 0)) { /* top zone: no force */
        for (int j = 0; j < dim; ++j) { force[r + j] = 0; active[r + j] = 0; }
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) { /* bottom zone: quadratic in every row */
        for (int j = 0; j < dim; ++j) {
          force[r + j] = -d->efc_D[r + j] * jar[r + j];
          active[r + j] = 1;
          cost += 0.5 * d->efc_D[r + j] * jar[r + j] * jar[r + j];
        }
      } else { /* middle zone */
        double NT = N - mu * T;
        cost += 0.5 * Dm * NT * NT;
        force[r] = -Dm * NT * mu;
        for (int j = 1; j < dim; ++j) force[r + j] = -force[r] / T * U[j] * fri[j - 1];
        for (int j = 0; j < dim; ++j) active[r + j] = 2;
        if (hcone) {
          /* Hessian of 0.5 Dm (N - mu T)^2 wrt U, then scaled back to jar */
          double *Hc = hcone[con];
          double scl[6];
          scl[0] = mu;
          for (int j = 1; j < dim; ++j) scl[j] = fri[j - 1];
          Hc[0] = 1;
          for (int j = 1; j < dim; ++j) Hc[j] = Hc[6 * j] = -mu * U[j] / T;
          for (int p = 1; p < dim; ++p)
            for (int q = 1; q < dim; ++q)
              Hc[6 * p + q] = mu * N / (T * T * T) * U[p] * U[q] + (p == q ? mu * mu - mu * N / T : 0);
          for (int p = 0; p < dim; ++p)
            for (int q = 0; q < dim; ++q) Hc[6 * p + q] *= Dm * scl[p] * scl[q];
        }
      }
      ++con;
      r += dim - 1;
    }
  }
  return cost;
}

/* ------------------------------------------------------------------------------------------ mj_forward + mj_Euler */
void mje_forward(const mjModelF *m, mjDataF *d) {
  mje_kinematics(m, d);
  mje_mass_matrix(m, d);
  if (mje_opt[1] == 1) for (int i = 0; i < m->nv; ++i) d->M[i][i] += m->timestep * m->dof_damping[i]; /* experiment: solver sees M + hB */
  mje_collision(m, d);
  mje_make_constraints(m, d);
  mje_passive(m, d);
  mje_bias(m, d);
  mje_actuation(m, d);
  for (int i = 0; i < m->nv; ++i) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
  mje_solve(m, d);
}

static void quat_integrate(double *q, const double *w_local, double h) {
  double ang = sqrt(dot3(w_local, w_local)) * h;
  if (ang < MJMINVAL) return;
  double ax[3] = {w_local[0] * h / ang, w_local[1] * h / ang, w_local[2] * h / ang}, dq[4];
  axisangle2quat(dq, ax, ang);
  mulquat(q, q, dq);
  normquat(q);
}

void mje_step(const mjModelF *m, mjDataF *d) {
  int nv = m->nv;
  double h = m->timestep;
  mje_forward(m, d);
  /* mj_Euler: implicit in joint damping when any damping > 0 */
  static double H[MJ_MAXV][MJ_MAXV];
  double qacc[MJ_MAXV];
  int damped = 0;
  for (int i = 0; i < nv; ++i) damped |= m->dof_damping[i] > 0;
  if (damped && mje_opt[1] != 1) {
    for (int i = 0; i < nv; ++i) { memcpy(H[i], d->M[i], sizeof(double) * nv); H[i][i] += h * m->dof_damping[i]; }
    for (int i = 0; i < nv; ++i) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    chol_factor(H, nv);
    chol_solve(H, nv, qacc);
    d->flops += (long long)nv * nv * nv / 3 + 2LL * nv * nv + 6LL * nv;
  } else {
    memcpy(qacc, d->qacc, sizeof(double) * nv);
  }
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * nv);
  for (int i = 0; i < nv; ++i) d->qvel[i] += h * qacc[i];
  for (int b = 1; b < m->nbody; ++b) {
    int j = m->body_jnt[b], qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == 0) {
      for (int k = 0; k < 3; ++k) d->qpos[qa + k] += h * d->qvel[da + k];
      quat_integrate(d->qpos + qa + 3, d->qvel + da + 3, h);
    } else {
      d->qpos[qa] += h * d->qvel[da];
    }
  }
  d->time += h;
}

long long mje_sizeof_data(void) { return (long long)sizeof(mjDataF); }

/* ------------------------------------------------------------------------------------------ accessors for the ctypes binding */
double *mje_ptr(mjDataF *d, int what) {
  switch (what) {
    case 0: return d->qpos;
    case 1: return d->qvel;
    case 2: return d->ctrl;
    case 3: return d->mocap_pos;
    case 4: return d->mocap_quat;
    case 5: return &d->xpos[0][0];
    case 6: return &d->site_xpos[0][0];
    case 7: return &d->geom_xpos[0][0];
    case 8: return d->qacc;
    case 9: return d->qacc_warmstart;
    case 10: return &d->M[0][0];
    case 11: return d->qfrc_bias;
    case 12: return d->efc_force;
    case 13: return d->efc_pos;
    case 14: return d->qfrc_constraint;
    case 15: return &d->xmat[0][0];
    case 16: return d->qfrc_smooth;
    case 17: return d->qacc_smooth;
    case 18: return &d->con_pos[0][0];
    case 19: return d->con_dist;
    case 20: return &d->con_frame[0][0];
  }
  return 0;
}
int mje_int(const mjDataF *d, int what) {
  switch (what) {
    case 0: return d->nefc;
    case 1: return d->ncon;
    case 2: return d->solver_iter;
  }
  return -1;
}
long long mje_flops(const mjDataF *d) { return d->flops; }
int mje_con_geoms(const mjDataF *d, int k) { return k < d->ncon ? d->con_geom1[k] * 1000 + d->con_geom2[k] : -1; }
void mje_multi_step(const mjModelF *m, mjDataF *d, int n) { for (int i = 0; i < n; ++i) mje_step(m, d); }
