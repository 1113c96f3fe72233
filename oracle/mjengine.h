/*
 * oracle/mjengine.h -- fp64 CPU restatement of the MuJoCo 2.1 `mj_step` subset the EARL Sawyer tasks exercise.
 * TEST INFRASTRUCTURE ONLY (see mjengine.c).
 */
#ifndef EARL_ORACLE_MJENGINE_H_
#define EARL_ORACLE_MJENGINE_H_
#include <stdint.h>

#define MJ_MAXB 24    /* fused bodies */
#define MJ_MAXV 32    /* dofs */
#define MJ_MAXQ 40
#define MJ_MAXG 160   /* geoms */
#define MJ_MAXS 40    /* sites */
#define MJ_MAXCON 64  /* contacts */
#define MJ_MAXEFC 320 /* constraint rows */

/* field order == earl_benchmark_b200/mjcf/compile.py: Model.FIELDS */
typedef struct {
  int nbody, nq, nv, ngeom, nsite, nu, nweld, nhullvert, iterations, cone_elliptic;
  double timestep, tolerance, impratio;
  const double *gravity;
  const int *body_parent;
  const double *body_pos, *body_quat, *body_mass, *body_ipos, *body_inertia;
  const int *body_jnt;
  const int *jnt_type, *jnt_body, *jnt_qposadr, *jnt_dofadr;
  const double *jnt_pos, *jnt_axis;
  const int *jnt_limited;
  const double *jnt_range, *jnt_margin, *jnt_solref, *jnt_solimp, *jnt_stiffness, *jnt_springref;
  const int *dof_body;
  const double *dof_damping, *dof_armature, *dof_frictionloss, *dof_invweight0, *qpos0;
  const int *geom_body, *geom_type;
  const double *geom_size, *geom_pos, *geom_quat;
  const int *geom_contype, *geom_conaffinity, *geom_condim, *geom_priority;
  const double *geom_friction, *geom_margin, *geom_gap, *geom_solref, *geom_solimp, *geom_solmix, *geom_invweight0,
      *geom_rbound;
  const int *geom_hulladr, *geom_hullnum, *geom_srcbody, *geom_srcparent;
  const double *hull_vert;
  const int *site_body;
  const double *site_pos, *site_quat;
  const int *act_dof, *act_qposadr;
  const double *act_kp, *act_ctrlrange;
  const int *act_ctrllimited;
  const double *act_forcerange;
  const int *act_forcelimited;
  const int *weld_body;
  const double *weld_pos, *weld_quat, *weld_relpose, *weld_solref, *weld_solimp, *weld_invweight;
  const double *mocap_pos0, *mocap_quat0;
  /* blob version 2 (models with joint equalities / friction loss: the kitchen); neq = 0 and defaults otherwise */
  int version, neq;
  const int *eq_qposadr, *eq_dofadr;
  const double *eq_polycoef, *eq_solref, *eq_solimp, *eq_invweight, *dof_solref_friction, *dof_solimp_friction;
  void *blob; /* owned copy of the serialized model */
} mjModelF;

typedef struct {
  /* state */
  double qpos[MJ_MAXQ], qvel[MJ_MAXV], qacc_warmstart[MJ_MAXV], ctrl[8], mocap_pos[3], mocap_quat[4];
  double time;
  /* position-dependent */
  double xpos[MJ_MAXB][3], xmat[MJ_MAXB][9], xquat[MJ_MAXB][4], xipos[MJ_MAXB][3];
  double dof_axis[MJ_MAXV][3], dof_anchor[MJ_MAXV][3];
  int dof_rot[MJ_MAXV];
  double geom_xpos[MJ_MAXG][3], geom_xmat[MJ_MAXG][9];
  double site_xpos[MJ_MAXS][3], site_xmat[MJ_MAXS][9];
  double M[MJ_MAXV][MJ_MAXV];
  /* forces */
  double qfrc_bias[MJ_MAXV], qfrc_passive[MJ_MAXV], qfrc_actuator[MJ_MAXV], qfrc_smooth[MJ_MAXV], qacc_smooth[MJ_MAXV];
  double qfrc_constraint[MJ_MAXV], qacc[MJ_MAXV];
  /* constraints */
  int nefc, ncon;
  double efc_J[MJ_MAXEFC][MJ_MAXV], efc_pos[MJ_MAXEFC], efc_aref[MJ_MAXEFC], efc_R[MJ_MAXEFC], efc_D[MJ_MAXEFC],
      efc_force[MJ_MAXEFC];
  int efc_type[MJ_MAXEFC];   /* 0 equality, 1 limit / frictionless / pyramid edge (force >= 0), 2 elliptic normal row (followed by
                                dim-1 friction rows of type 3), 4 dof friction loss (|force| <= efc_floss) */
  double efc_floss[MJ_MAXEFC];
  int efc_dim[MJ_MAXEFC];    /* for type 2: contact dimension */
  double efc_mu[MJ_MAXEFC];  /* for type 2: regularised friction coefficient */
  double efc_fri[MJ_MAXEFC][5];
  /* contacts */
  double con_pos[MJ_MAXCON][3], con_frame[MJ_MAXCON][9], con_dist[MJ_MAXCON], con_friction[MJ_MAXCON][5];
  double con_solref[MJ_MAXCON][2], con_solimp[MJ_MAXCON][5], con_margin[MJ_MAXCON];
  int con_geom1[MJ_MAXCON], con_geom2[MJ_MAXCON], con_dim[MJ_MAXCON];
  /* diagnostics */
  int solver_iter;
  long long flops; /* running count of fp64 multiply-adds in the engine (algorithmic flops) */
} mjDataF;

#endif
