// mj_model_host.hpp -- host side of the articulated-body engine: reads the serialized structure-of-arrays model
// (earl_benchmark_b200/mjcf/compile.py: Model.to_blob, fields in Model.FIELDS order) and fills the fixed-size
// fp32 `earl::mj::Model` the kernels read.  Plain C++, no CUDA.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <algorithm>
#include <cmath>

#include "mj_collide.cuh"
#include "mj_engine.cuh"

namespace earl {
namespace mj {

// task constants that are not part of the MJCF (metaworld SawyerXYZEnv / EARL env classes)
struct TaskSpec {
  int32_t frame_skip;      // SawyerXYZEnv frame_skip = 5
  int32_t hand_site;       // site 'body:hand'   (get_endeff_pos = body 'hand' xpos)
  int32_t ree_site;        // site 'rightEndEffector'
  int32_t lee_site;        // site 'leftEndEffector'
  int32_t obj_geom;        // door: geom 'handle' (sawyer_door._get_pos_objects); -1 if the object is a site
  int32_t obj_site;        // peg: site 'pegHead' (sawyer_peg.py:186-187); -1 otherwise
  int32_t max_newton;      // device cap on Newton iterations per substep
  int32_t obj_qpos_count;  // qpos entries of the object joint written by a reset (door angle: 1, peg xyz: 3)
  float mocap_low[3];      // hand_low  (sawyer_peg.py:66 / metaworld SawyerDoorEnvV2)
  float mocap_high[3];     // hand_high
  float action_scale;      // 1/100
  float success_radius;    // 0.02 door (sawyer_door.py:177), 0.05 peg (sawyer_peg.py:305)
  float obj_init_pos[3];   // dense door reward (sawyer_door.py:150)
  float hand_init_pos[3];  // dense door reward (sawyer_door.py:156)
  int32_t grasp_site, lpad_site, rpad_site, corner_site[4];  // dense peg reward (sawyer_peg.py:231-299); -1 when unused
};

class BlobReader {
 public:
  BlobReader(const void* blob, size_t nbytes) : p_(static_cast<const uint8_t*>(blob)), end_(p_ + nbytes) {}
  bool header() {
    if (end_ - p_ < 8) return false;
    int32_t h[2];
    memcpy(h, p_, 8);
    p_ += 8;
    version = h[1];
    return h[0] == 0x4C444D45 && (h[1] == 1 || h[1] == 2);
  }
  int version = 0;  // 1: door / peg; 2: + joint equalities and friction-loss parameters (kitchen)
  // next field as doubles / ints; returns element count or -1
  template <typename T>
  long field(std::vector<T>* out, int elem) {
    if (end_ - p_ < 4) return -1;
    int32_t ndim;
    memcpy(&ndim, p_, 4);
    if (ndim < 0 || ndim > 4 || end_ - p_ < 4 * (1 + ndim)) return -1;
    long n = 1;
    for (int i = 0; i < ndim; ++i) {
      int32_t d;
      memcpy(&d, p_ + 4 * (1 + i), 4);
      n *= d;
    }
    size_t hb = 4 * (size_t)(1 + ndim);
    if (hb % 8) hb += 4;
    p_ += hb;
    size_t db = (size_t)n * elem;
    size_t padded = db % 8 ? db + 8 - db % 8 : db;
    if ((size_t)(end_ - p_) < padded) return -1;
    out->resize((size_t)n);
    if (n) memcpy(out->data(), p_, db);
    p_ += padded;
    return n;
  }
  bool done() const { return p_ == end_; }

 private:
  const uint8_t* p_;
  const uint8_t* end_;
};

// hull vertices of mesh geoms are kept separately (too large for the fixed-size struct)
struct HostModel {
  Model m;
  std::vector<float> hull_vert;  // [nhull][3]
  double mocap_pos0[3] = {0, 0, 0}, mocap_quat0[4] = {1, 0, 0, 0};  // mocap pose after sim.reset()
};

// Compile-time pruning of candidate pairs (static geom gs, geom gm on a body whose only ancestor is the world and whose
// single hinge / slide joint is limited): the joint range, widened by 0.15, is swept in steps that move no point of
// the body by more than ~1 mm; if no sample brings the pair within margin + 4 mm the pair can never produce a contact
// and is dropped.  Box-box pairs are swept with the separating-axis test itself, everything else with the
// bounding-sphere-vs-box / sphere-sphere / sphere-plane bound the runtime broad phase uses.
inline bool never_touch(const Model& m, int gs, int gm) {
  if (m.geom_body[gs] != 0) return false;
  const int b = m.geom_body[gm];
  if (b == 0 || m.body_parent[b] != 0) return false;
  const int j = m.body_jnt[b];
  if (m.jnt_type[j] < 2 || !m.jnt_limited[j]) return false;
  const double lo = m.jnt_range[j][0] - 0.15, hi = m.jnt_range[j][1] + 0.15;
  const double slack = 0.004, margin = std::fmax(m.geom_margin[gs], m.geom_margin[gm]) + slack;
  const double reach = 1.0;  // bound on the distance of any body point from the joint (metres)
  const int nsamp = (int)((hi - lo) * reach / 0.001) + 2;
  Work* w = new Work();
  memset(w, 0, sizeof(Work));
  bool clear = true;
  for (int s = 0; s < nsamp && clear; ++s) {
    const double q = lo + (hi - lo) * s / (nsamp - 1);
    for (int k = 0; k < m.nq; ++k) w->qpos[k] = m.qpos0[k];
    w->qpos[m.jnt_qposadr[j]] = (real)q;
    // pose of body b alone (its parent is the world): same formulas as kinematics()
    real quat[4], R[9], pos[3];
    for (int k = 0; k < 4; ++k) quat[k] = m.body_quat[b][k];
    for (int k = 0; k < 3; ++k) pos[k] = m.body_pos[b][k];
    quat2mat(R, quat);
    if (m.jnt_type[j] == 3) {
      real anchor[3], t[3], qj[4];
      mulmatvec3(t, R, m.jnt_pos[j]);
      for (int k = 0; k < 3; ++k) anchor[k] = pos[k] + t[k];
      const real half = 0.5f * (real)(q - m.jnt_qpos0[j]);
      qj[0] = cosf(half);
      for (int k = 0; k < 3; ++k) qj[1 + k] = m.jnt_axis[j][k] * sinf(half);
      mulquat(quat, quat, qj);
      normquat(quat);
      quat2mat(R, quat);
      mulmatvec3(t, R, m.jnt_pos[j]);
      for (int k = 0; k < 3; ++k) pos[k] = anchor[k] - t[k];
    } else {
      real axis[3];
      mulmatvec3(axis, R, m.jnt_axis[j]);
      for (int k = 0; k < 3; ++k) pos[k] += axis[k] * (real)(q - m.jnt_qpos0[j]);
    }
    real gp[3], gR[9], t[3];
    mulmatvec3(t, R, m.geom_pos[gm]);
    for (int k = 0; k < 3; ++k) gp[k] = pos[k] + t[k];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) gR[3 * a + c] = R[3 * a] * m.geom_mat[gm][c] + R[3 * a + 1] * m.geom_mat[gm][3 + c] + R[3 * a + 2] * m.geom_mat[gm][6 + c];
    const int ts = m.geom_type[gs], tm = m.geom_type[gm];
    const real* sp = m.geom_pos[gs];
    const real* sR = m.geom_mat[gs];
    if (ts == GEOM_BOX && tm == GEOM_BOX) {
      NarrowScratch* S = reinterpret_cast<NarrowScratch*>(&w->H[0][0]);
      if (box_box(sp, sR, m.geom_size[gs], gp, gR, m.geom_size[gm], (real)margin, S->rc, S) > 0) clear = false;
    } else if (ts == GEOM_PLANE) {
      const real n[3] = {sR[2], sR[5], sR[8]};
      real rel[3];
      sub3(rel, gp, sp);
      if (dot3(rel, n) <= m.geom_rbound[gm] + margin) clear = false;
    } else if (ts == GEOM_BOX) {
      const double r = m.geom_rbound[gm] + margin;
      if (point_box_dist2(gp, sp, sR, m.geom_size[gs]) <= r * r) clear = false;
    } else if (tm == GEOM_BOX) {
      const double r = m.geom_rbound[gs] + margin;
      if (point_box_dist2(sp, gp, gR, m.geom_size[gm]) <= r * r) clear = false;
    } else {
      real rel[3];
      sub3(rel, gp, sp);
      const double r = m.geom_rbound[gs] + m.geom_rbound[gm] + margin;
      if (dot3(rel, rel) <= r * r) clear = false;
    }
  }
  delete w;
  return clear;
}

inline bool build_model(const void* blob, size_t nbytes, const TaskSpec& task, HostModel* out, std::string* err) {
  BlobReader rd(blob, nbytes);
  auto bad = [&](const char* what) { if (err) *err = what; return false; };
  if (!rd.header()) return bad("model blob: bad magic / version");
  std::vector<int32_t> I;
  std::vector<double> D;
  Model& m = out->m;
  memset(&m, 0, sizeof(m));
#define RI(dst) do { if (rd.field(&I, 4) != 1) return bad("model blob: scalar int field"); dst = I[0]; } while (0)
#define RD(dst) do { if (rd.field(&D, 8) != 1) return bad("model blob: scalar double field"); dst = (real)D[0]; } while (0)
  int nhullvert = 0, cone_elliptic = 0;
  RI(m.nbody); RI(m.nq); RI(m.nv); RI(m.ngeom); RI(m.nsite); RI(m.nu); RI(m.nweld); RI(nhullvert); RI(m.iterations);
  RI(cone_elliptic);
  RD(m.timestep);
  if (rd.field(&D, 8) != 1) return bad("tolerance");  // <option tolerance>: the fp32 solver uses its own termination rule
  RD(m.impratio);
  m.cone_elliptic = cone_elliptic;
#if !defined(MJ_CAPSET_KITCHEN)
  if (!cone_elliptic) return bad("model blob: only elliptic friction cones are built in this capacity set");
#endif
  if (m.nbody > MAXB || m.nv > MAXV || m.nq > MAXQ || m.ngeom > MAXG || m.nsite > MAXS || m.nu > MAXU || m.nweld > MAXW)
    return bad("model blob: model exceeds the engine's fixed sizes");
  const int nb = m.nbody, nv = m.nv, ng = m.ngeom, ns = m.nsite, nu = m.nu, nw = m.nweld;
#define VI(n, expr) do { if (rd.field(&I, 4) != (long)(n)) return bad("model blob: int array size"); for (long k = 0; k < (long)(n); ++k) { expr; } } while (0)
#define VD(n, expr) do { if (rd.field(&D, 8) != (long)(n)) return bad("model blob: double array size"); for (long k = 0; k < (long)(n); ++k) { expr; } } while (0)
  VD(3, m.gravity[k] = (real)D[k]);
  VI(nb, m.body_parent[k] = I[k]);
  VD(nb * 3, m.body_pos[k / 3][k % 3] = (real)D[k]);
  VD(nb * 4, m.body_quat[k / 4][k % 4] = (real)D[k]);
  VD(nb, m.body_mass[k] = (real)D[k]);
  VD(nb * 3, m.body_ipos[k / 3][k % 3] = (real)D[k]);
  VD(nb * 6, m.body_inertia[k / 6][k % 6] = (real)D[k]);
  VI(nb, m.body_jnt[k] = I[k]);
  // joints: count from the next field
  long nj = rd.field(&I, 4);
  if (nj < 0 || nj > MAXJ) return bad("model blob: joint count");
  m.njnt = (int)nj;
  for (long k = 0; k < nj; ++k) m.jnt_type[k] = I[k];
  VI(nj, m.jnt_body[k] = I[k]);
  VI(nj, m.jnt_qposadr[k] = I[k]);
  VI(nj, m.jnt_dofadr[k] = I[k]);
  VD(nj * 3, m.jnt_pos[k / 3][k % 3] = (real)D[k]);
  VD(nj * 3, m.jnt_axis[k / 3][k % 3] = (real)D[k]);
  VI(nj, m.jnt_limited[k] = I[k]);
  VD(nj * 2, m.jnt_range[k / 2][k % 2] = (real)D[k]);
  VD(nj, m.jnt_margin[k] = (real)D[k]);
  VD(nj * 2, m.jnt_solref[k / 2][k % 2] = (real)D[k]);
  VD(nj * 5, m.jnt_solimp[k / 5][k % 5] = (real)D[k]);
  VD(nj, m.jnt_stiffness[k] = (real)D[k]);
  VD(nj, m.jnt_springref[k] = (real)D[k]);
  VI(nv, m.dof_body[k] = I[k]);
  VD(nv, m.dof_damping[k] = (real)D[k]);
  VD(nv, m.dof_armature[k] = (real)D[k]);
  VD(nv, m.dof_frictionloss[k] = (real)D[k]);
#if !defined(MJ_CAPSET_KITCHEN)
  for (int k = 0; k < nv; ++k) if (D[k] != 0.0) return bad("model blob: frictionloss is not built in this capacity set");
#endif
  for (int k = 0; k < nv; ++k) {  // defaults (blob version 1 carries none)
    m.dof_solref_friction[k][0] = 0.02f; m.dof_solref_friction[k][1] = 1.0f;
    const real si[5] = {0.9f, 0.95f, 0.001f, 0.5f, 2.0f};
    for (int q = 0; q < 5; ++q) m.dof_solimp_friction[k][q] = si[q];
  }
  VD(nv, m.dof_invweight0[k] = (real)D[k]);
  VD(m.nq, m.qpos0[k] = (real)D[k]);
  VI(ng, m.geom_body[k] = I[k]);
  VI(ng, m.geom_type[k] = I[k]);
  VD(ng * 3, m.geom_size[k / 3][k % 3] = (real)D[k]);
  VD(ng * 3, m.geom_pos[k / 3][k % 3] = (real)D[k]);
  if (rd.field(&D, 8) != (long)ng * 4) return bad("model blob: geom_quat");
  for (int g = 0; g < ng; ++g) {  // local rotation matrix of the geom frame (row-major), from its unit quaternion
    const double w = D[4 * g], x = D[4 * g + 1], y = D[4 * g + 2], z = D[4 * g + 3];
    const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z),
                         2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
    for (int k = 0; k < 9; ++k) m.geom_mat[g][k] = (real)R[k];
  }
  std::vector<int32_t> contype(ng), conaff(ng);
  VI(ng, contype[k] = I[k]);
  VI(ng, conaff[k] = I[k]);
  VI(ng, m.geom_condim[k] = I[k]);
  VI(ng, m.geom_priority[k] = I[k]);
  VD(ng * 3, m.geom_friction[k / 3][k % 3] = (real)D[k]);
  VD(ng, m.geom_margin[k] = (real)D[k]);
  VD(ng, m.geom_gap[k] = (real)D[k]);
  VD(ng * 2, m.geom_solref[k / 2][k % 2] = (real)D[k]);
  VD(ng * 5, m.geom_solimp[k / 5][k % 5] = (real)D[k]);
  VD(ng, m.geom_solmix[k] = (real)D[k]);
  VD(ng * 2, m.geom_invweight0[k / 2][k % 2] = (real)D[k]);
  VD(ng, m.geom_rbound[k] = (real)D[k]);
  VI(ng, m.geom_hulladr[k] = I[k]);
  VI(ng, m.geom_hullnum[k] = I[k]);
  VI(ng, (void)I[k]);  // geom_srcbody
  VI(ng, (void)I[k]);  // geom_srcparent
  if (rd.field(&D, 8) != (long)nhullvert * 3) return bad("model blob: hull_vert");
  out->hull_vert.resize(D.size());
  for (size_t k = 0; k < D.size(); ++k) out->hull_vert[k] = (float)D[k];
  m.nhull = nhullvert;
  VI(ns, m.site_body[k] = I[k]);
  VD(ns * 3, m.site_pos[k / 3][k % 3] = (real)D[k]);
  VD(ns * 4, (void)D[k]);  // site_quat (orientation of sites is not observed by the door / peg tasks)
  VI(nu, m.act_dof[k] = I[k]);
  VI(nu, m.act_qposadr[k] = I[k]);
  VD(nu, m.act_kp[k] = (real)D[k]);
  VD(nu * 2, m.act_ctrlrange[k / 2][k % 2] = (real)D[k]);
  VI(nu, m.act_ctrllimited[k] = I[k]);
  VD(nu * 2, m.act_forcerange[k / 2][k % 2] = (real)D[k]);
  VI(nu, m.act_forcelimited[k] = I[k]);
  VI(nw, m.weld_body[k] = I[k]);
  VD(nw * 3, m.weld_pos[k / 3][k % 3] = (real)D[k]);
  VD(nw * 4, m.weld_quat[k / 4][k % 4] = (real)D[k]);
  VD(nw * 7, m.weld_relpose[k / 7][k % 7] = (real)D[k]);
  VD(nw * 2, m.weld_solref[k / 2][k % 2] = (real)D[k]);
  VD(nw * 5, m.weld_solimp[k / 5][k % 5] = (real)D[k]);
  VD(nw * 2, m.weld_invweight[k / 2][k % 2] = (real)D[k]);
  VD(3, out->mocap_pos0[k] = D[k]);
  VD(4, out->mocap_quat0[k] = D[k]);
  m.neq = 0;
  if (rd.version == 2) {
    RI(m.neq);
    if (m.neq < 0 || m.neq > MAXEQ) return bad("model blob: too many joint equalities");
    const int ne = m.neq;
    VI(ne * 2, m.eq_qposadr[k / 2][k % 2] = I[k]);
    VI(ne * 2, m.eq_dofadr[k / 2][k % 2] = I[k]);
    VD(ne * 5, m.eq_polycoef[k / 5][k % 5] = (real)D[k]);
    VD(ne * 2, m.eq_solref[k / 2][k % 2] = (real)D[k]);
    VD(ne * 5, m.eq_solimp[k / 5][k % 5] = (real)D[k]);
    VD(ne, m.eq_invweight[k] = (real)D[k]);
    VD(nv * 2, m.dof_solref_friction[k / 2][k % 2] = (real)D[k]);
    VD(nv * 5, m.dof_solimp_friction[k / 5][k % 5] = (real)D[k]);
  }
#undef RI
#undef RD
#undef VI
#undef VD
  if (!rd.done()) return bad("model blob: trailing bytes");
  {  // the device impedance function only carries the power-1 and power-2 sigmoids
    auto pw_ok = [](real p) { return p == 1.0f || p == 2.0f; };
    for (int j = 0; j < m.njnt; ++j) if (!pw_ok(m.jnt_solimp[j][4])) return bad("solimp power must be 1 or 2");
    for (int g = 0; g < ng; ++g) if (!pw_ok(m.geom_solimp[g][4])) return bad("solimp power must be 1 or 2");
    for (int k = 0; k < nw; ++k) if (!pw_ok(m.weld_solimp[k][4])) return bad("solimp power must be 1 or 2");
  }
  // derived tables
  for (int b = 0; b < nb; ++b) {
    unsigned mask = 0;
    for (int c = b; c > 0; c = m.body_parent[c]) mask |= 1u << c;
    m.body_anc[b] = mask;
  }
  for (int j = 0; j < m.njnt; ++j) {
    const int da = m.jnt_dofadr[j], b = m.jnt_body[j];
    m.jnt_qpos0[j] = m.qpos0[m.jnt_qposadr[j]];
    int pd = -1;  // last dof of the nearest moving ancestor
    const int pb = m.body_parent[b];
    if (pb > 0) {
      const int pj = m.body_jnt[pb];
      pd = m.jnt_dofadr[pj] + (m.jnt_type[pj] == 0 ? 5 : 0);
    }
    if (m.jnt_type[j] == 0) {
      for (int k = 0; k < 6; ++k) { m.dof_rot[da + k] = k >= 3; m.dof_parent[da + k] = k == 0 ? pd : da + k - 1; }
    } else if (m.jnt_type[j] == 1) {
      return bad("ball joints are not built");
    } else {
      m.dof_rot[da] = m.jnt_type[j] == 3;
      m.dof_parent[da] = pd;
    }
  }
  // dof_lever (cached broad phase): how far any geom CENTRE of the dof's subtree can move per unit motion of the dof.
  // reach[b] bounds the distance from body b's origin to every geom centre of its subtree in every configuration.
  {
    std::vector<double> reach(nb, 0.0);
    auto norm3 = [](const real* v) { return std::sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]); };
    for (int g = 0; g < ng; ++g) {
      const int b = m.geom_body[g];
      if (b > 0) reach[b] = std::max(reach[b], norm3(m.geom_pos[g]));
    }
    for (int b = nb - 1; b > 0; --b) {  // parents precede children
      const int pb = m.body_parent[b], j = m.body_jnt[b];
      if (pb <= 0) continue;
      double ext = norm3(m.body_pos[b]) + reach[b];
      if (j >= 0) {
        ext += 2.0 * norm3(m.jnt_pos[j]);
        if (m.jnt_type[j] == 2) ext += std::max(std::fabs((double)m.jnt_range[j][0]), std::fabs((double)m.jnt_range[j][1]));
        if (m.jnt_type[j] == 2 && !m.jnt_limited[j]) ext += 1e3;  // unlimited slide: no bound, the cache never holds
      }
      reach[pb] = std::max(reach[pb], ext);
    }
    for (int j = 0; j < m.njnt; ++j) {
      const int da = m.jnt_dofadr[j], b = m.jnt_body[j];
      if (m.jnt_type[j] == 0) {
        for (int k = 0; k < 3; ++k) { m.dof_lever[da + k] = 1.0f; m.dof_lever[da + 3 + k] = (real)reach[b]; }
      } else if (m.jnt_type[j] == 3) {
        m.dof_lever[da] = (real)(norm3(m.jnt_pos[j]) + reach[b]);
      } else {
        m.dof_lever[da] = 1.0f;
      }
    }
  }
  // bounding boxes in the geom frames (broad-phase cull)
  for (int g = 0; g < ng; ++g) {
    real* sz = m.geom_obb_size[g];
    real* off = m.geom_obb_off[g];
    off[0] = off[1] = off[2] = 0;
    if (m.geom_type[g] == GEOM_BOX) { for (int k = 0; k < 3; ++k) sz[k] = m.geom_size[g][k]; }
    else if (m.geom_type[g] == GEOM_CYLINDER) { sz[0] = sz[1] = m.geom_size[g][0]; sz[2] = m.geom_size[g][1]; }
    else if (m.geom_type[g] == GEOM_CAPSULE) { sz[0] = sz[1] = m.geom_size[g][0]; sz[2] = m.geom_size[g][0] + m.geom_size[g][1]; }
    else if (m.geom_type[g] == GEOM_MESH && m.geom_hullnum[g] > 0) {
      real lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
      for (int v = 0; v < m.geom_hullnum[g]; ++v)
        for (int k = 0; k < 3; ++k) {
          const real x = out->hull_vert[3 * (m.geom_hulladr[g] + v) + k];
          lo[k] = x < lo[k] ? x : lo[k];
          hi[k] = x > hi[k] ? x : hi[k];
        }
      for (int k = 0; k < 3; ++k) { sz[k] = 0.5f * (hi[k] - lo[k]); off[k] = 0.5f * (hi[k] + lo[k]); }
    } else { sz[0] = sz[1] = sz[2] = m.geom_rbound[g]; }
  }
  // geoms on moving bodies get a pose slot; static geoms keep their constant world pose in geom_pos / geom_mat
  m.nmgeom = 0;
  for (int g = 0; g < ng; ++g) {
    m.geom_slot[g] = -1;
    if (m.geom_body[g] != 0) {
      if (m.nmgeom >= MAXMG) return bad("too many geoms on moving bodies");
      m.geom_slot[g] = m.nmgeom;
      m.mgeom[m.nmgeom++] = g;
    }
  }
  // candidate collision pairs: contype/conaffinity, same fused body, fused parent-child unless one side is the world
  // (MuJoCo engine_collision_driver.c: filterBodyPair on weld ids, filterBitmask)
  int np = 0;
  for (int g1 = 0; g1 < ng; ++g1)
    for (int g2 = g1 + 1; g2 < ng; ++g2) {
      if (!((contype[g1] & conaff[g2]) || (contype[g2] & conaff[g1]))) continue;
      const int b1 = m.geom_body[g1], b2 = m.geom_body[g2];
      if (b1 == b2) continue;
      if (b1 != 0 && b2 != 0 && (m.body_parent[b1] == b2 || m.body_parent[b2] == b1)) continue;
      if (np >= MAXPAIR) return bad("too many candidate collision pairs");
      const int cd = m.geom_condim[g1] > m.geom_condim[g2] ? m.geom_condim[g1] : m.geom_condim[g2];
      if (cone_elliptic ? (cd != 3 && cd != 4) : (cd != 1 && cd != 3 && cd != 4 && cd != 6))
        return bad("contact dimension not built (elliptic: 3, 4; pyramidal: 1, 3, 4, 6)");
      if (never_touch(m, g1, g2) || never_touch(m, g2, g1)) continue;
      const bool swap = m.geom_type[g1] > m.geom_type[g2];  // MuJoCo orders a pair by geom type
      m.pair_g1[np] = swap ? g2 : g1;
      m.pair_g2[np] = swap ? g1 : g2;
      ++np;
    }
  m.npair = np;
  m.frame_skip = task.frame_skip;
  if (task.max_newton > 0 && task.max_newton < m.iterations) m.iterations = task.max_newton;
  m.obs_hand_site = task.hand_site;
  m.obs_ree_site = task.ree_site;
  m.obs_lee_site = task.lee_site;
  m.obs_obj_geom = task.obj_geom;
  m.obs_obj_site = task.obj_site;
  for (int k = 0; k < 3; ++k) { m.mocap_low[k] = task.mocap_low[k]; m.mocap_high[k] = task.mocap_high[k]; }
  m.action_scale = task.action_scale;
  m.success_radius = task.success_radius;
  for (int k = 0; k < 3; ++k) { m.obj_init_pos[k] = task.obj_init_pos[k]; m.hand_init_pos[k] = task.hand_init_pos[k]; }
  m.grasp_site = task.grasp_site; m.lpad_site = task.lpad_site; m.rpad_site = task.rpad_site;
  for (int k = 0; k < 4; ++k) m.corner_site[k] = task.corner_site[k];
  {
    const int ids[7] = {task.grasp_site, task.lpad_site, task.rpad_site, task.corner_site[0], task.corner_site[1], task.corner_site[2], task.corner_site[3]};
    for (int k = 0; k < 7; ++k) if (ids[k] >= ns) return bad("task spec: dense-reward site index out of range");
  }
  if (task.hand_site < 0 || task.hand_site >= ns || task.ree_site < 0 || task.ree_site >= ns || task.lee_site < 0 ||
      task.lee_site >= ns || (task.obj_geom < 0 && (task.obj_site < 0 || task.obj_site >= ns)) || task.obj_geom >= ng)
    return bad("task spec: observation site / geom index out of range");
  return true;
}

}  // namespace mj
}  // namespace earl
