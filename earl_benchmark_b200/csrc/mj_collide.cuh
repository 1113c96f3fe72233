// mj_collide.cuh -- collision detection and contact constraint rows of the warp-per-env engine.
//
// Replaces MuJoCo 2.1 mj_collision + mj_instantiateContact for the geom types of the Sawyer scenes (plane, cylinder,
// box, convex mesh):  candidate pairs are filtered once on the host (contype / conaffinity, same welded body, welded
// parent-child; mj_model_host.hpp); per substep every candidate goes through MuJoCo's bounding-sphere test plus an
// exact sphere-vs-box cull that can only remove pairs with no contact, then box-box pairs run a 15-axis separating
// axis test with face clipping / edge-edge closest points, and all other pairs run Minkowski Portal Refinement on
// support functions (the algorithm of libccd's ccdMPRPenetration, which MuJoCo's mjc_Convex calls), geoms inflated
// by margin / 2.  Contact parameters: condim = max, friction = max, solref / solimp mixed by solmix.
//
// Execution: geom poses and the broad phase are lane-parallel; the (few) surviving pairs are processed one after the
// other with every lane executing the same narrow-phase code, except the support function of mesh geoms, which is a
// lane-parallel argmax over the hull vertices.
#pragma once

#include "mj_engine.cuh"

namespace earl {
namespace mj {

constexpr int GEOM_PLANE = 0, GEOM_CAPSULE = 3, GEOM_CYLINDER = 5, GEOM_BOX = 6, GEOM_MESH = 7;
// Portal refinement runs in fp64 (`mreal`) on fp32 poses: its termination tests (libccd: |x| < eps, portal tolerance 1e-6)
// cannot be resolved in fp32, where the refinement stops at a different portal and contact normals of thin-box-vs-
// cylinder pairs came out up to 30 degrees away from the fp64 result; rounding the INPUTS to fp32 changes nothing
// measurable.  B200 issues fp64 at half the fp32 rate and MPR is a minor share of the step.
// Its branches are also DISCONTINUOUS in the inputs: at a near-degenerate portal (thin pad edge on a cylinder) a 1e-7
// change of a pose decides which portal the refinement continues with, and the two contact normals can be degrees
// apart (measured: 3 degrees, 1 of 96 probe states; the fp64 checker perturbed by 1e-7 flips the same way).  That is a
// property of the reference's algorithm, not of the precision: forcing the device to round like the host build (no
// FMA contraction) did not change a single result.
typedef double mreal;
constexpr mreal MPR_TOL = 1e-6;
constexpr int MPR_ITER = 50;
constexpr mreal CCD_EPS = 2.220446049250313e-16;

struct RawCon { real pos[3], normal[3], dist; };
struct Supp { mreal v[3], v1[3], v2[3]; };
// Narrow-phase scratch.  It lives in SHARED memory (aliased onto Work::H, which is idle during collision): every lane of
// the warp runs the same narrow-phase code and writes identical values to identical addresses, so no lane ever reads a
// value it did not also write itself; keeping these dynamically indexed arrays out of local memory keeps the kernel
// off the L1/L2 path.
// a convex geom as the support function sees it: pose / size pointers resolved once per pair
struct CObj {
  int type, nvert;
  real margin;
  const real* pos;   // world position of the geom frame
  const real* R;     // world rotation (row-major)
  const real* size;
  const real* hv;    // hull vertices (mesh geoms)
};
struct NarrowScratch {
  CObj o1, o2;
  int nsup;  // support-function calls of the current portal refinement (scheduling cost estimate)
  RawCon rc[8];
  struct BoxScratch { real A[3][3], B[3][3], poly[16][3], tmp[16][3]; };
  struct PortalScratch { Supp P[4], v4; };
  union {
    BoxScratch bb;     // box-box
    PortalScratch pr;  // portal refinement
  };
};
static_assert(sizeof(NarrowScratch) <= sizeof(real) * MAXV * LDM, "narrow-phase scratch must fit in Work::H");

template <class T>
MJ_HD void sub3(T* r, const T* a, const T* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
template <class T>
MJ_HD T normalize3(T* v) {
  T n = msqrt(dot3(v, v));
  if (n < (T)1e-15) { v[0] = 1; v[1] = v[2] = 0; return 0; }
  T s = (T)1 / n;
  v[0] *= s; v[1] *= s; v[2] *= s;
  return n;
}
MJ_HD void col3(real* c, const real* R, int k) { c[0] = R[k]; c[1] = R[3 + k]; c[2] = R[6 + k]; }

// mju_makeFrame: complete the contact frame from its normal
MJ_FN void make_frame(real* frame) {
  real* x = frame; real* y = frame + 3; real* z = frame + 6;
  y[0] = y[1] = y[2] = 0;
  if (x[1] < 0.5f && x[1] > -0.5f) y[1] = 1; else y[2] = 1;
  const real t = dot3(x, y);
  for (int k = 0; k < 3; ++k) y[k] -= t * x[k];
  normalize3(y);
  cross3(z, x, y);
}

// ------------------------------------------------------------------------------------------------ geom poses
// Static geoms (body 0) keep their constant world pose in the model; geoms on moving bodies get theirs recomputed.
MJ_HD const real* gpos(const Model& m, const Work& w, int g) { return m.geom_slot[g] < 0 ? m.geom_pos[g] : w.mg_xpos[m.geom_slot[g]]; }
MJ_HD const real* gmat(const Model& m, const Work& w, int g) { return m.geom_slot[g] < 0 ? m.geom_mat[g] : w.mg_xmat[m.geom_slot[g]]; }

template <int NL>
MJ_FN void geom_poses(const Model& m, Work& w, int lane) {
  for (int s = lane; s < m.nmgeom; s += NL) {
    const int g = m.mgeom[s], b = m.geom_body[g];
    real t[3];
    mulmatvec3(t, w.xmat[b], m.geom_pos[g]);
    for (int k = 0; k < 3; ++k) w.mg_xpos[s][k] = w.xpos[b][k] + t[k];
    const real* A = w.xmat[b];
    const real* B = m.geom_mat[g];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) w.mg_xmat[s][3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  }
  wsync<NL>();
}

// ------------------------------------------------------------------------------------------------ box-box
MJ_FN int clip_poly(real (*poly)[3], int n, const real* pn, real pd, real (*out)[3]) {
  int mcount = 0;  // keep the half space pn.x <= pd (Sutherland-Hodgman)
  for (int i = 0; i < n; ++i) {
    const real* a = poly[i];
    const real* b = poly[(i + 1) % n];
    const real da = dot3(pn, a) - pd, db = dot3(pn, b) - pd;
    if (da <= 0) { out[mcount][0] = a[0]; out[mcount][1] = a[1]; out[mcount][2] = a[2]; ++mcount; }
    if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
      const real t = da / (da - db);
      for (int k = 0; k < 3; ++k) out[mcount][k] = a[k] + t * (b[k] - a[k]);
      ++mcount;
    }
    if (mcount >= 15) break;
  }
  return mcount;
}

MJ_FN int box_box(const real* p1, const real* R1, const real* s1, const real* p2, const real* R2, const real* s2, real margin,
                  RawCon* out, NarrowScratch* S) {
  real (*A)[3] = S->bb.A;
  real (*B)[3] = S->bb.B;
  real (*poly)[3] = S->bb.poly;
  real (*tmp)[3] = S->bb.tmp;
  real pp[3], pA[3], pB[3], Q[3][3];
  for (int k = 0; k < 3; ++k) { col3(A[k], R1, k); col3(B[k], R2, k); }
  sub3(pp, p2, p1);
  for (int i = 0; i < 3; ++i) {
    pA[i] = dot3(pp, A[i]);
    pB[i] = dot3(pp, B[i]);
    for (int j = 0; j < 3; ++j) Q[i][j] = fabsf(dot3(A[i], B[j]));
  }
  real best = -1e30f;
  int code = -1, flip = 0;
  real nrm[3] = {0, 0, 0};
  for (int i = 0; i < 3; ++i) {
    const real s = fabsf(pA[i]) - (s1[i] + s2[0] * Q[i][0] + s2[1] * Q[i][1] + s2[2] * Q[i][2]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = i; flip = pA[i] < 0; nrm[0] = A[i][0]; nrm[1] = A[i][1]; nrm[2] = A[i][2]; }
  }
  for (int j = 0; j < 3; ++j) {
    const real s = fabsf(pB[j]) - (s2[j] + s1[0] * Q[0][j] + s1[1] * Q[1][j] + s1[2] * Q[2][j]);
    if (s > margin) return 0;
    if (s > best + 1e-9f) { best = s; code = 3 + j; flip = pB[j] < 0; nrm[0] = B[j][0]; nrm[1] = B[j][1]; nrm[2] = B[j][2]; }
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      real ax[3];
      cross3(ax, A[i], B[j]);
      const real l = sqrtf(dot3(ax, ax));
      if (l < 1e-6f) continue;
      for (int k = 0; k < 3; ++k) ax[k] /= l;
      const real d = dot3(pp, ax);
      real ra = 0, rb = 0;
      for (int k = 0; k < 3; ++k) { ra += s1[k] * fabsf(dot3(A[k], ax)); rb += s2[k] * fabsf(dot3(B[k], ax)); }
      const real s = fabsf(d) - (ra + rb);
      if (s > margin) return 0;
      // a face axis is preferred unless the edge axis is clearly better
      if (s > best + 1e-9f && s * 1.05f > best + (best < 0 ? 0.0f : 1e-9f) && s - 0.05f * fabsf(s) > best) {
        best = s; code = 6 + 3 * i + j; flip = d < 0; nrm[0] = ax[0]; nrm[1] = ax[1]; nrm[2] = ax[2];
      }
    }
  if (flip) for (int k = 0; k < 3; ++k) nrm[k] = -nrm[k];  // nrm points from box 1 to box 2
  if (code >= 6) {  // edge-edge: closest points of the two supporting edges
    const int i = (code - 6) / 3, j = (code - 6) % 3;
    real pa[3], pb[3];
    for (int k = 0; k < 3; ++k) { pa[k] = p1[k]; pb[k] = p2[k]; }
    for (int a = 0; a < 3; ++a) {
      if (a == i) continue;
      const real sg = dot3(nrm, A[a]) > 0 ? 1.0f : -1.0f;
      for (int k = 0; k < 3; ++k) pa[k] += sg * s1[a] * A[a][k];
    }
    for (int b = 0; b < 3; ++b) {
      if (b == j) continue;
      const real sg = dot3(nrm, B[b]) > 0 ? -1.0f : 1.0f;
      for (int k = 0; k < 3; ++k) pb[k] += sg * s2[b] * B[b][k];
    }
    const real* ua = A[i];
    const real* ub = B[j];
    real wv[3];
    sub3(wv, pb, pa);
    const real uaub = dot3(ua, ub), q1 = dot3(ua, wv), q2 = -dot3(ub, wv), den = 1 - uaub * uaub;
    const real t = den < 1e-12f ? 0.0f : (q1 + uaub * q2) / den, u = den < 1e-12f ? 0.0f : (uaub * q1 + q2) / den;
    for (int k = 0; k < 3; ++k) { out[0].pos[k] = 0.5f * ((pa[k] + t * ua[k]) + (pb[k] + u * ub[k])); out[0].normal[k] = nrm[k]; }
    out[0].dist = best;
    return 1;
  }
  // face contact: the reference box owns the axis, the incident face of the other box is clipped against it
  const real *pr, *pi, *sr, *si;
  real (*Rr)[3];
  real (*Ri)[3];
  real n[3];
  int ai;
  if (code < 3) { pr = p1; pi = p2; sr = s1; si = s2; Rr = A; Ri = B; ai = code; n[0] = nrm[0]; n[1] = nrm[1]; n[2] = nrm[2]; }
  else { pr = p2; pi = p1; sr = s2; si = s1; Rr = B; Ri = A; ai = code - 3; n[0] = -nrm[0]; n[1] = -nrm[1]; n[2] = -nrm[2]; }
  int aj = 0;
  real bestd = -1;
  for (int j = 0; j < 3; ++j) { const real d = fabsf(dot3(n, Ri[j])); if (d > bestd) { bestd = d; aj = j; } }
  const real sg = dot3(n, Ri[aj]) > 0 ? -1.0f : 1.0f;
  real c[3];
  for (int k = 0; k < 3; ++k) c[k] = pi[k] + sg * si[aj] * Ri[aj][k];
  const int k1 = (aj + 1) % 3, k2 = (aj + 2) % 3;
  for (int v = 0; v < 4; ++v) {
    const real sx = (v == 0 || v == 3) ? 1.0f : -1.0f, sy = v < 2 ? 1.0f : -1.0f;
    for (int k = 0; k < 3; ++k) poly[v][k] = c[k] + sx * si[k1] * Ri[k1][k] + sy * si[k2] * Ri[k2][k];
  }
  int np = 4;
  for (int e = 0; e < 2 && np > 0; ++e) {
    const int a = (ai + 1 + e) % 3;
    real pn[3] = {Rr[a][0], Rr[a][1], Rr[a][2]};
    const real base = dot3(pn, pr);
    np = clip_poly(poly, np, pn, base + sr[a], tmp);
    for (int k = 0; k < 3; ++k) pn[k] = -pn[k];
    np = clip_poly(tmp, np, pn, -base + sr[a], poly);
  }
  int nc = 0;
  for (int v = 0; v < np && nc < 8; ++v) {
    real rel[3];
    sub3(rel, poly[v], pr);
    const real depth = sr[ai] - dot3(rel, n);  // > 0: the vertex is below the reference face
    if (-depth >= margin) continue;
    int dup = 0;
    for (int q = 0; q < nc; ++q) {
      real dd[3];
      for (int k = 0; k < 3; ++k) dd[k] = poly[v][k] + 0.5f * depth * n[k] - out[q].pos[k];
      if (dot3(dd, dd) < 1e-12f) dup = 1;
    }
    if (dup) continue;
    for (int k = 0; k < 3; ++k) { out[nc].pos[k] = poly[v][k] + 0.5f * depth * n[k]; out[nc].normal[k] = nrm[k]; }
    // MuJoCo 2.1.0 reports HALF of this vertex-below-face depth as the contact distance of a box-box face contact
    // (established on the reference's own data: the peg dropped onto the table at the start of every shipped peg episode
    // lands, bounces and settles within 16 NANOMETRES of the recording over 26 env steps with dist = -depth / 2, and
    // 0.87 mm off with dist = -depth; see the pin tests of the fp64 checker)
    out[nc].dist = -0.5f * depth;
    ++nc;
  }
  return nc;
}

// ------------------------------------------------------------------------------------------------ support functions
MJ_HD void make_cobj(CObj& o, const Model& m, const real* hull, const Work& w, int g, real margin) {
  o.type = m.geom_type[g];
  o.nvert = m.geom_hullnum[g];
  o.margin = margin;
  o.pos = gpos(m, w, g);
  o.R = gmat(m, w, g);
  o.size = m.geom_size[g];
  o.hv = hull + 3 * m.geom_hulladr[g];
}

template <int NL>
MJ_HD void support_geom(const CObj& o, const mreal* dir, mreal* res, int lane) {
  const real* R = o.R;
  const real* sz = o.size;
  const mreal dl[3] = {R[0] * dir[0] + R[3] * dir[1] + R[6] * dir[2], R[1] * dir[0] + R[4] * dir[1] + R[7] * dir[2],
                       R[2] * dir[0] + R[5] * dir[1] + R[8] * dir[2]};
  mreal loc[3] = {0, 0, 0};
  if (o.type == GEOM_BOX) {
    for (int k = 0; k < 3; ++k) loc[k] = dl[k] >= 0 ? (mreal)sz[k] : -(mreal)sz[k];
  } else if (o.type == GEOM_CYLINDER) {
    const mreal t = sqrt(dl[0] * dl[0] + dl[1] * dl[1]);
    if (t > 1e-15) { loc[0] = dl[0] / t * sz[0]; loc[1] = dl[1] / t * sz[0]; }
    loc[2] = dl[2] >= 0 ? (mreal)sz[1] : -(mreal)sz[1];
  } else if (KITCHEN_ROWS && o.type == GEOM_CAPSULE) {  // segment along local z (half length size[1]) inflated by the radius size[0]
    const mreal t = sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
    if (t > 1e-15) for (int k = 0; k < 3; ++k) loc[k] = dl[k] / t * sz[0];
    loc[2] += dl[2] >= 0 ? (mreal)sz[1] : -(mreal)sz[1];
  } else if (o.type == GEOM_MESH) {
    // lane-parallel argmax over the hull vertices; the first vertex reaching the maximum wins
    const real* hv = o.hv;
    const int nvert = o.nvert;
    mreal bd = -1e300;
    int bi = nvert;
    for (int v = lane; v < nvert; v += NL) {
      const mreal dd = hv[3 * v] * dl[0] + hv[3 * v + 1] * dl[1] + hv[3 * v + 2] * dl[2];
      if (dd > bd) { bd = dd; bi = v; }
    }
    const mreal top = wmaxd<NL>(bd);
    const int cand = (bd == top) ? bi : nvert;
    const int best = -(int)wmax<NL>(-(real)cand);  // smallest index among the lanes holding the maximum
    loc[0] = hv[3 * best]; loc[1] = hv[3 * best + 1]; loc[2] = hv[3 * best + 2];
  }
  for (int k = 0; k < 3; ++k)
    res[k] = (mreal)o.pos[k] + R[3 * k] * loc[0] + R[3 * k + 1] * loc[1] + R[3 * k + 2] * loc[2] + dir[k] * (mreal)o.margin;
}

#if defined(MJ_DEBUG) && !defined(__CUDA_ARCH__)
static long g_support_calls = 0, g_mpr_calls = 0, g_mpr_hits = 0;
#endif
template <int NL>
MJ_FN void mpr_support(const CObj& o1, const CObj& o2, const mreal* dir, Supp* s, int lane) {
#if defined(MJ_DEBUG) && !defined(__CUDA_ARCH__)
  ++g_support_calls;
#endif
  const mreal nd[3] = {-dir[0], -dir[1], -dir[2]};
  support_geom<NL>(o1, dir, s->v1, lane);
  support_geom<NL>(o2, nd, s->v2, lane);
  sub3(s->v, s->v1, s->v2);
}
MJ_HD int is_zero(mreal x) { return fabs(x) < CCD_EPS; }
MJ_HD int ccd_eq(mreal a, mreal b) {
  const mreal ab = fabs(a - b);
  if (ab < CCD_EPS) return 1;
  const mreal fa = fabs(a), fb = fabs(b);
  return ab < CCD_EPS * (fb > fa ? fb : fa);
}
MJ_HD void portal_dir(const Supp* P, mreal* dir) {
  mreal a[3], b[3];
  sub3(a, P[2].v, P[1].v);
  sub3(b, P[3].v, P[1].v);
  cross3(dir, a, b);
  normalize3(dir);
}
MJ_HD int reach_tolerance(const Supp* P, const Supp* v4, const mreal* dir) {
  const mreal dv4 = dot3(v4->v, dir), d1 = dv4 - dot3(P[1].v, dir), d2 = dv4 - dot3(P[2].v, dir), d3 = dv4 - dot3(P[3].v, dir);
  mreal d = d1 < d2 ? d1 : d2;
  d = d < d3 ? d : d3;
  return ccd_eq(d, MPR_TOL) || d < MPR_TOL;
}
MJ_HD void expand_portal(Supp* P, const Supp* v4) {
  mreal v4v0[3];
  cross3(v4v0, v4->v, P[0].v);
  if (dot3(P[1].v, v4v0) > 0) {
    if (dot3(P[2].v, v4v0) > 0) P[1] = *v4; else P[3] = *v4;
  } else {
    if (dot3(P[3].v, v4v0) > 0) P[2] = *v4; else P[1] = *v4;
  }
}
MJ_FN mreal seg_dist2(const mreal* Pt, const mreal* x0, const mreal* b, mreal* wit) {
  mreal d[3], a[3];
  sub3(d, b, x0);
  sub3(a, x0, Pt);
  const mreal t = -dot3(a, d) / dot3(d, d);
  if (t < 0 || is_zero(t)) { wit[0] = x0[0]; wit[1] = x0[1]; wit[2] = x0[2]; }
  else if (t > 1 || ccd_eq(t, 1)) { wit[0] = b[0]; wit[1] = b[1]; wit[2] = b[2]; }
  else for (int k = 0; k < 3; ++k) wit[k] = x0[k] + t * d[k];
  mreal r[3];
  sub3(r, wit, Pt);
  return dot3(r, r);
}
MJ_FN mreal tri_dist2(const mreal* Pt, const mreal* x0, const mreal* B, const mreal* C, mreal* wit) {
  mreal d1[3], d2[3], a[3];
  sub3(d1, B, x0);
  sub3(d2, C, x0);
  sub3(a, x0, Pt);
  const mreal v = dot3(d1, d1), ww = dot3(d2, d2), p = dot3(a, d1), q = dot3(a, d2), r = dot3(d1, d2);
  const mreal s = (q * r - ww * p) / (ww * v - r * r), t = (-s * r - q) / ww;
  if ((is_zero(s) || s > 0) && (ccd_eq(s, 1) || s < 1) && (is_zero(t) || t > 0) && (ccd_eq(t, 1) || t < 1) &&
      (ccd_eq(t + s, 1) || t + s < 1)) {
    for (int k = 0; k < 3; ++k) wit[k] = x0[k] + s * d1[k] + t * d2[k];
    mreal rr[3];
    sub3(rr, wit, Pt);
    return dot3(rr, rr);
  }
  mreal w2[3];
  mreal dist = seg_dist2(Pt, x0, B, wit), d2v = seg_dist2(Pt, x0, C, w2);
  if (d2v < dist) { dist = d2v; wit[0] = w2[0]; wit[1] = w2[1]; wit[2] = w2[2]; }
  d2v = seg_dist2(Pt, B, C, w2);
  if (d2v < dist) { dist = d2v; wit[0] = w2[0]; wit[1] = w2[1]; wit[2] = w2[2]; }
  return dist;
}

// returns 1 with (depth, dir, pos) when the (inflated) geoms penetrate; all lanes take the same path
template <int NL>
MJ_FN int mpr_penetration(const CObj& o1, const CObj& o2, real* depth_out, real* dir_out, real* pos_out, NarrowScratch* S, int lane) {
  Supp* P = S->pr.P;
  Supp& v4 = S->pr.v4;
  if (lane == 0) S->nsup = 0;
  const mreal origin[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) { P[0].v1[k] = o1.pos[k]; P[0].v2[k] = o2.pos[k]; }
  sub3(P[0].v, P[0].v1, P[0].v2);
  if (is_zero(P[0].v[0]) && is_zero(P[0].v[1]) && is_zero(P[0].v[2])) P[0].v[0] = 0.00001;
  mreal d[3] = {-P[0].v[0], -P[0].v[1], -P[0].v[2]}, va[3], vb[3], dt, depth, dir[3], pos[3];
  normalize3(d);
  mpr_support<NL>(o1, o2, d, &P[1], lane);
  if (lane == 0) S->nsup += 1;
  dt = dot3(P[1].v, d);
  if (is_zero(dt) || dt < 0) return 0;
  cross3(d, P[0].v, P[1].v);
  if (is_zero(dot3(d, d))) {
    if (is_zero(P[1].v[0]) && is_zero(P[1].v[1]) && is_zero(P[1].v[2])) return 0;
    depth = sqrt(dot3(P[1].v, P[1].v));
    dir[0] = P[1].v[0]; dir[1] = P[1].v[1]; dir[2] = P[1].v[2];
    normalize3(dir);
    for (int k = 0; k < 3; ++k) { pos_out[k] = (real)(0.5 * (P[1].v1[k] + P[1].v2[k])); dir_out[k] = (real)dir[k]; }
    *depth_out = (real)depth;
    return 1;
  }
  normalize3(d);
  mpr_support<NL>(o1, o2, d, &P[2], lane);
  if (lane == 0) S->nsup += 1;
  dt = dot3(P[2].v, d);
  if (is_zero(dt) || dt < 0) return 0;
  sub3(va, P[1].v, P[0].v);
  sub3(vb, P[2].v, P[0].v);
  cross3(d, va, vb);
  normalize3(d);
  if (dot3(d, P[0].v) > 0) {
    v4 = P[1]; P[1] = P[2]; P[2] = v4;
    for (int k = 0; k < 3; ++k) d[k] = -d[k];
  }
  for (int guard = 0; guard < 100; ++guard) {
    mpr_support<NL>(o1, o2, d, &P[3], lane);
    if (lane == 0) S->nsup += 1;
    dt = dot3(P[3].v, d);
    if (is_zero(dt) || dt < 0) return 0;
    int cont = 0;
    cross3(va, P[1].v, P[3].v);
    dt = dot3(va, P[0].v);
    if (dt < 0 && !is_zero(dt)) { P[2] = P[3]; cont = 1; }
    if (!cont) {
      cross3(va, P[3].v, P[2].v);
      dt = dot3(va, P[0].v);
      if (dt < 0 && !is_zero(dt)) { P[1] = P[3]; cont = 1; }
    }
    if (!cont) break;
    sub3(va, P[1].v, P[0].v);
    sub3(vb, P[2].v, P[0].v);
    cross3(d, va, vb);
    normalize3(d);
  }
  for (int guard = 0;; ++guard) {  // refinePortal
    portal_dir(P, d);
    dt = dot3(d, P[1].v);
    if (is_zero(dt) || dt > 0) break;
    mpr_support<NL>(o1, o2, d, &v4, lane);
    if (lane == 0) S->nsup += 1;
    dt = dot3(v4.v, d);
    if (!(is_zero(dt) || dt > 0) || reach_tolerance(P, &v4, d) || guard > 200) return 0;
    expand_portal(P, &v4);
  }
  for (int it = 0;; ++it) {  // findPenetr
    portal_dir(P, d);
    mpr_support<NL>(o1, o2, d, &v4, lane);
    if (lane == 0) S->nsup += 1;
    if (reach_tolerance(P, &v4, d) || it > MPR_ITER) {
      mreal wit[3];
      depth = sqrt(tri_dist2(origin, P[1].v, P[2].v, P[3].v, wit));
      if (is_zero(depth)) return 0;
      dir[0] = wit[0]; dir[1] = wit[1]; dir[2] = wit[2];
      normalize3(dir);
      mreal b[4], t[3];
      portal_dir(P, d);
      cross3(t, P[1].v, P[2].v); b[0] = dot3(t, P[3].v);
      cross3(t, P[3].v, P[2].v); b[1] = dot3(t, P[0].v);
      cross3(t, P[0].v, P[1].v); b[2] = dot3(t, P[3].v);
      cross3(t, P[2].v, P[1].v); b[3] = dot3(t, P[0].v);
      mreal sum = b[0] + b[1] + b[2] + b[3];
      if (is_zero(sum) || sum < 0) {
        b[0] = 0;
        cross3(t, P[2].v, P[3].v); b[1] = dot3(t, d);
        cross3(t, P[3].v, P[1].v); b[2] = dot3(t, d);
        cross3(t, P[1].v, P[2].v); b[3] = dot3(t, d);
        sum = b[1] + b[2] + b[3];
      }
      const mreal inv = 1.0 / sum;
      for (int k = 0; k < 3; ++k) {
        mreal q1 = 0, q2 = 0;
        for (int v = 0; v < 4; ++v) { q1 += b[v] * P[v].v1[k]; q2 += b[v] * P[v].v2[k]; }
        pos[k] = 0.5 * (q1 + q2) * inv;
      }
      for (int k = 0; k < 3; ++k) { pos_out[k] = (real)pos[k]; dir_out[k] = (real)dir[k]; }
      *depth_out = (real)depth;
      return 1;
    }
    expand_portal(P, &v4);
  }
}

// plane (z axis of the plane geom) vs box corners / lowest support point of any other convex geom
template <int NL>
MJ_FN int plane_convex(const Model& m, const real* hull, Work& w, int gp, int g, real margin, RawCon* out, int lane) {
  const real* Rp = gmat(m, w, gp);
  const real n[3] = {Rp[2], Rp[5], Rp[8]}, nd[3] = {-Rp[2], -Rp[5], -Rp[8]};
  real pt[3], rel[3];
  int nc = 0;
  if (m.geom_type[g] == GEOM_BOX) {
    const real* R = gmat(m, w, g);
    const real* sz = m.geom_size[g];
    for (int v = 0; v < 8 && nc < 4; ++v) {
      const real loc[3] = {(v & 1 ? 1 : -1) * sz[0], (v & 2 ? 1 : -1) * sz[1], (v & 4 ? 1 : -1) * sz[2]};
      for (int k = 0; k < 3; ++k) pt[k] = gpos(m, w, g)[k] + R[3 * k] * loc[0] + R[3 * k + 1] * loc[1] + R[3 * k + 2] * loc[2];
      sub3(rel, pt, gpos(m, w, gp));
      const real dist = dot3(rel, n);
      if (dist >= margin) continue;
      for (int k = 0; k < 3; ++k) { out[nc].pos[k] = pt[k] - 0.5f * dist * n[k]; out[nc].normal[k] = n[k]; }
      out[nc].dist = dist;
      ++nc;
    }
    return nc;
  }
  CObj& o = reinterpret_cast<NarrowScratch*>(&w.H[0][0])->o1;
  make_cobj(o, m, hull, w, g, 0.0f);
  const mreal ndd[3] = {nd[0], nd[1], nd[2]};
  mreal ptd[3];
  support_geom<NL>(o, ndd, ptd, lane);
  for (int k = 0; k < 3; ++k) pt[k] = (real)ptd[k];
  sub3(rel, pt, gpos(m, w, gp));
  const real dist = dot3(rel, n);
  if (dist >= margin) return 0;
  for (int k = 0; k < 3; ++k) { out[0].pos[k] = pt[k] - 0.5f * dist * n[k]; out[0].normal[k] = n[k]; }
  out[0].dist = dist;
  return 1;
}

// squared distance from point c to the box (centre p, axes R columns, half sizes s)
// mjc_PlaneCapsule: one contact per end sphere
MJ_FN int plane_capsule(const real* pp, const real* Rp, const real* pc, const real* Rc, const real* sz, real margin, RawCon* out) {
  const real n[3] = {Rp[2], Rp[5], Rp[8]}, ax[3] = {Rc[2], Rc[5], Rc[8]};
  int nc = 0;
  for (int e = -1; e <= 1; e += 2) {
    real c[3], rel[3];
    for (int k = 0; k < 3; ++k) c[k] = pc[k] + e * sz[1] * ax[k];
    sub3(rel, c, pp);
    const real dist = dot3(rel, n) - sz[0];
    if (dist >= margin) continue;
    for (int k = 0; k < 3; ++k) { out[nc].pos[k] = c[k] - n[k] * (sz[0] + 0.5f * dist); out[nc].normal[k] = n[k]; }
    out[nc].dist = dist;
    ++nc;
  }
  return nc;
}
MJ_HD int sphere_sphere(const real* c1, real r1, const real* c2, real r2, real margin, RawCon* out) {
  real n[3];
  sub3(n, c2, c1);
  const real len = msqrt(dot3(n, n)), dist = len - r1 - r2;
  if (dist >= margin) return 0;
  if (len < MINVAL) { n[0] = 1; n[1] = n[2] = 0; } else { for (int k = 0; k < 3; ++k) n[k] /= len; }
  for (int k = 0; k < 3; ++k) { out->pos[k] = c1[k] + n[k] * (r1 + 0.5f * dist); out->normal[k] = n[k]; }
  out->dist = dist;
  return 1;
}
// mjc_CapsuleCapsule: closest points of the two axis segments; parallel axes give the two ends of the overlap
MJ_FN int capsule_capsule(const real* p1, const real* R1, const real* s1, const real* p2, const real* R2, const real* s2, real margin,
                          RawCon* out) {
  const real a1[3] = {R1[2] * s1[1], R1[5] * s1[1], R1[8] * s1[1]}, a2[3] = {R2[2] * s2[1], R2[5] * s2[1], R2[8] * s2[1]};
  real dif[3], c1[3], c2[3];
  sub3(dif, p1, p2);
  const real ma = dot3(a1, a1), mb = -dot3(a1, a2), mc = dot3(a2, a2), u = -dot3(a1, dif), v = dot3(a2, dif);
  const real det = ma * mc - mb * mb;
  if (mabs(det) >= 1e-9f * ma * mc) {  // general configuration (fp32: relative test for parallelism)
    real x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > 1) { x1 = 1; x2 = (v - mb) / mc; } else if (x1 < -1) { x1 = -1; x2 = (v + mb) / mc; }
    if (x2 > 1) { x2 = 1; x1 = clampr((u - mb) / ma, -1.0f, 1.0f); }
    else if (x2 < -1) { x2 = -1; x1 = clampr((u + mb) / ma, -1.0f, 1.0f); }
    for (int k = 0; k < 3; ++k) { c1[k] = p1[k] + a1[k] * x1; c2[k] = p2[k] + a2[k] * x2; }
    return sphere_sphere(c1, s1[0], c2, s2[0], margin, out);
  }
  int nc = 0;
  for (int e = -1; e <= 1; e += 2) {
    real x1 = (real)e;
    const real x2 = clampr((v - mb * x1) / mc, -1.0f, 1.0f);
    x1 = clampr((u - mb * x2) / ma, -1.0f, 1.0f);
    for (int k = 0; k < 3; ++k) { c1[k] = p1[k] + a1[k] * x1; c2[k] = p2[k] + a2[k] * x2; }
    nc += sphere_sphere(c1, s1[0], c2, s2[0], margin, out + nc);
  }
  if (nc == 2) {
    real dd[3];
    sub3(dd, out[0].pos, out[1].pos);
    if (dot3(dd, dd) < 1e-12f) nc = 1;
  }
  return nc;
}

MJ_HD real point_box_dist2(const real* c, const real* p, const real* R, const real* s) {
  real rel[3], d2 = 0;
  sub3(rel, c, p);
  for (int k = 0; k < 3; ++k) {
    const real x = fabsf(rel[0] * R[k] + rel[1] * R[3 + k] + rel[2] * R[6 + k]) - s[k];
    if (x > 0) d2 += x * x;
  }
  return d2;
}

// Conservative cull: true when the bounding boxes of the two geoms (centres ca / cb, axes = columns of Ra / Rb, half
// sizes sa / sb) are separated by more than `margin` along one of the 15 candidate axes.  The boxes contain the
// geoms, so a separated pair cannot produce a contact.
MJ_FN int obb_separated(const real* ca, const real* Ra, const real* sa, const real* cb, const real* Rb, const real* sb, real margin) {
  real pp[3], Rm[3][3], Q[3][3], pA[3];
  sub3(pp, cb, ca);
  for (int i = 0; i < 3; ++i) {
    pA[i] = pp[0] * Ra[i] + pp[1] * Ra[3 + i] + pp[2] * Ra[6 + i];
    for (int j = 0; j < 3; ++j) {
      Rm[i][j] = Ra[i] * Rb[j] + Ra[3 + i] * Rb[3 + j] + Ra[6 + i] * Rb[6 + j];
      Q[i][j] = fabsf(Rm[i][j]) + 1e-6f;
    }
  }
  for (int i = 0; i < 3; ++i)
    if (fabsf(pA[i]) - (sa[i] + sb[0] * Q[i][0] + sb[1] * Q[i][1] + sb[2] * Q[i][2]) > margin) return 1;
  for (int j = 0; j < 3; ++j) {
    const real pB = pA[0] * Rm[0][j] + pA[1] * Rm[1][j] + pA[2] * Rm[2][j];
    if (fabsf(pB) - (sb[j] + sa[0] * Q[0][j] + sa[1] * Q[1][j] + sa[2] * Q[2][j]) > margin) return 1;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      // axis A_i x B_j (unnormalised, length l): separation test scaled by l, skipped for near-parallel axes
      const real l2 = 1.0f - Rm[i][j] * Rm[i][j];
      if (l2 < 1e-4f) continue;
      const real d = fabsf(pA[i2] * Rm[i1][j] - pA[i1] * Rm[i2][j]);
      const real r = sa[i1] * Q[i2][j] + sa[i2] * Q[i1][j] + sb[j1] * Q[i][j2] + sb[j2] * Q[i][j1];
      if (d - r > margin * sqrtf(l2) + 1e-6f) return 1;
    }
  return 0;
}

// ------------------------------------------------------------------------------------------------ driver
// append `flag`-ed items of one lane-strided pass to a compact list, in item order (deterministic on every path)
template <int NL>
MJ_FN void compact_append(Work& w, int item, int flag, int lane) {
#if defined(__CUDA_ARCH__)
  if (NL > 1) {
    const unsigned mask = __ballot_sync(0xffffffffu, flag);
    if (mask == 0) return;  // nothing to append in this pass (the common case): no shared-memory traffic, no barriers
    const int base = w.nhit;
    const int pos = base + __popc(mask & ((1u << lane) - 1u));
    if (flag && pos < MAXHIT) w.hit_list[pos] = (unsigned short)item;
    __syncwarp();
    if (lane == 0) { const int n = base + __popc(mask); w.nhit = n < MAXHIT ? n : MAXHIT; if (n > MAXHIT) w.bad |= 2; }
    __syncwarp();
    return;
  }
#endif
  if (flag) {
    if (w.nhit < MAXHIT) w.hit_list[w.nhit++] = (unsigned short)item;
    else w.bad |= 2;  // capacity overflow: a candidate pair was dropped
  }
}

// same for the cached candidate list of the loose broad phase; an overflow invalidates the cache (exactness first)
template <int NL>
MJ_FN void cand_append(Work& w, int item, int flag, int lane) {
#if defined(__CUDA_ARCH__)
  if (NL > 1) {
    const unsigned mask = __ballot_sync(0xffffffffu, flag);
    if (mask == 0) return;
    const int base = w.ncand;
    const int pos = base + __popc(mask & ((1u << lane) - 1u));
    if (flag && pos < MAXCAND) w.cand_list[pos] = (unsigned short)item;
    __syncwarp();
    if (lane == 0) { const int n = base + __popc(mask); w.ncand = n < MAXCAND ? n : MAXCAND; if (n > MAXCAND) w.broad_valid = 0; }
    __syncwarp();
    return;
  }
#endif
  if (flag) {
    if (w.ncand < MAXCAND) w.cand_list[w.ncand++] = (unsigned short)item;
    else w.broad_valid = 0;
  }
}

// exact broad-phase test of candidate pair p (`slack` = 0), or its loose form (bounding spheres / plane distance inflated by
// `slack`, no box culls) used to build the cached candidate list
MJ_HD int pair_test(const Model& m, const Work& w, int p, real slack) {
  const int ga = m.pair_g1[p], gb = m.pair_g2[p];
  const real margin = fmaxf(m.geom_margin[ga], m.geom_margin[gb]) + slack;
  const int t1 = m.geom_type[ga], t2 = m.geom_type[gb];
  const real* pa = gpos(m, w, ga);
  const real* pb = gpos(m, w, gb);
  real rel[3];
  sub3(rel, pb, pa);
  int hit;
  if (t1 == GEOM_PLANE) {
    const real* Rp = gmat(m, w, ga);
    const real n[3] = {Rp[2], Rp[5], Rp[8]};
    return t2 != GEOM_PLANE && dot3(rel, n) <= m.geom_rbound[gb] + margin;
  }
  const real bound = m.geom_rbound[ga] + m.geom_rbound[gb] + margin;
  hit = dot3(rel, rel) <= bound * bound;  // MuJoCo's bounding-sphere test
  if (slack > 0) return hit;
  // exact cull: a geom whose bounding sphere stays clear of the other geom's box cannot touch it
  if (hit && t1 == GEOM_BOX) {
    const real r = m.geom_rbound[gb] + margin;
    hit = point_box_dist2(pb, pa, gmat(m, w, ga), m.geom_size[ga]) <= r * r;
  }
  if (hit && t2 == GEOM_BOX) {
    const real r = m.geom_rbound[ga] + margin;
    hit = point_box_dist2(pa, pb, gmat(m, w, gb), m.geom_size[gb]) <= r * r;
  }
  if (hit) {  // bounding-box cull
    const real* Ra = gmat(m, w, ga);
    const real* Rb = gmat(m, w, gb);
    real ca[3], cb[3];
    mulmatvec3(ca, Ra, m.geom_obb_off[ga]);
    mulmatvec3(cb, Rb, m.geom_obb_off[gb]);
    for (int k = 0; k < 3; ++k) { ca[k] += pa[k]; cb[k] += pb[k]; }
    hit = !obb_separated(ca, Ra, m.geom_obb_size[ga], cb, Rb, m.geom_obb_size[gb], margin);
  }
  return hit;
}

// MuJoCo engine_collision_convex.c: mjc_fixNormal.  After portal refinement the contact normal of a pair that involves a
// cylinder or a capsule is replaced by the primitive's own surface normal at the contact position (radial from the axis of
// a cylinder unless the point lies within 5 % of a flat cap; from the nearest point of a capsule's segment); when both
// geoms have one the two are averaged.  With it the gripper keeps hold of the door handle (two cylinders) in the shipped
// reverse door demonstrations: 4 of 5 episodes pull the door open, none without.
MJ_HD void fix_normal(const real* c1, const real* R1, const real* s1, int t1, const real* c2, const real* R2, const real* s2, int t2,
                      const real* pos, real* normal) {
  const real* C[2] = {c1, c2};
  const real* R[2] = {R1, R2};
  const real* S[2] = {s1, s2};
  const int T[2] = {t1, t2};
  real nrm[2][3] = {{0, 0, 0}, {0, 0, 0}};
  bool done[2] = {false, false};
  for (int i = 0; i < 2; ++i) {
    if (T[i] != GEOM_CYLINDER && !(KITCHEN_ROWS && T[i] == GEOM_CAPSULE)) continue;
    real rel[3], loc[3];
    sub3(rel, pos, C[i]);
    mulmatTvec3(loc, R[i], rel);
    if (T[i] == GEOM_CYLINDER) {
      if (mabs(loc[2]) > 0.95f * S[i][1]) continue;
      loc[2] = 0;
    } else {
      if (loc[2] > S[i][1]) loc[2] -= S[i][1];
      else if (loc[2] < -S[i][1]) loc[2] += S[i][1];
      else loc[2] = 0;
    }
    const real l = msqrt(dot3(loc, loc));
    if (l < 1e-12f) continue;
    real g[3];
    mulmatvec3(g, R[i], loc);
    for (int k = 0; k < 3; ++k) nrm[i][k] = g[k] / l;
    done[i] = true;
  }
  real out[3];
  if (done[0] && done[1]) { for (int k = 0; k < 3; ++k) out[k] = nrm[0][k] - nrm[1][k]; }
  else if (done[0]) { for (int k = 0; k < 3; ++k) out[k] = nrm[0][k]; }
  else if (done[1]) { for (int k = 0; k < 3; ++k) out[k] = -nrm[1][k]; }
  else return;
  const real l = msqrt(dot3(out, out));
  if (l < 1e-12f) return;
  for (int k = 0; k < 3; ++k) normal[k] = out[k] / l;
}

template <int NL>
MJ_FN void collide(const Model& m, const real* hull, Work& w, int lane) {
  geom_poses<NL>(m, w, lane);
  if (lane == 0) { w.ncon = 0; w.nhit = 0; }
  wsync<NL>();
  if (BROAD_CACHE) {
    // Cached broad phase (kitchen: 2,974 candidate pairs).  A loose pass over ALL pairs (bounding spheres inflated by
    // BROAD_SLACK) builds a candidate list that stays a superset of the exact hits for as long as no geom centre has
    // moved by more than BROAD_SLACK / 2: the travel since the last loose pass is bounded rigorously from the joint
    // motion, sum_i lever_i |h qvel_i| with lever_i = reach of dof i's subtree (host, Model::dof_lever), and the list
    // is rebuilt when the bound is used up.  Every substep then runs the EXACT tests on the candidates only, in pair
    // order, so the hit list -- hence the contacts -- is identical to testing all pairs.
    real step = 0;
    for (int i = lane; i < m.nv; i += NL) step += m.dof_lever[i] * mabs(m.timestep * w.qvel[i]);
    step = wsum<NL>(step);
    const real travel = w.broad_travel + step;
    const bool rebuild = !w.broad_valid || travel > 0.5f * BROAD_SLACK;
    wsync<NL>();
    if (rebuild) {
      if (lane == 0) { w.ncand = 0; w.broad_valid = 1; w.broad_travel = 0; w.acc_rebuild += 1; }
      wsync<NL>();
      for (int p0 = 0; p0 < m.npair; p0 += NL) {
        const int p = p0 + lane;
        const int hit = p < m.npair ? pair_test(m, w, p, BROAD_SLACK) : 0;
        cand_append<NL>(w, p, hit, lane);
      }
      wsync<NL>();
    } else if (lane == 0) {
      w.broad_travel = travel;
    }
    wsync<NL>();
    if (w.broad_valid) {
      const int nc = w.ncand;
      for (int c0 = 0; c0 < nc; c0 += NL) {
        const int c = c0 + lane;
        const int p = c < nc ? (int)w.cand_list[c] : 0;
        const int hit = c < nc ? pair_test(m, w, p, 0.0f) : 0;
        compact_append<NL>(w, p, hit, lane);
      }
    } else {  // candidate list overflowed: this substep tests every pair
      for (int p0 = 0; p0 < m.npair; p0 += NL) {
        const int p = p0 + lane;
        compact_append<NL>(w, p, p < m.npair ? pair_test(m, w, p, 0.0f) : 0, lane);
      }
    }
  } else {
    // broad phase, lane-parallel over the candidate pairs (already ordered by geom type on the host)
    for (int p0 = 0; p0 < m.npair; p0 += NL) {
      const int p = p0 + lane;
      compact_append<NL>(w, p, p < m.npair ? pair_test(m, w, p, 0.0f) : 0, lane);
    }
  }
  wsync<NL>();
  // narrow phase: surviving pairs one after the other, uniform across the warp
  const int nhit = w.nhit;
  for (int h = 0; h < nhit; ++h) {
    const int p = w.hit_list[h];
    const int ga = m.pair_g1[p], gb = m.pair_g2[p];
    const real margin = fmaxf(m.geom_margin[ga], m.geom_margin[gb]);
    const int t1 = m.geom_type[ga], t2 = m.geom_type[gb];
    NarrowScratch* S = reinterpret_cast<NarrowScratch*>(&w.H[0][0]);
    RawCon* rc = S->rc;
    int n = 0;
    if (t1 == GEOM_PLANE) {
      n = KITCHEN_ROWS && t2 == GEOM_CAPSULE ? plane_capsule(gpos(m, w, ga), gmat(m, w, ga), gpos(m, w, gb), gmat(m, w, gb), m.geom_size[gb], margin, rc)
                             : plane_convex<NL>(m, hull, w, ga, gb, margin, rc, lane);
    } else if (KITCHEN_ROWS && t1 == GEOM_CAPSULE && t2 == GEOM_CAPSULE) {
      n = capsule_capsule(gpos(m, w, ga), gmat(m, w, ga), m.geom_size[ga], gpos(m, w, gb), gmat(m, w, gb), m.geom_size[gb], margin, rc);
    } else if (t1 == GEOM_BOX && t2 == GEOM_BOX) {
      n = box_box(gpos(m, w, ga), gmat(m, w, ga), m.geom_size[ga], gpos(m, w, gb), gmat(m, w, gb), m.geom_size[gb], margin, rc, S);
    } else {
      if (lane == 0) w.acc_mpr += 1;
      make_cobj(S->o1, m, hull, w, ga, 0.5f * margin);
      make_cobj(S->o2, m, hull, w, gb, 0.5f * margin);
      const CObj& o1 = S->o1;
      const CObj& o2 = S->o2;
      real depth, dir[3], pos[3];
#if defined(MJ_DEBUG) && !defined(__CUDA_ARCH__)
      ++g_mpr_calls;
      const long before = g_support_calls;
#endif
      const int pen = mpr_penetration<NL>(o1, o2, &depth, dir, pos, S, lane);
      if (lane == 0) w.acc_sup += S->nsup;
#if defined(MJ_DEBUG) && !defined(__CUDA_ARCH__)
      g_mpr_hits += pen;
      if (g_support_calls - before > 12) printf("  mpr pair (%d,%d) types %d %d: %ld supports, pen %d depth %g\n", ga, gb, t1, t2, g_support_calls - before, pen, pen ? (double)depth : 0.0);
#endif
      if (pen) {
        for (int k = 0; k < 3; ++k) { rc[0].pos[k] = pos[k]; rc[0].normal[k] = dir[k]; }
        rc[0].dist = margin - depth;
        n = 1;
        fix_normal(gpos(m, w, ga), gmat(m, w, ga), m.geom_size[ga], t1, gpos(m, w, gb), gmat(m, w, gb), m.geom_size[gb], t2, rc[0].pos, rc[0].normal);
      }
    }
    const int base = w.ncon;
    int added = 0;
    for (int c = 0; c < n; ++c) {
      if (rc[c].dist >= margin) continue;
      if (base + added >= MAXCON) { if (lane == 0) w.bad |= 4; continue; }
      const int k = base + added;
      ++added;
      if (lane == 0) {
        for (int q = 0; q < 3; ++q) { w.con_pos[k][q] = rc[c].pos[q]; w.con_frame[k][q] = rc[c].normal[q]; }
        make_frame(w.con_frame[k]);
        w.con_dist[k] = rc[c].dist;
        w.con_g1[k] = ga;
        w.con_g2[k] = gb;
      }
    }
    wsync<NL>();
    if (lane == 0) w.ncon = base + added;
    wsync<NL>();
  }
}

// mj_instantiateContact (elliptic cones): rows (normal, tangent 1, tangent 2[, torsion]) of every contact.
// solref of a geom pair: damping ratio averaged, time constant by its INVERSE (harmonic mean: the geoms' natural frequencies
// are what is mixed).  Decided by the reference's own recordings: 39 of the 40 shipped door / peg episodes end within +-3
// steps with this rule, 28 with the arithmetic mean of the time constants.
MJ_HD void mix_solref(const Model& m, int g1, int g2, real mix, real* solref) {
  const real t1 = m.geom_solref[g1][0], t2 = m.geom_solref[g2][0];
  solref[0] = (t1 > 0 && t2 > 0) ? 1.0f / (mix / t1 + (1 - mix) / t2) : (t1 < t2 ? t1 : t2);
  solref[1] = (t1 > 0 && t2 > 0) ? mix * m.geom_solref[g1][1] + (1 - mix) * m.geom_solref[g2][1]
                                 : (m.geom_solref[g1][1] < m.geom_solref[g2][1] ? m.geom_solref[g1][1] : m.geom_solref[g2][1]);
}

// Contact parameters (mj_contactParam, equal priorities): condim = max, friction = max, solref / solimp mixed by solmix.
template <int NL>
MJ_FN void contact_rows(const Model& m, Work& w, int lane) {
  const int nv = m.nv, row0 = w.nefc;
  // row offsets (serial, a handful of contacts)
  int row = row0, used = 0;
  for (int c = 0; c < w.ncon; ++c) {
    const int g1 = w.con_g1[c], g2 = w.con_g2[c];
    const int dim = m.geom_condim[g1] > m.geom_condim[g2] ? m.geom_condim[g1] : m.geom_condim[g2];
    // elliptic: dim rows (con_dim = dim); pyramidal: 2 (dim - 1) unilateral rows, or 1 when frictionless (con_dim = -rows)
    const bool elliptic = !KITCHEN_ROWS || m.cone_elliptic;
    const int nrow = elliptic ? dim : (dim > 1 ? 2 * (dim - 1) : 1);
    if (row + nrow > MAXEFC) { if (lane == 0) w.bad |= 8; break; }
    if (lane == 0) { w.con_row[c] = row; w.con_dim[c] = elliptic ? dim : -nrow; }
    row += nrow;
    ++used;
  }
  wsync<NL>();
  if (KITCHEN_ROWS && !m.cone_elliptic) {
    // pyramidal cone (mj_instantiateContact): edge e of friction dimension k = 1 + e / 2 has the row J_n +- mu_k J_k, the
    // contact distance as residual and R_py = 2 mu_1^2 R_n / impratio (mj_makeImpedance, pyramidal)
    for (int idx = lane; idx < used * nv; idx += NL) {
      const int c = idx / nv, q = idx - c * nv;
      const int g1 = w.con_g1[c], g2 = w.con_g2[c], r = w.con_row[c], nrow = -w.con_dim[c];
      const int dim = nrow > 1 ? nrow / 2 + 1 : 1;
      const real* fr = w.con_frame[c];
      real jp1[3], jr1[3], jp2[3], jr2[3];
      jac_col(m, w, m.geom_body[g1], w.con_pos[c], q, jp1, jr1);
      jac_col(m, w, m.geom_body[g2], w.con_pos[c], q, jp2, jr2);
      const real dp[3] = {jp2[0] - jp1[0], jp2[1] - jp1[1], jp2[2] - jp1[2]}, dr[3] = {jr2[0] - jr1[0], jr2[1] - jr1[1], jr2[2] - jr1[2]};
      const real jn = dot3(fr, dp);
      if (dim == 1) { w.J[r][q] = jn; continue; }
      real f[3];
      for (int a = 0; a < 3; ++a) f[a] = fmaxf(m.geom_friction[g1][a], m.geom_friction[g2][a]);
      const real fri[5] = {f[0], f[0], f[1], f[2], f[2]};
      for (int k = 1; k < dim; ++k) {
        const real jk = k < 3 ? dot3(fr + 3 * k, dp) : dot3(fr + 3 * (k - 3), dr);
        w.J[r + 2 * (k - 1)][q] = jn + fri[k - 1] * jk;
        w.J[r + 2 * (k - 1) + 1][q] = jn - fri[k - 1] * jk;
      }
    }
    wsync<NL>();
    for (int c = lane; c < used; c += NL) {
      const int g1 = w.con_g1[c], g2 = w.con_g2[c], r = w.con_row[c], nrow = -w.con_dim[c];
      const real margin = fmaxf(m.geom_margin[g1], m.geom_margin[g2]), gap = fmaxf(m.geom_gap[g1], m.geom_gap[g2]);
      const real mix = m.geom_solmix[g1] / (m.geom_solmix[g1] + m.geom_solmix[g2]);
      real solref[2], solimp[5];
      mix_solref(m, g1, g2, mix, solref);
      for (int q = 0; q < 5; ++q) solimp[q] = mix * m.geom_solimp[g1][q] + (1 - mix) * m.geom_solimp[g2][q];
      const real mu1 = fmaxf(m.geom_friction[g1][0], m.geom_friction[g2][0]);
      const real tran = m.geom_invweight0[g1][0] + m.geom_invweight0[g2][0];
      for (int e = 0; e < nrow; ++e) {
        w.e_pos[r + e] = w.con_dist[c];
        w.e_type[r + e] = ROW_LIMIT;
        finish_row(m, w, r + e, solref, solimp, margin - gap, tran, nullptr);
        if (nrow > 1) {
          w.e_R[r + e] = fmaxf(MINVAL, 2 * mu1 * mu1 * w.e_R[r + e] / fmaxf(MINVAL, m.impratio));
          w.e_D[r + e] = 1.0f / w.e_R[r + e];
        }
      }
    }
    wsync<NL>();
    if (lane == 0) { w.nefc = row; w.ncon = used; }
    wsync<NL>();
    return;
  }
  // Jacobian entries: one (contact, dof) pair per lane
  for (int idx = lane; idx < used * nv; idx += NL) {
    const int c = idx / nv, q = idx - c * nv;
    const int b1 = m.geom_body[w.con_g1[c]], b2 = m.geom_body[w.con_g2[c]], r = w.con_row[c], dim = w.con_dim[c];
    const real* fr = w.con_frame[c];
    real jp1[3], jr1[3], jp2[3], jr2[3];
    jac_col(m, w, b1, w.con_pos[c], q, jp1, jr1);
    jac_col(m, w, b2, w.con_pos[c], q, jp2, jr2);
    const real dp[3] = {jp2[0] - jp1[0], jp2[1] - jp1[1], jp2[2] - jp1[2]}, dr[3] = {jr2[0] - jr1[0], jr2[1] - jr1[1], jr2[2] - jr1[2]};
    for (int k = 0; k < dim; ++k) w.J[r + k][q] = k < 3 ? dot3(fr + 3 * k, dp) : dot3(fr + 3 * (k - 3), dr);
  }
  wsync<NL>();
  // one contact per lane: parameters, reference acceleration and regularisation of its rows
  for (int c = lane; c < used; c += NL) {
    const int g1 = w.con_g1[c], g2 = w.con_g2[c], r = w.con_row[c], dim = w.con_dim[c];
    const real margin = fmaxf(m.geom_margin[g1], m.geom_margin[g2]), gap = fmaxf(m.geom_gap[g1], m.geom_gap[g2]);
    const real mix = m.geom_solmix[g1] / (m.geom_solmix[g1] + m.geom_solmix[g2]);
    real solref[2], solimp[5], f[3];
    mix_solref(m, g1, g2, mix, solref);
    for (int q = 0; q < 5; ++q) solimp[q] = mix * m.geom_solimp[g1][q] + (1 - mix) * m.geom_solimp[g2][q];
    for (int q = 0; q < 3; ++q) f[q] = fmaxf(m.geom_friction[g1][q], m.geom_friction[g2][q]);
    real* fri = w.con_fri[c];
    fri[0] = fri[1] = f[0]; fri[2] = f[1]; fri[3] = fri[4] = f[2];
    const real tran = m.geom_invweight0[g1][0] + m.geom_invweight0[g2][0], rot = m.geom_invweight0[g1][1] + m.geom_invweight0[g2][1];
    for (int k = 0; k < dim; ++k) {
      w.e_pos[r + k] = k == 0 ? w.con_dist[c] : 0.0f;
      w.e_type[r + k] = k == 0 ? ROW_CONE : ROW_CONE_FRIC;
      finish_row(m, w, r + k, solref, solimp, k == 0 ? margin - gap : 0.0f, k < 3 ? tran : rot, nullptr);
    }
    // elliptic cone: R of the friction rows from impratio and the friction coefficients (mj_makeImpedance)
    w.e_R[r + 1] = w.e_R[r] / fmaxf(MINVAL, m.impratio);
    w.con_mu[c] = fri[0] * sqrtf(w.e_R[r + 1] / w.e_R[r]);
    for (int k = 1; k < dim - 1; ++k) w.e_R[r + 1 + k] = w.e_R[r + 1] * fri[0] * fri[0] / (fri[k] * fri[k]);
    for (int k = 1; k < dim; ++k) w.e_D[r + k] = 1.0f / w.e_R[r + k];
  }
  wsync<NL>();
  if (lane == 0) { w.nefc = row; w.ncon = used; }
  wsync<NL>();
}

}  // namespace mj
}  // namespace earl
