/*
 * oracle/mjcollide.c -- collision detection + contact constraint rows for oracle/mjengine.c (fp64, TEST INFRASTRUCTURE ONLY).
 *
 * Restates, for the geom types of the EARL Sawyer scenes (plane, cylinder, box, convex mesh):
 *   - MuJoCo 2.1 engine_collision_driver.c: pair filters (contype/conaffinity bitmask, same welded body, welded
 *     parent-child unless one side is the world), bounding-sphere rejection, per-pair contact parameters (condim =
 *     max, friction = max, solref / solimp mixed by solmix, margin / gap = max);
 *   - box-box: separating-axis test over the 15 axes, face contacts by clipping the incident face against the
 *     reference face (up to 8 points), edge-edge contact at the closest points.  MuJoCo's own mjc_BoxBox is a
 *     different implementation of the same geometry; point placement inside a face patch may differ from it;
 *   - every other pair: Minkowski Portal Refinement on support functions, as libccd's ccdMPRPenetration (which
 *     MuJoCo's mjc_Convex calls with mpr_tolerance 1e-6, mpr_iterations 50), one contact per pair, geoms inflated
 *     by margin / 2 each;
 *   - engine_core_constraint.c mj_instantiateContact for elliptic cones: rows (normal, tangents, torsion), diagApprox
 *     from body_invweight0, R of the friction rows from impratio and the friction coefficients.
 * Parity status: see mjengine.c (PARTIAL: pinned only at the observation level by the shipped demonstrations).
 */
#include "mjengine.h"

#include <float.h>
#include <math.h>
#include <string.h>

#define GEOM_PLANE 0
#define GEOM_CAPSULE 3
#define GEOM_CYLINDER 5
#define GEOM_BOX 6
#define GEOM_MESH 7
#define MINVAL 1e-15
#define MPR_TOL 1e-6
#define MPR_ITER 50

static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double *r, const double *a, const double *b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void sub3(double *r, const double *a, const double *b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline double normalize3(double *v) {
  double n = sqrt(dot3(v, v));
  if (n < MINVAL) { v[0] = 1; v[1] = v[2] = 0; return 0; }
  v[0] /= n; v[1] /= n; v[2] /= n;
  return n;
}
static inline void col3(double *c, const double *R, int k) { c[0] = R[k]; c[1] = R[3 + k]; c[2] = R[6 + k]; }

typedef struct { double pos[3], normal[3], dist; } RawCon;

/* mju_makeFrame: complete the contact frame from its normal */
static void make_frame(double *frame) {
  double *x = frame, *y = frame + 3, *z = frame + 6;
  y[0] = y[1] = y[2] = 0;
  if (x[1] < 0.5 && x[1] > -0.5) y[1] = 1; else y[2] = 1;
  double t = dot3(x, y);
  for (int k = 0; k < 3; ++k) y[k] -= t * x[k];
  normalize3(y);
  cross3(z, x, y);
}

/* ------------------------------------------------------------------------------------------ box-box */
static int clip_poly(double (*poly)[3], int n, const double *pn, double pd, double (*out)[3]) {
  /* keep the half space pn.x <= pd (Sutherland-Hodgman) */
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const double *a = poly[i], *b = poly[(i + 1) % n];
    double da = dot3(pn, a) - pd, db = dot3(pn, b) - pd;
    if (da <= 0) { memcpy(out[m++], a, 3 * sizeof(double)); }
    if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
      double t = da / (da - db);
      for (int k = 0; k < 3; ++k) out[m][k] = a[k] + t * (b[k] - a[k]);
      ++m;
    }
    if (m >= 15) break;
  }
  return m;
}

/* MuJoCo 2.1.0 reports HALF of the vertex-below-face depth as the distance of a box-box face contact: pinned by the peg
 * landing at the start of every shipped peg episode (16 nm over 26 env steps; 0.87 mm with the full depth). */
double g_boxbox_face_scale = 0.5;
void mje_debug_boxbox_face_scale(double v) { g_boxbox_face_scale = v; }
static int box_box(const double *p1, const double *R1, const double *s1, const double *p2, const double *R2, const double *s2,
                   double margin, RawCon *out) {
  double A[3][3], B[3][3], pp[3], pA[3], pB[3], Rm[3][3], Q[3][3];
  for (int k = 0; k < 3; ++k) { col3(A[k], R1, k); col3(B[k], R2, k); }
  sub3(pp, p2, p1);
  for (int i = 0; i < 3; ++i) {
    pA[i] = dot3(pp, A[i]);
    pB[i] = dot3(pp, B[i]);
    for (int j = 0; j < 3; ++j) { Rm[i][j] = dot3(A[i], B[j]); Q[i][j] = fabs(Rm[i][j]); }
  }
  double best = -1e300;
  int code = -1, flip = 0;
  double nrm[3] = {0, 0, 0};
  /* face axes of box 1, then of box 2 */
  for (int i = 0; i < 3; ++i) {
    double s = fabs(pA[i]) - (s1[i] + s2[0] * Q[i][0] + s2[1] * Q[i][1] + s2[2] * Q[i][2]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = i; flip = pA[i] < 0; memcpy(nrm, A[i], sizeof nrm); }
  }
  for (int j = 0; j < 3; ++j) {
    double s = fabs(pB[j]) - (s2[j] + s1[0] * Q[0][j] + s1[1] * Q[1][j] + s1[2] * Q[2][j]);
    if (s > margin) return 0;
    if (s > best + 1e-9) { best = s; code = 3 + j; flip = pB[j] < 0; memcpy(nrm, B[j], sizeof nrm); }
  }
  /* edge x edge axes; a face axis is preferred unless the edge axis is clearly better */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double ax[3];
      cross3(ax, A[i], B[j]);
      double l = sqrt(dot3(ax, ax));
      if (l < 1e-6) continue;
      for (int k = 0; k < 3; ++k) ax[k] /= l;
      double d = dot3(pp, ax), ra = 0, rb = 0;
      for (int k = 0; k < 3; ++k) { ra += s1[k] * fabs(dot3(A[k], ax)); rb += s2[k] * fabs(dot3(B[k], ax)); }
      double s = fabs(d) - (ra + rb);
      if (s > margin) return 0;
      if (s > best + 1e-9 && s * 1.05 > best + (best < 0 ? 0 : 1e-9) && s - 0.05 * fabs(s) > best) {
        best = s; code = 6 + 3 * i + j; flip = d < 0; memcpy(nrm, ax, sizeof nrm);
      }
    }
  if (flip) for (int k = 0; k < 3; ++k) nrm[k] = -nrm[k];  /* nrm points from box 1 to box 2 */
  if (code >= 6) { /* edge-edge: closest points of the two supporting edges */
    int i = (code - 6) / 3, j = (code - 6) % 3;
    double pa[3], pb[3];
    for (int k = 0; k < 3; ++k) { pa[k] = p1[k]; pb[k] = p2[k]; }
    for (int a = 0; a < 3; ++a) {
      if (a == i) continue;
      double sg = dot3(nrm, A[a]) > 0 ? 1 : -1;
      for (int k = 0; k < 3; ++k) pa[k] += sg * s1[a] * A[a][k];
    }
    for (int b = 0; b < 3; ++b) {
      if (b == j) continue;
      double sg = dot3(nrm, B[b]) > 0 ? -1 : 1;
      for (int k = 0; k < 3; ++k) pb[k] += sg * s2[b] * B[b][k];
    }
    /* closest points of lines pa + t ua, pb + u ub */
    const double *ua = A[i], *ub = B[j];
    double w[3];
    sub3(w, pb, pa);
    double uaub = dot3(ua, ub), q1 = dot3(ua, w), q2 = -dot3(ub, w), den = 1 - uaub * uaub;
    double t = den < 1e-12 ? 0 : (q1 + uaub * q2) / den, u = den < 1e-12 ? 0 : (uaub * q1 + q2) / den;
    for (int k = 0; k < 3; ++k) out[0].pos[k] = 0.5 * ((pa[k] + t * ua[k]) + (pb[k] + u * ub[k]));
    memcpy(out[0].normal, nrm, sizeof nrm);
    out[0].dist = best;
    return 1;
  }
  /* face contact: reference box owns the axis, incident face of the other box is clipped against it */
  const double *pr, *pi, *sr, *si;
  double (*Rr)[3], (*Ri)[3], n[3];
  int ai;
  if (code < 3) { pr = p1; pi = p2; sr = s1; si = s2; Rr = A; Ri = B; ai = code; memcpy(n, nrm, sizeof n); }
  else { pr = p2; pi = p1; sr = s2; si = s1; Rr = B; Ri = A; ai = code - 3; for (int k = 0; k < 3; ++k) n[k] = -nrm[k]; }
  /* incident face: the face of the incident box most anti-parallel to n */
  int aj = 0;
  double bestd = -1;
  for (int j = 0; j < 3; ++j) { double d = fabs(dot3(n, Ri[j])); if (d > bestd) { bestd = d; aj = j; } }
  double sg = dot3(n, Ri[aj]) > 0 ? -1 : 1, c[3];
  for (int k = 0; k < 3; ++k) c[k] = pi[k] + sg * si[aj] * Ri[aj][k];
  int k1 = (aj + 1) % 3, k2 = (aj + 2) % 3;
  double poly[16][3], tmp[16][3];
  static const double sx[4] = {1, -1, -1, 1}, sy[4] = {1, 1, -1, -1};
  for (int v = 0; v < 4; ++v)
    for (int k = 0; k < 3; ++k) poly[v][k] = c[k] + sx[v] * si[k1] * Ri[k1][k] + sy[v] * si[k2] * Ri[k2][k];
  int np = 4;
  for (int e = 0; e < 2 && np > 0; ++e) {
    int a = (ai + 1 + e) % 3;
    double pn[3] = {Rr[a][0], Rr[a][1], Rr[a][2]}, base = dot3(pn, pr);
    np = clip_poly(poly, np, pn, base + sr[a], tmp);
    for (int k = 0; k < 3; ++k) pn[k] = -pn[k];
    np = clip_poly(tmp, np, pn, -base + sr[a], poly);
  }
  int nc = 0;
  for (int v = 0; v < np && nc < 8; ++v) {
    double rel[3];
    sub3(rel, poly[v], pr);
    double depth = sr[ai] - dot3(rel, n);  /* > 0: the vertex is below the reference face */
    if (-depth >= margin) continue;
    int dup = 0;  /* drop duplicates produced by clipping through a vertex */
    for (int q = 0; q < nc; ++q) {
      double dd[3];
      for (int k = 0; k < 3; ++k) dd[k] = poly[v][k] + 0.5 * depth * n[k] - out[q].pos[k];
      if (dot3(dd, dd) < 1e-16) dup = 1;
    }
    if (dup) continue;
    for (int k = 0; k < 3; ++k) out[nc].pos[k] = poly[v][k] + 0.5 * depth * n[k];
    memcpy(out[nc].normal, nrm, sizeof nrm);
    out[nc].dist = -(g_boxbox_face_scale) * depth;
    ++nc;
  }
  return nc;
}

/* ------------------------------------------------------------------------------------------ MPR (libccd mpr.c) */
typedef struct { const mjModelF *m; const mjDataF *d; int g; double margin; } CObj;
typedef struct { double v[3], v1[3], v2[3]; } Supp;

static void support_geom(const CObj *o, const double *dir, double *res) {
  const mjModelF *m = o->m;
  int g = o->g;
  const double *R = o->d->geom_xmat[g], *sz = m->geom_size + 3 * g;
  double dl[3] = {R[0] * dir[0] + R[3] * dir[1] + R[6] * dir[2], R[1] * dir[0] + R[4] * dir[1] + R[7] * dir[2],
                  R[2] * dir[0] + R[5] * dir[1] + R[8] * dir[2]}, loc[3];
  switch (m->geom_type[g]) {
    case GEOM_BOX:
      for (int k = 0; k < 3; ++k) loc[k] = dl[k] >= 0 ? sz[k] : -sz[k];
      break;
    case GEOM_CYLINDER: {
      double t = sqrt(dl[0] * dl[0] + dl[1] * dl[1]);
      if (t > MINVAL) { loc[0] = dl[0] / t * sz[0]; loc[1] = dl[1] / t * sz[0]; } else { loc[0] = loc[1] = 0; }
      loc[2] = dl[2] >= 0 ? sz[1] : -sz[1];
      break;
    }
    case GEOM_CAPSULE: { /* segment along local z (half length size[1]) inflated by the radius size[0] */
      double t = sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
      for (int k = 0; k < 3; ++k) loc[k] = t > MINVAL ? dl[k] / t * sz[0] : 0;
      loc[2] += dl[2] >= 0 ? sz[1] : -sz[1];
      break;
    }
    case GEOM_MESH: {
      const double *hv = m->hull_vert + 3 * m->geom_hulladr[g];
      int best = 0;
      double bd = -1e300;
      for (int v = 0; v < m->geom_hullnum[g]; ++v) {
        double dd = hv[3 * v] * dl[0] + hv[3 * v + 1] * dl[1] + hv[3 * v + 2] * dl[2];
        if (dd > bd) { bd = dd; best = v; }
      }
      memcpy(loc, hv + 3 * best, sizeof loc);
      break;
    }
    default:
      loc[0] = loc[1] = loc[2] = 0;
  }
  for (int k = 0; k < 3; ++k)
    res[k] = o->d->geom_xpos[g][k] + R[3 * k] * loc[0] + R[3 * k + 1] * loc[1] + R[3 * k + 2] * loc[2] + dir[k] * o->margin;
}

static long long g_flops; /* narrow-phase flop counter of the current mje_collision call */
static long long support_cost(const CObj *o) {
  return o->m->geom_type[o->g] == GEOM_MESH ? 40 + 6LL * o->m->geom_hullnum[o->g] : 60;
}
static void mpr_support(const CObj *o1, const CObj *o2, const double *dir, Supp *s) {
  double nd[3] = {-dir[0], -dir[1], -dir[2]};
  g_flops += support_cost(o1) + support_cost(o2) + 60;
  support_geom(o1, dir, s->v1);
  support_geom(o2, nd, s->v2);
  sub3(s->v, s->v1, s->v2);
}
static inline int is_zero(double x) { return fabs(x) < DBL_EPSILON; }
static inline int ccd_eq(double a, double b) {
  double ab = fabs(a - b);
  if (ab < DBL_EPSILON) return 1;
  double fa = fabs(a), fb = fabs(b);
  return ab < DBL_EPSILON * (fb > fa ? fb : fa);
}

static void portal_dir(const Supp *P, double *dir) {
  double a[3], b[3];
  sub3(a, P[2].v, P[1].v);
  sub3(b, P[3].v, P[1].v);
  cross3(dir, a, b);
  normalize3(dir);
}
static int reach_tolerance(const Supp *P, const Supp *v4, const double *dir) {
  double dv4 = dot3(v4->v, dir), d1 = dv4 - dot3(P[1].v, dir), d2 = dv4 - dot3(P[2].v, dir), d3 = dv4 - dot3(P[3].v, dir);
  double d = d1 < d2 ? d1 : d2;
  d = d < d3 ? d : d3;
  return ccd_eq(d, MPR_TOL) || d < MPR_TOL;
}
static void expand_portal(Supp *P, const Supp *v4) {
  double v4v0[3];
  cross3(v4v0, v4->v, P[0].v);
  if (dot3(P[1].v, v4v0) > 0) {
    if (dot3(P[2].v, v4v0) > 0) P[1] = *v4; else P[3] = *v4;
  } else {
    if (dot3(P[3].v, v4v0) > 0) P[2] = *v4; else P[1] = *v4;
  }
}
static double seg_dist2(const double *P, const double *x0, const double *b, double *wit) {
  double d[3], a[3];
  sub3(d, b, x0);
  sub3(a, x0, P);
  double t = -dot3(a, d) / dot3(d, d);
  if (t < 0 || is_zero(t)) { memcpy(wit, x0, 3 * sizeof(double)); }
  else if (t > 1 || ccd_eq(t, 1)) { memcpy(wit, b, 3 * sizeof(double)); }
  else for (int k = 0; k < 3; ++k) wit[k] = x0[k] + t * d[k];
  double r[3];
  sub3(r, wit, P);
  return dot3(r, r);
}
static double tri_dist2(const double *P, const double *x0, const double *B, const double *C, double *wit) {
  double d1[3], d2[3], a[3];
  sub3(d1, B, x0);
  sub3(d2, C, x0);
  sub3(a, x0, P);
  double v = dot3(d1, d1), w = dot3(d2, d2), p = dot3(a, d1), q = dot3(a, d2), r = dot3(d1, d2);
  double s = (q * r - w * p) / (w * v - r * r), t = (-s * r - q) / w;
  if ((is_zero(s) || s > 0) && (ccd_eq(s, 1) || s < 1) && (is_zero(t) || t > 0) && (ccd_eq(t, 1) || t < 1) &&
      (ccd_eq(t + s, 1) || t + s < 1)) {
    for (int k = 0; k < 3; ++k) wit[k] = x0[k] + s * d1[k] + t * d2[k];
    double rr[3];
    sub3(rr, wit, P);
    return dot3(rr, rr);
  }
  double w2[3], dist = seg_dist2(P, x0, B, wit), d2v = seg_dist2(P, x0, C, w2);
  if (d2v < dist) { dist = d2v; memcpy(wit, w2, sizeof w2); }
  d2v = seg_dist2(P, B, C, w2);
  if (d2v < dist) { dist = d2v; memcpy(wit, w2, sizeof w2); }
  return dist;
}

/* returns 1 with (depth, dir, pos) when the (inflated) geoms penetrate */
static int mpr_penetration(const CObj *o1, const CObj *o2, double *depth, double *dir, double *pos) {
  Supp P[4], v4;
  static const double origin[3] = {0, 0, 0};
  /* discoverPortal */
  memcpy(P[0].v1, o1->d->geom_xpos[o1->g], 3 * sizeof(double));
  memcpy(P[0].v2, o2->d->geom_xpos[o2->g], 3 * sizeof(double));
  sub3(P[0].v, P[0].v1, P[0].v2);
  if (is_zero(P[0].v[0]) && is_zero(P[0].v[1]) && is_zero(P[0].v[2])) P[0].v[0] = 0.00001;
  double d[3] = {-P[0].v[0], -P[0].v[1], -P[0].v[2]}, va[3], vb[3], dt;
  normalize3(d);
  mpr_support(o1, o2, d, &P[1]);
  dt = dot3(P[1].v, d);
  if (is_zero(dt) || dt < 0) return 0;
  cross3(d, P[0].v, P[1].v);
  if (is_zero(dot3(d, d))) {
    if (is_zero(P[1].v[0]) && is_zero(P[1].v[1]) && is_zero(P[1].v[2])) return 0;  /* touching contact: depth 0, no normal */
    /* origin on the segment v0-v1 */
    *depth = sqrt(dot3(P[1].v, P[1].v));
    memcpy(dir, P[1].v, 3 * sizeof(double));
    normalize3(dir);
    for (int k = 0; k < 3; ++k) pos[k] = 0.5 * (P[1].v1[k] + P[1].v2[k]);
    return 1;
  }
  normalize3(d);
  mpr_support(o1, o2, d, &P[2]);
  dt = dot3(P[2].v, d);
  if (is_zero(dt) || dt < 0) return 0;
  sub3(va, P[1].v, P[0].v);
  sub3(vb, P[2].v, P[0].v);
  cross3(d, va, vb);
  normalize3(d);
  if (dot3(d, P[0].v) > 0) {
    Supp t = P[1]; P[1] = P[2]; P[2] = t;
    for (int k = 0; k < 3; ++k) d[k] = -d[k];
  }
  for (int guard = 0; guard < 100; ++guard) {
    mpr_support(o1, o2, d, &P[3]);
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
  /* refinePortal */
  for (int guard = 0;; ++guard) {
    portal_dir(P, d);
    dt = dot3(d, P[1].v);
    if (is_zero(dt) || dt > 0) break; /* portal encapsules the origin */
    mpr_support(o1, o2, d, &v4);
    dt = dot3(v4.v, d);
    if (!(is_zero(dt) || dt > 0) || reach_tolerance(P, &v4, d) || guard > 200) return 0;
    expand_portal(P, &v4);
  }
  /* findPenetr */
  for (int it = 0;; ++it) {
    portal_dir(P, d);
    mpr_support(o1, o2, d, &v4);
    if (reach_tolerance(P, &v4, d) || it > MPR_ITER) {
      double wit[3];
      *depth = sqrt(tri_dist2(origin, P[1].v, P[2].v, P[3].v, wit));
      if (is_zero(*depth)) return 0; /* touching: direction undefined, MuJoCo drops the contact */
      memcpy(dir, wit, sizeof wit);
      normalize3(dir);
      /* findPos: barycentric coordinates of the origin in the portal tetrahedron */
      double b[4], t[3];
      portal_dir(P, d);
      cross3(t, P[1].v, P[2].v); b[0] = dot3(t, P[3].v);
      cross3(t, P[3].v, P[2].v); b[1] = dot3(t, P[0].v);
      cross3(t, P[0].v, P[1].v); b[2] = dot3(t, P[3].v);
      cross3(t, P[2].v, P[1].v); b[3] = dot3(t, P[0].v);
      double sum = b[0] + b[1] + b[2] + b[3];
      if (is_zero(sum) || sum < 0) {
        b[0] = 0;
        cross3(t, P[2].v, P[3].v); b[1] = dot3(t, d);
        cross3(t, P[3].v, P[1].v); b[2] = dot3(t, d);
        cross3(t, P[1].v, P[2].v); b[3] = dot3(t, d);
        sum = b[1] + b[2] + b[3];
      }
      double inv = 1 / sum;
      for (int k = 0; k < 3; ++k) {
        double q1 = 0, q2 = 0;
        for (int v = 0; v < 4; ++v) { q1 += b[v] * P[v].v1[k]; q2 += b[v] * P[v].v2[k]; }
        pos[k] = 0.5 * (q1 + q2) * inv;
      }
      return 1;
    }
    expand_portal(P, &v4);
  }
}

/* ------------------------------------------------------------------------------------------ plane-* (mjc_PlaneBox etc.) */
static int plane_convex(const mjModelF *m, const mjDataF *d, int gp, int g, double margin, RawCon *out) {
  /* plane normal = z axis of the plane geom; report the lowest support point(s) below the margin */
  const double *Rp = d->geom_xmat[gp];
  double n[3] = {Rp[2], Rp[5], Rp[8]}, nd[3] = {-n[0], -n[1], -n[2]}, pt[3], rel[3];
  CObj o = {m, d, g, 0.0};
  int nc = 0;
  if (m->geom_type[g] == GEOM_BOX) {
    const double *R = d->geom_xmat[g], *sz = m->geom_size + 3 * g;
    for (int v = 0; v < 8 && nc < 4; ++v) {
      double loc[3] = {(v & 1 ? 1 : -1) * sz[0], (v & 2 ? 1 : -1) * sz[1], (v & 4 ? 1 : -1) * sz[2]};
      for (int k = 0; k < 3; ++k) pt[k] = d->geom_xpos[g][k] + R[3 * k] * loc[0] + R[3 * k + 1] * loc[1] + R[3 * k + 2] * loc[2];
      sub3(rel, pt, d->geom_xpos[gp]);
      double dist = dot3(rel, n);
      if (dist >= margin) continue;
      for (int k = 0; k < 3; ++k) { out[nc].pos[k] = pt[k] - 0.5 * dist * n[k]; out[nc].normal[k] = n[k]; }
      out[nc].dist = dist;
      ++nc;
    }
    return nc;
  }
  support_geom(&o, nd, pt);
  sub3(rel, pt, d->geom_xpos[gp]);
  double dist = dot3(rel, n);
  if (dist >= margin) return 0;
  for (int k = 0; k < 3; ++k) { out[0].pos[k] = pt[k] - 0.5 * dist * n[k]; out[0].normal[k] = n[k]; }
  out[0].dist = dist;
  return 1;
}

/* squared distance from point c to a box (centre p, axes = columns of R, half sizes s) */
static double point_box_dist2(const double *c, const double *p, const double *R, const double *s) {
  double rel[3], d2 = 0;
  sub3(rel, c, p);
  for (int k = 0; k < 3; ++k) {
    double x = fabs(rel[0] * R[k] + rel[1] * R[3 + k] + rel[2] * R[6 + k]) - s[k];
    if (x > 0) d2 += x * x;
  }
  return d2;
}

/* ------------------------------------------------------------------------------------------ driver */
/* sensitivity probe (tests only): round the geom poses to fp32 before collision, i.e. give the fp64 narrow phase the
 * inputs the fp32 engine has */
/* mjc_PlaneCapsule: one contact per end sphere */
static int plane_capsule(const mjModelF *m, const mjDataF *d, int gp, int g, double margin, RawCon *out) {
  const double *Rp = d->geom_xmat[gp], *R = d->geom_xmat[g], *sz = m->geom_size + 3 * g;
  double n[3] = {Rp[2], Rp[5], Rp[8]}, ax[3] = {R[2], R[5], R[8]};
  int nc = 0;
  for (int e = -1; e <= 1; e += 2) {
    double c[3], rel[3];
    for (int k = 0; k < 3; ++k) c[k] = d->geom_xpos[g][k] + e * sz[1] * ax[k];
    sub3(rel, c, d->geom_xpos[gp]);
    double dist = dot3(rel, n) - sz[0];
    if (dist >= margin) continue;
    for (int k = 0; k < 3; ++k) { out[nc].pos[k] = c[k] - n[k] * (sz[0] + 0.5 * dist); out[nc].normal[k] = n[k]; }
    out[nc].dist = dist;
    ++nc;
  }
  return nc;
}

/* sphere-sphere kernel shared by the capsule routines (mjraw_SphereSphere) */
static int sphere_sphere(const double *c1, double r1, const double *c2, double r2, double margin, RawCon *out) {
  double n[3];
  sub3(n, c2, c1);
  double len = sqrt(dot3(n, n));
  double dist = len - r1 - r2;
  if (dist >= margin) return 0;
  if (len < MINVAL) { n[0] = 1; n[1] = n[2] = 0; } else { for (int k = 0; k < 3; ++k) n[k] /= len; }
  for (int k = 0; k < 3; ++k) { out->pos[k] = c1[k] + n[k] * (r1 + 0.5 * dist); out->normal[k] = n[k]; }
  out->dist = dist;
  return 1;
}

/* mjc_CapsuleCapsule: closest points of the two axis segments; parallel axes give two contacts (the ends of the overlap) */
static int capsule_capsule(const mjModelF *m, const mjDataF *d, int g1, int g2, double margin, RawCon *out) {
  const double *R1 = d->geom_xmat[g1], *R2 = d->geom_xmat[g2], *s1 = m->geom_size + 3 * g1, *s2 = m->geom_size + 3 * g2;
  const double *p1 = d->geom_xpos[g1], *p2 = d->geom_xpos[g2];
  double a1[3] = {R1[2] * s1[1], R1[5] * s1[1], R1[8] * s1[1]}, a2[3] = {R2[2] * s2[1], R2[5] * s2[1], R2[8] * s2[1]}, dif[3];
  sub3(dif, p1, p2);
  double ma = dot3(a1, a1), mb = -dot3(a1, a2), mc = dot3(a2, a2), u = -dot3(a1, dif), v = dot3(a2, dif);
  double det = ma * mc - mb * mb;
  if (fabs(det) >= MINVAL) { /* general configuration */
    double x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > 1) { x1 = 1; x2 = (v - mb) / mc; } else if (x1 < -1) { x1 = -1; x2 = (v + mb) / mc; }
    if (x2 > 1) { x2 = 1; x1 = (u - mb) / ma; if (x1 > 1) x1 = 1; else if (x1 < -1) x1 = -1; }
    else if (x2 < -1) { x2 = -1; x1 = (u + mb) / ma; if (x1 > 1) x1 = 1; else if (x1 < -1) x1 = -1; }
    double c1[3], c2[3];
    for (int k = 0; k < 3; ++k) { c1[k] = p1[k] + a1[k] * x1; c2[k] = p2[k] + a2[k] * x2; }
    return sphere_sphere(c1, s1[0], c2, s2[0], margin, out);
  }
  /* parallel axes: contacts at both ends of the projected overlap */
  int nc = 0;
  double x1, x2, c1[3], c2[3];
  for (int e = -1; e <= 1; e += 2) {
    x1 = e;                               /* end of capsule 1 projected on capsule 2 */
    x2 = (v - mb * x1) / mc;
    if (x2 > 1) x2 = 1; else if (x2 < -1) x2 = -1;
    x1 = (u - mb * x2) / ma;
    if (x1 > 1) x1 = 1; else if (x1 < -1) x1 = -1;
    for (int k = 0; k < 3; ++k) { c1[k] = p1[k] + a1[k] * x1; c2[k] = p2[k] + a2[k] * x2; }
    nc += sphere_sphere(c1, s1[0], c2, s2[0], margin, out + nc);
  }
  if (nc == 2) { /* identical points: keep one */
    double dd[3];
    sub3(dd, out[0].pos, out[1].pos);
    if (dot3(dd, dd) < 1e-20) nc = 1;
  }
  return nc;
}

static int g_round_poses = 0;
void mje_debug_round_poses(int on) { g_round_poses = on; }


/* MuJoCo engine_collision_convex.c: mjc_fixNormal.  After portal refinement the contact normal of a pair that involves a
 * capsule or a cylinder (spheres / ellipsoids do not occur in these scenes) is replaced by the primitive's own surface
 * normal at the contact position: radial from the axis for a cylinder (unless the point lies within 5 % of a flat cap),
 * from the nearest point of the segment for a capsule; if both geoms have one, the two are averaged. */
int g_fix_normal = 1;
void mje_debug_fix_normal(int on) { g_fix_normal = on; }
static void fix_normal(const mjModelF *m, const mjDataF *d, int ga, int gb, const double *pos, double *normal) {
  int g[2] = {ga, gb}, done[2] = {0, 0};
  double nrm[2][3] = {{0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < 2; ++i) {
    const double *R = d->geom_xmat[g[i]], *c = d->geom_xpos[g[i]], *sz = m->geom_size + 3 * g[i];
    double rel[3], loc[3];
    sub3(rel, pos, c);
    for (int k = 0; k < 3; ++k) loc[k] = R[k] * rel[0] + R[3 + k] * rel[1] + R[6 + k] * rel[2];
    int t = m->geom_type[g[i]];
    if (t == GEOM_CYLINDER) {
      if (fabs(loc[2]) > 0.95 * sz[1]) continue;
      loc[2] = 0;
    } else if (t == GEOM_CAPSULE) {
      if (loc[2] > sz[1]) loc[2] -= sz[1];
      else if (loc[2] < -sz[1]) loc[2] += sz[1];
      else loc[2] = 0;
    } else continue;
    double l = sqrt(dot3(loc, loc));
    if (l < 1e-12) continue;
    for (int k = 0; k < 3; ++k) nrm[i][k] = (R[3 * k] * loc[0] + R[3 * k + 1] * loc[1] + R[3 * k + 2] * loc[2]) / l;
    done[i] = 1;
  }
  double out[3];
  if (done[0] && done[1]) for (int k = 0; k < 3; ++k) out[k] = nrm[0][k] - nrm[1][k];
  else if (done[0]) memcpy(out, nrm[0], sizeof out);
  else if (done[1]) for (int k = 0; k < 3; ++k) out[k] = -nrm[1][k];
  else return;
  double l = sqrt(dot3(out, out));
  if (l < 1e-12) return;
  for (int k = 0; k < 3; ++k) normal[k] = out[k] / l;
}

void mje_collision(const mjModelF *m, mjDataF *d) {
  d->ncon = 0;
  g_flops = 0;
  if (g_round_poses)
    for (int g = 0; g < m->ngeom; ++g) {
      for (int k = 0; k < 3; ++k) d->geom_xpos[g][k] = (double)(float)d->geom_xpos[g][k];
      for (int k = 0; k < 9; ++k) d->geom_xmat[g][k] = (double)(float)d->geom_xmat[g][k];
    }
  for (int g1 = 0; g1 < m->ngeom; ++g1)
    for (int g2 = g1 + 1; g2 < m->ngeom; ++g2) {
      if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) continue;
      int b1 = m->geom_body[g1], b2 = m->geom_body[g2];
      if (b1 == b2) continue;
      if (b1 != 0 && b2 != 0 && (m->body_parent[b1] == b2 || m->body_parent[b2] == b1)) continue;
      double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]), gap = fmax(m->geom_gap[g1], m->geom_gap[g2]);
      int t1 = m->geom_type[g1], t2 = m->geom_type[g2], ga = g1, gb = g2;
      if (t1 > t2) { int t = t1; t1 = t2; t2 = t; ga = g2; gb = g1; } /* MuJoCo orders the pair by geom type */
      /* bounding-sphere / plane rejection */
      if (t1 == GEOM_PLANE) {
        const double *Rp = d->geom_xmat[ga];
        double n[3] = {Rp[2], Rp[5], Rp[8]}, rel[3];
        sub3(rel, d->geom_xpos[gb], d->geom_xpos[ga]);
        if (dot3(rel, n) > m->geom_rbound[gb] + margin) continue;
      } else {
        double rel[3];
        sub3(rel, d->geom_xpos[gb], d->geom_xpos[ga]);
        double bound = m->geom_rbound[ga] + m->geom_rbound[gb] + margin;
        if (dot3(rel, rel) > bound * bound) continue;
        /* exact cull (not in MuJoCo; cannot change the contact set): a geom whose bounding sphere stays clear of the
         * other geom's box cannot touch it.  Keeps the flop count that of a sensible implementation. */
        double r;
        if (t1 == GEOM_BOX) {
          r = m->geom_rbound[gb] + margin;
          if (point_box_dist2(d->geom_xpos[gb], d->geom_xpos[ga], d->geom_xmat[ga], m->geom_size + 3 * ga) > r * r) continue;
        }
        if (t2 == GEOM_BOX) {
          r = m->geom_rbound[ga] + margin;
          if (point_box_dist2(d->geom_xpos[ga], d->geom_xpos[gb], d->geom_xmat[gb], m->geom_size + 3 * gb) > r * r) continue;
        }
      }
      RawCon rc[8];
      int n = 0;
      g_flops += 25; /* bounding-sphere test */
      if (t1 == GEOM_BOX && t2 == GEOM_BOX) g_flops += 700;
      if (t1 == GEOM_PLANE) {
        if (t2 == GEOM_PLANE) continue;
        n = t2 == GEOM_CAPSULE ? plane_capsule(m, d, ga, gb, margin, rc) : plane_convex(m, d, ga, gb, margin, rc);
      } else if (t1 == GEOM_CAPSULE && t2 == GEOM_CAPSULE) {
        n = capsule_capsule(m, d, ga, gb, margin, rc);
      } else if (t1 == GEOM_BOX && t2 == GEOM_BOX) {
        n = box_box(d->geom_xpos[ga], d->geom_xmat[ga], m->geom_size + 3 * ga, d->geom_xpos[gb], d->geom_xmat[gb],
                    m->geom_size + 3 * gb, margin, rc);
      } else {
        CObj o1 = {m, d, ga, 0.5 * margin}, o2 = {m, d, gb, 0.5 * margin};
        double depth, dir[3], pos[3];
        if (mpr_penetration(&o1, &o2, &depth, dir, pos)) {
          memcpy(rc[0].pos, pos, sizeof pos);
          memcpy(rc[0].normal, dir, sizeof dir);
          rc[0].dist = margin - depth;
          n = 1;
          if (g_fix_normal) fix_normal(m, d, ga, gb, rc[0].pos, rc[0].normal);
        }
      }
      for (int c = 0; c < n && d->ncon < MJ_MAXCON; ++c) {
        if (rc[c].dist >= margin) continue;
        int k = d->ncon++;
        memcpy(d->con_pos[k], rc[c].pos, 3 * sizeof(double));
        memcpy(d->con_frame[k], rc[c].normal, 3 * sizeof(double));
        make_frame(d->con_frame[k]);
        d->con_dist[k] = rc[c].dist;
        d->con_geom1[k] = ga;
        d->con_geom2[k] = gb;
        /* contact parameters (mj_contactParam; all priorities are equal in these scenes) */
        int cd = m->geom_condim[ga] > m->geom_condim[gb] ? m->geom_condim[ga] : m->geom_condim[gb];
        d->con_dim[k] = cd;
        double f[3];
        for (int q = 0; q < 3; ++q) f[q] = fmax(m->geom_friction[3 * ga + q], m->geom_friction[3 * gb + q]);
        d->con_friction[k][0] = d->con_friction[k][1] = f[0];
        d->con_friction[k][2] = f[1];
        d->con_friction[k][3] = d->con_friction[k][4] = f[2];
        double mix = m->geom_solmix[ga] / (m->geom_solmix[ga] + m->geom_solmix[gb]);
        /* solref: the damping ratio is averaged; of the time constant the INVERSE is averaged (harmonic mean: the two geoms'
         * natural frequencies / damping coefficients B = 2 / (dmax tc) are what is mixed).  MuJoCo's documentation says
         * "weighted average" without saying of what; the reference's own MuJoCo 2.1.0 recordings decide: with the harmonic
         * mean 39 of the 40 shipped door / peg episodes end within +-3 steps of the recording (door 10 of 10), with the
         * arithmetic mean 28 (door 0 of 10: the panel-table pair (0.02, 0.01) then has 11 % too little friction damping and
         * a 21 % smaller cone) -- tests/test_engine_oracle.py.  Every mixed pair in these scenes is (0.02, 0.01). */
        {
          extern double mje_opt[16];
          double t1 = m->geom_solref[2 * ga], t2 = m->geom_solref[2 * gb], tc;
          if (mje_opt[8] == 1) tc = mix * t1 + (1 - mix) * t2;              /* experiment switches: arithmetic, */
          else if (mje_opt[8] == 2) tc = t1 < t2 ? t1 : t2;                  /* min, */
          else if (mje_opt[8] == 3) tc = sqrt(t1 * t2);                      /* geometric, */
          else if (mje_opt[8] == 4) tc = t1 > t2 ? t1 : t2;                  /* max, */
          else if (mje_opt[8] == 5 && t1 != t2) tc = mje_opt[9];             /* explicit value for mixed pairs */
          else tc = 1.0 / (mix / t1 + (1 - mix) / t2);
          d->con_solref[k][0] = tc;
          d->con_solref[k][1] = mix * m->geom_solref[2 * ga + 1] + (1 - mix) * m->geom_solref[2 * gb + 1];
        }
        for (int q = 0; q < 5; ++q) d->con_solimp[k][q] = mix * m->geom_solimp[5 * ga + q] + (1 - mix) * m->geom_solimp[5 * gb + q];
        d->con_margin[k] = margin - gap; /* includemargin */
      }
    }
  d->flops += g_flops;
}

/* Jacobian columns of body b at a world point (mjengine.c) */
void mje_body_jac(const mjModelF *m, const mjDataF *d, int b, const double *pt, double jp[3][MJ_MAXV], double jr[3][MJ_MAXV]);
void mje_finish_row(const mjModelF *m, mjDataF *d, int i, const double *solref, const double *solimp, double margin, double diag);

extern int mje_row_is_contact;
int mje_contact_rows_impl(const mjModelF *m, mjDataF *d, int row);
int mje_contact_rows(const mjModelF *m, mjDataF *d, int row) {
  mje_row_is_contact = 1;
  int r = mje_contact_rows_impl(m, d, row);
  mje_row_is_contact = 0;
  return r;
}
int mje_contact_rows_impl(const mjModelF *m, mjDataF *d, int row) {
  static double jp1[3][MJ_MAXV], jr1[3][MJ_MAXV], jp2[3][MJ_MAXV], jr2[3][MJ_MAXV];
  int nv = m->nv;
  for (int c = 0; c < d->ncon; ++c) {
    int dim = d->con_dim[c], g1 = d->con_geom1[c], g2 = d->con_geom2[c];
    mje_row_is_contact = (g1 == 1 || g2 == 1) ? 2 : 1;
    if (!m->cone_elliptic) {
      /* pyramidal cone (mj_instantiateContact): 2 (dim - 1) rows J_n +- mu_k J_k, every one a unilateral row with the
       * contact distance as residual; frictionless contacts keep the single normal row.
       * R (mj_makeImpedance, pyramidal): R_py = 2 mu_1^2 R_n / impratio with R_n from the translational weights. */
      int nrow = dim > 1 ? 2 * (dim - 1) : 1;
      if (row + nrow > MJ_MAXEFC) break;
      d->flops += (long long)nv * (30 + 6 * nrow);
      mje_body_jac(m, d, m->geom_body[g1], d->con_pos[c], jp1, jr1);
      mje_body_jac(m, d, m->geom_body[g2], d->con_pos[c], jp2, jr2);
      const double *fr = d->con_frame[c], *f = d->con_friction[c];
      double jn[MJ_MAXV], jk[MJ_MAXV];
      for (int q = 0; q < nv; ++q) {
        double sn = 0;
        for (int a = 0; a < 3; ++a) sn += fr[a] * (jp2[a][q] - jp1[a][q]);
        jn[q] = sn;
      }
      double tran = m->geom_invweight0[2 * g1] + m->geom_invweight0[2 * g2];
      for (int e = 0; e < nrow; ++e) {
        int k = 1 + e / 2;                /* friction dimension of this edge */
        double sgn = (e & 1) ? -1.0 : 1.0, mu = dim > 1 ? f[k - 1] : 0.0;
        if (dim > 1) {
          const double *ax = k < 3 ? fr + 3 * k : fr + 3 * (k - 3);
          for (int q = 0; q < nv; ++q) {
            double sk = 0;
            if (k < 3) for (int a = 0; a < 3; ++a) sk += ax[a] * (jp2[a][q] - jp1[a][q]);
            else for (int a = 0; a < 3; ++a) sk += ax[a] * (jr2[a][q] - jr1[a][q]);
            jk[q] = sk;
          }
        }
        for (int q = 0; q < nv; ++q) d->efc_J[row + e][q] = jn[q] + (dim > 1 ? sgn * mu * jk[q] : 0.0);
        d->efc_pos[row + e] = d->con_dist[c];
        d->efc_type[row + e] = 1;
        d->efc_dim[row + e] = 1;
        mje_finish_row(m, d, row + e, d->con_solref[c], d->con_solimp[c], d->con_margin[c], tran);
        if (dim > 1) {
          d->efc_R[row + e] = fmax(MINVAL, 2 * f[0] * f[0] * d->efc_R[row + e] / fmax(MINVAL, m->impratio));
          d->efc_D[row + e] = 1 / d->efc_R[row + e];
        }
      }
      row += nrow;
      continue;
    }
    if (row + dim > MJ_MAXEFC) break;
    d->flops += (long long)nv * (30 + 6 * dim);
    mje_body_jac(m, d, m->geom_body[g1], d->con_pos[c], jp1, jr1);
    mje_body_jac(m, d, m->geom_body[g2], d->con_pos[c], jp2, jr2);
    const double *fr = d->con_frame[c];
    for (int k = 0; k < dim; ++k) {
      /* rows 0..2: frame axis k on the relative translational velocity; row 3: torsion (normal axis, rotational);
       * rows 4..5: rolling (tangent axes, rotational) */
      const double *ax = k < 3 ? fr + 3 * k : fr + 3 * (k - 3);
      for (int q = 0; q < nv; ++q) {
        double s = 0;
        if (k < 3) for (int a = 0; a < 3; ++a) s += ax[a] * (jp2[a][q] - jp1[a][q]);
        else for (int a = 0; a < 3; ++a) s += ax[a] * (jr2[a][q] - jr1[a][q]);
        d->efc_J[row + k][q] = s;
      }
      d->efc_pos[row + k] = k == 0 ? d->con_dist[c] : 0;
      d->efc_type[row + k] = k == 0 ? 2 : 3;
      d->efc_dim[row + k] = dim;
    }
    double tran = m->geom_invweight0[2 * g1] + m->geom_invweight0[2 * g2], rot = m->geom_invweight0[2 * g1 + 1] + m->geom_invweight0[2 * g2 + 1];
    mje_finish_row(m, d, row, d->con_solref[c], d->con_solimp[c], d->con_margin[c], tran);
    for (int k = 1; k < dim; ++k) mje_finish_row(m, d, row + k, d->con_solref[c], d->con_solimp[c], 0.0, k < 3 ? tran : rot);
    /* elliptic cone: R of the friction rows from impratio and the friction coefficients (mj_makeImpedance) */
    const double *f = d->con_friction[c];
    if (dim > 1) {
      d->efc_R[row + 1] = d->efc_R[row] / fmax(MINVAL, m->impratio);
      d->efc_mu[row] = f[0] * sqrt(d->efc_R[row + 1] / d->efc_R[row]);
      for (int k = 1; k < dim - 1; ++k) d->efc_R[row + 1 + k] = d->efc_R[row + 1] * f[0] * f[0] / (f[k] * f[k]);
      for (int k = 1; k < dim; ++k) d->efc_D[row + k] = 1 / d->efc_R[row + k];
    } else {
      d->efc_mu[row] = 0;
    }
    for (int k = 0; k < 5; ++k) d->efc_fri[row][k] = f[k];
    row += dim;
  }
  return row;
}

/* direct entry points for the geometry unit tests (tests/test_collision_geometry.py) */
int mje_test_box_box(const double *p1, const double *R1, const double *s1, const double *p2, const double *R2, const double *s2,
                     double margin, double *out /* [8][7]: pos, normal, dist */) {
  RawCon rc[8];
  int n = box_box(p1, R1, s1, p2, R2, s2, margin, rc);
  for (int c = 0; c < n; ++c) {
    memcpy(out + 7 * c, rc[c].pos, 3 * sizeof(double));
    memcpy(out + 7 * c + 3, rc[c].normal, 3 * sizeof(double));
    out[7 * c + 6] = rc[c].dist;
  }
  return n;
}

int mje_test_mpr(int t1, const double *size1, const double *p1, const double *R1, int t2, const double *size2, const double *p2,
                 const double *R2, double margin, double *out /* depth, dir[3], pos[3] */) {
  static mjModelF m;
  static mjDataF d;
  static int types[2];
  static double sizes[6];
  types[0] = t1; types[1] = t2;
  memcpy(sizes, size1, 3 * sizeof(double));
  memcpy(sizes + 3, size2, 3 * sizeof(double));
  m.geom_type = types;
  m.geom_size = sizes;
  memcpy(d.geom_xpos[0], p1, 3 * sizeof(double)); memcpy(d.geom_xpos[1], p2, 3 * sizeof(double));
  memcpy(d.geom_xmat[0], R1, 9 * sizeof(double)); memcpy(d.geom_xmat[1], R2, 9 * sizeof(double));
  CObj o1 = {&m, &d, 0, 0.5 * margin}, o2 = {&m, &d, 1, 0.5 * margin};
  return mpr_penetration(&o1, &o2, out, out + 1, out + 4);
}
