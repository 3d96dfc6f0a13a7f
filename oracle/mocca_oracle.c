/*
 * mocca_oracle.c -- CPU float64 restatement of the mocca_envs hot path.  TEST INFRASTRUCTURE ONLY.
 * See mocca_oracle.h for the parity statement (env layer pinned by reference-run traces, Bullet arithmetic unpinned)
 * and the list of what may call this file.
 *
 * Layout of this file
 *   1. small linear algebra
 *   2. kinematics (btMultiBody link conventions: COM-centred link frames, parent->this rotations)
 *   3. articulated-body forward dynamics (btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof)
 *      and unit-impulse response (calcAccelerationDeltasMultiDof)
 *   4. world-frame RNEA + mass matrix (independent formulation, used to cross-check 3)
 *   5. narrow phase (sphere/capsule vs plane z=0 and vs static boxes)
 *   6. constraint rows + projected Gauss-Seidel (btMultiBodyConstraintSolver semantics)
 *   7. integration (btMultiBody::stepPositionsMultiDof) and the stepSimulation driver
 *   8. MT19937 / NumPy legacy RandomState draws
 *   9. Walker3DCustomEnv (reference mocca_envs/env_locomotion.py:37-222, robots.py:31-95,179-227)
 */
#include "mocca_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846

typedef double v3[3];
typedef double m3[3][3];
typedef double v6[6];
typedef double m6[6][6];

/* ------------------------------------------------------------------ 1. linear algebra */
static void v3set(v3 a, double x, double y, double z) { a[0] = x; a[1] = y; a[2] = z; }
static void v3copy(v3 a, const v3 b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
static double v3dot(const v3 a, const v3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double v3norm(const v3 a) { return sqrt(v3dot(a, a)); }
static void v3cross(const v3 a, const v3 b, v3 c) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  c[0] = x; c[1] = y; c[2] = z;
}
static void m3vec(const m3 R, const v3 a, v3 out) {
  double x = R[0][0] * a[0] + R[0][1] * a[1] + R[0][2] * a[2];
  double y = R[1][0] * a[0] + R[1][1] * a[1] + R[1][2] * a[2];
  double z = R[2][0] * a[0] + R[2][1] * a[1] + R[2][2] * a[2];
  out[0] = x; out[1] = y; out[2] = z;
}
static void m3Tvec(const m3 R, const v3 a, v3 out) {
  double x = R[0][0] * a[0] + R[1][0] * a[1] + R[2][0] * a[2];
  double y = R[0][1] * a[0] + R[1][1] * a[1] + R[2][1] * a[2];
  double z = R[0][2] * a[0] + R[1][2] * a[1] + R[2][2] * a[2];
  out[0] = x; out[1] = y; out[2] = z;
}
static void m3mul(const m3 A, const m3 B, m3 C) {
  m3 T;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
  memcpy(C, T, sizeof(m3));
}
/* xyzw quaternion -> rotation matrix (btMatrix3x3::setRotation; normalises implicitly) */
static void quat_to_mat(const double q[4], m3 R) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double d = x * x + y * y + z * z + w * w, s = 2.0 / d;
  double xs = x * s, ys = y * s, zs = z * s;
  double wx = w * xs, wy = w * ys, wz = w * zs, xx = x * xs, xy = x * ys, xz = x * zs;
  double yy = y * ys, yz = y * zs, zz = z * zs;
  R[0][0] = 1 - (yy + zz); R[0][1] = xy - wz; R[0][2] = xz + wy;
  R[1][0] = xy + wz; R[1][1] = 1 - (xx + zz); R[1][2] = yz - wx;
  R[2][0] = xz - wy; R[2][1] = yz + wx; R[2][2] = 1 - (xx + yy);
}
/* rotation about unit axis by angle (Rodrigues) */
static void axis_angle_mat(const v3 a, double ang, m3 R) {
  double c = cos(ang), s = sin(ang), t = 1 - c;
  R[0][0] = t * a[0] * a[0] + c; R[0][1] = t * a[0] * a[1] - s * a[2]; R[0][2] = t * a[0] * a[2] + s * a[1];
  R[1][0] = t * a[0] * a[1] + s * a[2]; R[1][1] = t * a[1] * a[1] + c; R[1][2] = t * a[1] * a[2] - s * a[0];
  R[2][0] = t * a[0] * a[2] - s * a[1]; R[2][1] = t * a[1] * a[2] + s * a[0]; R[2][2] = t * a[2] * a[2] + c;
}
static double v6dot(const v6 a, const v6 b) {
  double s = 0;
  for (int i = 0; i < 6; i++) s += a[i] * b[i];
  return s;
}
static void m6vec(const m6 A, const v6 x, v6 y) {
  v6 t;
  for (int i = 0; i < 6; i++) {
    double s = 0;
    for (int j = 0; j < 6; j++) s += A[i][j] * x[j];
    t[i] = s;
  }
  memcpy(y, t, sizeof(v6));
}
/* Gaussian elimination with partial pivoting: solve A x = b (n<=6), A destroyed */
static void solve_n(int n, double A[6][6], double* b, double* x) {
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++)
      if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
    if (piv != c) {
      for (int k = 0; k < n; k++) { double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
      double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    for (int r = c + 1; r < n; r++) {
      double f = A[r][c] / A[c][c];
      for (int k = c; k < n; k++) A[r][k] -= f * A[c][k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; r--) {
    double s = b[r];
    for (int k = r + 1; k < n; k++) s -= A[r][k] * x[k];
    x[r] = s / A[r][r];
  }
}
static void invert6(const m6 A, m6 Ainv) {
  for (int c = 0; c < 6; c++) {
    double T[6][6], b[6] = {0, 0, 0, 0, 0, 0}, x[6];
    memcpy(T, A, sizeof(T));
    b[c] = 1;
    solve_n(6, T, b, x);
    for (int r = 0; r < 6; r++) Ainv[r][c] = x[r];
  }
}

/* ------------------------------------------------------------------ 2. kinematics */
typedef struct {
  m3 R0;               /* world -> base (Bullet rot_from_parent[0]) */
  m3 Rp[ORC_MAXL];     /* parent -> this */
  v3 r[ORC_MAXL];      /* parent COM -> this COM, in this frame (m_cachedRVector) */
  m3 Rw[ORC_MAXL + 1]; /* world -> link, [0] = base */
  v3 pw[ORC_MAXL + 1]; /* link COM in world */
  v6 S[ORC_MAXL];      /* joint motion subspace in link frame (m_axisTop, m_axisBottom) */
  /* ABA cache (valid after aba()) */
  m6 IA[ORC_MAXL + 1];
  v6 h[ORC_MAXL];
  double Dinv[ORC_MAXL];
  m6 IA0inv;
} orc_cache;

static void kin(const orc_model* m, const orc_state* s, orc_cache* c) {
  m3 Q;
  quat_to_mat(s->quat, Q); /* local -> world */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c->R0[i][j] = Q[j][i];
  memcpy(c->Rw[0], c->R0, sizeof(m3));
  v3copy(c->pw[0], s->pos);
  for (int i = 0; i < m->n_links; i++) {
    m3 Rz;
    quat_to_mat(m->rot_p2t[i], Rz);
    if (m->joint_type[i] == ORC_JOINT_REVOLUTE) {
      /* m_cachedRotParentToThis = btQuaternion(axis, -q) * m_zeroRotParentToThis */
      m3 Rj;
      axis_angle_mat(m->axis[i], -s->q[m->dof_of_link[i]], Rj);
      m3mul(Rj, Rz, c->Rp[i]);
      v3copy(c->S[i], m->axis[i]);
      v3cross(m->axis[i], m->d_vec[i], &c->S[i][3]);
    } else {
      memcpy(c->Rp[i], Rz, sizeof(m3));
      memset(c->S[i], 0, sizeof(v6));
    }
    /* m_cachedRVector = quatRotate(rotParentToThis, eVector) + dVector */
    m3vec(c->Rp[i], m->e_vec[i], c->r[i]);
    for (int k = 0; k < 3; k++) c->r[i][k] += m->d_vec[i][k];
    int p = m->parent[i] + 1;
    m3mul(c->Rp[i], c->Rw[p], c->Rw[i + 1]);
    v3 rw;
    m3Tvec(c->Rw[i + 1], c->r[i], rw);
    for (int k = 0; k < 3; k++) c->pw[i + 1][k] = c->pw[p][k] + rw[k];
  }
}

void orc_fk(const orc_model* m, const orc_state* s, double link_pos[][3], double link_rot[][9]) {
  orc_cache* c = (orc_cache*)malloc(sizeof(orc_cache));
  kin(m, s, c);
  for (int i = 0; i <= m->n_links; i++) {
    v3copy(link_pos[i], c->pw[i]);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) link_rot[i][3 * a + b] = c->Rw[i][b][a]; /* local -> world, row-major */
  }
  free(c);
}

/* spatial transforms; vectors are (angular[3], linear[3]) */
static void xform_motion(const m3 R, const v3 r, const v6 in, v6 out) {
  v3 w, v, t;
  m3vec(R, in, w);
  m3vec(R, in + 3, v);
  v3cross(r, w, t);
  out[0] = w[0]; out[1] = w[1]; out[2] = w[2];
  out[3] = v[0] - t[0]; out[4] = v[1] - t[1]; out[5] = v[2] - t[2];
}
static void xform_force_T(const m3 R, const v3 r, const v6 in, v6 out) {
  /* child -> parent: torque about parent COM = n + r x f */
  v3 t, n;
  v3cross(r, in + 3, t);
  for (int k = 0; k < 3; k++) n[k] = in[k] + t[k];
  m3Tvec(R, n, out);
  m3Tvec(R, in + 3, out + 3);
}
static void xform_matrix(const m3 R, const v3 r, m6 X) {
  /* motion transform parent -> child as 6x6: [[R,0],[-[r]x R, R]] */
  memset(X, 0, sizeof(m6));
  double rx[3][3] = {{0, -r[2], r[1]}, {r[2], 0, -r[0]}, {-r[1], r[0], 0}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      X[i][j] = R[i][j];
      X[i + 3][j + 3] = R[i][j];
      double s = 0;
      for (int k = 0; k < 3; k++) s += rx[i][k] * R[k][j];
      X[i + 3][j] = -s;
    }
}
static void motion_cross(const v6 a, const v6 b, v6 out) {
  v3 t1, t2, t3;
  v3cross(a, b, t1);
  v3cross(a + 3, b, t2);
  v3cross(a, b + 3, t3);
  for (int k = 0; k < 3; k++) { out[k] = t1[k]; out[3 + k] = t2[k] + t3[k]; }
}

/* ------------------------------------------------------------------ 3. ABA */
static void rigid_bias(double mass, const v3 I, const v6 vel, const v3 fext_local, const orc_params* p,
                       int with_damping, v6 Z) {
  /* zero-acceleration force of one link: -external + damping + gyroscopic + m w x v */
  const double* w = vel;
  const double* v = vel + 3;
  v3 Iw = {I[0] * w[0], I[1] * w[1], I[2] * w[2]};
  for (int k = 0; k < 3; k++) { Z[k] = 0; Z[3 + k] = -fext_local[k]; }
  if (with_damping) {
    double wn = v3norm(w), vn = v3norm(v);
    for (int k = 0; k < 3; k++) {
      Z[k] += Iw[k] * (p->ang_damping + p->ang_damping * wn);
      Z[3 + k] += mass * v[k] * (p->lin_damping + p->lin_damping * vn);
    }
  }
  if (p->gyro) {
    v3 g;
    v3cross(w, Iw, g);
    for (int k = 0; k < 3; k++) Z[k] += g[k];
  }
  v3 wv;
  v3cross(w, v, wv);
  for (int k = 0; k < 3; k++) Z[3 + k] += mass * wv[k];
}

static void aba(const orc_model* m, const orc_params* p, const orc_state* s, const double* tau, int with_damping,
                orc_cache* c, double* acc) {
  int nl = m->n_links;
  static _Thread_local v6 vel[ORC_MAXL + 1], Z[ORC_MAXL + 1], cor[ORC_MAXL], a[ORC_MAXL + 1];
  static _Thread_local double Y[ORC_MAXL];
  v3 gw = {0, 0, -p->gravity};
  /* base */
  m3vec(c->R0, s->omega, vel[0]);
  m3vec(c->R0, s->vel, vel[0] + 3);
  v3 f, fl;
  for (int k = 0; k < 3; k++) f[k] = gw[k] * m->base_mass;
  m3vec(c->R0, f, fl);
  rigid_bias(m->base_mass, m->base_inertia, vel[0], fl, p, with_damping, Z[0]);
  memset(c->IA[0], 0, sizeof(m6));
  for (int k = 0; k < 3; k++) { c->IA[0][k][k] = m->base_inertia[k]; c->IA[0][3 + k][3 + k] = m->base_mass; }
  for (int i = 0; i < nl; i++) {
    int pp = m->parent[i] + 1;
    xform_motion(c->Rp[i], c->r[i], vel[pp], vel[i + 1]);
    v6 vj;
    double qd = m->joint_type[i] == ORC_JOINT_REVOLUTE ? s->qd[m->dof_of_link[i]] : 0.0;
    for (int k = 0; k < 6; k++) { vj[k] = c->S[i][k] * qd; vel[i + 1][k] += vj[k]; }
    motion_cross(vel[i + 1], vj, cor[i]);
    for (int k = 0; k < 3; k++) f[k] = gw[k] * m->mass[i];
    m3vec(c->Rw[i + 1], f, fl);
    rigid_bias(m->mass[i], m->inertia[i], vel[i + 1], fl, p, with_damping, Z[i + 1]);
    memset(c->IA[i + 1], 0, sizeof(m6));
    for (int k = 0; k < 3; k++) { c->IA[i + 1][k][k] = m->inertia[i][k]; c->IA[i + 1][3 + k][3 + k] = m->mass[i]; }
  }
  /* inward pass */
  for (int i = nl - 1; i >= 0; i--) {
    int pp = m->parent[i] + 1;
    m6 Ia;
    v6 Za, Ic;
    memcpy(Ia, c->IA[i + 1], sizeof(m6));
    m6vec(c->IA[i + 1], cor[i], Ic);
    for (int k = 0; k < 6; k++) Za[k] = Z[i + 1][k] + Ic[k];
    if (m->joint_type[i] == ORC_JOINT_REVOLUTE) {
      int d = m->dof_of_link[i];
      m6vec(c->IA[i + 1], c->S[i], c->h[i]);
      double D = v6dot(c->S[i], c->h[i]) + m->armature[d];
      c->Dinv[i] = 1.0 / D;
      Y[i] = tau[d] - v6dot(c->S[i], Z[i + 1]) - v6dot(cor[i], c->h[i]);
      for (int a_ = 0; a_ < 6; a_++) {
        for (int b = 0; b < 6; b++) Ia[a_][b] -= c->h[i][a_] * c->h[i][b] * c->Dinv[i];
        Za[a_] += c->h[i][a_] * Y[i] * c->Dinv[i];
      }
    }
    /* IA[parent] += X^T Ia X ; Z[parent] += X^T Za */
    m6 X, T;
    xform_matrix(c->Rp[i], c->r[i], X);
    for (int a_ = 0; a_ < 6; a_++)
      for (int b = 0; b < 6; b++) {
        double sum = 0;
        for (int k = 0; k < 6; k++) sum += Ia[a_][k] * X[k][b];
        T[a_][b] = sum;
      }
    for (int a_ = 0; a_ < 6; a_++)
      for (int b = 0; b < 6; b++) {
        double sum = 0;
        for (int k = 0; k < 6; k++) sum += X[k][a_] * T[k][b];
        c->IA[pp][a_][b] += sum;
      }
    v6 Zp;
    xform_force_T(c->Rp[i], c->r[i], Za, Zp);
    for (int k = 0; k < 6; k++) Z[pp][k] += Zp[k];
  }
  invert6(c->IA[0], c->IA0inv);
  for (int k = 0; k < 6; k++) {
    double sum = 0;
    for (int j = 0; j < 6; j++) sum += c->IA0inv[k][j] * Z[0][j];
    a[0][k] = -sum;
  }
  /* outward pass */
  for (int i = 0; i < nl; i++) {
    int pp = m->parent[i] + 1;
    xform_motion(c->Rp[i], c->r[i], a[pp], a[i + 1]);
    if (m->joint_type[i] == ORC_JOINT_REVOLUTE) {
      int d = m->dof_of_link[i];
      double qdd = (Y[i] - v6dot(c->h[i], a[i + 1])) * c->Dinv[i];
      acc[6 + d] = qdd;
      for (int k = 0; k < 6; k++) a[i + 1][k] += cor[i][k] + c->S[i][k] * qdd;
    }
  }
  /* base acceleration back to world; classical linear acceleration = spatial + w x v */
  v3 wv, lin;
  v3cross(vel[0], vel[0] + 3, wv);
  for (int k = 0; k < 3; k++) lin[k] = a[0][3 + k] + wv[k];
  m3Tvec(c->R0, a[0], acc);
  m3Tvec(c->R0, lin, acc + 3);
}

/* M^-1 f using the cached articulated inertias (calcAccelerationDeltasMultiDof) */
static void minv(const orc_model* m, const orc_cache* c, const double* f, double* out) {
  int nl = m->n_links;
  static _Thread_local v6 Z[ORC_MAXL + 1], a[ORC_MAXL + 1];
  static _Thread_local double Y[ORC_MAXL];
  v3 t;
  m3vec(c->R0, f, t);
  for (int k = 0; k < 3; k++) Z[0][k] = -t[k];
  m3vec(c->R0, f + 3, t);
  for (int k = 0; k < 3; k++) Z[0][3 + k] = -t[k];
  for (int i = 0; i < nl; i++) memset(Z[i + 1], 0, sizeof(v6));
  for (int i = nl - 1; i >= 0; i--) {
    int pp = m->parent[i] + 1;
    v6 Za, Zp;
    memcpy(Za, Z[i + 1], sizeof(v6));
    if (m->joint_type[i] == ORC_JOINT_REVOLUTE) {
      Y[i] = f[6 + m->dof_of_link[i]] - v6dot(c->S[i], Z[i + 1]);
      for (int k = 0; k < 6; k++) Za[k] += c->h[i][k] * Y[i] * c->Dinv[i];
    }
    xform_force_T(c->Rp[i], c->r[i], Za, Zp);
    for (int k = 0; k < 6; k++) Z[pp][k] += Zp[k];
  }
  for (int k = 0; k < 6; k++) {
    double sum = 0;
    for (int j = 0; j < 6; j++) sum += c->IA0inv[k][j] * Z[0][j];
    a[0][k] = -sum;
  }
  for (int i = 0; i < nl; i++) {
    int pp = m->parent[i] + 1;
    xform_motion(c->Rp[i], c->r[i], a[pp], a[i + 1]);
    if (m->joint_type[i] == ORC_JOINT_REVOLUTE) {
      double qdd = (Y[i] - v6dot(c->h[i], a[i + 1])) * c->Dinv[i];
      out[6 + m->dof_of_link[i]] = qdd;
      for (int k = 0; k < 6; k++) a[i + 1][k] += c->S[i][k] * qdd;
    }
  }
  m3Tvec(c->R0, a[0], out);
  m3Tvec(c->R0, a[0] + 3, out + 3);
}

void orc_default_params(orc_params* p) {
  p->gravity = 9.8;
  p->dt = 1.0 / 240.0;
  p->substeps = 4;
  p->iterations = 5;
  p->erp_contact = 0.9;
  p->erp_joint = 0.2;
  p->linear_slop = 1e-5;
  p->lin_damping = 0.04;
  p->ang_damping = 0.04;
  p->max_coord_vel = 100.0;
  p->warmstart = 0.0;
  p->limit_max_impulse = 100.0;
  p->split_threshold = -0.04;
  p->residual_threshold = 1e-7;
  p->limit_rows_always = 0;
  p->gyro = 1;
  p->has_ground = 1;
  p->ground_friction = 0.8;
  p->self_collision = 1;
  p->persistent_manifold = 0;
}

void orc_forward_dynamics(const orc_model* m, const orc_params* p, const orc_state* s, const double* tau,
                          int with_damping, double* acc) {
  orc_cache* c = (orc_cache*)malloc(sizeof(orc_cache));
  kin(m, s, c);
  aba(m, p, s, tau, with_damping, c, acc);
  free(c);
}

void orc_minv_mult(const orc_model* m, const orc_params* p, const orc_state* s, const double* f, double* out) {
  orc_cache* c = (orc_cache*)malloc(sizeof(orc_cache));
  double tau[ORC_MAXD] = {0}, acc[ORC_MAXU];
  kin(m, s, c);
  aba(m, p, s, tau, 0, c, acc);
  minv(m, c, f, out);
  free(c);
}

/* ------------------------------------------------------------------ 4. world-frame RNEA / mass matrix */
/* Independent of section 3: spatial vectors about the WORLD origin, world axes.
 * Generalised coordinates u = [omega_world, v_baseCOM_world, qd];  tau = M(q) udot + C(q,u) + G(q). */
void orc_rnea(const orc_model* m, const orc_state* s, const double* acc, double gravity, double* tau) {
  int nl = m->n_links;
  orc_cache* c = (orc_cache*)malloc(sizeof(orc_cache));
  kin(m, s, c);
  static _Thread_local v6 V[ORC_MAXL + 1], A[ORC_MAXL + 1], F[ORC_MAXL + 1], Sw[ORC_MAXL];
  v3 t;
  /* base: V = (w, v - w x p), A = (wd, vd - wd x p - w x v) + gravity trick */
  v3cross(s->omega, s->pos, t);
  for (int k = 0; k < 3; k++) { V[0][k] = s->omega[k]; V[0][3 + k] = s->vel[k] - t[k]; }
  v3 t2;
  v3cross(acc, s->pos, t);
  v3cross(s->omega, s->vel, t2);
  for (int k = 0; k < 3; k++) { A[0][k] = acc[k]; A[0][3 + k] = acc[3 + k] - t[k] - t2[k]; }
  A[0][5] += gravity;
  for (int i = 0; i < nl; i++) {
    int pp = m->parent[i] + 1;
    memcpy(V[i + 1], V[pp], sizeof(v6));
    memcpy(A[i + 1], A[pp], sizeof(v6));
    memset(Sw[i], 0, sizeof(v6));
    if (m->joint_type[i] == ORC_JOINT_REVOLUTE) {
      int d = m->dof_of_link[i];
      v3 aw, dw, piv;
      m3Tvec(c->Rw[i + 1], m->axis[i], aw);
      m3Tvec(c->Rw[i + 1], m->d_vec[i], dw);
      for (int k = 0; k < 3; k++) piv[k] = c->pw[i + 1][k] - dw[k];
      v3copy(Sw[i], aw);
      v3cross(piv, aw, &Sw[i][3]);
      v6 vj, cr;
      for (int k = 0; k < 6; k++) { vj[k] = Sw[i][k] * s->qd[d]; V[i + 1][k] += vj[k]; }
      motion_cross(V[i + 1], vj, cr);
      for (int k = 0; k < 6; k++) A[i + 1][k] += Sw[i][k] * acc[6 + d] + cr[k];
    }
  }
  for (int i = 0; i <= nl; i++) {
    double mass = i == 0 ? m->base_mass : m->mass[i - 1];
    const double* Id = i == 0 ? m->base_inertia : m->inertia[i - 1];
    /* world inertia about COM: R^T diag R with R = world->link */
    m3 Iw;
    for (int a_ = 0; a_ < 3; a_++)
      for (int b = 0; b < 3; b++) {
        double sum = 0;
        for (int k = 0; k < 3; k++) sum += c->Rw[i][k][a_] * Id[k] * c->Rw[i][k][b];
        Iw[a_][b] = sum;
      }
    /* momentum about origin: h = (Iw w + m c x vc, m vc), vc = v_O + w x c */
    const double* pc = c->pw[i];
    v3 vc, wc, ac, tmp;
    v3cross(V[i], pc, wc);
    for (int k = 0; k < 3; k++) vc[k] = V[i][3 + k] + wc[k];
    /* classical COM acceleration: a_O + wd x c + w x (w x c) + w x v_O ... use spatial form:
       f = I A + V x* (I V) with I about origin */
    /* I*X for X = (w, vO): lin = m (vO + w x c) ; ang = Iw w + m c x (vO + w x c) */
    v6 IV, IAc;
    m3vec(Iw, V[i], tmp);
    v3 cxv;
    v3cross(pc, vc, cxv);
    for (int k = 0; k < 3; k++) { IV[k] = tmp[k] + mass * cxv[k]; IV[3 + k] = mass * vc[k]; }
    v3cross(A[i], pc, wc);
    for (int k = 0; k < 3; k++) ac[k] = A[i][3 + k] + wc[k];
    m3vec(Iw, A[i], tmp);
    v3cross(pc, ac, cxv);
    for (int k = 0; k < 3; k++) { IAc[k] = tmp[k] + mass * cxv[k]; IAc[3 + k] = mass * ac[k]; }
    /* force cross: V x* F = (w x n + v x f, w x f) */
    v3 c1, c2, c3;
    v3cross(V[i], IV, c1);
    v3cross(V[i] + 3, IV + 3, c2);
    v3cross(V[i], IV + 3, c3);
    for (int k = 0; k < 3; k++) { F[i][k] = IAc[k] + c1[k] + c2[k]; F[i][3 + k] = IAc[3 + k] + c3[k]; }
  }
  for (int i = nl - 1; i >= 0; i--) {
    int pp = m->parent[i] + 1;
    if (m->joint_type[i] == ORC_JOINT_REVOLUTE) tau[6 + m->dof_of_link[i]] = v6dot(Sw[i], F[i + 1]);
    for (int k = 0; k < 6; k++) F[pp][k] += F[i + 1][k];
  }
  v3cross(s->pos, F[0] + 3, t);
  for (int k = 0; k < 3; k++) { tau[k] = F[0][k] - t[k]; tau[3 + k] = F[0][3 + k]; }
  free(c);
}

void orc_mass_matrix(const orc_model* m, const orc_state* s, double* M) {
  int nu = 6 + m->n_dof;
  orc_state z = *s;
  memset(z.omega, 0, sizeof(z.omega));
  memset(z.vel, 0, sizeof(z.vel));
  memset(z.qd, 0, sizeof(z.qd));
  for (int cidx = 0; cidx < nu; cidx++) {
    double acc[ORC_MAXU] = {0}, tau[ORC_MAXU];
    acc[cidx] = 1.0;
    orc_rnea(m, &z, acc, 0.0, tau);
    for (int r = 0; r < nu; r++) M[r * nu + cidx] = tau[r];
  }
}

void orc_energy_momentum(const orc_model* m, const orc_state* s, double gravity, double* out) {
  orc_cache* c = (orc_cache*)malloc(sizeof(orc_cache));
  kin(m, s, c);
  static _Thread_local v6 vel[ORC_MAXL + 1];
  m3vec(c->R0, s->omega, vel[0]);
  m3vec(c->R0, s->vel, vel[0] + 3);
  double KE = 0, PE = 0, P[3] = {0, 0, 0}, L[3] = {0, 0, 0};
  for (int i = 0; i <= m->n_links; i++) {
    if (i > 0) {
      int l = i - 1, pp = m->parent[l] + 1;
      xform_motion(c->Rp[l], c->r[l], vel[pp], vel[i]);
      double qd = m->joint_type[l] == ORC_JOINT_REVOLUTE ? s->qd[m->dof_of_link[l]] : 0.0;
      for (int k = 0; k < 6; k++) vel[i][k] += c->S[l][k] * qd;
    }
    double mass = i == 0 ? m->base_mass : m->mass[i - 1];
    const double* Id = i == 0 ? m->base_inertia : m->inertia[i - 1];
    v3 Iw_l = {Id[0] * vel[i][0], Id[1] * vel[i][1], Id[2] * vel[i][2]}, Iw, vw, t;
    KE += 0.5 * (v3dot(vel[i], Iw_l) + mass * v3dot(vel[i] + 3, vel[i] + 3));
    PE += mass * gravity * c->pw[i][2];
    m3Tvec(c->Rw[i], Iw_l, Iw);
    m3Tvec(c->Rw[i], vel[i] + 3, vw);
    v3cross(c->pw[i], vw, t);
    for (int k = 0; k < 3; k++) { P[k] += mass * vw[k]; L[k] += Iw[k] + mass * t[k]; }
  }
  out[0] = KE; out[1] = PE;
  for (int k = 0; k < 3; k++) { out[2 + k] = P[k]; out[5 + k] = L[k]; }
  free(c);
}

/* ------------------------------------------------------------------ 5. narrow phase */
static void add_point(orc_contacts* c, int pid, int link, int partner, const v3 pa, const v3 n, double dist,
                      double mu, double erp, double cfm) {
  if (c->n >= ORC_MAXP) return;
  int k = c->n++;
  c->point_id[k] = pid; c->link[k] = link; c->partner[k] = partner; c->link_b[k] = -2;
  v3set(c->pos_b[k], 0, 0, 0);
  v3copy(c->pos_a[k], pa); v3copy(c->normal[k], n);
  c->dist[k] = dist; c->friction[k] = mu; c->erp[k] = erp; c->cfm[k] = cfm; c->impulse[k] = 0;
}

/* sphere (centre cw, radius r) vs static box; returns 1 and fills (pa, n, dist) if dist < thresh */
/* sphere vs capped cylinder (Pillar, bullet_objects.py:86-90; pillar.urdf) in the record's local frame */
static int sphere_cyl(const v3 cw, double r, const orc_box* b, double thresh, v3 pa, v3 n, double* dist) {
  v3 d, cl, nl;
  for (int k = 0; k < 3; k++) d[k] = cw[k] - b->center[k];
  for (int k = 0; k < 3; k++) cl[k] = b->R[0][k] * d[0] + b->R[1][k] * d[1] + b->R[2][k] * d[2];
  double rad = b->half[0], h = b->half[2];
  double rho = sqrt(cl[0] * cl[0] + cl[1] * cl[1]);
  double ux = rho > 1e-12 ? cl[0] / rho : 0.0, uy = rho > 1e-12 ? cl[1] / rho : 0.0;
  int in_r = rho <= rad, in_z = fabs(cl[2]) <= h;
  if (!(in_r && in_z)) {
    double qr = rho < rad ? rho : rad, qz = cl[2] > h ? h : (cl[2] < -h ? -h : cl[2]);
    double dr = rho - qr, dz = cl[2] - qz, len = sqrt(dr * dr + dz * dz);
    *dist = len - r;
    if (*dist >= thresh) return 0;
    nl[0] = ux * dr / len; nl[1] = uy * dr / len; nl[2] = dz / len;
  } else {
    double pen_r = rad - rho, pen_z = h - fabs(cl[2]);
    if (pen_z <= pen_r) { nl[0] = 0; nl[1] = 0; nl[2] = cl[2] >= 0 ? 1.0 : -1.0; *dist = -pen_z - r; }
    else { nl[0] = ux; nl[1] = uy; nl[2] = 0; *dist = -pen_r - r; }
  }
  for (int k = 0; k < 3; k++) n[k] = b->R[k][0] * nl[0] + b->R[k][1] * nl[1] + b->R[k][2] * nl[2];
  for (int k = 0; k < 3; k++) pa[k] = cw[k] - r * n[k];
  return 1;
}

static int sphere_box(const v3 cw, double r, const orc_box* b, double thresh, v3 pa, v3 n, double* dist) {
  if (b->cylinder) return sphere_cyl(cw, r, b, thresh, pa, n, dist);
  v3 d, cl, q, nl;
  for (int k = 0; k < 3; k++) d[k] = cw[k] - b->center[k];
  for (int k = 0; k < 3; k++) cl[k] = b->R[0][k] * d[0] + b->R[1][k] * d[1] + b->R[2][k] * d[2];
  int inside = 1;
  for (int k = 0; k < 3; k++) {
    q[k] = cl[k];
    if (q[k] > b->half[k]) { q[k] = b->half[k]; inside = 0; }
    if (q[k] < -b->half[k]) { q[k] = -b->half[k]; inside = 0; }
  }
  if (!inside) {
    v3 diff = {cl[0] - q[0], cl[1] - q[1], cl[2] - q[2]};
    double len = v3norm(diff);
    *dist = len - r;
    if (*dist >= thresh) return 0;
    for (int k = 0; k < 3; k++) nl[k] = diff[k] / len;
  } else {
    int ax = 0;
    double best = 1e30;
    for (int k = 0; k < 3; k++) {
      double pen = b->half[k] - fabs(cl[k]);
      if (pen < best) { best = pen; ax = k; }
    }
    v3set(nl, 0, 0, 0);
    nl[ax] = cl[ax] >= 0 ? 1.0 : -1.0;
    *dist = -best - r;
  }
  for (int k = 0; k < 3; k++) n[k] = b->R[k][0] * nl[0] + b->R[k][1] * nl[1] + b->R[k][2] * nl[2];
  for (int k = 0; k < 3; k++) pa[k] = cw[k] - r * n[k];
  return 1;
}

static void point_world(const orc_model* m, const orc_cache* c, int g, int e, v3 cw) {
  int link = m->geom_link[g];
  const double* pl = e == 0 ? m->geom_p0[g] : m->geom_p1[g];
  v3 t;
  m3Tvec(c->Rw[link + 1], pl, t);
  for (int k = 0; k < 3; k++) cw[k] = c->pw[link + 1][k] + t[k];
}

/* Candidate points: sphere centres and both capsule end spheres.  Contacts are listed obstacle-major:
 * the ground plane for every point, then box 0 for every point, ... (the CUDA kernel uses the same order). */
/* sphere (centre cw, radius r) vs bar-as-capsule */
static int sphere_bar(const v3 cw, double r, const orc_bar* b, double thresh, v3 pa, v3 n, double* dist) {
  v3 d = {cw[0] - b->center[0], cw[1] - b->center[1], cw[2] - b->center[2]};
  double t = v3dot(d, b->axis);
  if (t > b->halflen) t = b->halflen;
  if (t < -b->halflen) t = -b->halflen;
  v3 q = {b->center[0] + t * b->axis[0], b->center[1] + t * b->axis[1], b->center[2] + t * b->axis[2]};
  v3 df = {cw[0] - q[0], cw[1] - q[1], cw[2] - q[2]};
  double len = v3norm(df);
  *dist = len - r - b->radius;
  if (*dist >= thresh || len < 1e-12) return 0;
  for (int k = 0; k < 3; k++) { n[k] = df[k] / len; pa[k] = cw[k] - r * n[k]; }
  return 1;
}

/* robot box geom vs bar.  Bullet runs GJK/EPA (one point per frame); restated as the deepest of 5 spheres of the
 * bar's radius sampled 3 cm apart along its axis around the point nearest to the box centre (documented
 * approximation). */
static int box_bar(const orc_box* bx, const orc_bar* b, double thresh, v3 pa, v3 n, double* dist) {
  v3 d = {bx->center[0] - b->center[0], bx->center[1] - b->center[1], bx->center[2] - b->center[2]};
  double t0 = v3dot(d, b->axis);
  int found = 0;
  double best = 1e30;
  /* visiting order 0, -1, +1, -2, +2; an outer sample replaces the current one only if it is deeper by more than
   * 1e-5 m, so that a bar lying parallel to a box face (all samples equally deep) yields the central point */
  static const int order[5] = {0, -1, 1, -2, 2};
  for (int kk = 0; kk < 5; kk++) {
    int k = order[kk];
    double t = t0 + 0.03 * k;
    if (t > b->halflen) t = b->halflen;
    if (t < -b->halflen) t = -b->halflen;
    v3 q = {b->center[0] + t * b->axis[0], b->center[1] + t * b->axis[1], b->center[2] + t * b->axis[2]};
    v3 ps, ns;
    double ds;
    if (sphere_box(q, b->radius, bx, thresh, ps, ns, &ds) && ds < best - 1e-5) {
      best = ds;
      found = 1;
      for (int i = 0; i < 3; i++) { n[i] = -ns[i]; pa[i] = ps[i] - ds * ns[i]; }
    }
  }
  *dist = best;
  return found;
}

static int collide_bars(const orc_model* m, const orc_params* p, const orc_cache* c, const orc_bar* bars, int n_bars,
                        orc_contacts* out) {
  for (int ob = 0; ob < n_bars; ob++) {
    for (int g = 0; g < m->n_geoms; g++) {
      int link = m->geom_link[g];
      double thresh = m->link_thresh[link + 1];
      if (m->geom_type[g] == ORC_GEOM_BOX) continue;
      int nends = m->geom_type[g] == ORC_GEOM_CAPSULE ? 2 : 1;
      for (int e = 0; e < nends; e++) {
        v3 cw, pa, n;
        double dist;
        point_world(m, c, g, e, cw);
        if (sphere_bar(cw, m->geom_size[g][0], &bars[ob], thresh, pa, n, &dist))
          add_point(out, 2 * g + e, link, bars[ob].id, pa, n, dist, m->geom_friction[g] * bars[ob].friction,
                    p->erp_contact, 0.0);
      }
    }
    for (int g = 0; g < m->n_geoms; g++) {
      if (m->geom_type[g] != ORC_GEOM_BOX) continue;
      int link = m->geom_link[g];
      orc_box bx;
      v3 t;
      m3 Rg, Rl;
      m3Tvec(c->Rw[link + 1], m->geom_p0[g], t);
      for (int k = 0; k < 3; k++) bx.center[k] = c->pw[link + 1][k] + t[k];
      quat_to_mat(m->geom_quat[g], Rg);
      for (int a = 0; a < 3; a++)
        for (int b2 = 0; b2 < 3; b2++) Rl[a][b2] = c->Rw[link + 1][b2][a]; /* link -> world */
      m3mul(Rl, Rg, bx.R);
      for (int k = 0; k < 3; k++) bx.half[k] = m->geom_size[g][k];
      bx.cylinder = 0;
      v3 pa, n;
      double dist;
      if (box_bar(&bx, &bars[ob], m->link_thresh[link + 1], pa, n, &dist))
        add_point(out, 2 * g, link, bars[ob].id, pa, n, dist, m->geom_friction[g] * bars[ob].friction,
                  p->erp_contact, 0.0);
    }
  }
  return out->n;
}

/* closest points of two segments (Ericson, Real-Time Collision Detection 5.1.9); a sphere is a zero-length segment */
static void seg_seg(const v3 p1, const v3 q1, const v3 p2, const v3 q2, v3 c1, v3 c2) {
  v3 d1 = {q1[0] - p1[0], q1[1] - p1[1], q1[2] - p1[2]}, d2 = {q2[0] - p2[0], q2[1] - p2[1], q2[2] - p2[2]};
  v3 r = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
  double a = v3dot(d1, d1), e = v3dot(d2, d2), f = v3dot(d2, r), s, t;
  const double EPS = 1e-12;
  if (a <= EPS && e <= EPS) { s = t = 0; }
  else if (a <= EPS) { s = 0; t = f / e; t = t < 0 ? 0 : (t > 1 ? 1 : t); }
  else {
    double c = v3dot(d1, r);
    if (e <= EPS) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
    else {
      double b = v3dot(d1, d2), den = a * e - b * b;
      s = den > EPS ? (b * f - c * e) / den : 0.0;
      s = s < 0 ? 0 : (s > 1 ? 1 : s);
      t = (b * s + f) / e;
      if (t < 0) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
      else if (t > 1) { t = 1; s = (b - c) / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
    }
  }
  for (int k = 0; k < 3; k++) { c1[k] = p1[k] + d1[k] * s; c2[k] = p2[k] + d2[k] * t; }
}

/* Self-collision (robots.py:259-264).  Bullet: sphere-sphere, capsule-capsule (capsuleCapsuleDistance) and the GJK
 * pairs all yield the closest points of the two core segments; one point per geom pair and substep.  Contact while
 * the distance is below the smaller of the two links' breaking thresholds; combined friction = product. */
static void collide_self(const orc_model* m, const orc_params* p, const orc_cache* c, orc_contacts* out) {
  for (int k = 0; k < m->n_self; k++) {
    int ga = m->self_a[k], gb = m->self_b[k];
    int la = m->geom_link[ga], lb = m->geom_link[gb];
    v3 a0, a1, b0, b1, c1, c2;
    point_world(m, c, ga, 0, a0); point_world(m, c, ga, 1, a1);
    point_world(m, c, gb, 0, b0); point_world(m, c, gb, 1, b1);
    seg_seg(a0, a1, b0, b1, c1, c2);
    v3 d = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
    double len = v3norm(d), ra = m->geom_size[ga][0], rb = m->geom_size[gb][0];
    double dist = len - ra - rb;
    double thresh = m->link_thresh[la + 1] < m->link_thresh[lb + 1] ? m->link_thresh[la + 1] : m->link_thresh[lb + 1];
    if (dist >= thresh || len < 1e-9) continue;
    v3 n = {d[0] / len, d[1] / len, d[2] / len}; /* on B, pointing towards A */
    v3 pa = {c1[0] - ra * n[0], c1[1] - ra * n[1], c1[2] - ra * n[2]};
    if (out->n >= ORC_MAXP) return;
    int idx = out->n;
    add_point(out, 2 * ga, la, 1000 + lb + 1, pa, n, dist, m->geom_friction[ga] * m->geom_friction[gb],
              p->erp_contact, 0.0);
    out->link_b[idx] = lb;
    for (int i = 0; i < 3; i++) out->pos_b[idx][i] = c2[i] + rb * n[i];
  }
}

/* ---- persistent_manifold switch: btPersistentManifold semantics for robot link vs ground plane (oracle only) ----------
 * State per link in man[16 * (link + 1)]: 4 slots x {candidate id + 1, cached plane point x, y, pad}, compact at the front.
 * Per substep, as btCollisionDispatcher / btCompoundCollisionAlgorithm / btConvexPlaneCollisionAlgorithm do it:
 *   refreshContactPoints: a cached point is dropped when its distance exceeds the link's breaking threshold or when it has
 *     drifted along the plane by more than the threshold from where it was cached (removeContactPoint swaps the last in);
 *   every child shape reports ONE point, its support vertex towards the plane (a capsule: the deeper end), if closer than
 *     the threshold; getCacheEntry replaces the cached point nearest to it within the threshold, else it is appended, and
 *     a fifth point replaces the slot sortCachedPoints picks (keep the deepest, maximise calcArea4Points). */
static __thread double* orc_tls_manifold = NULL;

static double area4(const v3 p0, const v3 p1, const v3 p2, const v3 p3) {
  v3 a[3], b[3], t;
  for (int k = 0; k < 3; k++) {
    a[0][k] = p0[k] - p1[k]; b[0][k] = p2[k] - p3[k];
    a[1][k] = p0[k] - p2[k]; b[1][k] = p1[k] - p3[k];
    a[2][k] = p0[k] - p3[k]; b[2][k] = p1[k] - p2[k];
  }
  double best = 0;
  for (int i = 0; i < 3; i++) {
    v3cross(a[i], b[i], t);
    double l2 = v3dot(t, t);
    if (l2 > best) best = l2;
  }
  return best;
}

static void candidate_now(const orc_model* m, const orc_cache* c, int id, v3 pa, double* dist) {
  int g = id >> 1, e = id & 1;
  v3 cw;
  point_world(m, c, g, e, cw);
  double r = m->geom_size[g][0];
  pa[0] = cw[0]; pa[1] = cw[1]; pa[2] = cw[2] - r;
  *dist = cw[2] - r;
}

static void collide_ground_manifold(const orc_model* m, const orc_params* p, const orc_cache* c, double* man,
                                    orc_contacts* out) {
  for (int li = -1; li < m->n_links; li++) {
    double* M = man + 16 * (li + 1);
    double thresh = m->link_thresh[li + 1];
    int n = 0;
    while (n < 4 && M[4 * n] > 0) n++;
    /* refreshContactPoints (back to front, like Bullet) */
    for (int k = n - 1; k >= 0; k--) {
      v3 pa;
      double dist;
      candidate_now(m, c, (int)M[4 * k] - 1, pa, &dist);
      double dx = pa[0] - M[4 * k + 1], dy = pa[1] - M[4 * k + 2];
      if (dist > thresh || dx * dx + dy * dy > thresh * thresh) {
        n--;
        for (int q = 0; q < 4; q++) { M[4 * k + q] = M[4 * n + q]; M[4 * n + q] = 0; }
      }
    }
    /* narrow phase: one support point per child shape */
    for (int g = 0; g < m->n_geoms; g++) {
      if (m->geom_link[g] != li || m->geom_type[g] == ORC_GEOM_BOX) continue;
      int nends = m->geom_type[g] == ORC_GEOM_CAPSULE ? 2 : 1, id = 2 * g;
      v3 pa, pb;
      double dist, d1;
      candidate_now(m, c, 2 * g, pa, &dist);
      if (nends == 2) {
        candidate_now(m, c, 2 * g + 1, pb, &d1);
        if (d1 < dist) { dist = d1; id = 2 * g + 1; v3copy(pa, pb); }
      }
      if (!(dist < thresh)) continue;
      int idx = -1;
      double shortest = thresh * thresh;
      v3 q[4];
      double qd[4];
      for (int k = 0; k < n; k++) {
        candidate_now(m, c, (int)M[4 * k] - 1, q[k], &qd[k]);
        double d2 = (q[k][0] - pa[0]) * (q[k][0] - pa[0]) + (q[k][1] - pa[1]) * (q[k][1] - pa[1]) +
                    (q[k][2] - pa[2]) * (q[k][2] - pa[2]);
        if (d2 < shortest) { shortest = d2; idx = k; }
      }
      if (idx < 0) {
        if (n < 4) idx = n++;
        else { /* sortCachedPoints */
          int deepest = -1;
          double maxpen = dist;
          for (int k = 0; k < 4; k++)
            if (qd[k] < maxpen) { deepest = k; maxpen = qd[k]; }
          double res[4] = {0, 0, 0, 0};
          if (deepest != 0) res[0] = area4(pa, q[1], q[2], q[3]);
          if (deepest != 1) res[1] = area4(pa, q[0], q[2], q[3]);
          if (deepest != 2) res[2] = area4(pa, q[0], q[1], q[3]);
          if (deepest != 3) res[3] = area4(pa, q[0], q[1], q[2]);
          idx = 0;
          for (int k = 1; k < 4; k++)
            if (res[k] > res[idx]) idx = k;
        }
      }
      M[4 * idx] = id + 1; M[4 * idx + 1] = pa[0]; M[4 * idx + 2] = pa[1]; M[4 * idx + 3] = 0;
    }
    /* the manifold's points are this link's contacts */
    for (int k = 0; k < n; k++) {
      int id = (int)M[4 * k] - 1;
      v3 pa, nrm = {0, 0, 1};
      double dist;
      candidate_now(m, c, id, pa, &dist);
      add_point(out, id, li, 0, pa, nrm, dist, m->geom_friction[id >> 1] * p->ground_friction, p->erp_contact, 0.0);
    }
  }
}

int orc_collide_cached(const orc_model* m, const orc_params* p, const orc_cache* c, const orc_box* boxes,
                       int n_boxes, orc_contacts* out) {
  out->n = 0;
  for (int ob = -1; ob < n_boxes; ob++) {
    if (ob < 0 && !p->has_ground) continue;
    if (ob < 0 && orc_tls_manifold) { collide_ground_manifold(m, p, c, orc_tls_manifold, out); continue; }
    for (int g = 0; g < m->n_geoms; g++) {
      int link = m->geom_link[g];
      if (m->geom_type[g] == ORC_GEOM_BOX) continue; /* robot box geoms (Monkey3D) not handled yet */
      int nends = m->geom_type[g] == ORC_GEOM_CAPSULE ? 2 : 1;
      double r = m->geom_size[g][0];
      double thresh = m->link_thresh[link + 1];
      for (int e = 0; e < nends; e++) {
        v3 cw;
        point_world(m, c, g, e, cw);
        if (ob < 0) {
          double dist = cw[2] - r;
          if (dist < thresh) {
            v3 n = {0, 0, 1}, pa = {cw[0], cw[1], cw[2] - r};
            add_point(out, 2 * g + e, link, 0, pa, n, dist, m->geom_friction[g] * p->ground_friction,
                      p->erp_contact, 0.0);
          }
        } else {
          v3 pa, n;
          double dist;
          if (sphere_box(cw, r, &boxes[ob], thresh, pa, n, &dist)) {
            double erp = p->erp_contact, cfm = 0.0;
            if (boxes[ob].stiffness > 0) {
              /* soft contact, SURVEY App. B.4 (bullet_objects.py:64-72) */
              double denom = p->dt * boxes[ob].stiffness + boxes[ob].damping;
              if (denom < 1.1920929e-07) denom = 1.1920929e-07;
              cfm = 1.0 / denom;
              erp = p->dt * boxes[ob].stiffness / denom;
            }
            add_point(out, 2 * g + e, link, boxes[ob].id, pa, n, dist, m->geom_friction[g] * boxes[ob].friction,
                      erp, cfm);
          }
        }
      }
    }
  }
  return out->n;
}

/* ---- mesh-hull self-collision (Cassie): GJK distance between the <= 32-vertex hulls of two links, Bullet's
 * btGjkPairDetector on two btConvexHullShapes with margin m->hull_margin each.  Sub-distance step: every face of the
 * <= 4-point simplex (closest point of its affine hull, valid when its barycentric weights are >= 0; the nearest valid one
 * is the closest point of the simplex).  Cores that overlap fall back to the overlap along the centre line. */
static int gjk_closest(double w[4][3], double a[4][3], double lam[4], int n, v3 v) {
  double best = 1e300, bl[4] = {0, 0, 0, 0};
  int bmask = 0;
  for (int mask = 1; mask < (1 << n); mask++) {
    int id[4], k = 0;
    for (int i = 0; i < n; i++)
      if ((mask >> i) & 1) id[k++] = i;
    const double* p0 = w[id[0]];
    double mu[3] = {0, 0, 0};
    int ok = 1;
    if (k > 1) {
      double e[3][3], G[3][3], b[3];
      for (int i = 0; i < k - 1; i++)
        for (int c = 0; c < 3; c++) e[i][c] = w[id[i + 1]][c] - p0[c];
      for (int i = 0; i < k - 1; i++) {
        b[i] = -v3dot(e[i], p0);
        for (int j = 0; j < k - 1; j++) G[i][j] = v3dot(e[i], e[j]);
      }
      if (k == 2) {
        ok = G[0][0] > 1e-20;
        mu[0] = ok ? b[0] / G[0][0] : 0.0;
      } else if (k == 3) {
        double det = G[0][0] * G[1][1] - G[0][1] * G[1][0];
        ok = det > 1e-5 * G[0][0] * G[1][1];  /* (a float32 determinant is noise below that) */
        if (ok) { mu[0] = (b[0] * G[1][1] - b[1] * G[0][1]) / det; mu[1] = (G[0][0] * b[1] - G[1][0] * b[0]) / det; }
      } else {
        double c00 = G[1][1] * G[2][2] - G[1][2] * G[2][1], c01 = G[1][2] * G[2][0] - G[1][0] * G[2][2];
        double c02 = G[1][0] * G[2][1] - G[1][1] * G[2][0];
        double det = G[0][0] * c00 + G[0][1] * c01 + G[0][2] * c02;
        ok = det > 1e-4 * G[0][0] * G[1][1] * G[2][2];
        if (ok) {
          mu[0] = (b[0] * c00 + b[1] * (G[0][2] * G[2][1] - G[0][1] * G[2][2]) + b[2] * (G[0][1] * G[1][2] - G[0][2] * G[1][1])) / det;
          mu[1] = (b[0] * c01 + b[1] * (G[0][0] * G[2][2] - G[0][2] * G[2][0]) + b[2] * (G[0][2] * G[1][0] - G[0][0] * G[1][2])) / det;
          mu[2] = (b[0] * c02 + b[1] * (G[0][1] * G[2][0] - G[0][0] * G[2][1]) + b[2] * (G[0][0] * G[1][1] - G[0][1] * G[1][0])) / det;
        }
      }
    }
    if (!ok) continue;
    double l4[4] = {1.0 - mu[0] - mu[1] - mu[2], mu[0], mu[1], mu[2]};
    int inside = 1;
    for (int i = 0; i < k; i++)
      if (l4[i] < -1e-6) inside = 0;
    if (!inside) continue;
    v3 pnt = {0, 0, 0};
    if (k == 3) { /* triangle interior: foot of the perpendicular through the plane normal (better conditioned) */
      v3 e0, e1, nn;
      for (int c = 0; c < 3; c++) { e0[c] = w[id[1]][c] - p0[c]; e1[c] = w[id[2]][c] - p0[c]; }
      v3cross(e0, e1, nn);
      double sc = v3dot(nn, p0) / v3dot(nn, nn);
      for (int c = 0; c < 3; c++) pnt[c] = nn[c] * sc;
    } else {
      for (int i = 0; i < k; i++)
        for (int c = 0; c < 3; c++) pnt[c] += l4[i] * w[id[i]][c];
    }
    double d2 = v3dot(pnt, pnt);
    if (d2 < best) {
      best = d2; bmask = mask;
      for (int i = 0; i < 4; i++) bl[i] = i < k ? (l4[i] > 0 ? l4[i] : 0.0) : 0.0;
      v3copy(v, pnt);
    }
  }
  int k = 0;
  for (int i = 0; i < n; i++)
    if ((bmask >> i) & 1) {
      for (int c = 0; c < 3; c++) { w[k][c] = w[i][c]; a[k][c] = a[i][c]; }
      lam[k] = bl[k];
      k++;
    }
  return k;
}

static int hull_support(double V[ORC_HULLV][3], const v3 d) {
  int best = 0;
  double bv = v3dot(V[0], d);
  for (int i = 1; i < ORC_HULLV; i++) {
    double x = v3dot(V[i], d);
    if (x > bv) { bv = x; best = i; }
  }
  return best;
}

static void collide_hulls(const orc_model* m, const orc_params* p, const orc_cache* c, orc_contacts* out) {
  for (int k = 0; k < m->n_hpairs; k++) {
    int ha = m->hpair_a[k], hb = m->hpair_b[k];
    int la = m->hull_link[ha], lb = m->hull_link[hb];
    double thresh = m->link_thresh[la + 1] < m->link_thresh[lb + 1] ? m->link_thresh[la + 1] : m->link_thresh[lb + 1];
    v3 ca, cb, t;
    m3Tvec(c->Rw[la + 1], m->hull_center[ha], t);
    for (int i = 0; i < 3; i++) ca[i] = c->pw[la + 1][i] + t[i];
    m3Tvec(c->Rw[lb + 1], m->hull_center[hb], t);
    for (int i = 0; i < 3; i++) cb[i] = c->pw[lb + 1][i] + t[i];
    double reach = m->hull_radius[ha] + m->hull_radius[hb] + thresh + 2 * m->hull_margin;
    v3 dc = {ca[0] - cb[0], ca[1] - cb[1], ca[2] - cb[2]};
    if (v3dot(dc, dc) >= reach * reach) continue;
    double A[ORC_HULLV][3], B[ORC_HULLV][3];
    for (int i = 0; i < ORC_HULLV; i++) {
      m3Tvec(c->Rw[la + 1], m->hull_verts[ha][i], t);
      for (int j = 0; j < 3; j++) A[i][j] = c->pw[la + 1][j] + t[j];
      m3Tvec(c->Rw[lb + 1], m->hull_verts[hb][i], t);
      for (int j = 0; j < 3; j++) B[i][j] = c->pw[lb + 1][j] + t[j];
    }
    double w[4][3], a[4][3], lam[4];
    v3 v = {A[0][0] - B[0][0], A[0][1] - B[0][1], A[0][2] - B[0][2]};
    int n = 0, overlap = 0;
    for (int it = 0; it < 32; it++) {
      double vv = v3dot(v, v);
      if (vv < 1e-14) { overlap = 1; break; }
      v3 nv = {-v[0], -v[1], -v[2]};
      int sa = hull_support(A, nv), sb = hull_support(B, v);
      v3 ws = {A[sa][0] - B[sb][0], A[sa][1] - B[sb][1], A[sa][2] - B[sb][2]};
      if (vv - v3dot(v, ws) <= 1e-7 * vv && n > 0) break;
      v3copy(w[n], ws); v3copy(a[n], A[sa]);
      n = gjk_closest(w, a, lam, n + 1, v);
      if (n == 4) { overlap = 1; break; }
    }
    v3 nrm, pa;
    double dist;
    if (!overlap) {
      double len = v3norm(v);
      dist = len - 2 * m->hull_margin;
      if (dist >= thresh) continue;
      for (int i = 0; i < 3; i++) {
        nrm[i] = v[i] / len;
        double acc = 0;
        for (int j = 0; j < n; j++) acc += lam[j] * a[j][i];
        pa[i] = acc - m->hull_margin * nrm[i];
      }
    } else {
      double len = v3norm(dc);
      if (len < 1e-9) continue;
      for (int i = 0; i < 3; i++) nrm[i] = dc[i] / len;
      v3 nn = {-nrm[0], -nrm[1], -nrm[2]};
      int sa = hull_support(A, nn), sb = hull_support(B, nrm);
      dist = (A[sa][0] - B[sb][0]) * nrm[0] + (A[sa][1] - B[sb][1]) * nrm[1] + (A[sa][2] - B[sb][2]) * nrm[2] - 2 * m->hull_margin;
      for (int i = 0; i < 3; i++) pa[i] = A[sa][i] - m->hull_margin * nrm[i];
    }
    if (out->n >= ORC_MAXP) return;
    int idx = out->n;
    add_point(out, 320 + k, la, 1000 + lb + 1, pa, nrm, dist, m->hull_friction[ha] * m->hull_friction[hb], p->erp_contact, 0.0);
    out->link_b[idx] = lb;
    for (int i = 0; i < 3; i++) out->pos_b[idx][i] = pa[i] - dist * nrm[i];
  }
}

int orc_collide(const orc_model* m, const orc_params* p, const orc_state* s, const orc_box* boxes, int n_boxes,
                orc_contacts* out) {
  orc_cache* c = (orc_cache*)malloc(sizeof(orc_cache));
  kin(m, s, c);
  orc_collide_cached(m, p, c, boxes, n_boxes, out);
  if (p->self_collision) { collide_self(m, p, c, out); collide_hulls(m, p, c, out); }
  free(c);
  return out->n;
}

/* ------------------------------------------------------------------ 6. rows + PGS */
typedef struct {
  double J[ORC_MAXU];
  double MinvJ[ORC_MAXU];
  double rhs, cfm, lo, hi, jinv, applied, mu;
  int normal_index; /* friction rows: index of their normal row */
} orc_row;

/* btPlaneSpace1 */
static void plane_space(const v3 n, v3 p, v3 q) {
  if (fabs(n[2]) > 0.7071067811865475244008443621048490) {
    double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
    p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
    q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
  } else {
    double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
    p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
    q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
  }
}

static void point_jacobian(const orc_model* m, const orc_state* s, const orc_cache* c, int link, const v3 pt,
                           const v3 dir, double* J) {
  int nu = 6 + m->n_dof;
  for (int k = 0; k < nu; k++) J[k] = 0;
  v3 rel, t;
  for (int k = 0; k < 3; k++) rel[k] = pt[k] - s->pos[k];
  v3cross(rel, dir, t);
  for (int k = 0; k < 3; k++) { J[k] = t[k]; J[3 + k] = dir[k]; }
  for (int l = link; l >= 0; l = m->parent[l]) {
    if (m->joint_type[l] != ORC_JOINT_REVOLUTE) continue;
    v3 aw, dw, piv, arm, vel;
    m3Tvec(c->Rw[l + 1], m->axis[l], aw);
    m3Tvec(c->Rw[l + 1], m->d_vec[l], dw);
    for (int k = 0; k < 3; k++) { piv[k] = c->pw[l + 1][k] - dw[k]; arm[k] = pt[k] - piv[k]; }
    v3cross(aw, arm, vel);
    J[6 + m->dof_of_link[l]] = v3dot(dir, vel);
  }
}

static double dotn(const double* a, const double* b, int n) {
  double s = 0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}

/* one contact/friction row (btMultiBodyConstraintSolver::setupMultiBodyContactConstraint) */
static void setup_contact_row(const orc_model* m, const orc_params* p, const orc_state* s, const orc_cache* c,
                              const double* u, const orc_contacts* ct, int k, const v3 dir, int is_friction,
                              orc_row* row) {
  int nu = 6 + m->n_dof;
  double inv_dt = 1.0 / p->dt;
  double cfm = is_friction ? 0.0 : ct->cfm[k] * inv_dt;
  double erp = ct->erp[k];
  point_jacobian(m, s, c, ct->link[k], ct->pos_a[k], dir, row->J);
  minv(m, c, row->J, row->MinvJ);
  double d = dotn(row->J, row->MinvJ, nu) + cfm;
  if (ct->link_b[k] > -2) {
    /* both bodies are links of the multibody: setupMultiBodyContactConstraint fills jacobian B with -dir and sums
     * the two denominators (no coupling term) */
    double JB[ORC_MAXU], MB[ORC_MAXU];
    v3 neg = {-dir[0], -dir[1], -dir[2]};
    point_jacobian(m, s, c, ct->link_b[k], ct->pos_b[k], neg, JB);
    minv(m, c, JB, MB);
    d += dotn(JB, MB, nu);
    for (int i = 0; i < nu; i++) { row->J[i] += JB[i]; row->MinvJ[i] += MB[i]; }
  }
  row->jinv = d > 1.1920929e-07 ? 1.0 / d : 0.0;
  double rel_vel = dotn(row->J, u, nu);
  double distance = is_friction ? 0.0 : ct->dist[k] + p->linear_slop;
  double positional = 0.0, velocity_err = -rel_vel; /* combined restitution = 0 (robot links) */
  if (is_friction) {
    positional = 0.0;
  } else if (distance > 0) {
    velocity_err -= distance * inv_dt;
  } else {
    positional = -distance * erp * inv_dt;
  }
  row->rhs = positional * row->jinv + velocity_err * row->jinv;
  row->cfm = cfm * row->jinv;
  row->applied = 0.0;
  row->mu = ct->friction[k];
  if (is_friction) { row->lo = -row->mu; row->hi = row->mu; }
  else { row->lo = 0.0; row->hi = 1e10; }
}

typedef struct {
  orc_row limit[2 * ORC_MAXD];
  orc_row normal[ORC_MAXP];
  orc_row fric[2 * ORC_MAXP];
} orc_rows;

static void apply_row(double* dv, const orc_row* r, double imp, int nu) {
  for (int i = 0; i < nu; i++) dv[i] += r->MinvJ[i] * imp;
}

static double resolve_single(orc_row* r, double* dv, int nu) {
  double delta = r->rhs - r->applied * r->cfm;
  delta -= dotn(r->J, dv, nu) * r->jinv;
  double sum = r->applied + delta;
  if (sum < r->lo) { delta = r->lo - r->applied; r->applied = r->lo; }
  else if (sum > r->hi) { delta = r->hi - r->applied; r->applied = r->hi; }
  else r->applied = sum;
  apply_row(dv, r, delta, nu);
  return r->jinv != 0.0 ? delta / r->jinv : 0.0;
}

/* btMultiBodyConstraintSolver::resolveConeFrictionConstraintRows */
static double resolve_cone(orc_row* a, orc_row* b, double* dv, int nu) {
  double dB = b->rhs - b->applied * b->cfm - dotn(b->J, dv, nu) * b->jinv;
  double sumB = b->applied + dB;
  double dA = a->rhs - a->applied * a->cfm - dotn(a->J, dv, nu) * a->jinv;
  double sumA = a->applied + dA;
  if (sumA * sumA + sumB * sumB >= a->lo * b->lo) {
    double angle = atan2(sumA, sumB);
    double clipA = fabs(a->lo * sin(angle)), clipB = fabs(b->lo * cos(angle));
    if (sumA < -clipA) { dA = -clipA - a->applied; a->applied = -clipA; }
    else if (sumA > clipA) { dA = clipA - a->applied; a->applied = clipA; }
    else a->applied = sumA;
    if (sumB < -clipB) { dB = -clipB - b->applied; b->applied = -clipB; }
    else if (sumB > clipB) { dB = clipB - b->applied; b->applied = clipB; }
    else b->applied = sumB;
  } else {
    a->applied = sumA;
    b->applied = sumB;
  }
  apply_row(dv, a, dA, nu);
  apply_row(dv, b, dB, nu);
  double res = 0;
  if (a->jinv != 0.0) res += dA / a->jinv;
  if (b->jinv != 0.0) res += dB / b->jinv;
  return res;
}

static void clamp_u(double* u, int nu, double vmax) {
  for (int i = 0; i < nu; i++) {
    if (u[i] > vmax) u[i] = vmax;
    if (u[i] < -vmax) u[i] = -vmax;
  }
}

static void pack_u(const orc_model* m, const orc_state* s, double* u) {
  for (int k = 0; k < 3; k++) { u[k] = s->omega[k]; u[3 + k] = s->vel[k]; }
  for (int d = 0; d < m->n_dof; d++) u[6 + d] = s->qd[d];
}
static void unpack_u(const orc_model* m, orc_state* s, const double* u) {
  for (int k = 0; k < 3; k++) { s->omega[k] = u[k]; s->vel[k] = u[3 + k]; }
  for (int d = 0; d < m->n_dof; d++) s->qd[d] = u[6 + d];
}

/* ------------------------------------------------------------------ 7. integration + driver */
static void integrate_positions(const orc_model* m, const orc_params* p, orc_state* s) {
  double dt = p->dt;
  for (int k = 0; k < 3; k++) s->pos[k] += dt * s->vel[k];
  /* btMultiBody::stepPositionsMultiDof quaternion exponential (world-frame angular velocity) */
  double ang = v3norm(s->omega);
  if (ang * dt > 0.25 * PI) ang = 0.25 * PI / dt;
  double f;
  if (ang < 0.001) f = 0.5 * dt - dt * dt * dt * 0.020833333333 * ang * ang;
  else f = sin(0.5 * ang * dt) / ang;
  double ax = s->omega[0] * f, ay = s->omega[1] * f, az = s->omega[2] * f, aw = cos(ang * dt * 0.5);
  double x = s->quat[0], y = s->quat[1], z = s->quat[2], w = s->quat[3];
  /* q_new = dq * q */
  double nx = aw * x + ax * w + ay * z - az * y;
  double ny = aw * y - ax * z + ay * w + az * x;
  double nz = aw * z + ax * y - ay * x + az * w;
  double nw = aw * w - ax * x - ay * y - az * z;
  double n = sqrt(nx * nx + ny * ny + nz * nz + nw * nw);
  s->quat[0] = nx / n; s->quat[1] = ny / n; s->quat[2] = nz / n; s->quat[3] = nw / n;
  for (int d = 0; d < m->n_dof; d++) s->q[d] += dt * s->qd[d];
}

static void substep_impl(const orc_model* m, const orc_params* p, orc_state* s, const double* tau,
                         const orc_box* boxes, int n_boxes, const orc_bar* bars, int n_bars, double* warm,
                         orc_contacts* ct, int* out_rows);

/* test diagnostics: contact points summed over every substep taken by this thread since the last take (the device
 * keeps the same sum in its record, ER_CONTACTS) */
static __thread long orc_diag_contacts = 0;
/* ... and the joint angles at the start of (up to the last 64 of) those substeps, so that a test can evaluate the
 * conditioning of M(q) along the step the oracle actually took */
static __thread double orc_diag_q[64][ORC_MAXD];
static __thread int orc_diag_nq = 0;
long orc_diag_contacts_take(void) { long v = orc_diag_contacts; orc_diag_contacts = 0; return v; }
int orc_diag_q_take(double* out /* [64][ORC_MAXD] */) {
  int n = orc_diag_nq < 64 ? orc_diag_nq : 64;
  for (int i = 0; i < n; i++)
    for (int d = 0; d < ORC_MAXD; d++) out[i * ORC_MAXD + d] = orc_diag_q[i][d];
  orc_diag_nq = 0;
  return n;
}

void orc_substep(const orc_model* m, const orc_params* p, orc_state* s, const double* tau, const orc_box* boxes,
                 int n_boxes, double* warm, orc_contacts* ct, int* out_rows) {
  substep_impl(m, p, s, tau, boxes, n_boxes, NULL, 0, warm, ct, out_rows);
}

static void substep_impl(const orc_model* m, const orc_params* p, orc_state* s, const double* tau,
                         const orc_box* boxes, int n_boxes, const orc_bar* bars, int n_bars, double* warm,
                         orc_contacts* ct, int* out_rows) {
  int nu = 6 + m->n_dof;
  /* per-thread scratch, allocated once (the CPU baseline steps millions of substeps: no malloc in the loop) */
  static __thread orc_cache* c = NULL;
  static __thread orc_rows* rows = NULL;
  if (!c) { c = (orc_cache*)malloc(sizeof(orc_cache)); rows = (orc_rows*)malloc(sizeof(orc_rows)); }
  double u[ORC_MAXU], acc[ORC_MAXU], dv[ORC_MAXU];
  /* collision detection at start-of-substep poses */
  kin(m, s, c);
  orc_tls_manifold = (p->persistent_manifold && warm) ? warm + ORC_MAXW : NULL;
  orc_collide_cached(m, p, c, boxes, n_boxes, ct);
  orc_tls_manifold = NULL;
  if (n_bars > 0) collide_bars(m, p, c, bars, n_bars, ct);
  if (p->self_collision) { collide_self(m, p, c, ct); collide_hulls(m, p, c, ct); }
  orc_diag_contacts += ct->n;
  for (int d = 0; d < m->n_dof; d++) orc_diag_q[orc_diag_nq & 63][d] = s->q[d];
  orc_diag_nq++;
  /* forward dynamics, velocity update */
  aba(m, p, s, tau, 1, c, acc);
  pack_u(m, s, u);
  for (int i = 0; i < nu; i++) u[i] += acc[i] * p->dt;
  clamp_u(u, nu, p->max_coord_vel);
  for (int i = 0; i < nu; i++) dv[i] = 0;
  /* non-contact rows: joint limits (btMultiBodyJointLimitConstraint::createConstraintRows) */
  int nlim = 0;
  for (int d = 0; d < m->n_dof; d++) {
    if (m->lower[d] > m->upper[d]) continue;
    for (int r = 0; r < 2; r++) {
      double pen = r == 0 ? s->q[d] - m->lower[d] : m->upper[d] - s->q[d];
      if (pen > 0 && !p->limit_rows_always) continue;
      orc_row* row = &rows->limit[nlim++];
      double dir = r ? -1.0 : 1.0;
      for (int k = 0; k < nu; k++) row->J[k] = 0;
      row->J[6 + d] = dir;
      minv(m, c, row->J, row->MinvJ);
      double dd = dotn(row->J, row->MinvJ, nu);
      row->jinv = dd > 1.1920929e-07 ? 1.0 / dd : 0.0;
      double rel_vel = dotn(row->J, u, nu);
      double positional = 0, velocity_err = -rel_vel;
      double erp = pen > p->split_threshold ? p->erp_joint : p->erp_contact;
      int split = !(pen > p->split_threshold); /* position part goes to the (unapplied) split impulse */
      if (pen > 0) velocity_err = -pen / p->dt;
      else positional = -pen * erp / p->dt;
      row->rhs = velocity_err * row->jinv + (split ? 0.0 : positional * row->jinv);
      row->cfm = 0; row->lo = 0; row->hi = p->limit_max_impulse; row->applied = 0; row->mu = 0;
    }
  }
  /* loop closures: btMultiBodyPoint2Point::createConstraintRows -- three world-axis rows per constraint, appended to
   * the non-contact list after the limit rows (PyBullet creates them after the URDF's limit constraints).  The
   * denominator is JA M^-1 JA^T + JB M^-1 JB^T: fillMultiBodyConstraint sums the two bodies' terms and ignores their
   * coupling even when both links belong to the same multibody; erp = m_erp. */
  for (int cidx = 0; cidx < m->n_p2p; cidx++) {
    int la = m->p2p_link_a[cidx], lb = m->p2p_link_b[cidx];
    v3 pa, pb, t;
    m3Tvec(c->Rw[la + 1], m->p2p_pivot_a[cidx], t);
    for (int k = 0; k < 3; k++) pa[k] = c->pw[la + 1][k] + t[k];
    m3Tvec(c->Rw[lb + 1], m->p2p_pivot_b[cidx], t);
    for (int k = 0; k < 3; k++) pb[k] = c->pw[lb + 1][k] + t[k];
    for (int ax = 0; ax < 3; ax++) {
      orc_row* row = &rows->limit[nlim++];
      v3 nrm = {0, 0, 0}, neg = {0, 0, 0};
      nrm[ax] = -1.0; neg[ax] = 1.0;
      double JA[ORC_MAXU], JB[ORC_MAXU], MA[ORC_MAXU], MB[ORC_MAXU];
      point_jacobian(m, s, c, la, pa, nrm, JA);
      point_jacobian(m, s, c, lb, pb, neg, JB);
      minv(m, c, JA, MA);
      minv(m, c, JB, MB);
      double dd = dotn(JA, MA, nu) + dotn(JB, MB, nu);
      for (int k = 0; k < nu; k++) { row->J[k] = JA[k] + JB[k]; row->MinvJ[k] = MA[k] + MB[k]; }
      row->jinv = dd > 1.1920929e-07 ? 1.0 / dd : 0.0;
      double rel_vel = dotn(row->J, u, nu);
      double pos_error = (pa[0] - pb[0]) * nrm[0] + (pa[1] - pb[1]) * nrm[1] + (pa[2] - pb[2]) * nrm[2];
      double positional = -pos_error * p->erp_joint / p->dt;
      row->rhs = (positional - rel_vel) * row->jinv;
      row->cfm = 0; row->lo = -m->p2p_max_impulse[cidx]; row->hi = m->p2p_max_impulse[cidx];
      row->applied = 0; row->mu = 0;
    }
  }
  /* contact rows */
  int nc = ct->n;
  for (int k = 0; k < nc; k++) {
    v3 t1, t2;
    setup_contact_row(m, p, s, c, u, ct, k, ct->normal[k], 0, &rows->normal[k]);
    if (p->warmstart > 0 && warm) {
      double imp = warm[ct->point_id[k]] * p->warmstart;
      rows->normal[k].applied = imp;
      if (imp != 0.0) apply_row(dv, &rows->normal[k], imp, nu);
    }
    plane_space(ct->normal[k], t1, t2);
    setup_contact_row(m, p, s, c, u, ct, k, t1, 1, &rows->fric[2 * k]);
    setup_contact_row(m, p, s, c, u, ct, k, t2, 1, &rows->fric[2 * k + 1]);
    rows->fric[2 * k].normal_index = rows->fric[2 * k + 1].normal_index = k;
  }
  /* PGS (btMultiBodyConstraintSolver::solveSingleIteration order) */
  for (int it = 0; it < p->iterations; it++) {
    double res2 = 0;
    for (int j = 0; j < nlim; j++) {
      int idx = (it & 1) ? j : nlim - 1 - j;
      double r = resolve_single(&rows->limit[idx], dv, nu);
      if (r * r > res2) res2 = r * r;
    }
    for (int k = 0; k < nc; k++) {
      double r = resolve_single(&rows->normal[k], dv, nu);
      if (r * r > res2) res2 = r * r;
    }
    for (int k = 0; k < nc; k++) {
      orc_row *a = &rows->fric[2 * k], *b = &rows->fric[2 * k + 1];
      double total = rows->normal[k].applied;
      a->lo = -(a->mu * total); a->hi = a->mu * total;
      b->lo = -(b->mu * total); b->hi = b->mu * total;
      double r = resolve_cone(a, b, dv, nu);
      if (r * r > res2) res2 = r * r;
    }
    if (res2 <= p->residual_threshold) break;
  }
  for (int i = 0; i < nu; i++) u[i] += dv[i];
  clamp_u(u, nu, p->max_coord_vel);
  unpack_u(m, s, u);
  if (warm) {
    for (int i = 0; i < ORC_MAXW; i++) warm[i] = 0;
    for (int k = 0; k < nc; k++) warm[ct->point_id[k]] = rows->normal[k].applied;
  }
  for (int k = 0; k < nc; k++) ct->impulse[k] = rows->normal[k].applied;
  if (out_rows) *out_rows = nlim + 3 * nc;
  integrate_positions(m, p, s);
}

/* one pybullet.stepSimulation(): applied torques + PyBullet's joint damping torque are computed once and
 * held over the substeps */
void orc_step_physics(const orc_model* m, const orc_params* p, orc_state* s, const double* tau_applied,
                      const orc_box* boxes, int n_boxes, double* warm, orc_contacts* last, int* rows_sum) {
  double tau[ORC_MAXD];
  for (int d = 0; d < m->n_dof; d++) tau[d] = tau_applied[d] - m->damping[d] * s->qd[d];
  orc_contacts* ct = last ? last : (orc_contacts*)malloc(sizeof(orc_contacts));
  int total = 0;
  for (int k = 0; k < p->substeps; k++) {
    int r = 0;
    orc_substep(m, p, s, tau, boxes, n_boxes, warm, ct, &r);
    total += r;
  }
  if (rows_sum) *rows_sum = total;
  if (!last) free(ct);
}

void orc_step_physics_bars(const orc_model* m, const orc_params* p, orc_state* s, const double* tau_applied,
                           const orc_bar* bars, int n_bars, orc_contacts* last, int* rows_sum) {
  double tau[ORC_MAXD];
  for (int d = 0; d < m->n_dof; d++) tau[d] = tau_applied[d] - m->damping[d] * s->qd[d];
  int total = 0;
  for (int k = 0; k < p->substeps; k++) {
    int r = 0;
    substep_impl(m, p, s, tau, NULL, 0, bars, n_bars, NULL, last, &r);
    total += r;
  }
  if (rows_sum) *rows_sum = total;
}

/* ------------------------------------------------------------------ 8. MT19937 */
static void mt_init_genrand(orc_rng* r, uint32_t s) {
  r->mt[0] = s;
  for (int i = 1; i < 624; i++) r->mt[i] = 1812433253U * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
  r->pos = 624;
}
void orc_rng_seed_array(orc_rng* r, const uint32_t* key, int len) {
  mt_init_genrand(r, 19650218U);
  int i = 1, j = 0, k = 624 > len ? 624 : len;
  for (; k; k--) {
    r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1664525U)) + key[j] + (uint32_t)j;
    i++; j++;
    if (i >= 624) { r->mt[0] = r->mt[623]; i = 1; }
    if (j >= len) j = 0;
  }
  for (k = 623; k; k--) {
    r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1566083941U)) - (uint32_t)i;
    i++;
    if (i >= 624) { r->mt[0] = r->mt[623]; i = 1; }
  }
  r->mt[0] = 0x80000000U;
  r->pos = 624;
}
uint32_t orc_rng_u32(orc_rng* r) {
  if (r->pos >= 624) {
    uint32_t* mt = r->mt;
    int kk;
    for (kk = 0; kk < 624 - 397; kk++) {
      uint32_t y = (mt[kk] & 0x80000000U) | (mt[kk + 1] & 0x7fffffffU);
      mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    }
    for (; kk < 623; kk++) {
      uint32_t y = (mt[kk] & 0x80000000U) | (mt[kk + 1] & 0x7fffffffU);
      mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    }
    uint32_t y = (mt[623] & 0x80000000U) | (mt[0] & 0x7fffffffU);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    r->pos = 0;
  }
  uint32_t y = r->mt[r->pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680U;
  y ^= (y << 15) & 0xefc60000U;
  y ^= (y >> 18);
  return y;
}
double orc_rng_double(orc_rng* r) { /* RandomState.random_sample / rand() */
  uint32_t a = orc_rng_u32(r) >> 5, b = orc_rng_u32(r) >> 6;
  return (a * 67108864.0 + b) / 9007199254740992.0;
}
double orc_rng_uniform(orc_rng* r, double lo, double hi) { return lo + (hi - lo) * orc_rng_double(r); }

/* ------------------------------------------------------------------ 9. Walker3DCustomEnv */
/* pybullet.getEulerFromQuaternion */
static void euler_from_quat(const double qin[4], double rpy[3]) {
  double len = sqrt(qin[0] * qin[0] + qin[1] * qin[1] + qin[2] * qin[2] + qin[3] * qin[3]);
  double q[4] = {qin[0] / len, qin[1] / len, qin[2] / len, qin[3] / len};
  double sqx = q[0] * q[0], sqy = q[1] * q[1], sqz = q[2] * q[2], squ = q[3] * q[3];
  double sarg = -2 * (q[0] * q[2] - q[3] * q[1]);
  if (sarg <= -0.99999) {
    rpy[0] = 0; rpy[1] = -0.5 * PI; rpy[2] = 2 * atan2(q[0], -q[1]);
  } else if (sarg >= 0.99999) {
    rpy[0] = 0; rpy[1] = 0.5 * PI; rpy[2] = 2 * atan2(-q[0], q[1]);
  } else {
    rpy[0] = atan2(2 * (q[1] * q[2] + q[3] * q[0]), squ - sqx - sqy + sqz);
    rpy[1] = asin(sarg);
    rpy[2] = atan2(2 * (q[0] * q[1] + q[3] * q[2]), squ + sqx - sqy - sqz);
  }
}

static orc_rng* robot_rng(orc_w3d_env* e) { return e->rng_aliased ? &e->env_rng : &e->robot_rng; }

void orc_w3d_seed(orc_w3d_env* e, const uint32_t* key, int len, int at_construction) {
  /* env_base.py:164-166: seed() rebinds only the env's RandomState; the robot keeps the object it was
   * given at construction (env_base.py:93) -- quirk Q1 */
  if (!at_construction && e->rng_aliased) {
    e->robot_rng = e->env_rng;
    e->rng_aliased = 0;
  }
  orc_rng_seed_array(&e->env_rng, key, len);
  if (at_construction) e->rng_aliased = 1;
}

static double f32(double x) { return (double)(float)x; }

/* robots.py:42-95 */
static void w3d_calc_state(const orc_model* m, orc_w3d_env* e, const orc_contacts* ground_contacts) {
  int A = m->n_dof, F = m->n_feet;
  orc_cache* c = (orc_cache*)malloc(sizeof(orc_cache));
  kin(m, &e->s, c);
  double* st = e->robot_state;
  e->joints_at_limit = 0;
  for (int d = 0; d < A; d++) {
    float q = (float)e->s.q[d], qd = (float)e->s.qd[d];
    float bias = (float)m->lower[d], weight = (float)(m->upper[d] - m->lower[d]);
    float nrm = 2.0f * (q - bias) / weight - 1.0f;
    float sp = 0.1f * qd;
    st[6 + d] = nrm;
    st[6 + A + d] = sp;
    e->joint_speeds[d] = sp;
    if (fabsf(nrm) > 0.99f) e->joints_at_limit++;
  }
  v3copy(e->body_xyz, e->s.pos);
  euler_from_quat(e->s.quat, e->body_rpy);
  double yaw = e->body_rpy[2];
  double cy = cos(-yaw), sy = sin(-yaw);
  e->body_vel[0] = cy * e->s.vel[0] - sy * e->s.vel[1];
  e->body_vel[1] = sy * e->s.vel[0] + cy * e->s.vel[1];
  e->body_vel[2] = e->s.vel[2];
  double minz = 1e30;
  for (int f = 0; f < F; f++) {
    v3copy(e->feet_xyz[f], c->pw[m->foot_link[f] + 1]);
    if (e->feet_xyz[f][2] < minz) minz = e->feet_xyz[f][2];
  }
  if (ground_contacts) {
    for (int f = 0; f < F; f++) {
      e->feet_contact[f] = 0;
      for (int k = 0; k < ground_contacts->n; k++)
        if (ground_contacts->link[k] == m->foot_link[f] && ground_contacts->partner[k] == 0) e->feet_contact[f] = 1;
    }
  }
  st[0] = f32(e->body_xyz[2] - minz);
  st[1] = f32(e->body_vel[0]); st[2] = f32(e->body_vel[1]); st[3] = f32(e->body_vel[2]);
  st[4] = f32(e->body_rpy[0]); st[5] = f32(e->body_rpy[1]);
  for (int f = 0; f < F; f++) st[6 + 2 * A + f] = e->feet_contact[f];
  for (int k = 0; k < 6 + 2 * A + F; k++) {
    if (st[k] > 5) st[k] = 5;
    if (st[k] < -5) st[k] = -5;
  }
  free(c);
}

/* robots.py:179-210 */
static void robot_reset_ex(const orc_model* m, orc_w3d_env* e, const double* pos, const double* vel, int random_pose);
static void w3d_robot_reset(const orc_model* m, orc_w3d_env* e, const double* pos) {
  robot_reset_ex(m, e, pos, NULL, 1);
}
static void robot_reset_ex(const orc_model* m, orc_w3d_env* e, const double* pos, const double* vel, int random_pose) {
  int A = m->n_dof;
  double ang[ORC_MAXD];
  for (int d = 0; d < A; d++) ang[d] = m->base_joint_angles[d];
  orc_rng* rr = robot_rng(e);
  if (orc_rng_double(rr) < 0.5) {
    e->mirrored = 1;
    double tmp[ORC_MAXD];
    memcpy(tmp, ang, sizeof(tmp));
    for (int k = 0; k < m->n_right; k++) {
      ang[m->right_idx[k]] = tmp[m->left_idx[k]];
      ang[m->left_idx[k]] = tmp[m->right_idx[k]];
    }
    for (int k = 0; k < m->n_neg; k++) ang[m->neg_idx[k]] *= -1;
  } else {
    e->mirrored = 0;
  }
  /* random_pose=True: +-0.1 rad noise, clipped to +-0.95 of the normalised range (robots.py:190-194) */
  if (random_pose) {
    double ds[ORC_MAXD];
    for (int d = 0; d < A; d++) ds[d] = orc_rng_uniform(rr, -0.1, 0.1);
    for (int d = 0; d < A; d++) {
      double bias = (double)(float)m->lower[d], weight = (double)(float)(m->upper[d] - m->lower[d]);
      double ps = 2 * (ang[d] + ds[d] - bias) / weight - 1;
      if (ps > 0.95) ps = 0.95;
      if (ps < -0.95) ps = -0.95;
      ang[d] = weight * (ps + 1) / 2 + bias;
    }
  }
  for (int d = 0; d < A; d++) { e->s.q[d] = ang[d]; e->s.qd[d] = 0; }
  for (int k = 0; k < 3; k++) { e->s.pos[k] = pos[k]; e->s.omega[k] = 0; e->s.vel[k] = vel ? vel[k] : 0; }
  /* "quat = quat or self.base_orientation" (robots.py:199): the un-mirrored attribute, not the mirrored copy */
  for (int k = 0; k < 4; k++) e->s.quat[k] = m->base_orientation[k];
  for (int f = 0; f < 4; f++) { e->feet_contact[f] = 0; v3set(e->feet_xyz[f], 0, 0, 0); }
  for (int i = 0; i < ORC_WARMSZ; i++) e->warm[i] = 0;
  w3d_calc_state(m, e, NULL);
}

static void w3d_randomize_target(orc_w3d_env* e) { /* env_locomotion.py:67-74 */
  if (e->eval_mode) { e->dist = 4; e->angle = 0; }
  else {
    e->dist = orc_rng_uniform(&e->env_rng, 3, 5);
    e->angle = orc_rng_uniform(&e->env_rng, -PI / 2, PI / 2);
  }
  e->stop_frames = (orc_rng_u32(&e->env_rng) & 1U) ? 60.0 : 30.0;
}

static void w3d_calc_potential(orc_w3d_env* e, double scene_dt) { /* env_locomotion.py:143-158 */
  double dx = e->walk_target[0] - e->body_xyz[0], dy = e->walk_target[1] - e->body_xyz[1];
  e->angle_to_target = atan2(dy, dx) - e->body_rpy[2];
  e->distance_to_target = sqrt(dx * dx + dy * dy);
  e->linear_potential = -e->distance_to_target / scene_dt;
  e->angular_potential = cos(e->angle_to_target);
}

static void w3d_obs(const orc_model* m, const orc_w3d_env* e, double* obs) {
  int n = 6 + 2 * m->n_dof + m->n_feet;
  for (int k = 0; k < n; k++) obs[k] = e->robot_state[k];
  double s_ = e->distance_to_target * sin(e->angle_to_target);
  double c_ = e->distance_to_target * cos(e->angle_to_target);
  obs[n] = s_ / (1 + fabs(s_));
  obs[n + 1] = c_ / (1 + fabs(c_));
}

void orc_w3d_reset(const orc_model* m, const orc_params* p, orc_w3d_env* e, double* obs) {
  e->done = 0;
  e->elapsed = 0;
  w3d_randomize_target(e);
  e->walk_target[0] = e->dist * cos(e->angle);
  e->walk_target[1] = e->dist * sin(e->angle);
  e->walk_target[2] = 1.0;
  e->close_count = 0;
  w3d_robot_reset(m, e, m->base_position);
  w3d_calc_potential(e, p->dt * p->substeps);
  w3d_obs(m, e, obs);
  if (m->planar_env != 0) { /* Walker2DCustomEnv.reset (env_locomotion.py:289-299): (robot_state, [0], [0]) */
    int n = 6 + 2 * m->n_dof + m->n_feet;
    obs[n] = 0;
    obs[n + 1] = 0;
  }
}

void orc_w3d_step(const orc_model* m, const orc_params* p, orc_w3d_env* e, const double* action, double* obs,
                  double* reward, int* done, int* truncated) {
  int A = m->n_dof;
  double tau[ORC_MAXD];
  /* robots.py:31-40 apply_action (applied_gain = 1 for Walker3DCustomEnv) */
  for (int d = 0; d < A; d++) {
    double a = action[d];
    if (a > 1) a = 1;
    if (a < -1) a = -1;
    tau[d] = m->gain[d] * a;
  }
  int rows = 0;
  orc_step_physics(m, p, &e->s, tau, NULL, 0, e->warm, &e->last_contacts, &rows);
  e->rows_sum = rows;
  if (e->eval_mode) { /* env_locomotion.py:115-116 uses the body_xyz of the previous calc_state */
    e->walk_target[0] = e->body_xyz[0] + 4; e->walk_target[1] = 0; e->walk_target[2] = 1.0;
  }
  w3d_calc_state(m, e, &e->last_contacts);
  /* calc_env_state (env_locomotion.py:204-222) */
  int nstate = 6 + 2 * A + m->n_feet;
  for (int k = 0; k < nstate; k++)
    if (!isfinite(e->robot_state[k])) e->done = 1;
  double old_lin = e->linear_potential;
  w3d_calc_potential(e, p->dt * p->substeps);
  e->progress = e->linear_potential - old_lin;
  e->posture_penalty = 0;
  double pitch = e->body_rpy[1], roll = e->body_rpy[0];
  if (!(-0.2 < pitch && pitch < 0.4)) e->posture_penalty = fabs(pitch);
  if (!(-0.4 < roll && roll < 0.4)) e->posture_penalty += fabs(roll);
  double s1 = 0, s2 = 0;
  for (int d = 0; d < A; d++) { s1 += fabs(action[d] * e->joint_speeds[d]); s2 += action[d] * action[d]; }
  e->energy_penalty = 4.5 * (s1 / A) + 0.225 * (s2 / A);
  e->joints_penalty = 0.1 * e->joints_at_limit;
  e->tall_bonus = e->robot_state[0] > m->termination_height ? 2.0 : -1.0;
  if (e->tall_bonus < 0) e->done = 1;
  if (m->planar_env != 0) {
    /* Walker2DCustomEnv.step (env_locomotion.py:301-305): "self.done = False" after super().step(); a non-finite
     * state still ends the episode here (shared deviation with the kernel: the reference would stay broken) */
    e->done = 0;
    for (int k = 0; k < nstate; k++)
      if (!isfinite(e->robot_state[k])) e->done = 1;
  }
  e->target_bonus = 0;
  if (e->distance_to_target < 0.15) { e->close_count++; e->target_bonus = 2; }
  if (e->close_count >= e->stop_frames) {
    e->close_count = 0;
    w3d_randomize_target(e);
    e->walk_target[0] += e->dist * cos(e->angle);
    e->walk_target[1] += e->dist * sin(e->angle);
    w3d_calc_potential(e, p->dt * p->substeps);
  }
  *reward = e->progress + e->target_bonus - e->energy_penalty + e->tall_bonus - e->posture_penalty - e->joints_penalty;
  w3d_obs(m, e, obs);
  /* gym TimeLimit(max_episode_steps=1000), reference mocca_envs/__init__.py:52-56 */
  e->elapsed++;
  *truncated = 0;
  *done = e->done;
  if (e->elapsed >= 1000) { *truncated = !e->done; *done = 1; }
}

void orc_w3d_step_batch(const orc_model* m, const orc_params* p, orc_w3d_env* envs, int n, const double* actions,
                        double* obs, double* rewards, int* dones, int n_threads) {
  int A = m->n_dof, O = 6 + 2 * A + m->n_feet + 2;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int i = 0; i < n; i++) {
    int trunc;
    orc_w3d_step(m, p, &envs[i], actions + (size_t)i * A, obs + (size_t)i * O, &rewards[i], &dones[i], &trunc);
    if (dones[i]) orc_w3d_reset(m, p, &envs[i], obs + (size_t)i * O);
  }
}

/* ------------------------------------------------------------------ 10. Walker3DStepperEnv */
static double linspace10(double a, double b, int i) { /* np.linspace(a, b, 10)[i] */
  if (i >= 9) return b;
  return a + i * ((b - a) / 9.0);
}

/* pybullet.getQuaternionFromEuler([roll, pitch, yaw]) -> rotation matrix (local -> world) */
static void euler_to_mat(double roll, double pitch, double yaw, m3 R) {
  double cr = cos(roll * 0.5), sr = sin(roll * 0.5), cp = cos(pitch * 0.5), sp = sin(pitch * 0.5);
  double cy = cos(yaw * 0.5), sy = sin(yaw * 0.5);
  double q[4] = {sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                 cr * cp * cy + sr * sp * sy};
  quat_to_mat(q, R);
}

/* Walker3DStepperEnv.set_step_state + BaseStep.set_position (env_locomotion.py:461-465, bullet_objects.py:77-83):
 * LargePlank scaled by 2*step_radius = 0.5; base box centre = pos + (0,0,-0.1375) (offset NOT rotated, quirk Q15),
 * cover centre = base centre + R (0,0,0.125); friction 1.0, contactStiffness 30000, contactDamping 1000. */
static void stepper_place_plank(orc_stepper_env* e, int info_index, int plank) {
  const double* t = e->terrain[info_index];
  m3 R;
  euler_to_mat(t[4], t[5], t[3], R);
  orc_box* b = &e->boxes[2 * plank];
  orc_box* c = &e->boxes[2 * plank + 1];
  b->center[0] = t[0]; b->center[1] = t[1]; b->center[2] = t[2] - 0.1375;
  memcpy(b->R, R, sizeof(m3));
  /* plank_large.urdf: box 1 x 20 x 0.45 / 0.05, plank.urdf: 1 x 1.5 x 0.45 / 0.05, globalScaling 2 * step_radius */
  /* pillar.urdf: cylinders radius 1, length 0.9 / 0.1, globalScaling step_radius = 0.25 */
  double half_y = e->plank_class == 1 ? 0.375 : (e->plank_class == 2 ? 0.25 : 5.0);
  b->cylinder = c->cylinder = e->plank_class == 2;
  b->half[0] = 0.25; b->half[1] = half_y; b->half[2] = 0.1125;
  memcpy(c->R, R, sizeof(m3));
  for (int k = 0; k < 3; k++) c->center[k] = b->center[k] + R[k][2] * 0.125;
  c->half[0] = 0.25; c->half[1] = half_y; c->half[2] = 0.0125;
  for (int k = 0; k < 2; k++) {
    orc_box* x = k ? c : b;
    x->friction = 1.0;
    x->stiffness = 30000.0;
    x->damping = 1000.0 + 0.1; /* Bullet sums both bodies' contact damping; a link's default is 0.1 */
    x->id = 10 + 2 * plank + k;
  }
  e->plank_index[plank] = info_index;
}

/* env_locomotion.py:395-441 */
static void stepper_generate_placements(orc_stepper_env* e) {
  orc_rng* r = &e->base.env_rng;
  int c = e->curriculum > 9 ? 9 : e->curriculum;
  e->curriculum = c;
  const double D2R = PI / 180;
  double ratio = c / 9.0;
  double dist_lo = 0.65, dist_hi = linspace10(0.65, 1.25, c);
  double yaw_lo = -20 * ratio * D2R, yaw_hi = 20 * ratio * D2R;
  double pit_lo = -30 * ratio * D2R + PI / 2, pit_hi = 30 * ratio * D2R + PI / 2;
  double til_lo = -15 * ratio * D2R, til_hi = 15 * ratio * D2R;
  double dr[ORC_NSTEPS], dphi[ORC_NSTEPS], dth[ORC_NSTEPS], xt[ORC_NSTEPS], yt[ORC_NSTEPS];
  for (int i = 0; i < ORC_NSTEPS; i++) dr[i] = orc_rng_uniform(r, dist_lo, dist_hi);
  for (int i = 0; i < ORC_NSTEPS; i++) dphi[i] = orc_rng_uniform(r, yaw_lo, yaw_hi);
  for (int i = 0; i < ORC_NSTEPS; i++) dth[i] = orc_rng_uniform(r, pit_lo, pit_hi);
  for (int i = 0; i < ORC_NSTEPS; i++) xt[i] = orc_rng_uniform(r, til_lo, til_hi);
  for (int i = 0; i < ORC_NSTEPS; i++) yt[i] = orc_rng_uniform(r, til_lo, til_hi);
  dr[0] = 0; dphi[0] = 0; dth[0] = PI / 2;
  for (int i = 1; i < 3; i++) { dr[i] = 0.75; dphi[i] = 0; dth[i] = PI / 2; }
  for (int i = 0; i < 3; i++) { xt[i] = 0; yt[i] = 0; }
  double acc = 0;
  for (int i = 0; i < ORC_NSTEPS; i++) { acc += dphi[i]; dphi[i] = acc; }
  double dx[ORC_NSTEPS], dy[ORC_NSTEPS], dz[ORC_NSTEPS];
  for (int i = 0; i < ORC_NSTEPS; i++) {
    dx[i] = dr[i] * sin(dth[i]) * cos(dphi[i]);
    dy[i] = dr[i] * sin(dth[i]) * sin(dphi[i]);
    dz[i] = dr[i] * cos(dth[i]);
  }
  for (int i = 2; i < ORC_NSTEPS; i++) {
    double ax = fabs(dx[i]);
    double mx = ax > 0.25 * 2.5 ? ax : 0.25 * 2.5;
    double sg = dx[i] > 0 ? 1.0 : (dx[i] < 0 ? -1.0 : 0.0);
    dx[i] = sg * (mx < 1.25 ? mx : 1.25);
  }
  double x = 0, y = 0, z = 0;
  for (int i = 0; i < ORC_NSTEPS; i++) {
    x += dx[i]; y += dy[i]; z += dz[i];
    e->terrain[i][0] = x; e->terrain[i][1] = y; e->terrain[i][2] = z;
    e->terrain[i][3] = dphi[i]; e->terrain[i][4] = xt[i]; e->terrain[i][5] = yt[i];
  }
}

/* env_locomotion.py:712-759 */
static void stepper_targets(orc_stepper_env* e) {
  int N = e->next_step_index, idx[3];
  if (!e->stop_on_next_step) {
    for (int k = 0; k < 3; k++) { idx[k] = N - 1 + k; if (idx[k] > ORC_NSTEPS - 1) idx[k] = ORC_NSTEPS - 1; }
  } else {
    idx[0] = N - 1; idx[1] = N; idx[2] = N;
  }
  orc_w3d_env* b = &e->base;
  for (int k = 0; k < 3; k++) b->walk_target[k] = e->terrain[idx[2]][k];
  for (int k = 0; k < 3; k++) {
    const double* t = e->terrain[idx[k]];
    double dx = t[0] - b->body_xyz[0], dy = t[1] - b->body_xyz[1], dz = t[2] - b->body_xyz[2];
    double ang = atan2(dy, dx) - b->body_rpy[2], d = sqrt(dx * dx + dy * dy);
    e->targets[k][0] = sin(ang) * d; e->targets[k][1] = cos(ang) * d; e->targets[k][2] = dz;
    e->targets[k][3] = t[4]; e->targets[k][4] = t[5];
  }
}

static void stepper_obs(const orc_model* m, const orc_stepper_env* e, double* obs) {
  int n = 6 + 2 * m->n_dof + m->n_feet;
  for (int k = 0; k < n; k++) obs[k] = e->base.robot_state[k];
  for (int k = 0; k < 3; k++)
    for (int j = 0; j < 5; j++) obs[n + 5 * k + j] = e->targets[k][j];
}

void orc_stepper_seed(orc_stepper_env* e, const uint32_t* key, int len, int at_construction) {
  orc_w3d_seed(&e->base, key, len, at_construction);
}

void orc_stepper_reset(const orc_model* m, const orc_params* p, orc_stepper_env* e, double* obs) {
  orc_w3d_env* b = &e->base;
  e->timestep = 0; b->done = 0; b->elapsed = 0; e->target_reached_count = 0;
  e->set_stop_on_next_step = 0; e->stop_on_next_step = 0; e->steps_reached = -1;
  e->gain_curriculum = e->curriculum > 9 ? 9 : e->curriculum;
  double pos[3] = {m->stepper_init_position[0], m->stepper_init_position[1],
                   m->stepper_init_position[2]}; /* env_locomotion.py:339,845 */
  w3d_robot_reset(m, b, pos);
  /* quirk Q5: the reference runs calc_feet_state() here on Bullet's STALE contact points of the previous episode;
   * the batched simulator defines "no contacts after reset" (documented deviation) */
  e->target_reached = 0;
  stepper_generate_placements(e);
  for (int k = 0; k < 3; k++) stepper_place_plank(e, k, k);
  e->next_step_index = 1;
  stepper_targets(e);
  w3d_calc_potential(b, p->dt * p->substeps);
  stepper_obs(m, e, obs);
}

void orc_stepper_step(const orc_model* m, const orc_params* p, orc_stepper_env* e, const double* action,
                      double* obs, double* reward, int* done, int* truncated) {
  orc_w3d_env* b = &e->base;
  int A = m->n_dof;
  double tau[ORC_MAXD];
  double gain = linspace10(1.0, 1.2, e->gain_curriculum);
  e->timestep += 1;
  for (int d = 0; d < A; d++) {
    double a = action[d];
    if (a > 1) a = 1;
    if (a < -1) a = -1;
    tau[d] = m->gain[d] * (gain * a);
  }
  orc_params pp = *p;
  pp.has_ground = 0; /* remove_ground=True (env_locomotion.py:359) */
  int rows = 0;
  orc_step_physics(m, &pp, &b->s, tau, e->boxes, 6, b->warm, &b->last_contacts, &rows);
  b->rows_sum = rows;
  e->set_stop_on_next_step = (e->next_step_index == 6 || e->next_step_index == 7 || e->next_step_index == 13 ||
                              e->next_step_index == 14);
  w3d_calc_state(m, b, NULL); /* obs foot contacts lag one step (quirk Q6) */
  int nstate = 6 + 2 * A + m->n_feet;
  for (int k = 0; k < nstate; k++)
    if (!isfinite(b->robot_state[k])) b->done = 1;
  int cur = e->next_step_index;
  /* calc_feet_state (env_locomotion.py:632-674) */
  int cover_id = 10 + 2 * (e->next_step_index % 3) + 1;
  e->target_reached = 0;
  for (int f = 0; f < 2; f++) {
    double dx = b->feet_xyz[f][0] - e->terrain[e->next_step_index][0];
    double dy = b->feet_xyz[f][1] - e->terrain[e->next_step_index][1];
    e->foot_dist_to_target[f] = sqrt(dx * dx + dy * dy);
    int contact = 0;
    for (int k = 0; k < b->last_contacts.n; k++) {
      /* "contact = 1.0 if contact_ids" (env_locomotion.py:645-646): any contact point of the foot link counts, a
       * self-contact included (either side) */
      if (b->last_contacts.link[k] == m->foot_link[f]) {
        contact = 1;
        if (b->last_contacts.partner[k] == cover_id) e->target_reached = 1;
      } else if (b->last_contacts.link_b[k] == m->foot_link[f]) contact = 1;
    }
    b->feet_contact[f] = contact;
  }
  if (e->target_reached) {
    e->target_reached_count += 1;
    if (e->target_reached_count > 120) { e->stop_on_next_step = 0; e->set_stop_on_next_step = 0; }
    if (e->target_reached_count >= 2) {
      if (!e->stop_on_next_step) {
        e->next_step_index += 1;
        e->target_reached_count = 0;
        if (e->next_step_index >= 3) { /* update_steps (env_locomotion.py:472-479) */
          int oldest = e->next_step_index % 3;
          int nxt = e->next_step_index < ORC_NSTEPS - 1 ? e->next_step_index : ORC_NSTEPS - 1;
          stepper_place_plank(e, nxt, oldest);
        }
      }
      e->stop_on_next_step = e->set_stop_on_next_step;
    }
    if (e->next_step_index >= ORC_NSTEPS) e->next_step_index -= 1;
  }
  /* calc_base_reward (env_locomotion.py:598-630) */
  double old_lin = b->linear_potential;
  w3d_calc_potential(b, p->dt * p->substeps);
  b->progress = b->linear_potential - old_lin;
  b->posture_penalty = 0;
  double pitch = b->body_rpy[1], roll = b->body_rpy[0];
  if (!(-0.2 < pitch && pitch < 0.4)) b->posture_penalty = fabs(pitch);
  if (!(-0.4 < roll && roll < 0.4)) b->posture_penalty += fabs(roll);
  double sp = sqrt(b->body_vel[0] * b->body_vel[0] + b->body_vel[1] * b->body_vel[1] + b->body_vel[2] * b->body_vel[2]);
  e->speed_penalty = sp - 1.6 > 0 ? sp - 1.6 : 0;
  double s1 = 0, s2 = 0;
  for (int d = 0; d < A; d++) { s1 += fabs(action[d] * b->joint_speeds[d]); s2 += action[d] * action[d]; }
  b->energy_penalty = 4.5 * (s1 / A) + 0.225 * (s2 / A);
  b->joints_penalty = 0.1 * b->joints_at_limit;
  double terminal_height = linspace10(0.75, 0.45, e->curriculum);
  b->tall_bonus = b->robot_state[0] > terminal_height ? 2.0 : -1.0;
  if (b->tall_bonus < 0) b->done = 1;
  /* calc_step_reward (env_locomotion.py:676-693) */
  e->step_bonus = 0;
  if (e->target_reached && e->target_reached_count == 1 && e->next_step_index != ORC_NSTEPS - 1) {
    double dist = e->foot_dist_to_target[0] < e->foot_dist_to_target[1] ? e->foot_dist_to_target[0]
                                                                        : e->foot_dist_to_target[1];
    e->step_bonus = 50 * pow(2.718, -dist / 0.25);
  }
  b->target_bonus = 0;
  int last_step = e->next_step_index == ORC_NSTEPS - 1;
  if ((last_step || e->stop_on_next_step) && b->distance_to_target < 0.15) b->target_bonus = 2.0;
  stepper_targets(e);
  if (cur != e->next_step_index) w3d_calc_potential(b, p->dt * p->substeps);
  *reward = b->progress - b->energy_penalty + e->step_bonus + b->target_bonus - e->speed_penalty * 0 + b->tall_bonus -
            b->posture_penalty - b->joints_penalty;
  if (e->random_reward) { /* env_locomotion.py:532-547 */
    double term[8] = {b->progress, -b->energy_penalty, e->step_bonus, b->target_bonus, -e->speed_penalty * 0,
                      b->tall_bonus, -b->posture_penalty, -b->joints_penalty};
    double acc = 0;
    for (int i = 0; i < 8; i++) acc += orc_rng_uniform(&b->env_rng, 0.8, 1.2) * term[i];
    *reward = acc;
  }
  stepper_obs(m, e, obs);
  e->steps_reached = (b->done || e->timestep == 999) ? e->next_step_index : -1;
  b->elapsed++;
  *truncated = 0;
  *done = b->done;
  if (b->elapsed >= 1000) { *truncated = !b->done; *done = 1; }
}

void orc_stepper_step_batch(const orc_model* m, const orc_params* p, orc_stepper_env* envs, int n,
                            const double* actions, double* obs, double* rewards, int* dones, int n_threads) {
  int A = m->n_dof, O = 6 + 2 * A + m->n_feet + 15;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int i = 0; i < n; i++) {
    int trunc;
    orc_stepper_step(m, p, &envs[i], actions + (size_t)i * A, obs + (size_t)i * O, &rewards[i], &dones[i], &trunc);
    if (dones[i]) orc_stepper_reset(m, p, &envs[i], obs + (size_t)i * O);
  }
}

/* ------------------------------------------------------------------ 11. Monkey3DCustomEnv */
/* btMatrix3x3::getRotation on the link->world rotation */
static void mat_to_quat(m3 R, double q[4]) {
  double tr = R[0][0] + R[1][1] + R[2][2];
  if (tr > 0) {
    double s = sqrt(tr + 1.0);
    q[3] = s * 0.5;
    s = 0.5 / s;
    q[0] = (R[2][1] - R[1][2]) * s; q[1] = (R[0][2] - R[2][0]) * s; q[2] = (R[1][0] - R[0][1]) * s;
  } else {
    int i = R[0][0] < R[1][1] ? (R[1][1] < R[2][2] ? 2 : 1) : (R[0][0] < R[2][2] ? 2 : 0);
    int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
    q[i] = s * 0.5;
    s = 0.5 / s;
    q[3] = (R[k][j] - R[j][k]) * s; q[j] = (R[j][i] + R[i][j]) * s; q[k] = (R[k][i] + R[i][k]) * s;
  }
}

/* BodyPart.pose() of the palm links (bullet_utils.py:100-107): COM frame == body frame for these MJCF bodies */
static void monkey_palms(const orc_model* m, orc_monkey_env* e) {
  orc_cache* c = (orc_cache*)malloc(sizeof(orc_cache));
  kin(m, &e->base.s, c);
  for (int h = 0; h < 2; h++) {
    int li = m->palm_link[h] + 1;
    v3copy(e->palm_xyz[h], c->pw[li]);
    m3 Rl;
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) Rl[a][b] = c->Rw[li][b][a];
    mat_to_quat(Rl, e->palm_quat[h]);
  }
  free(c);
}

/* set_step_state (env_locomotion.py:1244-1248): cylinder axis = local z rotated by euler(90 deg, 0, phi) */
static void monkey_place_bar(orc_monkey_env* e, int info_index, int bar) {
  const double* t = e->terrain[info_index];
  orc_bar* b = &e->bars[bar];
  m3 R;
  euler_to_mat(90 * (PI / 180), 0.0, t[3], R);
  for (int k = 0; k < 3; k++) { b->center[k] = t[k]; b->axis[k] = R[k][2]; }
  b->halflen = 2.5;   /* bar_length = 5 (env_locomotion.py:1143) */
  b->radius = 0.015;  /* step_radius (env_locomotion.py:1148) */
  b->friction = 0.5;  /* Bullet default: the changeDynamics call is commented out (bullet_objects.py:172-179) */
  b->id = 20 + bar;
  e->bar_index[bar] = info_index;
}

/* env_locomotion.py:1183-1228 with n_steps=32, yaw_limit=pitch_limit=0 */
static void monkey_generate_placements(orc_monkey_env* e) {
  orc_rng* r = &e->base.env_rng;
  const double D2R = PI / 180;
  double dr[ORC_NBARS], dphi[ORC_NBARS], dth[ORC_NBARS], dx[ORC_NBARS], dy[ORC_NBARS], dz[ORC_NBARS], phi[ORC_NBARS];
  double ylo = -0.0 * D2R, yhi = 0.0 * D2R, plo = (90 - 0) * D2R, phi_ = (90 + 0) * D2R;
  for (int i = 0; i < ORC_NBARS; i++) dr[i] = orc_rng_uniform(r, 0.3, 0.5);
  for (int i = 0; i < ORC_NBARS; i++) dphi[i] = orc_rng_uniform(r, ylo, yhi);
  for (int i = 0; i < ORC_NBARS; i++) dth[i] = orc_rng_uniform(r, plo, phi_);
  dphi[0] = 0; dphi[1] = 0;
  double acc = 0;
  for (int i = 0; i < ORC_NBARS; i++) { acc += dphi[i]; phi[i] = acc; }
  /* base_phi (env_locomotion.py:1298-1301) */
  double sgn = e->base.mirrored ? -1.0 : 1.0;
  for (int i = 0; i < ORC_NBARS; i++) {
    double deg = i == 0 ? -10 : (i == ORC_NBARS - 1 ? 10 : ((i & 1) ? 20 : -20));
    double bp = (D2R * deg) * sgn;
    dx[i] = dr[i] * sin(dth[i]) * cos(phi[i] + bp);
    double ax = fabs(dx[i]), mx = ax > 0.015 * 2.5 ? ax : 0.015 * 2.5;
    double sg = dx[i] > 0 ? 1.0 : (dx[i] < 0 ? -1.0 : 0.0);
    dx[i] = sg * (mx < 0.5 ? mx : 0.5);
    dy[i] = dr[i] * sin(dth[i]) * sin(phi[i] + bp);
    dz[i] = dr[i] * cos(dth[i]);
  }
  const double (*f)[3] = e->base.feet_xyz;
  int i0 = f[1][0] < f[0][0] ? 1 : 0; /* np.argmin: first minimum */
  int j0 = f[1][0] > f[0][0] ? 1 : 0; /* np.argmax: first maximum */
  dx[0] = f[i0][0]; dy[0] = f[i0][1]; dz[0] = f[i0][2];
  dx[1] = f[j0][0] - dx[0] + 0.01;
  dy[1] = f[j0][1] - dy[0];
  dz[1] = f[j0][2] - dz[0] - 0.02;
  dx[0] += 0.04;
  dz[0] += -20 + 0.04;
  double x = 0, y = 0, z = 0;
  for (int i = 0; i < ORC_NBARS; i++) {
    x += dx[i]; y += dy[i]; z += dz[i];
    e->terrain[i][0] = x; e->terrain[i][1] = y; e->terrain[i][2] = z + 20; e->terrain[i][3] = phi[i];
  }
  e->swing_leg = i0;
  e->pivot_leg = j0;
}

/* delta_to_k_targets(k=2) (env_locomotion.py:1489-1516) */
static void monkey_targets(orc_monkey_env* e) {
  orc_w3d_env* b = &e->base;
  for (int k = 0; k < 2; k++) {
    int idx = e->next_step_index + k;
    if (idx > ORC_NBARS - 1) idx = ORC_NBARS - 1;
    const double* t = e->terrain[idx];
    if (k == 0) v3copy(b->walk_target, t);
    double dx = t[0] - b->body_xyz[0], dy = t[1] - b->body_xyz[1], dz = t[2] - b->body_xyz[2];
    double ang = atan2(dy, dx) - b->body_rpy[2], d = sqrt(dx * dx + dy * dy);
    e->targets[k][0] = sin(ang) * d; e->targets[k][1] = cos(ang) * d; e->targets[k][2] = dz;
  }
}

/* env_locomotion.py:1351-1364 */
static void monkey_calc_potential(const orc_model* m, orc_monkey_env* e, double scene_dt) {
  orc_w3d_env* b = &e->base;
  double dx = b->walk_target[0] - b->body_xyz[0], dy = b->walk_target[1] - b->body_xyz[1];
  b->distance_to_target = sqrt(dx * dx + dy * dy);
  monkey_palms(m, e);
  const double* pxyz = e->palm_xyz[e->swing_leg == 0 ? 0 : 1];
  double d[3] = {b->walk_target[0] - pxyz[0], b->walk_target[1] - pxyz[1], b->walk_target[2] - pxyz[2]};
  b->linear_potential = -b->distance_to_target / scene_dt;
  e->swing_potential = -sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) / scene_dt;
}

/* get_observation_component (env_locomotion.py:1268-1281) */
static void monkey_obs(const orc_model* m, orc_monkey_env* e, double* obs) {
  orc_w3d_env* b = &e->base;
  int n = 6 + 2 * m->n_dof; /* robot_state[:-2] */
  for (int k = 0; k < n; k++) obs[k] = b->robot_state[k];
  obs[n] = b->feet_contact[0]; obs[n + 1] = b->feet_contact[1];
  for (int k = 0; k < 2; k++)
    for (int j = 0; j < 3; j++) obs[n + 2 + 3 * k + j] = e->targets[k][j];
  obs[n + 8] = e->swing_leg; obs[n + 9] = e->pivot_leg;
  monkey_palms(m, e);
  int h = e->swing_leg == 0 ? 0 : 1;
  for (int k = 0; k < 3; k++) obs[n + 10 + k] = b->walk_target[k] - e->palm_xyz[h][k];
  for (int k = 0; k < 4; k++) obs[n + 13 + k] = e->palm_quat[h][k];
}

void orc_monkey_seed(orc_monkey_env* e, const uint32_t* key, int len, int at_construction) {
  orc_w3d_seed(&e->base, key, len, at_construction);
}

/* calc_feet_state (env_locomotion.py:1404-1452).  contacts == NULL: the reset-time call, which in the reference
 * reads Bullet's STALE contact points of the previous episode; the batched simulator defines "no contacts after
 * reset" (documented deviation, same policy as the Stepper's quirk Q5). */
static void monkey_feet_state(const orc_model* m, orc_monkey_env* e, const orc_contacts* ct) {
  orc_w3d_env* b = &e->base;
  /* next_step and p_xyz are bound BEFORE the loop and stay stale if the index advances at i == 0 */
  int target_id = 20 + (e->next_step_index % 4);
  double px = e->terrain[e->next_step_index][0], py = e->terrain[e->next_step_index][1];
  for (int i = 0; i < 2; i++) {
    int contact = 0;
    if (ct)
      for (int k = 0; k < ct->n; k++)
        /* every static partner (bars, ground) is in all_contact_object_ids; robot links are not */
        if (ct->link[k] == m->foot_link[i] && ct->partner[k] < 1000) contact = 1;
    b->feet_contact[i] = contact;
    if (i != e->swing_leg) continue;
    double dx = b->feet_xyz[e->swing_leg][0] - px, dy = b->feet_xyz[e->swing_leg][1] - py;
    e->foot_dist_to_target = sqrt(dx * dx + dy * dy);
    int palm = m->palm_link[e->swing_leg == 0 ? 0 : 1];
    int hit = 0;
    if (ct)
      for (int k = 0; k < ct->n; k++)
        if (ct->link[k] == palm && ct->partner[k] == target_id) hit = 1;
    e->target_reached_count += hit;
    e->target_reached = e->target_reached_count >= 1;
    if (!e->target_reached) continue;
    e->target_reached_count = 0;
    e->next_step_index += 1;
    if (e->next_step_index > ORC_NBARS - 1) e->next_step_index = ORC_NBARS - 1;
    if (e->next_step_index >= 4) { /* update_steps (env_locomotion.py:1255-1266) */
      int oldest = e->next_step_index % 4;
      int nxt = e->next_step_index < ORC_NBARS - 1 ? e->next_step_index : ORC_NBARS - 1;
      monkey_place_bar(e, nxt, oldest);
    }
    e->pivot_leg = e->swing_leg;
    e->swing_leg = (e->swing_leg + 1) % 2;
  }
}

void orc_monkey_reset(const orc_model* m, const orc_params* p, orc_monkey_env* e, double* obs) {
  orc_w3d_env* b = &e->base;
  b->done = 0; b->elapsed = 0;
  e->free_fall_count = 0; e->target_reached_count = 0; e->timestep = 0; e->target_reached = 0;
  e->next_step_index = 2;
  double pos[3] = {0, 0, 20}, vel[3] = {3, 0, -1}; /* env_locomotion.py:1153-1154 */
  robot_reset_ex(m, b, pos, vel, 0);
  monkey_generate_placements(e);
  for (int k = 0; k < 4; k++) monkey_place_bar(e, k, k);
  monkey_feet_state(m, e, NULL);
  monkey_targets(e);
  monkey_calc_potential(m, e, p->dt * p->substeps);
  monkey_obs(m, e, obs);
}

void orc_monkey_step(const orc_model* m, const orc_params* p, orc_monkey_env* e, const double* action_in, double* obs,
                     double* reward, int* done, int* truncated) {
  orc_w3d_env* b = &e->base;
  int A = m->n_dof;
  double action[ORC_MAXD], tau[ORC_MAXD];
  for (int d = 0; d < A; d++) action[d] = action_in[d];
  e->timestep += 1;
  /* env_locomotion.py:1322-1323 (the reference mutates the caller's array, quirk Q11) */
  action[e->swing_leg == 0 ? 17 : 22] = 1;
  action[e->pivot_leg == 0 ? 17 : 22] = -1;
  for (int d = 0; d < A; d++) {
    double a = action[d];
    if (a > 1) a = 1;
    if (a < -1) a = -1;
    tau[d] = m->gain[d] * a;
  }
  int rows = 0;
  orc_step_physics_bars(m, p, &b->s, tau, e->bars, 4, &b->last_contacts, &rows);
  b->rows_sum = rows;
  w3d_calc_state(m, b, NULL);
  int nstate = 6 + 2 * A + m->n_feet;
  for (int k = 0; k < nstate; k++)
    if (!isfinite(b->robot_state[k])) b->done = 1;
  int cur = e->next_step_index;
  monkey_feet_state(m, e, &b->last_contacts);
  /* calc_base_reward (env_locomotion.py:1366-1402): only the swing progress and the free-fall test survive the
   * zero weights of step() */
  double old_swing = e->swing_potential;
  monkey_calc_potential(m, e, p->dt * p->substeps);
  b->progress = e->swing_potential - old_swing;
  if (e->free_fall_count > 30) b->done = 1;
  /* calc_step_reward (env_locomotion.py:1454-1466) */
  e->step_bonus = e->target_reached ? 50 * exp(-e->foot_dist_to_target / 0.25) : 0.0;
  monkey_targets(e);
  int mask = (b->feet_contact[0] + b->feet_contact[1]) == 0;
  e->free_fall_count = mask * e->free_fall_count + mask;
  if (cur != e->next_step_index) monkey_calc_potential(m, e, p->dt * p->substeps);
  *reward = b->progress + e->step_bonus - b->feet_contact[e->swing_leg];
  if (e->timestep > 180 && e->next_step_index <= 2) b->done = 1;
  monkey_obs(m, e, obs);
  b->elapsed++;
  *truncated = 0;
  *done = b->done;
  if (b->elapsed >= 1000) { *truncated = !b->done; *done = 1; }
}

void orc_monkey_step_batch(const orc_model* m, const orc_params* p, orc_monkey_env* envs, int n,
                           const double* actions, double* obs, double* rewards, int* dones, int n_threads) {
  int A = m->n_dof, O = 6 + 2 * A + 17;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int i = 0; i < n; i++) {
    int trunc;
    orc_monkey_step(m, p, &envs[i], actions + (size_t)i * A, obs + (size_t)i * O, &rewards[i], &dones[i], &trunc);
    if (dones[i]) orc_monkey_reset(m, p, &envs[i], obs + (size_t)i * O);
  }
}

int orc_sizeof_monkey_env(void) { return (int)sizeof(orc_monkey_env); }

/* ------------------------------------------------------------------ 12. CassieEnv (env_cassie.py) */
void orc_cassie_params(orc_params* p) {
  orc_default_params(p);
  p->dt = 0.03 / 50 / 1; /* control_step / llc_frame_skip / sim_frame_skip (env_cassie.py:287-289, env_base.py:81) */
  p->substeps = 1;
}

/* Cassie.calc_state (env_cassie.py:238-276) */
static void cassie_calc_state(const orc_model* m, orc_cassie_env* e) {
  orc_w3d_env* b = &e->base;
  orc_cache* c = (orc_cache*)malloc(sizeof(orc_cache));
  kin(m, &b->s, c);
  int A = m->n_ordered;
  double* st = e->robot_state;
  b->joints_at_limit = 0;
  for (int k = 0; k < A; k++) {
    int d = m->ordered_dof[k];
    double lo = m->lower[d], hi = m->upper[d], mid = 0.5 * (lo + hi);
    float nrm = (float)(2 * (b->s.q[d] - mid) / (hi - lo));
    float sp = (float)b->s.qd[d];
    st[6 + k] = nrm;
    st[6 + A + k] = sp;
    /* to_radians (env_cassie.py:204-213): float64 weights, but `thetas + 1` is a float32 array plus a Python int and
     * stays float32 -- pinned by the reference-generated trace tests/golden/ref_cassie_*.npz */
    e->rad_angles[k] = (hi - lo) * (double)(float)(nrm + 1.0f) / 2 + lo;
    e->speeds[k] = sp;
    if (fabsf(nrm) > 0.99f) b->joints_at_limit++;
  }
  v3copy(b->body_xyz, b->s.pos);
  if (isnan(e->initial_z)) e->initial_z = b->body_xyz[2];
  euler_from_quat(b->s.quat, b->body_rpy);
  double yaw = b->body_rpy[2], cy = cos(-yaw), sy = sin(-yaw);
  b->body_vel[0] = cy * b->s.vel[0] - sy * b->s.vel[1];
  b->body_vel[1] = sy * b->s.vel[0] + cy * b->s.vel[1];
  b->body_vel[2] = b->s.vel[2];
  st[0] = f32(b->body_xyz[2] - e->initial_z);
  st[1] = f32(b->body_vel[0]); st[2] = f32(b->body_vel[1]); st[3] = f32(b->body_vel[2]);
  st[4] = f32(b->body_rpy[0]); st[5] = f32(b->body_rpy[1]);
  for (int f = 0; f < m->n_feet; f++) v3copy(b->feet_xyz[f], c->pw[m->foot_link[f] + 1]);
  free(c);
}

static double cassie_potential(const orc_cassie_env* e) { /* env_cassie.py:348-354 */
  const orc_w3d_env* b = &e->base;
  double dx = b->walk_target[0] - b->body_xyz[0], dy = b->walk_target[1] - b->body_xyz[1];
  return -sqrt(dy * dy + dx * dx) / 0.03;
}

static void cassie_obs(const orc_model* m, const orc_cassie_env* e, double* obs) { /* env_cassie.py:416-431 */
  const orc_w3d_env* b = &e->base;
  int n = 6 + 2 * m->n_ordered;
  for (int k = 0; k < n; k++) obs[k] = e->robot_state[k];
  double dx = b->walk_target[0] - b->body_xyz[0], dy = b->walk_target[1] - b->body_xyz[1];
  double dth = atan2(dy, dx) - b->body_rpy[2];
  double cs = cos(-dth), sn = sin(-dth);
  obs[n] = cs * b->walk_target[0] - sn * b->walk_target[1];
  obs[n + 1] = sn * b->walk_target[0] + cs * b->walk_target[1];
}

void orc_cassie_reset(const orc_model* m, const orc_params* p, orc_cassie_env* e, double* obs) {
  (void)p;
  orc_w3d_env* b = &e->base;
  b->done = 0; b->elapsed = 0;
  b->walk_target[0] = 1000.0; b->walk_target[1] = 0.0; b->walk_target[2] = 0.0;
  /* restoreState + resetJoints + reset_velocity (env_cassie.py:363-370): the saved state has the base INERTIAL
   * frame at base_position with identity orientation (resetBasePositionAndOrientation, env_cassie.py:104-106) */
  for (int d = 0; d < m->n_dof; d++) { b->s.q[d] = m->base_joint_angles[d]; b->s.qd[d] = 0; }
  for (int k = 0; k < 3; k++) { b->s.pos[k] = m->base_position[k]; b->s.omega[k] = 0; b->s.vel[k] = 0; }
  b->s.quat[0] = b->s.quat[1] = b->s.quat[2] = 0; b->s.quat[3] = 1;
  for (int i = 0; i < ORC_WARMSZ; i++) b->warm[i] = 0;
  for (int k = 0; k < 16; k++) e->jvel[k] = 0;
  e->initial_z = NAN;
  cassie_calc_state(m, e);
  e->potential = cassie_potential(e);
  cassie_obs(m, e, obs);
}

void orc_cassie_step(const orc_model* m, const orc_params* p, orc_cassie_env* e, const double* action, double* obs,
                     double* reward, int* done, int* truncated) {
  orc_w3d_env* b = &e->base;
  int A = m->n_ordered, NP = m->n_pd;
  double target[16], jpos0[16];
  /* residual control: base angles of the powered joints + a; 0 for the springs (env_cassie.py:434-443) */
  for (int k = 0; k < NP; k++) {
    int oj = m->pd_ordered_index[k];
    target[k] = k < NP - 2 ? m->base_joint_angles[m->ordered_dof[oj]] + action[k] : 0.0;
  }
  for (int k = 0; k < A; k++) jpos0[k] = e->rad_angles[k];
  int rows_total = 0;
  for (int it = 0; it < 50; it++) { /* llc_frame_skip (env_cassie.py:288,450) */
    /* jvel_alpha = 10/50 (env_cassie.py:319,451-453).  `alpha * joint_speeds` is a Python float times a float32
     * array: the product is rounded to float32 before it joins the float64 running value (pinned by the
     * reference-generated trace tests/golden/ref_cassie_*.npz) */
    for (int k = 0; k < A; k++) e->jvel[k] = (1 - 0.2) * e->jvel[k] + (double)(0.2f * (float)e->speeds[k]);
    double tau[ORC_MAXD];
    for (int d = 0; d < m->n_dof; d++) tau[d] = 0;
    for (int k = 0; k < NP; k++) { /* pd_control (env_cassie.py:380-393) + apply_action clip (:225-230) */
      int oj = m->pd_ordered_index[k], d = m->ordered_dof[oj];
      double perr = target[k] - e->rad_angles[oj];
      double verr = 0.0 - e->jvel[oj];
      if (verr > 5) verr = 5;
      if (verr < -5) verr = -5;
      double t = m->pd_kp[k] * perr + m->pd_kd[k] * verr;
      double lim = m->gain[d];
      if (t > lim) t = lim;
      if (t < -lim) t = -lim;
      tau[d] = t;
    }
    int rows = 0;
    orc_step_physics(m, p, &b->s, tau, NULL, 0, b->warm, &b->last_contacts, &rows);
    rows_total += rows;
    cassie_calc_state(m, e);
  }
  b->rows_sum = rows_total;
  for (int k = 0; k < A; k++) e->jvel[k] = (e->rad_angles[k] - jpos0[k]) / 0.03;
  int n = 6 + 2 * A;
  for (int k = 0; k < n; k++)
    if (!isfinite(e->robot_state[k])) b->done = 1;
  /* compute_rewards (env_cassie.py:401-414) */
  double old = e->potential;
  e->potential = cassie_potential(e);
  e->progress_rew = e->potential - old;
  double minz = b->feet_xyz[0][2] < b->feet_xyz[1][2] ? b->feet_xyz[0][2] : b->feet_xyz[1][2];
  e->alive_rew = b->body_xyz[2] - minz > 0.6 ? 2.0 : -1.0;
  if (e->alive_rew < 0) b->done = 1;
  *reward = e->alive_rew + e->progress_rew;
  cassie_obs(m, e, obs);
  b->elapsed++;
  *truncated = 0;
  *done = b->done;
  if (b->elapsed >= 1000) { *truncated = !b->done; *done = 1; }
}

void orc_cassie_step_batch(const orc_model* m, const orc_params* p, orc_cassie_env* envs, int n,
                           const double* actions, double* obs, double* rewards, int* dones, int n_threads) {
  int A = 10, O = 36;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int i = 0; i < n; i++) {
    int trunc;
    orc_cassie_step(m, p, &envs[i], actions + (size_t)i * A, obs + (size_t)i * O, &rewards[i], &dones[i], &trunc);
    if (dones[i]) orc_cassie_reset(m, p, &envs[i], obs + (size_t)i * O);
  }
}

int orc_sizeof_cassie_env(void) { return (int)sizeof(orc_cassie_env); }

int orc_sizeof_stepper_env(void) { return (int)sizeof(orc_stepper_env); }
int orc_sizeof_w3d_env(void) { return (int)sizeof(orc_w3d_env); }
int orc_sizeof_model(void) { return (int)sizeof(orc_model); }
