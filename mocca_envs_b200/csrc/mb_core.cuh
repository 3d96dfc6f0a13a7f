// mb_core.cuh -- batched articulated-body simulator core, one environment per warp.
//
// Replaces, for the batched path, what the reference reaches through pybullet.stepSimulation()
// (reference mocca_envs/bullet_utils.py:352-353, physics parameters :343-350): per substep
//   collision -> forward dynamics -> PGS over limits/contacts -> semi-implicit Euler.
// It is NOT a port of btMultiBody.  Mathematically equivalent formulation chosen for a 32-lane warp:
//   * world-frame spatial algebra about the base COM (no per-link frame transforms),
//   * CRBA mass matrix + RNEA bias instead of ABA, factorised leaf-to-root as M = L^T L (no fill-in on a tree),
//   * constraint rows r are half-solved once, Y_r = L^-T J_r^T, and the projected Gauss-Seidel runs in the
//     transformed space z = L dv:  J_r dv = Y_r . z,  dv = L^-1 z.  Same row order / clamps / cone projection as
//     btMultiBodyConstraintSolver, so results equal Bullet's velocity-space PGS up to rounding.
//
// SPMD style: code inside MB_LANES(l) ... MB_END is per-lane; code outside is warp-uniform.  Under nvcc the
// lane loop is the hardware warp; under g++ (tests/emu only) it is a 32-iteration loop, which lets the very same
// source be diffed against the CPU oracle on a box without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#include "mb_tables.h"

// MB_SYNC: 0 = no CTA barrier; 1 = all warps of the CTA meet at the start of every substep (keeps them in the same phase
// of the step code so instruction-cache lines are shared); 3 = only at the first substep of an env step; 4 = every
// substep, but the CTA's warps meet in two halves (named barriers 1 / 2)
#if defined(__CUDACC__) && defined(MB_SYNC) && MB_SYNC == 1
#define MB_BLOCK_BARRIER(sub) __syncthreads()
#elif defined(__CUDACC__) && defined(MB_SYNC) && MB_SYNC == 3
#define MB_BLOCK_BARRIER(sub) do { if ((sub) == 0) __syncthreads(); } while (0)
#elif defined(__CUDACC__) && defined(MB_SYNC) && MB_SYNC == 4
#define MB_BLOCK_BARRIER(sub)                                                                         \
  do {                                                                                                \
    const unsigned hw_ = (blockDim.x >> 6) << 5; /* threads of the lower half (whole warps) */        \
    if (threadIdx.x < hw_) asm volatile("bar.sync 1, %0;" ::"r"(hw_));                               \
    else asm volatile("bar.sync 2, %0;" ::"r"(blockDim.x - hw_));                                    \
  } while (0)
#else
#define MB_BLOCK_BARRIER(sub)
#endif
#if defined(__CUDACC__) && defined(MB_SYNC) && MB_SYNC >= 2
#define MB_BLOCK_BARRIER2() __syncthreads()
#else
#define MB_BLOCK_BARRIER2()
#endif
#ifdef __CUDACC__
#define MB_NOINLINE __device__ __noinline__
#define MB_LANES(l) { const int l = (int)(threadIdx.x & 31);
#define MB_END } __syncwarp();
#define MB_END_REG }  /* the block touched registers only: no memory ordering needed */
#define MB_WARP_SYNC() __syncwarp()
template <typename T> struct LaneVar {
  T v;
  MB_HD T& operator[](int) { return v; }
  MB_HD const T& operator[](int) const { return v; }
};
MB_HD float warp_sum(const LaneVar<float>& x) {
  float v = x.v;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
MB_HD float warp_bcast(const LaneVar<float>& x, int src) { return __shfl_sync(0xffffffffu, x.v, src); }
MB_HD void warp_gather(const LaneVar<float>& x, const LaneVar<int>& src, LaneVar<float>& out) {
  out.v = __shfl_sync(0xffffffffu, x.v, src.v);
}
MB_HD unsigned warp_ballot(const LaneVar<int>& p) { return __ballot_sync(0xffffffffu, p.v != 0); }
MB_HD int mb_popc(unsigned x) { return __popc(x); }
MB_HD int mb_ffs(unsigned x) { return __ffs((int)x); }  // 1-based index of the lowest set bit, 0 for x = 0
MB_HD int mb_nth_bit(unsigned x, int n) { return (int)__fns(x, 0u, n + 1); }  // position of the n-th (0-based) set bit
// lane holding the largest value (the lowest such lane on ties)
MB_HD int warp_argmax(const LaneVar<float>& x) {
  float v = x.v;
  int i = (int)(threadIdx.x & 31);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
    if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
  }
  return i;
}
#else
#define MB_NOINLINE
#define MB_LANES(l) for (int l = 0; l < 32; ++l) {
#define MB_END }
#define MB_END_REG }
#define MB_WARP_SYNC()
template <typename T> struct LaneVar {
  T v[32];
  T& operator[](int l) { return v[l]; }
  const T& operator[](int l) const { return v[l]; }
};
inline float warp_sum(const LaneVar<float>& x) {
  float t[32];
  for (int l = 0; l < 32; ++l) t[l] = x.v[l];
  for (int o = 16; o > 0; o >>= 1) {
    float n[32];
    for (int l = 0; l < 32; ++l) n[l] = t[l] + t[l ^ o];
    for (int l = 0; l < 32; ++l) t[l] = n[l];
  }
  return t[0];
}
inline float warp_bcast(const LaneVar<float>& x, int src) { return x.v[src]; }
inline void warp_gather(const LaneVar<float>& x, const LaneVar<int>& src, LaneVar<float>& out) {
  float t[32];
  for (int l = 0; l < 32; ++l) t[l] = x.v[src.v[l] & 31];
  for (int l = 0; l < 32; ++l) out.v[l] = t[l];
}
inline unsigned warp_ballot(const LaneVar<int>& p) {
  unsigned m = 0;
  for (int l = 0; l < 32; ++l)
    if (p.v[l]) m |= 1u << l;
  return m;
}
inline int mb_popc(unsigned x) { return __builtin_popcount(x); }
inline int mb_ffs(unsigned x) { return __builtin_ffs((int)x); }
inline int mb_nth_bit(unsigned x, int n) {
  for (int i = 0; i < n; ++i) x &= x - 1u;
  return __builtin_ffs((int)x) - 1;
}
inline int warp_argmax(const LaneVar<float>& x) {
  int best = 0;
  for (int l = 1; l < 32; ++l)
    if (x.v[l] > x.v[best]) best = l;
  return best;
}
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#endif

#define MB_UNLIKELY(x) __builtin_expect(!!(x), 0)
#ifdef __CUDACC__
#define MB_ASSUME_SHARED(S) __builtin_assume(__isShared(&(S)))  /* out-of-line helpers keep LDS/STS addressing */
#define MB_FDIV(a, b) __fdividef((a), (b))
#define MB_FMUL(a, b) __fmul_rn((a), (b))  /* a product the compiler must not contract into an FMA */
#else
#define MB_FMUL(a, b) ((a) * (b))
#define MB_ASSUME_SHARED(S)
#define MB_FDIV(a, b) ((a) / (b))
#endif
// sin and cos of a joint angle: Cody-Waite reduction by pi/2 (three-part constant) and the cephes minimax polynomials
// on [-pi/4, pi/4]; |error| < 1e-7 for |x| < 300.  ~30 instructions: libdevice's sincosf drags its (never taken)
// Payne-Hanek slow path into the hot loop, 4 KB of instruction-cache footprint.
MB_HD void mb_sincos(float x, float* s, float* c) {
  const float k = rintf(x * 0.636619772f);
  float r = fmaf(k, -1.5703125f, x);
  r = fmaf(k, -4.837512969970703125e-4f, r);
  r = fmaf(k, -7.54978995489188e-8f, r);
  const int q = (int)k & 3;
  const float r2 = r * r;
  const float sp = r + r * r2 * (-1.6666654611e-1f + r2 * (8.3321608736e-3f + r2 * -1.9515295891e-4f));
  const float cp = 1.0f - 0.5f * r2 +
                   r2 * r2 * (4.166664568298827e-2f + r2 * (-1.388731625493765e-3f + r2 * 2.443315711809948e-5f));
  const float sv = (q & 1) ? cp : sp, cv = (q & 1) ? sp : cp;
  *s = (q & 2) ? -sv : sv;
  *c = ((q + 1) & 2) ? -cv : cv;
}

#define MB_MAXC 16    /* contact points kept per substep */
#define MB_POINT_WIDTH 10 /* floats per contact point of the debug output (substep points_out) */
#define MB_MAXROW 48  /* constraint rows per substep (limits + 3 per contact) */
#define MB_YSTRIDE 15 /* compact row: 6 base + <= 8 chain entries (+1 pad, odd stride = conflict-free) */
#define MB_ROW_DUAL 0x80000000u /* r_mask flag: the row continues in a second compact row (other link, same multiplier) */
#define MB_ROW_SUP 0x1FFFFFFFu  /* r_mask bits that are generalised coordinates (NU <= 29) */
#define MB_MAXBOX 6   /* static box obstacles per env (3 planks x {base, cover}) */
#define MB_MAXBAR 4   /* static bars per env (Monkey3D rendered_step_count, env_locomotion.py:1149) */
#define MB_OBST_BOXES 1
#define MB_OBST_BARS 2
#define MB_OBST_CYLS 4 /* the box records are capped cylinders about their local z axis (Pillar stepping stones) */
#define MB_PI_F 3.14159265358979323846f

// Physics constants of the reference's Bullet world (citations in include/mocca_b200.h: mb200_physics)
struct MbPhysics {
  float dt;              // env_base.py:81  control_step / llc_frame_skip / sim_frame_skip
  int substeps;          // bullet_utils.py:349 numSubSteps
  int iterations;        // bullet_utils.py:340 numSolverIterations
  float gravity;         // env_base.py:80
  float erp_contact;     // bullet_utils.py:345 setDefaultContactERP (m_erp2)
  float erp_joint;       // Bullet m_erp
  float linear_slop;     // PyBullet m_linearSlop
  float lin_damping;     // btMultiBody m_linearDamping
  float ang_damping;     // btMultiBody m_angularDamping
  float max_coord_vel;   // btMultiBody m_maxCoordinateVelocity
  float limit_max_impulse;
  float split_threshold;
  float residual_threshold;
  float ground_friction; // bullet_utils.py:371
  int has_ground;
  // static box obstacles (stepping-stone planks, bullet_objects.py:64-72): friction and the per-contact ERP / CFM
  // Bullet derives from contactStiffness / contactDamping
  float box_friction;
  float box_erp;
  float box_cfm;
  float bar_friction;  // MonkeyBar keeps Bullet's default lateral friction (bullet_objects.py:172-179 is commented out)
  int self_collision;  // robots.py:259-264 URDF_USE_SELF_COLLISION | URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS
  // Bullet-version switch: multibody contact warm starting.  0 = off (Bullet <= 2.8x and the
  // "(DISABLES CODE)" of later versions, the default); f > 0: a contact normal row starts from f times the impulse
  // its candidate point carried in the previous substep (btContactSolverInfo::m_warmstartingFactor), also across env
  // steps: the impulses live in HBM (mb200_env::warm, one slot per candidate id).
  float warmstart;
};

// contact bookkeeping word cfoot: low byte = foot index of the point (-1 = not a foot), upper bits = candidate id of
// the point (2 * geom + end, the key of the warm-start impulses)
MB_HD int mb_pack_foot(int foot, int pid) { return (foot & 255) | (pid << 8); }
MB_HD int mb_foot(int v) { return (int)(signed char)(v & 255); }
MB_HD int mb_pid(int v) { return v >> 8; }
#define MB_NWARM 384 /* warm-start slots per env: candidate ids 2 * geom + end, geoms <= 192 */
MB_HD constexpr int tri(int i, int j) { return (i * (i + 1)) / 2 + j; }  // packed lower-triangular index, j <= i
// 1 / sqrt(x) for a pivot (positive, normal): the bare MUFU.RSQ -- rsqrtf() wraps it in a denormal-range rescue (a
// compare and two predicated multiplies on the critical path of every pivot)
MB_HD float mb_rsqrt_pivot(float x) {
#ifdef __CUDACC__
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / sqrtf(x);
#endif
}
MB_HD unsigned mb_byte(unsigned x, int r) {  // byte r of x, zero-extended (one PRMT)
#ifdef __CUDACC__
  return __byte_perm(x, 0u, 0x4440u | (unsigned)r);
#else
  return (x >> (8 * r)) & 255u;
#endif
}
MB_HD bool mb_finite(float x) { return fabsf(x) <= 3.402823466e38f; }   // false for NaN and +-inf

// per-row solver constants, read with one 128-bit shared-memory load per row visit
struct alignas(16) MbRowPar {
  float rhs;   // (positional + velocity error) * jinv
  float cfm;   // cfm * jinv
  float jinv;  // 1 / (J M^-1 J^T + cfm)
  float den;   // J M^-1 J^T + cfm (0 for a degenerate row): residual = delta impulse * den
};

template <class M> struct WarpMem {
  // ---- state (generalised velocity u = [omega_w, v_w, qd])
  float u[32];
  float q[32];
  float tau[32];
  float quat[4];
  float pos[4];
  float Rb[9];
  float js[M::NJ][6];  // joint motion subspaces about O (needed until the rows are built)
  // Kinematics / body scratch is dead once the mass matrix is assembled; the constraint rows are born after
  // the factorisation.  They share storage.
  union {
    struct {
      // ---- kinematics (world axes, positions relative to the base COM)
      float jR[M::NJ][9];
      float jp[M::NJ][3];
      float jV[M::NJ + 1][6];  // [0] = base
      float jA[M::NJ + 1][6];
      union {
        struct alignas(16) {
          // ---- bodies: one 64-byte record per body, [0..9] = m, h[3], I_O{xx,yy,zz,xy,xz,yz}, [10..15] = bias wrench
          // about O (n, f).  composites() then sums the records up the body forest in place, one component per lane,
          // so that the record of body bstart(j) becomes the composite of joint j's subtree.
          float bIF[M::NB][16];
        } b;
        // contact candidate points (collision runs before the body pass); models with many candidates (hull
        // vertices, Cassie) test them on the fly against the ground plane and store nothing
        float pt[(M::NPT <= 64 ? M::NPT : 1)][3];
      } u2;
    } k;
    // ---- rows: Y_r = L^-T J_r^T stored compactly over its support (base block + ancestor chain)
    float Yc[MB_MAXROW][MB_YSTRIDE];
    // ---- tail behind the rows: the list of violated joint limits (find_limits -> setup_rows; the kinematics / body
    // scratch it overlaps is dead by then).  During collision the same words serve collide_self as its candidate list:
    // there they alias the part of u2 behind the candidate points, unused until bodies() (checked in collide_self).
    struct {
      float rows_[MB_MAXROW * MB_YSTRIDE];
      int r_dof[32];
      float r_dir[32];
    } t;
  } w;
  // ---- dynamics
  // M, then its factor L (M = L^T L), compact: row i keeps only its support in chain order
  // [base block 0..5 | ancestors root->parent | diagonal]; an ancestor's support is a prefix of its descendants'
  // After factorize(): rows are left UNSCALED (U[k][t] = pivot-time M entries, diagonal slot = pivot d_k) with
  // Ldinv = d^-1/2 and Ldi2 = 1/d beside them; L = diag(Ldinv) U.  Nothing ever reads a scaled row, so the
  // factorisation has no scaling pass and one warp barrier per pivot.
  float L[M::LSIZE];
  float Ldinv[32];
  float Ldi2[32];
  float rhs[32];
  // ---- contacts
  float cP[MB_MAXC][3];
  float cn[MB_MAXC][3];
  float cdist[MB_MAXC];
  float cmu[MB_MAXC];
  float cerp[MB_MAXC];
  float ccfm[MB_MAXC];
  int clink[MB_MAXC];
  int cfoot[MB_MAXC];
  int cpartner[MB_MAXC];
  // ---- row parameters
  // The row parameters live from setup_rows() to the end of the substep.  Before that (collision) the same bytes
  // hold the env's static obstacles, re-staged from the record at the start of every substep; between steps they
  // are scratch for the epilogue / reset.
  union {
    struct {
      MbRowPar r_par[MB_MAXROW];
      float r_app[MB_MAXROW];
      float r_mu[MB_MAXROW];
      unsigned r_mask[MB_MAXROW];  // support of the row over the generalised coordinates
    } r;
    // static box obstacles: centre[3], axes R[9] (row-major, columns = box axes), half[3], pad
    float box[MB_MAXBOX][16];
    // static bars (MonkeyBar, bullet_objects.py:148-187): centre[3], unit axis[3], half length, radius
    float bar[MB_MAXBAR][8];
    // behind the bars: the robot's own box geoms (Monkey3D fingers / hands) in world axes, staged once per substep by
    // collide(): centre[3] relative to the base COM, axes R[9], half[3], bounding radius
    struct {
      float bars_[MB_MAXBAR][8];
      float xbox[(M::NXBOX > 0 ? M::NXBOX : 1)][16];
    } xb;
    float scratch[64];
  } rc;
  int nbox;
  int nbar;
  int step_rows;  // constraint rows of the env step just taken: the scheduler's sort key (mb200.cu step_body)
  float* warm;    // this env's warm-start impulses in HBM [MB_NWARM], nullptr = warm starting off (MbPhysics::warmstart)
  // ---- loop-closure pivots (btMultiBodyPoint2Point, Cassie): world axes, relative to the base COM; [2c] on link A,
  // [2c + 1] on link B
  float lcP[(M::NLOOP > 0 ? 2 * M::NLOOP : 1)][3];
};

// ------------------------------------------------------------------------------------------------ helpers
MB_HD void mb_cross(const float* a, const float* b, float* c) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  c[0] = x; c[1] = y; c[2] = z;
}
MB_HD float mb_dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
MB_HD void mb_matvec(const float* R, const float* a, float* o) {  // row-major 3x3
  float x = R[0] * a[0] + R[1] * a[1] + R[2] * a[2];
  float y = R[3] * a[0] + R[4] * a[1] + R[5] * a[2];
  float z = R[6] * a[0] + R[7] * a[1] + R[8] * a[2];
  o[0] = x; o[1] = y; o[2] = z;
}
MB_HD void mb_matmul(const float* A, const float* B, float* C) {
  float T[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = T[i];
}
MB_HD void mb_quat_to_mat(const float* q, float* R) {  // xyzw, local->world, row-major
  float x = q[0], y = q[1], z = q[2], w = q[3];
  float d = x * x + y * y + z * z + w * w, s = 2.0f / d;
  float xs = x * s, ys = y * s, zs = z * s;
  float wx = w * xs, wy = w * ys, wz = w * zs, xx = x * xs, xy = x * ys, xz = x * zs;
  float yy = y * ys, yz = y * zs, zz = z * zs;
  R[0] = 1 - (yy + zz); R[1] = xy - wz; R[2] = xz + wy;
  R[3] = xy + wz; R[4] = 1 - (xx + zz); R[5] = yz - wx;
  R[6] = xz - wy; R[7] = yz + wx; R[8] = 1 - (xx + yy);
}
// btPlaneSpace1
MB_HD void mb_plane_space(const float* n, float* p, float* q) {
  if (fabsf(n[2]) > 0.70710678f) {
    float a = n[1] * n[1] + n[2] * n[2], k = 1.0f / sqrtf(a);
    p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
    q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
  } else {
    float a = n[0] * n[0] + n[1] * n[1], k = 1.0f / sqrtf(a);
    p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
    q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
  }
}

// dense view of the compact symmetric storage (debug / parity kernels only)
template <class M> MB_HD float mb_Lget(const float* L, int i, int j) {
  if (j > i) { const int t = i; i = j; j = t; }
  if (i == j) return L[M::rowoff(i) + M::rowlen(i) - 1];
  const unsigned sup = M::rowmask(i);
  return ((sup >> j) & 1u) ? L[M::rowoff(i) + mb_popc(sup & ((1u << j) - 1u))] : 0.0f;
}

// (t, s) of the packed lower-triangle entries p = l, l + 32, l + 64, as byte offsets 4 t / 4 s, one byte per round
struct MbPairTab {
  unsigned t4[32], s4[32];
  constexpr MbPairTab() : t4(), s4() {
    for (int l = 0; l < 32; ++l) {
      unsigned pt = 0u, ps = 0u;
      for (int r = 0; r < 3; ++r) {
        const int p = l + 32 * r;
        int t = 0;
        while ((t + 1) * (t + 2) / 2 <= p) ++t;
        pt |= (unsigned)(4 * t) << (8 * r);
        ps |= (unsigned)(4 * (p - t * (t + 1) / 2)) << (8 * r);
      }
      t4[l] = pt; s4[l] = ps;
    }
  }
};
MB_TABLE MbPairTab mb_pairtab = MbPairTab();

// ------------------------------------------------------------------------------------------------ simulator
template <class M> struct Sim {
  typedef WarpMem<M> Mem;
#if defined(MB_DISABLE_SELF) && MB_DISABLE_SELF
  enum { NJ = M::NJ, NB = M::NB, NU = M::NU, NPT = M::NPT, NSELF = 0 };  // ablation build (tools/abbench.sh)
#else
  enum { NJ = M::NJ, NB = M::NB, NU = M::NU, NPT = M::NPT, NSELF = M::NSELF };
#endif

  // ---- lane constants -----------------------------------------------------------------------------------------
  // Only what the constraint solver's inner loop reads per row visit stays in registers for the whole kernel; the
  // factorisation's pair table and the forward substitution's tree tables are (re)loaded where they are used -- the
  // Cassie and Monkey3D kernels are register-starved in the solver (round 2).
  struct LaneConst {
    LaneVar<int> tl;        // |support(l)| = slot of column l in every descendant's compact row = rowlen(l) - 1
    LaneVar<unsigned> bit;  // 1 << l: "is coordinate l in the row's support" is one LOP3 with predicate output
  };
  MB_HD static void init_lane_const(LaneConst& C) {
    MB_LANES(l)
      C.bit[l] = 1u << l;
      C.tl[l] = l < NU ? M::rowlen(l) - 1 : 0;
    MB_END_REG
  }

  // ---- A. kinematics (+ velocities / bias accelerations when with_vel) --------------------------------------
  // Pose-only kinematics (reset, observation epilogue): once per env step at most, so one out-of-line copy per kernel
  // instead of two more inline copies of the unrolled chain walk in the instruction stream.
  MB_NOINLINE static void kinematics_pose(Mem& S, const MbPhysics& P) {
    MB_ASSUME_SHARED(S);
    kinematics_impl<false>(S, P);
  }
  MB_HD static void kinematics(Mem& S, const MbPhysics& P, const LaneConst&, bool with_vel) {
    if (with_vel) kinematics_impl<true>(S, P);
    else kinematics_pose(S, P);
  }
  template <bool with_vel> MB_HD static void kinematics_impl(Mem& S, const MbPhysics& P) {
    MB_LANES(l)
      if (l == 0) {
        mb_quat_to_mat(S.quat, S.Rb);
        if (with_vel) {
          // base spatial velocity about O = base COM; bias acceleration with udot = 0 and the gravity trick:
          // A0 = (0, -w x v - g_vec)
          float wv[3];
          mb_cross(&S.u[0], &S.u[3], wv);
          for (int k = 0; k < 3; ++k) {
            S.w.k.jV[0][k] = S.u[k]; S.w.k.jV[0][3 + k] = S.u[3 + k];
            S.w.k.jA[0][k] = 0.0f; S.w.k.jA[0][3 + k] = -wv[k];
          }
          S.w.k.jA[0][5] += P.gravity;
        }
      }
    MB_END
    // sin / cos of every joint angle once, one joint per lane (the chain walk below reads them back: the three lanes of
    // a chain would otherwise each run the 30-instruction polynomial).  Scratch: Ldinv / Ldi2, dead until factorize().
    MB_LANES(l)
      if (l < NJ) {
        float sn, cs;
        mb_sincos(S.q[l], &sn, &cs);
        S.Ldinv[l] = sn;
        S.Ldi2[l] = cs;
      }
    MB_END
    // Chain walk (round 2).  Joint frames are parallel to the base frame at q = 0 (codegen.py), so a joint's frame is
    // its parent's times ONE Rodrigues rotation about the axis a as it points at q = 0:
    //   row_c(R) = cos row_c(B) + sin (row_c(B) x a) + (1 - cos)(row_c(B) . a) a,   world axis component c = row_c(B) . a.
    // Row c of a rotation depends on row c of the parent's only, and so does component c of the pivot position: lane
    // 3 ch + c carries them in registers down the ch-th root-to-leaf chain -- no level loop, no table look-ups by joint
    // (one 32-byte record per step and chain), no shared-memory round trip between a joint and its child.  The cross
    // products of the motion subspace / velocity / bias-acceleration recursion need the two other components, which
    // live in the two neighbouring lanes: eight shuffles per step.  Chains that share a prefix (the two legs below the
    // abdomen) both walk it; the first one stores it.  Lanes behind the last chain walk an idle record.
    LaneVar<float> b0, b1, b2, pc, vw, vv, aw, al;
    LaneVar<int> s1, s2;
    MB_LANES(l)
      const int g = l / 3, c = l - 3 * g;
      s1[l] = (3 * g + (c == 2 ? 0 : c + 1)) & 31;  // lanes of the cyclic successors: (x y)_c = x[i1] y[i2] - x[i2] y[i1]
      s2[l] = (3 * g + (c == 0 ? 2 : c - 1)) & 31;
      b0[l] = S.Rb[3 * c]; b1[l] = S.Rb[3 * c + 1]; b2[l] = S.Rb[3 * c + 2];
      pc[l] = 0.0f;
      vw[l] = with_vel ? S.w.k.jV[0][c] : 0.0f; vv[l] = with_vel ? S.w.k.jV[0][3 + c] : 0.0f;
      aw[l] = with_vel ? S.w.k.jA[0][c] : 0.0f; al[l] = with_vel ? S.w.k.jA[0][3 + c] : 0.0f;
    MB_END_REG
#pragma unroll
    for (int st = 0; st < M::NLEVEL; ++st) {
      LaneVar<float> ac, p1, p2, a1, a2, w1, w2, v1, v2;
      LaneVar<int> jj;
      MB_LANES(l)
        const int g = l / 3, c = l - 3 * g;
        const MbKinRec* kr = M::kin(st, g < M::NCH ? g : (int)M::NCH);
        const int j = kr->j;
        const float ax0 = kr->ax[0], ax1 = kr->ax[1], ax2 = kr->ax[2];
        const float sn = S.Ldinv[j], cs = S.Ldi2[j];
        const float B0 = b0[l], B1 = b1[l], B2 = b2[l];
        // (explicit fmaf: a sum of two products can be contracted either way round, and the device-buffer and host-buffer
        // instantiations of a step kernel must round alike -- their outputs are compared bit for bit)
        const float pn = fmaf(B2, kr->off[2], fmaf(B1, kr->off[1], fmaf(B0, kr->off[0], pc[l])));
        const float ba = fmaf(B2, ax2, fmaf(B1, ax1, B0 * ax0));
        const float x0 = fmaf(B1, ax2, -(B2 * ax1)), x1 = fmaf(B2, ax0, -(B0 * ax2)), x2 = fmaf(B0, ax1, -(B1 * ax0));
        const float kk = (1.0f - cs) * ba;
        const float r0 = fmaf(kk, ax0, fmaf(sn, x0, cs * B0)), r1 = fmaf(kk, ax1, fmaf(sn, x1, cs * B1));
        const float r2 = fmaf(kk, ax2, fmaf(sn, x2, cs * B2));
        jj[l] = kr->store ? j : -1;
        if (kr->store) {
          S.w.k.jR[j][3 * c] = r0; S.w.k.jR[j][3 * c + 1] = r1; S.w.k.jR[j][3 * c + 2] = r2;
          S.w.k.jp[j][c] = pn;
          S.js[j][c] = ba;
        }
        b0[l] = r0; b1[l] = r1; b2[l] = r2; pc[l] = pn; ac[l] = ba;
      MB_END_REG
      warp_gather(pc, s1, p1); warp_gather(pc, s2, p2);
      warp_gather(ac, s1, a1); warp_gather(ac, s2, a2);
      if (with_vel) {
        warp_gather(vw, s1, w1); warp_gather(vw, s2, w2);
        warp_gather(vv, s1, v1); warp_gather(vv, s2, v2);
      }
      MB_LANES(l)
        const int g = l / 3, c = l - 3 * g;
        const int j = jj[l];
        // linear part of the motion subspace sl = p x a: own component for the store, the two others for the crosses
        const float slc = fmaf(p1[l], a2[l], -(p2[l] * a1[l]));
        if (j >= 0) S.js[j][3 + c] = slc;
        if (with_vel) {
          const MbKinRec* kr = M::kin(st, g < M::NCH ? g : (int)M::NCH);
          const float sl1 = fmaf(p2[l], ac[l], -(pc[l] * a2[l])), sl2 = fmaf(pc[l], a1[l], -(p1[l] * ac[l]));
          const float qd = S.u[6 + kr->j];
          // A = Ap + Vp x vj (motion cross; vj x vj = 0), component c only: vj = (a, sl) qd
          const float c1c = fmaf(w1[l], a2[l], -(w2[l] * a1[l])) * qd;
          const float c23 = fmaf(-v2[l], a1[l], fmaf(v1[l], a2[l], fmaf(-w2[l], sl1, w1[l] * sl2))) * qd;
          vw[l] = fmaf(ac[l], qd, vw[l]); vv[l] = fmaf(slc, qd, vv[l]); aw[l] += c1c; al[l] += c23;
          if (j >= 0) {
            S.w.k.jV[j + 1][c] = vw[l]; S.w.k.jV[j + 1][3 + c] = vv[l];
            S.w.k.jA[j + 1][c] = aw[l]; S.w.k.jA[j + 1][3 + c] = al[l];
          }
        }
      MB_END_REG
    }
    MB_WARP_SYNC();
  }

  // ---- B. per-body spatial inertia about O and bias wrench (gyroscopic + Bullet velocity damping + gravity) --
  MB_HD static void bodies(Mem& S, const MbPhysics& P) {
    MB_LANES(l)
      if (l < NB) {
        const int o = M::bowner(l);
        const float* R = o < 0 ? S.Rb : S.w.k.jR[o];
        float com[3] = {M::bcom(l, 0), M::bcom(l, 1), M::bcom(l, 2)};
        float c[3];
        mb_matvec(R, com, c);
        if (o >= 0) { c[0] += S.w.k.jp[o][0]; c[1] += S.w.k.jp[o][1]; c[2] += S.w.k.jp[o][2]; }
        // Ic = R Ib R^T
        const float ixx = M::binertia(l, 0), iyy = M::binertia(l, 1), izz = M::binertia(l, 2);
        const float ixy = M::binertia(l, 3), ixz = M::binertia(l, 4), iyz = M::binertia(l, 5);
        float Ib[9] = {ixx, ixy, ixz, ixy, iyy, iyz, ixz, iyz, izz};
        float T[9], Ic[9];
        mb_matmul(R, Ib, T);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) Ic[3 * i + j] = T[3 * i] * R[3 * j] + T[3 * i + 1] * R[3 * j + 1] + T[3 * i + 2] * R[3 * j + 2];
        const float m = M::bmass(l);
        const float* V = S.w.k.jV[o + 1];
        const float* A = S.w.k.jA[o + 1];
        const float* w = V;
        float wxc[3], vc[3], t1[3], t2[3], t3[3], ac[3], Iw[3], Ia[3], g[3], f[3], nc[3], nO[3];
        mb_cross(w, c, wxc);
#pragma unroll
        for (int k = 0; k < 3; ++k) vc[k] = V[3 + k] + wxc[k];
        mb_cross(A, c, t1);
        mb_cross(w, V + 3, t2);
        mb_cross(w, wxc, t3);
#pragma unroll
        for (int k = 0; k < 3; ++k) ac[k] = A[3 + k] + t1[k] + t2[k] + t3[k];
        mb_matvec(Ic, w, Iw);
        mb_matvec(Ic, A, Ia);
        mb_cross(w, Iw, g);
        const float wn = sqrtf(mb_dot3(w, w)), vn = sqrtf(mb_dot3(vc, vc));
        const float kl = P.lin_damping + P.lin_damping * vn, ka = P.ang_damping + P.ang_damping * wn;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          f[k] = m * ac[k] + m * vc[k] * kl;
          nc[k] = Ia[k] + g[k] + Iw[k] * ka;
        }
        mb_cross(c, f, nO);
#pragma unroll
        for (int k = 0; k < 3; ++k) { S.w.k.u2.b.bIF[l][10 + k] = nc[k] + nO[k]; S.w.k.u2.b.bIF[l][13 + k] = f[k]; }
        float* rec = S.w.k.u2.b.bIF[l];
        const float cc = mb_dot3(c, c);
        rec[0] = m;
        rec[1] = m * c[0]; rec[2] = m * c[1]; rec[3] = m * c[2];
        rec[4] = Ic[0] + m * (cc - c[0] * c[0]);
        rec[5] = Ic[4] + m * (cc - c[1] * c[1]);
        rec[6] = Ic[8] + m * (cc - c[2] * c[2]);
        rec[7] = Ic[1] - m * c[0] * c[1];
        rec[8] = Ic[2] - m * c[0] * c[2];
        rec[9] = Ic[5] - m * c[1] * c[2];
      }
    MB_END
  }

  // ---- B2. composite inertias and summed bias wrenches, leaf to root over the body forest (codegen: c_bparent), in
  // place, one of the 16 components per lane: NB - 1 dependent adds instead of every joint lane summing its whole
  // subtree (round 2: the per-lane sums were 4.8 % of the step kernel's instructions at 5 active lanes).
  template <int B> MB_HD static void composite_from(Mem& S, int l) {  // (parents are immediates: four instructions per body)
    S.w.k.u2.b.bIF[M::k_bparent[B]][l] += S.w.k.u2.b.bIF[B][l];
    if constexpr (B > 1) composite_from<B - 1>(S, l);
  }
  MB_HD static void composites(Mem& S) {
    MB_LANES(l)
      if (l < 16) composite_from<NB - 1>(S, l);
    MB_END
  }

  // ---- C. composite inertias -> mass matrix rows (packed lower) and generalised rhs = tau - bias ------------
  MB_HD static void mass_matrix_and_rhs(Mem& S) {
    composites(S);
    MB_LANES(l)
      if (l <= NJ) {
        const float* rec = S.w.k.u2.b.bIF[l < NJ ? M::bstart(l) : 0];  // composite record (composites())
        float I[10], F[6];
#pragma unroll
        for (int k = 0; k < 10; ++k) I[k] = rec[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) F[k] = rec[10 + k];
        const float m = I[0];
        const float* h = &I[1];
        if (l < NJ) {
          const float* s = S.js[l];
          // G = I^c s : n = I_O a + h x lv ; f = m lv + a x h
          float G[6], hx[3], ah[3];
          float IOa[3] = {I[4] * s[0] + I[7] * s[1] + I[8] * s[2], I[7] * s[0] + I[5] * s[1] + I[9] * s[2],
                          I[8] * s[0] + I[9] * s[1] + I[6] * s[2]};
          mb_cross(h, s + 3, hx);
          mb_cross(s, h, ah);
#pragma unroll
          for (int k = 0; k < 3; ++k) { G[k] = IOa[k] + hx[k]; G[3 + k] = m * s[3 + k] + ah[k]; }
          float bias = 0.0f;
#pragma unroll
          for (int k = 0; k < 6; ++k) bias += s[k] * F[k];
          const int row = 6 + l;
          float* Lr = &S.L[M::rowoff(row)];
#pragma unroll
          for (int k = 0; k < 6; ++k) Lr[k] = G[k];
          const int depth = M::jdepth(l);
          const int t1 = M::ft1(row), c1 = M::fc1(row), t2 = M::RSTEPS > 1 ? M::ft2(row) : 15, c2 = M::RSTEPS > 1 ? M::fc2(row) : 0;
          for (int t = 0; t <= depth; ++t) {  // (chain joint t = t + the affine steps of row 6 + l, see setup_rows())
            const float* si = S.js[t + (6 + t >= t1 ? c1 : 0) + (M::RSTEPS > 1 && 6 + t >= t2 ? c2 : 0)];
            float v = 0.0f;
#pragma unroll
            for (int k = 0; k < 6; ++k) v += si[k] * G[k];
            Lr[6 + t] = v;
          }
          Lr[6 + depth] += M::armature(l);
          S.rhs[row] = S.tau[l] - bias;
        } else {
          // base 6x6 block: [[I_O, [h]x], [-[h]x, m 1]], lower triangle
          float* L0 = S.L;
          L0[tri(0, 0)] = I[4];
          L0[tri(1, 0)] = I[7]; L0[tri(1, 1)] = I[5];
          L0[tri(2, 0)] = I[8]; L0[tri(2, 1)] = I[9]; L0[tri(2, 2)] = I[6];
          L0[tri(3, 0)] = 0.0f; L0[tri(3, 1)] = h[2]; L0[tri(3, 2)] = -h[1]; L0[tri(3, 3)] = m;
          L0[tri(4, 0)] = -h[2]; L0[tri(4, 1)] = 0.0f; L0[tri(4, 2)] = h[0]; L0[tri(4, 3)] = 0.0f; L0[tri(4, 4)] = m;
          L0[tri(5, 0)] = h[1]; L0[tri(5, 1)] = -h[0]; L0[tri(5, 2)] = 0.0f; L0[tri(5, 3)] = 0.0f; L0[tri(5, 4)] = 0.0f;
          L0[tri(5, 5)] = m;
#pragma unroll
          for (int k = 0; k < 6; ++k) S.rhs[k] = -F[k];
        }
      }
    MB_END
  }

  // ---- D. M = L^T L, processed leaf-to-root so the tree sparsity of M is preserved (no fill-in) -------------
  // Compact rows: entry t of row k belongs to column i_t = (t < 6 ? t : 6 + chain_k[t-6]), and row i_t has exactly
  // t off-diagonal entries occupying the same slots 0..t-1 -- every update L[i_t][s] -= L[k][t] L[k][s] (s <= t)
  // is a contiguous prefix.  The nk(nk+1)/2 updates of step k are spread over the 32 lanes (<= 3 rounds).
  // With RHS the backward substitution L^T y = rhs rides along (same visiting order, k descending): the pivot row
  // pushes rhs_col -= U[k][col] rhs_k / d_k.  S.rhs is left holding d^1/2 y, which is exactly the pre-scaled input
  // solve_L<true> wants, so the forward-dynamics solve never touches a square root besides the pivot's rsqrt.
  // Addressing (round 2): the rows along a chain are stored like a packed dense triangle, so the pair (t, s) with packed
  // index p = t (t + 1) / 2 + s = l + 32 r updates word p of L, shifted by a per-pivot constant behind each branch point
  // of the pivot's chain (codegen: c_ft / c_fd / c_fc, at most FSTEPS = 2 steps): no table look-up on the lane path.
  template <bool RHS> MB_HD static void factorize(Mem& S, const LaneConst& C) {
#if !defined(MB_FACT_UNROLL) || MB_FACT_UNROLL
    pivots_from<RHS, NU - 1>(S, C);
#else
#pragma unroll 1
    for (int k = NU - 1; k >= 0; --k) {
      const int offk = M::c_rowoff(k), nk = M::c_rowlen(k) - 1;
      const int t1 = M::c_ft1(k), d1 = M::c_fd1(k), c1 = M::c_fc1(k), p1 = (t1 * (t1 + 1)) >> 1;
      const int t2 = M::FSTEPS > 1 ? M::c_ft2(k) : 15, d2 = M::FSTEPS > 1 ? M::c_fd2(k) : 0;
      const int c2 = M::FSTEPS > 1 ? M::c_fc2(k) : 0, p2 = (t2 * (t2 + 1)) >> 1;
      const float dkk = S.L[offk + nk];
      const float inv = mb_rsqrt_pivot(dkk), invd = inv * inv;
      const float ck = RHS ? S.rhs[k] * invd : 0.0f;
      const int npairs = (nk * (nk + 1)) >> 1;
      // byte addressing throughout: one add per operand address (the 4 t / 4 s fields are lane constants, the row base
      // and the step sizes are per-pivot uniforms).  Lanes behind the last pair compute on in-bounds garbage (at most
      // word LSIZE + 53 of the L / Ldinv / Ldi2 block) and only their store is predicated: no divergence in the loop.
      const char* Lk = (const char*)&S.L[offk];
      char* L0 = (char*)&S.L[0];
      const int d14 = 4 * d1, d24 = 4 * d2;
      MB_LANES(l)
        if (l == 31) { S.Ldinv[k] = inv; S.Ldi2[k] = invd; }
        if (RHS && l < nk) {
          int col = l + (l >= t1 ? c1 : 0);
          if (M::FSTEPS > 1) col += l >= t2 ? c2 : 0;
          S.rhs[col] -= *(const float*)(Lk + 4 * l) * ck;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          if (32 * r < npairs) {  // uniform: whole rounds are skipped for short rows
            const int p = l + 32 * r;
            const unsigned t4 = mb_byte(mb_pairtab.t4[l], r), s4 = mb_byte(mb_pairtab.s4[l], r);
            int dst4 = 4 * p + (p >= p1 ? d14 : 0);
            if (M::FSTEPS > 1) dst4 += p >= p2 ? d24 : 0;
            float* dst = (float*)(L0 + dst4);
            const float v = *dst - (*(const float*)(Lk + t4) * invd) * *(const float*)(Lk + s4);
            if (p < npairs) *dst = v;
          }
        }
      MB_END
    }
#endif
  }

  // Unrolled form (MB_FACT_UNROLL, the default): one template instance per pivot, so the row offset, the row length,
  // the number of rounds, the branch-point steps and every shared-memory offset are immediates -- no constant-bank
  // loads, no index arithmetic, no loop: ~32 instead of ~65 instructions per pivot for 14 KB of straight-line code.
  template <bool RHS, int K> MB_HD static void pivot(Mem& S, const LaneConst& C) {
    constexpr int offk = M::k_rowoff[K], nk = M::k_rowlen[K] - 1, npairs = (nk * (nk + 1)) / 2;
    constexpr int t1 = M::k_ft1[K], c1 = M::k_fc1[K], p1 = (t1 * (t1 + 1)) / 2, d14 = 4 * M::k_fd1[K];
    constexpr int t2 = M::k_ft2[K], c2 = M::k_fc2[K], p2 = (t2 * (t2 + 1)) / 2, d24 = 4 * M::k_fd2[K];
    const float dkk = S.L[offk + nk];
    const float inv = mb_rsqrt_pivot(dkk), invd = inv * inv;
    const float ck = RHS ? S.rhs[K] * invd : 0.0f;
    const char* Lk = (const char*)&S.L[offk];
    char* L0 = (char*)&S.L[0];
    MB_LANES(l)
      if (l == 31) { S.Ldinv[K] = inv; S.Ldi2[K] = invd; }
      if (RHS && nk > 0) {  // (branch-free like the rounds: lanes behind the row compute on in-bounds garbage)
        int col = l;
        if (t1 < nk) col += l >= t1 ? c1 : 0;
        if (t2 < nk) col += l >= t2 ? c2 : 0;
        const float v = S.rhs[col] - *(const float*)(Lk + 4 * l) * ck;
        if (l < nk) S.rhs[col] = v;
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        if (32 * r < npairs) {
          const int p = l + 32 * r;
          const unsigned t4 = mb_byte(mb_pairtab.t4[l], r), s4 = mb_byte(mb_pairtab.s4[l], r);
          int dst4 = 4 * p;
          if (p1 < npairs) dst4 += p >= p1 ? d14 : 0;
          if (p2 < npairs) dst4 += p >= p2 ? d24 : 0;
          float* dst = (float*)(L0 + dst4);
          const float v = *dst - (*(const float*)(Lk + t4) * invd) * *(const float*)(Lk + s4);
          if (32 * r + 32 <= npairs || p < npairs) *dst = v;
        }
      }
    MB_END
  }
  template <bool RHS, int K> MB_HD static void pivots_from(Mem& S, const LaneConst& C) {
    pivot<RHS, K>(S, C);
    if constexpr (K > 0) pivots_from<RHS, K - 1>(S, C);
  }

  // ---- E. single right-hand-side solve, one generalised coordinate per lane ----------------------------------
  // L y = x, forward substitution on the unscaled rows: with xs = d^1/2 x the recurrence is
  // y_i = xs_i / d_i, xs_l -= U[l][i] y_i.  PRESCALED: x already holds xs (the fused factorisation leaves it so).
  // The six base coordinates go one by one; after that the joints of one tree level are independent of each other,
  // so a level is one step: its lanes finalise, every deeper lane fetches the value of ITS ancestor on that level
  // (one indexed shuffle) and subtracts U[l][ancestor], which sits at slot 6 + level of the lane's own compact row.
  template <bool PRESCALED> MB_HD static void solve_L(Mem& S, const LaneConst& C, LaneVar<float>& x) {
    LaneVar<float> di2;
    LaneVar<int> off, dep;
    LaneVar<unsigned> anc0, anc1;
    MB_LANES(l)
      di2[l] = S.Ldi2[l];
      off[l] = l < NU ? M::rowoff(l) : 0;
      dep[l] = l < NU ? M::cdepth(l) : 99;   // tree depth of coordinate l (-1 base block, 99 unused lane)
      anc0[l] = l < NU ? M::canc0(l) : 0u;   // coordinate index of l's ancestor at each level, 5 bits per level
      anc1[l] = l < NU ? M::canc1(l) : 0u;
      if (!PRESCALED && l < NU) x[l] *= S.L[off[l] + C.tl[l]] * S.Ldinv[l];
    MB_END_REG
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float xi = warp_bcast(x, i) * S.Ldi2[i];
      MB_LANES(l)
        if (l == i) x[l] = xi;
        else if (l > i && l < NU) x[l] -= S.L[off[l] + i] * xi;
      MB_END_REG
    }
#pragma unroll
    for (int d = 0; d < M::NLEVEL; ++d) {
      LaneVar<int> src;
      LaneVar<float> xa;
      MB_LANES(l)
        if (dep[l] == d) x[l] *= di2[l];
        src[l] = (int)(((d < 6 ? anc0[l] >> (5 * d) : anc1[l] >> (5 * (d - 6)))) & 31u);
      MB_END_REG
      warp_gather(x, src, xa);
      MB_LANES(l)
        if (dep[l] > d && dep[l] < 99) x[l] -= S.L[off[l] + 6 + d] * xa[l];
      MB_END_REG
    }
  }

  MB_HD static void clamp_u(Mem& S, const MbPhysics& P) {
    MB_LANES(l)
      if (l < NU) S.u[l] = fminf(fmaxf(S.u[l], -P.max_coord_vel), P.max_coord_vel);
    MB_END
  }

  // ---- F. narrow phase: sphere / capsule-end candidates vs the ground plane z = 0 and static boxes ------------
  // Contacts are listed obstacle-major (ground for every point, then box 0, ...) like the oracle.
  MB_HD static bool sphere_box(const float* c, float r, const float* bx, float thresh, float* pa, float* n,
                               float* dist) {
    const float* R = bx + 3;
    const float* half = bx + 12;
    const float d[3] = {c[0] - bx[0], c[1] - bx[1], c[2] - bx[2]};
    float cl[3], q[3], nl[3];
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      cl[k] = R[k] * d[0] + R[3 + k] * d[1] + R[6 + k] * d[2];
      q[k] = fminf(fmaxf(cl[k], -half[k]), half[k]);
      if (q[k] != cl[k]) inside = false;
    }
    if (!inside) {
      const float df[3] = {cl[0] - q[0], cl[1] - q[1], cl[2] - q[2]};
      const float len = sqrtf(df[0] * df[0] + df[1] * df[1] + df[2] * df[2]);
      *dist = len - r;
      if (*dist >= thresh) return false;
      const float il = 1.0f / len;
      nl[0] = df[0] * il; nl[1] = df[1] * il; nl[2] = df[2] * il;
    } else {
      int ax = 0;
      float best = 1e30f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float pen = half[k] - fabsf(cl[k]);
        if (pen < best) { best = pen; ax = k; }
      }
      nl[0] = nl[1] = nl[2] = 0.0f;
      const float sgn = (ax == 0 ? cl[0] : (ax == 1 ? cl[1] : cl[2])) >= 0.0f ? 1.0f : -1.0f;
      if (ax == 0) nl[0] = sgn; else if (ax == 1) nl[1] = sgn; else nl[2] = sgn;
      *dist = -best - r;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      n[k] = R[3 * k] * nl[0] + R[3 * k + 1] * nl[1] + R[3 * k + 2] * nl[2];
      pa[k] = c[k] - r * n[k];
    }
    return true;
  }

  // sphere vs capped cylinder about the record's local z axis (Pillar, bullet_objects.py:86-90; pillar.urdf): radius
  // half[0], half length half[2].  Same conventions as sphere_box; Bullet runs its convex-convex pipeline here.
  MB_HD static bool sphere_cyl(const float* c, float r, const float* bx, float thresh, float* pa, float* n,
                               float* dist) {
    const float* R = bx + 3;
    const float rad = bx[12], h = bx[14];
    const float d[3] = {c[0] - bx[0], c[1] - bx[1], c[2] - bx[2]};
    float cl[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) cl[k] = R[k] * d[0] + R[3 + k] * d[1] + R[6 + k] * d[2];
    const float rho = sqrtf(cl[0] * cl[0] + cl[1] * cl[1]);
    const float ir = rho > 1e-12f ? 1.0f / rho : 0.0f;
    const float ux = cl[0] * ir, uy = cl[1] * ir;  // radial unit vector (0 on the axis)
    const bool in_r = rho <= rad, in_z = fabsf(cl[2]) <= h;
    float nl[3];
    if (!(in_r && in_z)) {
      const float qr = fminf(rho, rad), qz = fminf(fmaxf(cl[2], -h), h);
      const float dr = rho - qr, dz = cl[2] - qz;
      const float len = sqrtf(dr * dr + dz * dz);
      *dist = len - r;
      if (*dist >= thresh) return false;
      const float il = 1.0f / len;
      nl[0] = ux * dr * il; nl[1] = uy * dr * il; nl[2] = dz * il;
    } else {
      const float pen_r = rad - rho, pen_z = h - fabsf(cl[2]);
      if (pen_z <= pen_r) { nl[0] = 0.0f; nl[1] = 0.0f; nl[2] = cl[2] >= 0.0f ? 1.0f : -1.0f; *dist = -pen_z - r; }
      else { nl[0] = ux; nl[1] = uy; nl[2] = 0.0f; *dist = -pen_r - r; }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      n[k] = R[3 * k] * nl[0] + R[3 * k + 1] * nl[1] + R[3 * k + 2] * nl[2];
      pa[k] = c[k] - r * n[k];
    }
    return true;
  }

  // sphere (centre c, radius r) vs bar treated as a capsule around its axis segment; everything relative to the
  // base COM (bc = bar centre - base position) so that a world far from the origin costs no precision
  MB_HD static bool sphere_bar(const float* c, float r, const float* bc, const float* bar, float thresh, float* pa,
                               float* n, float* dist) {
    const float d[3] = {c[0] - bc[0], c[1] - bc[1], c[2] - bc[2]};
    float t = d[0] * bar[3] + d[1] * bar[4] + d[2] * bar[5];
    t = fminf(fmaxf(t, -bar[6]), bar[6]);
    const float df[3] = {d[0] - t * bar[3], d[1] - t * bar[4], d[2] - t * bar[5]};
    const float len = sqrtf(df[0] * df[0] + df[1] * df[1] + df[2] * df[2]);
    *dist = len - r - bar[7];
    if (*dist >= thresh || len < 1e-12f) return false;
    const float il = 1.0f / len;
#pragma unroll
    for (int k = 0; k < 3; ++k) { n[k] = df[k] * il; pa[k] = c[k] - r * n[k]; }
    return true;
  }

  // robot box geom (Monkey3D fingers / hands) vs bar.  Bullet runs GJK/EPA (one point per frame); restated as the
  // deepest of 5 spheres of the bar's radius sampled 3 cm apart along its axis around the point nearest to the box
  // centre (same restatement as the oracle's box_bar) -- evaluated five lanes per box inside collide().

  // closest points of two segments (Ericson, Real-Time Collision Detection 5.1.9; same restatement as the oracle's
  // seg_seg): Bullet's sphere-sphere, capsuleCapsuleDistance and GJK on two capsules all reduce to this
  MB_HD static void seg_seg(const float* p1, const float* q1, const float* p2, const float* q2, float* c1, float* c2) {
    const float d1[3] = {q1[0] - p1[0], q1[1] - p1[1], q1[2] - p1[2]};
    const float d2[3] = {q2[0] - p2[0], q2[1] - p2[1], q2[2] - p2[2]};
    const float r[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    const float a = mb_dot3(d1, d1), e = mb_dot3(d2, d2), f = mb_dot3(d2, r), c = mb_dot3(d1, r);
    const float EPS = 1e-12f;
    float s2 = 0.0f, t = 0.0f;
    if (a <= EPS && e <= EPS) {
    } else if (a <= EPS) {
      t = fminf(fmaxf(MB_FDIV(f, e), 0.0f), 1.0f);
    } else if (e <= EPS) {
      s2 = fminf(fmaxf(MB_FDIV(-c, a), 0.0f), 1.0f);
    } else {
      const float b = mb_dot3(d1, d2), den = a * e - b * b;
      s2 = den > EPS ? fminf(fmaxf(MB_FDIV(b * f - c * e, den), 0.0f), 1.0f) : 0.0f;
      t = MB_FDIV(b * s2 + f, e);
      if (t < 0.0f) { t = 0.0f; s2 = fminf(fmaxf(MB_FDIV(-c, a), 0.0f), 1.0f); }
      else if (t > 1.0f) { t = 1.0f; s2 = fminf(fmaxf(MB_FDIV(b - c, a), 0.0f), 1.0f); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { c1[k] = p1[k] + d1[k] * s2; c2[k] = p2[k] + d2[k] * t; }
  }

  // narrow phase over the (<= 32) candidate pairs listed in S.w.t.r_dof; appends at contact slot `at`
  MB_HD static int collide_self_narrow(Mem& S, int cnt, int at, float erp) {
    LaneVar<int> hit, pair;
    LaneVar<float> px, py, pz, nx, ny, nz, dd;
    MB_LANES(l)
      hit[l] = 0;
      if (l < cnt) {
        const int k = S.w.t.r_dof[l];
        pair[l] = k;
        const unsigned pk = M::sp_pack(k);
        const int a0 = pk & 255u, a1 = (pk >> 8) & 255u, b0 = (pk >> 16) & 255u, b1 = pk >> 24;
        float c1[3], c2[3];
        seg_seg(S.w.k.u2.pt[a0], S.w.k.u2.pt[a1], S.w.k.u2.pt[b0], S.w.k.u2.pt[b1], c1, c2);
        const float d[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
        const float len = sqrtf(mb_dot3(d, d)), ra = M::pradius(a0), rb = M::pradius(b0);
        const float dist = len - ra - rb;
        if (dist < M::sp_thresh(k) && len >= 1e-9f) {
          const float il = 1.0f / len;
          hit[l] = 1;
          nx[l] = d[0] * il; ny[l] = d[1] * il; nz[l] = d[2] * il;
          px[l] = c1[0] - ra * nx[l]; py[l] = c1[1] - ra * ny[l]; pz[l] = c1[2] - ra * nz[l];
          dd[l] = dist;
        }
      }
    MB_END
    const unsigned mask = warp_ballot(hit);
    if (mask == 0u) return 0;
    MB_LANES(l)
      if (hit[l]) {
        const int pr = pair[l];
        const int k = at + mb_popc(mask & ((1u << l) - 1u));
        if (k < MB_MAXC) {
          S.cP[k][0] = px[l]; S.cP[k][1] = py[l]; S.cP[k][2] = pz[l];
          S.cn[k][0] = nx[l]; S.cn[k][1] = ny[l]; S.cn[k][2] = nz[l];
          S.cdist[k] = dd[l];
          S.cmu[k] = M::sp_mu(pr);
          S.cerp[k] = erp;
          S.ccfm[k] = 0.0f;
          S.clink[k] = (M::sp_own(pr) & 255) - 1;
          S.cfoot[k] = mb_pack_foot(-1, M::pid((int)(M::sp_pack(pr) & 255u)));
          S.cpartner[k] = 1000 + pr;
        }
      }
    MB_END
    return mb_popc(mask);
  }

  // self-collision (robots.py:259-264): one point per candidate geom pair, listed after the contacts with the
  // static world in pair order.  cP = point on link A, cn = normal on B towards A; the point on B is cP - cdist cn;
  // cpartner = 1000 + pair index.  Broad phase: bounding spheres of the two core segments (+ radii + breaking
  // threshold, tabulated as sp_reach); the survivors are compacted into S.w.t.r_dof (free until find_limits) so that
  // the closest-point routine usually runs once per substep instead of NSELF / 32 times.
  MB_NOINLINE static int collide_self(Mem& S, float erp_contact, int nc) {
    MB_ASSUME_SHARED(S);
    // the candidate list (w.t.r_dof) must not overlap the live kinematics arrays or the candidate points
    static_assert(NSELF == 0 || sizeof(S.w.k.jR) + sizeof(S.w.k.jp) + sizeof(S.w.k.jV) + sizeof(S.w.k.jA) +
                                        sizeof(float) * 3 * NPT <= sizeof(float) * MB_MAXROW * MB_YSTRIDE,
                  "collide_self's candidate list would alias live kinematics data");
    int ns = 0, cnt = 0;
#pragma unroll 1
    for (int pass = 0; pass * 32 < NSELF; ++pass) {
      LaneVar<int> near;
      MB_LANES(l)
        const int k = pass * 32 + l;
        near[l] = 0;
        if (k < NSELF) {
          const unsigned pk = M::sp_pack(k);
          const float* a0 = S.w.k.u2.pt[pk & 255u];
          const float* a1 = S.w.k.u2.pt[(pk >> 8) & 255u];
          const float* b0 = S.w.k.u2.pt[(pk >> 16) & 255u];
          const float* b1 = S.w.k.u2.pt[pk >> 24];
          const float dx = (a0[0] + a1[0]) - (b0[0] + b1[0]), dy = (a0[1] + a1[1]) - (b0[1] + b1[1]);
          const float dz = (a0[2] + a1[2]) - (b0[2] + b1[2]);  // twice the centre distance
          const float reach = M::sp_reach(k);
          near[l] = dx * dx + dy * dy + dz * dz < 4.0f * reach * reach;
        }
      MB_END_REG
      const unsigned mask = warp_ballot(near);
      if (mask == 0u) continue;
      const int add = mb_popc(mask);
      if (cnt + add > 32) {  // flush (rare): keeps pair order
        ns += collide_self_narrow(S, cnt, nc + ns, erp_contact);
        cnt = 0;
      }
      MB_LANES(l)
        if (near[l]) S.w.t.r_dof[cnt + mb_popc(mask & ((1u << l) - 1u))] = pass * 32 + l;
      MB_END
      cnt += add;
    }
    if (cnt > 0) ns += collide_self_narrow(S, cnt, nc + ns, erp_contact);
    MB_WARP_SYNC();  // the candidate points are dead from here on: bodies() overwrites them (union u2)
    return ns;
  }

  // ---- F1b. mesh-hull self-collision (Cassie: env_cassie.py:81-85 loads the URDF with URDF_USE_SELF_COLLISION |
  // ..._EXCLUDE_ALL_PARENTS; the only non-ancestor link pairs are left-leg vs right-leg links) -------------------------
  // Bullet runs GJK between the two btConvexHullShapes (margin 1 mm each) and reports one point per frame.  Here a hull
  // is the convex hull of <= 32 support vertices of the mesh (urdf_compiler.hull_fan_vertices), ONE VERTEX PER LANE: the
  // support function of GJK is a dot product per lane and a warp arg-max.  The sub-distance step enumerates the <= 15
  // faces of the <= 4-point simplex (closest point of each face's affine hull, valid if its barycentric weights are
  // non-negative; the nearest valid one is the closest point of the simplex).  Contact distance = core distance - 2
  // margins; cores that overlap (deeper than 2 mm: ERP 0.9 pushes contacts out long before) fall back to the overlap of
  // the two hulls along the line of their bounding-sphere centres.  Cold path, written for size; scratch in S.rc.
  struct Gjk {  // view of S.rc.scratch: simplex points of the Minkowski difference and their A-side points, weights
    float* w;    // [4][3]
    float* a;    // [4][3]
    float* lam;  // [4]
  };
  // closest point of the simplex (n points) to the origin: writes v, compacts the simplex to the supporting face
  MB_HD static int gjk_closest(const Gjk& g, int n, float* v) {
    float best = 3.0e38f, bl[4] = {0, 0, 0, 0};
    int bmask = 0;
#pragma unroll 1
    for (int mask = 1; mask < (1 << n); ++mask) {
      int id[4], k = 0;
      for (int i = 0; i < n; ++i)
        if ((mask >> i) & 1) id[k++] = i;
      const float* p0 = g.w + 3 * id[0];
      float mu[3] = {0, 0, 0};
      bool ok = true;
      if (k > 1) {
        float e[3][3], G[3][3], b[3];
        for (int i = 0; i < k - 1; ++i)
          for (int c = 0; c < 3; ++c) e[i][c] = g.w[3 * id[i + 1] + c] - p0[c];
        for (int i = 0; i < k - 1; ++i) {
          b[i] = -(e[i][0] * p0[0] + e[i][1] * p0[1] + e[i][2] * p0[2]);
          for (int j = 0; j < k - 1; ++j) G[i][j] = e[i][0] * e[j][0] + e[i][1] * e[j][1] + e[i][2] * e[j][2];
        }
        if (k == 2) {
          ok = G[0][0] > 1e-20f;
          mu[0] = ok ? b[0] / G[0][0] : 0.0f;
        } else if (k == 3) {
          const float det = G[0][0] * G[1][1] - G[0][1] * G[1][0];
          ok = det > 1e-5f * G[0][0] * G[1][1];  /* (a float32 determinant is noise below that) */
          if (ok) { mu[0] = (b[0] * G[1][1] - b[1] * G[0][1]) / det; mu[1] = (G[0][0] * b[1] - G[1][0] * b[0]) / det; }
        } else {
          const float c00 = G[1][1] * G[2][2] - G[1][2] * G[2][1], c01 = G[1][2] * G[2][0] - G[1][0] * G[2][2];
          const float c02 = G[1][0] * G[2][1] - G[1][1] * G[2][0];
          const float det = G[0][0] * c00 + G[0][1] * c01 + G[0][2] * c02;
          ok = det > 1e-4f * G[0][0] * G[1][1] * G[2][2];
          if (ok) {
            mu[0] = (b[0] * c00 + b[1] * (G[0][2] * G[2][1] - G[0][1] * G[2][2]) + b[2] * (G[0][1] * G[1][2] - G[0][2] * G[1][1])) / det;
            mu[1] = (b[0] * c01 + b[1] * (G[0][0] * G[2][2] - G[0][2] * G[2][0]) + b[2] * (G[0][2] * G[1][0] - G[0][0] * G[1][2])) / det;
            mu[2] = (b[0] * c02 + b[1] * (G[0][1] * G[2][0] - G[0][0] * G[2][1]) + b[2] * (G[0][0] * G[1][1] - G[0][1] * G[1][0])) / det;
          }
        }
      }
      if (!ok) continue;
      float l4[4] = {1.0f - mu[0] - mu[1] - mu[2], mu[0], mu[1], mu[2]};
      bool inside = true;
      for (int i = 0; i < k; ++i)
        if (l4[i] < -1e-6f) inside = false;
      if (!inside) continue;
      float p[3] = {0, 0, 0};
      if (k == 3) {
        // triangle interior: the foot of the perpendicular through the plane normal (the barycentric combination
        // squares the condition number of the edge matrix; the normal of the contact is this vector's direction)
        const float* q1 = g.w + 3 * id[1];
        const float* q2 = g.w + 3 * id[2];
        const float e0[3] = {q1[0] - p0[0], q1[1] - p0[1], q1[2] - p0[2]}, e1[3] = {q2[0] - p0[0], q2[1] - p0[1], q2[2] - p0[2]};
        float nn[3];
        mb_cross(e0, e1, nn);
        const float sc = (nn[0] * p0[0] + nn[1] * p0[1] + nn[2] * p0[2]) / (nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
        p[0] = nn[0] * sc; p[1] = nn[1] * sc; p[2] = nn[2] * sc;
      } else {
        for (int i = 0; i < k; ++i)
          for (int c = 0; c < 3; ++c) p[c] += l4[i] * g.w[3 * id[i] + c];
      }
      const float d2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
      if (d2 < best) {
        best = d2; bmask = mask;
        for (int i = 0; i < 4; ++i) bl[i] = i < k ? fmaxf(l4[i], 0.0f) : 0.0f;
        v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
      }
    }
    // compact the simplex to the supporting face: one lane writes (the enumeration above ran redundantly on every lane;
    // 32 lanes storing the same words is what compute-sanitizer racecheck reports as intra-warp hazards)
    MB_WARP_SYNC();
    MB_LANES(l)
      if (l == 0) {
        int k = 0;
        for (int i = 0; i < n; ++i)
          if ((bmask >> i) & 1) {
            for (int c = 0; c < 3; ++c) { g.w[3 * k + c] = g.w[3 * i + c]; g.a[3 * k + c] = g.a[3 * i + c]; }
            g.lam[k] = bl[k];
            ++k;
          }
      }
    MB_END
    return mb_popc((unsigned)bmask & ((1u << n) - 1u));
  }
  // narrow phase of one hull pair; appends a contact at slot `at` (returns 1) or nothing (0)
  MB_HD static int hull_pair(Mem& S, int pr, int at, float erp) {
    const unsigned pk = M::sp_pack(pr);
    const int ha = pk & 255u, hb = (pk >> 8) & 255u;
    LaneVar<float> ax, ay, az, bx, by, bz, key;
    MB_LANES(l)
      {
        const int o = M::hown(ha);
        const float* R = o < 0 ? S.Rb : S.w.k.jR[o];
        const float loc[3] = {M::hv(ha, l, 0), M::hv(ha, l, 1), M::hv(ha, l, 2)};
        float c[3];
        mb_matvec(R, loc, c);
        if (o >= 0) { c[0] += S.w.k.jp[o][0]; c[1] += S.w.k.jp[o][1]; c[2] += S.w.k.jp[o][2]; }
        ax[l] = c[0]; ay[l] = c[1]; az[l] = c[2];
      }
      {
        const int o = M::hown(hb);
        const float* R = o < 0 ? S.Rb : S.w.k.jR[o];
        const float loc[3] = {M::hv(hb, l, 0), M::hv(hb, l, 1), M::hv(hb, l, 2)};
        float c[3];
        mb_matvec(R, loc, c);
        if (o >= 0) { c[0] += S.w.k.jp[o][0]; c[1] += S.w.k.jp[o][1]; c[2] += S.w.k.jp[o][2]; }
        bx[l] = c[0]; by[l] = c[1]; bz[l] = c[2];
      }
    MB_END_REG
    Gjk g;
    g.w = S.rc.scratch; g.a = S.rc.scratch + 12; g.lam = S.rc.scratch + 24;
    float v[3] = {warp_bcast(ax, 0) - warp_bcast(bx, 0), warp_bcast(ay, 0) - warp_bcast(by, 0),
                  warp_bcast(az, 0) - warp_bcast(bz, 0)};
    int n = 0;
    bool overlap = false;
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
      const float vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
      if (vv < 1e-14f) { overlap = true; break; }
      MB_LANES(l)
        key[l] = -(ax[l] * v[0] + ay[l] * v[1] + az[l] * v[2]);
      MB_END_REG
      const int sa = warp_argmax(key);
      MB_LANES(l)
        key[l] = bx[l] * v[0] + by[l] * v[1] + bz[l] * v[2];
      MB_END_REG
      const int sb = warp_argmax(key);
      const float pa[3] = {warp_bcast(ax, sa), warp_bcast(ay, sa), warp_bcast(az, sa)};
      const float w[3] = {pa[0] - warp_bcast(bx, sb), pa[1] - warp_bcast(by, sb), pa[2] - warp_bcast(bz, sb)};
      if (vv - (v[0] * w[0] + v[1] * w[1] + v[2] * w[2]) <= 1e-7f * vv && n > 0) break;  // no progress: v is the closest point
      MB_LANES(l)
        if (l < 3) { g.w[3 * n + l] = l == 0 ? w[0] : (l == 1 ? w[1] : w[2]); g.a[3 * n + l] = l == 0 ? pa[0] : (l == 1 ? pa[1] : pa[2]); }
      MB_END
      n = gjk_closest(g, n + 1, v);
      MB_WARP_SYNC();
      if (n == 4) { overlap = true; break; }  // the origin is inside the tetrahedron
    }
    const float margin = M::hull_margin();
    float nrm[3], pA[3], dist;
    if (!overlap) {
      const float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      dist = len - 2.0f * margin;
      if (dist >= M::sp_thresh(pr)) return 0;
      const float il = 1.0f / len;
      for (int c = 0; c < 3; ++c) {
        nrm[c] = v[c] * il;
        float acc = 0.0f;
        for (int i = 0; i < n; ++i) acc += g.lam[i] * g.a[3 * i + c];
        pA[c] = acc - margin * nrm[c];
      }
    } else {
      // overlapping cores: separate along the line of the bounding-sphere centres
      float ca[3], cb[3];
      for (int side = 0; side < 2; ++side) {
        const int h = side ? hb : ha, o = M::hown(h);
        const float* R = o < 0 ? S.Rb : S.w.k.jR[o];
        const float loc[3] = {M::hcen(h, 0), M::hcen(h, 1), M::hcen(h, 2)};
        float* c = side ? cb : ca;
        mb_matvec(R, loc, c);
        if (o >= 0) { c[0] += S.w.k.jp[o][0]; c[1] += S.w.k.jp[o][1]; c[2] += S.w.k.jp[o][2]; }
      }
      const float d[3] = {ca[0] - cb[0], ca[1] - cb[1], ca[2] - cb[2]};
      const float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      if (len < 1e-9f) return 0;
      for (int c = 0; c < 3; ++c) nrm[c] = d[c] / len;
      MB_LANES(l)
        key[l] = -(ax[l] * nrm[0] + ay[l] * nrm[1] + az[l] * nrm[2]);
      MB_END_REG
      const int sa = warp_argmax(key);
      MB_LANES(l)
        key[l] = bx[l] * nrm[0] + by[l] * nrm[1] + bz[l] * nrm[2];
      MB_END_REG
      const int sb = warp_argmax(key);
      const float pa[3] = {warp_bcast(ax, sa), warp_bcast(ay, sa), warp_bcast(az, sa)};
      const float pb[3] = {warp_bcast(bx, sb), warp_bcast(by, sb), warp_bcast(bz, sb)};
      dist = (pa[0] - pb[0]) * nrm[0] + (pa[1] - pb[1]) * nrm[1] + (pa[2] - pb[2]) * nrm[2] - 2.0f * margin;
      for (int c = 0; c < 3; ++c) pA[c] = pa[c] - margin * nrm[c];
    }
    if (at >= MB_MAXC) return 1;
    MB_LANES(l)
      if (l == 0) {
        S.cP[at][0] = pA[0]; S.cP[at][1] = pA[1]; S.cP[at][2] = pA[2];
        S.cn[at][0] = nrm[0]; S.cn[at][1] = nrm[1]; S.cn[at][2] = nrm[2];
        S.cdist[at] = dist;
        S.cmu[at] = M::sp_mu(pr);
        S.cerp[at] = erp;
        S.ccfm[at] = 0.0f;
        S.clink[at] = (M::sp_own(pr) & 255) - 1;
        S.cfoot[at] = mb_pack_foot(-1, 320 + pr);  // (warm-start slots 320.. : behind every candidate id 2 * geom + end)
        S.cpartner[at] = 1000 + pr;
      }
    MB_END
    return 1;
  }
  MB_NOINLINE static int collide_self_hulls_narrow(Mem& S, unsigned mask, float erp_contact, int nc) {
    MB_ASSUME_SHARED(S);
    int ns = 0;
#pragma unroll 1
    while (mask) {
      int pr = 0;
      while (!((mask >> pr) & 1u)) ++pr;
      mask &= mask - 1u;
      ns += hull_pair(S, pr, nc + ns, erp_contact);
    }
    return ns;
  }
  // broad phase inline (every substep), narrow phase out of line (only when two capsules come within reach)
  MB_HD static int collide_self_hulls(Mem& S, float erp_contact, int nc) {
    static_assert(NSELF <= 32, "one hull pair per lane in the broad phase");
    // broad phase, one pair per lane: bounding capsules (principal-axis segment + radius of each hull; the links are
    // elongated, bounding spheres of two shins side by side overlap all the time) -- distance of the two segments
    // against the tabulated reach = both radii + breaking threshold + both margins
    LaneVar<int> near;
    MB_LANES(l)
      near[l] = 0;
      if (l < NSELF) {
        const unsigned pk = M::sp_pack(l);
        float e[2][2][3];
        for (int side = 0; side < 2; ++side) {
          const int h = side ? (int)((pk >> 8) & 255u) : (int)(pk & 255u), o = M::hown(h);
          const float* R = o < 0 ? S.Rb : S.w.k.jR[o];
          for (int end = 0; end < 2; ++end) {
            const float loc[3] = {M::hseg(h, 3 * end), M::hseg(h, 3 * end + 1), M::hseg(h, 3 * end + 2)};
            mb_matvec(R, loc, e[side][end]);
            if (o >= 0) { e[side][end][0] += S.w.k.jp[o][0]; e[side][end][1] += S.w.k.jp[o][1]; e[side][end][2] += S.w.k.jp[o][2]; }
          }
        }
        float c1[3], c2[3];
        seg_seg(e[0][0], e[0][1], e[1][0], e[1][1], c1, c2);
        const float dx = c1[0] - c2[0], dy = c1[1] - c2[1], dz = c1[2] - c2[2];
        const float reach = M::sp_reach(l);
        near[l] = dx * dx + dy * dy + dz * dz < reach * reach;
      }
    MB_END_REG
    const unsigned mask = warp_ballot(near);
    return MB_UNLIKELY(mask != 0u) ? collide_self_hulls_narrow(S, mask, erp_contact, nc) : 0;
  }

  template <int OBST> MB_HD static int collide(Mem& S, const MbPhysics& P, int* overflow, int* ns_out) {
    constexpr bool BOXES = (OBST & (MB_OBST_BOXES | MB_OBST_CYLS)) != 0;
    constexpr bool CYLS = (OBST & MB_OBST_CYLS) != 0;
    constexpr bool STORE = NPT <= 64;  // else: ground plane only, points are transformed inside the test pass
    static_assert(STORE || OBST == 0, "on-the-fly candidate points support the ground plane only");
    // world positions of the candidate points (relative to the base COM)
    MB_LANES(l)
      for (int pt = l; STORE && pt < NPT; pt += 32) {
        const int o = M::powner(pt);
        const float* R = o < 0 ? S.Rb : S.w.k.jR[o];
        float loc[3] = {M::ppos(pt, 0), M::ppos(pt, 1), M::ppos(pt, 2)};
        float c[3];
        mb_matvec(R, loc, c);
        if (o >= 0) { c[0] += S.w.k.jp[o][0]; c[1] += S.w.k.jp[o][1]; c[2] += S.w.k.jp[o][2]; }
        S.w.k.u2.pt[pt][0] = c[0]; S.w.k.u2.pt[pt][1] = c[1]; S.w.k.u2.pt[pt][2] = c[2];
      }
    MB_END
    int nc = 0;
    if (P.has_ground) {  // the stadium plane z = 0
#pragma unroll 1
      for (int pass = 0; pass * 32 < NPT; ++pass) {
        LaneVar<int> hit;
        LaneVar<float> px, py, pz, dd;
        MB_LANES(l)
          const int pt = pass * 32 + l;
          hit[l] = 0;
          if (pt < NPT) {
            const float r = M::pradius(pt);
            float cfly[3];
            if (!STORE) {
              const int o = M::powner(pt);
              const float* R = o < 0 ? S.Rb : S.w.k.jR[o];
              const float loc[3] = {M::ppos(pt, 0), M::ppos(pt, 1), M::ppos(pt, 2)};
              mb_matvec(R, loc, cfly);
              if (o >= 0) { cfly[0] += S.w.k.jp[o][0]; cfly[1] += S.w.k.jp[o][1]; cfly[2] += S.w.k.jp[o][2]; }
            }
            const float* c = STORE ? S.w.k.u2.pt[STORE ? pt : 0] : cfly;
            const float dist = (S.pos[2] + c[2]) - r;
            if (dist < M::pthresh(pt)) {
              hit[l] = 1;
              px[l] = c[0]; py[l] = c[1]; pz[l] = c[2] - r; dd[l] = dist;
            }
          }
        MB_END
        const unsigned mask = warp_ballot(hit);
        if (mask == 0u) continue;
        MB_LANES(l)
          if (hit[l]) {
            const int pt = pass * 32 + l;
            const int k = nc + mb_popc(mask & ((1u << l) - 1u));
            if (k < MB_MAXC) {
              S.cP[k][0] = px[l]; S.cP[k][1] = py[l]; S.cP[k][2] = pz[l];
              S.cn[k][0] = 0.0f; S.cn[k][1] = 0.0f; S.cn[k][2] = 1.0f;
              S.cdist[k] = dd[l];
              S.cmu[k] = M::pfriction(pt) * P.ground_friction;
              S.cerp[k] = P.erp_contact;
              S.ccfm[k] = 0.0f;
              S.clink[k] = M::powner(pt);
              S.cfoot[k] = mb_pack_foot(M::pfoot(pt), M::pid(pt));
              S.cpartner[k] = 0;
            }
          }
        MB_END
        nc += mb_popc(mask);
      }
    }
    if (BOXES) {
      // Static boxes (stepping stones).  A cheap uniform cull first: the robot (all points within ~1.2 m of the base)
      // cannot reach a box whose local slab is farther away than that.  The (box, point) pairs of the boxes in reach are
      // then flattened over the lanes, box-major like the oracle's contact list (round 2: NPT = 34 points per box took
      // two passes with two live lanes in the second one).
      unsigned boxlist = 0u;
      int nact = 0;
      const int nbox = S.nbox;
#pragma unroll 1
      for (int ob = 0; ob < nbox; ++ob) {
        const float* bx = S.rc.box[ob];
        const float d[3] = {S.pos[0] - bx[0], S.pos[1] - bx[1], S.pos[2] - bx[2]};
        bool far = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float cl = bx[3 + k] * d[0] + bx[6 + k] * d[1] + bx[9 + k] * d[2];
          if (fabsf(cl) > bx[12 + k] + 1.6f) far = true;
        }
        if (!far) { boxlist |= (unsigned)ob << (4 * nact); ++nact; }
      }
      const int ntask = nact * NPT;
#pragma unroll 1
      for (int base = 0; base < ntask; base += 32) {
        LaneVar<int> hit, ptl, obl;
        LaneVar<float> px, py, pz, nx, ny, nz, dd;
        MB_LANES(l)
          const int idx = base + l;
          hit[l] = 0;
          if (idx < ntask) {
            const int bi = idx / NPT, pt = idx - bi * NPT, ob = (int)((boxlist >> (4 * bi)) & 15u);
            ptl[l] = pt; obl[l] = ob;
            const float r = M::pradius(pt);
            const float* c = S.w.k.u2.pt[STORE ? pt : 0];
            const float cw[3] = {c[0] + S.pos[0], c[1] + S.pos[1], c[2] + S.pos[2]};
            float pa[3], n[3], dist;
            if (CYLS ? sphere_cyl(cw, r, S.rc.box[ob], M::pthresh(pt), pa, n, &dist)
                     : sphere_box(cw, r, S.rc.box[ob], M::pthresh(pt), pa, n, &dist)) {
              hit[l] = 1;
              px[l] = pa[0] - S.pos[0]; py[l] = pa[1] - S.pos[1]; pz[l] = pa[2] - S.pos[2]; dd[l] = dist;
              nx[l] = n[0]; ny[l] = n[1]; nz[l] = n[2];
            }
          }
        MB_END
        const unsigned mask = warp_ballot(hit);
        if (mask == 0u) continue;
        MB_LANES(l)
          if (hit[l]) {
            const int pt = ptl[l];
            const int k = nc + mb_popc(mask & ((1u << l) - 1u));
            if (k < MB_MAXC) {
              S.cP[k][0] = px[l]; S.cP[k][1] = py[l]; S.cP[k][2] = pz[l];
              S.cn[k][0] = nx[l]; S.cn[k][1] = ny[l]; S.cn[k][2] = nz[l];
              S.cdist[k] = dd[l];
              S.cmu[k] = M::pfriction(pt) * P.box_friction;
              S.cerp[k] = P.box_erp;
              S.ccfm[k] = P.box_cfm;
              S.clink[k] = M::powner(pt);
              S.cfoot[k] = mb_pack_foot(M::pfoot(pt), M::pid(pt));
              S.cpartner[k] = 10 + obl[l];
            }
          }
        MB_END
        nc += mb_popc(mask);
      }
    }
    if (OBST & MB_OBST_BARS) {
      if (M::NXBOX > 0) {  // the robot's box geoms in world axes, once per substep (they do not depend on the bar)
        MB_LANES(l)
          if (l < M::NXBOX) {
            const int o = M::xowner(l);
            const float* R = o < 0 ? S.Rb : S.w.k.jR[o];
            float bx[16];
            const float loc[3] = {M::xpos(l, 0), M::xpos(l, 1), M::xpos(l, 2)};
            mb_matvec(R, loc, bx);
            if (o >= 0) { bx[0] += S.w.k.jp[o][0]; bx[1] += S.w.k.jp[o][1]; bx[2] += S.w.k.jp[o][2]; }
            float Rx[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Rx[i] = M::xrot(l, i);
            mb_matmul(R, Rx, bx + 3);
            bx[12] = M::xhalf(l, 0); bx[13] = M::xhalf(l, 1); bx[14] = M::xhalf(l, 2);
            bx[15] = sqrtf(bx[12] * bx[12] + bx[13] * bx[13] + bx[14] * bx[14]);
#pragma unroll
            for (int i = 0; i < 16; ++i) S.rc.xb.xbox[l][i] = bx[i];
          }
        MB_END
      }
#pragma unroll 1
      for (int ob = 0; ob < S.nbar; ++ob) {
        const float* bar = S.rc.bar[ob];
        const float bc[3] = {bar[0] - S.pos[0], bar[1] - S.pos[1], bar[2] - S.pos[2]};
        // uniform cull: every robot point lies within ~1.3 m of the base COM; compare with the distance from the
        // base to the bar's axis line
        const float ta = bc[0] * bar[3] + bc[1] * bar[4] + bc[2] * bar[5];
        const float px0 = bc[0] - ta * bar[3], py0 = bc[1] - ta * bar[4], pz0 = bc[2] - ta * bar[5];
        if (px0 * px0 + py0 * py0 + pz0 * pz0 > 1.5f * 1.5f) continue;
#pragma unroll 1
        for (int pass = 0; pass * 32 < NPT; ++pass) {
          LaneVar<int> hit;
          LaneVar<float> px, py, pz, nx, ny, nz, dd;
          MB_LANES(l)
            const int pt = pass * 32 + l;
            hit[l] = 0;
            if (pt < NPT) {
              float pa[3], n[3], dist;
              if (sphere_bar(S.w.k.u2.pt[pt], M::pradius(pt), bc, bar, M::pthresh(pt), pa, n, &dist)) {
                hit[l] = 1;
                px[l] = pa[0]; py[l] = pa[1]; pz[l] = pa[2]; dd[l] = dist;
                nx[l] = n[0]; ny[l] = n[1]; nz[l] = n[2];
              }
            }
          MB_END
          const unsigned mask = warp_ballot(hit);
          if (mask == 0u) continue;
          MB_LANES(l)
            if (hit[l]) {
              const int pt = pass * 32 + l;
              const int k = nc + mb_popc(mask & ((1u << l) - 1u));
              if (k < MB_MAXC) {
                S.cP[k][0] = px[l]; S.cP[k][1] = py[l]; S.cP[k][2] = pz[l];
                S.cn[k][0] = nx[l]; S.cn[k][1] = ny[l]; S.cn[k][2] = nz[l];
                S.cdist[k] = dd[l];
                S.cmu[k] = M::pfriction(pt) * P.bar_friction;
                S.cerp[k] = P.erp_contact;
                S.ccfm[k] = 0.0f;
                S.clink[k] = M::powner(pt);
                S.cfoot[k] = mb_pack_foot(M::pfoot(pt), M::pid(pt));
                S.cpartner[k] = 20 + ob;
              }
            }
          MB_END
          nc += mb_popc(mask);
        }
        if (M::NXBOX > 0) {
          // Robot box geoms vs this bar (box_bar: the deepest of five bar-radius spheres sampled along the axis).  A
          // conservative reach test per box first -- no sample can touch a box whose centre is farther from the axis
          // segment than its bounding radius + bar radius + threshold -- and a uniform skip when no box is near (the
          // bars the monkey is not holding); the (box, sample) pairs of the near boxes are then flattened over the lanes,
          // five neighbouring lanes per box, and the sequential "deeper by 1e-5 wins" choice is replayed from the five
          // distances (round 2: 8 live lanes walking 5 samples in turn were 16 % of the Monkey3D kernel).
          LaneVar<int> near;
          MB_LANES(l)
            near[l] = 0;
            if (l < M::NXBOX) {
              const float* bx = S.rc.xb.xbox[l];
              const float d[3] = {bx[0] - bc[0], bx[1] - bc[1], bx[2] - bc[2]};
              const float t = fminf(fmaxf(d[0] * bar[3] + d[1] * bar[4] + d[2] * bar[5], -bar[6]), bar[6]);
              const float e[3] = {d[0] - t * bar[3], d[1] - t * bar[4], d[2] - t * bar[5]};
              const float reach = bx[15] + bar[7] + M::xthresh(l) + 1e-4f;
              near[l] = e[0] * e[0] + e[1] * e[1] + e[2] * e[2] < reach * reach;
            }
          MB_END_REG
          const unsigned nearmask = warp_ballot(near);
          const int nnear = mb_popc(nearmask);
#pragma unroll 1
          for (int base = 0; base < nnear; base += 6) {
            LaneVar<int> hit, boxl, src0;
            LaneVar<float> px, py, pz, nx, ny, nz, dd, dse, e0, e1, e2, e3, e4;
            MB_LANES(l)
              const int slot = l / 5, kk = l - 5 * slot;
              hit[l] = 0; dse[l] = 1e30f; boxl[l] = 0; src0[l] = (5 * slot) & 31;
              if (slot < 6 && base + slot < nnear) {
                const int bxi = mb_nth_bit(nearmask, base + slot);
                boxl[l] = bxi;
                const float* bx = S.rc.xb.xbox[bxi];
                const float d[3] = {bx[0] - bc[0], bx[1] - bc[1], bx[2] - bc[2]};
                const float t0 = d[0] * bar[3] + d[1] * bar[4] + d[2] * bar[5];
                // visiting order 0, -1, +1, -2, +2
                const int k = kk == 0 ? 0 : ((kk & 1) ? -((kk + 1) >> 1) : (kk >> 1));
                const float t = fminf(fmaxf(t0 + 0.03f * k, -bar[6]), bar[6]);
                const float q[3] = {bc[0] + t * bar[3], bc[1] + t * bar[4], bc[2] + t * bar[5]};
                float ps[3], ns[3], ds;
                if (sphere_box(q, bar[7], bx, M::xthresh(bxi), ps, ns, &ds)) {
                  dse[l] = ds; dd[l] = ds;
                  nx[l] = -ns[0]; ny[l] = -ns[1]; nz[l] = -ns[2];
                  px[l] = ps[0] - ds * ns[0]; py[l] = ps[1] - ds * ns[1]; pz[l] = ps[2] - ds * ns[2];
                }
              }
            MB_END_REG
            LaneVar<int> s1, s2, s3, s4;
            MB_LANES(l)
              s1[l] = (src0[l] + 1) & 31; s2[l] = (src0[l] + 2) & 31; s3[l] = (src0[l] + 3) & 31; s4[l] = (src0[l] + 4) & 31;
            MB_END_REG
            warp_gather(dse, src0, e0); warp_gather(dse, s1, e1); warp_gather(dse, s2, e2);
            warp_gather(dse, s3, e3); warp_gather(dse, s4, e4);
            MB_LANES(l)
              const int slot = l / 5, kk = l - 5 * slot;
              // an outer sample wins only if deeper by more than 1e-5 m, so a bar lying parallel to a box face (all
              // samples equally deep) yields the central point
              float best = 1e30f;
              int sel = -1;
              if (e0[l] < best - 1e-5f) { best = e0[l]; sel = 0; }
              if (e1[l] < best - 1e-5f) { best = e1[l]; sel = 1; }
              if (e2[l] < best - 1e-5f) { best = e2[l]; sel = 2; }
              if (e3[l] < best - 1e-5f) { best = e3[l]; sel = 3; }
              if (e4[l] < best - 1e-5f) { best = e4[l]; sel = 4; }
              hit[l] = slot < 6 && base + slot < nnear && sel == kk;
            MB_END_REG
            const unsigned mask = warp_ballot(hit);
            if (mask != 0u) {
              MB_LANES(l)
                if (hit[l]) {
                  const int bxi = boxl[l];
                  const int k = nc + mb_popc(mask & ((1u << l) - 1u));
                  if (k < MB_MAXC) {
                    S.cP[k][0] = px[l]; S.cP[k][1] = py[l]; S.cP[k][2] = pz[l];
                    S.cn[k][0] = nx[l]; S.cn[k][1] = ny[l]; S.cn[k][2] = nz[l];
                    S.cdist[k] = dd[l];
                    S.cmu[k] = M::xfriction(bxi) * P.bar_friction;
                    S.cerp[k] = P.erp_contact;
                    S.ccfm[k] = 0.0f;
                    S.clink[k] = M::xowner(bxi);
                    S.cfoot[k] = mb_pack_foot(M::xfoot(bxi), M::xpid(bxi));
                    S.cpartner[k] = 20 + ob;
                  }
                }
              MB_END
              nc += mb_popc(mask);
            }
          }
        }
      }
    }
    if (nc > MB_MAXC) { *overflow += 1; nc = MB_MAXC; }
    int ns = 0;
    if constexpr (NSELF > 0 && M::SELF_HULLS) {
      static_assert(OBST == 0, "the hull narrow phase borrows the obstacle staging area as scratch");
      if (P.self_collision) {
        ns = collide_self_hulls(S, P.erp_contact, nc);
        if (nc + ns > MB_MAXC) { *overflow += 1; ns = MB_MAXC - nc; }
      }
    } else if constexpr (NSELF > 0 && STORE) {
      if (P.self_collision) {
        ns = collide_self(S, P.erp_contact, nc);
        if (nc + ns > MB_MAXC) { *overflow += 1; ns = MB_MAXC - nc; }
      }
    }
    *ns_out = ns;
    return nc + ns;
  }

  // ---- F2. loop-closure pivots in world axes (needs the kinematics of this substep) ---------------------------
  MB_HD static void loop_pivots(Mem& S) {
    if (M::NLOOP == 0) return;
    MB_LANES(l)
      if (l < 2 * M::NLOOP) {
        const int o = M::lc_owner(l);
        const float* R = o < 0 ? S.Rb : S.w.k.jR[o];
        const float loc[3] = {M::lc_pos(l, 0), M::lc_pos(l, 1), M::lc_pos(l, 2)};
        float c[3];
        mb_matvec(R, loc, c);
        if (o >= 0) { c[0] += S.w.k.jp[o][0]; c[1] += S.w.k.jp[o][1]; c[2] += S.w.k.jp[o][2]; }
        S.lcP[l][0] = c[0]; S.lcP[l][1] = c[1]; S.lcP[l][2] = c[2];
      }
    MB_END
  }

  // ---- G. constraint rows: J, Y = L^-T J^T, effective mass, rhs ---------------------------------------------
  // rows [0, nlim) joint limits, [nlim, nlim+nc) contact normals, then 2 friction rows per contact.
  MB_HD static int find_limits(Mem& S) {
    LaneVar<int> lim;
    MB_LANES(l)
      lim[l] = 0;
      if (l < NJ) {
        const float lo = M::lower(l), hi = M::upper(l), q = S.q[l];
        if (lo <= hi) {
          if (q - lo <= 0.0f) lim[l] = 1;
          else if (hi - q <= 0.0f) lim[l] = 2;
        }
      }
    MB_END
    const unsigned mask = warp_ballot(lim);
    MB_LANES(l)
      if (lim[l]) {
        const int k = mb_popc(mask & ((1u << l) - 1u));
        S.w.t.r_dof[k] = l;
        S.w.t.r_dir[k] = lim[l] == 1 ? 1.0f : -1.0f;
      }
    MB_END
    return mb_popc(mask);
  }

  // rows [0, nlim) joint limits; [nlim, nlim + NLC) loop closures: three world axes per constraint, each stored
  // as TWO compact rows (the part on link A and the part on link B -- their union is a tree, not a chain) that share
  // one multiplier; then contact normals and friction pairs.
  enum { NLC = 6 * M::NLOOP };
  // ---- G2. self-contact rows.  A self-contact couples two links, so like a loop row each of its three rows is stored
  // as two compact rows sharing one multiplier (flag MB_ROW_DUAL on the first): normals at S0 + 2s + {A, B}, friction
  // at S0 + 2 ncs + 4s + {t1 A, t2 A, t1 B, t2 B} (S0 = first row after the static-world contacts, s = self-contact
  // index, contact slot nc + s).  They are built by setup_rows() in the same pass as every other row -- one more lane
  // each, the unrolled register code is shared.  (Until round 2 a size-optimised out-of-line routine built them in
  // ~1 500 instructions; a substep with a self-contact then made its whole CTA wait at the next barrier, and with 14
  // envs per CTA nearly every substep had one.)
  // number of compact rows the row pass builds for (nlim, nc, ncs)
  MB_HD static int compact_rows(int nlim, int nc, int ncs) { return nlim + NLC + 3 * nc + (NSELF > 0 ? 6 * ncs : 0); }

  // One compact row, one lane: J over the row's support, Y = L^-T J^T in registers, effective mass, right-hand side.
  // S is the env the row belongs to -- under the cooperative row pass (coop_rows) that is not the calling warp's env.
  MB_HD static void row_build(Mem& S, const MbPhysics& P, int nlim, int nc, int ncs, int r) {
    const int n0 = nlim + NLC;
    const int S0 = n0 + 3 * nc;
    const float inv_dt = 1.0f / P.dt;
    // everything lives on the row's support: base block + chain (root -> constrained joint)
    float b[M::MAXSUP];
    float W[6];
    float cfm = 0.0f, mu = 0.0f, pen = 0.0f, dist = 0.0f, erp = 0.0f, dir = 0.0f;
    int kind;  // 0 limit, 1 normal, 2 friction
    int cj;    // constrained joint, -1 = base link
    if (r < nlim) {
      kind = 0;
      cj = S.w.t.r_dof[r];
      dir = S.w.t.r_dir[r];
      pen = dir > 0.0f ? S.q[cj] - M::lower(cj) : M::upper(cj) - S.q[cj];
#pragma unroll
      for (int i = 0; i < 6; ++i) W[i] = 0.0f;
    } else if (r < n0) {
      // btMultiBodyPoint2Point::createConstraintRows: contactNormalOnB = -e_ax for link A, +e_ax for link B
      kind = 3;
      const int i = r - nlim, side = i & 1, ax = (i % 6) >> 1, sd = 2 * (i / 6) + side;
      float dirv[3] = {0.0f, 0.0f, 0.0f};
      const float sg = side ? 1.0f : -1.0f;
      if (ax == 0) dirv[0] = sg; else if (ax == 1) dirv[1] = sg; else dirv[2] = sg;
      mb_cross(S.lcP[sd], dirv, W);
      W[3] = dirv[0]; W[4] = dirv[1]; W[5] = dirv[2];
      cj = M::lc_owner(sd);
    } else {
      int k, fr = -1;       // contact slot; friction direction (-1 = the normal)
      bool sideB = false;   // the part of a self-contact row on the partner link
      if (r < n0 + nc) { kind = 1; k = r - n0; }
      else if (NSELF == 0 || r < S0) { kind = 2; k = (r - n0 - nc) >> 1; fr = (r - n0 - nc) & 1; }
      else {
        kind = 4;
        const int i = r - S0;
        if (i < 2 * ncs) { k = nc + (i >> 1); sideB = (i & 1) != 0; }
        else { const int i2 = i - 2 * ncs; k = nc + (i2 >> 2); fr = i2 & 1; sideB = (i2 & 2) != 0; }
      }
      float dirv[3] = {S.cn[k][0], S.cn[k][1], S.cn[k][2]};
      if (fr >= 0) {
        float t1[3], t2[3];
        mb_plane_space(S.cn[k], t1, t2);
        dirv[0] = fr ? t2[0] : t1[0]; dirv[1] = fr ? t2[1] : t1[1]; dirv[2] = fr ? t2[2] : t1[2];
      }
      if (kind == 1) { cfm = S.ccfm[k] * inv_dt; erp = S.cerp[k]; dist = S.cdist[k] + P.linear_slop; }
      mu = S.cmu[k];
      cj = S.clink[k];
      float pcv[3] = {S.cP[k][0], S.cP[k][1], S.cP[k][2]};
      if (NSELF > 0 && sideB) {
        // setupMultiBodyContactConstraint: jacobian B is built with -direction at the point on B = pA - dist n
        const float dk = S.cdist[k];
        pcv[0] -= dk * S.cn[k][0]; pcv[1] -= dk * S.cn[k][1]; pcv[2] -= dk * S.cn[k][2];
        dirv[0] = -dirv[0]; dirv[1] = -dirv[1]; dirv[2] = -dirv[2];
        cj = ((M::sp_own(S.cpartner[k] - 1000) >> 8) & 255) - 1;
      }
      mb_cross(pcv, dirv, W);
      W[3] = dirv[0]; W[4] = dirv[1]; W[5] = dirv[2];
    }
    const int depth = cj >= 0 ? M::jdepth(cj) : -1;
    const int n = 7 + depth;  // support size
    // Affine chain addressing (see factorize()): slot t of the row's support is coordinate t, and its compact row
    // of L starts at word t (t + 1) / 2 -- both shifted by a per-row constant behind each branch point of the
    // chain.  No per-slot table look-up sits on the serial path of the half solve.
    const int kc = 6 + (cj >= 0 ? cj : 0);
    const int t1 = cj >= 0 ? M::ft1(kc) : 15, d1 = cj >= 0 ? M::fd1(kc) : 0, c1 = cj >= 0 ? M::fc1(kc) : 0;
    const int t2 = (M::RSTEPS > 1 && cj >= 0) ? M::ft2(kc) : 15, d2 = (M::RSTEPS > 1 && cj >= 0) ? M::fd2(kc) : 0;
    const int c2 = (M::RSTEPS > 1 && cj >= 0) ? M::fc2(kc) : 0;
    float rel_vel = 0.0f;
#pragma unroll
    for (int t = 0; t < 6; ++t) { b[t] = W[t]; rel_vel += W[t] * S.u[t]; }
#pragma unroll
    for (int t = 0; t < M::MAXSUP - 6; ++t) {
      float v = 0.0f;
      if (t <= depth) {
        int a = t + (6 + t >= t1 ? c1 : 0);
        if (M::RSTEPS > 1) a += 6 + t >= t2 ? c2 : 0;
        if (kind == 0) v = t == depth ? dir : 0.0f;
        else {
          const float* sj = S.js[a];
          v = sj[0] * W[0] + sj[1] * W[1] + sj[2] * W[2] + sj[3] * W[3] + sj[4] * W[4] + sj[5] * W[5];
        }
        rel_vel += v * S.u[6 + a];
      }
      b[6 + t] = v;
    }
    // half solve L^T y = J^T restricted to the support (prefix property of the compact factor)
#pragma unroll
    for (int t = M::MAXSUP - 1; t >= 0; --t) {
      if (t < n) {
        int it = t, ro = (t * (t + 1)) / 2;
        if (t >= 6) {
          if (t >= t1) { it += c1; ro += d1; }
          if (M::RSTEPS > 1 && t >= t2) { it += c2; ro += d2; }
        }
        const float ci = b[t] * S.Ldi2[it];  // rows are unscaled: L[it][s] y = U[it][s] (b_t / d_it)
        b[t] *= S.Ldinv[it];
        const float* Li = &S.L[ro];
#pragma unroll
        for (int s2 = 0; s2 < t; ++s2) b[s2] -= Li[s2] * ci;
      }
    }
    float dd = cfm;
#pragma unroll
    for (int t = 0; t < M::MAXSUP; ++t) dd += b[t] * b[t];
    const float jinv = dd > 1.1920929e-07f ? 1.0f / dd : 0.0f;
    float positional = 0.0f, verr = -rel_vel;
    if (kind == 0) {
      // btMultiBodyJointLimitConstraint: erp = m_erp unless deeper than the split-impulse threshold, in
      // which case the positional part is routed to the (never applied) split impulse
      if (pen > 0.0f) verr = -pen * inv_dt;
      else if (pen > P.split_threshold) positional = -pen * P.erp_joint * inv_dt;
    } else if (kind == 1) {
      if (dist > 0.0f) verr -= dist * inv_dt;
      else positional = -dist * erp * inv_dt;
    }
    float* Yr = S.w.Yc[r];
#pragma unroll
    for (int t = 0; t < M::MAXSUP; ++t)
      if (t < n) Yr[t] = b[t];
    S.rc.r.r_mask[r] = 0x3Fu | (cj >= 0 ? (M::janc(cj) << 6) : 0u);
    MbRowPar par;
    if (kind >= 3) {  // partial sums; the two parts of a loop / self-contact row are combined below
      par.rhs = rel_vel; par.jinv = dd; par.den = 0.0f;
    } else {
      par.rhs = (positional + verr) * jinv; par.jinv = jinv; par.den = jinv != 0.0f ? dd : 0.0f;
    }
    par.cfm = cfm * jinv;
    S.rc.r.r_par[r] = par;
    S.rc.r.r_app[r] = 0.0f;
    S.rc.r.r_mu[r] = mu;
  }

  // the rows of one env, built by its own warp (lane-loop emulation; kernels without the cooperative pass)
  MB_HD static void setup_rows(Mem& S, const MbPhysics& P, int nlim, int nc, int ncs) {
    const int R = compact_rows(nlim, nc, ncs);
#pragma unroll 1
    for (int base = 0; base < R; base += 32) {
      MB_LANES(l)
        if (base + l < R) row_build(S, P, nlim, nc, ncs, base + l);
      MB_END
    }
    rows_combine(S, P, nlim, nc, ncs);
  }

#if defined(__CUDACC__) && defined(MB_COOP_ROWS) && MB_COOP_ROWS
  // Cooperative row pass (round 2; compile with -DMB_COOP_ROWS=1 -- measured and NOT kept as the default: Walker3D +-0,
  // Stepper +1.3 %, Monkey3D +-0, Cassie -1.8 %, profiles/r3o_coop_rows_ab.txt: the instructions it saves are paid back
  // by the two extra CTA barriers and by the few working warps running latency-bound while the others wait).
  // One lane builds one row in ~800 instructions, and an env has ~8 rows (Walker3D) to ~40
  // (Cassie): a warp building only its own env's rows runs the pass a quarter full.  The warps of a CTA are in lockstep
  // anyway (barrier per substep), so the CTA pools its rows: every warp publishes its env's row count, and warp w builds
  // the pooled rows 32 w .. 32 w + 31 -- of whichever envs they belong to, through those envs' WarpMem -- while the warps
  // without a share wait at the closing barrier and issue nothing.  Same arithmetic per row, fewer instructions per CTA.
  // EVERY warp of the CTA must call this once per substep (it contains two __syncthreads()).
  MB_HD static void coop_rows(Mem& S, const MbPhysics& P, int nlim, int nc, int ncs, int Rc) {
    __shared__ int s_cnt[32], s_par[32];
    const int lane = (int)(threadIdx.x & 31), W = (int)(blockDim.x >> 5);
    const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (lane == 0) { s_cnt[w] = Rc; s_par[w] = nlim | (nc << 8) | (ncs << 16); }
    __syncthreads();
    const int c = lane < W ? s_cnt[lane] : 0;
    int inc = c;  // inclusive prefix sum of the counts over the warps, one warp per lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    const int T = __shfl_sync(0xffffffffu, inc, 31);
    Mem* const S0 = &S - w;
#pragma unroll 1
    for (int base = 32 * w; base < T; base += 32 * W) {
      const int g = base + lane;
      int e = 0;  // owner of pooled row g: the number of warps whose rows end at or before it
#pragma unroll 1
      for (int k = 0; k < W; ++k) e += g >= __shfl_sync(0xffffffffu, inc, k) ? 1 : 0;
      e = e < W ? e : W - 1;
      const int first = __shfl_sync(0xffffffffu, inc - c, e);
      if (g < T) {
        const int par = s_par[e];
        row_build(S0[e], P, par & 255, (par >> 8) & 255, par >> 16, g - first);
      }
    }
    __syncthreads();
  }
#endif

  // second, short pass of the dual rows (loop closures, self-contacts): the two parts' partial sums -> one multiplier
  MB_HD static void rows_combine(Mem& S, const MbPhysics& P, int nlim, int nc, int ncs) {
    const int n0 = nlim + NLC;
    const int S0 = n0 + 3 * nc;
    const float inv_dt = 1.0f / P.dt;
    if (NLC > 0) {
      // fillMultiBodyConstraint: denominator = JA M^-1 JA^T + JB M^-1 JB^T (no coupling term, even for two links of
      // the same multibody), erp = m_erp, impulse bounds +-maxAppliedImpulse
      MB_LANES(l)
        if (l < NLC / 2) {
          const int ra = nlim + 2 * l, ax = l % 3, c = l / 3;
          const float dd = S.rc.r.r_par[ra].jinv + S.rc.r.r_par[ra + 1].jinv;
          const float jinv = dd > 1.1920929e-07f ? 1.0f / dd : 0.0f;
          const float rel_vel = S.rc.r.r_par[ra].rhs + S.rc.r.r_par[ra + 1].rhs;
          const float pos_error = -(S.lcP[2 * c][ax] - S.lcP[2 * c + 1][ax]);
          const float positional = -pos_error * P.erp_joint * inv_dt;
          MbRowPar par;
          par.rhs = (positional - rel_vel) * jinv; par.cfm = 0.0f; par.jinv = jinv; par.den = jinv != 0.0f ? dd : 0.0f;
          S.rc.r.r_par[ra] = par;
          S.rc.r.r_mu[ra] = M::lc_maximp(c);
          S.rc.r.r_mask[ra] |= MB_ROW_DUAL;
        }
      MB_END
    }
    if (NSELF > 0 && ncs > 0) {
      // fillMultiBodyConstraint: denominator = JA M^-1 JA^T + JB M^-1 JB^T (no coupling term, as for loop closures)
      MB_LANES(l)
        for (int i = l; i < 3 * ncs; i += 32) {
          const int sidx = i / 3, w = i - 3 * sidx;  // w: 0 normal, 1 / 2 friction directions
          const int ra = S0 + (w == 0 ? 2 * sidx : 2 * ncs + 4 * sidx + (w - 1));
          const int rb = ra + (w == 0 ? 1 : 2);
          const float dd = S.rc.r.r_par[ra].jinv + S.rc.r.r_par[rb].jinv;
          const float jinv = dd > 1.1920929e-07f ? 1.0f / dd : 0.0f;
          const float rel_vel = S.rc.r.r_par[ra].rhs + S.rc.r.r_par[rb].rhs;
          float positional = 0.0f, verr = -rel_vel;
          if (w == 0) {
            const float dist = S.cdist[nc + sidx] + P.linear_slop;
            if (dist > 0.0f) verr -= dist * inv_dt;
            else positional = -dist * S.cerp[nc + sidx] * inv_dt;
          }
          MbRowPar par;
          par.rhs = (positional + verr) * jinv; par.cfm = 0.0f; par.jinv = jinv; par.den = jinv != 0.0f ? dd : 0.0f;
          S.rc.r.r_par[ra] = par;
          if (w < 2) S.rc.r.r_mask[ra] |= MB_ROW_DUAL;  // friction pairs carry the flag on their first row
        }
      MB_END
    }
  }


  // ---- G3. contact warm starting (cold path: MbPhysics::warmstart > 0 only) --------------------------------------------
  // setupMultiBodyContactConstraint with SOLVER_USE_WARMSTARTING: appliedImpulse = factor * previous impulse of the point,
  // applied to the velocity before the first iteration (z += Y_r lambda_r); friction rows start from zero.
  MB_NOINLINE static void warm_start(Mem& S, float factor, int n0, int S0, int nc, int ncs, float* zout) {
    MB_ASSUME_SHARED(S);
    LaneVar<float> z;
    MB_LANES(l)
      z[l] = 0.0f;
    MB_END_REG
#pragma unroll 1
    for (int k = 0; k < nc + ncs; ++k) {
      const int ra = k < nc ? n0 + k : S0 + 2 * (k - nc);
      const float imp = factor * S.warm[mb_pid(S.cfoot[k])];
      if (imp == 0.0f) continue;
      const bool dual = k >= nc;
      const unsigned supA = S.rc.r.r_mask[ra] & MB_ROW_SUP, supB = dual ? S.rc.r.r_mask[ra + 1] : 0u;
      MB_LANES(l)
        const int tl = l < NU ? M::rowlen(l) - 1 : 0;
        if ((supA >> l) & 1u) z[l] = fmaf(S.w.Yc[ra][tl], imp, z[l]);
        if ((supB >> l) & 1u) z[l] = fmaf(S.w.Yc[ra + 1][tl], imp, z[l]);
        if (l == 0) S.rc.r.r_app[ra] = imp;
      MB_END
    }
    MB_LANES(l)
      zout[l] = z[l];
    MB_END
  }
  MB_NOINLINE static void warm_store(Mem& S, int n0, int S0, int nc, int ncs, bool solved) {
    MB_ASSUME_SHARED(S);
    MB_LANES(l)
      for (int i = l; i < MB_NWARM; i += 32) S.warm[i] = 0.0f;
    MB_END
    if (!solved) return;
#pragma unroll 1
    for (int k = 0; k < nc + ncs; ++k) {  // in contact order: a later contact of the same candidate overwrites
      const int ra = k < nc ? n0 + k : S0 + 2 * (k - nc);
      MB_LANES(l)
        if (l == 0) S.warm[mb_pid(S.cfoot[k])] = S.rc.r.r_app[ra];
      MB_END
    }
  }

  // ---- H. projected Gauss-Seidel in z-space (btMultiBodyConstraintSolver::solveSingleIteration order) --------
  // Row r of Y is stored over its support; the entry of coordinate l sits at slot tl(l) (prefix property).
  // DUAL: 0 = the row is one compact row (joint limits; contacts of a model without self-collision), 1 = always two
  // (loop closures), 2 = look at the row's MB_ROW_DUAL flag (contacts of a model with self-collision)
  template <int DUAL>
  MB_HD static float pgs_single(Mem& S, const LaneConst& C, int ra, float lo, float hi, LaneVar<float>& z, float& applied) {
    const unsigned supA = S.rc.r.r_mask[ra];
    LaneVar<float> ya, ta;
    MB_LANES(l)
      ya[l] = ((DUAL ? supA & MB_ROW_SUP : supA) & C.bit[l]) ? S.w.Yc[ra][C.tl[l]] : 0.0f;
    MB_END_REG
    if (DUAL == 1 || (DUAL == 2 && (supA & MB_ROW_DUAL))) {  // loop closure / self-contact: part on the other link
      const unsigned supB = S.rc.r.r_mask[ra + 1];
      MB_LANES(l)
        if (supB & C.bit[l]) ya[l] += S.w.Yc[ra + 1][C.tl[l]];
      MB_END_REG
    }
    MB_LANES(l)
      ta[l] = ya[l] * z[l];
    MB_END_REG
    const float dotA = warp_sum(ta);
    const MbRowPar pA = S.rc.r.r_par[ra];
    const float appA = S.rc.r.r_app[ra];
    float dA = pA.rhs - appA * pA.cfm - dotA * pA.jinv;
    const float sumA = appA + dA;
    // clamp to [lo, hi]; the delta is recomputed only when the clamp bites (same values as the if / else if chain of
    // btMultiBodyConstraintSolver::resolveSingleConstraintRowGeneric, five instructions instead of twelve)
    const float nA = fminf(fmaxf(sumA, lo), hi);
    if (nA != sumA) dA = nA - appA;
    MB_WARP_SYNC();  // every lane has read r_app[ra] before lane 0 replaces it (compute-sanitizer racecheck: WAR)
    MB_LANES(l)
      z[l] += ya[l] * dA;
      if (l == 0) S.rc.r.r_app[ra] = nA;
    MB_END
    applied = nA;
    return dA * pA.den;  // deltaImpulse * (1 / jacDiagABInv)
  }
  // friction pair with btMultiBodyConstraintSolver::resolveConeFrictionConstraintRows' projection;
  // sin/cos(atan2(a, b)) are written as a/|(a,b)|, b/|(a,b)|.  A self-contact's pair continues in rows ra + 2, ra + 3.
  template <bool SELF>
  MB_HD static float pgs_pair(Mem& S, const LaneConst& C, int ra, float cone, LaneVar<float>& z, bool& loaded) {
    const int rb = ra + 1;
    const unsigned supA = S.rc.r.r_mask[ra];  // both rows of a contact share the support
    LaneVar<float> ya, yb, ta, tb;
    MB_LANES(l)
      const bool in = ((SELF ? supA & MB_ROW_SUP : supA) & C.bit[l]) != 0u;
      ya[l] = in ? S.w.Yc[ra][C.tl[l]] : 0.0f;
      yb[l] = in ? S.w.Yc[rb][C.tl[l]] : 0.0f;
    MB_END_REG
    if (SELF && (supA & MB_ROW_DUAL)) {
      const unsigned supB = S.rc.r.r_mask[ra + 2];
      MB_LANES(l)
        if (supB & C.bit[l]) { ya[l] += S.w.Yc[ra + 2][C.tl[l]]; yb[l] += S.w.Yc[ra + 3][C.tl[l]]; }
      MB_END_REG
    }
    MB_LANES(l)
      ta[l] = ya[l] * z[l];
      tb[l] = yb[l] * z[l];
    MB_END_REG
    const float dotA = warp_sum(ta), dotB = warp_sum(tb);
    const MbRowPar pA = S.rc.r.r_par[ra], pB = S.rc.r.r_par[rb];
    const float appA = S.rc.r.r_app[ra], appB = S.rc.r.r_app[rb];
    float dA = pA.rhs - appA * pA.cfm - dotA * pA.jinv;
    float dB = pB.rhs - appB * pB.cfm - dotB * pB.jinv;
    const float sumA = appA + dA, sumB = appB + dB;
    float nA = sumA, nB = sumB;
    const float n2 = sumA * sumA + sumB * sumB;
    if (n2 >= cone * cone) {
      const float sc = n2 > 0.0f ? fabsf(cone) * rsqrtf(n2) : 0.0f;
      const float clipA = fabsf(sumA) * sc, clipB = fabsf(sumB) * sc;
      if (sumA < -clipA) { dA = -clipA - appA; nA = -clipA; }
      else if (sumA > clipA) { dA = clipA - appA; nA = clipA; }
      if (sumB < -clipB) { dB = -clipB - appB; nB = -clipB; }
      else if (sumB > clipB) { dB = clipB - appB; nB = clipB; }
    }
    MB_WARP_SYNC();
    MB_LANES(l)
      z[l] += ya[l] * dA + yb[l] * dB;
      if (l == 0) { S.rc.r.r_app[ra] = nA; S.rc.r.r_app[rb] = nB; }
    MB_END
    loaded = nA != 0.0f || nB != 0.0f;
    return dA * pA.den + dB * pB.den;
  }

  // btMultiBodyConstraintSolver::solveSingleIteration order: non-contact rows (limits, then loop closures;
  // alternating direction), normals, friction.  Contact k < nc is a static-world contact, nc <= k < nc + ncs a
  // self-contact (rows behind S0, see setup_rows); one loop serves both so that the row code exists once.
  // SELF: the model has self-collision pairs, so a contact row may be a dual row.  (Running substeps without
  // self-contacts through a SELF = false instantiation and keeping the SELF = true copy out of line was measured in
  // round 2: the extra call site cost Walker3D 8 % through register allocation, profiles/README.md r2l.)
  template <bool SELF>
  MB_HD static void solve_constraints(Mem& S, const MbPhysics& P, const LaneConst& C, int nlim, int nc, int ncs,
                                      LaneVar<float>& z) {
    const int nnc = nlim + NLC / 2, n0 = nlim + NLC, S0 = n0 + 3 * nc, nct = nc + (SELF ? ncs : 0);
    // A contact that carries no normal impulse has a zero friction cone: with nothing applied on its friction pair
    // either, the projection returns exactly zero for both rows (deltas 0, residual 0), so the visit is skipped with
    // identical results.  Which contacts are loaded is kept in two bit masks (normal impulse / friction impulses non-zero,
    // refreshed by the visits themselves), and the friction sweep walks the set bits: an idle contact costs nothing
    // (round 2; until then every contact paid 13 instructions per iteration to find out -- Cassie's speculative
    // hull-vertex contacts are mostly idle).
    unsigned nmask = 0u, fmask = 0u;
#pragma unroll 1
    for (int it = 0; it < P.iterations; ++it) {
      float res2 = 0.0f, app;
      // (alternating sweep direction over the non-contact rows, as solveSingleIteration does)
      int idx = (it & 1) ? 0 : nnc - 1;
      const int dir = (it & 1) ? 1 : -1;
#pragma unroll 1
      for (int v = 0; v < nnc; ++v, idx += dir) {
        float rr;
        if (NLC == 0 || idx < nlim) rr = pgs_single<0>(S, C, idx, 0.0f, P.limit_max_impulse, z, app);
        else {
          const int ra = nlim + 2 * (idx - nlim);
          const float lim = S.rc.r.r_mu[ra];
          rr = pgs_single<1>(S, C, ra, -lim, lim, z, app);
        }
        res2 = fmaxf(res2, rr * rr);
      }
#pragma unroll 1
      for (int k = 0; k < nct; ++k) {
        const int ra = (SELF && k >= nc) ? S0 + 2 * (k - nc) : n0 + k;
        const float rr = pgs_single<(SELF ? 2 : 0)>(S, C, ra, 0.0f, 1e10f, z, app);
        nmask = app != 0.0f ? nmask | (1u << k) : nmask & ~(1u << k);
        res2 = fmaxf(res2, rr * rr);
      }
#pragma unroll 1
      for (unsigned m = nmask | fmask; m != 0u; m &= m - 1u) {
        const int k = mb_ffs(m) - 1;
        const bool self = SELF && k >= nc;
        const int ra = self ? S0 + 2 * ncs + 4 * (k - nc) : n0 + nc + 2 * k;
        const int rn = self ? S0 + 2 * (k - nc) : n0 + k;
        const float cone = S.rc.r.r_mu[ra] * S.rc.r.r_app[rn];
        bool loaded;
        const float rr = pgs_pair<SELF>(S, C, ra, cone, z, loaded);
        fmask = loaded ? fmask | (1u << k) : fmask & ~(1u << k);
        res2 = fmaxf(res2, rr * rr);
      }
      if (res2 <= P.residual_threshold) break;
    }
  }
  // ---- I. integrate positions (btMultiBody::stepPositionsMultiDof) --------------------------------------------
  MB_HD static void integrate(Mem& S, const MbPhysics& P) {
    MB_LANES(l)
      if (l < NJ) S.q[l] += P.dt * S.u[6 + l];
      if (l == 31) {
        const float dt = P.dt;
        S.pos[0] += dt * S.u[3]; S.pos[1] += dt * S.u[4]; S.pos[2] += dt * S.u[5];
        const float wx = S.u[0], wy = S.u[1], wz = S.u[2];
        float ang = sqrtf(wx * wx + wy * wy + wz * wz);
        if (ang * dt > 0.25f * MB_PI_F) ang = 0.25f * MB_PI_F / dt;
        // half angle h <= pi/8: sin(h)/ang = (dt/2) sinc(h), cos(h) -- short Taylor polynomials (rel. err < 1e-8;
        // Bullet's own small-angle branch is the first two terms of the same series)
        const float h = 0.5f * ang * dt, h2 = h * h;
        const float f = 0.5f * dt * (1.0f + h2 * (-1.0f / 6.0f + h2 * (1.0f / 120.0f + h2 * (-1.0f / 5040.0f))));
        const float aw = 1.0f + h2 * (-0.5f + h2 * (1.0f / 24.0f + h2 * (-1.0f / 720.0f + h2 * (1.0f / 40320.0f))));
        const float ax = wx * f, ay = wy * f, az = wz * f;
        const float x = S.quat[0], y = S.quat[1], zq = S.quat[2], w = S.quat[3];
        float nx = aw * x + ax * w + ay * zq - az * y;
        float ny = aw * y - ax * zq + ay * w + az * x;
        float nz = aw * zq + ax * y - ay * x + az * w;
        float nw = aw * w - ax * x - ay * y - az * zq;
        const float n = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
        S.quat[0] = nx * n; S.quat[1] = ny * n; S.quat[2] = nz * n; S.quat[3] = nw * n;
      }
    MB_END
  }

  // ---- one Bullet substep.  Returns the number of constraint rows; contact list of this substep stays in S ----
  template <int OBST>
  // points_out (debug / golden-vector kernels only; nullptr in the env step kernels, where the block folds away):
  // [MB_MAXC][MB_POINT_WIDTH] = world position on the robot link, normal (on the partner, towards the link), distance,
  // applied normal impulse, link index (-1 = base), partner code (0 ground, 10 + box, 20 + bar, 1000 + self pair) --
  // the fields pybullet.getContactPoints reports ([5..9], [3], [2 / 4]); unused slots carry link = -2.
  MB_HD static int substep(Mem& S, const MbPhysics& P, LaneConst& C, int* nc_out, int* overflow, int sub = 0,
                           float* points_out = nullptr) {
    MB_BLOCK_BARRIER(sub);  // keeps the warps of a CTA in the same phase so instruction-cache lines are shared
    kinematics(S, P, C, true);
    int ns_all = 0;
    const int nc_all = collide<OBST>(S, P, overflow, &ns_all);  // static-world contacts, then ns_all self-contacts
    loop_pivots(S);
    bodies(S, P);
    mass_matrix_and_rhs(S);
    factorize<true>(S, C);  // also turns S.rhs into d^1/2 L^-T rhs
    // forward dynamics: udot = M^-1 (tau - bias); u += dt udot, clamped like btMultiBody::applyDeltaVeeMultiDof.
    // (The clamp is live in practice: Bullet ignores the MJCF armature, so the light arm links reach 100 rad/s
    // under full torque -- which is why the two forward substitutions of a substep cannot be merged into one.)
    LaneVar<float> x;
    MB_LANES(l)
      x[l] = l < NU ? S.rhs[l] : 0.0f;
    MB_END
    solve_L<true>(S, C, x);
    MB_LANES(l)
      if (l < NU) S.u[l] = fminf(fmaxf(S.u[l] + P.dt * x[l], -P.max_coord_vel), P.max_coord_vel);
    MB_END
    const int nlim = find_limits(S);
    int nc = nc_all - ns_all, ncs = ns_all;
    if (nlim + NLC + 3 * nc + 6 * ncs > MB_MAXROW) {  // self-contacts are dropped first
      const int avail = MB_MAXROW - nlim - NLC;
      if (3 * nc > avail) { nc = avail / 3; ncs = 0; }
      else ncs = (avail - 3 * nc) / 6;
      *overflow += 1;
    }
    const int R = nlim + NLC / 2 + 3 * (nc + ncs);  // as Bullet counts them (a loop / self-contact row is one row)
#if defined(__CUDACC__) && defined(MB_COOP_ROWS) && MB_COOP_ROWS
    coop_rows(S, P, nlim, nc, ncs, R > 0 ? compact_rows(nlim, nc, ncs) : 0);
#endif
    if (R > 0) {
      // (a size-optimised rolled version of setup_rows serving every row kind was measured too: 12.7 KB less hot
      // code, but 9 % slower on Walker3D and 11 % on Cassie -- the unrolled register version stays)
#if defined(__CUDACC__) && defined(MB_COOP_ROWS) && MB_COOP_ROWS
      rows_combine(S, P, nlim, nc, ncs);
#else
      setup_rows(S, P, nlim, nc, ncs);
#endif
      // (an impulse-space variant of the PGS -- Gram matrix G = Y Y^T of the rows in the unused tail of the row matrix,
      // one register w_c = sum_s G_cs lambda_s per lane, a row visit = one shuffle + uniform delta + one LDS / FFMA per
      // lane -- was measured in round 2: 26 instead of 31 instructions per visit, but building G and assembling z cost
      // what the visits saved: -3 % on every Walker-family env (profiles/README.md, r2d); the z-space solver stays)
      LaneVar<float> z;
      MB_LANES(l)
        z[l] = 0.0f;
      MB_END
      // (an exact skip of idle speculative rows -- |Y_r . z| <= |Y_r| |z| with a running bound of |z| -- was measured in
      // round 2: the bound is too loose to fire often and its bookkeeping cost 6 % (Walker3D) to 11 % (Cassie): dropped)
      if (MB_UNLIKELY(P.warmstart > 0.0f)) {  // (a kernel parameter: the test costs the default path one predicate)
        warm_start(S, P.warmstart, nlim + NLC, nlim + NLC + 3 * nc, nc, ncs, S.rhs);  // (S.rhs is free after the FD solve)
        init_lane_const(C);
        MB_LANES(l)
          z[l] = S.rhs[l];
        MB_END
      }
      // (hull models -- Cassie, 34 rows per substep -- run substeps without self-contacts through the SELF = false
      // instantiation: +3.6 %; for the segment-pair models the second inline copy costs 2 % of instruction cache, r2n)
      if (M::SELF_HULLS && ncs == 0) solve_constraints<false>(S, P, C, nlim, nc, 0, z);
      else solve_constraints<(NSELF > 0)>(S, P, C, nlim, nc, ncs, z);
      solve_L<false>(S, C, z);
      MB_LANES(l)
        if (l < NU) S.u[l] = fminf(fmaxf(S.u[l] + z[l], -P.max_coord_vel), P.max_coord_vel);
      MB_END
    }
    if (MB_UNLIKELY(P.warmstart > 0.0f)) {
      warm_store(S, nlim + NLC, nlim + NLC + 3 * nc, nc, ncs, R > 0);
      init_lane_const(C);
    }
    if (points_out) {
      const int n0 = nlim + NLC, S0 = n0 + 3 * nc;
      MB_LANES(l)
        if (l < MB_MAXC) {
          float* o = points_out + MB_POINT_WIDTH * l;
          const bool live = l < nc + ncs;  // (contacts dropped at the row cap carry no row)
          const int row = l < nc ? n0 + l : S0 + 2 * (l - nc);
          o[0] = live ? S.cP[l][0] + S.pos[0] : 0.0f; o[1] = live ? S.cP[l][1] + S.pos[1] : 0.0f;
          o[2] = live ? S.cP[l][2] + S.pos[2] : 0.0f;
          o[3] = live ? S.cn[l][0] : 0.0f; o[4] = live ? S.cn[l][1] : 0.0f; o[5] = live ? S.cn[l][2] : 0.0f;
          o[6] = live ? S.cdist[l] : 0.0f;
          o[7] = live && R > 0 ? S.rc.r.r_app[row] : 0.0f;
          o[8] = live ? (float)S.clink[l] : -2.0f;
          o[9] = live ? (float)S.cpartner[l] : 0.0f;
        }
      MB_END
    }
    integrate(S, P);
    *nc_out = nc_all < MB_MAXC ? nc_all : MB_MAXC;
    return R;
  }
};
