/*
 * mocca_oracle.h -- CPU float64 restatement of the mocca_envs hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY, two layers:
 *   - ENV LAYER (everything the reference's Python computes: apply_action, calc_state, reset draws, targets, terrain
 *     and bar generators, foot / palm contact logic, rewards, termination, Cassie's PD loop): PINNED against the
 *     reference's own code.  tools/gen_reference_golden.py imports the unmodified env_base.py / env_locomotion.py /
 *     robots.py / bullet_utils.py / bullet_objects.py / env_cassie.py with stand-ins for gym and pybullet whose Bullet
 *     client is served by this file's physics, and records traces (tests/golden/ref_*.npz); replaying the recorded
 *     actions through the orc_*_env functions reproduces observation / reward / done to float64 rounding (1e-15;
 *     Cassie bit-identical) for all six env classes (tests/test_reference_golden.py).
 *   - BULLET ARITHMETIC: PARITY UNPINNED.  It lives in the third-party `pybullet` C-extension (reference setup.py:11,
 *     un-pinned, not vendored, not installable here), reached through `stepSimulation` (reference
 *     mocca_envs/bullet_utils.py:352-353); the reference ships no tests or golden vectors.  This file
 *     restates Bullet's published multibody pipeline (btMultiBody ABA, btMultiBodyConstraintSolver PGS, semi-implicit
 *     Euler).  That part is pinned only against: NumPy RandomState streams (bit-exact), an
 *     independent Jacobian-based mass matrix, ID(FD(tau)) round trips and conservation laws (tests/test_oracle_*.py)
 *     -- NOT against PyBullet outputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 */
#ifndef MOCCA_ORACLE_H
#define MOCCA_ORACLE_H
#include <stdint.h>

#define ORC_MAXL 40   /* links (without base) */
#define ORC_MAXD 40   /* joint dofs */
#define ORC_MAXG 192  /* geoms (Cassie: 158 hull support vertices) */
#define ORC_MAXP 96   /* contact points kept per substep */
#define ORC_MAXW (2 * ORC_MAXG) /* warm-start slots, indexed by candidate id = 2 * geom + end */
/* per-env persistent contact state: [0, ORC_MAXW) the impulses above; behind them the ground-plane manifolds of the
 * persistent_manifold switch, 4 slots x {candidate id + 1 (0 = empty), cached point on the plane x, y, pad} per link */
#define ORC_WARMSZ (ORC_MAXW + 16 * (ORC_MAXL + 1))
#define ORC_MAXROW (3 * ORC_MAXP + 2 * ORC_MAXD)
#define ORC_MAXU (6 + ORC_MAXD)
#define ORC_MAXSP 256 /* self-collision candidate pairs */
#define ORC_MAXHULL 24 /* links with a mesh hull */
#define ORC_HULLV 32   /* vertices per hull */
#define ORC_MAXHPAIR 64
#define ORC_MAXSTEPS 32 /* stepping stones / bars in a terrain table */

enum { ORC_GEOM_SPHERE = 0, ORC_GEOM_CAPSULE = 1, ORC_GEOM_BOX = 2 };
enum { ORC_JOINT_FIXED = 0, ORC_JOINT_REVOLUTE = 1 };

typedef struct {
  int n_links, n_dof, n_geoms, n_feet;
  int parent[ORC_MAXL];      /* -1 = base */
  int joint_type[ORC_MAXL];
  int dof_of_link[ORC_MAXL]; /* -1 for fixed */
  int link_of_dof[ORC_MAXD];
  double axis[ORC_MAXL][3];     /* joint axis, link frame */
  double rot_p2t[ORC_MAXL][4];  /* xyzw, parent->this at q=0 (Bullet zeroRotParentToThis) */
  double e_vec[ORC_MAXL][3];    /* parent COM -> pivot, parent frame */
  double d_vec[ORC_MAXL][3];    /* pivot -> COM, this frame */
  double mass[ORC_MAXL];
  double inertia[ORC_MAXL][3];
  double base_mass;
  double base_inertia[3];
  double lower[ORC_MAXD], upper[ORC_MAXD], damping[ORC_MAXD], armature[ORC_MAXD], gain[ORC_MAXD];
  double link_thresh[ORC_MAXL + 1]; /* contact breaking threshold; [0] = base, [i+1] = link i */
  int link_group[ORC_MAXL + 1], link_mask[ORC_MAXL + 1];
  int geom_link[ORC_MAXG];  /* -1 = base */
  int geom_type[ORC_MAXG];
  double geom_p0[ORC_MAXG][3], geom_p1[ORC_MAXG][3]; /* sphere centre / capsule ends, link frame */
  double geom_quat[ORC_MAXG][4];
  double geom_size[ORC_MAXG][3];
  double geom_friction[ORC_MAXG];
  int foot_link[4];
  double base_joint_angles[ORC_MAXD];
  double base_position[3];
  double base_orientation[4];       /* xyzw; set_base_pose (robots.py:276,311-323) */
  double termination_height;        /* env class attribute (env_locomotion.py:44,320) */
  double planar_env;                /* != 0: Walker2DCustomEnv / Crab2DCustomEnv (env_locomotion.py:285-314): done forced
                                       to False, reset() returns zeros in the target slots */
  double stepper_init_position[3];  /* robot_init_position of the stepper env (env_locomotion.py:339,845) */
  int n_right, right_idx[ORC_MAXD], left_idx[ORC_MAXD]; /* mirroring tables robots.py:282-290 */
  int n_neg, neg_idx[8];
  int palm_link[2]; /* Monkey3D: right_palm, left_palm (env_locomotion.py:1269,1424); -1 otherwise */
  /* loop closures: btMultiBodyPoint2Point between two links of the robot (Cassie, env_cassie.py:114-137);
   * pivots in the links' inertial frames */
  int n_p2p, p2p_link_a[2], p2p_link_b[2];
  double p2p_pivot_a[2][3], p2p_pivot_b[2][3], p2p_max_impulse[2];
  /* Cassie bookkeeping (env_cassie.py:59-60,192-202): dofs of the 14 ordered joints, PD joint list, gains */
  /* self-collision candidate geom pairs (robots.py:259-264; compiled by model_compiler.self_collision_pairs) */
  int n_self, self_a[ORC_MAXSP], self_b[ORC_MAXSP];
  /* mesh-hull self-collision (Cassie, env_cassie.py:81-85): per link hull ORC_HULLV support vertices in the link's
   * inertial frame (urdf_compiler.hull_fan_vertices), candidate hull pairs, collision margin of a hull */
  int n_hulls, hull_link[ORC_MAXHULL];
  double hull_verts[ORC_MAXHULL][ORC_HULLV][3];
  double hull_center[ORC_MAXHULL][3], hull_radius[ORC_MAXHULL], hull_friction[ORC_MAXHULL];
  int n_hpairs, hpair_a[ORC_MAXHPAIR], hpair_b[ORC_MAXHPAIR];
  double hull_margin;
  int n_ordered, ordered_dof[ORC_MAXD];
  int n_pd, pd_ordered_index[16]; /* powered + spring joints, as indices into the ordered joints */
  double pd_kp[16], pd_kd[16];
} orc_model;

typedef struct {
  double gravity;        /* 9.8   env_base.py:80 */
  double dt;             /* 1/240 env_base.py:81 */
  int substeps;          /* 4     bullet_utils.py:349 */
  int iterations;        /* 5     bullet_utils.py:340 */
  double erp_contact;    /* 0.9   bullet_utils.py:345 (m_erp2) */
  double erp_joint;      /* 0.2   Bullet m_erp */
  double linear_slop;    /* 1e-5  PyBullet createEmptyDynamicsWorld */
  double lin_damping;    /* 0.04  btMultiBody m_linearDamping */
  double ang_damping;    /* 0.04 */
  double max_coord_vel;  /* 100   btMultiBody m_maxCoordinateVelocity */
  double warmstart;      /* 0.1   PyBullet m_warmstartingFactor; 0 disables */
  double limit_max_impulse; /* 100 btMultiBodyConstraint m_maxAppliedImpulse */
  double split_threshold;   /* -0.04 m_splitImpulsePenetrationThreshold (limit rows) */
  double residual_threshold; /* 1e-7 m_leastSquaresResidualThreshold */
  int limit_rows_always;     /* 0: rows only when the limit is violated (Bullet >= 2.88) */
  int gyro;                  /* 1 */
  int has_ground;            /* 1: infinite plane z=0 (plane_stadium.sdf) */
  double ground_friction;    /* 0.8 bullet_utils.py:371 */
  int self_collision;        /* 1: URDF_USE_SELF_COLLISION | ..._EXCLUDE_ALL_PARENTS (robots.py:259-264) */
  int persistent_manifold;   /* 0 (default): every sphere / capsule end within the breaking threshold is a contact, as the
                                CUDA kernel has it.  1: btPersistentManifold semantics against the ground plane (SURVEY App.
                                B.2, OQ8): per link <= 4 cached points, refreshed every substep (dropped beyond the breaking
                                threshold or after drifting more than it along the plane), and each geom reports only its
                                deepest point per substep.  Oracle-only hypothesis switch (tools/switch_deltas.py). */
} orc_params;

typedef struct {
  double pos[3];   /* base COM, world */
  double quat[4];  /* xyzw, base orientation local->world (as PyBullet reports it) */
  double omega[3]; /* world */
  double vel[3];   /* world, base COM */
  double q[ORC_MAXD];
  double qd[ORC_MAXD];
} orc_state;

typedef struct {
  int n;                  /* number of contact points this substep */
  int point_id[ORC_MAXP]; /* candidate id: geom*2 + end */
  int link[ORC_MAXP];     /* -1 = base */
  int partner[ORC_MAXP];  /* 0 = ground plane, 10+k = box k, 20+k = bar k, 1000 + (link_b + 1) = robot link */
  int link_b[ORC_MAXP];   /* self-contact: the other robot link (-1 = base); -2 = static partner */
  double pos_a[ORC_MAXP][3];
  double pos_b[ORC_MAXP][3]; /* self-contact: contact point on link_b */
  double normal[ORC_MAXP][3]; /* on B, pointing towards A */
  double dist[ORC_MAXP];
  double friction[ORC_MAXP];
  double erp[ORC_MAXP], cfm[ORC_MAXP]; /* per-contact (soft contacts on planks) */
  double impulse[ORC_MAXP];            /* applied normal impulse after the solve */
} orc_contacts;

/* static box obstacle (plank base / cover), world frame */
typedef struct {
  double center[3];
  double R[3][3]; /* box axes as columns */
  double half[3];
  double friction;
  double stiffness, damping; /* <=0: rigid */
  int id;                    /* partner code reported in orc_contacts */
  int cylinder;              /* 1: capped cylinder about the local z axis, radius half[0], half length half[2] (Pillar) */
} orc_box;

/* static thin cylinder (MonkeyBar, bullet_objects.py:148-187), treated as a capsule around its axis segment */
typedef struct {
  double center[3];
  double axis[3]; /* unit */
  double halflen, radius, friction;
  int id;
} orc_bar;

typedef struct {
  uint32_t mt[624];
  int pos;
} orc_rng;

/* ---- dynamics primitives (tests) ---- */
void orc_default_params(orc_params* p);
void orc_fk(const orc_model* m, const orc_state* s, double link_pos[][3], double link_rot[][9]);
void orc_forward_dynamics(const orc_model* m, const orc_params* p, const orc_state* s, const double* tau,
                          int with_damping, double* acc /* [6+n] */);
void orc_rnea(const orc_model* m, const orc_state* s, const double* acc, double gravity, double* tau /* [6+n] */);
void orc_mass_matrix(const orc_model* m, const orc_state* s, double* M /* [(6+n)^2] row-major */);
void orc_minv_mult(const orc_model* m, const orc_params* p, const orc_state* s, const double* f, double* out);
int orc_collide(const orc_model* m, const orc_params* p, const orc_state* s, const orc_box* boxes, int n_boxes,
                orc_contacts* c);
void orc_step_physics_bars(const orc_model* m, const orc_params* p, orc_state* s, const double* tau_applied,
                           const orc_bar* bars, int n_bars, orc_contacts* last_contacts, int* rows_sum);
void orc_substep(const orc_model* m, const orc_params* p, orc_state* s, const double* tau, const orc_box* boxes,
                 int n_boxes, double* warm /* [ORC_MAXW], or [ORC_WARMSZ] with persistent_manifold */, orc_contacts* out_contacts, int* out_rows);
void orc_step_physics(const orc_model* m, const orc_params* p, orc_state* s, const double* tau_applied,
                      const orc_box* boxes, int n_boxes, double* warm, orc_contacts* last_contacts, int* rows_sum);
void orc_energy_momentum(const orc_model* m, const orc_state* s, double gravity, double* out /* KE,PE,P[3],L[3] */);

/* ---- RNG (NumPy legacy RandomState compatible) ---- */
void orc_rng_seed_array(orc_rng* r, const uint32_t* key, int len);
uint32_t orc_rng_u32(orc_rng* r);
double orc_rng_double(orc_rng* r);
double orc_rng_uniform(orc_rng* r, double lo, double hi);

/* ---- Walker3DCustomEnv (env_locomotion.py:37-282) ---- */
typedef struct {
  orc_state s;
  double warm[ORC_WARMSZ];
  /* robot (robots.py:13-227) */
  double feet_contact[4];
  double feet_xyz[4][3];
  double body_xyz[3], body_rpy[3], body_vel[3];
  double joint_speeds[ORC_MAXD];
  int joints_at_limit;
  int mirrored;
  double robot_state[6 + 2 * ORC_MAXD + 4];
  /* env */
  double dist, angle, stop_frames;
  double walk_target[3];
  int close_count;
  double linear_potential, angular_potential, distance_to_target, angle_to_target;
  int done;
  int eval_mode;
  int elapsed;  /* gym TimeLimit counter (__init__.py:55) */
  double progress, posture_penalty, energy_penalty, joints_penalty, tall_bonus, target_bonus;
  orc_rng env_rng, robot_rng;
  int rng_aliased; /* env_base.py:93 quirk Q1: robot draws from the env stream until seed() rebinds */
  double rows_sum; /* diagnostics: constraint rows over the last step */
  orc_contacts last_contacts;
} orc_w3d_env;

void orc_w3d_seed(orc_w3d_env* e, const uint32_t* key, int len, int at_construction);
void orc_w3d_reset(const orc_model* m, const orc_params* p, orc_w3d_env* e, double* obs /* [52] */);
void orc_w3d_step(const orc_model* m, const orc_params* p, orc_w3d_env* e, const double* action, double* obs,
                  double* reward, int* done, int* truncated);
/* batched convenience for the CPU baseline: n independent envs, OpenMP over envs, auto-reset on done */
void orc_w3d_step_batch(const orc_model* m, const orc_params* p, orc_w3d_env* envs, int n, const double* actions,
                        double* obs, double* rewards, int* dones, int n_threads);
/* ---- Walker3DStepperEnv (env_locomotion.py:330-840; planks bullet_objects.py:47-103) ---- */
#define ORC_NSTEPS 20
typedef struct {
  orc_w3d_env base; /* robot + physics state + RNG streams (walk_target / potentials reused) */
  int curriculum;   /* 0..9, set through set_env_params (env_locomotion.py:362-369) */
  int gain_curriculum; /* curriculum latched for robot.applied_gain at reset (env_locomotion.py:489) */
  double terrain[ORC_NSTEPS][6]; /* x y z phi x_tilt y_tilt (env_locomotion.py:395-441) */
  int plank_index[3];            /* which terrain row each of the 3 physical planks shows */
  orc_box boxes[6];              /* plank p: boxes[2p] = base, boxes[2p+1] = cover */
  int next_step_index, target_reached_count, stop_on_next_step, set_stop_on_next_step, timestep, target_reached;
  double foot_dist_to_target[2];
  double targets[3][5];
  double step_bonus, speed_penalty;
  int steps_reached; /* info["steps_reached"] when reported, else -1 */
  int random_reward; /* constructor kwarg (env_locomotion.py:355) */
  int plank_class;   /* constructor kwarg (env_locomotion.py:342,356-357): 0 LargePlank, 1 Plank, 2 Pillar */
} orc_stepper_env;

void orc_stepper_seed(orc_stepper_env* e, const uint32_t* key, int len, int at_construction);
void orc_stepper_reset(const orc_model* m, const orc_params* p, orc_stepper_env* e, double* obs /* [65] */);
void orc_stepper_step(const orc_model* m, const orc_params* p, orc_stepper_env* e, const double* action,
                      double* obs, double* reward, int* done, int* truncated);
void orc_stepper_step_batch(const orc_model* m, const orc_params* p, orc_stepper_env* envs, int n,
                            const double* actions, double* obs, double* rewards, int* dones, int n_threads);
/* ---- Monkey3DCustomEnv (env_locomotion.py:1136-1516) ---- */
#define ORC_NBARS 32
typedef struct {
  orc_w3d_env base;
  double terrain[ORC_NBARS][4]; /* x y z phi */
  orc_bar bars[4];
  int bar_index[4];
  int next_step_index, target_reached_count, free_fall_count, timestep, swing_leg, pivot_leg, target_reached;
  double foot_dist_to_target, swing_potential;
  double targets[2][3];
  double palm_xyz[2][3], palm_quat[2][4];
  double step_bonus;
} orc_monkey_env;

void orc_monkey_seed(orc_monkey_env* e, const uint32_t* key, int len, int at_construction);
void orc_monkey_reset(const orc_model* m, const orc_params* p, orc_monkey_env* e, double* obs /* [69] */);
void orc_monkey_step(const orc_model* m, const orc_params* p, orc_monkey_env* e, const double* action, double* obs,
                     double* reward, int* done, int* truncated);
void orc_monkey_step_batch(const orc_model* m, const orc_params* p, orc_monkey_env* envs, int n,
                           const double* actions, double* obs, double* rewards, int* dones, int n_threads);
int orc_sizeof_monkey_env(void);
int orc_diag_q_take(double* out); /* test diagnostics: joint angles at the start of the last <= 64 substeps */
long orc_diag_contacts_take(void); /* test diagnostics: contact points over the substeps since the last call */
/* ---- CassieEnv (env_cassie.py:285-479) ---- */
typedef struct {
  orc_w3d_env base;
  double jvel[16];       /* low-pass joint velocity of the ordered joints (env_cassie.py:319,451-453) */
  double rad_angles[16]; /* robot.rad_joint_angles of the last calc_state */
  double speeds[16];     /* robot.joint_speeds (raw rad/s) */
  double potential, initial_z;
  double alive_rew, progress_rew;
  double robot_state[40];
} orc_cassie_env;
void orc_cassie_params(orc_params* p); /* dt = 0.03 / 50, one substep per stepSimulation */
void orc_cassie_reset(const orc_model* m, const orc_params* p, orc_cassie_env* e, double* obs /* [36] */);
void orc_cassie_step(const orc_model* m, const orc_params* p, orc_cassie_env* e, const double* action /* [10] */,
                     double* obs, double* reward, int* done, int* truncated);
void orc_cassie_step_batch(const orc_model* m, const orc_params* p, orc_cassie_env* envs, int n,
                           const double* actions, double* obs, double* rewards, int* dones, int n_threads);
int orc_sizeof_cassie_env(void);
int orc_sizeof_stepper_env(void);
int orc_sizeof_w3d_env(void);
int orc_sizeof_model(void);

#endif
