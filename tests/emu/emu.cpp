// Lane-loop emulation of the kernel source (mb_core.cuh / mb_env.cuh compiled by g++ with MB_LANES as a
// 32-iteration loop).  DEBUG/TEST HARNESS ONLY: lets the exact kernel source be diffed against the CPU oracle on
// a box without a GPU.  Never loaded by the mocca_envs_b200 package; the product path is libmocca_b200.so.
#include <string.h>

#include "../../mocca_envs_b200/csrc/generated/walker3d_model.h"
#include "../../mocca_envs_b200/csrc/generated/monkey3d_model.h"
#include "../../mocca_envs_b200/csrc/generated/cassie_model.h"
#include "../../mocca_envs_b200/csrc/generated/child3d_model.h"
#include "../../mocca_envs_b200/csrc/generated/mike_model.h"
#include "../../mocca_envs_b200/csrc/generated/walker2d_model.h"
#include "../../mocca_envs_b200/csrc/generated/crab2d_model.h"
#include "../../mocca_envs_b200/csrc/mb_env.cuh"

typedef W3D_Model WM;
typedef W3DEnv<WM> WEnv;
typedef StepperEnv<WM> SEnv;
typedef WarpMem<WM> WMem;
typedef MK3D_Model MM;
typedef MonkeyEnv<MM> MEnv;
typedef WarpMem<MM> MMem;
typedef CAS_Model CM;
typedef CassieEnv<CM> CEnv;
typedef WarpMem<CM> CMem;

static void default_phys(MbPhysics* p) {
  p->dt = 1.0f / 240.0f; p->substeps = 4; p->iterations = 5; p->gravity = 9.8f; p->erp_contact = 0.9f;
  p->erp_joint = 0.2f; p->linear_slop = 1e-5f; p->lin_damping = 0.04f; p->ang_damping = 0.04f;
  p->max_coord_vel = 100.0f; p->limit_max_impulse = 100.0f; p->split_threshold = -0.04f;
  p->residual_threshold = 1e-7f; p->ground_friction = 0.8f; p->has_ground = 1;
  p->box_friction = 1.0f; p->box_erp = 0.9f; p->box_cfm = 0.0f; p->bar_friction = 0.5f;
  p->self_collision = 1;
  p->warmstart = 0.0f;
}

extern "C" {
int emu_sizeof_phys() { return (int)sizeof(MbPhysics); }
int emu_sizeof_warpmem() { return (int)sizeof(WMem); }
void emu_default_phys(MbPhysics* p) { default_phys(p); }

void emu_step_physics(const MbPhysics* p, float* state, const float* tau, int* rows, int* contacts) {
  static WMem S;
  memset(&S, 0, sizeof(S));
  WEnv::load_state(S, state);
  for (int j = 0; j < WM::NJ; ++j) S.tau[j] = tau[j];
  int r = 0, nc = 0, ov = 0;
  Sim<WM>::LaneConst C;
  Sim<WM>::init_lane_const(C);
  for (int k = 0; k < p->substeps; ++k) r += Sim<WM>::substep<0>(S, *p, C, &nc, &ov);
  WEnv::store_state(S, state);
  *rows = r;
  *contacts = nc;
}

// stepSimulation with the warm-start switch: `warm` [MB_NWARM] plays the env's HBM impulse array
void emu_step_physics_warm(const MbPhysics* p, float* state, const float* tau, float* warm, int* rows, int* contacts) {
  static WMem S;
  memset(&S, 0, sizeof(S));
  S.warm = warm;
  WEnv::load_state(S, state);
  for (int j = 0; j < WM::NJ; ++j) S.tau[j] = tau[j];
  int r = 0, nc = 0, ov = 0;
  Sim<WM>::LaneConst C;
  Sim<WM>::init_lane_const(C);
  for (int k = 0; k < p->substeps; ++k) r += Sim<WM>::substep<0>(S, *p, C, &nc, &ov);
  WEnv::store_state(S, state);
  *rows = r;
  *contacts = nc;
}

void emu_mass_matrix(const MbPhysics* p, const float* state, float* Mout, float* bias) {
  static WMem S;
  memset(&S, 0, sizeof(S));
  WEnv::load_state(S, state);
  Sim<WM>::LaneConst C;
  Sim<WM>::init_lane_const(C);
  Sim<WM>::kinematics(S, *p, C, true);
  Sim<WM>::bodies(S, *p);
  Sim<WM>::mass_matrix_and_rhs(S);
  const int NU = WM::NU;
  for (int i = 0; i < NU; ++i) {
    for (int j = 0; j < NU; ++j) Mout[i * NU + j] = mb_Lget<WM>(S.L, i, j);
    bias[i] = -S.rhs[i];
  }
}

void emu_w3d_reset(const MbPhysics* p, float* state, float* rec, uint32_t* mt_env, uint32_t* mt_robot, float* obs) {
  static WMem S;
  memset(&S, 0, sizeof(S));
  WEnv::reset(S, *p, rec, mt_env, mt_robot, obs);
  WEnv::store_state(S, state);
}

void emu_w3d_step(const MbPhysics* p, float* state, float* rec, uint32_t* mt_env, uint32_t* mt_robot,
                  const float* act, float* obs, float* rew, uint8_t* done, uint8_t* trunc, float* final_obs,
                  double* stats_out) {
  static WMem S;
  memset(&S, 0, sizeof(S));
  MbStats st;
  memset(&st, 0, sizeof(st));
  WEnv::step(S, *p, state, rec, mt_env, mt_robot, act, obs, rew, done, trunc, final_obs, &st);
  stats_out[0] = (double)st.episodes; stats_out[1] = st.ret_sum; stats_out[2] = st.len_sum;
  stats_out[3] = (double)st.nonfinite;
}

// ---- Walker3DStepperEnv
void emu_stepper_phys(MbPhysics* p) {
  default_phys(p);
  p->has_ground = 0;
  const float kp = 30000.0f, kd = 1000.0f + 0.1f, denom = p->dt * kp + kd;
  p->box_friction = 1.0f; p->box_erp = p->dt * kp / denom; p->box_cfm = 1.0f / denom;
}
int emu_stepper_rec_stride() { return (int)SEnv::REC_STRIDE; }

void emu_stepper_reset(const MbPhysics* p, float* state, float* rec, uint32_t* mt_env, uint32_t* mt_robot, float* obs) {
  static WMem S;
  memset(&S, 0, sizeof(S));
  SEnv::reset(S, *p, rec, mt_env, mt_robot, obs);
  WEnv::store_state(S, state);
}

void emu_stepper_step(const MbPhysics* p, float* state, float* rec, uint32_t* mt_env, uint32_t* mt_robot,
                      const float* act, float* obs, float* rew, uint8_t* done, uint8_t* trunc, float* final_obs,
                      double* stats_out) {
  static WMem S;
  memset(&S, 0, sizeof(S));
  MbStats st;
  memset(&st, 0, sizeof(st));
  SEnv::step(S, *p, state, rec, mt_env, mt_robot, act, obs, rew, done, trunc, final_obs, &st);
  stats_out[0] = (double)st.episodes; stats_out[1] = st.ret_sum; stats_out[2] = st.len_sum;
  stats_out[3] = (double)st.nonfinite;
}

// stepSimulation with the planks described by a Stepper record
void emu_stepper_step_physics(const MbPhysics* p, float* state, const float* rec, const float* tau, int* rows,
                              int* contacts) {
  static WMem S;
  memset(&S, 0, sizeof(S));
  WEnv::load_state(S, state);
  for (int j = 0; j < WM::NJ; ++j) S.tau[j] = tau[j];
  int r = 0, nc = 0, ov = 0;
  Sim<WM>::LaneConst C;
  Sim<WM>::init_lane_const(C);
  for (int k = 0; k < p->substeps; ++k) {
    SEnv::load_obstacles(S, rec);
    r += Sim<WM>::substep<MB_OBST_BOXES>(S, *p, C, &nc, &ov);
  }
  WEnv::store_state(S, state);
  *rows = r;
  *contacts = nc;
}

// ---- Monkey3DCustomEnv
int emu_monkey_rec_stride() { return (int)MEnv::REC_STRIDE; }
int emu_sizeof_monkey_warpmem() { return (int)sizeof(MMem); }

void emu_monkey_reset(const MbPhysics* p, float* state, float* rec, uint32_t* mt_env, uint32_t* mt_robot, float* obs) {
  static MMem S;
  memset(&S, 0, sizeof(S));
  MEnv::reset(S, *p, rec, mt_env, mt_robot, obs);
  MEnv::store_state(S, state);
}

void emu_monkey_step(const MbPhysics* p, float* state, float* rec, uint32_t* mt_env, uint32_t* mt_robot,
                     const float* act, float* obs, float* rew, uint8_t* done, uint8_t* trunc, float* final_obs,
                     double* stats_out) {
  static MMem S;
  memset(&S, 0, sizeof(S));
  MbStats st;
  memset(&st, 0, sizeof(st));
  MEnv::step(S, *p, state, rec, mt_env, mt_robot, act, obs, rew, done, trunc, final_obs, &st);
  stats_out[0] = (double)st.episodes; stats_out[1] = st.ret_sum; stats_out[2] = st.len_sum;
  stats_out[3] = (double)st.nonfinite;
}

// stepSimulation with the bars described by a Monkey record
void emu_monkey_step_physics(const MbPhysics* p, float* state, const float* rec, const float* tau, int* rows,
                             int* contacts) {
  static MMem S;
  memset(&S, 0, sizeof(S));
  MEnv::load_state(S, state);
  for (int j = 0; j < MM::NJ; ++j) S.tau[j] = tau[j];
  int r = 0, nc = 0, ov = 0;
  Sim<MM>::LaneConst C;
  Sim<MM>::init_lane_const(C);
  for (int k = 0; k < p->substeps; ++k) {
    MEnv::load_obstacles(S, rec);
    r += Sim<MM>::substep<MB_OBST_BARS>(S, *p, C, &nc, &ov);
  }
  MEnv::store_state(S, state);
  *rows = r;
  *contacts = nc;
}

void emu_monkey_mass_matrix(const MbPhysics* p, const float* state, float* Mout, float* bias) {
  static MMem S;
  memset(&S, 0, sizeof(S));
  MEnv::load_state(S, state);
  Sim<MM>::LaneConst C;
  Sim<MM>::init_lane_const(C);
  Sim<MM>::kinematics(S, *p, C, true);
  Sim<MM>::bodies(S, *p);
  Sim<MM>::mass_matrix_and_rhs(S);
  const int NU = MM::NU;
  for (int i = 0; i < NU; ++i) {
    for (int j = 0; j < NU; ++j) Mout[i * NU + j] = mb_Lget<MM>(S.L, i, j);
    bias[i] = -S.rhs[i];
  }
}

// ---- CassieEnv
void emu_cassie_phys(MbPhysics* p) {
  default_phys(p);
  p->dt = 0.03f / 50.0f;
  p->substeps = 1;
}
int emu_cassie_rec_stride() { return (int)CEnv::REC_STRIDE; }
int emu_sizeof_cassie_warpmem() { return (int)sizeof(CMem); }

void emu_cassie_reset(const MbPhysics* p, float* state, float* rec, float* obs) {
  static CMem S;
  memset(&S, 0, sizeof(S));
  CEnv::reset(S, *p, rec, nullptr, nullptr, obs);
  CEnv::store_state(S, state);
}

void emu_cassie_step(const MbPhysics* p, float* state, float* rec, const float* act, float* obs, float* rew,
                     uint8_t* done, uint8_t* trunc, float* final_obs, double* stats_out) {
  static CMem S;
  memset(&S, 0, sizeof(S));
  MbStats st;
  memset(&st, 0, sizeof(st));
  CEnv::step(S, *p, state, rec, nullptr, nullptr, act, obs, rew, done, trunc, final_obs, &st);
  stats_out[0] = (double)st.episodes; stats_out[1] = st.ret_sum; stats_out[2] = st.len_sum;
  stats_out[3] = (double)st.nonfinite;
}

void emu_cassie_step_physics(const MbPhysics* p, float* state, const float* tau, int* rows, int* contacts) {
  static CMem S;
  memset(&S, 0, sizeof(S));
  CEnv::load_state(S, state);
  for (int j = 0; j < CM::NJ; ++j) S.tau[j] = tau[j];
  int r = 0, nc = 0, ov = 0;
  Sim<CM>::LaneConst C;
  Sim<CM>::init_lane_const(C);
  for (int k = 0; k < p->substeps; ++k) r += Sim<CM>::substep<0>(S, *p, C, &nc, &ov);
  CEnv::store_state(S, state);
  *rows = r;
  *contacts = nc;
}

// the same with the contact points of the last collision pass ([MB_MAXC][MB_POINT_WIDTH], see Sim::substep)
void emu_cassie_step_physics_points(const MbPhysics* p, float* state, const float* tau, int* rows, int* contacts,
                                    float* points) {
  static CMem S;
  memset(&S, 0, sizeof(S));
  CEnv::load_state(S, state);
  for (int j = 0; j < CM::NJ; ++j) S.tau[j] = tau[j];
  int r = 0, nc = 0, ov = 0;
  Sim<CM>::LaneConst C;
  Sim<CM>::init_lane_const(C);
  for (int k = 0; k < p->substeps; ++k)
    r += Sim<CM>::substep<0>(S, *p, C, &nc, &ov, k, k == p->substeps - 1 ? points : nullptr);
  CEnv::store_state(S, state);
  *rows = r;
  *contacts = nc;
}

void emu_cassie_mass_matrix(const MbPhysics* p, const float* state, float* Mout, float* bias) {
  static CMem S;
  memset(&S, 0, sizeof(S));
  CEnv::load_state(S, state);
  Sim<CM>::LaneConst C;
  Sim<CM>::init_lane_const(C);
  Sim<CM>::kinematics(S, *p, C, true);
  Sim<CM>::bodies(S, *p);
  Sim<CM>::mass_matrix_and_rhs(S);
  const int NU = CM::NU;
  for (int i = 0; i < NU; ++i) {
    for (int j = 0; j < NU; ++j) Mout[i * NU + j] = mb_Lget<CM>(S.L, i, j);
    bias[i] = -S.rhs[i];
  }
}

// ---- more model tables through the same env templates (Child3DCustomEnv-v0, MikeStepperEnv-v0)
#define EMU_ENV(PFX, MODEL, ENV, OBST)                                                                                \
  void emu_##PFX##_reset(const MbPhysics* p, float* state, float* rec, uint32_t* mt_env, uint32_t* mt_robot,         \
                         float* obs) {                                                                                \
    static WarpMem<MODEL> S;                                                                                          \
    memset(&S, 0, sizeof(S));                                                                                         \
    ENV::reset(S, *p, rec, mt_env, mt_robot, obs);                                                                    \
    W3DEnv<MODEL>::store_state(S, state);                                                                             \
  }                                                                                                                   \
  void emu_##PFX##_step(const MbPhysics* p, float* state, float* rec, uint32_t* mt_env, uint32_t* mt_robot,          \
                        const float* act, float* obs, float* rew, uint8_t* done, uint8_t* trunc, float* final_obs,   \
                        double* stats_out) {                                                                          \
    static WarpMem<MODEL> S;                                                                                          \
    memset(&S, 0, sizeof(S));                                                                                         \
    MbStats st;                                                                                                       \
    memset(&st, 0, sizeof(st));                                                                                       \
    ENV::step(S, *p, state, rec, mt_env, mt_robot, act, obs, rew, done, trunc, final_obs, &st);                       \
    stats_out[0] = (double)st.episodes; stats_out[1] = st.ret_sum; stats_out[2] = st.len_sum;                         \
    stats_out[3] = (double)st.nonfinite;                                                                              \
  }                                                                                                                   \
  void emu_##PFX##_step_physics(const MbPhysics* p, float* state, const float* rec, const float* tau, int* rows,     \
                                int* contacts) {                                                                      \
    static WarpMem<MODEL> S;                                                                                          \
    memset(&S, 0, sizeof(S));                                                                                         \
    W3DEnv<MODEL>::load_state(S, state);                                                                              \
    for (int j = 0; j < MODEL::NJ; ++j) S.tau[j] = tau[j];                                                            \
    int r = 0, nc = 0, ov = 0;                                                                                        \
    Sim<MODEL>::LaneConst C;                                                                                          \
    Sim<MODEL>::init_lane_const(C);                                                                                   \
    for (int k = 0; k < p->substeps; ++k) {                                                                           \
      ENV::load_obstacles(S, rec);                                                                                    \
      r += Sim<MODEL>::substep<OBST>(S, *p, C, &nc, &ov);                                                             \
    }                                                                                                                 \
    W3DEnv<MODEL>::store_state(S, state);                                                                             \
    *rows = r;                                                                                                        \
    *contacts = nc;                                                                                                   \
  }                                                                                                                   \
  void emu_##PFX##_mass_matrix(const MbPhysics* p, const float* state, float* Mout, float* bias) {                   \
    static WarpMem<MODEL> S;                                                                                          \
    memset(&S, 0, sizeof(S));                                                                                         \
    W3DEnv<MODEL>::load_state(S, state);                                                                              \
    Sim<MODEL>::LaneConst C;                                                                                          \
    Sim<MODEL>::init_lane_const(C);                                                                                   \
    Sim<MODEL>::kinematics(S, *p, C, true);                                                                           \
    Sim<MODEL>::bodies(S, *p);                                                                                        \
    Sim<MODEL>::mass_matrix_and_rhs(S);                                                                               \
    const int NU = MODEL::NU;                                                                                         \
    for (int i = 0; i < NU; ++i) {                                                                                    \
      for (int j = 0; j < NU; ++j) Mout[i * NU + j] = mb_Lget<MODEL>(S.L, i, j);                                      \
      bias[i] = -S.rhs[i];                                                                                            \
    }                                                                                                                 \
  }
EMU_ENV(child, CH3D_Model, W3DEnv<CH3D_Model>, 0)
EMU_ENV(mike, MIKE_Model, StepperEnv<MIKE_Model>, MB_OBST_BOXES)
// the planar walkers (Walker2DCustomEnv-v0, Crab2DCustomEnv-v0)
EMU_ENV(walker2d, W2D_Model, W3DEnv<W2D_Model>, 0)
EMU_ENV(crab2d, CR2D_Model, W3DEnv<CR2D_Model>, 0)
// plank_class = "Pillar": the PILLAR instantiation of the stepper template
typedef StepperEnv<WM, true> SEnvPillar;
EMU_ENV(pillar, WM, SEnvPillar, MB_OBST_CYLS)
}
