// MikeStepperEnv-v0: kernels of this env kind (reference env_locomotion.py:843-851).
#include "../generated/mike_model.h"
#include "../mb_kind.cuh"
typedef StepperEnv<MIKE_Model> KindEnv;
MB_DEFINE_KIND(mike_stepper, "MikeStepperEnv-v0", "", KindEnv, MB_WARPS_DEFAULT)
