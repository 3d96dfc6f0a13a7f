// MikeStepperEnv-v0 (pillar): kernels of this env kind (plank_class = Pillar on the Mike table).
#include "../generated/mike_model.h"
#include "../mb_kind.cuh"
typedef StepperEnv<MIKE_Model, true> KindEnv;
MB_DEFINE_KIND(mike_stepper_pillar, "MikeStepperEnv-v0", "pillar", KindEnv, MB_WARPS_DEFAULT)
