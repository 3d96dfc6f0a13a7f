// Walker3DStepperEnv-v0 (pillar): kernels of this env kind (plank_class = Pillar, reference bullet_objects.py:86-90).
#include "../generated/walker3d_model.h"
#include "../mb_kind.cuh"
typedef StepperEnv<W3D_Model, true> KindEnv;
MB_DEFINE_KIND(walker3d_stepper_pillar, "Walker3DStepperEnv-v0", "pillar", KindEnv, MB_WARPS_DEFAULT)
