// Walker3DStepperEnv-v0: kernels of this env kind (reference env_locomotion.py:330-840).
#include "../generated/walker3d_model.h"
#include "../mb_kind.cuh"
typedef StepperEnv<W3D_Model> KindEnv;
MB_DEFINE_KIND(walker3d_stepper, "Walker3DStepperEnv-v0", "", KindEnv, MB_WARPS_DEFAULT)
