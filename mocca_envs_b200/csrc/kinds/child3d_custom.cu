// Child3DCustomEnv-v0: kernels of this env kind (reference env_locomotion.py:317-327).
#include "../generated/child3d_model.h"
#include "../mb_kind.cuh"
typedef W3DEnv<CH3D_Model> KindEnv;
MB_DEFINE_KIND(child3d_custom, "Child3DCustomEnv-v0", "", KindEnv, MB_WARPS_DEFAULT)
