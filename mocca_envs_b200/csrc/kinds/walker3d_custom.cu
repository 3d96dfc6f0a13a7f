// Walker3DCustomEnv-v0: kernels of this env kind (reference env_locomotion.py:27-222).
#include "../generated/walker3d_model.h"
#include "../mb_kind.cuh"
typedef W3DEnv<W3D_Model> KindEnv;
MB_DEFINE_KIND(walker3d_custom, "Walker3DCustomEnv-v0", "", KindEnv, MB_WARPS_DEFAULT)
