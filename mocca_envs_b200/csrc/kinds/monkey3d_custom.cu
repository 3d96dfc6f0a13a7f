// Monkey3DCustomEnv-v0: kernels of this env kind (reference env_locomotion.py:1136-1516).
#include "../generated/monkey3d_model.h"
#include "../mb_kind.cuh"
typedef MonkeyEnv<MK3D_Model> KindEnv;
MB_DEFINE_KIND(monkey3d_custom, "Monkey3DCustomEnv-v0", "", KindEnv, MB_WARPS_DEFAULT)
