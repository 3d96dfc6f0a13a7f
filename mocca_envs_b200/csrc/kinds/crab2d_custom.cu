// Crab2DCustomEnv-v0: kernels of this env kind (reference env_locomotion.py:312-314).
#include "../generated/crab2d_model.h"
#include "../mb_kind.cuh"
typedef W3DEnv<CR2D_Model> KindEnv;
MB_DEFINE_KIND(crab2d_custom, "Crab2DCustomEnv-v0", "", KindEnv, MB_WARPS_DEFAULT)
