// Walker2DCustomEnv-v0: kernels of this env kind (reference env_locomotion.py:285-310).
#include "../generated/walker2d_model.h"
#include "../mb_kind.cuh"
typedef W3DEnv<W2D_Model> KindEnv;
MB_DEFINE_KIND(walker2d_custom, "Walker2DCustomEnv-v0", "", KindEnv, MB_WARPS_DEFAULT)
