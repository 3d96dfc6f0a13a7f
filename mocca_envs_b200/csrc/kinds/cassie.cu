// CassieEnv-v0: kernels of this env kind (reference env_cassie.py:285-479).
#include "../generated/cassie_model.h"
#include "../mb_kind.cuh"
typedef CassieEnv<CAS_Model> KindEnv;
MB_DEFINE_KIND(cassie, "CassieEnv-v0", "", KindEnv, MB_WARPS_DEFAULT)
