// rollout_f32_softmin.cu -- rollout_kernel<float, *, MODE_SOFTMIN, *, *> instantiations (see rollout_tu.inc)
#define TU_REAL float
#define TU_MODE MODE_SOFTMIN
#define TU_NAME(x) rollout_f32_softmin_##x
#define TU_HAS_LEAN 1
#include "rollout_tu.inc"
