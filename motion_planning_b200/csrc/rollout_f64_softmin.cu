// rollout_f64_softmin.cu -- rollout_kernel<double, *, MODE_SOFTMIN, *, *> instantiations (see rollout_tu.inc)
#define TU_REAL double
#define TU_MODE MODE_SOFTMIN
#define TU_NAME(x) rollout_f64_softmin_##x
#include "rollout_tu.inc"
