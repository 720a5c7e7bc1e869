// rollout_f32_screen.cu -- rollout_kernel<float, *, MODE_SCREEN, *, *> instantiations (see rollout_tu.inc)
#define TU_REAL float
#define TU_MODE MODE_SCREEN
#define TU_NAME(x) rollout_f32_screen_##x
#define TU_HAS_LEAN 1
#include "rollout_tu.inc"
