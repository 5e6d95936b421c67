// env_step_kernel instantiations for visible_range = 5
#include "env_kernel.cuh"
namespace maexp {
KernelFn pick_v5(int epw, int noise, int nspec) { return pick_epw<5>(epw, noise, nspec); }
}  // namespace maexp
