// env_step_kernel instantiations for visible_range = 7
#include "env_kernel.cuh"
namespace maexp {
KernelFn pick_v7(int epw, int noise, int nspec) { return pick_epw<7>(epw, noise, nspec); }
}  // namespace maexp
