// env_step_kernel instantiations for visible_range = 9
#include "env_kernel.cuh"
namespace maexp {
KernelFn pick_v9(int epw, int noise, int nspec) { return pick_epw<9>(epw, noise, nspec); }
}  // namespace maexp
