// te_footprint.cu — footprint sweep (placeholder until the kernel lands).
#include "te_footprint.h"
namespace te {
void FootprintState::release() {
  if (d_spiral) cudaFree(d_spiral);
  if (d_block) cudaFree(d_block);
  d_spiral = d_block = nullptr;
  spiral_cap = block_cap = 0;
  valid = false;
}
int footprint_halo(const te_geometry*, const te_footprint_params*) { return 0; }
int launch_footprint(FootprintState& st, const SlabView&, const te_geometry*, const te_footprint_params*, const std::vector<double>&,
                     const std::vector<double>&, const float*, const float*, const float*, const float*, float*, float*, float*, int,
                     cudaStream_t, int*) {
  st.why = "footprint sweep not built yet";
  return TE_ERR_UNSUPPORTED;
}
}  // namespace te
