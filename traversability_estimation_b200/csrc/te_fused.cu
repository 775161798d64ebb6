// te_fused.cu — fused chain stencil (placeholder until the specialised kernel lands).
#include "te_fused.h"
namespace te {
void FusedState::release() {
  if (d_rowmask) cudaFree(d_rowmask);
  if (d_colmask) cudaFree(d_colmask);
  d_rowmask = d_colmask = nullptr;
  rowmask_cap = colmask_cap = 0;
  valid = false;
}
bool fused_eligible(FusedState& st, const std::vector<double>&, const std::vector<double>&, const te_geometry*, const te_chain_params*) {
  st.why = "fused stencil not built yet";
  return false;
}
int launch_chain_fused(FusedState& st, const SlabView&, const ChainDev&, const float*, const ChainOut&, unsigned*, unsigned*, unsigned,
                       int, cudaStream_t) {
  st.why = "fused stencil not built yet";
  return 1;
}
}  // namespace te
