// placeholder replaced by the tcgen05 kernel below in this commit series
#include "attn.cuh"
namespace rmem {
size_t long_attn_tc_workspace(int HW, int HWp, int nslots, int Dv) { return 256; }
int long_attn_tc(const LongAttnArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  set_error("long_attn_tc: tcgen05 kernel not built yet");
  return RMEM_ERR_STATE;
}
}
