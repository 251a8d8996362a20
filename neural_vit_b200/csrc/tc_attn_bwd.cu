#include "common.cuh"
namespace tvit {
size_t tc_attn_bwd_workspace(int, int, int, int) { return 0; }
int tc_attn_bwd(const void*, const void*, const void*, const float*, void*, void*, size_t, int, int, int, int,
                const tvit_dropout*, cudaStream_t) {
  return fail(TVIT_ERR_UNSUPPORTED, "tcgen05 attention backward not built");
}
}  // namespace tvit
