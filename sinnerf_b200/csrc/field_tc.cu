// field_tc.cu -- tensor-core (tcgen05) field pass.  Placeholder until the UMMA kernel lands:
// every entry point reports SNB_ERR_UNSUPPORTED (never a silent fallback).
#include "common.cuh"

namespace snb {
size_t tc_packed_bytes(int) { return 0; }
int launch_pack_tc(const float* const*, int precision, int, void*, cudaStream_t) {
  return fail(SNB_ERR_UNSUPPORTED, "precision mode %d (tensor-core field pass) is not built yet", precision);
}
int field_forward_tc(const void*, int precision, const float*, const float*, int64_t, int, int, float*, cudaStream_t) {
  return fail(SNB_ERR_UNSUPPORTED, "precision mode %d (tensor-core field pass) is not built yet", precision);
}
int mlp_forward_tc(const void*, int precision, const float*, int64_t, int64_t, int, float*, cudaStream_t) {
  return fail(SNB_ERR_UNSUPPORTED, "precision mode %d (tensor-core field pass) is not built yet", precision);
}
}  // namespace snb
