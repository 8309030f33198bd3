// Backward twins of the GRU kernels (placeholder until the BPTT kernels land).
#include "gru.cuh"

extern "C" int ign_gru_cell_bwd(const float*, const float*, int64_t, int, int, const float*, const float*,
                                const float*, const float*, float*, float*, float*, float*, float*, void*) {
  ign_set_error("IGNNITION: gru_cell_bwd is not built yet");
  return IGN_ERR_UNSUPPORTED;
}
extern "C" int ign_gru_seq_bwd(const int32_t*, const int32_t*, const int32_t*, int, const float* const*, int,
                               const float*, const float*, int64_t, int, const float*, const float*, const float*,
                               const float*, float*, float*, float*, float*, float*, void*) {
  ign_set_error("IGNNITION: gru_seq_bwd is not built yet");
  return IGN_ERR_UNSUPPORTED;
}
