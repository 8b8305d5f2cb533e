// CTC beam search (K.ctc_decode(greedy=False) -> TF CTCBeamSearchDecoder semantics).  See beam
// section of DESIGN.md.  [placeholder translation unit: implemented in a following milestone]
#include "common.cuh"

extern "C" int gr_ctc_beam_workspace_bytes(int N, int T, int C, int beam_width, size_t* bytes_out) {
  if (!bytes_out || N <= 0 || T <= 0 || C < 2 || beam_width <= 0) return gr::set_error(GR_EINVAL, "beam_workspace_bytes: bad argument");
  *bytes_out = 256;
  return GR_OK;
}
extern "C" int gr_ctc_beam_f32(const float*, int, int, int, const int32_t*, float, int, int, int, int32_t*, int32_t*,
                               float*, void*, size_t, void*) {
  return gr::set_error(GR_EUNSUPPORTED, "gr_ctc_beam_f32: not built yet");
}
