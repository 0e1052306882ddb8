// Extend / chunked-prefill attention -- placeholder entry point until the kernel lands.
#include "common.cuh"
extern "C" int b200_attn_extend(const b200_extend_params* p, b200_stream_t stream) {
  (void)p; (void)stream;
  return B200_ERR_UNSUPPORTED;
}
