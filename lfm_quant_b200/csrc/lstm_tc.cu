#include "lstm_tc.h"

namespace lfmq {
void tc_layout(TcState&, const lfmq_config&, char*, size_t&) {}
int tc_init(TcState&, const lfmq_config& cfg) {
  if (cfg.precision == LFMQ_PREC_BF16) {
    LFMQ_SET_ERR("LFMQ_PREC_BF16 not built yet");
    return LFMQ_ERR_UNSUPPORTED;
  }
  return 0;
}
void tc_destroy(TcState&) {}
int tc_forward(TcState&, const lfmq_config&, const float*, const float*, int, int64_t, int64_t, float*, bool,
               cudaStream_t) {
  LFMQ_SET_ERR("LFMQ_PREC_BF16 not built yet");
  return LFMQ_ERR_UNSUPPORTED;
}
int tc_backward(TcState&, const lfmq_config&, const float*, float*, const float*, const float*, int, int64_t, int64_t,
                const float*, float*, cudaStream_t) {
  LFMQ_SET_ERR("LFMQ_PREC_BF16 not built yet");
  return LFMQ_ERR_UNSUPPORTED;
}
}  // namespace lfmq
