// Batched small-circuit kernel (VQLS Hadamard tests) -- placeholder until the kernel lands.
#include "qls_common.cuh"

extern "C" int qls_batch_expect_z(const QlsProgram*, uint32_t, int, int, const double*, uint32_t, uint32_t, double*, void*) {
  qls_set_error("qls_batch_expect_z: not built yet");
  return QLS_ERR_UNSUPPORTED;
}
