// tcgen05 / TMEM / TMA implicit-GEMM convolution -- placeholder until the kernel lands.
#include <stdio.h>

#include "kernels.h"

namespace accel {

struct TcPlan { int unused; };
bool tc_supported(const ConvParams&) { return false; }
TcPlan* tc_plan_create(const ConvParams&, int, char* err, int errlen) {
  if (err && errlen > 0) snprintf(err, errlen, "tcgen05 engine not built");
  return nullptr;
}
void tc_plan_destroy(TcPlan* p) { delete p; }
size_t tc_plan_partial_bytes(const TcPlan*) { return 0; }
void tc_plan_set_partial(TcPlan*, float*) {}
cudaError_t launch_conv_tc_ext(const TcPlan*, float*, cudaStream_t) { return cudaErrorNotSupported; }
int tc_plan_launches(const TcPlan*) { return 0; }

}  // namespace accel
