// tcgen05 / TMA dense-layer GEMMs (placeholder until the kernel lands: reports "unsupported" so that every
// call site uses the SIMT kernels).
#include "kernels.h"

namespace vaeassoc {
struct TcPlan { int unused; };
bool tc_supported(int, const GemmArgs&) { return false; }
TcPlan* tc_plan_create(int, const GemmArgs&, char*, int) { return nullptr; }
void tc_plan_destroy(TcPlan* p) { delete p; }
void launch_gemm_tc(const TcPlan*, cudaStream_t) {}
}  // namespace vaeassoc
