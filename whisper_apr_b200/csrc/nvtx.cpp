// NVTX ranges around the two pipeline stages, named after the reference's renacer trace spans: `step_f_mel`
// (src/audio/mel.rs:234, MelFilterbank::compute) and `step_g_encode` (.renacer.toml:11-32).  NVTX v3 is header-only: without a
// profiler attached (no NVTX_INJECTION64_PATH) a push/pop is one predictable branch; under nsys / ncu --nvtx the ranges show up.
#include <nvtx3/nvToolsExt.h>

namespace wb {
void nvtx_push(const char* name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }
}  // namespace wb
