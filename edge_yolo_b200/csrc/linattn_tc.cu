// tcgen05 / TMEM path of the linear-attention core (bf16 / fp16 inputs).  Placeholder until the
// tensor-core kernel lands: reports "unsupported" so el_linattn_fwd takes the fp32 CUDA-core kernel.
#include "el_common.cuh"

namespace el {
struct AttnArgs;
bool linattn_tc_supported(const AttnArgs&, int) { return false; }
int linattn_tc_launch(const AttnArgs&, int, int, cudaStream_t) { return EL_ERR_UNSUPPORTED; }
}  // namespace el
