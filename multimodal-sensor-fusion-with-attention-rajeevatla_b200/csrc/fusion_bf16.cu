// HybridFusion forward / backward on the tensor-core path (MSF_PREC_BF16).
// Placeholder until the tcgen05 pipeline lands: reports the shape as ineligible
// so callers get MSF_E_UNSUPPORTED rather than a silent fallback.
#include "msf_common.cuh"

namespace msf {
bool fusion_bf16_eligible(const Layout&) { return false; }
size_t fusion_bf16_workspace_bytes(const Layout&, int64_t) { return 0; }
size_t fusion_bf16_arena_bytes(const Layout&) { return 0; }
int fusion_bf16_pack(const Layout&, const float*, void*, cudaStream_t) { return MSF_E_UNSUPPORTED; }
int fusion_bf16_forward(const Layout&, const msf_fusion_call*, cudaStream_t) { return MSF_E_UNSUPPORTED; }
int fusion_bf16_backward(const Layout&, const msf_fusion_call*, cudaStream_t) { return MSF_E_UNSUPPORTED; }
}  // namespace msf

