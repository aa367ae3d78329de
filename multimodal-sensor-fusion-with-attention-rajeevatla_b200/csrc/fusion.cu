// C-ABI entry points of the HybridFusion path: validation + precision dispatch.
#include <stdlib.h>

#include "msf_common.cuh"

namespace msf {
size_t fusion_f32_workspace_bytes(const Layout& L, int64_t B);
int fusion_f32_forward(const Layout& L, const msf_fusion_call* c, cudaStream_t st);
bool proj_eligible(int H, int M, const int* D);
int fusion_f32_backward(const Layout& L, const msf_fusion_call* c, cudaStream_t st);

// tensor-core path (fusion_bf16.cu)
bool fusion_bf16_eligible(const Layout& L);
size_t fusion_bf16_workspace_bytes(const Layout& L, int64_t B);
size_t fusion_bf16_arena_bytes(const Layout& L);
int fusion_bf16_pack(const Layout& L, const float* params, void* arena, cudaStream_t st);
int fusion_bf16_forward(const Layout& L, const msf_fusion_call* c, cudaStream_t st);
int fusion_bf16_backward(const Layout& L, const msf_fusion_call* c, cudaStream_t st);
bool fusion_bf16_head_fused(const Layout& L);
int head_debug_stamps(long long* out16);
int chain_debug_stamps(long long* out16);
int proj_debug_stamps(long long* out16);
int fusion_bf16_train(const Layout& L, const msf_fusion_call* c, const int64_t* labels, float smoothing,
                      float grad_scale, float* row_loss, float* loss_out, int flags, cudaStream_t st);
int fusion_bf16_infer(const Layout& L, const msf_fusion_call* c, float* conf, int64_t* pred, unsigned present_hint,
                      cudaStream_t st);
int fusion_bf16_infer_folded(const Layout& L, const msf_fusion_call* c, const void* wov, const float* bias_sum,
                             unsigned present, int flags, float* conf, int64_t* pred, cudaStream_t st);

static int check_call(const Layout& L, const msf_fusion_call* c, bool backward) {
  MSF_REQUIRE(c != nullptr, "null call");
  MSF_REQUIRE(c->batch >= 1, "batch must be >= 1 (got %d)", c->batch);
  MSF_REQUIRE(c->precision == MSF_PREC_F32 || c->precision == MSF_PREC_BF16, "unknown precision %d",
              c->precision);
  MSF_REQUIRE(c->dropout_p >= 0.0f && c->dropout_p < 1.0f, "dropout_p must be in [0, 1)");
  MSF_REQUIRE(c->params != nullptr, "null params");
  for (int m = 0; m < L.M; ++m) MSF_REQUIRE(c->x[m] != nullptr, "null features for modality %d", m);
  MSF_REQUIRE(c->workspace != nullptr, "null workspace");
  if (!backward) MSF_REQUIRE(c->logits != nullptr, "null logits");
  if (c->x_bf16 && c->precision != MSF_PREC_BF16) {
    set_error("msf_fusion_call.x_bf16 needs BF16 precision (the fp32 parity path reads fp32 features)");
    return MSF_E_UNSUPPORTED;
  }
  if (c->precision == MSF_PREC_BF16) {
    MSF_REQUIRE(c->params_bf16 != nullptr, "BF16 precision needs params_bf16 (msf_fusion_pack_bf16)");
    if (!fusion_bf16_eligible(L)) {
      set_error("shape not eligible for the tensor-core path (needs hidden %% 64 == 0 and every in_dim %% 8 == 0)");
      return MSF_E_UNSUPPORTED;
    }
  }
  return MSF_OK;
}
}  // namespace msf

extern "C" {

int msf_fusion_workspace_bytes(const msf_fusion_shape* shape, int32_t batch, int32_t precision, size_t* bytes) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  MSF_REQUIRE(bytes != nullptr && batch >= 1, "msf_fusion_workspace_bytes: bad arguments");
  if (precision == MSF_PREC_F32) {
    *bytes = msf::fusion_f32_workspace_bytes(L, batch);
    return MSF_OK;
  }
  if (precision == MSF_PREC_BF16) {
    if (!msf::fusion_bf16_eligible(L)) {
      msf::set_error("shape not eligible for the tensor-core path");
      return MSF_E_UNSUPPORTED;
    }
    *bytes = msf::fusion_bf16_workspace_bytes(L, batch);
    return MSF_OK;
  }
  msf::set_error("unknown precision %d", precision);
  return MSF_E_INVALID;
}

int msf_fusion_bf16_arena_bytes(const msf_fusion_shape* shape, size_t* bytes) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  MSF_REQUIRE(bytes != nullptr, "null bytes");
  if (!msf::fusion_bf16_eligible(L)) {
    msf::set_error("shape not eligible for the tensor-core path");
    return MSF_E_UNSUPPORTED;
  }
  *bytes = msf::fusion_bf16_arena_bytes(L);
  return MSF_OK;
}

int msf_fusion_pack_bf16(const msf_fusion_shape* shape, const float* params, void* params_bf16, void* stream) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  MSF_REQUIRE(params && params_bf16, "msf_fusion_pack_bf16: null pointer");
  if (!msf::fusion_bf16_eligible(L)) {
    msf::set_error("shape not eligible for the tensor-core path");
    return MSF_E_UNSUPPORTED;
  }
  return msf::fusion_bf16_pack(L, params, params_bf16, (cudaStream_t)stream);
}

int msf_fusion_layer_norm_fused(const msf_fusion_shape* shape, int32_t precision) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  return (precision == MSF_PREC_BF16 && msf::fusion_bf16_eligible(L) && msf::proj_eligible(L.H, L.M, L.D) &&
          !getenv("MSF_NO_PROJ")) ? 1 : 0;
}

int msf_fusion_forward(const msf_fusion_shape* shape, const msf_fusion_call* call, void* stream) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  if ((rc = msf::check_call(L, call, false))) return rc;
  if (call->precision == MSF_PREC_F32)
    for (int m = 0; m < L.M; ++m)
      MSF_REQUIRE(call->ln_weight[m] == nullptr && call->ln_bias[m] == nullptr,
                  "msf_fusion_call.ln_* is only applied on the tensor-core path (msf_fusion_layer_norm_fused): run "
                  "msf_layer_norm_forward first");
  return call->precision == MSF_PREC_F32 ? msf::fusion_f32_forward(L, call, (cudaStream_t)stream)
                                         : msf::fusion_bf16_forward(L, call, (cudaStream_t)stream);
}

int msf_fusion_backward(const msf_fusion_shape* shape, const msf_fusion_call* call, void* stream) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  if ((rc = msf::check_call(L, call, true))) return rc;
  return call->precision == MSF_PREC_F32 ? msf::fusion_f32_backward(L, call, (cudaStream_t)stream)
                                         : msf::fusion_bf16_backward(L, call, (cudaStream_t)stream);
}

int msf_fusion_train_pass(const msf_fusion_shape* shape, const msf_fusion_call* call, const int64_t* labels,
                          float smoothing, float grad_scale, float* row_loss, float* loss_out,
                          float* grad_logits_scratch, int32_t flags, void* stream) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  if ((rc = msf::check_call(L, call, false))) return rc;
  MSF_REQUIRE(labels && row_loss && call->grad_params, "msf_fusion_train_pass: labels, row_loss and grad_params are required");
  if (call->precision == MSF_PREC_BF16 && msf::fusion_bf16_head_fused(L))
    return msf::fusion_bf16_train(L, call, labels, smoothing, grad_scale, row_loss, loss_out, flags, (cudaStream_t)stream);
  // un-fused composition (fp32 parity path, shapes outside the head kernel): same three steps
  MSF_REQUIRE(call->grad_sq == nullptr, "msf_fusion_train_pass: grad_sq needs the fused path (msf_fusion_train_pass_is_fused)");
  MSF_REQUIRE(grad_logits_scratch != nullptr, "msf_fusion_train_pass: this precision / shape needs grad_logits_scratch");
  if ((rc = msf_fusion_forward(shape, call, stream))) return rc;
  if ((rc = msf_cross_entropy(call->logits, labels, call->batch, L.C, smoothing, grad_scale, row_loss, loss_out,
                              grad_logits_scratch, stream)))
    return rc;
  msf_fusion_call back = *call;
  back.grad_logits = grad_logits_scratch;
  return msf_fusion_backward(shape, &back, stream);
}

int msf_fusion_train_pass_is_fused(const msf_fusion_shape* shape, int32_t precision) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  return (precision == MSF_PREC_BF16 && msf::fusion_bf16_head_fused(L)) ? 1 : 0;
}

int msf_fusion_infer_pass(const msf_fusion_shape* shape, const msf_fusion_call* call, float* conf, int64_t* pred,
                          uint32_t present_hint, void* stream) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  if ((rc = msf::check_call(L, call, false))) return rc;
  MSF_REQUIRE(conf && pred, "msf_fusion_infer_pass: conf and pred are required");
  MSF_REQUIRE(present_hint < (1u << L.M), "msf_fusion_infer_pass: present_hint has bits beyond the modalities");
  if (call->precision == MSF_PREC_BF16 && msf::fusion_bf16_head_fused(L))
    return msf::fusion_bf16_infer(L, call, conf, pred, present_hint, (cudaStream_t)stream);
  if ((rc = msf_fusion_forward(shape, call, stream))) return rc;
  return msf_softmax_conf_pred(call->logits, call->batch, L.C, conf, pred, stream);
}

int msf_fusion_infer_folded(const msf_fusion_shape* shape, const msf_fusion_call* call, const void* wov_bf16,
                            const float* bias_sum, uint32_t present, int32_t flags, float* conf, int64_t* pred,
                            void* stream) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  if ((rc = msf::check_call(L, call, false))) return rc;
  if (call->precision != MSF_PREC_BF16) {
    msf::set_error("msf_fusion_infer_folded runs on the tensor-core path only");
    return MSF_E_UNSUPPORTED;
  }
  return msf::fusion_bf16_infer_folded(L, call, wov_bf16, bias_sum, present, flags, conf, pred, (cudaStream_t)stream);
}

int msf_debug_chain_stamps(int64_t* out16) {
  MSF_REQUIRE(out16 != nullptr, "msf_debug_chain_stamps: null output");
  return msf::chain_debug_stamps(reinterpret_cast<long long*>(out16));
}

int msf_debug_proj_stamps(int64_t* out16) {
  MSF_REQUIRE(out16 != nullptr, "msf_debug_proj_stamps: null output");
  return msf::proj_debug_stamps(reinterpret_cast<long long*>(out16));
}

int msf_debug_head_stamps(int64_t* out16) {
  MSF_REQUIRE(out16 != nullptr, "msf_debug_head_stamps: null output");
  return msf::head_debug_stamps(reinterpret_cast<long long*>(out16));
}

}  // extern "C"
