// Pieces shared by the fp32 and the tensor-core HybridFusion paths: the
// adaptive-weighting arithmetic of one window (src/fusion.py:429-479) and the
// dropout configuration of a call.
#pragma once

#include <math.h>

#include "msf_common.cuh"

namespace msf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// weights for one window from scores s[] and mask mk[]; returns branch taken (1 = softmax branch)
__device__ __forceinline__ int adaptive_weights_row(const float* s, const float* mk, int M, float* soft,
                                                    float* w) {
  float mx = -INFINITY;
  for (int q = 0; q < M; ++q)
    if (mk[q] > 0.0f) mx = fmaxf(mx, s[q]);  // masked_fill(mask <= 0, -inf)   fusion.py:464
  float den = 0.0f;
  for (int q = 0; q < M; ++q) {
    // all-masked row: softmax of all -inf is NaN -> nan_to_num -> 0          fusion.py:465-466
    soft[q] = (mk[q] > 0.0f && mx > -INFINITY) ? expf(s[q] - mx) : 0.0f;
    den += soft[q];
  }
  float sum_w = 0.0f, mask_sum = 0.0f;
  for (int q = 0; q < M; ++q) {
    soft[q] = den > 0.0f ? soft[q] / den : 0.0f;
    w[q] = soft[q] * mk[q];  // fusion.py:467
    sum_w += w[q];
    mask_sum += mk[q];
  }
  if (sum_w > 0.0f) {  // fusion.py:476-478
    for (int q = 0; q < M; ++q) w[q] = w[q] / (sum_w + 1e-8f);
    return 1;
  }
  for (int q = 0; q < M; ++q)  // fusion.py:471-475
    w[q] = mask_sum > 0.0f ? mk[q] / (mask_sum + 1e-8f) : 1.0f / (float)M;
  return 0;
}

static inline DropCfg make_drop(const msf_fusion_call* c) {
  DropCfg d;
  d.seed = c->seed;
  d.offset = c->offset;
  d.p = c->dropout_p;
  d.active = (c->training && c->dropout_p > 0.0f) ? 1 : 0;
  d.scale = d.active ? 1.0f / (1.0f - c->dropout_p) : 1.0f;
  d.state = reinterpret_cast<const unsigned long long*>(c->rng_state);
  return d;
}

}  // namespace msf
