// Generic multi-head scaled-dot-product core of CrossModalAttention for
// q_len, k_len >= 1 (src/attention.py:108-139): scores, key mask, softmax,
// NaN -> 0, dropout on the weights, weights . V.  fp32, one warp per
// (window, head, query position).  HybridFusion itself never calls this (its
// 1x1 attention is folded into the fused path as a gate); it backs the
// stand-alone nn.Module API.
#include <math.h>

#include "msf_common.cuh"

namespace msf {
namespace {

constexpr int SITE_CORE = 4;

struct CoreArgs {
  const float* q;     // (B, Lq, H) projected
  const float* k;     // (B, Lk, H)
  const float* v;     // (B, Lk, H)
  const float* mask;  // (B, Lk) or null
  long long B;
  int Lq, Lk, H, heads;
  float scale;
  DropCfg drop;
  float* weights;     // (B, heads, Lq, Lk) post-dropout
  float* out;         // (B, Lq, H)
  // backward
  const float* gout;  // (B, Lq, H)
  float* scratch;     // (B, heads, Lq, Lk)
  float* dq;
  float* dk;
  float* dv;
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// softmax over keys of one (b, head, qi) row into dst[0..Lk) (pre-dropout)
__device__ __forceinline__ void softmax_row(const CoreArgs& a, long long b, int head, int qi, int lane,
                                            float* dst) {
  const int hd = a.H / a.heads;
  const float* qv = a.q + ((b * a.Lq + qi) * (long long)a.H) + head * hd;
  float mx = -INFINITY;
  for (int j = lane; j < a.Lk; j += 32) {
    const float* kv = a.k + ((b * a.Lk + j) * (long long)a.H) + head * hd;
    float s = 0.0f;
    for (int d = 0; d < hd; ++d) s = fmaf(__ldg(qv + d), __ldg(kv + d), s);
    s *= a.scale;
    if (a.mask && __ldg(a.mask + b * a.Lk + j) == 0.0f) s = -INFINITY;  // attention.py:124
    dst[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = wmax(mx);
  float den = 0.0f;
  for (int j = lane; j < a.Lk; j += 32) {
    const float e = (mx > -INFINITY) ? expf(dst[j] - mx) : 0.0f;  // all masked: NaN -> 0 (attention.py:127-129)
    dst[j] = e;
    den += e;
  }
  den = wsum(den);
  const float inv = den > 0.0f ? 1.0f / den : 0.0f;
  for (int j = lane; j < a.Lk; j += 32) dst[j] *= inv;
  __syncwarp();
}

__global__ void __launch_bounds__(256) core_fwd_kernel(const __grid_constant__ CoreArgs a) {
  const int lane = threadIdx.x & 31;
  const long long w = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = a.B * a.heads * a.Lq;
  if (w >= total) return;
  const int qi = (int)(w % a.Lq);
  const int head = (int)((w / a.Lq) % a.heads);
  const long long b = w / ((long long)a.Lq * a.heads);
  const int hd = a.H / a.heads;
  float* wr = a.weights + w * a.Lk;  // (b, head, qi, :)
  softmax_row(a, b, head, qi, lane, wr);
  if (a.drop.active) {
    for (int j = lane; j < a.Lk; j += 32) wr[j] *= drop1(a.drop, SITE_CORE, head, b * a.Lq + qi, j);
    __syncwarp();
  }
  float* o = a.out + ((b * a.Lq + qi) * (long long)a.H) + head * hd;
  for (int d = lane; d < hd; d += 32) {
    float acc = 0.0f;
    for (int j = 0; j < a.Lk; ++j)
      acc = fmaf(wr[j], __ldg(a.v + ((b * a.Lk + j) * (long long)a.H) + head * hd + d), acc);
    o[d] = acc;
  }
}

__global__ void __launch_bounds__(256) core_bwd_kernel(const __grid_constant__ CoreArgs a) {
  const int lane = threadIdx.x & 31;
  const long long w = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = a.B * a.heads * a.Lq;
  if (w >= total) return;
  const int qi = (int)(w % a.Lq);
  const int head = (int)((w / a.Lq) % a.heads);
  const long long b = w / ((long long)a.Lq * a.heads);
  const int hd = a.H / a.heads;
  float* pre = a.scratch + w * a.Lk;
  const float* post = a.weights + w * a.Lk;
  const float* go = a.gout + ((b * a.Lq + qi) * (long long)a.H) + head * hd;
  const float* qv = a.q + ((b * a.Lq + qi) * (long long)a.H) + head * hd;
  softmax_row(a, b, head, qi, lane, pre);
  // dWpre[j] = (sum_d gout[d] v[j,d]) * dropmask[j];  dot = sum_j pre[j] dWpre[j]
  float dot = 0.0f;
  for (int j = lane; j < a.Lk; j += 32) {
    const float* vv = a.v + ((b * a.Lk + j) * (long long)a.H) + head * hd;
    float g = 0.0f;
    for (int d = 0; d < hd; ++d) g = fmaf(__ldg(go + d), __ldg(vv + d), g);
    if (a.drop.active) g *= drop1(a.drop, SITE_CORE, head, b * a.Lq + qi, j);
    dot = fmaf(pre[j], g, dot);  // g is recomputed in the second pass (rows may exceed registers)
  }
  dot = wsum(dot);
  for (int j = lane; j < a.Lk; j += 32) {
    const float* vv = a.v + ((b * a.Lk + j) * (long long)a.H) + head * hd;
    float g = 0.0f;
    for (int d = 0; d < hd; ++d) g = fmaf(__ldg(go + d), __ldg(vv + d), g);
    if (a.drop.active) g *= drop1(a.drop, SITE_CORE, head, b * a.Lq + qi, j);
    pre[j] = pre[j] * (g - dot) * a.scale;  // dS[j] * scale
  }
  __syncwarp();
  for (int d = lane; d < hd; d += 32) {
    float acc = 0.0f;
    const float qd = __ldg(qv + d), gd = __ldg(go + d);
    for (int j = 0; j < a.Lk; ++j) {
      const long long kv = ((b * a.Lk + j) * (long long)a.H) + head * hd + d;
      const float ds = pre[j];
      acc = fmaf(ds, __ldg(a.k + kv), acc);
      if (ds != 0.0f) atomicAdd(a.dk + kv, ds * qd);
      const float pw = post[j];
      if (pw != 0.0f) atomicAdd(a.dv + kv, pw * gd);
    }
    a.dq[((b * a.Lq + qi) * (long long)a.H) + head * hd + d] = acc;
  }
}

int fill(CoreArgs* a, const float* q, const float* k, const float* v, const float* mask, int64_t batch,
         int32_t q_len, int32_t k_len, int32_t hidden, int32_t heads, float p, int32_t training,
         uint64_t seed, uint64_t offset) {
  MSF_REQUIRE(q && k && v && batch >= 0 && q_len >= 1 && k_len >= 1 && hidden >= 1 && heads >= 1 &&
                  hidden % heads == 0 && heads <= 255, "attention core: bad arguments");
  MSF_REQUIRE(p >= 0.0f && p < 1.0f, "attention core: dropout_p must be in [0,1)");
  memset(a, 0, sizeof(*a));
  a->q = q; a->k = k; a->v = v; a->mask = mask;
  a->B = batch; a->Lq = q_len; a->Lk = k_len; a->H = hidden; a->heads = heads;
  a->scale = 1.0f / sqrtf((float)(hidden / heads));
  a->drop.seed = seed; a->drop.offset = offset; a->drop.p = p;
  a->drop.active = (training && p > 0.0f) ? 1 : 0;
  a->drop.scale = a->drop.active ? 1.0f / (1.0f - p) : 1.0f;
  a->drop.state = nullptr;
  return MSF_OK;
}

}  // namespace
}  // namespace msf

extern "C" {

int msf_attention_core_forward(const float* q, const float* k, const float* v, const float* mask,
                               int64_t batch, int32_t q_len, int32_t k_len, int32_t hidden, int32_t heads,
                               float dropout_p, int32_t training, uint64_t seed, uint64_t offset,
                               float* weights, float* out, void* stream) {
  msf::CoreArgs a;
  int rc = msf::fill(&a, q, k, v, mask, batch, q_len, k_len, hidden, heads, dropout_p, training, seed, offset);
  if (rc) return rc;
  MSF_REQUIRE(weights && out, "msf_attention_core_forward: null output");
  if (batch == 0) return MSF_OK;
  a.weights = weights; a.out = out;
  const long long warps = batch * heads * (long long)q_len;
  msf::core_fwd_kernel<<<(unsigned)msf::ceil_div(warps, 8), 256, 0, (cudaStream_t)stream>>>(a);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

int msf_attention_core_backward(const float* q, const float* k, const float* v, const float* mask,
                                int64_t batch, int32_t q_len, int32_t k_len, int32_t hidden, int32_t heads,
                                float dropout_p, int32_t training, uint64_t seed, uint64_t offset,
                                const float* weights, const float* grad_out, float* scratch, float* dq,
                                float* dk, float* dv, void* stream) {
  msf::CoreArgs a;
  int rc = msf::fill(&a, q, k, v, mask, batch, q_len, k_len, hidden, heads, dropout_p, training, seed, offset);
  if (rc) return rc;
  MSF_REQUIRE(weights && grad_out && scratch && dq && dk && dv, "msf_attention_core_backward: null pointer");
  if (batch == 0) return MSF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  a.weights = const_cast<float*>(weights); a.gout = grad_out; a.scratch = scratch;
  a.dq = dq; a.dk = dk; a.dv = dv;
  MSF_CHECK_CUDA(cudaMemsetAsync(dk, 0, sizeof(float) * (size_t)batch * k_len * hidden, st));
  MSF_CHECK_CUDA(cudaMemsetAsync(dv, 0, sizeof(float) * (size_t)batch * k_len * hidden, st));
  const long long warps = batch * heads * (long long)q_len;
  msf::core_bwd_kernel<<<(unsigned)msf::ceil_div(warps, 8), 256, 0, st>>>(a);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

}  // extern "C"
