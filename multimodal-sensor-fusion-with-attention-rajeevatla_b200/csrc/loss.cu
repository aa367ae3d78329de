// Cross-entropy with label smoothing (fwd + grad) and softmax -> (confidence, prediction).
//   src/train.py:185-186,310   nn.CrossEntropyLoss(label_smoothing=s), mean reduction
//   src/eval.py:89-90, src/train.py:341-342   probs = softmax(logits); conf, pred = max(probs)
// One warp per window; classes are strided over lanes.
#include <math.h>

#include "msf_common.cuh"

namespace msf {
namespace {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// blocks that have finished their rows; the last one to arrive reduces row_loss (fixed order:
// deterministic) and resets the ticket, so the mean needs no second launch
__device__ unsigned int g_ce_ticket = 0;

__device__ unsigned long long g_bad_labels = 0ull;   // rows whose label was outside [0, C) since the last reset

__device__ void ce_rows(const float* __restrict__ logits, const int64_t* __restrict__ labels, long long B, int C,
                        float smoothing, float grad_scale, float* __restrict__ row_loss, float* __restrict__ grad) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* z = logits + row * C;
  float mx = -INFINITY, zsum = 0.0f;
  for (int c = lane; c < C; c += 32) {
    const float v = __ldg(z + c);
    mx = fmaxf(mx, v);
    zsum += v;
  }
  mx = warp_max(mx);
  zsum = warp_sum(zsum);
  float se = 0.0f;
  for (int c = lane; c < C; c += 32) se += expf(__ldg(z + c) - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  // a label outside [0, C) (e.g. an ignore_index of -100) must not index the logits: the row is treated as having
  // no target class (loss = lse, gradient = softmax - smoothing / C) and counted (msf_bad_label_count)
  const long long y64 = labels[row];
  const bool y_ok = y64 >= 0 && y64 < (long long)C;
  const int y = y_ok ? (int)y64 : -1;
  if (lane == 0 && !y_ok) atomicAdd(&g_bad_labels, 1ull);
  if (lane == 0 && row_loss) {
    const float nll = lse - (y_ok ? __ldg(z + y) : 0.0f);
    const float smooth = lse - zsum / (float)C;  // mean_c(-log p_c)
    row_loss[row] = (1.0f - smoothing) * nll + smoothing * smooth;
  }
  if (grad) {
    const float inv = 1.0f / se;
    const float off = smoothing / (float)C;
    for (int c = lane; c < C; c += 32) {
      const float p = expf(__ldg(z + c) - mx) * inv;
      const float t = off + ((c == y) ? (1.0f - smoothing) : 0.0f);
      grad[row * C + c] = (p - t) * grad_scale;
    }
  }
}

__global__ void __launch_bounds__(256) ce_kernel(const float* __restrict__ logits,
                                                 const int64_t* __restrict__ labels, long long B, int C,
                                                 float smoothing, float grad_scale,
                                                 float* __restrict__ row_loss, float* __restrict__ grad,
                                                 float* __restrict__ loss_out) {
  ce_rows(logits, labels, B, C, smoothing, grad_scale, row_loss, grad);
  if (loss_out == nullptr) return;
  __shared__ bool last;
  __shared__ double sh[256];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&g_ce_ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  double s = 0.0;
  for (long long i = threadIdx.x; i < B; i += blockDim.x) s += (double)__ldcg(row_loss + i);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    loss_out[0] = (float)(sh[0] / (double)B);
    g_ce_ticket = 0;
  }
}

__global__ void __launch_bounds__(256) conf_pred_kernel(const float* __restrict__ logits, long long B, int C,
                                                        float* __restrict__ conf, int64_t* __restrict__ pred) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* z = logits + row * C;
  // first-max tie rule of torch.max.  A NaN logit makes every probability NaN:
  // torch.max then returns (NaN, 0).
  float best = -INFINITY;
  int idx = 0x7fffffff;
  bool has_nan = false;
  for (int c = lane; c < C; c += 32) {
    const float v = __ldg(z + c);
    has_nan |= (v != v);
    if (v > best || idx == 0x7fffffff) { best = v; idx = c; }
  }
  has_nan = __any_sync(0xffffffffu, has_nan);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (oi != 0x7fffffff && (idx == 0x7fffffff || ob > best || (ob == best && oi < idx))) {
      best = ob;
      idx = oi;
    }
  }
  if (has_nan) {
    if (lane == 0) { conf[row] = nanf(""); pred[row] = 0; }
    return;
  }
  float se = 0.0f;
  for (int c = lane; c < C; c += 32) se += expf(__ldg(z + c) - best);
  se = warp_sum(se);
  if (lane == 0) {
    conf[row] = 1.0f / se;  // exp(best - best) / sum = max softmax probability
    pred[row] = idx;
  }
}

}  // namespace
}  // namespace msf

extern "C" {

int msf_cross_entropy(const float* logits, const int64_t* labels, int64_t batch, int32_t classes,
                      float smoothing, float grad_scale, float* row_loss, float* loss_out,
                      float* grad_logits, void* stream) {
  MSF_REQUIRE(logits && labels && batch > 0 && classes > 0, "msf_cross_entropy: bad arguments");
  MSF_REQUIRE(loss_out == nullptr || row_loss != nullptr, "msf_cross_entropy: loss_out needs row_loss scratch");
  cudaStream_t st = (cudaStream_t)stream;
  msf::ce_kernel<<<(unsigned)msf::ceil_div(batch, 8), 256, 0, st>>>(logits, labels, batch, classes, smoothing,
                                                                     grad_scale, row_loss, grad_logits, loss_out);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

int msf_softmax_conf_pred(const float* logits, int64_t batch, int32_t classes, float* conf, int64_t* pred,
                          void* stream) {
  MSF_REQUIRE(logits && conf && pred && batch > 0 && classes > 0, "msf_softmax_conf_pred: bad arguments");
  msf::conf_pred_kernel<<<(unsigned)msf::ceil_div(batch, 8), 256, 0, (cudaStream_t)stream>>>(logits, batch,
                                                                                            classes, conf, pred);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

}  // extern "C"


extern "C" int msf_bad_label_count(int64_t* count, int32_t reset) {
  using namespace msf;
  MSF_REQUIRE(count != nullptr, "msf_bad_label_count: null output");
  unsigned long long v = 0ull;
  MSF_CHECK_CUDA(cudaMemcpyFromSymbol(&v, g_bad_labels, sizeof(v)));
  if (reset) {
    const unsigned long long zero = 0ull;
    MSF_CHECK_CUDA(cudaMemcpyToSymbol(g_bad_labels, &zero, sizeof(zero)));
  }
  *count = (int64_t)v;
  return MSF_OK;
}
