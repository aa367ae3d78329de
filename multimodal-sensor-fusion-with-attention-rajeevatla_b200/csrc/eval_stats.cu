// Evaluation epilogue in one pass over the logits (src/eval.py:39-130 evaluate_model; src/uncertainty.py:495-553):
// per window  softmax -> (confidence, first arg-max), NLL = lse - z[label]
// accumulated  confusion[label][pred] (accuracy and macro-F1 follow from it), the NLL sum, and the ECE /
//              reliability bins (count, correct, confidence sum) of src/uncertainty.py:113-126
// — what the reference gathers by concatenating every batch on the host and calling sklearn.  All accumulators are
// integers (NLL and confidences in fixed point), so the result does not depend on the order of the atomics and
// shards merge exactly with one integer all-reduce.  100 B of logits + 8 B of label per window: HBM-bound.
//
// One warp per window (lanes stride the classes); every block keeps its confusion matrix (classes <= 32) and bins in
// shared memory and flushes the non-zero cells once.
#include "msf_common.cuh"

namespace msf {

constexpr int EV_MAX_BINS = 128;
struct EvEdges {
  double e[EV_MAX_BINS + 1];
};

namespace {

constexpr int EV_THREADS = 256;
constexpr int EV_SMEM_CLASSES = 32;   // confusion matrix kept in shared memory up to this many classes

__device__ __forceinline__ int ev_bin_of(float cf, const double* __restrict__ e, int nb) {
  const double c = (double)cf;
  if (!(c >= e[0] && c <= e[nb])) return nb;  // NaN / out of range: dropped
  int g = (int)(cf * (float)nb);
  g = g < 0 ? 0 : (g > nb - 1 ? nb - 1 : g);
  while (g > 0 && c < e[g]) --g;               // bin i: e[i] <= c < e[i+1]
  while (g < nb - 1 && c >= e[g + 1]) ++g;     // last bin also takes c == e[nb]
  return g;
}

__global__ void __launch_bounds__(EV_THREADS) eval_stats_kernel(
    const float* __restrict__ logits, const int64_t* __restrict__ labels, long long B, int C,
    const __grid_constant__ EvEdges edges, int nb, float* __restrict__ conf_out, int64_t* __restrict__ pred_out,
    unsigned long long* __restrict__ confusion, unsigned long long* __restrict__ scalars,
    unsigned long long* __restrict__ bins) {
  __shared__ unsigned int cm[EV_SMEM_CLASSES * EV_SMEM_CLASSES];
  __shared__ unsigned int bcount[EV_MAX_BINS], bcorrect[EV_MAX_BINS];
  __shared__ unsigned long long bsum[EV_MAX_BINS];
  __shared__ unsigned long long nll_q24, n_rows, n_bad;
  __shared__ double e[EV_MAX_BINS + 1];
  const bool cm_smem = C <= EV_SMEM_CLASSES;
  for (int i = threadIdx.x; i < C * C && cm_smem; i += EV_THREADS) cm[i] = 0u;
  for (int i = threadIdx.x; i < nb; i += EV_THREADS) { bcount[i] = 0u; bcorrect[i] = 0u; bsum[i] = 0ull; }
  for (int i = threadIdx.x; i <= nb; i += EV_THREADS) e[i] = edges.e[i];
  if (threadIdx.x == 0) { nll_q24 = 0ull; n_rows = 0ull; n_bad = 0ull; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = EV_THREADS / 32;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < B; row += (long long)gridDim.x * wpb) {
    const float* z = logits + row * C;
    // first-max tie rule of torch.max; a NaN logit makes every probability NaN: torch.max then returns (NaN, 0)
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
      if (oi != 0x7fffffff && (idx == 0x7fffffff || ob > best || (ob == best && oi < idx))) { best = ob; idx = oi; }
    }
    float se = 0.0f;
    for (int c = lane; c < C; c += 32) se += expf(__ldg(z + c) - best);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if (lane == 0) {
      const long long y = labels[row];
      const bool y_ok = y >= 0 && y < (long long)C;
      const float cf = has_nan ? nanf("") : 1.0f / se;
      const int pr = has_nan ? 0 : idx;
      if (conf_out != nullptr) conf_out[row] = cf;
      if (pred_out != nullptr) pred_out[row] = pr;
      atomicAdd(&n_rows, 1ull);
      if (!y_ok) {
        atomicAdd(&n_bad, 1ull);
      } else {
        if (cm_smem) atomicAdd(&cm[(int)y * C + pr], 1u);
        else atomicAdd(confusion + y * C + pr, 1ull);
        if (!has_nan) {
          const float nll = best + logf(se) - __ldg(z + y);
          atomicAdd(&nll_q24, (unsigned long long)__double2ll_rn((double)fmaxf(nll, 0.0f) * 16777216.0));
        }
        const int b = ev_bin_of(cf, e, nb);
        if (b < nb) {
          atomicAdd(&bcount[b], 1u);
          if (pr == (int)y) atomicAdd(&bcorrect[b], 1u);
          atomicAdd(&bsum[b], (unsigned long long)__double2ll_rn((double)cf * 4294967296.0));
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C && cm_smem; i += EV_THREADS)
    if (cm[i]) atomicAdd(confusion + i, (unsigned long long)cm[i]);
  for (int i = threadIdx.x; i < nb; i += EV_THREADS)
    if (bcount[i]) {
      atomicAdd(bins + i, (unsigned long long)bcount[i]);
      atomicAdd(bins + nb + i, (unsigned long long)bcorrect[i]);
      atomicAdd(bins + 2 * nb + i, bsum[i]);
    }
  if (threadIdx.x == 0) {
    if (n_rows) atomicAdd(scalars + 0, n_rows);
    if (nll_q24) atomicAdd(scalars + 1, nll_q24);
    if (n_bad) atomicAdd(scalars + 2, n_bad);
  }
}

}  // namespace
}  // namespace msf

extern "C" int msf_eval_accumulate(const float* logits, const int64_t* labels, int64_t batch, int32_t classes,
                                   const double* edges, int32_t num_bins, float* conf, int64_t* pred,
                                   int64_t* confusion, int64_t* scalars, int64_t* bins, void* stream) {
  using namespace msf;
  MSF_REQUIRE(logits && labels && confusion && scalars && bins && edges && batch >= 0 && classes >= 1,
              "msf_eval_accumulate: bad arguments");
  MSF_REQUIRE(num_bins >= 1 && num_bins <= EV_MAX_BINS, "msf_eval_accumulate: num_bins %d out of range [1, %d]", num_bins,
              EV_MAX_BINS);
  for (int i = 0; i < num_bins; ++i) MSF_REQUIRE(edges[i] <= edges[i + 1], "msf_eval_accumulate: edges must be ascending");
  if (batch == 0) return MSF_OK;
  EvEdges ed;
  for (int i = 0; i <= num_bins; ++i) ed.e[i] = edges[i];
  long long blocks = ceil_div(batch, EV_THREADS / 32);
  if (blocks > 148 * 8) blocks = 148 * 8;
  eval_stats_kernel<<<(unsigned)blocks, EV_THREADS, 0, (cudaStream_t)stream>>>(
      logits, labels, batch, classes, ed, num_bins, conf, pred, reinterpret_cast<unsigned long long*>(confusion),
      reinterpret_cast<unsigned long long*>(scalars), reinterpret_cast<unsigned long long*>(bins));
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}
