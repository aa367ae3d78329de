// ECE / MCE / reliability-diagram binning in one streaming pass.
//   src/uncertainty.py:113-126 (fp32 torch.linspace edges), :231-241 (f64 np.linspace edges)
// The reference makes one boolean-mask pass per bin (15 passes over the data on
// the host); here each sample is read once: 20 B/sample, HBM-bound.
//
// Layout of the per-block histogram: every thread owns a private column
// [bin][tid] in shared memory, so updates are plain LDS/STS with no bank
// conflicts and no atomics in the streaming loop (shared-memory atomics would
// cap the kernel far below HBM speed).  Count and correct share one u32
// (16 bits each, flushed before they can overflow); the confidence sum is kept
// in Q32 fixed point so the final integer atomics give an order-independent,
// bit-reproducible result that shards merge exactly.
#include "msf_common.cuh"

namespace msf {

constexpr int ECE_ITERS_PER_FLUSH = 8000;  // x (4 samples x 2 unroll) < 65535 per thread

constexpr int ECE_MAX_BINS = 128;
struct EceEdges {
  double e[ECE_MAX_BINS + 1];  // by value: no device allocation, CUDA-graph capturable
};

__device__ __forceinline__ int bin_of(float cf, const double* __restrict__ e, int nb, float nbf) {
  const double c = (double)cf;
  if (!(c >= e[0] && c <= e[nb])) return nb;  // NaN / out of range: garbage column
  int g = (int)(cf * nbf);
  g = g < 0 ? 0 : (g > nb - 1 ? nb - 1 : g);
  while (g > 0 && c < e[g]) --g;               // bin i: e[i] <= c < e[i+1]
  while (g < nb - 1 && c >= e[g + 1]) ++g;     // last bin also takes c == e[nb]
  return g;
}

template <int T>
__global__ void __launch_bounds__(T) ece_bin_kernel(const float* __restrict__ conf,
                                                    const int64_t* __restrict__ pred,
                                                    const int64_t* __restrict__ label, long long n,
                                                    const __grid_constant__ EceEdges edges, int nb,
                                                    int64_t* __restrict__ count, int64_t* __restrict__ correct,
                                                    unsigned long long* __restrict__ conf_sum, int vec_ok) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* cs = reinterpret_cast<unsigned long long*>(smem_raw);          // [nb+1][T]
  unsigned int* cc = reinterpret_cast<unsigned int*>(cs + (size_t)(nb + 1) * T);      // [nb+1][T]
  double* e = reinterpret_cast<double*>(cc + (size_t)(nb + 1) * T);                   // [nb+1]
  const int tid = threadIdx.x;
  for (int i = tid; i <= nb; i += T) e[i] = edges.e[i];
  for (int i = tid; i < (nb + 1) * T; i += T) { cs[i] = 0ull; cc[i] = 0u; }
  __syncthreads();
  const float nbf = (float)nb;

  auto add = [&](float cf, long long p, long long l) {
    const int b = bin_of(cf, e, nb, nbf);
    const int idx = b * T + tid;
    cc[idx] += 1u + ((p == l) ? 65536u : 0u);
    cs[idx] += (unsigned long long)__double2ll_rn((double)cf * 4294967296.0);
  };
  auto flush = [&]() {
    __syncthreads();
    for (int b = tid >> 5; b < nb; b += (T >> 5)) {
      unsigned long long s = 0ull, cnt = 0ull, cor = 0ull;
      for (int t = tid & 31; t < T; t += 32) {
        const unsigned int v = cc[b * T + t];
        cnt += v & 0xffffu;
        cor += v >> 16;
        s += cs[b * T + t];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        cor += __shfl_xor_sync(0xffffffffu, cor, o);
      }
      if ((tid & 31) == 0 && cnt) {
        atomicAdd(reinterpret_cast<unsigned long long*>(count + b), cnt);
        atomicAdd(reinterpret_cast<unsigned long long*>(correct + b), cor);
        atomicAdd(conf_sum + b, s);
      }
    }
    __syncthreads();
    for (int i = tid; i < (nb + 1) * T; i += T) { cs[i] = 0ull; cc[i] = 0u; }
    __syncthreads();
  };

  const long long nquad = vec_ok ? (n >> 2) : 0;
  const long long stride = (long long)gridDim.x * T;
  long long q = blockIdx.x * (long long)T + tid;
  const float4* c4 = reinterpret_cast<const float4*>(conf);
  const longlong2* p2 = reinterpret_cast<const longlong2*>(pred);
  const longlong2* l2 = reinterpret_cast<const longlong2*>(label);
  int iters = 0;
  // block-uniform trip count so the periodic flush can use __syncthreads
  const long long q_base = blockIdx.x * (long long)T;
  for (long long qb = q_base; qb < nquad; qb += 2 * stride, q += 2 * stride) {
    const long long q1 = q + stride;
    const bool ok0 = q < nquad, ok1 = q1 < nquad;
    float4 c0 = make_float4(0, 0, 0, 0), c1 = c0;
    longlong2 pa0 = {0, 0}, pb0 = {0, 0}, la0 = {0, 0}, lb0 = {0, 0};
    longlong2 pa1 = {0, 0}, pb1 = {0, 0}, la1 = {0, 0}, lb1 = {0, 0};
    if (ok0) {  // issue every load before the first use
      c0 = __ldcs(c4 + q);
      pa0 = __ldcs(p2 + 2 * q); pb0 = __ldcs(p2 + 2 * q + 1);
      la0 = __ldcs(l2 + 2 * q); lb0 = __ldcs(l2 + 2 * q + 1);
    }
    if (ok1) {
      c1 = __ldcs(c4 + q1);
      pa1 = __ldcs(p2 + 2 * q1); pb1 = __ldcs(p2 + 2 * q1 + 1);
      la1 = __ldcs(l2 + 2 * q1); lb1 = __ldcs(l2 + 2 * q1 + 1);
    }
    if (ok0) {
      add(c0.x, pa0.x, la0.x); add(c0.y, pa0.y, la0.y);
      add(c0.z, pb0.x, lb0.x); add(c0.w, pb0.y, lb0.y);
    }
    if (ok1) {
      add(c1.x, pa1.x, la1.x); add(c1.y, pa1.y, la1.y);
      add(c1.z, pb1.x, lb1.x); add(c1.w, pb1.y, lb1.y);
    }
    if (++iters == ECE_ITERS_PER_FLUSH) { flush(); iters = 0; }
  }
  // scalar remainder (or everything, when the pointers are not 16-byte aligned)
  flush();
  iters = 0;
  const long long start = nquad << 2;
  for (long long ib = start + blockIdx.x * (long long)T; ib < n; ib += stride) {
    const long long i = ib + tid;
    if (i < n) add(__ldcs(conf + i), __ldcs(pred + i), __ldcs(label + i));
    if (++iters == 60000) { flush(); iters = 0; }
  }
  flush();
}

}  // namespace msf

extern "C" int msf_ece_bin(const float* conf, const int64_t* pred, const int64_t* label, int64_t n,
                           const double* edges, int32_t num_bins, int64_t* count, int64_t* correct,
                           uint64_t* conf_sum_q32, void* stream) {
  MSF_REQUIRE(num_bins >= 1 && num_bins <= msf::ECE_MAX_BINS, "msf_ece_bin: num_bins %d out of range [1, %d]", num_bins,
              msf::ECE_MAX_BINS);
  MSF_REQUIRE(n >= 0 && edges && count && correct && conf_sum_q32, "msf_ece_bin: bad arguments");
  for (int i = 0; i < num_bins; ++i)
    MSF_REQUIRE(edges[i] <= edges[i + 1], "msf_ece_bin: edges must be ascending");
  if (n == 0) return MSF_OK;
  MSF_REQUIRE(conf && pred && label, "msf_ece_bin: null input");
  cudaStream_t st = (cudaStream_t)stream;

  msf::EceEdges ed;
  for (int i = 0; i <= num_bins; ++i) ed.e[i] = edges[i];

  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  auto smem_for = [&](int T) { return (size_t)(num_bins + 1) * T * 12 + (size_t)(num_bins + 1) * 8; };
  int T = 256;
  while (T > 32 && smem_for(T) > 100 * 1024) T >>= 1;
  const size_t smem = smem_for(T);
  MSF_REQUIRE(smem <= 200 * 1024, "msf_ece_bin: num_bins %d needs too much shared memory", num_bins);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(conf) | reinterpret_cast<uintptr_t>(pred) |
                       reinterpret_cast<uintptr_t>(label)) & 15) == 0;
  const long long work = (n + 7) / 8;
  long long blocks = (work + T - 1) / T;
  const long long max_blocks = (long long)sms * (smem > 56 * 1024 ? 2 : 4);
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  cudaError_t e = cudaSuccess;
#define MSF_ECE_LAUNCH(TT)                                                                              \
  do {                                                                                                  \
    e = cudaFuncSetAttribute(msf::ece_bin_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                             (int)smem);                                                                \
    if (e == cudaSuccess)                                                                               \
      msf::ece_bin_kernel<TT><<<(unsigned)blocks, TT, smem, st>>>(                                      \
          conf, pred, label, n, ed, num_bins, count, correct,                                          \
          reinterpret_cast<unsigned long long*>(conf_sum_q32), vec_ok);                                 \
  } while (0)
  if (T == 256) MSF_ECE_LAUNCH(256);
  else if (T == 128) MSF_ECE_LAUNCH(128);
  else if (T == 64) MSF_ECE_LAUNCH(64);
  else MSF_ECE_LAUNCH(32);
#undef MSF_ECE_LAUNCH
  MSF_CHECK_CUDA(e);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}
