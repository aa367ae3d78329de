// LSTM recurrence of SequenceEncoder (src/encoders.py:54-65,135-166: nn.LSTM(F, H, batch_first), one layer, no
// packing) on the tensor cores: per time step ONE grouped-GEMM launch computes, for every sequence encoder,
//
//     pre = h_{t-1} W_hh^T + x_t W_ih^T + (b_ih + b_hh)        tcgen05.mma, H/64 + 1 K-segments summed in TMEM
//     c_t = f*c_{t-1} + i*g,   h_t = o*tanh(c_t)                in the epilogue (TC_EPI_LSTM, tc_gemm.cu)
//
// with the weight rows gate-interleaved (row 4u+g = gate g of unit u) so that the four gates of a unit sit in
// adjacent accumulator columns of one thread.  h ping-pongs between two bf16 buffers laid out k-block-major
// ([H/64][B][64]: exactly the A operand the next step loads by TMA), c stays fp32 in place.  Every launch is a
// programmatic dependent of the previous one; the caller can capture the whole sequence in a CUDA graph.
#include <string.h>

#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace msf {
bool lstm_seq_eligible(int hidden, int n, long long batch, int sms);   // lstm_seq.cu: one persistent launch
int lstm_seq_launch(const msf_lstm_seq* seqs, int n, long long batch, int steps, int hidden, cudaStream_t st);
}  // namespace msf

namespace msf {
namespace {
// (B, T, F) fp32 windows -> the recurrence's A operand [T][B][64] bf16: features zero-padded to 64 columns, optionally
// 1.0 in column F (the ones of the bias gradient).  One thread per (t, window): a full 128-byte line written with four
// 256-bit stores, windows fastest so that a warp writes 4 KB contiguously.
__global__ void __launch_bounds__(256) lstm_pack_input_kernel(const float* __restrict__ x, long long B, int T, int F, int ones,
                                                              __nv_bfloat16* __restrict__ out) {
  const long long total = B * T;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long t = e / B, b = e % B;
    const float* src = x + (b * T + t) * F;
    __nv_bfloat16* dst = out + e * 64;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = 16 * q + 2 * k;
        const float lo = c < F ? __ldg(src + c) : (ones && c == F ? 1.0f : 0.0f);
        const float hi = c + 1 < F ? __ldg(src + c + 1) : (ones && c + 1 == F ? 1.0f : 0.0f);
        __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
        w[k] = *reinterpret_cast<uint32_t*>(&p);
      }
      st_global_v8(dst + 16 * q, w);
    }
  }
}
}  // namespace
}  // namespace msf

extern "C" int msf_lstm_pack_input(const float* x, int64_t batch, int32_t steps, int32_t features, int32_t ones_column,
                                   void* out_bf16, void* stream) {
  using namespace msf;
  MSF_REQUIRE(x && out_bf16 && batch >= 1 && steps >= 1 && features >= 1 && features <= 64,
              "msf_lstm_pack_input: bad arguments (1..64 features)");
  MSF_REQUIRE((reinterpret_cast<uintptr_t>(out_bf16) & 31) == 0, "msf_lstm_pack_input: output not 32-byte aligned");
  long long blocks = ceil_div((long long)batch * steps, 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  lstm_pack_input_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, batch, steps, features, ones_column ? 1 : 0,
                                                                            static_cast<__nv_bfloat16*>(out_bf16));
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

extern "C" int msf_lstm_forward(const msf_lstm_seq* seqs, int32_t n, int64_t batch, int32_t steps, int32_t hidden,
                                void* stream) {
  using namespace msf;
  MSF_REQUIRE(seqs != nullptr && n >= 1 && n <= MSF_LSTM_MAX_SEQS, "msf_lstm_forward: 1..%d sequences per call", MSF_LSTM_MAX_SEQS);
  MSF_REQUIRE(batch >= 1 && batch < (1 << 24) && steps >= 1, "msf_lstm_forward: bad batch / steps");
  MSF_REQUIRE(hidden % 64 == 0 && hidden >= 64 && hidden <= 384, "msf_lstm_forward: hidden %d (needs a multiple of 64, <= 384)", hidden);
  const int KBH = hidden / 64;
  MSF_REQUIRE(KBH + 1 <= TC_MAX_SEG, "msf_lstm_forward: hidden %d needs too many K-segments", hidden);
  cudaStream_t st = (cudaStream_t)stream;
  {   // hidden <= 256: the whole sequence as one persistent launch, weights resident in shared memory (lstm_seq.cu);
      // MSF_LSTM_STEPS=1 keeps the launch-per-step path below
    static int sms = 0;
    if (sms == 0) {
      int dev = 0;
      MSF_CHECK_CUDA(cudaGetDevice(&dev));
      MSF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (lstm_seq_eligible(hidden, n, batch, sms)) return lstm_seq_launch(seqs, n, batch, steps, hidden, st);
    for (int i = 0; i < n; ++i)
      if (seqs[i].lengths != nullptr || seqs[i].h_all != nullptr || seqs[i].z_in != nullptr || seqs[i].cell_type != 0) {
        set_error("msf_lstm_forward: per-window lengths, stacked layers, the GRU cell and the training mode need the persistent kernel (hidden <= 256, MSF_LSTM_STEPS unset)");
        return MSF_E_UNSUPPORTED;
      }
  }
  const long long B = batch, N4 = 4LL * hidden, slice = B * 64;
  // two launch descriptions: even steps read h_a / write h_b, odd steps the other way round
  TcBuilder even(false, 256, no_dropout(), st, "LSTM step"), odd(false, 256, no_dropout(), st, "LSTM step");
  for (int i = 0; i < n; ++i) {
    const msf_lstm_seq& S = seqs[i];
    MSF_REQUIRE(S.x_bf16 && S.w_hh && S.w_ih && S.bias && S.h_a && S.h_b && S.cell && S.h_out,
                "msf_lstm_forward: null pointer in sequence %d", i);
    for (int par = 0; par < 2; ++par) {
      TcBuilder& tb = par ? odd : even;
      const void* h_in = par ? S.h_b : S.h_a;
      void* h_next = par ? S.h_a : S.h_b;
      const short m_h = (short)tb.add_map(h_in, B, 64, 64, KBH, slice, TC_BLOCK_M);
      const short m_whh = (short)tb.add_map(S.w_hh, N4, 64, 64, KBH, N4 * 64, 256);
      const short m_x = (short)tb.add_map(S.x_bf16, B, 64, 64, steps, slice, TC_BLOCK_M);
      const short m_wih = (short)tb.add_map(S.w_ih, N4, 64, 64, 1, 0, 256);
      TcProblem p = tc_blank_problem();
      for (int kb = 0; kb < KBH; ++kb) {
        p.seg[kb].a_map = m_h; p.seg[kb].a_z = kb;
        p.seg[kb].b_map = m_whh; p.seg[kb].b_z = kb;
      }
      p.seg[KBH].a_map = m_x; p.seg[KBH].a_z = 0;
      p.seg[KBH].b_map = m_wih; p.seg[KBH].b_z = 0;
      p.nseg = KBH + 1;
      p.bias[0] = S.bias;
      p.M = (int)B; p.N = (int)N4; p.K = 64;
      p.C = h_next; p.ldc = 64; p.c_bf16 = 1;
      p.epi = TC_EPI_LSTM;
      p.cell = S.cell; p.h32 = nullptr; p.h_slice = slice;
      int rc = tb.add_problem(p);
      if (rc) return rc;
    }
  }
  if (even.status) return even.status;
  if (odd.status) return odd.status;
  for (int t = 0; t < steps; ++t) {
    TcBuilder& tb = (t & 1) ? odd : even;
    for (int i = 0; i < n; ++i) {
      tb.L.p[i].seg[KBH].a_z = t;                                    // x_t
      tb.L.p[i].h32 = (t == steps - 1) ? seqs[i].h_out : nullptr;    // h_T in fp32 for the caller
    }
    int rc = tb.flush(true);
    if (rc) return rc;
  }
  return MSF_OK;
}
