// fp32 parity path of SequenceEncoder's LSTM recurrence (src/encoders.py:54-65,135-166: nn.LSTM, batch_first, optional
// packed windows, any number of layers — one call per layer), forward and backward, on the FFMA grouped GEMM
// (simt_gemm.cu) and two pointwise cell kernels.  This is what the drop-in runs by default (precision = "fp32",
// max-abs <= 1e-5 against the reference); the tensor-core path is lstm_seq.cu / lstm_bwd.cu.
//
//   forward :  z = x W_ih^T + b_ih for ALL steps in one GEMM, written into the gate buffer; per step
//              pre = z_t + h_{t-1} W_hh^T + b_hh (one GEMM), then the cell kernel turns pre into the gate activations
//              in place and writes c_t, h_t.  Rows whose window is over (t >= lengths[b]) carry (h, c) unchanged.
//   backward:  per step (walking back) the cell kernel turns d h_t, d c_t into d pre_t (over the gate activations, in
//              place) and d c_{t-1}; d h_{t-1} = d pre_t W_hh (one GEMM).  Afterwards d W_hh = sum d pre_t^T h_{t-1},
//              d W_ih = sum d pre_t^T x_t, d x = d pre W_ih as single GEMMs over all steps and d b as a column sum.
//
// Layouts: time-major, contiguous: x [T][B][F], h_seq / c_seq [T+1][B][H] (slice t = state BEFORE step t, slice 0 zeros on
// entry), gates [T][B][4H] with nn.LSTM's gate-major columns (i | f | g | o).
#include "simt_gemm.cuh"

namespace msf {
namespace {

__device__ __forceinline__ float f32_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// gates[b, :] holds z_t (input share + b_ih); zh[b, :] = h_{t-1} W_hh^T + b_hh
__global__ void lstm_f32_cell_fwd_kernel(float* __restrict__ gates, const float* __restrict__ zh, const float* __restrict__ c_prev,
                                         const float* __restrict__ h_prev, float* __restrict__ c_next, float* __restrict__ h_next,
                                         const int* __restrict__ lengths, int t, long long rows, int H) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * H) return;
  const long long b = e / H;
  const int u = (int)(e % H);
  if (lengths != nullptr && t >= lengths[b]) {   // the window is over: the state stands still (pack_padded_sequence)
    c_next[e] = c_prev[e];
    h_next[e] = h_prev[e];
    return;
  }
  float* g = gates + b * 4 * H;
  const float* z = zh + b * 4 * H;
  const float gi = f32_sigmoid(g[u] + z[u]);
  const float gf = f32_sigmoid(g[H + u] + z[H + u]);
  const float gg = tanhf(g[2 * H + u] + z[2 * H + u]);
  const float go = f32_sigmoid(g[3 * H + u] + z[3 * H + u]);
  const float c = gf * c_prev[e] + gi * gg;
  g[u] = gi; g[H + u] = gf; g[2 * H + u] = gg; g[3 * H + u] = go;
  c_next[e] = c;
  h_next[e] = go * tanhf(c);
}

// d h_t = dh_rec (from step t+1's GEMM) + dh_pass (carried past finished windows) + d_h_seq[t] + [t == last] d_h_last
__global__ void lstm_f32_cell_bwd_kernel(float* __restrict__ gates, const float* __restrict__ c_prev, const float* __restrict__ c_next,
                                         const float* __restrict__ dh_rec, float* __restrict__ dh_pass, float* __restrict__ dc,
                                         const float* __restrict__ d_h_seq_t, const float* __restrict__ d_h_last,
                                         const int* __restrict__ lengths, int t, int steps, long long rows, int H, int first) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * H) return;
  const long long b = e / H;
  const int u = (int)(e % H);
  const int len = lengths != nullptr ? lengths[b] : steps;
  float dh = first ? 0.0f : dh_rec[e] + dh_pass[e];
  if (d_h_seq_t != nullptr) dh += d_h_seq_t[e];
  if (d_h_last != nullptr && t == len - 1) dh += d_h_last[e];
  float* g = gates + b * 4 * H;
  if (t >= len) {   // no step was taken: the gradient of the carried state passes through
    g[u] = 0.0f; g[H + u] = 0.0f; g[2 * H + u] = 0.0f; g[3 * H + u] = 0.0f;
    dh_pass[e] = dh;
    return;
  }
  const float gi = g[u], gf = g[H + u], gg = g[2 * H + u], go = g[3 * H + u];
  const float tc = tanhf(c_next[e]);
  const float dcs = dc[e] + dh * go * (1.0f - tc * tc);
  g[u] = dcs * gg * gi * (1.0f - gi);
  g[H + u] = dcs * c_prev[e] * gf * (1.0f - gf);
  g[2 * H + u] = dcs * gi * (1.0f - gg * gg);
  g[3 * H + u] = dh * tc * go * (1.0f - go);
  dc[e] = dcs * gf;
  dh_pass[e] = 0.0f;
}

// ---- GRU (nn.GRU, gate order r | z | n; src/encoders.py:66-72) ---------------------------------------------------
// gates[b, 0:3H] holds the input's share z_x (+ b_ih); zh[b, 0:3H] = h_{t-1} W_hh^T + b_hh.  Afterwards
// gates[b, :] = (r | z | n | zh_n): what the backward pass needs.
__global__ void gru_f32_cell_fwd_kernel(float* __restrict__ gates, const float* __restrict__ zh, const float* __restrict__ h_prev,
                                        float* __restrict__ h_next, const int* __restrict__ lengths, int t, long long rows, int H) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * H) return;
  const long long b = e / H;
  const int u = (int)(e % H);
  if (lengths != nullptr && t >= lengths[b]) {
    h_next[e] = h_prev[e];
    return;
  }
  float* g = gates + b * 4 * H;
  const float* z = zh + b * 3 * H;
  const float gr = f32_sigmoid(g[u] + z[u]);
  const float gz = f32_sigmoid(g[H + u] + z[H + u]);
  const float hn = z[2 * H + u];
  const float gn = tanhf(g[2 * H + u] + gr * hn);
  g[u] = gr; g[H + u] = gz; g[2 * H + u] = gn; g[3 * H + u] = hn;
  h_next[e] = (1.0f - gz) * gn + gz * h_prev[e];
}

// gates[b, 0:3H] <- d z_x = (d r, d z, d n) (pre-activations); dzh[b, 0:3H] <- d z_h = (d r, d z, d n * r);
// dh_pass <- the part of d h_t that reaches h_{t-1} directly (d h_t * z; everything for a finished window)
__global__ void gru_f32_cell_bwd_kernel(float* __restrict__ gates, float* __restrict__ dzh, const float* __restrict__ h_prev,
                                        const float* __restrict__ dh_rec, float* __restrict__ dh_pass,
                                        const float* __restrict__ d_h_seq_t, const float* __restrict__ d_h_last,
                                        const int* __restrict__ lengths, int t, int steps, long long rows, int H, int first) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * H) return;
  const long long b = e / H;
  const int u = (int)(e % H);
  const int len = lengths != nullptr ? lengths[b] : steps;
  float dh = first ? 0.0f : dh_rec[e] + dh_pass[e];
  if (d_h_seq_t != nullptr) dh += d_h_seq_t[e];
  if (d_h_last != nullptr && t == len - 1) dh += d_h_last[e];
  float* g = gates + b * 4 * H;
  float* d = dzh + b * 3 * H;
  if (t >= len) {
    g[u] = 0.0f; g[H + u] = 0.0f; g[2 * H + u] = 0.0f;
    d[u] = 0.0f; d[H + u] = 0.0f; d[2 * H + u] = 0.0f;
    dh_pass[e] = dh;
    return;
  }
  const float gr = g[u], gz = g[H + u], gn = g[2 * H + u], hn = g[3 * H + u];
  const float dn = dh * (1.0f - gz) * (1.0f - gn * gn);
  const float dz = dh * (h_prev[e] - gn) * gz * (1.0f - gz);
  const float dr = dn * hn * gr * (1.0f - gr);
  g[u] = dr; g[H + u] = dz; g[2 * H + u] = dn;
  d[u] = dr; d[H + u] = dz; d[2 * H + u] = dn * gr;
  dh_pass[e] = dh * gz;
}

SimtProblem f32_problem() {
  SimtProblem p;
  memset(&p, 0, sizeof(p));
  p.nseg = 1; p.scale = 1.0f; p.head_dim = 1; p.heads = 1;
  return p;
}
DropCfg f32_no_drop() {
  DropCfg d;
  d.seed = 0; d.offset = 0; d.p = 0.f; d.scale = 1.f; d.active = 0; d.state = nullptr;
  return d;
}
// C[M,N] = A[M,K] . B[N,K]^T + bias (row-major, leading dimensions lda / ldb / ldc)
int f32_gemm_nt(const float* A, long long lda, const float* B, long long ldb, const float* bias, float* C, long long ldc,
                long long M, int N, int K, cudaStream_t st) {
  SimtProblem p = f32_problem();
  p.A[0] = A; p.a_rs = lda; p.a_cs = 1;
  p.B[0] = B; p.b_rs = ldb; p.b_cs = 1;
  p.bias[0] = bias;
  p.M = (int)M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc; p.epi = EPI_STORE;
  return simt_gemm_launch(&p, 1, f32_no_drop(), st);
}

}  // namespace
}  // namespace msf

extern "C" int msf_lstm_f32_forward(const float* x, int32_t in_dim, const float* w_ih, const float* w_hh, const float* b_ih,
                                    const float* b_hh, const int32_t* lengths, int64_t batch, int32_t steps, int32_t hidden,
                                    float* h_seq, float* c_seq, float* gates, float* scratch, void* stream) {
  using namespace msf;
  MSF_REQUIRE(x && w_ih && w_hh && h_seq && c_seq && gates && scratch, "msf_lstm_f32_forward: null pointer");
  MSF_REQUIRE(batch >= 1 && steps >= 1 && hidden >= 1 && in_dim >= 1 && batch * (long long)steps < (1ll << 31),
              "msf_lstm_f32_forward: bad batch / steps / hidden / in_dim");
  cudaStream_t st = (cudaStream_t)stream;
  const long long B = batch, H = hidden, N4 = 4LL * hidden, BH = B * H;
  int rc;
  // z = x W_ih^T + b_ih for all steps at once, into the gate buffer
  if ((rc = f32_gemm_nt(x, in_dim, w_ih, in_dim, b_ih, gates, N4, B * steps, (int)N4, in_dim, st))) return rc;
  const unsigned blocks = (unsigned)ceil_div(BH, 256);
  for (int t = 0; t < steps; ++t) {
    const float* h_prev = h_seq + (long long)t * BH;
    if ((rc = f32_gemm_nt(h_prev, H, w_hh, H, b_hh, scratch, N4, B, (int)N4, hidden, st))) return rc;
    lstm_f32_cell_fwd_kernel<<<blocks, 256, 0, st>>>(gates + (long long)t * B * N4, scratch, c_seq + (long long)t * BH, h_prev,
                                                     c_seq + (long long)(t + 1) * BH, h_seq + (long long)(t + 1) * BH,
                                                     lengths, t, B, hidden);
    MSF_LAUNCH_CHECK();
  }
  return MSF_OK;
}

extern "C" int msf_lstm_f32_backward(const float* x, int32_t in_dim, const float* w_ih, const float* w_hh,
                                     const int32_t* lengths, int64_t batch, int32_t steps, int32_t hidden,
                                     const float* h_seq, const float* c_seq, float* gates, const float* d_h_last,
                                     const float* d_h_seq, float* scratch, float* d_x, float* d_w_ih, float* d_w_hh,
                                     float* d_bias, void* stream) {
  using namespace msf;
  MSF_REQUIRE(x && w_ih && w_hh && h_seq && c_seq && gates && scratch && d_w_ih && d_w_hh, "msf_lstm_f32_backward: null pointer");
  MSF_REQUIRE(d_h_last || d_h_seq, "msf_lstm_f32_backward: no incoming gradient");
  MSF_REQUIRE(batch >= 1 && steps >= 1 && hidden >= 1 && in_dim >= 1 && batch * (long long)steps < (1ll << 31),
              "msf_lstm_f32_backward: bad batch / steps / hidden / in_dim");
  cudaStream_t st = (cudaStream_t)stream;
  const long long B = batch, H = hidden, N4 = 4LL * hidden, BH = B * H;
  float* dh_rec = scratch;            // [B][H]  d pre_{t+1} W_hh
  float* dh_pass = scratch + BH;      // [B][H]  gradient carried past finished windows
  float* dc = scratch + 2 * BH;       // [B][H]
  MSF_CHECK_CUDA(cudaMemsetAsync(dc, 0, sizeof(float) * (size_t)BH, st));
  const unsigned blocks = (unsigned)ceil_div(BH, 256);
  int rc;
  for (int t = steps - 1; t >= 0; --t) {
    float* g_t = gates + (long long)t * B * N4;
    lstm_f32_cell_bwd_kernel<<<blocks, 256, 0, st>>>(g_t, c_seq + (long long)t * BH, c_seq + (long long)(t + 1) * BH, dh_rec,
                                                     dh_pass, dc, d_h_seq ? d_h_seq + (long long)t * BH : nullptr, d_h_last,
                                                     lengths, t, steps, B, hidden, t == steps - 1 ? 1 : 0);
    MSF_LAUNCH_CHECK();
    if (t > 0) {   // d h_{t-1} (recurrent part) = d pre_t W_hh: A = d pre_t [B][4H], "B" = W_hh^T viewed as [H][4H]
      SimtProblem p = f32_problem();
      p.A[0] = g_t; p.a_rs = N4; p.a_cs = 1;
      p.B[0] = w_hh; p.b_rs = 1; p.b_cs = H;
      p.M = (int)B; p.N = hidden; p.K = (int)N4;
      p.C = dh_rec; p.ldc = H; p.epi = EPI_STORE;
      if ((rc = simt_gemm_launch(&p, 1, f32_no_drop(), st))) return rc;
    }
  }
  // weight gradients: contraction over all (t, window) rows
  const long long R = B * steps;
  SimtProblem probs[3];
  int n = 0;
  {
    SimtProblem p = f32_problem();   // d W_hh[o, k] = sum_r d pre[r, o] h_seq[r, k]   (h_seq slice t = h_{t-1})
    p.A[0] = gates; p.a_rs = 1; p.a_cs = N4;
    p.B[0] = h_seq; p.b_rs = 1; p.b_cs = H;
    p.M = (int)N4; p.N = hidden; p.K = (int)R;
    p.C = d_w_hh; p.ldc = H; p.epi = EPI_STORE;
    probs[n++] = p;
  }
  {
    SimtProblem p = f32_problem();   // d W_ih[o, f] = sum_r d pre[r, o] x[r, f]
    p.A[0] = gates; p.a_rs = 1; p.a_cs = N4;
    p.B[0] = x; p.b_rs = 1; p.b_cs = in_dim;
    p.M = (int)N4; p.N = in_dim; p.K = (int)R;
    p.C = d_w_ih; p.ldc = in_dim; p.epi = EPI_STORE;
    probs[n++] = p;
  }
  if ((rc = simt_gemm_launch(probs, n, f32_no_drop(), st))) return rc;
  if (d_x != nullptr) {   // d x[r, f] = sum_o d pre[r, o] W_ih[o, f]
    SimtProblem p = f32_problem();
    p.A[0] = gates; p.a_rs = N4; p.a_cs = 1;
    p.B[0] = w_ih; p.b_rs = 1; p.b_cs = in_dim;
    p.M = (int)R; p.N = in_dim; p.K = (int)N4;
    p.C = d_x; p.ldc = in_dim; p.epi = EPI_STORE;
    if ((rc = simt_gemm_launch(&p, 1, f32_no_drop(), st))) return rc;
  }
  if (d_bias != nullptr) {
    ColsumProblem c{gates, N4, (int)R, (int)N4, d_bias};
    if ((rc = colsum_launch(&c, 1, st))) return rc;
  }
  return MSF_OK;
}

// ---- GRU, same conventions; gates is [T][B][4H] ((r | z | n | zh_n) after the forward pass, d z_x in the first 3H columns
// after the backward pass), dzh [T][B][3H] receives d z_h in the backward pass -------------------------------------------
extern "C" int msf_gru_f32_forward(const float* x, int32_t in_dim, const float* w_ih, const float* w_hh, const float* b_ih,
                                   const float* b_hh, const int32_t* lengths, int64_t batch, int32_t steps, int32_t hidden,
                                   float* h_seq, float* gates, float* scratch, void* stream) {
  using namespace msf;
  MSF_REQUIRE(x && w_ih && w_hh && h_seq && gates && scratch, "msf_gru_f32_forward: null pointer");
  MSF_REQUIRE(batch >= 1 && steps >= 1 && hidden >= 1 && in_dim >= 1 && batch * (long long)steps < (1ll << 31),
              "msf_gru_f32_forward: bad batch / steps / hidden / in_dim");
  cudaStream_t st = (cudaStream_t)stream;
  const long long B = batch, H = hidden, N3 = 3LL * hidden, N4 = 4LL * hidden, BH = B * H;
  int rc;
  if ((rc = f32_gemm_nt(x, in_dim, w_ih, in_dim, b_ih, gates, N4, B * steps, (int)N3, in_dim, st))) return rc;
  const unsigned blocks = (unsigned)ceil_div(BH, 256);
  for (int t = 0; t < steps; ++t) {
    const float* h_prev = h_seq + (long long)t * BH;
    if ((rc = f32_gemm_nt(h_prev, H, w_hh, H, b_hh, scratch, N3, B, (int)N3, hidden, st))) return rc;
    gru_f32_cell_fwd_kernel<<<blocks, 256, 0, st>>>(gates + (long long)t * B * N4, scratch, h_prev,
                                                    h_seq + (long long)(t + 1) * BH, lengths, t, B, hidden);
    MSF_LAUNCH_CHECK();
  }
  return MSF_OK;
}

extern "C" int msf_gru_f32_backward(const float* x, int32_t in_dim, const float* w_ih, const float* w_hh,
                                    const int32_t* lengths, int64_t batch, int32_t steps, int32_t hidden, const float* h_seq,
                                    float* gates, float* dzh, const float* d_h_last, const float* d_h_seq, float* scratch,
                                    float* d_x, float* d_w_ih, float* d_w_hh, float* d_b_ih, float* d_b_hh, void* stream) {
  using namespace msf;
  MSF_REQUIRE(x && w_ih && w_hh && h_seq && gates && dzh && scratch && d_w_ih && d_w_hh, "msf_gru_f32_backward: null pointer");
  MSF_REQUIRE(d_h_last || d_h_seq, "msf_gru_f32_backward: no incoming gradient");
  MSF_REQUIRE(batch >= 1 && steps >= 1 && hidden >= 1 && in_dim >= 1 && batch * (long long)steps < (1ll << 31),
              "msf_gru_f32_backward: bad batch / steps / hidden / in_dim");
  cudaStream_t st = (cudaStream_t)stream;
  const long long B = batch, H = hidden, N3 = 3LL * hidden, N4 = 4LL * hidden, BH = B * H;
  float* dh_rec = scratch;         // [B][H]
  float* dh_pass = scratch + BH;   // [B][H]
  const unsigned blocks = (unsigned)ceil_div(BH, 256);
  int rc;
  for (int t = steps - 1; t >= 0; --t) {
    float* g_t = gates + (long long)t * B * N4;
    float* d_t = dzh + (long long)t * B * N3;
    gru_f32_cell_bwd_kernel<<<blocks, 256, 0, st>>>(g_t, d_t, h_seq + (long long)t * BH, dh_rec, dh_pass,
                                                    d_h_seq ? d_h_seq + (long long)t * BH : nullptr, d_h_last, lengths, t,
                                                    steps, B, hidden, t == steps - 1 ? 1 : 0);
    MSF_LAUNCH_CHECK();
    if (t > 0) {   // d h_{t-1} (through the recurrent weights) = d z_h W_hh
      SimtProblem p = f32_problem();
      p.A[0] = d_t; p.a_rs = N3; p.a_cs = 1;
      p.B[0] = w_hh; p.b_rs = 1; p.b_cs = H;
      p.M = (int)B; p.N = hidden; p.K = (int)N3;
      p.C = dh_rec; p.ldc = H; p.epi = EPI_STORE;
      if ((rc = simt_gemm_launch(&p, 1, f32_no_drop(), st))) return rc;
    }
  }
  const long long R = B * steps;
  SimtProblem probs[2];
  {
    SimtProblem p = f32_problem();   // d W_hh[o, k] = sum_r d z_h[r, o] h_seq[r, k]
    p.A[0] = dzh; p.a_rs = 1; p.a_cs = N3;
    p.B[0] = h_seq; p.b_rs = 1; p.b_cs = H;
    p.M = (int)N3; p.N = hidden; p.K = (int)R;
    p.C = d_w_hh; p.ldc = H; p.epi = EPI_STORE;
    probs[0] = p;
  }
  {
    SimtProblem p = f32_problem();   // d W_ih[o, f] = sum_r d z_x[r, o] x[r, f]
    p.A[0] = gates; p.a_rs = 1; p.a_cs = N4;
    p.B[0] = x; p.b_rs = 1; p.b_cs = in_dim;
    p.M = (int)N3; p.N = in_dim; p.K = (int)R;
    p.C = d_w_ih; p.ldc = in_dim; p.epi = EPI_STORE;
    probs[1] = p;
  }
  if ((rc = simt_gemm_launch(probs, 2, f32_no_drop(), st))) return rc;
  if (d_x != nullptr) {   // d x[r, f] = sum_o d z_x[r, o] W_ih[o, f]
    SimtProblem p = f32_problem();
    p.A[0] = gates; p.a_rs = N4; p.a_cs = 1;
    p.B[0] = w_ih; p.b_rs = 1; p.b_cs = in_dim;
    p.M = (int)R; p.N = in_dim; p.K = (int)N3;
    p.C = d_x; p.ldc = in_dim; p.epi = EPI_STORE;
    if ((rc = simt_gemm_launch(&p, 1, f32_no_drop(), st))) return rc;
  }
  ColsumProblem cs[2];
  int nc = 0;
  if (d_b_ih != nullptr) cs[nc++] = ColsumProblem{gates, N4, (int)R, (int)N3, d_b_ih};
  if (d_b_hh != nullptr) cs[nc++] = ColsumProblem{dzh, N3, (int)R, (int)N3, d_b_hh};
  if (nc > 0 && (rc = colsum_launch(cs, nc, st))) return rc;
  return MSF_OK;
}
