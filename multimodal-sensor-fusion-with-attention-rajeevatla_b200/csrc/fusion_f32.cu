// HybridFusion forward / backward, MSF_PREC_F32 (fp32 FFMA) path.
//
// Restates src/fusion.py:331-479 + src/attention.py:68-146 stage by stage
// (SURVEY.md §8 a-2..a-4).  Inside HybridFusion every CrossModalAttention call
// has q_len = k_len = 1, so its softmax is the per-(window, head) gate
// g = 1[mask_k != 0] * dropout and the query/key projections are dead: they are
// never evaluated and their gradient slots are written as exact zeros.
//
// Stages (each one launch of the grouped GEMM in simt_gemm.cu unless noted):
//   F0 prep      xt_m  = drop0(x_m * mask_m)                       [elementwise]
//   F1 proj      P_m   = drop1(relu(xt_m Wp_m^T + bp_m))
//   F2 value     U_qk  = g_qk * (P_k Wv_qk^T + bv_qk)              (records g)
//   F3 out+mean  agg_q = (P_q + sum_k (U_qk Wo_qk^T + bo_qk)) / cnt_q * mask_q
//   F4 tail      s_q = agg_q.wg_q + bg_q; masked softmax + fallbacks; fused = sum w_q agg_q  [warp/row]
//   F5 cls1      Hr    = drop3(relu(fused W1^T + b1))
//   F6 cls2      logits = Hr W2^T + b2
// Backward mirrors it (B1..B9 below).
#include "fusion_common.cuh"
#include "simt_gemm.cuh"

namespace msf {

// ---------------------------------------------------------------------------
// workspace carve-up (fp32 elements), shared by forward and backward
// ---------------------------------------------------------------------------
struct WsF32 {
  float* xt[MSF_MAX_MODALITIES];
  float* P[MSF_MAX_MODALITIES];
  float* U;      // [pairs][B][H]
  float* G;      // [pairs][B][heads]
  float* agg;    // [M][B][H]
  float* soft;   // [B][M] softmax before mask/renorm
  float* w;      // [B][M] fusion weights
  float* fused;  // [B][H]
  float* Hr;     // [B][H]
  // backward temporaries
  float* dH1;    // [B][H]
  float* dfused; // [B][H]
  float* dS;     // [M][B][H]
  float* dV;     // [pairs][B][H]
  float* dZ;     // [M][B][H]
  size_t bytes;
};

static void carve_f32(const Layout& L, int64_t B, void* base, WsF32* ws) {
  size_t off = 0;
  auto take = [&](size_t n) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += align_up(n * sizeof(float), 256);
    return p;
  };
  const size_t BH = (size_t)B * L.H;
  for (int m = 0; m < L.M; ++m) ws->xt[m] = take((size_t)B * L.D[m]);
  for (int m = 0; m < L.M; ++m) ws->P[m] = take(BH);
  ws->U = take(BH * L.num_pairs());
  ws->G = take((size_t)B * L.heads * L.num_pairs());
  ws->agg = take(BH * L.M);
  ws->soft = take((size_t)B * L.M);
  ws->w = take((size_t)B * L.M);
  ws->fused = take(BH);
  ws->Hr = take(BH);
  ws->dH1 = take(BH);
  ws->dfused = take(BH);
  ws->dS = take(BH * L.M);
  ws->dV = take(BH * L.num_pairs());
  ws->dZ = take(BH * L.M);
  ws->bytes = off;
}

size_t fusion_f32_workspace_bytes(const Layout& L, int64_t B) {
  WsF32 ws;
  carve_f32(L, B, nullptr, &ws);
  return ws.bytes;
}

// ---------------------------------------------------------------------------
// F0: xt = drop0(x * mask)           (fusion.py:370-374)
// ---------------------------------------------------------------------------
struct PrepArgs {
  const float* x[MSF_MAX_MODALITIES];
  float* xt[MSF_MAX_MODALITIES];
  int D[MSF_MAX_MODALITIES];
  int M;
  long long B;
  const float* mask;
  DropCfg drop;
};

__global__ void __launch_bounds__(256) prep_kernel(const __grid_constant__ PrepArgs a) {
  const int m = blockIdx.y;
  const int D = a.D[m];
  const DropCfg drop = resolve_drop(a.drop);
  const int quads = (D + 3) >> 2;
  const long long total = a.B * quads;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / quads;
    const int c4 = (int)(i % quads);
    const float mk = a.mask ? __ldg(a.mask + row * a.M + m) : 1.0f;
    float dm[4] = {1.f, 1.f, 1.f, 1.f};
    if (drop.active) drop4(drop, SITE_INPUT, m, row, c4, dm);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = c4 * 4 + j;
      if (col < D) a.xt[m][row * D + col] = __ldg(a.x[m] + row * D + col) * mk * dm[j];
    }
  }
}

// ---------------------------------------------------------------------------
// F4: gating, masked softmax with fallbacks, weighted sum   (fusion.py:410-418, 429-479)
// one warp per window
// ---------------------------------------------------------------------------
struct TailArgs {
  const float* agg;    // [M][B][H]
  const float* gate_w[MSF_MAX_MODALITIES];
  const float* gate_b[MSF_MAX_MODALITIES];
  const float* mask;   // (B, M) or null
  float* soft;         // (B, M)
  float* w;            // (B, M)
  float* w_out;        // optional user copy
  float* fused;        // (B, H)
  long long B;
  int M, H;
};

__global__ void __launch_bounds__(256) tail_fwd_kernel(const __grid_constant__ TailArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.B) return;
  float s[MSF_MAX_MODALITIES], mk[MSF_MAX_MODALITIES], soft[MSF_MAX_MODALITIES], w[MSF_MAX_MODALITIES];
  for (int q = 0; q < a.M; ++q) {
    const float* ag = a.agg + ((long long)q * a.B + row) * a.H;
    float part = 0.0f;
    for (int c = lane; c < a.H; c += 32) part = fmaf(__ldg(ag + c), __ldg(a.gate_w[q] + c), part);
    s[q] = warp_sum(part) + __ldg(a.gate_b[q]);  // fusion.py:459
    mk[q] = a.mask ? __ldg(a.mask + row * a.M + q) : 1.0f;
  }
  adaptive_weights_row(s, mk, a.M, soft, w);
  if (lane == 0)
    for (int q = 0; q < a.M; ++q) {
      if (a.soft) a.soft[row * a.M + q] = soft[q];
      if (a.w) a.w[row * a.M + q] = w[q];
      if (a.w_out) a.w_out[row * a.M + q] = w[q];
    }
  if (a.fused == nullptr) return;
  for (int c = lane; c < a.H; c += 32) {
    float f = 0.0f;
    for (int q = 0; q < a.M; ++q) f = fmaf(__ldg(a.agg + ((long long)q * a.B + row) * a.H + c), w[q], f);
    a.fused[row * a.H + c] = f;  // fusion.py:416-418
  }
}

// ---------------------------------------------------------------------------
// B4: backward of F4.  dfused -> dS_q (already scaled by mask_q / cnt_q), gating grads.
// ---------------------------------------------------------------------------
struct TailBwdArgs {
  const float* agg;
  const float* dfused;
  const float* gate_w[MSF_MAX_MODALITIES];
  const float* mask;
  const float* soft;
  const float* w;
  float* dS;                                // [M][B][H]
  float* d_gate_w[MSF_MAX_MODALITIES];      // accumulated with atomics (pre-zeroed)
  float* d_gate_b[MSF_MAX_MODALITIES];
  float inv_cnt[MSF_MAX_MODALITIES];
  long long B;
  int M, H;
};

__global__ void __launch_bounds__(256) tail_bwd_kernel(const __grid_constant__ TailBwdArgs a) {
  extern __shared__ float sh[];  // [M][H] gate-weight grads + [M] bias grads
  const int lane = threadIdx.x & 31;
  const int nwarp = blockDim.x >> 5;
  for (int i = threadIdx.x; i < a.M * a.H + a.M; i += blockDim.x) sh[i] = 0.0f;
  __syncthreads();
  for (long long row = blockIdx.x * (long long)nwarp + (threadIdx.x >> 5); row < a.B;
       row += (long long)gridDim.x * nwarp) {
    float mk[MSF_MAX_MODALITIES], p[MSF_MAX_MODALITIES], w[MSF_MAX_MODALITIES], dw[MSF_MAX_MODALITIES];
    float ds[MSF_MAX_MODALITIES];
    for (int q = 0; q < a.M; ++q) {
      mk[q] = a.mask ? __ldg(a.mask + row * a.M + q) : 1.0f;
      p[q] = __ldg(a.soft + row * a.M + q);
      w[q] = __ldg(a.w + row * a.M + q);
      const float* ag = a.agg + ((long long)q * a.B + row) * a.H;
      float part = 0.0f;
      for (int c = lane; c < a.H; c += 32) part = fmaf(__ldg(ag + c), __ldg(a.dfused + row * a.H + c), part);
      dw[q] = warp_sum(part);  // d fused / d w_q
    }
    // w = n / (S + 1e-8), n = p * mask (only when S > 0; the fallbacks are constants)
    float S = 0.0f;
    for (int q = 0; q < a.M; ++q) S += p[q] * mk[q];
    if (S > 0.0f) {
      const float inv = 1.0f / (S + 1e-8f);
      float dot = 0.0f;
      for (int q = 0; q < a.M; ++q) dot += dw[q] * p[q] * mk[q];
      float dp[MSF_MAX_MODALITIES], pdot = 0.0f;
      for (int q = 0; q < a.M; ++q) {
        dp[q] = (dw[q] * inv - dot * inv * inv) * mk[q];
        pdot += dp[q] * p[q];
      }
      for (int q = 0; q < a.M; ++q) ds[q] = (mk[q] > 0.0f) ? p[q] * (dp[q] - pdot) : 0.0f;
    } else {
      for (int q = 0; q < a.M; ++q) ds[q] = 0.0f;
    }
    for (int q = 0; q < a.M; ++q) {
      const float* ag = a.agg + ((long long)q * a.B + row) * a.H;
      float* out = a.dS + ((long long)q * a.B + row) * a.H;
      const float sc = mk[q] * a.inv_cnt[q];
      for (int c = lane; c < a.H; c += 32) {
        const float dagg = fmaf(w[q], __ldg(a.dfused + row * a.H + c), ds[q] * __ldg(a.gate_w[q] + c));
        out[c] = dagg * sc;
        if (ds[q] != 0.0f) atomicAdd(&sh[q * a.H + c], ds[q] * __ldg(ag + c));
      }
      if (lane == 0 && ds[q] != 0.0f) atomicAdd(&sh[a.M * a.H + q], ds[q]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.M * a.H; i += blockDim.x)
    if (sh[i] != 0.0f) atomicAdd(a.d_gate_w[i / a.H] + (i % a.H), sh[i]);
  for (int q = threadIdx.x; q < a.M; q += blockDim.x)
    if (sh[a.M * a.H + q] != 0.0f) atomicAdd(a.d_gate_b[q], sh[a.M * a.H + q]);
}

__global__ void copy_gates_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
static SimtProblem blank_problem() {
  SimtProblem p;
  memset(&p, 0, sizeof(p));
  p.nseg = 1;
  p.scale = 1.0f;
  p.head_dim = 1;
  p.heads = 1;
  return p;
}

int fusion_f32_forward(const Layout& L, const msf_fusion_call* c, cudaStream_t st) {
  const int64_t B = c->batch;
  const int M = L.M, H = L.H;
  WsF32 ws;
  carve_f32(L, B, c->workspace, &ws);
  if (c->workspace == nullptr || c->workspace_bytes < ws.bytes) {
    set_error("workspace too small: need %zu bytes, have %zu", ws.bytes, c->workspace_bytes);
    return MSF_E_WORKSPACE;
  }
  const float* W = c->params;
  const DropCfg drop = make_drop(c);
  const size_t BH = (size_t)B * H;
  int rc;

  {  // F0
    PrepArgs a;
    memset(&a, 0, sizeof(a));
    int maxd = 1;
    for (int m = 0; m < M; ++m) {
      a.x[m] = c->x[m];
      a.xt[m] = ws.xt[m];
      a.D[m] = L.D[m];
      if (L.D[m] > maxd) maxd = L.D[m];
    }
    a.M = M;
    a.B = B;
    a.mask = c->mask;
    a.drop = drop;
    const long long work = B * ((maxd + 3) / 4);
    dim3 grid((unsigned)(ceil_div(work, 256) < 2048 ? ceil_div(work, 256) : 2048), (unsigned)M);
    prep_kernel<<<grid, 256, 0, st>>>(a);
    MSF_LAUNCH_CHECK();
  }

  SimtProblem probs[MSF_MAX_MODALITIES * MSF_MAX_MODALITIES];
  int n = 0;
  for (int m = 0; m < M; ++m) {  // F1
    SimtProblem p = blank_problem();
    p.A[0] = ws.xt[m]; p.a_rs = L.D[m]; p.a_cs = 1;
    p.B[0] = W + L.proj_w[m]; p.b_rs = L.D[m]; p.b_cs = 1;
    p.bias[0] = W + L.proj_b[m];
    p.M = (int)B; p.N = H; p.K = L.D[m];
    p.C = ws.P[m]; p.ldc = H;
    p.epi = EPI_BIAS_RELU_DROP; p.site = SITE_PROJ; p.sub = m;
    probs[n++] = p;
  }
  if ((rc = simt_gemm_launch(probs, n, drop, st))) return rc;

  n = 0;
  for (int q = 0; q < M; ++q)  // F2
    for (int k = 0; k < M; ++k) {
      if (q == k || !L.has_pair(q, k)) continue;
      const int pi = L.pair_index(q, k);
      SimtProblem p = blank_problem();
      p.A[0] = ws.P[k]; p.a_rs = H; p.a_cs = 1;
      p.B[0] = W + L.pair_w(pi, 2); p.b_rs = H; p.b_cs = 1;
      p.bias[0] = W + L.pair_b(pi, 2);
      p.M = (int)B; p.N = H; p.K = H;
      p.C = ws.U + (size_t)pi * BH; p.ldc = H;
      p.epi = EPI_VALUE_GATE;
      p.mask = c->mask; p.mask_ld = M; p.mask_col = k;
      p.gate_out = ws.G + (size_t)pi * B * L.heads;
      p.head_dim = H / L.heads; p.heads = L.heads;
      p.sub = q * M + k;
      probs[n++] = p;
    }
  if ((rc = simt_gemm_launch(probs, n, drop, st))) return rc;

  n = 0;
  for (int q = 0; q < M; ++q) {  // F3
    SimtProblem p = blank_problem();
    int seg = 0;
    for (int k = 0; k < M; ++k) {
      if (q == k || !L.has_pair(q, k)) continue;
      const int pi = L.pair_index(q, k);
      p.A[seg] = ws.U + (size_t)pi * BH;
      p.B[seg] = W + L.pair_w(pi, 3);
      p.bias[seg] = W + L.pair_b(pi, 3);
      ++seg;
    }
    p.a_rs = H; p.a_cs = 1; p.b_rs = H; p.b_cs = 1;
    p.M = (int)B; p.N = H; p.K = seg ? H : 0;
    if (seg == 0) { p.A[0] = ws.P[q]; p.B[0] = W; seg = 1; }  // K = 0: epilogue only
    p.nseg = seg;
    p.C = ws.agg + (size_t)q * BH; p.ldc = H;
    p.epi = EPI_OUT_MEAN; p.scale = (float)L.mean_count(q);
    p.aux = ws.P[q]; p.ld_aux = H;
    p.mask = c->mask; p.mask_ld = M; p.mask_col = q;
    probs[n++] = p;
  }
  if ((rc = simt_gemm_launch(probs, n, drop, st))) return rc;

  {  // F4
    TailArgs a;
    memset(&a, 0, sizeof(a));
    a.agg = ws.agg;
    for (int m = 0; m < M; ++m) {
      a.gate_w[m] = W + L.gate_w[m];
      a.gate_b[m] = W + L.gate_b[m];
    }
    a.mask = c->mask; a.soft = ws.soft; a.w = ws.w; a.w_out = c->fusion_weights; a.fused = ws.fused;
    a.B = B; a.M = M; a.H = H;
    tail_fwd_kernel<<<(unsigned)ceil_div(B, 8), 256, 0, st>>>(a);
    MSF_LAUNCH_CHECK();
  }

  {  // F5
    SimtProblem p = blank_problem();
    p.A[0] = ws.fused; p.a_rs = H; p.a_cs = 1;
    p.B[0] = W + L.cls_w1; p.b_rs = H; p.b_cs = 1;
    p.bias[0] = W + L.cls_b1;
    p.M = (int)B; p.N = H; p.K = H;
    p.C = ws.Hr; p.ldc = H;
    p.epi = EPI_BIAS_RELU_DROP; p.site = SITE_CLS; p.sub = 0;
    if ((rc = simt_gemm_launch(&p, 1, drop, st))) return rc;
  }
  {  // F6
    SimtProblem p = blank_problem();
    p.A[0] = ws.Hr; p.a_rs = H; p.a_cs = 1;
    p.B[0] = W + L.cls_w2; p.b_rs = H; p.b_cs = 1;
    p.bias[0] = W + L.cls_b2;
    p.M = (int)B; p.N = L.C; p.K = H;
    p.C = c->logits; p.ldc = L.C;
    p.epi = EPI_STORE;
    if ((rc = simt_gemm_launch(&p, 1, drop, st))) return rc;
  }
  if (c->attn_gates) {
    const long long ng = (long long)L.num_pairs() * B * L.heads;
    // absent pairs have no map in the reference; their slots are zero-filled here
    for (int q = 0; q < M; ++q)
      for (int k = 0; k < M; ++k)
        if (q != k && !L.has_pair(q, k))
          MSF_CHECK_CUDA(cudaMemsetAsync(ws.G + (size_t)L.pair_index(q, k) * B * L.heads, 0,
                                         (size_t)B * L.heads * sizeof(float), st));
    if (ng > 0) {
      copy_gates_kernel<<<(unsigned)(ceil_div(ng, 256) < 1024 ? ceil_div(ng, 256) : 1024), 256, 0, st>>>(
          ws.G, c->attn_gates, ng);
      MSF_LAUNCH_CHECK();
    }
  }
  return MSF_OK;
}

int fusion_f32_backward(const Layout& L, const msf_fusion_call* c, cudaStream_t st) {
  const int64_t B = c->batch;
  const int M = L.M, H = L.H, C = L.C;
  WsF32 ws;
  carve_f32(L, B, c->workspace, &ws);
  if (c->workspace == nullptr || c->workspace_bytes < ws.bytes) {
    set_error("workspace too small: need %zu bytes, have %zu", ws.bytes, c->workspace_bytes);
    return MSF_E_WORKSPACE;
  }
  MSF_REQUIRE(c->grad_logits && c->grad_params, "backward needs grad_logits and grad_params");
  const float* W = c->params;
  float* dW = c->grad_params;
  const DropCfg drop = make_drop(c);
  const size_t BH = (size_t)B * H;
  int rc;

  // dead query/key projections and every slot not written below: exact zeros
  MSF_CHECK_CUDA(cudaMemsetAsync(dW, 0, (size_t)L.total * sizeof(float), st));

  SimtProblem probs[2 * MSF_MAX_MODALITIES * MSF_MAX_MODALITIES];
  ColsumProblem cs[2 * MSF_MAX_MODALITIES * MSF_MAX_MODALITIES];
  int n = 0, nc = 0;

  {  // B1: classifier.3 — dHr -> dH1 (through relu/dropout), dW2, db2
    SimtProblem p = blank_problem();
    p.A[0] = c->grad_logits; p.a_rs = C; p.a_cs = 1;
    p.B[0] = W + L.cls_w2; p.b_rs = 1; p.b_cs = H;      // B(n=h, k=c) = W2[c*H + h]
    p.M = (int)B; p.N = H; p.K = C;
    p.C = ws.dH1; p.ldc = H;
    p.epi = EPI_RELU_GRAD; p.scale = drop.scale; p.aux = ws.Hr; p.ld_aux = H;
    probs[n++] = p;
    SimtProblem g = blank_problem();                     // dW2[c,h] = sum_b dlogits[b,c] Hr[b,h]
    g.A[0] = c->grad_logits; g.a_rs = 1; g.a_cs = C;
    g.B[0] = ws.Hr; g.b_rs = 1; g.b_cs = H;
    g.M = C; g.N = H; g.K = (int)B;
    g.C = dW + L.cls_w2; g.ldc = H; g.epi = EPI_STORE;
    probs[n++] = g;
    if ((rc = simt_gemm_launch(probs, n, drop, st))) return rc;
    cs[nc++] = ColsumProblem{c->grad_logits, C, (int)B, C, dW + L.cls_b2};
  }
  {  // B2: classifier.0 — dfused, dW1, db1
    n = 0;
    SimtProblem p = blank_problem();
    p.A[0] = ws.dH1; p.a_rs = H; p.a_cs = 1;
    p.B[0] = W + L.cls_w1; p.b_rs = 1; p.b_cs = H;
    p.M = (int)B; p.N = H; p.K = H;
    p.C = ws.dfused; p.ldc = H; p.epi = EPI_STORE;
    probs[n++] = p;
    SimtProblem g = blank_problem();
    g.A[0] = ws.dH1; g.a_rs = 1; g.a_cs = H;
    g.B[0] = ws.fused; g.b_rs = 1; g.b_cs = H;
    g.M = H; g.N = H; g.K = (int)B;
    g.C = dW + L.cls_w1; g.ldc = H; g.epi = EPI_STORE;
    probs[n++] = g;
    if ((rc = simt_gemm_launch(probs, n, drop, st))) return rc;
    cs[nc++] = ColsumProblem{ws.dH1, H, (int)B, H, dW + L.cls_b1};
  }
  {  // B4: tail backward
    TailBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.agg = ws.agg; a.dfused = ws.dfused; a.mask = c->mask; a.soft = ws.soft; a.w = ws.w; a.dS = ws.dS;
    for (int m = 0; m < M; ++m) {
      a.gate_w[m] = W + L.gate_w[m];
      a.d_gate_w[m] = dW + L.gate_w[m];
      a.d_gate_b[m] = dW + L.gate_b[m];
      a.inv_cnt[m] = 1.0f / (float)L.mean_count(m);
    }
    a.B = B; a.M = M; a.H = H;
    const size_t sh = (size_t)(M * H + M) * sizeof(float);
    int grid = (int)(ceil_div(B, 8) < 592 ? ceil_div(B, 8) : 592);
    if (sh > 48 * 1024)
      MSF_CHECK_CUDA(cudaFuncSetAttribute(tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    tail_bwd_kernel<<<grid, 256, sh, st>>>(a);
    MSF_LAUNCH_CHECK();
  }
  {  // B5/B6: out_proj — dU_qk = dS_q Wo_qk (x gate -> dV_qk); dWo_qk = dS_q^T U_qk; dbo_qk = colsum dS_q
    n = 0;
    for (int q = 0; q < M; ++q)
      for (int k = 0; k < M; ++k) {
        if (q == k || !L.has_pair(q, k)) continue;
        const int pi = L.pair_index(q, k);
        SimtProblem p = blank_problem();
        p.A[0] = ws.dS + (size_t)q * BH; p.a_rs = H; p.a_cs = 1;
        p.B[0] = W + L.pair_w(pi, 3); p.b_rs = 1; p.b_cs = H;
        p.M = (int)B; p.N = H; p.K = H;
        p.C = ws.dV + (size_t)pi * BH; p.ldc = H;
        p.epi = EPI_GATE_MUL; p.gate_in = ws.G + (size_t)pi * B * L.heads;
        p.head_dim = H / L.heads; p.heads = L.heads;
        probs[n++] = p;
        SimtProblem g = blank_problem();
        g.A[0] = ws.dS + (size_t)q * BH; g.a_rs = 1; g.a_cs = H;
        g.B[0] = ws.U + (size_t)pi * BH; g.b_rs = 1; g.b_cs = H;
        g.M = H; g.N = H; g.K = (int)B;
        g.C = dW + L.pair_w(pi, 3); g.ldc = H; g.epi = EPI_STORE;
        probs[n++] = g;
        cs[nc++] = ColsumProblem{ws.dS + (size_t)q * BH, H, (int)B, H, dW + L.pair_b(pi, 3)};
      }
    if ((rc = simt_gemm_launch(probs, n, drop, st))) return rc;
  }
  {  // B7/B8: value_proj — dWv_qk = dV_qk^T P_k, dbv_qk; dP_k = dS_k + sum_q dV_qk Wv_qk -> dZ_k
    n = 0;
    for (int k = 0; k < M; ++k) {
      SimtProblem p = blank_problem();
      int seg = 0;
      for (int q = 0; q < M; ++q) {
        if (q == k || !L.has_pair(q, k)) continue;
        const int pi = L.pair_index(q, k);
        p.A[seg] = ws.dV + (size_t)pi * BH;
        p.B[seg] = W + L.pair_w(pi, 2);
        ++seg;
        SimtProblem g = blank_problem();
        g.A[0] = ws.dV + (size_t)pi * BH; g.a_rs = 1; g.a_cs = H;
        g.B[0] = ws.P[k]; g.b_rs = 1; g.b_cs = H;
        g.M = H; g.N = H; g.K = (int)B;
        g.C = dW + L.pair_w(pi, 2); g.ldc = H; g.epi = EPI_STORE;
        probs[n++] = g;
        cs[nc++] = ColsumProblem{ws.dV + (size_t)pi * BH, H, (int)B, H, dW + L.pair_b(pi, 2)};
      }
      p.a_rs = H; p.a_cs = 1; p.b_rs = 1; p.b_cs = H;
      p.M = (int)B; p.N = H; p.K = seg ? H : 0;
      if (seg == 0) { p.A[0] = ws.P[k]; p.B[0] = W; seg = 1; }
      p.nseg = seg;
      p.C = ws.dZ + (size_t)k * BH; p.ldc = H;
      p.epi = EPI_ADD_RELU_GRAD; p.scale = drop.scale;
      p.aux = ws.dS + (size_t)k * BH; p.ld_aux = H;
      p.aux2 = ws.P[k]; p.ld_aux2 = H;
      probs[n++] = p;
    }
    if ((rc = simt_gemm_launch(probs, n, drop, st))) return rc;
  }
  {  // B9: projections — dWp_m = dZ_m^T xt_m, dbp_m, dx_m = (dZ_m Wp_m) * mask_m * drop0
    n = 0;
    for (int m = 0; m < M; ++m) {
      SimtProblem g = blank_problem();
      g.A[0] = ws.dZ + (size_t)m * BH; g.a_rs = 1; g.a_cs = H;
      g.B[0] = ws.xt[m]; g.b_rs = 1; g.b_cs = L.D[m];
      g.M = H; g.N = L.D[m]; g.K = (int)B;
      g.C = dW + L.proj_w[m]; g.ldc = L.D[m]; g.epi = EPI_STORE;
      probs[n++] = g;
      cs[nc++] = ColsumProblem{ws.dZ + (size_t)m * BH, H, (int)B, H, dW + L.proj_b[m]};
      if (c->grad_x[m]) {
        SimtProblem p = blank_problem();
        p.A[0] = ws.dZ + (size_t)m * BH; p.a_rs = H; p.a_cs = 1;
        p.B[0] = W + L.proj_w[m]; p.b_rs = 1; p.b_cs = L.D[m];
        p.M = (int)B; p.N = L.D[m]; p.K = H;
        p.C = c->grad_x[m]; p.ldc = L.D[m];
        p.epi = EPI_DX; p.site = SITE_INPUT; p.sub = m;
        p.mask = c->mask; p.mask_ld = M; p.mask_col = m;
        probs[n++] = p;
      }
    }
    if ((rc = simt_gemm_launch(probs, n, drop, st))) return rc;
  }
  return colsum_launch(cs, nc, st);
}

}  // namespace msf

extern "C" int msf_adaptive_weights(const float* agg, const float* gate_w, const float* gate_b,
                                    const float* mask, int64_t batch, int32_t num_modalities, int32_t hidden,
                                    float* weights, void* stream) {
  MSF_REQUIRE(agg && gate_w && gate_b && weights && batch >= 0 && hidden >= 1, "msf_adaptive_weights: bad arguments");
  MSF_REQUIRE(num_modalities >= 1 && num_modalities <= MSF_MAX_MODALITIES, "msf_adaptive_weights: bad M");
  if (batch == 0) return MSF_OK;
  msf::TailArgs a;
  memset(&a, 0, sizeof(a));
  a.agg = agg;
  for (int m = 0; m < num_modalities; ++m) {
    a.gate_w[m] = gate_w + (size_t)m * hidden;
    a.gate_b[m] = gate_b + m;
  }
  a.mask = mask; a.w_out = weights;
  a.B = batch; a.M = num_modalities; a.H = hidden;
  msf::tail_fwd_kernel<<<(unsigned)msf::ceil_div(batch, 8), 256, 0, (cudaStream_t)stream>>>(a);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}
