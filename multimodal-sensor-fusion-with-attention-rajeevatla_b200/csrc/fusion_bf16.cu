// HybridFusion forward / backward on the tensor-core path (MSF_PREC_BF16).
//
// Same stage structure as the fp32 path (fusion_f32.cu; src/fusion.py:331-479,
// src/attention.py:68-146) with bf16 activations / weights, fp32 accumulation in
// TMEM (tcgen05.mma) and every elementwise step fused into a GEMM epilogue:
//
//   F0 prep      xt_m  = bf16(drop0(x_m * mask_m))                          [elementwise]
//   F1 proj      P_m   = drop1(relu(xt_m Wp_m^T + bp_m))                    tc_gemm, 1 launch, M problems
//   F2 value     U_qk  = g_qk * (P_k Wv_qk^T + bv_qk)      (records g)      tc_gemm, 1 launch, M(M-1) problems
//   F3 out+mean  agg_q = (P_q + sum_k (U_qk Wo_qk^T + bo_qk)) / cnt_q * mask_q   tc_gemm, K-segments summed in TMEM
//   F4 tail      gating dot, masked softmax + fallbacks, weighted sum        [warp per window]
//   F5 cls1      Hr = drop3(relu(fused W1^T + b1));  F6 logits = Hr W2^T + b2
//
// Backward: dgrad GEMMs are K-major launches against transposed bf16 weight
// copies kept in the compute arena; all weight gradients (dY^T . X, contraction
// over the windows) go into ONE MN-major launch at the end; bias gradients are
// column sums.  Dead query/key projection slots are exact zeros (memset).
#include <stdlib.h>

#include "chain_gemm.cuh"
#include "fusion_common.cuh"
#include "fusion_bf16_layout.cuh"
#include "head_gemm.cuh"
#include "proj_gemm.cuh"
#include "tc_gemm.cuh"
#include "wg2_gemm.cuh"

namespace msf {


struct WsBf16 {
  bf16* xt[MSF_MAX_MODALITIES];  // [B][D_m]
  bf16* P;       // [M][B][H]
  bf16* U;       // [pairs][B][H]
  float* G;      // [pairs][B][heads]
  bf16* agg;     // [M][B][H]
  float* soft;   // [B][M]
  float* w;      // [B][M]
  bf16* fused;   // [B][H]
  bf16* Hr;      // [B][H]
  // backward
  bf16* dlog;    // [B][Cp]
  bf16* dH1;     // [B][H]
  bf16* dfused;  // [B][H]
  float* ds;     // [B][M]  d loss / d gating score
  bf16* dS;      // [M][B][H]
  bf16* dV;      // [pairs][B][H]
  bf16* dZ;      // [M][B][H]
  size_t bytes;
};

static void carve_bf16(const Layout& L, int64_t B, void* base, WsBf16* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? reinterpret_cast<char*>(base) + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const size_t BH = (size_t)B * L.H;
  const size_t pairs = L.num_pairs() > 0 ? L.num_pairs() : 1;
  const int Cp = (int)align_up(L.C, 8);
  for (int m = 0; m < L.M; ++m) ws->xt[m] = (bf16*)take((size_t)B * L.D[m] * 2);
  ws->P = (bf16*)take(BH * L.M * 2);
  ws->U = (bf16*)take(BH * pairs * 2);
  ws->G = (float*)take((size_t)B * L.heads * pairs * 4);
  ws->agg = (bf16*)take(BH * L.M * 2);
  ws->soft = (float*)take((size_t)B * L.M * 4);
  ws->w = (float*)take((size_t)B * L.M * 4);
  ws->fused = (bf16*)take(BH * 2);
  ws->Hr = (bf16*)take(BH * 2);
  ws->dlog = (bf16*)take((size_t)B * Cp * 2);
  ws->dH1 = (bf16*)take(BH * 2);
  ws->dfused = (bf16*)take(BH * 2);
  ws->ds = (float*)take((size_t)B * L.M * 4);
  ws->dS = (bf16*)take(BH * L.M * 2);
  ws->dV = (bf16*)take(BH * pairs * 2);
  ws->dZ = (bf16*)take(BH * L.M * 2);
  ws->bytes = off;
}

bool fusion_bf16_eligible(const Layout& L) {
  if (L.H % 64 != 0) return false;
  for (int m = 0; m < L.M; ++m)
    if (L.D[m] % 8 != 0) return false;
  return true;
}
size_t fusion_bf16_workspace_bytes(const Layout& L, int64_t B) {
  WsBf16 ws;
  carve_bf16(L, B, nullptr, &ws);
  return ws.bytes;
}
size_t fusion_bf16_arena_bytes(const Layout& L) { return arena_layout(L).total * sizeof(bf16); }

static int block_n_for(int n) { return n >= 256 ? 256 : (int)align_up(n, 64); }

// Tile width for a launch of `problems` GEMMs with `rows` output rows and n columns each: wide tiles
// (fewer operand re-reads) when that already fills the 148 SMs, otherwise narrower ones so that a
// small launch spreads over more SMs and each CTA's epilogue is shorter.
static int block_n_fill(int n, long long rows, int problems) {
  int bn = block_n_for(n);
  const long long m_tiles = ceil_div(rows, TC_BLOCK_M) * problems;
  while (bn > 64 && m_tiles * ceil_div(n, bn) < 120 && (bn / 2) % 64 == 0) bn /= 2;
  return bn;
}

// ---------------------------------------------------------------------------
// pack: fp32 master -> bf16 compute arena (plain and transposed copies)
// ---------------------------------------------------------------------------
struct PackJob {
  long long src, src_batch;   // element offsets in the master arena
  long long dst, dst_batch;   // element offsets in the compute arena
  int rows, cols;             // source matrix (rows x cols, row-major)
  int dst_ld;                 // destination row pitch
  int batch;
  int transpose;              // destination holds cols x rows
  int tile_begin;
};
constexpr int PACK_MAX_JOBS = 4 * MSF_MAX_MODALITIES + 8;
struct PackList {
  PackJob j[PACK_MAX_JOBS];
  int count, total_tiles;
};

// one 32x32 tile per block iteration; transposes go through shared memory so both sides coalesce
__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackList list,
                                                   const float* __restrict__ master, bf16* __restrict__ arena) {
  __shared__ float tile[32][33];
  for (int t = blockIdx.x; t < list.total_tiles; t += gridDim.x) {
    int ji = 0;
    while (ji + 1 < list.count && t >= list.j[ji + 1].tile_begin) ++ji;
    const PackJob& J = list.j[ji];
    const int tr = (J.rows + 31) >> 5, tc = (J.cols + 31) >> 5;
    int local = t - J.tile_begin;
    const int b = local / (tr * tc);
    local -= b * tr * tc;
    const int r0 = (local / tc) << 5, c0 = (local % tc) << 5;
    const float* src = master + J.src + (long long)b * J.src_batch;
    bf16* dst = arena + J.dst + (long long)b * J.dst_batch;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (!J.transpose) {
      for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        if (r < J.rows && c < J.cols) dst[(long long)r * J.dst_ld + c] = __float2bfloat16_rn(__ldg(src + (long long)r * J.cols + c));
      }
    } else {
      for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < J.rows && c < J.cols) ? __ldg(src + (long long)r * J.cols + c) : 0.0f;
      }
      __syncthreads();
      for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;  // destination row = source column
        if (c < J.cols && r < J.rows) dst[(long long)c * J.dst_ld + r] = __float2bfloat16_rn(tile[tx][i]);
      }
      __syncthreads();
    }
  }
}

int fusion_bf16_pack(const Layout& L, const float* params, void* arena_v, cudaStream_t st) {
  const ArenaBf16 A = arena_layout(L);
  bf16* arena = reinterpret_cast<bf16*>(arena_v);
  PackList list;
  memset(&list, 0, sizeof(list));
  const long long H = L.H;
  auto add = [&](long long src, long long sb, long long dst, long long db, int rows, int cols, int dst_ld,
                 int batch, int transpose) {
    if (batch <= 0) return;
    PackJob& J = list.j[list.count++];
    J.src = src; J.src_batch = sb; J.dst = dst; J.dst_batch = db;
    J.rows = rows; J.cols = cols; J.dst_ld = dst_ld; J.batch = batch; J.transpose = transpose;
    J.tile_begin = list.total_tiles;
    list.total_tiles += batch * ((rows + 31) / 32) * ((cols + 31) / 32);
  };
  for (int m = 0; m < L.M; ++m) {
    add(L.proj_w[m], 0, (long long)A.wp[m], 0, L.H, L.D[m], L.D[m], 1, 0);
    add(L.proj_w[m], 0, (long long)A.wpT[m], 0, L.H, L.D[m], L.H, 1, 1);
  }
  const int pairs = L.num_pairs();
  if (pairs > 0) {
    add(L.pair_w(0, 2), L.pair_stride, (long long)A.wv, H * H, L.H, L.H, L.H, pairs, 0);
    add(L.pair_w(0, 3), L.pair_stride, (long long)A.wo, H * H, L.H, L.H, L.H, pairs, 0);
    add(L.pair_w(0, 2), L.pair_stride, (long long)A.wvT, H * H, L.H, L.H, L.H, pairs, 1);
    add(L.pair_w(0, 3), L.pair_stride, (long long)A.woT, H * H, L.H, L.H, L.H, pairs, 1);
  }
  add(L.cls_w1, 0, (long long)A.w1, 0, L.H, L.H, L.H, 1, 0);
  add(L.cls_w1, 0, (long long)A.w1T, 0, L.H, L.H, L.H, 1, 1);
  add(L.cls_w2, 0, (long long)A.w2, 0, L.C, L.H, L.H, 1, 0);
  // w2T is (H x Cp) with zero padding columns C..Cp-1: clear it first
  MSF_CHECK_CUDA(cudaMemsetAsync(arena + A.w2T, 0, (size_t)L.H * A.Cp * sizeof(bf16), st));
  add(L.cls_w2, 0, (long long)A.w2T, 0, L.C, L.H, A.Cp, 1, 1);
  const int grid = list.total_tiles < 1184 ? list.total_tiles : 1184;
  pack_kernel<<<grid, 256, 0, st>>>(list, params, arena);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

// ---------------------------------------------------------------------------
// small vector helpers: 8 consecutive bf16 <-> 8 floats
// ---------------------------------------------------------------------------
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __bfloat1622float2(h[e]);
    v[2 * e] = f.x;
    v[2 * e + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 pk;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = pk;
}
// 8 fp32 values from the master arena (tensors there are only 4-byte aligned)
__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __ldg(p + j);
}

// ---------------------------------------------------------------------------
// F0: xt = bf16(drop0(x * mask))           (fusion.py:370-374)
// ---------------------------------------------------------------------------
struct PrepArgs16 {
  const float* x[MSF_MAX_MODALITIES];
  bf16* xt[MSF_MAX_MODALITIES];
  int D[MSF_MAX_MODALITIES];
  int M;
  long long B;
  const float* mask;
  DropCfg drop;
};

// Gradient slots that the backward pass ACCUMULATES into (bias / gating gradients via atomics), the dead
// query/key slots and the slots of deleted pair modules: cleared by the first kernel of the train pass
// instead of a 13 MB memset node over the whole arena (the GEMM weight gradients are plain stores).
struct ZeroRange {
  long long begin, count, stride;
  int batch;
};
constexpr int PREP_MAX_ZERO = 24;
struct ZeroList {
  ZeroRange r[PREP_MAX_ZERO];
  int n;
  float* base;
};

__device__ __forceinline__ void zero_ranges(const ZeroList& z, long long tid, long long nthreads) {
  for (int i = 0; i < z.n; ++i) {
    const ZeroRange R = z.r[i];
    for (int b = 0; b < R.batch; ++b) {
      float* p = z.base + R.begin + (long long)b * R.stride;
      if (((R.begin + (long long)b * R.stride) & 3) == 0) {
        const long long n4 = R.count >> 2;
        for (long long e = tid; e < n4; e += nthreads) reinterpret_cast<float4*>(p)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long e = (n4 << 2) + tid; e < R.count; e += nthreads) p[e] = 0.0f;
      } else {
        for (long long e = tid; e < R.count; e += nthreads) p[e] = 0.0f;
      }
    }
  }
}

__global__ void __launch_bounds__(256) prep16_kernel(const __grid_constant__ PrepArgs16 a,
                                                     const __grid_constant__ ZeroList z) {
  pdl_wait();
  pdl_launch();
  // a slice of the grid does the clearing (the ranges are small; every thread walking the list costs more)
  constexpr int kZeroCtas = 64;
  if (z.n > 0 && blockIdx.y == 0 && blockIdx.x < kZeroCtas) {
    const int ctas = (int)gridDim.x < kZeroCtas ? (int)gridDim.x : kZeroCtas;
    zero_ranges(z, blockIdx.x * (long long)blockDim.x + threadIdx.x, (long long)ctas * blockDim.x);
  }
  const int m = blockIdx.y;
  const int D = a.D[m];
  const DropCfg drop = resolve_drop(a.drop);
  const int quads = D >> 2;  // D % 8 == 0
  const long long total = a.B * quads;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / quads;
    const int c4 = (int)(i % quads);
    const float mk = a.mask ? __ldg(a.mask + row * a.M + m) : 1.0f;
    float dm[4] = {1.f, 1.f, 1.f, 1.f};
    if (drop.active) drop4(drop, SITE_INPUT, m, row, c4, dm);
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.x[m] + row * D) + c4);
    uint2 pk;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
    h[0] = __floats2bfloat162_rn(v.x * mk * dm[0], v.y * mk * dm[1]);
    h[1] = __floats2bfloat162_rn(v.z * mk * dm[2], v.w * mk * dm[3]);
    *reinterpret_cast<uint2*>(a.xt[m] + row * D + c4 * 4) = pk;
  }
}

// dlogits (B, C) fp32 -> (B, Cp) bf16, zero padded
__global__ void __launch_bounds__(256) cvt_pad_kernel(const float* __restrict__ src, bf16* __restrict__ dst,
                                                      long long rows, int cols, int ld) {
  const long long total = rows * ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld;
    const int c = (int)(i % ld);
    dst[i] = __float2bfloat16_rn(c < cols ? __ldg(src + r * cols + c) : 0.0f);
  }
}

// ---------------------------------------------------------------------------
// F4: gating, masked softmax with fallbacks, weighted sum   (fusion.py:410-418, 429-479)
// one warp per window; each lane owns 8-column chunks
// ---------------------------------------------------------------------------
struct Tail16Args {
  const bf16* agg;     // [M][B][H]
  const float* gate_w[MSF_MAX_MODALITIES];
  const float* gate_b[MSF_MAX_MODALITIES];
  const float* mask;
  float* soft;
  float* w;
  float* w_out;
  bf16* fused;         // (B, H)
  long long B;
  int M, H;
};

__global__ void __launch_bounds__(256) tail16_fwd_kernel(const __grid_constant__ Tail16Args a) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.B) return;
  float s[MSF_MAX_MODALITIES], mk[MSF_MAX_MODALITIES], soft[MSF_MAX_MODALITIES], w[MSF_MAX_MODALITIES];
  for (int q = 0; q < a.M; ++q) {
    const bf16* ag = a.agg + ((long long)q * a.B + row) * a.H;
    float part = 0.0f;
    for (int c = lane * 8; c < a.H; c += 256) {
      float v[8], g[8];
      load8(ag + c, v);
      load8f(a.gate_w[q] + c, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) part = fmaf(v[j], g[j], part);
    }
    s[q] = warp_sum(part) + __ldg(a.gate_b[q]);  // fusion.py:459
    mk[q] = a.mask ? __ldg(a.mask + row * a.M + q) : 1.0f;
  }
  adaptive_weights_row(s, mk, a.M, soft, w);
  if (lane == 0)
    for (int q = 0; q < a.M; ++q) {
      a.soft[row * a.M + q] = soft[q];
      a.w[row * a.M + q] = w[q];
      if (a.w_out) a.w_out[row * a.M + q] = w[q];
    }
  for (int c = lane * 8; c < a.H; c += 256) {
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < a.M; ++q) {
      float v[8];
      load8(a.agg + ((long long)q * a.B + row) * a.H + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(v[j], w[q], f[j]);
    }
    store8(a.fused + row * a.H + c, f);  // fusion.py:416-418
  }
}

// ---------------------------------------------------------------------------
// B4: backward of F4.  dfused -> dS_q (already scaled by mask_q / cnt_q) and the
// per-window score gradients ds (B, M); gating-layer gradients are weighted
// column sums of agg with ds, done by wcolsum16_kernel.
// ---------------------------------------------------------------------------
struct Tail16BwdArgs {
  const bf16* agg;
  const bf16* dfused;
  const float* gate_w[MSF_MAX_MODALITIES];
  const float* mask;
  const float* soft;
  const float* w;
  bf16* dS;       // [M][B][H]
  float* ds;      // (B, M)
  float inv_cnt[MSF_MAX_MODALITIES];
  long long B;
  int M, H;
};

__global__ void __launch_bounds__(256) tail16_bwd_kernel(const __grid_constant__ Tail16BwdArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.B) return;
  float mk[MSF_MAX_MODALITIES], p[MSF_MAX_MODALITIES], w[MSF_MAX_MODALITIES], dw[MSF_MAX_MODALITIES];
  float ds[MSF_MAX_MODALITIES];
  for (int q = 0; q < a.M; ++q) {
    mk[q] = a.mask ? __ldg(a.mask + row * a.M + q) : 1.0f;
    p[q] = __ldg(a.soft + row * a.M + q);
    w[q] = __ldg(a.w + row * a.M + q);
    const bf16* ag = a.agg + ((long long)q * a.B + row) * a.H;
    float part = 0.0f;
    for (int c = lane * 8; c < a.H; c += 256) {
      float v[8], g[8];
      load8(ag + c, v);
      load8(a.dfused + row * a.H + c, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) part = fmaf(v[j], g[j], part);
    }
    dw[q] = warp_sum(part);  // d loss / d w_q
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
  if (lane == 0)
    for (int q = 0; q < a.M; ++q) a.ds[row * a.M + q] = ds[q];
  for (int q = 0; q < a.M; ++q) {
    bf16* out = a.dS + ((long long)q * a.B + row) * a.H;
    const float sc = mk[q] * a.inv_cnt[q];
    for (int c = lane * 8; c < a.H; c += 256) {
      float g[8], gw[8], o[8];
      load8(a.dfused + row * a.H + c, g);
      load8f(a.gate_w[q] + c, gw);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(w[q], g[j], ds[q] * gw[j]) * sc;
      store8(out + c, o);
    }
  }
}

// ---------------------------------------------------------------------------
// (weighted) column sums of bf16 / fp32 matrices, accumulated with fp32 atomics
// into pre-zeroed destinations: bias gradients and gating-layer gradients.
//   dst[c] += sum_r coef[r*coef_ld] * src[r*ld + c]        (coef == nullptr: 1)
// grid = (row chunks, problems); a warp covers 256 columns of one row per step.
// ---------------------------------------------------------------------------
struct Colsum16Problem {
  const void* src;
  int src_f32;         // 1: fp32 source, 0: bf16
  long long ld;
  int rows, cols;
  const float* coef;
  int coef_ld;
  float* dst;
  float* dst_more[MSF_MAX_MODALITIES];   // further destinations of the same sums (nullptr-terminated)
};
constexpr int COLSUM16_MAX = 48;
constexpr int CS_ROWS = 8;   // rows per warp in flight
// Dynamic shared memory the column-sum blocks ask for without using it: with it a block no longer fits beside a
// resident tensor-core CTA (those leave less than one pipeline stage of shared memory free), so the column sums run
// on the SMs the tensor-core grid leaves idle instead of taking issue slots and shared-memory bandwidth from it.
constexpr int CS_PAD_BYTES = 40960;
struct Colsum16List {
  Colsum16Problem p[COLSUM16_MAX];
  int count;
  int rows_per_block;
  int tag;   // which launch of the step (timeline builds)
};

__global__ void __launch_bounds__(256) colsum16_kernel(const __grid_constant__ Colsum16List list) {
  TL_KERNEL(list.tag);
  const Colsum16Problem& P = list.p[blockIdx.y];
  const int r0 = blockIdx.x * list.rows_per_block;
  if (r0 >= P.rows) return;
  const int r1 = min(P.rows, r0 + list.rows_per_block);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float red[8][256];
  for (int cbase = 0; cbase < P.cols; cbase += 256) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int c = cbase + lane * 8;
    const bool vec = (c + 8 <= P.cols) && ((P.ld & 7) == 0);
    if (c < P.cols && vec && !P.src_f32) {
      // bf16 rows, 16 bytes per lane: CS_ROWS rows per warp in flight (the block may be one of only a few on its SM)
      const bf16* s0 = reinterpret_cast<const bf16*>(P.src) + c;
      for (int rb = r0 + warp; rb < r1; rb += 8 * CS_ROWS) {
        float kk[CS_ROWS];
        uint4 raw[CS_ROWS];
#pragma unroll
        for (int u = 0; u < CS_ROWS; ++u) {
          const int r = rb + 8 * u;
          kk[u] = r < r1 ? (P.coef ? __ldg(P.coef + (long long)r * P.coef_ld) : 1.0f) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < CS_ROWS; ++u) {
          raw[u] = make_uint4(0u, 0u, 0u, 0u);
          if (kk[u] != 0.0f) raw[u] = __ldg(reinterpret_cast<const uint4*>(s0 + (long long)(rb + 8 * u) * P.ld));
        }
#pragma unroll
        for (int u = 0; u < CS_ROWS; ++u) {
          if (kk[u] == 0.0f) continue;   // zero-weight rows are never read (they may hold anything)
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            acc[2 * e] = fmaf(kk[u], f.x, acc[2 * e]);
            acc[2 * e + 1] = fmaf(kk[u], f.y, acc[2 * e + 1]);
          }
        }
      }
    } else if (c < P.cols) {
      for (int r = r0 + warp; r < r1; r += 8) {
        const float k = P.coef ? __ldg(P.coef + (long long)r * P.coef_ld) : 1.0f;
        if (k == 0.0f) continue;
        float v[8];
        if (P.src_f32) {
          const float* s = reinterpret_cast<const float*>(P.src) + (long long)r * P.ld + c;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (c + j < P.cols) ? __ldg(s + j) : 0.0f;
        } else {
          const bf16* s = reinterpret_cast<const bf16*>(P.src) + (long long)r * P.ld + c;
          if (vec) {
            load8(s, v);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (c + j < P.cols) ? __bfloat162float(s[j]) : 0.0f;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(k, v[j], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int col = cbase + threadIdx.x;
    if (col < P.cols) {
      float t = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
      if (t != 0.0f) {
        atomicAdd(P.dst + col, t);
        for (int d = 0; d < MSF_MAX_MODALITIES && P.dst_more[d] != nullptr; ++d) atomicAdd(P.dst_more[d] + col, t);
      }
    }
    __syncthreads();
  }
}

// sum of squares of the bias / gating-layer gradient slots -> msf_fusion_call.grad_sq (after the column sums
// that fill them; the weight matrices' share comes from the weight-gradient GEMM's epilogue)
struct VecSqList {
  long long begin[2 * MSF_MAX_MODALITIES * MSF_MAX_MODALITIES + MSF_MAX_MODALITIES + 4];
  int count[2 * MSF_MAX_MODALITIES * MSF_MAX_MODALITIES + MSF_MAX_MODALITIES + 4];
  int n;
};
__global__ void __launch_bounds__(256) vec_sq_kernel(const __grid_constant__ VecSqList list, const float* __restrict__ g,
                                                     double* __restrict__ sq_out) {
  __shared__ double red[8];
  const float* p = g + list.begin[blockIdx.x];
  double sq = 0.0;
  for (int e = threadIdx.x; e < list.count[blockIdx.x]; e += 256) {
    const double x = (double)__ldcg(p + e);
    sq += x * x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    if (t != 0.0) atomicAdd(sq_out, t);
  }
}

// mean of the per-window losses in a fixed order (deterministic): one block beside the backward chain instead of a
// serial tail on the head kernel's last CTA
__global__ void __launch_bounds__(256) loss_mean_kernel(const float* __restrict__ row_loss, long long rows,
                                                        float* __restrict__ loss_out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (long long i = threadIdx.x; i < rows; i += 256) s += (double)__ldcg(row_loss + i);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = (float)(sh[0] / (double)rows);
}

static int colsum16_launch(const Colsum16Problem* probs, int count, cudaStream_t st, int tag = 0) {
  int done = 0;
  while (done < count) {
    Colsum16List list;
    const int n = (count - done) < COLSUM16_MAX ? (count - done) : COLSUM16_MAX;
    int max_rows = 0;
    for (int i = 0; i < n; ++i) {
      list.p[i] = probs[done + i];
      if (list.p[i].rows > max_rows) max_rows = list.p[i].rows;
    }
    list.count = n;
    list.tag = tag;
    // ~4 waves of blocks over 148 SMs x 8 resident blocks
    int chunks = (int)ceil_div(148 * 8, n);
    if (chunks < 1) chunks = 1;
    int rpb = (int)ceil_div(max_rows, chunks);
    if (rpb < 64) rpb = 64;
    list.rows_per_block = rpb;
    if (max_rows > 0) {
      // Highest launch priority: the block scheduler hands out grids of one priority in launch order, and the
      // tensor-core kernel launched ahead of this one (programmatic dependent launch) has blocks waiting for SMs its
      // predecessor still holds -- without the priority these small blocks queue behind them instead of running
      // on the idle SMs.  Set per launch so it also holds for the kernel node of a captured graph.
      int least = 0, greatest = 0;
      MSF_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributePriority;
      attr.val.priority = greatest;
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)ceil_div(max_rows, rpb), (unsigned)n);
      cfg.blockDim = dim3(256);
      // tag 0 runs beside the backward chain kernel (shared-memory bound: keep these blocks off its SMs with a
      // shared-memory request that does not fit beside a chain CTA); tag 1 runs beside the weight-gradient CTA
      // pairs, which wait on the L2 and leave their SMs' issue slots and shared-memory port idle: no padding, the
      // blocks spread over all SMs (MSF_CS_PAD=1 pads both)
      static const bool pad_all = getenv("MSF_CS_PAD") != nullptr;
      cfg.dynamicSmemBytes = (tag == 0 || pad_all) ? CS_PAD_BYTES : 0;
      cfg.stream = st;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      MSF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, colsum16_kernel, list));
      MSF_LAUNCH_CHECK();
    }
    done += n;
  }
  return MSF_OK;
}

__global__ void copy_gates16_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
// A second stream per device so independent kernels of one step overlap (the weight-gradient GEMM
// occupies ~110 SMs; the bias-gradient column sums run beside it).  Forked and joined with events,
// which also works while the caller's stream is being captured into a CUDA graph.
struct SideStream {
  cudaStream_t stream;
  cudaEvent_t fork, fork2, join;
};
static int side_stream(SideStream** out) {
  static SideStream cache[16];
  static bool made[16] = {false};
  int dev = 0;
  MSF_CHECK_CUDA(cudaGetDevice(&dev));
  MSF_REQUIRE(dev >= 0 && dev < 16, "device index %d out of range", dev);
  if (!made[dev]) {
    int least = 0, greatest = 0;   // see colsum16_launch: the side work must not queue behind pending tensor-core blocks
    MSF_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    MSF_CHECK_CUDA(cudaStreamCreateWithPriority(&cache[dev].stream, cudaStreamNonBlocking, greatest));
    MSF_CHECK_CUDA(cudaEventCreateWithFlags(&cache[dev].fork, cudaEventDisableTiming));
    MSF_CHECK_CUDA(cudaEventCreateWithFlags(&cache[dev].fork2, cudaEventDisableTiming));
    MSF_CHECK_CUDA(cudaEventCreateWithFlags(&cache[dev].join, cudaEventDisableTiming));
    made[dev] = true;
  }
  *out = &cache[dev];
  return MSF_OK;
}

static int check_ws(const WsBf16& ws, const msf_fusion_call* c) {
  if (c->workspace == nullptr || c->workspace_bytes < ws.bytes) {
    set_error("workspace too small: need %zu bytes, have %zu", ws.bytes, c->workspace_bytes);
    return MSF_E_WORKSPACE;
  }
  if (c->batch >= (1 << 24)) {
    set_error("batch %d too large for one call of the tensor-core path", c->batch);
    return MSF_E_INVALID;
  }
  return MSF_OK;
}

// F0..F3: inputs -> aggregated modality tokens (ws.agg), gates in ws.G
// present_hint: 0 = the mask varies per row; otherwise bit m says whether modality m is present in EVERY row
// (absent in every row if clear) — the subset sweep of src/eval.py:342-404.  Work for absent modalities is
// skipped: their projections, the chain items of absent queries (aggregated = 0 through the row mask) and the
// GEMMs of pairs with an absent key (gate 0: they contribute exactly their out_proj bias).
static int forward_front(const Layout& L, const msf_fusion_call* c, const WsBf16& ws, const ArenaBf16& A,
                         cudaStream_t st, const ZeroList* zero = nullptr, unsigned present_hint = 0u,
                         bool projections_only = false) {
  const int64_t B = c->batch;
  const int M = L.M, H = L.H;
  int rc = MSF_OK;
  const bf16* W16 = reinterpret_cast<const bf16*>(c->params_bf16);
  const float* W = c->params;
  const DropCfg drop = make_drop(c);
  const long long BH = (long long)B * H;
  const int pairs = L.num_pairs();
  const int bnH = block_n_for(H);

  if (proj_eligible(H, M, L.D) && !getenv("MSF_NO_PROJ")) {  // F0 + F1 in one kernel (proj_gemm.cu)
    ProjLaunch pl;
    memset(&pl, 0, sizeof(pl));
    pl.M = M; pl.H = H; pl.rows = (int)B;
    for (int m = 0; m < M; ++m) {
      MSF_REQUIRE((reinterpret_cast<uintptr_t>(c->x[m]) & 15) == 0, "features of modality %d are not 16-byte aligned", m);
      pl.D[m] = L.D[m];
      pl.x[m] = c->x[m];
      pl.bias[m] = W + L.proj_b[m];
      pl.ln_w[m] = c->ln_weight[m];
      pl.ln_b[m] = c->ln_bias[m];
      if ((rc = tc_encode_map(&pl.map_w[m], W16 + A.wp[m], H, L.D[m], L.D[m], 1, 0, 64, H))) return rc;
      if ((rc = tc_encode_map(&pl.map_xt[m], ws.xt[m], B, L.D[m], L.D[m], 1, 0, 64, 128))) return rc;
    }
    if ((rc = tc_encode_map(&pl.map_p, ws.P, B, H, H, M, BH, 64, 128))) return rc;
    pl.mask = c->mask;
    pl.x_bf16 = c->x_bf16 != 0;
    pl.ln_eps = c->ln_eps;
    pl.drop = drop;
    if (present_hint != 0u)
      for (int m = 0; m < M; ++m)
        if ((present_hint >> m) & 1u) pl.active[pl.n_active++] = (short)m;
    if (zero != nullptr && zero->n > 0) {
      MSF_REQUIRE(zero->n <= PROJ_MAX_ZERO, "too many gradient ranges to clear");
      for (int i = 0; i < zero->n; ++i) {
        pl.zero[i].begin = zero->r[i].begin; pl.zero[i].count = zero->r[i].count;
        pl.zero[i].stride = zero->r[i].stride; pl.zero[i].batch = zero->r[i].batch;
      }
      pl.nzero = zero->n;
      pl.zero_base = zero->base;
    }
    if ((rc = proj_launch(pl, st, "F0+F1 input prep + projections"))) return rc;
  } else {
  if (c->x_bf16) {
    set_error("msf_fusion_call.x_bf16 is only read by the fused projection kernel (in_dims %% 64 == 0, <= 256, hidden <= 256)");
    return MSF_E_UNSUPPORTED;
  }
  for (int m = 0; m < M; ++m)
    MSF_REQUIRE(c->ln_weight[m] == nullptr && c->ln_bias[m] == nullptr,
                "msf_fusion_call.ln_* is only applied by the fused projection kernel (msf_fusion_layer_norm_fused() == 0 "
                "for this shape): run msf_layer_norm_forward first");
  {  // F0
    PrepArgs16 a;
    memset(&a, 0, sizeof(a));
    int maxd = 8;
    for (int m = 0; m < M; ++m) {
      a.x[m] = c->x[m];
      a.xt[m] = ws.xt[m];
      a.D[m] = L.D[m];
      if (L.D[m] > maxd) maxd = L.D[m];
      MSF_REQUIRE((reinterpret_cast<uintptr_t>(c->x[m]) & 15) == 0, "features of modality %d are not 16-byte aligned", m);
    }
    a.M = M;
    a.B = B;
    a.mask = c->mask;
    a.drop = drop;
    const long long work = B * (maxd / 4);
    dim3 grid((unsigned)(ceil_div(work, 256) < 1184 ? ceil_div(work, 256) : 1184), (unsigned)M);
    ZeroList zl;
    memset(&zl, 0, sizeof(zl));
    if (zero != nullptr) zl = *zero;
    MSF_CHECK_CUDA(launch_pdl(prep16_kernel, grid, dim3(256), 0, st, a, zl));
    MSF_LAUNCH_CHECK();
  }

  {  // F1: P_m = drop1(relu(xt_m Wp_m^T + bp_m))
    TcBuilder tb(false, bnH, drop, st, "F1 projections");
    for (int m = 0; m < M; ++m) {
      TcProblem p = tc_blank_problem();
      p.seg[0].a_map = (short)tb.add_map(ws.xt[m], B, L.D[m], L.D[m], 1, 0, TC_BLOCK_M);
      p.seg[0].b_map = (short)tb.add_map(W16 + A.wp[m], H, L.D[m], L.D[m], 1, 0, bnH);
      p.bias[0] = W + L.proj_b[m];
      p.M = (int)B; p.N = H; p.K = L.D[m];
      p.C = ws.P + (long long)m * BH; p.ldc = H; p.c_bf16 = 1;
      p.epi = TC_EPI_BIAS_RELU_DROP; p.site = SITE_PROJ; p.sub = m;
      tb.add_problem(p);
    }
    if ((rc = tb.flush())) return rc;
  }

  }

  if (projections_only) return MSF_OK;
  const bool use_chain = chain_eligible(H, M) && !getenv("MSF_NO_CHAIN");
  if (use_chain) {  // F2 + F3 chained per (window tile, query): U never leaves the SM on its way to out_proj
    ChainLaunch C;
    memset(&C, 0, sizeof(C));
    C.mode = 0; C.M = M; C.H = H; C.heads = L.heads; C.head_dim = H / L.heads; C.rows = (int)B;
    C.store1 = 1;  // U is an operand of the out_proj weight gradient
    const int np = pairs > 0 ? pairs : 1;
    if ((rc = tc_encode_map(&C.map_a1, ws.P, B, H, H, M, BH, 64, 128))) return rc;
    if ((rc = tc_encode_map(&C.map_w1, pairs > 0 ? (const void*)(W16 + A.wv) : (const void*)ws.P, H, H, H, np, (long long)H * H, 64, chain_w1_box_rows(H, M, L.heads, B)))) return rc;
    if ((rc = tc_encode_map(&C.map_w2, pairs > 0 ? (const void*)(W16 + A.wo) : (const void*)ws.P, H, H, H, np, (long long)H * H, 64, chain_w2_box_rows(H, M, L.heads, B)))) return rc;
    if ((rc = tc_encode_map(&C.map_out1, ws.U, B, H, H, np, BH, 64, 128))) return rc;
    if ((rc = tc_encode_map(&C.map_out, ws.agg, B, H, H, M, BH, 64, 128))) return rc;
    for (int q = 0; q < M; ++q) {
      ChainOuter& O = C.outer[q];
      const bool q_absent = present_hint != 0u && !((present_hint >> q) & 1u);
      for (int k = 0; k < M; ++k) {
        if (q == k || !L.has_pair(q, k) || q_absent) continue;
        const int pi = L.pair_index(q, k);
        C.bias1[pi] = W + L.pair_b(pi, 2);
        C.bias2[pi] = W + L.pair_b(pi, 3);
        if (present_hint != 0u && !((present_hint >> k) & 1u)) {   // absent key: bias only
          O.bias_only[O.nb++] = (short)pi;
          continue;
        }
        O.inner[O.n] = (short)k; O.pair[O.n] = (short)pi; O.sub[O.n] = (short)(q * M + k); O.mask_col[O.n] = (short)k;
        ++O.n;
      }
      C.inv_cnt[q] = 1.0f / (float)L.mean_count(q);
      if (present_hint != 0u && !q_absent) C.active[C.n_active++] = (short)q;   // absent queries get no items at all
    }
    C.gate_out = ws.G; C.mask = c->mask; C.aux = ws.P; C.drop = drop;
    if ((rc = chain_launch(C, st, "F2+F3 value->out chain"))) return rc;
  } else {
  if (pairs > 0) {  // F2: U_qk = g_qk * (P_k Wv_qk^T + bv_qk)
      TcBuilder tb(false, bnH, drop, st, "F2 value_proj");
      const short mapP = (short)tb.add_map(ws.P, B, H, H, M, BH, TC_BLOCK_M);
      const short mapW = (short)tb.add_map(W16 + A.wv, H, H, H, pairs, (long long)H * H, bnH);
      for (int q = 0; q < M; ++q)
        for (int k = 0; k < M; ++k) {
          if (q == k || !L.has_pair(q, k)) continue;
          const int pi = L.pair_index(q, k);
          TcProblem p = tc_blank_problem();
          p.seg[0].a_map = mapP; p.seg[0].a_z = k;
          p.seg[0].b_map = mapW; p.seg[0].b_z = pi;
          p.bias[0] = W + L.pair_b(pi, 2);
          p.M = (int)B; p.N = H; p.K = H;
          p.C = ws.U + (long long)pi * BH; p.ldc = H; p.c_bf16 = 1;
          p.epi = TC_EPI_VALUE_GATE;
          p.mask = c->mask; p.mask_ld = M; p.mask_col = k;
          p.gate_out = ws.G + (long long)pi * B * L.heads;
          p.head_dim = H / L.heads; p.heads = L.heads;
          p.sub = q * M + k;
          tb.add_problem(p);
        }
      if ((rc = tb.flush())) return rc;
    }
  
    {  // F3: agg_q = (P_q + sum_k (U_qk Wo_qk^T + bo_qk)) / cnt_q * mask_q
      TcBuilder tb(false, bnH, drop, st, "F3 out_proj+mean");
      const short mapU = (short)tb.add_map(ws.U, B, H, H, pairs > 0 ? pairs : 1, BH, TC_BLOCK_M);
      const short mapW = pairs > 0 ? (short)tb.add_map(W16 + A.wo, H, H, H, pairs, (long long)H * H, bnH) : mapU;
      for (int q = 0; q < M; ++q) {
        TcProblem p = tc_blank_problem();
        int seg = 0;
        for (int k = 0; k < M; ++k) {
          if (q == k || !L.has_pair(q, k)) continue;
          const int pi = L.pair_index(q, k);
          p.seg[seg].a_map = mapU; p.seg[seg].a_z = pi;
          p.seg[seg].b_map = mapW; p.seg[seg].b_z = pi;
          p.bias[seg] = W + L.pair_b(pi, 3);
          ++seg;
        }
        p.M = (int)B; p.N = H; p.K = seg ? H : 0;
        if (seg == 0) { p.seg[0].a_map = mapU; p.seg[0].b_map = mapW; seg = 1; }  // epilogue only
        p.nseg = seg;
        p.C = ws.agg + (long long)q * BH; p.ldc = H; p.c_bf16 = 1;
        p.epi = TC_EPI_OUT_MEAN; p.scale = (float)L.mean_count(q);
        p.aux = ws.P + (long long)q * BH; p.ld_aux = H;
        p.mask = c->mask; p.mask_ld = M; p.mask_col = q;
        tb.add_problem(p);
      }
      if ((rc = tb.flush())) return rc;
    }
  
  }
  return MSF_OK;
}

// Head of the network as ONE kernel (head_gemm.cu).  train: labels etc. must be set by the caller in `hl`.
static int launch_head(const Layout& L, const msf_fusion_call* c, const WsBf16& ws, const ArenaBf16& A, HeadLaunch& hl,
                       cudaStream_t st, const char* label) {
  const int64_t B = c->batch;
  const int M = L.M, H = L.H;
  const bf16* W16 = reinterpret_cast<const bf16*>(c->params_bf16);
  const float* W = c->params;
  int rc;
  if ((rc = tc_encode_map(&hl.map_w1, W16 + A.w1, H, H, H, 1, 0, 64, H))) return rc;
  if ((rc = tc_encode_map(&hl.map_w2, W16 + A.w2, L.C, H, H, 1, 0, 64, 32))) return rc;
  if ((rc = tc_encode_map(&hl.map_w2t, W16 + A.w2T, H, A.Cp, A.Cp, 1, 0, 64, H))) return rc;
  if ((rc = tc_encode_map(&hl.map_w1t, W16 + A.w1T, H, H, H, 1, 0, 64, H))) return rc;
  if ((rc = tc_encode_map(&hl.map_agg, ws.agg, B, H, H, M, (long long)B * H, 64, 32))) return rc;
  const int tr = head_tile_rows(B);
  if ((rc = tc_encode_map(&hl.map_fused, ws.fused, B, H, H, 1, 0, 64, tr))) return rc;
  if ((rc = tc_encode_map(&hl.map_hr, ws.Hr, B, H, H, 1, 0, 64, tr))) return rc;
  if ((rc = tc_encode_map(&hl.map_dh1, ws.dH1, B, H, H, 1, 0, 64, tr))) return rc;
  hl.M = M; hl.H = H; hl.C = L.C; hl.Cp = A.Cp; hl.rows = (int)B;
  hl.agg = ws.agg;
  for (int m = 0; m < M; ++m) {
    hl.gate_w[m] = W + L.gate_w[m];
    hl.gate_b[m] = W + L.gate_b[m];
    hl.inv_cnt[m] = 1.0f / (float)L.mean_count(m);
  }
  hl.mask = c->mask; hl.soft = ws.soft; hl.w = ws.w; hl.w_out = c->fusion_weights;
  hl.b1 = W + L.cls_b1; hl.b2 = W + L.cls_b2;
  hl.logits = c->logits;
  hl.drop = make_drop(c);
  return head_launch(hl, st, label);
}

static bool use_head(const Layout& L) { return head_eligible(L.H, L.M, L.C) && !getenv("MSF_NO_HEAD"); }

static int export_gates(const Layout& L, const msf_fusion_call* c, const WsBf16& ws, cudaStream_t st);

int fusion_bf16_forward(const Layout& L, const msf_fusion_call* c, cudaStream_t st) {
  const int64_t B = c->batch;
  const int M = L.M, H = L.H;
  WsBf16 ws;
  carve_bf16(L, B, c->workspace, &ws);
  int rc = check_ws(ws, c);
  if (rc) return rc;
  const ArenaBf16 A = arena_layout(L);
  const bf16* W16 = reinterpret_cast<const bf16*>(c->params_bf16);
  const float* W = c->params;
  const DropCfg drop = make_drop(c);
  if ((rc = forward_front(L, c, ws, A, st))) return rc;

  if (use_head(L)) {  // F4 + F5 + F6 in one kernel
    HeadLaunch hl;
    memset(&hl, 0, sizeof(hl));
    hl.train = 0; hl.store_acts = 1;
    if ((rc = launch_head(L, c, ws, A, hl, st, "HEAD gating+classifier"))) return rc;
    return export_gates(L, c, ws, st);
  }

  {  // F4
    Tail16Args a;
    memset(&a, 0, sizeof(a));
    a.agg = ws.agg;
    for (int m = 0; m < M; ++m) {
      a.gate_w[m] = W + L.gate_w[m];
      a.gate_b[m] = W + L.gate_b[m];
    }
    a.mask = c->mask; a.soft = ws.soft; a.w = ws.w; a.w_out = c->fusion_weights; a.fused = ws.fused;
    a.B = B; a.M = M; a.H = H;
    tail16_fwd_kernel<<<(unsigned)ceil_div(B, 8), 256, 0, st>>>(a);
    MSF_LAUNCH_CHECK();
  }

  {  // F5: Hr = drop3(relu(fused W1^T + b1))
    const int bn1 = block_n_fill(H, B, 1);
    TcBuilder tb(false, bn1, drop, st, "F5 classifier.0");
    TcProblem p = tc_blank_problem();
    p.seg[0].a_map = (short)tb.add_map(ws.fused, B, H, H, 1, 0, TC_BLOCK_M);
    p.seg[0].b_map = (short)tb.add_map(W16 + A.w1, H, H, H, 1, 0, bn1);
    p.bias[0] = W + L.cls_b1;
    p.M = (int)B; p.N = H; p.K = H;
    p.C = ws.Hr; p.ldc = H; p.c_bf16 = 1;
    p.epi = TC_EPI_BIAS_RELU_DROP; p.site = SITE_CLS; p.sub = 0;
    tb.add_problem(p);
    if ((rc = tb.flush())) return rc;
  }
  {  // F6: logits = Hr W2^T + b2 (fp32 out)
    const int bnC = L.C <= 32 ? 32 : (L.C > 64 ? 128 : 64);  // fp32 output: at most 128 columns per tile
    TcBuilder tb(false, bnC, drop, st, "F6 classifier.3");
    TcProblem p = tc_blank_problem();
    p.seg[0].a_map = (short)tb.add_map(ws.Hr, B, H, H, 1, 0, TC_BLOCK_M);
    p.seg[0].b_map = (short)tb.add_map(W16 + A.w2, L.C, H, H, 1, 0, bnC);
    p.bias[0] = W + L.cls_b2;
    p.M = (int)B; p.N = L.C; p.K = H;
    p.C = c->logits; p.ldc = L.C; p.c_bf16 = 0;
    p.epi = TC_EPI_STORE;
    tb.add_problem(p);
    if ((rc = tb.flush())) return rc;
  }

  return export_gates(L, c, ws, st);
}

static int export_gates(const Layout& L, const msf_fusion_call* c, const WsBf16& ws, cudaStream_t st) {
  const int64_t B = c->batch;
  const int M = L.M, pairs = L.num_pairs();
  if (c->attn_gates && pairs > 0) {
    const long long ng = (long long)pairs * B * L.heads;
    for (int q = 0; q < M; ++q)
      for (int k = 0; k < M; ++k)
        if (q != k && !L.has_pair(q, k))
          MSF_CHECK_CUDA(cudaMemsetAsync(ws.G + (size_t)L.pair_index(q, k) * B * L.heads, 0,
                                         (size_t)B * L.heads * sizeof(float), st));
    copy_gates16_kernel<<<(unsigned)(ceil_div(ng, 256) < 1184 ? ceil_div(ng, 256) : 1184), 256, 0, st>>>(
        ws.G, c->attn_gates, ng);
    MSF_LAUNCH_CHECK();
  }
  return MSF_OK;
}

// Bias and gating-layer gradients that are (weighted) column sums of the head's outputs: classifier.0 bias (dH1),
// the gating layers (ds, agg) and the out_proj biases (dS_q, the same sum for every key of query q).
static int colsum_head(const Layout& L, const msf_fusion_call* c, const WsBf16& ws, cudaStream_t st, bool head_fused) {
  const int64_t B = c->batch;
  const int M = L.M, H = L.H, C = L.C;
  const long long BH = (long long)B * H;
  float* dW = c->grad_params;
  Colsum16Problem cs[3 * MSF_MAX_MODALITIES + 2];
  int nc = 0;
  auto add = [&](const void* src, int f32, long long ld, int cols, const float* coef, int coef_ld, float* dst) {
    Colsum16Problem p;
    memset(&p, 0, sizeof(p));
    p.src = src; p.src_f32 = f32; p.ld = ld; p.rows = (int)B; p.cols = cols;
    p.coef = coef; p.coef_ld = coef_ld; p.dst = dst;
    cs[nc++] = p;
  };
  if (!head_fused) add(c->grad_logits, 1, C, C, nullptr, 0, dW + L.cls_b2);
  add(ws.dH1, 0, H, H, nullptr, 0, dW + L.cls_b1);
  for (int q = 0; q < M; ++q) {
    add(ws.agg + (long long)q * BH, 0, H, H, ws.ds + q, M, dW + L.gate_w[q]);   // d gate_w_q = sum_r ds[r,q] agg_q[r,:]
    add(ws.ds + q, 1, M, 1, nullptr, 0, dW + L.gate_b[q]);
    int nd = 0;
    for (int k = 0; k < M; ++k) {
      if (q == k || !L.has_pair(q, k)) continue;
      float* dst = dW + L.pair_b(L.pair_index(q, k), 3);
      if (nd == 0) add(ws.dS + (long long)q * BH, 0, H, H, nullptr, 0, dst);
      else cs[nc - 1].dst_more[nd - 1] = dst;
      ++nd;
    }
  }
  return colsum16_launch(cs, nc, st);
}

static int backward_back(const Layout& L, const msf_fusion_call* c, const WsBf16& ws, const ArenaBf16& A,
                         cudaStream_t st, bool head_fused);

int fusion_bf16_backward(const Layout& L, const msf_fusion_call* c, cudaStream_t st) {
  const int64_t B = c->batch;
  const int M = L.M, H = L.H, C = L.C;
  WsBf16 ws;
  carve_bf16(L, B, c->workspace, &ws);
  int rc = check_ws(ws, c);
  if (rc) return rc;
  MSF_REQUIRE(c->grad_logits && c->grad_params, "backward needs grad_logits and grad_params");
  const ArenaBf16 A = arena_layout(L);
  const bf16* W16 = reinterpret_cast<const bf16*>(c->params_bf16);
  const float* W = c->params;
  float* dW = c->grad_params;
  const DropCfg drop = make_drop(c);
  const int Cp = A.Cp;

  // dead query/key projections and every slot accumulated below start from exact zeros
  MSF_CHECK_CUDA(cudaMemsetAsync(dW, 0, (size_t)L.total * sizeof(float), st));
  {
    const long long total = B * Cp;
    cvt_pad_kernel<<<(unsigned)(ceil_div(total, 256) < 1184 ? ceil_div(total, 256) : 1184), 256, 0, st>>>(
        c->grad_logits, ws.dlog, B, C, Cp);
    MSF_LAUNCH_CHECK();
  }

  const int bn1 = block_n_fill(H, B, 1);
  {  // B1: dH1 = (dlogits W2) * relu'(Hr) * drop3
    TcBuilder tb(false, bn1, drop, st, "B1 d classifier.3");
    TcProblem p = tc_blank_problem();
    p.seg[0].a_map = (short)tb.add_map(ws.dlog, B, Cp, Cp, 1, 0, TC_BLOCK_M);
    p.seg[0].b_map = (short)tb.add_map(W16 + A.w2T, H, Cp, Cp, 1, 0, bn1);
    p.M = (int)B; p.N = H; p.K = Cp;
    p.C = ws.dH1; p.ldc = H; p.c_bf16 = 1;
    p.epi = TC_EPI_RELU_GRAD; p.scale = drop.scale; p.aux = ws.Hr; p.ld_aux = H;
    tb.add_problem(p);
    if ((rc = tb.flush())) return rc;
  }
  {  // B2: dfused = dH1 W1
    TcBuilder tb(false, bn1, drop, st, "B2 d classifier.0");
    TcProblem p = tc_blank_problem();
    p.seg[0].a_map = (short)tb.add_map(ws.dH1, B, H, H, 1, 0, TC_BLOCK_M);
    p.seg[0].b_map = (short)tb.add_map(W16 + A.w1T, H, H, H, 1, 0, bn1);
    p.M = (int)B; p.N = H; p.K = H;
    p.C = ws.dfused; p.ldc = H; p.c_bf16 = 1;
    p.epi = TC_EPI_STORE;
    tb.add_problem(p);
    if ((rc = tb.flush())) return rc;
  }
  {  // B4: tail backward
    Tail16BwdArgs a;
    memset(&a, 0, sizeof(a));
    a.agg = ws.agg; a.dfused = ws.dfused; a.mask = c->mask; a.soft = ws.soft; a.w = ws.w;
    a.dS = ws.dS; a.ds = ws.ds;
    for (int m = 0; m < M; ++m) {
      a.gate_w[m] = W + L.gate_w[m];
      a.inv_cnt[m] = 1.0f / (float)L.mean_count(m);
    }
    a.B = B; a.M = M; a.H = H;
    tail16_bwd_kernel<<<(unsigned)ceil_div(B, 8), 256, 0, st>>>(a);
    MSF_LAUNCH_CHECK();
  }
  return backward_back(L, c, ws, A, st, false);
}

// From d aggregated (ws.dS) back to the inputs, plus every bias / weight gradient.  head_fused: the head
// kernel already produced the classifier.3 bias gradient.
static int backward_back(const Layout& L, const msf_fusion_call* c, const WsBf16& ws, const ArenaBf16& A,
                         cudaStream_t st, bool head_fused) {
  const int64_t B = c->batch;
  const int M = L.M, H = L.H, C = L.C;
  int rc = MSF_OK;
  const bf16* W16 = reinterpret_cast<const bf16*>(c->params_bf16);
  float* dW = c->grad_params;
  const DropCfg drop = make_drop(c);
  const long long BH = (long long)B * H;
  const int pairs = L.num_pairs();
  const int bnH = block_n_for(H);
  const int Cp = A.Cp;
  const bool use_chain = chain_eligible(H, M) && !getenv("MSF_NO_CHAIN");
  if (use_chain) {  // B5 + B7 chained per (window tile, key): dV goes straight from the epilogue into the value_proj dgrad
    ChainLaunch C;
    memset(&C, 0, sizeof(C));
    C.mode = 1; C.M = M; C.H = H; C.heads = L.heads; C.head_dim = H / L.heads; C.rows = (int)B;
    C.store1 = 1;  // dV is an operand of the value_proj weight / bias gradients
    const int np = pairs > 0 ? pairs : 1;
    if ((rc = tc_encode_map(&C.map_a1, ws.dS, B, H, H, M, BH, 64, 128))) return rc;
    if ((rc = tc_encode_map(&C.map_w1, pairs > 0 ? (const void*)(W16 + A.woT) : (const void*)ws.dS, H, H, H, np, (long long)H * H, 64, chain_w1_box_rows(H, M, L.heads, B)))) return rc;
    if ((rc = tc_encode_map(&C.map_w2, pairs > 0 ? (const void*)(W16 + A.wvT) : (const void*)ws.dS, H, H, H, np, (long long)H * H, 64, chain_w2_box_rows(H, M, L.heads, B)))) return rc;
    if ((rc = tc_encode_map(&C.map_out1, ws.dV, B, H, H, np, BH, 64, 128))) return rc;
    if ((rc = tc_encode_map(&C.map_out, ws.dZ, B, H, H, M, BH, 64, 128))) return rc;
    if ((rc = tc_encode_map(&C.map_aux2, ws.P, B, H, H, M, BH, 64, 128))) return rc;
    for (int k = 0; k < M; ++k) {
      ChainOuter& O = C.outer[k];
      for (int q = 0; q < M; ++q) {
        if (q == k || !L.has_pair(q, k)) continue;
        O.inner[O.n] = (short)q; O.pair[O.n] = (short)L.pair_index(q, k);
        ++O.n;
      }
    }
    C.gate_in = ws.G; C.aux = ws.dS; C.aux2 = ws.P; C.scale = drop.scale; C.drop = drop;
    if ((rc = chain_launch(C, st, "B5+B7 d-out->d-value chain"))) return rc;
  } else {
  if (pairs > 0) {  // B5: dV_qk = (dS_q Wo_qk) * g_qk
      TcBuilder tb(false, bnH, drop, st, "B5 d out_proj");
      const short mapS = (short)tb.add_map(ws.dS, B, H, H, M, BH, TC_BLOCK_M);
      const short mapW = (short)tb.add_map(W16 + A.woT, H, H, H, pairs, (long long)H * H, bnH);
      for (int q = 0; q < M; ++q)
        for (int k = 0; k < M; ++k) {
          if (q == k || !L.has_pair(q, k)) continue;
          const int pi = L.pair_index(q, k);
          TcProblem p = tc_blank_problem();
          p.seg[0].a_map = mapS; p.seg[0].a_z = q;
          p.seg[0].b_map = mapW; p.seg[0].b_z = pi;
          p.M = (int)B; p.N = H; p.K = H;
          p.C = ws.dV + (long long)pi * BH; p.ldc = H; p.c_bf16 = 1;
          p.epi = TC_EPI_GATE_MUL; p.gate_in = ws.G + (long long)pi * B * L.heads;
          p.head_dim = H / L.heads; p.heads = L.heads;
          tb.add_problem(p);
        }
      if ((rc = tb.flush())) return rc;
    }
    {  // B7: dZ_k = (dS_k + sum_q dV_qk Wv_qk) * relu'(P_k) * drop1
      TcBuilder tb(false, bnH, drop, st, "B7 d value_proj");
      const short mapV = (short)tb.add_map(ws.dV, B, H, H, pairs > 0 ? pairs : 1, BH, TC_BLOCK_M);
      const short mapW = pairs > 0 ? (short)tb.add_map(W16 + A.wvT, H, H, H, pairs, (long long)H * H, bnH) : mapV;
      for (int k = 0; k < M; ++k) {
        TcProblem p = tc_blank_problem();
        int seg = 0;
        for (int q = 0; q < M; ++q) {
          if (q == k || !L.has_pair(q, k)) continue;
          const int pi = L.pair_index(q, k);
          p.seg[seg].a_map = mapV; p.seg[seg].a_z = pi;
          p.seg[seg].b_map = mapW; p.seg[seg].b_z = pi;
          ++seg;
        }
        p.M = (int)B; p.N = H; p.K = seg ? H : 0;
        if (seg == 0) { p.seg[0].a_map = mapV; p.seg[0].b_map = mapW; seg = 1; }
        p.nseg = seg;
        p.C = ws.dZ + (long long)k * BH; p.ldc = H; p.c_bf16 = 1;
        p.epi = TC_EPI_ADD_RELU_GRAD; p.scale = drop.scale;
        p.aux = ws.dS + (long long)k * BH; p.ld_aux = H;
        p.aux2 = ws.P + (long long)k * BH; p.ld_aux2 = H;
        tb.add_problem(p);
      }
      if ((rc = tb.flush())) return rc;
    }
  }
  {  // B9a: dx_m = (dZ_m Wp_m) * mask_m * drop0   (fp32 out, only where requested)
    for (int m = 0; m < M; ++m) {
      if (!c->grad_x[m]) continue;
      const int bnD = L.D[m] > 64 ? 128 : 64;  // fp32 output: tiles of at most 128 columns
      TcBuilder tb(false, bnD, drop, st, "B9 d projections (dx)");
      TcProblem p = tc_blank_problem();
      p.seg[0].a_map = (short)tb.add_map(ws.dZ + (long long)m * BH, B, H, H, 1, 0, TC_BLOCK_M);
      p.seg[0].b_map = (short)tb.add_map(W16 + A.wpT[m], L.D[m], H, H, 1, 0, bnD);
      p.M = (int)B; p.N = L.D[m]; p.K = H;
      p.C = c->grad_x[m]; p.ldc = L.D[m]; p.c_bf16 = 0;
      p.epi = TC_EPI_DX; p.site = SITE_INPUT; p.sub = m;
      p.mask = c->mask; p.mask_ld = M; p.mask_col = m;
      tb.add_problem(p);
      if ((rc = tb.flush())) return rc;
    }
  }

  // fork: the column sums only read activations that are complete at this point.  Those over the head's outputs
  // were launched by the caller when the head is fused (colsum_head: they run under the chain kernel above).
  SideStream* side = nullptr;
  if ((rc = side_stream(&side))) return rc;
  MSF_CHECK_CUDA(cudaEventRecord(side->fork, st));
  MSF_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
  if (!head_fused && (rc = colsum_head(L, c, ws, side->stream, false))) return rc;
  {  // value_proj and projection biases: column sums of the chain's outputs
    Colsum16Problem cs[MSF_MAX_MODALITIES * MSF_MAX_MODALITIES + MSF_MAX_MODALITIES];
    int nc = 0;
    auto add = [&](const bf16* src, float* dst) {
      Colsum16Problem p;
      memset(&p, 0, sizeof(p));
      p.src = src; p.ld = H; p.rows = (int)B; p.cols = H; p.dst = dst;
      cs[nc++] = p;
    };
    for (int q = 0; q < M; ++q) {
      add(ws.dZ + (long long)q * BH, dW + L.proj_b[q]);
      for (int k = 0; k < M; ++k)
        if (q != k && L.has_pair(q, k)) add(ws.dV + (long long)L.pair_index(q, k) * BH, dW + L.pair_b(L.pair_index(q, k), 2));
    }
    if ((rc = colsum16_launch(cs, nc, side->stream, 1))) return rc;
  }
  if (head_fused && c->grad_sq != nullptr) {   // every bias / gating slot is complete on this stream now
    VecSqList vl;
    memset(&vl, 0, sizeof(vl));
    auto seg = [&](long long begin, long long count) {
      if (count > 0) { vl.begin[vl.n] = begin; vl.count[vl.n] = (int)count; ++vl.n; }
    };
    for (int m = 0; m < M; ++m) seg(L.proj_b[m], H);
    for (int p = 0; p < pairs; ++p) {
      seg(L.pair_b(p, 2), H);
      seg(L.pair_b(p, 3), H);
    }
    seg(L.gate_w[0], L.cls_w1 - L.gate_w[0]);
    seg(L.cls_b1, H);
    seg(L.cls_b2, C);
    vec_sq_kernel<<<vl.n, 256, 0, side->stream>>>(vl, dW, c->grad_sq);
    MSF_LAUNCH_CHECK();
  }
  // ---- all weight gradients: one MN-major launch, dW[out,in] = dY^T . X over the windows ----
  bool wg_pairs = wg2_enabled() && wg2_shape_ok(H, H, dW + L.cls_w1) && wg2_shape_ok(H, H, dW + L.cls_w2) &&
                  (pairs == 0 || (wg2_shape_ok(H, H, dW + L.pair_w(0, 2)) && L.pair_stride % 4 == 0 && (H * H + H) % 4 == 0));
  for (int m = 0; m < M; ++m) wg_pairs = wg_pairs && wg2_shape_ok(L.D[m], L.D[m], dW + L.proj_w[m]);
  wg_pairs = wg_pairs && 9 + M <= WG2_MAX_MAPS;   // more than WG2_MAX_PROBLEMS problems go out as several launches
  if (wg_pairs) {   // CTA pairs, 256 x 256 tiles, the contraction split in two halves (wg2_gemm.cu)
    Wg2Builder wb(B, st, "WG weight gradients");
    const short mapDlog = (short)wb.add_map(ws.dlog, B, Cp, Cp, 1, 0);
    const short mapHr = (short)wb.add_map(ws.Hr, B, H, H, 1, 0);
    const short mapDH1 = (short)wb.add_map(ws.dH1, B, H, H, 1, 0);
    const short mapFused = (short)wb.add_map(ws.fused, B, H, H, 1, 0);
    const short mapDS = (short)wb.add_map(ws.dS, B, H, H, M, BH);
    const short mapU = (short)wb.add_map(ws.U, B, H, H, pairs > 0 ? pairs : 1, BH);
    const short mapDV = (short)wb.add_map(ws.dV, B, H, H, pairs > 0 ? pairs : 1, BH);
    const short mapP = (short)wb.add_map(ws.P, B, H, H, M, BH);
    const short mapDZ = (short)wb.add_map(ws.dZ, B, H, H, M, BH);
    double* sq = head_fused ? c->grad_sq : nullptr;   // cleared by the head kernel of the same pass
    for (int q = 0; q < M; ++q)   // the heavy tiles first
      for (int k = 0; k < M; ++k) {
        if (q == k || !L.has_pair(q, k)) continue;
        const int pi = L.pair_index(q, k);
        wb.add_problem(mapDS, q, mapU, pi, H, H, dW + L.pair_w(pi, 3), H, sq);    // dWo_qk = dS_q^T U_qk
        wb.add_problem(mapDV, pi, mapP, k, H, H, dW + L.pair_w(pi, 2), H, sq);    // dWv_qk = dV_qk^T P_k
      }
    wb.add_problem(mapDH1, 0, mapFused, 0, H, H, dW + L.cls_w1, H, sq);           // dW1 = dH1^T fused
    for (int m = 0; m < M; ++m) {
      const short mapX = (short)wb.add_map(ws.xt[m], B, L.D[m], L.D[m], 1, 0);
      wb.add_problem(mapDZ, m, mapX, 0, H, L.D[m], dW + L.proj_w[m], L.D[m], sq);  // dWp_m = dZ_m^T xt_m
    }
    wb.add_problem(mapDlog, 0, mapHr, 0, C, H, dW + L.cls_w2, H, sq);             // dW2 = dlogits^T Hr
    if ((rc = wb.flush())) return rc;
  } else {
    const int bn = 128;
    TcBuilder tb(true, bn, drop, st, "WG weight gradients");
    const short mapDlog = (short)tb.add_map(ws.dlog, B, Cp, Cp, 1, 0, 0);
    const short mapHr = (short)tb.add_map(ws.Hr, B, H, H, 1, 0, 0);
    const short mapDH1 = (short)tb.add_map(ws.dH1, B, H, H, 1, 0, 0);
    const short mapFused = (short)tb.add_map(ws.fused, B, H, H, 1, 0, 0);
    const short mapDS = (short)tb.add_map(ws.dS, B, H, H, M, BH, 0);
    const short mapU = (short)tb.add_map(ws.U, B, H, H, pairs > 0 ? pairs : 1, BH, 0);
    const short mapDV = (short)tb.add_map(ws.dV, B, H, H, pairs > 0 ? pairs : 1, BH, 0);
    const short mapP = (short)tb.add_map(ws.P, B, H, H, M, BH, 0);
    const short mapDZ = (short)tb.add_map(ws.dZ, B, H, H, M, BH, 0);
    auto wgrad = [&](short am, int az, short bm, int bz, int rows, int cols, float* dst) {
      TcProblem g = tc_blank_problem();
      g.seg[0].a_map = am; g.seg[0].a_z = az;
      g.seg[0].b_map = bm; g.seg[0].b_z = bz;
      g.M = rows; g.N = cols; g.K = (int)B;
      g.C = dst; g.ldc = cols; g.c_bf16 = 0; g.epi = TC_EPI_STORE;
      g.sq = head_fused ? c->grad_sq : nullptr;   // cleared by the head kernel of the same pass
      tb.add_problem(g);
    };
    wgrad(mapDlog, 0, mapHr, 0, C, H, dW + L.cls_w2);      // dW2 = dlogits^T Hr
    wgrad(mapDH1, 0, mapFused, 0, H, H, dW + L.cls_w1);    // dW1 = dH1^T fused
    for (int q = 0; q < M; ++q)
      for (int k = 0; k < M; ++k) {
        if (q == k || !L.has_pair(q, k)) continue;
        const int pi = L.pair_index(q, k);
        wgrad(mapDS, q, mapU, pi, H, H, dW + L.pair_w(pi, 3));   // dWo_qk = dS_q^T U_qk
        wgrad(mapDV, pi, mapP, k, H, H, dW + L.pair_w(pi, 2));   // dWv_qk = dV_qk^T P_k
      }
    // projections: xt_m may have a different width per modality -> one descriptor each
    for (int m = 0; m < M; ++m) {
      if (tb.nmaps >= TC_MAX_MAPS) {  // descriptor table full: launch and start a new one
        if ((rc = tb.flush())) return rc;
        tb.nmaps = 9;  // keep the shared maps above
      }
      const short mapX = (short)tb.add_map(ws.xt[m], B, L.D[m], L.D[m], 1, 0, 0);
      wgrad(mapDZ, m, mapX, 0, H, L.D[m], dW + L.proj_w[m]);     // dWp_m = dZ_m^T xt_m
    }
    if ((rc = tb.flush())) return rc;
  }

  MSF_CHECK_CUDA(cudaEventRecord(side->join, side->stream));
  MSF_CHECK_CUDA(cudaStreamWaitEvent(st, side->join, 0));
  return MSF_OK;
}

// One training pass: forward, CE(label smoothing) and backward with the head of the network fused into
// one kernel (no logits / d logits round trip, 6 launches fewer than forward + msf_cross_entropy + backward).
int fusion_bf16_train(const Layout& L, const msf_fusion_call* c, const int64_t* labels, float smoothing,
                      float grad_scale, float* row_loss, float* loss_out, int flags, cudaStream_t st) {
  const int64_t B = c->batch;
  WsBf16 ws;
  carve_bf16(L, B, c->workspace, &ws);
  int rc = check_ws(ws, c);
  if (rc) return rc;
  MSF_REQUIRE(c->grad_params && c->logits && labels && row_loss, "train pass needs grad_params, logits, labels, row_loss");
  MSF_REQUIRE(use_head(L), "fused train pass not available for this shape");
  const ArenaBf16 A = arena_layout(L);
  float* dW = c->grad_params;
  // every gradient slot that is accumulated into (or never written) starts from zero; see ZeroList
  ZeroList zl;
  memset(&zl, 0, sizeof(zl));
  zl.base = dW;
  bool fits = true;
  auto zr = [&](long long begin, long long count, int batch, long long stride) {
    if (count <= 0 || batch <= 0) return;
    if (zl.n >= PREP_MAX_ZERO) { fits = false; return; }
    zl.r[zl.n].begin = begin; zl.r[zl.n].count = count; zl.r[zl.n].batch = batch; zl.r[zl.n].stride = stride;
    ++zl.n;
  };
  {
    const long long H = L.H, half = 2 * (H * H + H);
    const int pairs = L.num_pairs();
    for (int m = 0; m < L.M; ++m) zr(L.proj_b[m], H, 1, 0);
    if (pairs > 0) {
      zr(L.pair_b(0, 2), H, pairs, L.pair_stride);
      zr(L.pair_b(0, 3), H, pairs, L.pair_stride);
      if (!(flags & MSF_TRAIN_DEAD_SLOTS_ZERO)) zr(L.pair_w(0, 0), half, pairs, L.pair_stride);
    }
    for (int q = 0; q < L.M; ++q)
      for (int k = 0; k < L.M; ++k)
        if (q != k && !L.has_pair(q, k)) zr(L.pair_w(L.pair_index(q, k), 2), half, 1, 0);   // deleted module: no writer
    zr(L.gate_w[0], L.cls_w1 - L.gate_w[0], 1, 0);
    zr(L.cls_b1, H, 1, 0);
    zr(L.cls_b2, L.C, 1, 0);
  }
  if (!fits) {
    MSF_CHECK_CUDA(cudaMemsetAsync(dW, 0, (size_t)L.total * sizeof(float), st));
    zl.n = 0;
  }
  if ((rc = forward_front(L, c, ws, A, st, &zl))) return rc;
  HeadLaunch hl;
  memset(&hl, 0, sizeof(hl));
  hl.train = 1; hl.store_acts = 1;
  hl.labels = reinterpret_cast<const long long*>(labels);
  hl.smoothing = smoothing; hl.grad_scale = grad_scale; hl.row_loss = row_loss;
  hl.loss_out = nullptr;   // the mean is formed on the side stream below
  hl.dlog = ws.dlog; hl.db2 = dW + L.cls_b2; hl.dS = ws.dS; hl.ds = ws.ds;
  hl.grad_sq = c->grad_sq;
  if ((rc = launch_head(L, c, ws, A, hl, st, "HEAD gating+classifier+CE fwd/bwd"))) return rc;
  if ((rc = export_gates(L, c, ws, st))) return rc;
  {  // the column sums over the head's outputs run beside the d-out -> d-value chain
    SideStream* side = nullptr;
    if ((rc = side_stream(&side))) return rc;
    MSF_CHECK_CUDA(cudaEventRecord(side->fork2, st));
    MSF_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork2, 0));
    if (loss_out != nullptr) {
      loss_mean_kernel<<<1, 256, 0, side->stream>>>(row_loss, c->batch, loss_out);
      MSF_LAUNCH_CHECK();
    }
    if ((rc = colsum_head(L, c, ws, side->stream, true))) return rc;
  }
  return backward_back(L, c, ws, A, st, true);
}

// Inference pass: logits plus softmax -> (confidence, prediction) from the head kernel's epilogue.
int fusion_bf16_infer(const Layout& L, const msf_fusion_call* c, float* conf, int64_t* pred, unsigned present_hint,
                      cudaStream_t st) {
  WsBf16 ws;
  carve_bf16(L, c->batch, c->workspace, &ws);
  int rc = check_ws(ws, c);
  if (rc) return rc;
  MSF_REQUIRE(use_head(L), "fused inference pass not available for this shape");
  MSF_REQUIRE(present_hint == 0u || (c->attn_gates == nullptr && c->mask != nullptr && chain_eligible(L.H, L.M)),
              "a uniform-mask hint needs a mask, no attention-gate export and the chained pair kernel");
  const ArenaBf16 A = arena_layout(L);
  if ((rc = forward_front(L, c, ws, A, st, nullptr, present_hint))) return rc;
  HeadLaunch hl;
  memset(&hl, 0, sizeof(hl));
  hl.train = 0; hl.store_acts = 0;
  hl.conf = conf; hl.pred = reinterpret_cast<long long*>(pred);
  if ((rc = launch_head(L, c, ws, A, hl, st, "HEAD gating+classifier+softmax"))) return rc;
  return export_gates(L, c, ws, st);
}

// Inference for a subset of modalities that is present in EVERY row (the sweep of src/eval.py:342-404) with each
// attention module's value_proj -> out_proj pair folded into one matrix.  Without dropout the attention gate of a
// present key is 1 for every head, so out_proj(value_proj(P_k)) = P_k (Wo Wv)^T + (Wo bv + bo): per present query
// ONE GEMM whose K-segments are the present keys, accumulated in TMEM, with the mean epilogue of the un-fused path —
// half the FLOPs of the chained pair kernel and no U round trip.  `wov` = [pairs][H][H] bf16 (Wo Wv, row-major
// [out][in]); `bias_sum` = [M][H] fp32, row q = sum over present keys of (Wo bv + bo) + sum over absent keys of bo.
// flags bit 0: the projections P of the present modalities are still valid in the workspace (same features, an
// earlier call of the sweep); bit 1: compute the projections only.
int fusion_bf16_infer_folded(const Layout& L, const msf_fusion_call* c, const void* wov, const float* bias_sum,
                             unsigned present, int flags, float* conf, int64_t* pred, cudaStream_t st) {
  WsBf16 ws;
  carve_bf16(L, c->batch, c->workspace, &ws);
  int rc = check_ws(ws, c);
  if (rc) return rc;
  const int M = L.M, H = L.H;
  MSF_REQUIRE(use_head(L) && chain_eligible(H, M), "folded inference pass not available for this shape");
  MSF_REQUIRE(present != 0u && present < (1u << M), "msf_fusion_infer_folded: present must name at least one modality");
  MSF_REQUIRE(c->mask != nullptr && c->attn_gates == nullptr, "msf_fusion_infer_folded needs the subset's mask and no gate export");
  MSF_REQUIRE(!c->training, "msf_fusion_infer_folded is an inference pass (attention dropout makes the gates per head)");
  const ArenaBf16 A = arena_layout(L);
  const int64_t B = c->batch;
  const long long BH = (long long)B * H;
  if (!(flags & 1) && (rc = forward_front(L, c, ws, A, st, nullptr, present, true))) return rc;
  if (flags & 2) return MSF_OK;
  MSF_REQUIRE(wov != nullptr && bias_sum != nullptr && conf && pred, "msf_fusion_infer_folded: null argument");
  {
    const int bnH = block_n_for(H);
    const int pairs = L.num_pairs();
    TcBuilder tb(false, bnH, no_dropout(), st, "FOLD present keys -> aggregated tokens");
    const short mapP = (short)tb.add_map(ws.P, B, H, H, M, BH, TC_BLOCK_M);
    const short mapW = (short)tb.add_map(wov, H, H, H, pairs > 0 ? pairs : 1, (long long)H * H, bnH);
    for (int q = 0; q < M; ++q) {
      if (!((present >> q) & 1u)) continue;   // absent query: its token is zero through the row mask
      TcProblem p = tc_blank_problem();
      int seg = 0;
      for (int k = 0; k < M; ++k) {
        if (q == k || !L.has_pair(q, k) || !((present >> k) & 1u)) continue;
        p.seg[seg].a_map = mapP; p.seg[seg].a_z = k;
        p.seg[seg].b_map = mapW; p.seg[seg].b_z = L.pair_index(q, k);
        ++seg;
      }
      p.M = (int)B; p.N = H; p.K = seg ? H : 0;
      if (seg == 0) { p.seg[0].a_map = mapP; p.seg[0].b_map = mapW; seg = 1; }  // epilogue only
      p.nseg = seg;
      p.bias[0] = bias_sum + (long long)q * H;
      p.C = ws.agg + (long long)q * BH; p.ldc = H; p.c_bf16 = 1;
      p.epi = TC_EPI_OUT_MEAN; p.scale = (float)L.mean_count(q);
      p.aux = ws.P + (long long)q * BH; p.ld_aux = H;
      p.mask = c->mask; p.mask_ld = M; p.mask_col = q;
      tb.add_problem(p);
    }
    if ((rc = tb.flush())) return rc;
  }
  HeadLaunch hl;
  memset(&hl, 0, sizeof(hl));
  hl.train = 0; hl.store_acts = 0;
  hl.conf = conf; hl.pred = reinterpret_cast<long long*>(pred);
  return launch_head(L, c, ws, A, hl, st, "HEAD gating+classifier+softmax");
}

bool fusion_bf16_head_fused(const Layout& L) { return use_head(L); }

}  // namespace msf
