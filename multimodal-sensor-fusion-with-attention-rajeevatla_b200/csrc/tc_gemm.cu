// Persistent warp-specialised grouped GEMM: TMA -> 128B-swizzled smem ->
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue.
//
//   warp 0 (one lane)  TMA producer      : fills a 4-stage ring, arms full[stage] with expect_tx
//   warp 1 (one lane)  MMA issuer        : tcgen05.mma per 16-deep K slice, tcgen05.commit frees the stage
//   warp 2             TMEM allocator    : 2 x block_n columns (double-buffered accumulator)
//   warps 4..7         epilogue          : one thread per accumulator lane (= output row)
//
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps
// the MMAs of tile i+1.  Every mbarrier wait is bounded: a wedged pipeline traps
// instead of hanging the GPU.
#include "tc_gemm.cuh"

#include <stdlib.h>

#include "tc_ptx.cuh"
#include "wg2_gemm.cuh"

namespace msf {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

constexpr uint32_t A_STAGE_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;  // 16 KiB
__host__ __device__ constexpr uint32_t b_stage_bytes(int block_n) { return (uint32_t)block_n * TC_BLOCK_K * 2; }

struct TileCoord {
  int problem, m0, n0;
};

__device__ __forceinline__ TileCoord locate(const TcLaunch& L, int tile) {
  int pi = 0;
  while (pi + 1 < L.count && tile >= L.p[pi + 1].tile_begin) ++pi;
  const TcProblem& P = L.p[pi];
  const int local = tile - P.tile_begin;
  const int tiles_n = (P.N + L.block_n - 1) / L.block_n;
  TileCoord t;
  t.problem = pi;
  t.m0 = (local / tiles_n) * TC_BLOCK_M;
  t.n0 = (local % tiles_n) * L.block_n;
  return t;
}

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

// ---------------------------------------------------------------------------
// epilogue
// ---------------------------------------------------------------------------
// Everything an epilogue thread needs about its tile, copied out of the kernel
// parameters once per tile so the per-element code touches registers only.
struct EpiTile {
  int epi, M, N, nseg, c_bf16, head_dim, heads, site, sub, dbg;
  float scale;
  void* C;
  long long ldc;
  const __nv_bfloat16* aux;
  long long ld_aux;
  const __nv_bfloat16* aux2;
  long long ld_aux2;
  float* gate_out;
  const float* gate_in;
  float* cell;
  float* h32;
  long long h_slice;
  double* sq;
};

// 16 consecutive bf16 of one row, raw (two 16-byte loads) or zero when out of range.
struct Raw16 {
  uint4 lo, hi;
};
__device__ __forceinline__ Raw16 load_raw16(const __nv_bfloat16* row_ptr, int col, int n_cols, bool row_ok, bool vec_ok) {
  Raw16 r;
  r.lo = make_uint4(0u, 0u, 0u, 0u);
  r.hi = r.lo;
  if (!row_ok || col >= n_cols) return r;
  const __nv_bfloat16* src = row_ptr + col;
  if (vec_ok && col + 16 <= n_cols) {
    r.lo = __ldg(reinterpret_cast<const uint4*>(src));
    r.hi = __ldg(reinterpret_cast<const uint4*>(src) + 1);
  } else {  // ragged edge or unaligned rows
    unsigned short h[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = (col + j < n_cols) ? __bfloat16_as_ushort(src[j]) : (unsigned short)0;
    r.lo = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
    r.hi = make_uint4(h[8] | (h[9] << 16), h[10] | (h[11] << 16), h[12] | (h[13] << 16), h[14] | (h[15] << 16));
  }
  return r;
}
__device__ __forceinline__ void unpack16(const Raw16& r, float (&out)[16]) {
  const uint32_t w[8] = {r.lo.x, r.lo.y, r.lo.z, r.lo.w, r.hi.x, r.hi.y, r.hi.z, r.hi.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    out[2 * e] = __uint_as_float(w[e] << 16);
    out[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}

// Gate of one (row, column) when heads are narrower than a 16-column group.
__device__ __noinline__ float column_gate(bool value_gate, const DropCfg& drop, int sub, int head_dim, int heads,
                                          float* gate_out, const float* gate_in, int row, bool row_ok, int col,
                                          bool col_ok, float mrow) {
  const int head = col / head_dim;
  if (!value_gate) return (row_ok && head < heads) ? __ldg(gate_in + (long long)row * heads + head) : 0.0f;
  float gate = (mrow != 0.0f) ? 1.0f : 0.0f;
  if (drop.active && head < heads) gate *= drop1(drop, SITE_ATTN, sub, row, head);
  if (col == head * head_dim && col_ok && gate_out != nullptr && row_ok && head < heads)
    gate_out[(long long)row * heads + head] = gate;
  return gate;
}

// LSTM cell on a tile whose columns are gate-interleaved pre-activations (nn.LSTM gate order i, f, g, o;
// src/encoders.py:54-65,135-166): 16 accumulator columns = 4 hidden units.  sigmoid / tanh through __expf
// (abs error ~1e-6, far inside the bf16 path's tolerance).
// (MUFU.EX2 + MUFU.RCP: no IEEE-division slow path, the epilogue has to stay small enough for the I-cache)
__device__ __forceinline__ float lstm_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float lstm_tanh(float x) { return __fdividef(2.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }

__device__ __forceinline__ void lstm_tile(const EpiTile T, const float* bias_s, uint32_t tmem_acc, uint32_t tfull,
                                          uint32_t tfull_parity, int row, int n0, int ncols, int col_begin,
                                          int col_end) {
  const bool row_ok = row < T.M;
  const int hidden = T.N >> 2;
  const int my_end = min(ncols, col_end);
  // rolled loop (an unrolled body does not fit the instruction cache: ncu stall_no_inst); the previous cell
  // state of the NEXT group is fetched before the current group is processed
  float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row_ok && col_begin < my_end) nxt = *reinterpret_cast<const float4*>(T.cell + (long long)row * hidden + ((n0 + col_begin) >> 2));
  mbar_wait(tfull, tfull_parity);
  tc_fence_after();
#pragma unroll 1
  for (int c = col_begin; c < my_end; c += 16) {
    uint32_t acc[16];
    tmem_ld16_issue(tmem_acc + (uint32_t)c, acc);
    const float cp[4] = {nxt.x, nxt.y, nxt.z, nxt.w};
    if (row_ok && c + 16 < my_end) nxt = *reinterpret_cast<const float4*>(T.cell + (long long)row * hidden + ((n0 + c + 16) >> 2));
    tmem_wait16(acc);
    const int u0 = (n0 + c) >> 2;
    float cn[4], hn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c + 4 * j);
      const float gi = lstm_sigmoid(__uint_as_float(acc[4 * j + 0]) + b4.x);
      const float gf = lstm_sigmoid(__uint_as_float(acc[4 * j + 1]) + b4.y);
      const float gg = lstm_tanh(__uint_as_float(acc[4 * j + 2]) + b4.z);
      const float go = lstm_sigmoid(__uint_as_float(acc[4 * j + 3]) + b4.w);
      cn[j] = gf * cp[j] + gi * gg;
      hn[j] = go * lstm_tanh(cn[j]);
    }
    if (row_ok) {
      *reinterpret_cast<float4*>(T.cell + (long long)row * hidden + u0) = make_float4(cn[0], cn[1], cn[2], cn[3]);
      uint2 pk;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
      h2[0] = __floats2bfloat162_rn(hn[0], hn[1]);
      h2[1] = __floats2bfloat162_rn(hn[2], hn[3]);
      __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(T.C) + (long long)(u0 >> 6) * T.h_slice + (long long)row * 64 + (u0 & 63);
      *reinterpret_cast<uint2*>(hb) = pk;
      if (T.h32 != nullptr)
        *reinterpret_cast<float4*>(T.h32 + (long long)row * hidden + u0) = make_float4(hn[0], hn[1], hn[2], hn[3]);
    }
  }
}

// Everything a warp does for one output tile: its 16-column groups of one
// 32-row slab.  Kept as a rolled loop (an unrolled, branchy epilogue thrashed
// the instruction cache: ncu stall_no_inst), all per-tile state in registers,
// and the auxiliary rows of the NEXT group are fetched before the current
// group is processed so their L2 latency overlaps the arithmetic.
template <int EPI>
__device__ __forceinline__ void epilogue_tile(const EpiTile T, const float* bias_s, const DropCfg& drop,
                                              uint32_t tmem_acc, uint32_t tfull, uint32_t tfull_parity, int row,
                                              float mrow, int n0, int ncols, int col_begin, int col_end, bool has_acc,
                                              unsigned char* stage, int stage_pitch, int row0, int lane) {
  constexpr bool kBias = EPI == TC_EPI_STORE || EPI == TC_EPI_BIAS_RELU_DROP || EPI == TC_EPI_VALUE_GATE ||
                         EPI == TC_EPI_OUT_MEAN;
  constexpr bool kAux = EPI == TC_EPI_OUT_MEAN || EPI == TC_EPI_RELU_GRAD || EPI == TC_EPI_ADD_RELU_GRAD;
  constexpr bool kAux2 = EPI == TC_EPI_ADD_RELU_GRAD;
  constexpr bool kGate = EPI == TC_EPI_VALUE_GATE || EPI == TC_EPI_GATE_MUL;
  constexpr bool kDrop = EPI == TC_EPI_BIAS_RELU_DROP || EPI == TC_EPI_DX;

  const bool row_ok = row < T.M;
  const __nv_bfloat16* aux_row = kAux ? T.aux + (long long)row * T.ld_aux : nullptr;
  const __nv_bfloat16* aux2_row = kAux2 ? T.aux2 + (long long)row * T.ld_aux2 : nullptr;
  const bool aux_vec = kAux && ((reinterpret_cast<uintptr_t>(T.aux) & 15) == 0) && ((T.ld_aux & 7) == 0);
  const bool aux2_vec = kAux2 && ((reinterpret_cast<uintptr_t>(T.aux2) & 15) == 0) && ((T.ld_aux2 & 7) == 0);
  const bool c_vec = ((reinterpret_cast<uintptr_t>(T.C) & 15) == 0) && ((T.ldc & (T.c_bf16 ? 7 : 3)) == 0);
  const float inv_scale = 1.0f / T.scale;
  (void)inv_scale;

  // this warp owns tile columns [col_begin, col_end) (a contiguous range, so its rows leave as long runs);
  // group g covers 16 of them
  const int my_end = min(ncols, col_end);
  auto group_col = [&](int g) { return col_begin + g * 16; };
  const int esize = T.c_bf16 ? 2 : 4;
  unsigned char* my_row = stage + lane * stage_pitch;

  // everything that can be fetched before the accumulator is ready is fetched now
  Raw16 nxt_aux, nxt_aux2;
  if (kAux) nxt_aux = load_raw16(aux_row, n0 + group_col(0), T.N, row_ok && group_col(0) < my_end, aux_vec);
  if (kAux2) nxt_aux2 = load_raw16(aux2_row, n0 + group_col(0), T.N, row_ok && group_col(0) < my_end, aux2_vec);

  mbar_wait(tfull, tfull_parity);
  tc_fence_after();

  int cur_head = -1;
  float cur_gate = 0.0f;
  float sq = 0.0f;   // TC_EPI_STORE with T.sq: sum of squares of this thread's stored values
#pragma unroll 1
  for (int g = 0;; ++g) {
    const int c = group_col(g);
    if (c >= my_end) break;
    const int col0 = n0 + c;
    const int gcols = min(16, T.N - col0);
    const int cn = group_col(g + 1);

    uint32_t acc[16];
    if (has_acc) {
      if (!(T.dbg & 2)) {
        tmem_ld16_issue(tmem_acc + (uint32_t)c, acc);
        tmem_wait16(acc);
      } else {
        for (int j = 0; j < 16; ++j) acc[j] = 0u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0u;  // epilogue-only problem: nothing was accumulated
    }

    float aux[16], aux2[16];
    if (kAux) {
      unpack16(nxt_aux, aux);
      nxt_aux = load_raw16(aux_row, n0 + cn, T.N, row_ok && cn < my_end, aux_vec);
    }
    if (kAux2) {
      unpack16(nxt_aux2, aux2);
      nxt_aux2 = load_raw16(aux2_row, n0 + cn, T.N, row_ok && cn < my_end, aux2_vec);
    }

    // per-(row, head) gate; the group lies inside one head when head_dim is a multiple of 16
    const bool one_head = kGate && (T.head_dim & 15) == 0;
    if (kGate && one_head) {
      const int head = col0 / T.head_dim;
      if (head != cur_head) {
        cur_head = head;
        if (EPI == TC_EPI_VALUE_GATE) {
          cur_gate = (mrow != 0.0f) ? 1.0f : 0.0f;
          if (drop.active && head < T.heads) cur_gate *= drop1(drop, SITE_ATTN, T.sub, row, head);
          if (col0 == head * T.head_dim && T.gate_out != nullptr && row_ok && head < T.heads)
            T.gate_out[(long long)row * T.heads + head] = cur_gate;
        } else {
          cur_gate = (row_ok && head < T.heads) ? __ldg(T.gate_in + (long long)row * T.heads + head) : 0.0f;
        }
      }
    }

    float dm[16];
    if (kDrop) {
#pragma unroll
      for (int j = 0; j < 16; ++j) dm[j] = 1.0f;
      if (drop.active && row_ok) {
        float d8[8];
        drop8(drop, T.site, T.sub, row, col0 >> 3, d8);
#pragma unroll
        for (int j = 0; j < 8; ++j) dm[j] = d8[j];
        if (gcols > 8) {
          drop8(drop, T.site, T.sub, row, (col0 >> 3) + 1, d8);
#pragma unroll
          for (int j = 0; j < 8; ++j) dm[8 + j] = d8[j];
        }
      }
    }

    float bias[16];
    if (kBias) {  // broadcast reads of the tile's summed bias row staged in shared memory
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c + 4 * q);
        bias[4 * q] = b4.x; bias[4 * q + 1] = b4.y; bias[4 * q + 2] = b4.z; bias[4 * q + 3] = b4.w;
      }
    }

    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = __uint_as_float(acc[j]);
      const float b = kBias ? bias[j] : 0.0f;
      float gate = cur_gate;
      if (kGate && !one_head)  // narrow / unaligned heads: per-column gate (rare shapes, out of line)
        gate = column_gate(EPI == TC_EPI_VALUE_GATE, drop, T.sub, T.head_dim, T.heads, T.gate_out, T.gate_in, row,
                           row_ok, col0 + j, j < gcols, mrow);
      float r;
      if (EPI == TC_EPI_STORE) r = fmaf(a, T.scale, b);
      else if (EPI == TC_EPI_BIAS_RELU_DROP) r = fmaxf(a + b, 0.0f) * dm[j];
      else if (EPI == TC_EPI_VALUE_GATE) r = (a + b) * gate;
      else if (EPI == TC_EPI_OUT_MEAN) r = (a + b + aux[j]) * inv_scale * mrow;
      else if (EPI == TC_EPI_RELU_GRAD) r = a * (aux[j] > 0.0f ? T.scale : 0.0f);
      else if (EPI == TC_EPI_GATE_MUL) r = a * gate;
      else if (EPI == TC_EPI_ADD_RELU_GRAD) r = (a + aux[j]) * (aux2[j] > 0.0f ? T.scale : 0.0f);
      else r = a * mrow * dm[j];  // TC_EPI_DX
      v[j] = r;
    }
    if (EPI == TC_EPI_STORE && T.sq != nullptr && row_ok) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < gcols) sq = fmaf(v[j], v[j], sq);
    }

    // results go to this warp's shared-memory staging rows first (row pitch padded by 16 B: conflict-free)
    {
      unsigned char* dst = my_row + (c - col_begin) * esize;
      if (T.c_bf16) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint4 pk;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
          reinterpret_cast<uint4*>(dst)[q] = pk;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
  }

  if (EPI == TC_EPI_STORE && T.sq != nullptr) {   // one fp64 atomic per warp and tile
    double s = (double)sq;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && s != 0.0) atomicAdd(T.sq, s);
  }

  // flush: the warp walks its 32 x (col_end - col_begin) block row by row, 16 bytes per lane, so every
  // store instruction writes long contiguous runs (per-thread row stores were request-rate bound)
  __syncwarp();
  if (!(T.dbg & 1) && col_begin < my_end) {
    const int epc = 16 / esize;                               // elements per 16-byte chunk
    const int nch = ((my_end - col_begin) * esize + 15) >> 4;   // chunks per row that hold data
    const int total = 32 * nch;
#pragma unroll 1
    for (int f = lane; f < total; f += 32) {
      const int r = f / nch, ch = f - r * nch;
      const int grow = row0 + r;
      if (grow >= T.M) continue;
      const uint4 val = *reinterpret_cast<const uint4*>(stage + r * stage_pitch + ch * 16);
      const int gcol = n0 + col_begin + ch * epc;
      unsigned char* gdst = reinterpret_cast<unsigned char*>(T.C) + ((long long)grow * T.ldc + gcol) * esize;
      if (c_vec && gcol + epc <= T.N) {
        *reinterpret_cast<uint4*>(gdst) = val;
      } else if (T.c_bf16) {
        const unsigned short* hv = reinterpret_cast<const unsigned short*>(&val);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (gcol + e < T.N) reinterpret_cast<unsigned short*>(gdst)[e] = hv[e];
      } else {
        const float* fv = reinterpret_cast<const float*>(&val);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (gcol + e < T.N) reinterpret_cast<float*>(gdst)[e] = fv[e];
      }
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
template <bool MN_MAJOR>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcLaunch L) {
  TL_KERNEL(MN_MAJOR ? 1 : 0);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [A stages][B stages][barriers][tmem base]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t BN = (uint32_t)L.block_n;
  const uint32_t B_STAGE = b_stage_bytes(L.block_n);
  const int STAGES = L.stages;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + STAGES * A_STAGE_BYTES;
  const uint32_t stg_base = b_base + STAGES * B_STAGE;          // epilogue staging, TC_EPI_WARPS x 32 rows
  const uint32_t bar_base = stg_base + (uint32_t)(TC_EPI_WARPS * 32 * L.stage_pitch);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TC_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TC_MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC_MAX_STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  // two summed-bias rows (one per accumulator buffer), 16-byte aligned
  float* bias_smem = reinterpret_cast<float*>(smem_raw + (bar_base + 8u * (2 * TC_MAX_STAGES + 6) - smem_u32(smem_raw)));
  unsigned char* stage_smem = smem_raw + (stg_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < L.nmaps; ++i) tma_prefetch_desc(&L.maps[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), TC_EPI_WARPS);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();     TL_WAITED(MN_MAJOR ? 1 : 0);  // prologue overlapped the previous kernel (programmatic dependent launch, msf_common.cuh)
  pdl_launch();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < L.total_tiles; tile += gridDim.x) {
        const TileCoord t = locate(L, tile);
        const int nseg = L.p[t.problem].nseg;
        const int kb_per_seg = (L.p[t.problem].K + TC_BLOCK_K - 1) / TC_BLOCK_K;
        for (int s = 0; s < nseg; ++s) {
          const CUtensorMap* amap = &L.maps[L.p[t.problem].seg[s].a_map];
          const CUtensorMap* bmap = &L.maps[L.p[t.problem].seg[s].b_map];
          const int a_z = L.p[t.problem].seg[s].a_z, b_z = L.p[t.problem].seg[s].b_z;
          for (int kb = 0; kb < kb_per_seg; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), A_STAGE_BYTES + B_STAGE);
            const uint32_t a_dst = a_base + stage * A_STAGE_BYTES;
            const uint32_t b_dst = b_base + stage * B_STAGE;
            const int k0 = kb * TC_BLOCK_K;
            if (!MN_MAJOR) {
              tma_load_3d(a_dst, amap, k0, t.m0, a_z, full_bar(stage));
              tma_load_3d(b_dst, bmap, k0, t.n0, b_z, full_bar(stage));
            } else {
              // 64(mn) x 64(k) slabs, 8 KiB each, mn-slab major
              for (int j = 0; j < TC_BLOCK_M / 64; ++j)
                tma_load_3d(a_dst + j * 8192u, amap, t.m0 + 64 * j, k0, a_z, full_bar(stage));
              for (int j = 0; j < (int)BN / 64; ++j)
                tma_load_3d(b_dst + j * 8192u, bmap, t.n0 + 64 * j, k0, b_z, full_bar(stage));
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================
    if (lane == 0) {
      const uint32_t idesc = instr_desc((int)BN, MN_MAJOR, MN_MAJOR);
      const uint32_t lbo = MN_MAJOR ? 8192u : 16u;
      const uint32_t sbo = 1024u;
      const uint32_t kstep = MN_MAJOR ? 2048u : 32u;  // bytes per UMMA_K = 16
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < L.total_tiles; tile += gridDim.x, ++it) {
        const TileCoord t = locate(L, tile);
        const int total_kb = L.p[t.problem].nseg * ((L.p[t.problem].K + TC_BLOCK_K - 1) / TC_BLOCK_K);
        const int acc = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(tempty_bar(acc), (use & 1u) ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = a_base + stage * A_STAGE_BYTES;
          const uint32_t b_addr = b_base + stage * B_STAGE;
#pragma unroll
          for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
            const uint64_t da = smem_desc(a_addr + k * kstep, lbo, sbo);
            const uint64_t db = smem_desc(b_addr + k * kstep, lbo, sbo);
            tc_mma_bf16(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(empty_bar(stage));  // stage reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue ===============================
    const DropCfg drop = resolve_drop(L.drop);
    const int lq = warp & 3;               // TMEM lane quarter this warp may read
    const int colgroup = (warp - 4) >> 2;  // which contiguous column range of the tile it owns
    const int cols_per_warp = BN >= 64 ? (int)BN / TC_EPI_COLGROUPS : (int)BN;
    const int col_begin = colgroup * cols_per_warp;
    const int col_end = (BN >= 64 || colgroup == 0) ? col_begin + cols_per_warp : col_begin;
    unsigned char* stage = stage_smem + (warp - 4) * 32 * L.stage_pitch;
    const int et = threadIdx.x - 128;      // 0 .. 32*TC_EPI_WARPS-1
    int it = 0;
    for (int tile = blockIdx.x; tile < L.total_tiles; tile += gridDim.x, ++it) {
      const TileCoord t = locate(L, tile);
      const TcProblem& P = L.p[t.problem];
      EpiTile T;
      T.epi = P.epi; T.M = P.M; T.N = P.N; T.nseg = P.nseg; T.c_bf16 = P.c_bf16;
      T.head_dim = P.head_dim; T.heads = P.heads; T.site = P.site; T.sub = P.sub; T.scale = P.scale;
      T.C = P.C; T.ldc = P.ldc; T.aux = P.aux; T.ld_aux = P.ld_aux; T.aux2 = P.aux2; T.ld_aux2 = P.ld_aux2;
      T.gate_out = P.gate_out; T.gate_in = P.gate_in; T.dbg = L.dbg;
      T.cell = P.cell; T.h32 = P.h32; T.h_slice = P.h_slice; T.sq = P.sq;
      const bool has_acc = P.K > 0;
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int row = t.m0 + lq * 32 + lane;
      const float mrow = (P.mask != nullptr && row < T.M) ? __ldg(P.mask + (long long)row * P.mask_ld + P.mask_col) : 1.0f;
      const int ncols = min((int)BN, T.N - t.n0);
      // summed segment biases of this tile's columns -> shared memory (double-buffered by accumulator parity)
      float* bias_s = bias_smem + acc * 256;
      for (int e = et; e < (int)BN; e += 32 * TC_EPI_WARPS) {
        float bsum = 0.0f;
        const int col = t.n0 + e;
        if (col < T.N) {
#pragma unroll
          for (int s = 0; s < TC_MAX_SEG; ++s)
            if (s < T.nseg && P.bias[s] != nullptr) bsum += __ldg(P.bias[s] + col);
        }
        bias_s[e] = bsum;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");  // epilogue warps only
      const uint32_t tmem_acc = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)acc * BN;
      const uint32_t tf = tfull_bar(acc), tp = use & 1u;
      switch (T.epi) {
        default:
        case TC_EPI_STORE: epilogue_tile<TC_EPI_STORE>(T, bias_s, drop, tmem_acc, tf, tp, row, mrow, t.n0, ncols, col_begin, col_end, has_acc, stage, L.stage_pitch, t.m0 + lq * 32, lane); break;
        case TC_EPI_BIAS_RELU_DROP: epilogue_tile<TC_EPI_BIAS_RELU_DROP>(T, bias_s, drop, tmem_acc, tf, tp, row, mrow, t.n0, ncols, col_begin, col_end, has_acc, stage, L.stage_pitch, t.m0 + lq * 32, lane); break;
        case TC_EPI_VALUE_GATE: epilogue_tile<TC_EPI_VALUE_GATE>(T, bias_s, drop, tmem_acc, tf, tp, row, mrow, t.n0, ncols, col_begin, col_end, has_acc, stage, L.stage_pitch, t.m0 + lq * 32, lane); break;
        case TC_EPI_OUT_MEAN: epilogue_tile<TC_EPI_OUT_MEAN>(T, bias_s, drop, tmem_acc, tf, tp, row, mrow, t.n0, ncols, col_begin, col_end, has_acc, stage, L.stage_pitch, t.m0 + lq * 32, lane); break;
        case TC_EPI_RELU_GRAD: epilogue_tile<TC_EPI_RELU_GRAD>(T, bias_s, drop, tmem_acc, tf, tp, row, mrow, t.n0, ncols, col_begin, col_end, has_acc, stage, L.stage_pitch, t.m0 + lq * 32, lane); break;
        case TC_EPI_GATE_MUL: epilogue_tile<TC_EPI_GATE_MUL>(T, bias_s, drop, tmem_acc, tf, tp, row, mrow, t.n0, ncols, col_begin, col_end, has_acc, stage, L.stage_pitch, t.m0 + lq * 32, lane); break;
        case TC_EPI_ADD_RELU_GRAD: epilogue_tile<TC_EPI_ADD_RELU_GRAD>(T, bias_s, drop, tmem_acc, tf, tp, row, mrow, t.n0, ncols, col_begin, col_end, has_acc, stage, L.stage_pitch, t.m0 + lq * 32, lane); break;
        case TC_EPI_LSTM: lstm_tile(T, bias_s, tmem_acc, tf, tp, row, t.n0, ncols, col_begin, col_end); break;
        case TC_EPI_DX: epilogue_tile<TC_EPI_DX>(T, bias_s, drop, tmem_acc, tf, tp, row, mrow, t.n0, ncols, col_begin, col_end, has_acc, stage, L.stage_pitch, t.m0 + lq * 32, lane); break;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

constexpr size_t TC_SMEM_LIMIT = 232448;  // 227 KiB per CTA on sm_100

size_t tc_fixed_smem_bytes() { return 1024 + 8 * (2 * TC_MAX_STAGES + 6) + 2 * 256 * 4; }

}  // namespace

int tc_init() {
  if (g_encode != nullptr) return MSF_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  MSF_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return MSF_E_CUDA;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return MSF_OK;
}

int tc_encode_map(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld, long long depth,
                  long long slice, int box_cols, int box_rows) {
  int rc = tc_init();
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 8) || (depth > 1 && (slice % 8))) {
    set_error("tc_encode_map: operand not 16-byte aligned (base %p, ld %lld, slice %lld)", base, ld, slice);
    return MSF_E_INVALID;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(depth < 1 ? 1 : depth)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(depth > 1 ? slice : rows * ld) * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rows %lld cols %lld ld %lld depth %lld box %d x %d", (int)r, rows, cols,
              ld, depth, box_cols, box_rows);
    return MSF_E_CUDA;
  }
  return MSF_OK;
}

TcBuilder::TcBuilder(bool mn, int block_n, const DropCfg& drop, cudaStream_t st, const char* lbl)
    : nmaps(0), mn_major(mn), stream(st), status(MSF_OK), label(lbl) {
  memset(&L, 0, sizeof(L));
  L.block_n = block_n;
  L.drop = drop;
}

TcProblem tc_blank_problem() {
  TcProblem p;
  memset(&p, 0, sizeof(p));
  p.nseg = 1;
  p.scale = 1.0f;
  p.head_dim = 1;
  p.heads = 1;
  return p;
}

DropCfg no_dropout() {
  DropCfg d;
  memset(&d, 0, sizeof(d));
  d.scale = 1.0f;
  return d;
}

int TcBuilder::add_map(const void* base, long long rows, long long cols, long long ld, long long depth,
                       long long slice, int role_rows) {
  if (nmaps >= TC_MAX_MAPS) {
    set_error("tc_gemm: too many tensor maps in one launch");
    status = MSF_E_INVALID;
    return -1;
  }
  if (tc_init() != MSF_OK) {
    status = MSF_E_CUDA;
    return -1;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 8) || (depth > 1 && (slice % 8))) {
    set_error("tc_gemm: operand not 16-byte aligned (base %p, ld %lld, slice %lld)", base, ld, slice);
    status = MSF_E_INVALID;
    return -1;
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(depth < 1 ? 1 : depth)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(depth > 1 ? slice : rows * ld) * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)(mn_major ? 64 : role_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(&L.maps[nmaps], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rows %lld cols %lld ld %lld depth %lld box %u", (int)r, rows, cols,
              ld, depth, box[1]);
    status = MSF_E_CUDA;
    return -1;
  }
  return nmaps++;
}

int TcBuilder::add_problem(const TcProblem& p) {
  if (status != MSF_OK) return status;
  if (p.M <= 0 || p.N <= 0) return MSF_OK;
  if (L.count >= TC_MAX_PROBLEMS) {
    const int rc = flush();
    if (rc) return rc;
  }
  if (p.nseg < 1 || p.nseg > TC_MAX_SEG) {
    set_error("tc_gemm: nseg %d out of range", p.nseg);
    return status = MSF_E_INVALID;
  }
  for (int s = 0; s < p.nseg; ++s)
    if (p.seg[s].a_map < 0 || p.seg[s].b_map < 0) return status = MSF_E_INVALID;  // add_map failed earlier
  TcProblem q = p;
  q.tile_begin = L.total_tiles;
  L.total_tiles += (int)(ceil_div(p.M, TC_BLOCK_M) * ceil_div(p.N, L.block_n));
  L.p[L.count++] = q;
  return MSF_OK;
}

int TcBuilder::flush(bool keep) {
  if (status != MSF_OK) return status;
  if (L.total_tiles == 0) return MSF_OK;
  MSF_REQUIRE(L.block_n >= 32 && L.block_n <= 256 && L.block_n % 16 == 0 && (!mn_major || L.block_n % 64 == 0),
              "tc_gemm: block_n %d unsupported", L.block_n);
  // fill unused descriptor slots with a valid descriptor (they are prefetched)
  for (int i = nmaps; i < TC_MAX_MAPS; ++i) L.maps[i] = L.maps[0];
  L.nmaps = nmaps;
  { const char* e = getenv("MSF_TC_DEBUG"); L.dbg = e ? atoi(e) : 0; }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    MSF_CHECK_CUDA(cudaGetDevice(&dev));
    MSF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // epilogue staging: each epilogue warp stages 32 rows x its column range in the widest output type of the launch
  int esize = 2;
  for (int i = 0; i < L.count; ++i)
    if (!L.p[i].c_bf16) esize = 4;
  const int cols_per_warp = L.block_n >= 64 ? L.block_n / TC_EPI_COLGROUPS : L.block_n;
  L.stage_pitch = cols_per_warp * esize + 16;
  const size_t staging = (size_t)TC_EPI_WARPS * 32 * L.stage_pitch;
  const size_t per_stage = A_STAGE_BYTES + b_stage_bytes(L.block_n);
  int stages = (int)((TC_SMEM_LIMIT - tc_fixed_smem_bytes() - staging) / per_stage);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  MSF_REQUIRE(stages >= 2, "tc_gemm: block_n %d with fp32 output does not leave room for the operand pipeline", L.block_n);
  L.stages = stages;
  const size_t smem = tc_fixed_smem_bytes() + staging + (size_t)stages * per_stage;
  const int grid = L.total_tiles < sms ? L.total_tiles : sms;
  if (prof_enabled()) {
    double flops = 0.0;
    for (int i = 0; i < L.count; ++i) flops += 2.0 * L.p[i].M * L.p[i].N * (double)L.p[i].K * L.p[i].nseg;
    prof_begin(label, flops, stream);
  }
  if (mn_major) {
    MSF_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSF_CHECK_CUDA(launch_pdl(tc_gemm_kernel<true>, dim3(grid), dim3(TC_THREADS), smem, stream, L));
  } else {
    MSF_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSF_CHECK_CUDA(launch_pdl(tc_gemm_kernel<false>, dim3(grid), dim3(TC_THREADS), smem, stream, L));
  }
  MSF_LAUNCH_CHECK();
  prof_end(stream);
  if (!keep) {
    L.count = 0;
    L.total_tiles = 0;
  }
  return MSF_OK;
}

}  // namespace msf

// ---------------------------------------------------------------------------
// stand-alone entry point (tests, encoder/attention projections)
// ---------------------------------------------------------------------------
extern "C" int msf_gemm_bf16(const void* a, const void* b, void* d, int32_t d_is_bf16, int64_t m, int64_t n,
                             int64_t k, int64_t lda, int64_t ldb, int64_t ldd, int32_t mn_major,
                             const float* bias, int32_t relu, void* stream) {
  MSF_REQUIRE(a && b && d && m >= 1 && n >= 1 && k >= 1, "msf_gemm_bf16: bad arguments");
  MSF_REQUIRE(m < (1ll << 31) && n < (1ll << 31) && k < (1ll << 31), "msf_gemm_bf16: dimension too large");
  if (mn_major && !d_is_bf16 && bias == nullptr && !relu && msf::wg2_enabled() && msf::wg2_shape_ok((int)n, ldd, d)) {
    // weight-gradient shape: CTA pairs, 256 x 256 tiles, contraction split in two halves (wg2_gemm.cu)
    msf::Wg2Builder wb(k, (cudaStream_t)stream, "wg2_gemm");
    const short am = (short)wb.add_map(a, k, m, lda, 1, 0), bm = (short)wb.add_map(b, k, n, ldb, 1, 0);
    if (am < 0 || bm < 0) return wb.status;
    int rc = wb.add_problem(am, 0, bm, 0, (int)m, (int)n, static_cast<float*>(d), ldd, nullptr);
    if (rc) return rc;
    return wb.flush();
  }
  int bn = n > 128 ? 256 : n > 64 ? 128 : 64;
  if (!d_is_bf16 && bn > 128) bn = 128;  // fp32 tiles are staged in shared memory: 128 x 128 x 4 B
  if (!mn_major && n <= 32) bn = 32;
  msf::TcBuilder tb(mn_major != 0, bn, msf::no_dropout(), (cudaStream_t)stream);
  int am, bm;
  if (!mn_major) {
    am = tb.add_map(a, m, k, lda, 1, 0, msf::TC_BLOCK_M);   // A[m, k]
    bm = tb.add_map(b, n, k, ldb, 1, 0, bn);                // B[n, k]
  } else {
    am = tb.add_map(a, k, m, lda, 1, 0, 0);                 // A[k, m]  (contraction over rows)
    bm = tb.add_map(b, k, n, ldb, 1, 0, 0);                 // B[k, n]
  }
  if (am < 0 || bm < 0) return MSF_E_INVALID;
  msf::TcProblem p = msf::tc_blank_problem();
  p.seg[0].a_map = (short)am; p.seg[0].b_map = (short)bm;
  p.bias[0] = bias;
  p.M = (int)m; p.N = (int)n; p.K = (int)k;
  p.C = d; p.ldc = ldd; p.c_bf16 = d_is_bf16;
  p.epi = relu ? msf::TC_EPI_BIAS_RELU_DROP : msf::TC_EPI_STORE;
  int rc = tb.add_problem(p);
  if (rc) return rc;
  return tb.flush();
}
