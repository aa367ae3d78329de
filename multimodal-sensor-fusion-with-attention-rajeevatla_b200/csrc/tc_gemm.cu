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

namespace msf {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: ~2 s at 2 GHz, then trap (a CUDA error beats a hung GPU).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("msf_b200 tc_gemm: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------
// descriptors
// ---------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version bit set
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // layout type: SWIZZLE_128B
  return d;
}

// instruction descriptor: bf16 x bf16 -> fp32, M = 128
__device__ __forceinline__ uint32_t instr_desc(int n, bool a_mn, bool b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                       // C format: F32
  d |= 1u << 7;                       // A format: BF16
  d |= 1u << 10;                      // B format: BF16
  d |= (a_mn ? 1u : 0u) << 15;        // A major: 0 = K, 1 = MN
  d |= (b_mn ? 1u : 0u) << 16;        // B major
  d |= (uint32_t)(n >> 3) << 17;      // N / 8
  d |= (uint32_t)(TC_BLOCK_M >> 4) << 24;  // M / 16
  return d;
}

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
  int epi, M, N, nseg, c_bf16, head_dim, heads, site, sub;
  float scale;
  void* C;
  long long ldc;
  const __nv_bfloat16* aux;
  long long ld_aux;
  const __nv_bfloat16* aux2;
  long long ld_aux2;
  float* gate_out;
  const float* gate_in;
};

__device__ __forceinline__ void load_row32(const __nv_bfloat16* src, bool vec, int ncols, float (&out)[32]) {
  if (vec) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src) + q);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        out[q * 8 + 2 * e] = f.x;
        out[q * 8 + 2 * e + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) out[j] = (j < ncols) ? bf2f(src[j]) : 0.0f;
  }
}

// One chunk: 32 consecutive columns [col0, col0+32) of output row `row`.
// bias_lane: sum of the segment biases for column col0 + lane (0 beyond N).
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const EpiTile& T, const DropCfg& drop, int row, bool row_ok, int col0,
                                               const uint32_t (&acc)[32], float bias_lane, float mrow) {
  const int ncols = min(32, T.N - col0);
  const bool full = ncols == 32;
  float v[32];

  float aux[32], aux2[32];
  if (EPI == TC_EPI_OUT_MEAN || EPI == TC_EPI_RELU_GRAD || EPI == TC_EPI_ADD_RELU_GRAD) {
    if (row_ok) {
      const __nv_bfloat16* src = T.aux + (long long)row * T.ld_aux + col0;
      load_row32(src, full && ((reinterpret_cast<uintptr_t>(src) & 15) == 0), ncols, aux);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) aux[j] = 0.0f;
    }
  }
  if (EPI == TC_EPI_ADD_RELU_GRAD) {
    if (row_ok) {
      const __nv_bfloat16* src = T.aux2 + (long long)row * T.ld_aux2 + col0;
      load_row32(src, full && ((reinterpret_cast<uintptr_t>(src) & 15) == 0), ncols, aux2);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) aux2[j] = 0.0f;
    }
  }

  // per-(row, head) gate: heads are runs of head_dim columns
  int head = 0, rem = 0;
  float gate = 0.0f;
  if (EPI == TC_EPI_VALUE_GATE || EPI == TC_EPI_GATE_MUL) {
    head = col0 / T.head_dim;
    rem = col0 - head * T.head_dim;
    if (rem != 0) {  // chunk starts inside a head: fetch its gate (already recorded by the owner of its first column)
      if (EPI == TC_EPI_VALUE_GATE) {
        gate = (mrow != 0.0f) ? 1.0f : 0.0f;
        if (drop.active && head < T.heads) gate *= drop1(drop, SITE_ATTN, T.sub, row, head);
      } else {
        gate = (row_ok && head < T.heads) ? __ldg(T.gate_in + (long long)row * T.heads + head) : 0.0f;
      }
    }
  }

#pragma unroll
  for (int g4 = 0; g4 < 8; ++g4) {
    float dm[4] = {1.f, 1.f, 1.f, 1.f};
    if (EPI == TC_EPI_BIAS_RELU_DROP || EPI == TC_EPI_DX) {
      if (drop.active && row_ok && g4 * 4 < ncols) drop4(drop, T.site, T.sub, row, (col0 >> 2) + g4, dm);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = g4 * 4 + e;
      const float a = __uint_as_float(acc[j]);
      float bias = 0.0f;
      if (EPI == TC_EPI_STORE || EPI == TC_EPI_BIAS_RELU_DROP || EPI == TC_EPI_VALUE_GATE || EPI == TC_EPI_OUT_MEAN)
        bias = __shfl_sync(0xffffffffu, bias_lane, j);
      if (EPI == TC_EPI_VALUE_GATE || EPI == TC_EPI_GATE_MUL) {
        if (rem == 0) {  // first column of a head
          if (EPI == TC_EPI_VALUE_GATE) {
            gate = (mrow != 0.0f) ? 1.0f : 0.0f;
            if (drop.active && head < T.heads) gate *= drop1(drop, SITE_ATTN, T.sub, row, head);
            if (T.gate_out != nullptr && row_ok && head < T.heads && j < ncols)
              T.gate_out[(long long)row * T.heads + head] = gate;
          } else {
            gate = (row_ok && head < T.heads) ? __ldg(T.gate_in + (long long)row * T.heads + head) : 0.0f;
          }
        }
        if (++rem == T.head_dim) {
          rem = 0;
          ++head;
        }
      }
      float r;
      if (EPI == TC_EPI_STORE) r = fmaf(a, T.scale, bias);
      else if (EPI == TC_EPI_BIAS_RELU_DROP) r = fmaxf(a + bias, 0.0f) * dm[e];
      else if (EPI == TC_EPI_VALUE_GATE) r = (a + bias) * gate;
      else if (EPI == TC_EPI_OUT_MEAN) r = (a + bias + aux[j]) / T.scale * mrow;
      else if (EPI == TC_EPI_RELU_GRAD) r = a * (aux[j] > 0.0f ? T.scale : 0.0f);
      else if (EPI == TC_EPI_GATE_MUL) r = a * gate;
      else if (EPI == TC_EPI_ADD_RELU_GRAD) r = (a + aux[j]) * (aux2[j] > 0.0f ? T.scale : 0.0f);
      else r = a * mrow * dm[e];  // TC_EPI_DX
      v[j] = r;
    }
  }

  if (!row_ok) return;
  if (T.c_bf16) {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(T.C) + (long long)row * T.ldc + col0;
    if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 pk;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
        reinterpret_cast<uint4*>(dst)[q] = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) dst[j] = __float2bfloat16_rn(v[j]);
    }
  } else {
    float* dst = reinterpret_cast<float*>(T.C) + (long long)row * T.ldc + col0;
    if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) dst[j] = v[j];
    }
  }
}

// All chunks this warp owns in one tile: chunk c (32 columns) belongs to column group c % TC_EPI_COLGROUPS.
template <int EPI>
__device__ __forceinline__ void epilogue_tile(const EpiTile& T, const float* const (&bias)[TC_MAX_SEG], const DropCfg& drop,
                                              uint32_t tmem_acc, int row, float mrow, int n0, int ncols, int colgroup,
                                              bool has_acc, int lane) {
  const bool row_ok = row < T.M;
  for (int c = colgroup * 32; c < ncols; c += 32 * TC_EPI_COLGROUPS) {
    uint32_t acc[32];
    if (has_acc) {
      tmem_ld32(tmem_acc + (uint32_t)c, acc);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0u;  // epilogue-only problem: nothing was accumulated
    }
    float bias_lane = 0.0f;
    if (EPI == TC_EPI_STORE || EPI == TC_EPI_BIAS_RELU_DROP || EPI == TC_EPI_VALUE_GATE || EPI == TC_EPI_OUT_MEAN) {
      const int col = n0 + c + lane;
      if (col < T.N) {
#pragma unroll
        for (int s = 0; s < TC_MAX_SEG; ++s)
          if (s < T.nseg && bias[s] != nullptr) bias_lane += __ldg(bias[s] + col);
      }
    }
    epilogue_chunk<EPI>(T, drop, row, row_ok, n0 + c, acc, bias_lane, mrow);
  }
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
template <bool MN_MAJOR>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcLaunch L) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [A stages][B stages][barriers][tmem base]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t BN = (uint32_t)L.block_n;
  const uint32_t B_STAGE = b_stage_bytes(L.block_n);
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + TC_STAGES * A_STAGE_BYTES;
  const uint32_t bar_base = b_base + TC_STAGES * B_STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TC_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TC_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC_STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < L.nmaps; ++i) tma_prefetch_desc(&L.maps[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
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
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
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
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue ===============================
    const DropCfg drop = resolve_drop(L.drop);
    const int lq = warp & 3;               // TMEM lane quarter this warp may read
    const int colgroup = (warp - 4) >> 2;  // which 32-column chunks it owns
    int it = 0;
    for (int tile = blockIdx.x; tile < L.total_tiles; tile += gridDim.x, ++it) {
      const TileCoord t = locate(L, tile);
      const TcProblem& P = L.p[t.problem];
      EpiTile T;
      T.epi = P.epi; T.M = P.M; T.N = P.N; T.nseg = P.nseg; T.c_bf16 = P.c_bf16;
      T.head_dim = P.head_dim; T.heads = P.heads; T.site = P.site; T.sub = P.sub; T.scale = P.scale;
      T.C = P.C; T.ldc = P.ldc; T.aux = P.aux; T.ld_aux = P.ld_aux; T.aux2 = P.aux2; T.ld_aux2 = P.ld_aux2;
      T.gate_out = P.gate_out; T.gate_in = P.gate_in;
      const float* bias[TC_MAX_SEG];
#pragma unroll
      for (int s = 0; s < TC_MAX_SEG; ++s) bias[s] = P.bias[s];
      const bool has_acc = P.K > 0;
      const int row = t.m0 + lq * 32 + lane;
      const float mrow = (P.mask != nullptr && row < T.M) ? __ldg(P.mask + (long long)row * P.mask_ld + P.mask_col) : 1.0f;
      const int ncols = min((int)BN, T.N - t.n0);
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(tfull_bar(acc), use & 1u);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)acc * BN;
      switch (T.epi) {
        default:
        case TC_EPI_STORE: epilogue_tile<TC_EPI_STORE>(T, bias, drop, tmem_acc, row, mrow, t.n0, ncols, colgroup, has_acc, lane); break;
        case TC_EPI_BIAS_RELU_DROP: epilogue_tile<TC_EPI_BIAS_RELU_DROP>(T, bias, drop, tmem_acc, row, mrow, t.n0, ncols, colgroup, has_acc, lane); break;
        case TC_EPI_VALUE_GATE: epilogue_tile<TC_EPI_VALUE_GATE>(T, bias, drop, tmem_acc, row, mrow, t.n0, ncols, colgroup, has_acc, lane); break;
        case TC_EPI_OUT_MEAN: epilogue_tile<TC_EPI_OUT_MEAN>(T, bias, drop, tmem_acc, row, mrow, t.n0, ncols, colgroup, has_acc, lane); break;
        case TC_EPI_RELU_GRAD: epilogue_tile<TC_EPI_RELU_GRAD>(T, bias, drop, tmem_acc, row, mrow, t.n0, ncols, colgroup, has_acc, lane); break;
        case TC_EPI_GATE_MUL: epilogue_tile<TC_EPI_GATE_MUL>(T, bias, drop, tmem_acc, row, mrow, t.n0, ncols, colgroup, has_acc, lane); break;
        case TC_EPI_ADD_RELU_GRAD: epilogue_tile<TC_EPI_ADD_RELU_GRAD>(T, bias, drop, tmem_acc, row, mrow, t.n0, ncols, colgroup, has_acc, lane); break;
        case TC_EPI_DX: epilogue_tile<TC_EPI_DX>(T, bias, drop, tmem_acc, row, mrow, t.n0, ncols, colgroup, has_acc, lane); break;
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

size_t tc_smem_bytes(int block_n) {
  return 1024 + (size_t)TC_STAGES * (A_STAGE_BYTES + b_stage_bytes(block_n)) + 8 * (2 * TC_STAGES + 4) + 16;
}

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

TcBuilder::TcBuilder(bool mn, int block_n, const DropCfg& drop, cudaStream_t st)
    : nmaps(0), mn_major(mn), stream(st), status(MSF_OK) {
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

int TcBuilder::flush() {
  if (status != MSF_OK) return status;
  if (L.total_tiles == 0) return MSF_OK;
  MSF_REQUIRE(L.block_n >= 32 && L.block_n <= 256 && L.block_n % 16 == 0 && (!mn_major || L.block_n % 64 == 0),
              "tc_gemm: block_n %d unsupported", L.block_n);
  // fill unused descriptor slots with a valid descriptor (they are prefetched)
  for (int i = nmaps; i < TC_MAX_MAPS; ++i) L.maps[i] = L.maps[0];
  L.nmaps = nmaps;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    MSF_CHECK_CUDA(cudaGetDevice(&dev));
    MSF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const size_t smem = tc_smem_bytes(L.block_n);
  const int grid = L.total_tiles < sms ? L.total_tiles : sms;
  if (mn_major) {
    MSF_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_kernel<true><<<grid, TC_THREADS, smem, stream>>>(L);
  } else {
    MSF_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_kernel<false><<<grid, TC_THREADS, smem, stream>>>(L);
  }
  MSF_LAUNCH_CHECK();
  L.count = 0;
  L.total_tiles = 0;
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
  int bn = n >= 256 ? 256 : n > 128 ? 256 : n > 64 ? 128 : 64;
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
