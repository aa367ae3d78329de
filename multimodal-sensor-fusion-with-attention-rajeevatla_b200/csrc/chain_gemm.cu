// chain_kernel: see chain_gemm.cuh.  One persistent CTA per SM, 12 warps:
//
//   warp 0 (one lane)  TMA producer : A1 / W1 k-blocks for GEMM1, then W2 k-blocks for GEMM2, through a smem ring
//   warp 1 (one lane)  MMA issuer   : GEMM1 -> TMEM cols [0,H), GEMM2 (A = the staged intermediate) -> TMEM cols [H,2H)
//   warp 2             TMEM allocator
//   warp 3 (one lane)  TMA store    : intermediate / final tiles from the swizzled smem block to global memory
//   warps 4..11        epilogue     : thread = accumulator row; tcgen05.ld, math, bf16, 128B-swizzled st.shared
//
// The smem block the epilogue writes is at once the K-major A operand of GEMM2 (read by tcgen05.mma
// through a shared-memory descriptor) and the source box of the TMA store, so the intermediate is
// produced exactly once.  All mbarrier waits are bounded (trap instead of hanging the GPU).
#include "chain_gemm.cuh"

#include <stdlib.h>

#include "tc_ptx.cuh"

namespace msf {

namespace {

constexpr int CH_THREADS = 384;
constexpr int CH_EPI_WARPS = 8;
constexpr int CH_MAX_STAGES = 4;
constexpr uint32_t CH_A_BYTES = 128 * 64 * 2;  // one K-block of the A operand: 128 rows x 64 bf16
constexpr size_t CH_SMEM_LIMIT = 232448;

__host__ __device__ constexpr uint32_t ch_b_bytes(int H) { return (uint32_t)H * 64 * 2; }

struct Raw16 {
  uint4 lo, hi;
};
__device__ __forceinline__ Raw16 ld_row16(const __nv_bfloat16* p, bool ok) {
  Raw16 r;
  r.lo = make_uint4(0u, 0u, 0u, 0u);
  r.hi = r.lo;
  if (ok) {
    r.lo = __ldg(reinterpret_cast<const uint4*>(p));
    r.hi = __ldg(reinterpret_cast<const uint4*>(p) + 1);
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

// 16 results of row `trow` (0..127), tile columns [c, c+16) -> the K-major, 128B-swizzled block layout
// tcgen05.mma and the TMA box share: [c / 64][trow][64 bf16], 16-byte chunk index XOR (trow & 7).
__device__ __forceinline__ void st_swizzled16(unsigned char* blk, int trow, int c, const float (&v)[16]) {
  unsigned char* rowp = blk + (c >> 6) * CH_A_BYTES + trow * 128;
  const int j0 = (c & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint4 pk;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
    *reinterpret_cast<uint4*>(rowp + (((j0 + q) ^ (trow & 7)) << 4)) = pk;
  }
}

template <int MODE>
__global__ void __launch_bounds__(CH_THREADS, 1) chain_kernel(const __grid_constant__ ChainLaunch L) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int H = L.H, KB = L.H >> 6, STAGES = L.stages, M = L.M;
  const uint32_t B_BYTES = ch_b_bytes(H);
  const uint32_t a_base = smem_base;                                  // ring: A blocks
  const uint32_t b_base = a_base + STAGES * CH_A_BYTES;               // ring: B blocks
  const uint32_t u_base = b_base + STAGES * B_BYTES;                  // intermediate / output staging, KB blocks
  const uint32_t bar_base = u_base + KB * CH_A_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (CH_MAX_STAGES + s); };
  const uint32_t u_full = bar_base + 8u * (2 * CH_MAX_STAGES + 0);    // GEMM1 finished (MMA -> epilogue)
  const uint32_t u_empty = bar_base + 8u * (2 * CH_MAX_STAGES + 1);   // TMEM intermediate drained (epilogue -> MMA)
  const uint32_t us_ready = bar_base + 8u * (2 * CH_MAX_STAGES + 2);  // smem intermediate written (epilogue -> MMA, store)
  const uint32_t us_free = bar_base + 8u * (2 * CH_MAX_STAGES + 3);   // smem block reusable (MMA + store -> epilogue)
  const uint32_t acc_full = bar_base + 8u * (2 * CH_MAX_STAGES + 4);  // ACC complete (MMA -> epilogue)
  const uint32_t acc_empty = bar_base + 8u * (2 * CH_MAX_STAGES + 5); // ACC drained (epilogue -> MMA)
  const uint32_t out_ready = bar_base + 8u * (2 * CH_MAX_STAGES + 6); // final tile staged (epilogue -> store)
  const uint32_t tmem_slot = bar_base + 8u * (2 * CH_MAX_STAGES + 7);
  const uint32_t off0 = smem_u32(smem_raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - off0));
  float* bias_smem = reinterpret_cast<float*>(smem_raw + (bar_base + 8u * (2 * CH_MAX_STAGES + 8) - off0));  // 3 x 256
  unsigned char* u_smem = smem_raw + (u_base - off0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (2 * H <= 32) ? 32 : (2 * H <= 64) ? 64 : (2 * H <= 128) ? 128 : (2 * H <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&L.map_a1);
    tma_prefetch_desc(&L.map_w1);
    tma_prefetch_desc(&L.map_w2);
    tma_prefetch_desc(&L.map_out1);
    tma_prefetch_desc(&L.map_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(u_full, 1);
    mbar_init(u_empty, CH_EPI_WARPS);
    mbar_init(us_ready, CH_EPI_WARPS);
    mbar_init(us_free, 2);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, CH_EPI_WARPS);
    mbar_init(out_ready, CH_EPI_WARPS);
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
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < L.items; item += gridDim.x) {
        const int o = L.active[item % L.n_active], m0 = (item / L.n_active) * 128;
        const int n = L.outer[o].n;
        for (int i = 0; i < n; ++i) {
          const int zin = L.outer[o].inner[i], zp = L.outer[o].pair[i];
          for (int kb = 0; kb < KB; ++kb) {  // GEMM1 operands
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), CH_A_BYTES + B_BYTES);
            tma_load_3d(a_base + stage * CH_A_BYTES, &L.map_a1, kb * 64, m0, zin, full_bar(stage));
            tma_load_3d(b_base + stage * B_BYTES, &L.map_w1, kb * 64, 0, zp, full_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          for (int kb = 0; kb < KB; ++kb) {  // GEMM2: only the weights travel, A is already on chip
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), B_BYTES);
            tma_load_3d(b_base + stage * B_BYTES, &L.map_w2, kb * 64, 0, zp, full_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================
    if (lane == 0) {
      const uint32_t idesc = instr_desc(H, false, false);
      int stage = 0;
      uint32_t phase = 0, u_cnt = 0, item_cnt = 0;
      for (int item = blockIdx.x; item < L.items; item += gridDim.x, ++item_cnt) {
        const int n = L.outer[L.active[item % L.n_active]].n;
        mbar_wait(acc_empty, (item_cnt & 1u) ^ 1u);
        tc_fence_after();
        for (int i = 0; i < n; ++i, ++u_cnt) {
          mbar_wait(u_empty, (u_cnt & 1u) ^ 1u);
          tc_fence_after();
          for (int kb = 0; kb < KB; ++kb) {  // GEMM1 -> TMEM [0, H)
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = a_base + stage * CH_A_BYTES, b_addr = b_base + stage * B_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(tmem_base, smem_desc(a_addr + k * 32, 16, 1024), smem_desc(b_addr + k * 32, 16, 1024), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(empty_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          tc_commit(u_full);
          mbar_wait(us_ready, u_cnt & 1u);  // the epilogue has staged the bf16 intermediate
          tc_fence_after();
          for (int kb = 0; kb < KB; ++kb) {  // GEMM2 -> TMEM [H, 2H), A from the staged block
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = u_base + kb * CH_A_BYTES, b_addr = b_base + stage * B_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(tmem_base + (uint32_t)H, smem_desc(a_addr + k * 32, 16, 1024),
                          smem_desc(b_addr + k * 32, 16, 1024), idesc, (i > 0 || kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(empty_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          tc_commit(us_free);  // 1 of 2: GEMM2 no longer reads the staged block
        }
        tc_commit(acc_full);
      }
    }
  } else if (warp == 3) {
    // =========================== TMA store ==============================
    if (lane == 0) {
      uint32_t u_cnt = 0, item_cnt = 0;
      for (int item = blockIdx.x; item < L.items; item += gridDim.x, ++item_cnt) {
        const int o = L.active[item % L.n_active], m0 = (item / L.n_active) * 128;
        const int n = L.outer[o].n;
        for (int i = 0; i < n; ++i, ++u_cnt) {
          mbar_wait(us_ready, u_cnt & 1u);
          if (L.store1) {
            for (int kb = 0; kb < KB; ++kb)
              tma_store_3d(&L.map_out1, u_base + kb * CH_A_BYTES, kb * 64, m0, L.outer[o].pair[i]);
            tma_store_commit();
            tma_store_wait_read();
          }
          mbar_arrive(us_free);  // 2 of 2
        }
        mbar_wait(out_ready, item_cnt & 1u);
        for (int kb = 0; kb < KB; ++kb) tma_store_3d(&L.map_out, u_base + kb * CH_A_BYTES, kb * 64, m0, o);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(us_free);  // the final tile has no GEMM2 reader: both arrivals come from here
        mbar_arrive(us_free);
      }
      tma_store_wait_all();
    }
  } else if (warp >= 4) {
    // =========================== epilogue ===============================
    const DropCfg drop = resolve_drop(L.drop);
    const int lq = warp & 3, cg = (warp - 4) >> 2;
    const int trow = lq * 32 + lane;                 // row inside the tile = TMEM lane
    const int c_begin = cg * (H >> 1), c_end = c_begin + (H >> 1);
    const int et = threadIdx.x - 128;
    const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
    uint32_t u_cnt = 0, us_uses = 0, item_cnt = 0;
    for (int item = blockIdx.x; item < L.items; item += gridDim.x, ++item_cnt) {
      const int o = L.active[item % L.n_active], m0 = (item / L.n_active) * 128;
      const int n = L.outer[o].n;
      const long long row = (long long)m0 + trow;
      const bool row_ok = row < L.rows;
      for (int i = 0; i < n; ++i, ++u_cnt, ++us_uses) {
        const int zp = L.outer[o].pair[i];
        float* bias_s = bias_smem + (u_cnt & 1u) * 256;
        if (MODE == 0) {  // value_proj bias row -> shared memory
          const float* bsrc = L.bias1[zp];
          for (int e = et; e < H; e += 32 * CH_EPI_WARPS) bias_s[e] = bsrc ? __ldg(bsrc + e) : 0.0f;
          asm volatile("bar.sync 1, %0;" ::"n"(32 * CH_EPI_WARPS) : "memory");
        }
        float gate_mask = 1.0f;
        if (MODE == 0 && L.mask != nullptr && row_ok) gate_mask = __ldg(L.mask + row * M + L.outer[o].mask_col[i]);
        const int sub = L.outer[o].sub[i];
        float* gate_out = (MODE == 0 && L.gate_out) ? L.gate_out + ((long long)zp * L.rows + row) * L.heads : nullptr;
        const float* gate_in = (MODE == 1) ? L.gate_in + ((long long)zp * L.rows + row) * L.heads : nullptr;

        mbar_wait(u_full, u_cnt & 1u);
        tc_fence_after();
        mbar_wait(us_free, (us_uses & 1u) ^ 1u);
        int cur_head = -1;
        float cur_gate = 0.0f;
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 16) {
          uint32_t acc[16];
          tmem_ld16_issue(tmem_base + lane_base + (uint32_t)c, acc);
          tmem_wait16(acc);
          float v[16];
          if ((L.head_dim & 15) == 0) {  // the 16 columns lie inside one head
            const int head = c / L.head_dim;
            if (head != cur_head) {
              cur_head = head;
              if (MODE == 0) {
                cur_gate = (gate_mask != 0.0f) ? 1.0f : 0.0f;
                if (drop.active) cur_gate *= drop1(drop, SITE_ATTN, sub, row, head);
                if (gate_out != nullptr && row_ok && c == head * L.head_dim) gate_out[head] = cur_gate;
              } else {
                cur_gate = row_ok ? __ldg(gate_in + head) : 0.0f;
              }
            }
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (MODE == 0) b4 = *reinterpret_cast<const float4*>(bias_s + c + 4 * q4);
              v[4 * q4 + 0] = (__uint_as_float(acc[4 * q4 + 0]) + b4.x) * cur_gate;
              v[4 * q4 + 1] = (__uint_as_float(acc[4 * q4 + 1]) + b4.y) * cur_gate;
              v[4 * q4 + 2] = (__uint_as_float(acc[4 * q4 + 2]) + b4.z) * cur_gate;
              v[4 * q4 + 3] = (__uint_as_float(acc[4 * q4 + 3]) + b4.w) * cur_gate;
            }
          } else {  // narrow heads (head_dim not a multiple of 16): per-column gate
#pragma unroll 1
            for (int j = 0; j < 16; ++j) {
              const int head = (c + j) / L.head_dim;
              float g;
              if (MODE == 0) {
                g = (gate_mask != 0.0f) ? 1.0f : 0.0f;
                if (drop.active) g *= drop1(drop, SITE_ATTN, sub, row, head);
                if (gate_out != nullptr && row_ok && c + j == head * L.head_dim) gate_out[head] = g;
              } else {
                g = row_ok ? __ldg(gate_in + head) : 0.0f;
              }
              const float b = (MODE == 0) ? bias_s[c + j] : 0.0f;
              float a = 0.0f;
#pragma unroll
              for (int jj = 0; jj < 16; ++jj)
                if (jj == j) a = __uint_as_float(acc[jj]);
              const float r = (a + b) * g;
#pragma unroll
              for (int jj = 0; jj < 16; ++jj)
                if (jj == j) v[jj] = r;
            }
          }
          st_swizzled16(u_smem, trow, c, v);
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(u_empty);
          mbar_arrive(us_ready);
        }
      }

      // ---- final epilogue over ACC ----
      float* bias_s = bias_smem + 2 * 256;  // own buffer: the pair loop double-buffers the other two
      if (MODE == 0) {  // sum of the out_proj biases of this query's pairs
        for (int e = et; e < H; e += 32 * CH_EPI_WARPS) {
          float bsum = 0.0f;
          for (int i = 0; i < n; ++i) {
            const float* bsrc = L.bias2[L.outer[o].pair[i]];
            if (bsrc) bsum += __ldg(bsrc + e);
          }
          for (int i = 0; i < L.outer[o].nb; ++i) {
            const float* bsrc = L.bias2[L.outer[o].bias_only[i]];
            if (bsrc) bsum += __ldg(bsrc + e);
          }
          bias_s[e] = bsum;
        }
        asm volatile("bar.sync 2, %0;" ::"n"(32 * CH_EPI_WARPS) : "memory");
      }
      const __nv_bfloat16* aux_row = L.aux + ((long long)o * L.rows + row) * H;
      const __nv_bfloat16* aux2_row = (MODE == 1) ? L.aux2 + ((long long)o * L.rows + row) * H : nullptr;
      float rscale = 1.0f;
      if (MODE == 0) rscale = L.inv_cnt[o] * ((L.mask != nullptr && row_ok) ? __ldg(L.mask + row * M + o) : 1.0f);
      Raw16 nxt = ld_row16(aux_row + c_begin, row_ok), nxt2;
      if (MODE == 1) nxt2 = ld_row16(aux2_row + c_begin, row_ok);

      mbar_wait(acc_full, item_cnt & 1u);
      tc_fence_after();
      mbar_wait(us_free, (us_uses & 1u) ^ 1u);
      ++us_uses;
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 16) {
        uint32_t acc[16];
        if (n > 0) {
          tmem_ld16_issue(tmem_base + lane_base + (uint32_t)(H + c), acc);
          tmem_wait16(acc);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = 0u;  // no pair module: nothing was accumulated
        }
        float aux[16], aux2[16], v[16];
        unpack16(nxt, aux);
        const bool more = c + 16 < c_end;
        nxt = ld_row16(aux_row + c + 16, row_ok && more);
        if (MODE == 1) {
          unpack16(nxt2, aux2);
          nxt2 = ld_row16(aux2_row + c + 16, row_ok && more);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = __uint_as_float(acc[j]);
          if (MODE == 0) v[j] = rscale != 0.0f ? (a + bias_s[c + j] + aux[j]) * rscale : 0.0f;  // masked row: exact 0
          else v[j] = (a + aux[j]) * (aux2[j] > 0.0f ? L.scale : 0.0f);
        }
        st_swizzled16(u_smem, trow, c, v);
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc_empty);
        mbar_arrive(out_ready);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

size_t chain_fixed_smem() { return 1024 + 8 * (2 * CH_MAX_STAGES + 8) + 3 * 256 * 4; }

}  // namespace

bool chain2_eligible(int H, int M);                                          // chain2_gemm.cu
int chain2_launch(ChainLaunch& L, cudaStream_t stream, const char* label);

// kernel choice.  2 = chain2_kernel (single CTA, half-pair pipeline): the default for hidden % 128 == 0;
// 1 = chain_kernel (every other shape, or MSF_CHAIN=v1).  (A cluster variant that shared every weight fetch by TMA
// multicast, chain3_kernel, measured slower at B = 4096 — 30 vs 24 us per launch — and was removed.)
static int chain_variant(int H, int M, int heads, long long rows) {
  const char* e = getenv("MSF_CHAIN");
  if (getenv("MSF_CHAIN_V1") || (e && e[0] == 'v' && e[1] == '1')) return 1;
  return chain2_eligible(H, M) ? 2 : 1;
}
int chain_w1_box_rows(int H, int M, int heads, long long rows) {
  return chain_variant(H, M, heads, rows) == 2 ? H / 2 : H;
}
int chain_w2_box_rows(int H, int M, int heads, long long rows) {
  (void)M; (void)heads; (void)rows;
  return H;
}

bool chain_eligible(int H, int M) { return H % 64 == 0 && H >= 64 && H <= 256 && M >= 1 && M <= MSF_MAX_MODALITIES; }

int chain_launch(ChainLaunch& L, cudaStream_t stream, const char* label) {
  MSF_REQUIRE(chain_eligible(L.H, L.M), "chain_gemm: hidden %d / modalities %d not supported", L.H, L.M);
  MSF_REQUIRE(L.rows >= 1, "chain_gemm: empty batch");
  L.head_shift = -1;
  for (int sft = 0; sft < 16; ++sft)
    if ((1 << sft) == L.head_dim) L.head_shift = sft;
  const int variant = chain_variant(L.H, L.M, L.heads, L.rows);
  if (variant == 2) return chain2_launch(L, stream, label);
  L.row_tiles = (int)ceil_div(L.rows, 128);
  if (L.n_active <= 0) {   // default: every modality is an outer modality
    L.n_active = L.M;
    for (int m = 0; m < L.M; ++m) L.active[m] = (short)m;
  }
  L.items = L.row_tiles * L.n_active;
  const size_t per_stage = CH_A_BYTES + ch_b_bytes(L.H);
  const size_t ublock = (size_t)(L.H / 64) * CH_A_BYTES;
  int stages = (int)((CH_SMEM_LIMIT - chain_fixed_smem() - ublock) / per_stage);
  if (stages > CH_MAX_STAGES) stages = CH_MAX_STAGES;
  MSF_REQUIRE(stages >= 2, "chain_gemm: not enough shared memory for hidden %d", L.H);
  L.stages = stages;
  const size_t smem = chain_fixed_smem() + ublock + (size_t)stages * per_stage;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    MSF_CHECK_CUDA(cudaGetDevice(&dev));
    MSF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = L.items < sms ? L.items : sms;
  if (prof_enabled()) {
    double pairs = 0.0;
    for (int o = 0; o < L.M; ++o) pairs += L.outer[o].n;
    prof_begin(label, 2.0 * 2.0 * (double)L.rows * L.H * L.H * pairs, stream);
  }
  if (L.mode == 0) {
    MSF_CHECK_CUDA(cudaFuncSetAttribute(chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSF_CHECK_CUDA(launch_pdl(chain_kernel<0>, dim3(grid), dim3(CH_THREADS), smem, stream, L));
  } else {
    MSF_CHECK_CUDA(cudaFuncSetAttribute(chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MSF_CHECK_CUDA(launch_pdl(chain_kernel<1>, dim3(grid), dim3(CH_THREADS), smem, stream, L));
  }
  MSF_LAUNCH_CHECK();
  prof_end(stream);
  return MSF_OK;
}

}  // namespace msf
