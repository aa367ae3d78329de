// lstm_bwd_kernel: backward pass of the LSTM recurrence of SequenceEncoder in training mode (src/encoders.py:54-65,
// 135-166 under autograd: nn.LSTM(F, H, batch_first), one layer, zero initial state, optionally packed windows),
// as ONE persistent launch over all T time steps walked backwards, for up to MSF_LSTM_MAX_SEQS encoders.
//
// The forward pass in training mode (lstm_seq_kernel<true>) kept, per step t: the gate activations i, f, g, o
// (bf16, [T][B][4H], columns gate-interleaved 4u + gate), the cell state c_t (fp32) and h_{t-1} (bf16, [T+1][B][H]).
// With a_t the gate pre-activations:
//
//     d h_t  = [t == last valid step] d h_out  +  d a_{t+1} W_hh                    (tcgen05: K = 4H, N = H)
//     d c_t += d h_t o (1 - tanh^2 c_t);   d a_t = (d c_t g i(1-i), d c_t c_{t-1} f(1-f), d c_t i (1-g^2),
//     d h_t tanh(c_t) o(1-o));   d c_{t-1} = d c_t f                                (epilogue, thread = window)
//
//   * a CLUSTER of H/64 CTAs owns a fixed set of 128-window tiles of one encoder, exactly as in the forward kernel.
//     CTA j computes d h for hidden units [64j, 64j+64): its B operand, rows [64j, 64j+64) of W_hh^T ([H][4H]), stays
//     RESIDENT in shared memory for the whole sequence (128 KB at H = 256); per step only the A operand moves:
//     d a_{t+1} of the cluster's tiles (4H/64 k-blocks of 16 KB per tile) through a TMA ring.
//   * the epilogue warps turn d h_t of their 16 units into d a_t and write it OVER the gate activations of step t:
//     the same buffer is the next step's A operand and, afterwards, the A operand of the weight-gradient GEMMs.
//   * d a_t of a tile is produced by all CTAs of the cluster and consumed by all of them: one cluster barrier per step.
//   * the carried d c (fp32; GRU: the direct d h_t z_t path) of a thread's windows lives in TMEM columns behind the two
//     accumulators (tcgen05.ld / tcgen05.st, up to 6 tiles per cluster), not in global memory; the tape is read and d a
//     written with one 256-bit access per 32-byte sector (thread = window: a warp touches 32 different rows), and two
//     of a thread's four unit groups are requested before it waits for the tile's accumulator.
//
// After the recurrence:  d W_hh = sum_t d a_t^T h_{t-1},  d W_ih = sum_t d a_t^T x_t  run on the grouped tensor-core GEMM
// (tc_gemm_kernel<true>: contraction over (t, window)), split over chunks of steps into fp32 partial sums that
// lstm_wgrad_reduce_kernel adds up in fixed order (bit-reproducible) and scatters into nn.LSTM's gate-major row order.
// The bias gradient is the column `features` of d W_ih: x carries 1.0 there and the weight column is zero.
#include <stdlib.h>
#include <string.h>

#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace msf {
namespace {

constexpr int LB_EPI_WARPS = 16;
constexpr int LB_THREADS = 128 + 32 * LB_EPI_WARPS;
constexpr int LB_MAX_STAGES = 8;
constexpr uint32_t LB_A_BYTES = 128 * 64 * 2;   // one k-block of d a: 128 windows x 64 gate columns
constexpr uint32_t LB_W_BYTES = 64 * 64 * 2;    // one k-block of this CTA's rows of W_hh^T: 64 units x 64 gate columns
constexpr size_t LB_SMEM_LIMIT = 232448;
constexpr int LB_NBAR = 2 * LB_MAX_STAGES + 5;
constexpr int LB_ACC_COLS = 64;                 // accumulator: 128 windows x 64 hidden units
constexpr int LB_DC_COL0 = 2 * LB_ACC_COLS;     // behind the two accumulators: the carried d c of up to LB_DC_TILES tiles
constexpr int LB_DC_TILES = (512 - LB_DC_COL0) / 64;

struct LstmBwdMaps {
  CUtensorMap da, wt;
};
struct LstmBwdLaunch {
  LstmBwdMaps m[MSF_LSTM_MAX_SEQS];
  __nv_bfloat16* gates[MSF_LSTM_MAX_SEQS];
  const float* c_all[MSF_LSTM_MAX_SEQS];
  const __nv_bfloat16* h_all[MSF_LSTM_MAX_SEQS];    // GRU: h_{t-1} of every step
  float* dc[MSF_LSTM_MAX_SEQS];
  const float* d_h_out[MSF_LSTM_MAX_SEQS];          // gradient of the last valid hidden state, or nullptr
  const __nv_bfloat16* dh_all[MSF_LSTM_MAX_SEQS];   // [T][B][H] gradient of every step's hidden state (from the layer above), or nullptr
  const int* lengths[MSF_LSTM_MAX_SEQS];
  int n, rows, steps, hidden, kb4, cs, cps, row_tiles, stages;
  int dc_tmem;   // 1: d c (GRU: the direct d h z path) lives in TMEM columns [128, 128 + 64 tiles) instead of `dc`
  int dbg;   // MSF_LSTM_DBG: 16 = cycle stamps of CTA 0, 8 = no gate / cell traffic in the epilogue (timing only),
             // 4 = keep d c in global memory
};

__device__ __forceinline__ uint32_t lb_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void lb_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float lb_tanh(float x) {   // the forward kernel's tanh(c_t): same instruction
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// the forward kernel's cell-state layout (lstm_seq.cu: ls_cell_index)
__device__ __forceinline__ long long lb_cell_index(int tile, int r, int u, int H, bool ragged) {
  const long long base = (long long)tile * 128 * H;
  return ragged ? base + (long long)r * H + u : base + ((long long)(u >> 2) * 128 + r) * 4 + (u & 3);
}
__device__ __forceinline__ float lb_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float lb_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t lb_pack(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// cycles of CTA 0 summed over the steps (dbg & 16): [0] step start -> producer done, [1] -> MMAs issued,
// [2] -> first accumulator complete, [3] -> last accumulator complete, [4] -> epilogue done, [5] -> barrier passed
__device__ long long g_lb_stamps[8];

// GRU = true: the cell backward of nn.GRU (forward kernel with cell_type 1: columns (r, z, n_x, n_h) per unit, the tape
// holds (r, z, n, a_nh + b_hn)):  d n = d h (1-z)(1-n^2),  d z = d h (h_{t-1} - n) z(1-z),  d r = d n (a_nh + b_hn) r(1-r),
// d a = (d r, d z, d n, d n r), and d h_{t-1} gets d h z directly (carried in fp32 in the `dc` buffer) beside d a W_hh.
template <bool GRU>
__global__ void __launch_bounds__(LB_THREADS, 1) lstm_bwd_kernel(const __grid_constant__ LstmBwdLaunch L) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t off0 = smem_u32(smem_raw);
  const uint32_t base = (off0 + 1023u) & ~1023u;
  const int KB4 = L.kb4, STAGES = L.stages;
  const uint32_t w_base = base;                                    // KB4 weight k-blocks, resident
  const uint32_t ring_base = w_base + (uint32_t)KB4 * LB_W_BYTES;  // A k-blocks
  const uint32_t bar_base = ring_base + (uint32_t)STAGES * LB_A_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (LB_MAX_STAGES + s); };
  const uint32_t w_full = bar_base + 8u * (2 * LB_MAX_STAGES);
  auto acc_full = [&](int a) { return bar_base + 8u * (2 * LB_MAX_STAGES + 1 + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (2 * LB_MAX_STAGES + 3 + a); };
  const uint32_t tmem_slot = bar_base + 8u * LB_NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - off0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)lb_ctarank();
  const int cluster_id = (int)blockIdx.x / L.cs;
  const int seq = cluster_id / L.cps, slot = cluster_id % L.cps;   // which encoder, which share of its tiles
  const int my_tiles = (L.row_tiles - slot + L.cps - 1) / L.cps;   // tiles slot, slot + cps, ...
  const LstmBwdMaps& M = L.m[seq];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&M.da);
    tma_prefetch_desc(&M.wt);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(w_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), LB_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();
  pdl_launch();
  lb_cluster_sync();   // every CTA of the cluster is up before the first step

  int stage = 0;
  uint32_t phase = 0, cnt = 0;

  if (warp == 0 && lane == 0) {   // rows [64 rank, 64 rank + 64) of W_hh^T: resident for the whole sequence
    mbar_expect_tx(w_full, (uint32_t)KB4 * LB_W_BYTES);
    for (int kb = 0; kb < KB4; ++kb) tma_load_3d(w_base + kb * LB_W_BYTES, &M.wt, kb * 64, rank * 64, 0, w_full);
  }
  if (warp == 1 && lane == 0) {
    mbar_wait(w_full, 0u);
    tc_fence_after();
  }
  const uint32_t idesc = instr_desc(LB_ACC_COLS, false, false);
  const int lq = warp & 3, cg = (warp - 4) >> 2;   // epilogue: TMEM lane quarter, column quarter (16 columns = 16 units)
  const int H = L.hidden;
  const long long BH = (long long)L.rows * H;

#pragma unroll 1
  for (int s = 0; s < L.steps; ++s) {
    const int t = L.steps - 1 - s;
    const bool stamp = (L.dbg & 16) && blockIdx.x == 0 && lane == 0;
    const long long t_step = stamp ? clock64() : 0;
    if (warp == 0) {
      // ===== TMA producer: d a_{t+1} of every tile of this cluster (nothing at the last time step) =====
      if (lane == 0 && s > 0) {
        for (int i = 0; i < my_tiles; ++i) {
          const int row0 = (slot + i * L.cps) * 128;
          for (int kb = 0; kb < KB4; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), LB_A_BYTES);
            tma_load_3d(ring_base + stage * LB_A_BYTES, &M.da, kb * 64, row0, t + 1, full_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
        if (stamp) g_lb_stamps[0] += clock64() - t_step;
      }
    } else if (warp == 1) {
      // ===== MMA issuer: d h_t (recurrent part) = d a_{t+1} W_hh restricted to this CTA's 64 units =====
      if (lane == 0 && s > 0) {
        for (int i = 0; i < my_tiles; ++i, ++cnt) {
          const int acc = (int)(cnt & 1u);
          mbar_wait(acc_empty(acc), ((cnt >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * LB_ACC_COLS;
          for (int kb = 0; kb < KB4; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = ring_base + stage * LB_A_BYTES, b_addr = w_base + kb * LB_W_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(d_tmem, smem_desc(a_addr + k * 32, 16, 1024), smem_desc(b_addr + k * 32, 16, 1024), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(empty_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          tc_commit(acc_full(acc));
        }
        if (stamp) g_lb_stamps[1] += clock64() - t_step;
      }
    } else if (warp >= 4) {
      // ===== epilogue: cell backward on this CTA's 64 hidden units (16 per warp column group) =====
      __nv_bfloat16* gates_t = L.gates[seq] + (long long)t * BH * 4;
      const float* c_t = GRU ? nullptr : L.c_all[seq] + (long long)t * BH;
      const float* c_p = GRU ? nullptr : c_t - BH;   // step t - 1 (not read at t = 0: c_{-1} = 0)
      const __nv_bfloat16* h_p = GRU ? L.h_all[seq] + (long long)t * BH : nullptr;   // h_{t-1}
      float* dcp = L.dc[seq];
      const float* dho = L.d_h_out[seq];
      const __nv_bfloat16* dha = L.dh_all[seq] != nullptr ? L.dh_all[seq] + (long long)t * BH : nullptr;
      const int* lens = L.lengths[seq];
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = slot + i * L.cps;
        const int r = lq * 32 + lane, row = tile * 128 + r;
        const bool row_ok = row < L.rows;
        const bool ragged = (tile + 1) * 128 > L.rows;
        const int ubase = rank * 64 + cg * 16;   // first hidden unit of this thread's 16
        const int len = (lens != nullptr && row_ok) ? __ldg(lens + row) : L.steps;
        const bool live = row_ok && t < len;      // the window took this step in the forward pass
        const bool ext = row_ok && dho != nullptr && t == len - 1;  // h_out of the window is h_t
        // What the forward pass kept for these windows does not depend on this step's MMAs: the loads of two of the
        // thread's four unit groups are issued BEFORE the wait for the accumulator (one DRAM round trip hidden behind
        // the tile's operand loads and MMAs), the other two right after the first two are consumed.
        struct Kept {
          uint32_t g[8];    // gate activations of 4 units
          float4 ct, cp;    // c_t, c_{t-1}            (LSTM)
        };
        const bool use = live && !(L.dbg & 8);
        auto load_kept = [&](int g, Kept& K) {
          if (!use) return;
          const int u0 = ubase + 4 * g;
          ld_global_v8(gates_t + (long long)row * 4 * H + 4 * u0, K.g);
          if (!GRU) {
            const long long idx = lb_cell_index(tile, r, u0, H, ragged);
            K.ct = *reinterpret_cast<const float4*>(c_t + idx);
            K.cp = t > 0 ? *reinterpret_cast<const float4*>(c_p + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        // 16 units x bf16 = one 32-byte sector per thread: h_{t-1} (GRU) and the gradient from the layer above
        uint32_t hpv[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, dhv[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (GRU && use) ld_global_v8(h_p + (long long)row * H + ubase, hpv);
        if (dha != nullptr && use) ld_global_v8(dha + (long long)row * H + ubase, dhv);
        Kept K0, K1;
        load_kept(0, K0);
        load_kept(1, K1);
        uint32_t a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = 0u;
        if (s > 0) {
          const int acc = (int)(cnt & 1u);
          mbar_wait(acc_full(acc), (cnt >> 1) & 1u);
          tc_fence_after();
          if (stamp && warp == 4 && i == 0) g_lb_stamps[2] += clock64() - t_step;
          if (stamp && warp == 4 && i == my_tiles - 1) g_lb_stamps[3] += clock64() - t_step;
          const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)acc * LB_ACC_COLS + (uint32_t)(cg * 16);
          tmem_ld16_issue(taddr, a);
          tmem_wait16(a);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(acc));   // the accumulator is in registers: the next tile may use it
          ++cnt;
        }
        auto cell_backward = [&](int g, const Kept& K) {
          const int u0 = ubase + 4 * g;
          __nv_bfloat16* gp = gates_t + (long long)row * 4 * H + 4 * u0;
          // the carried d c of these 4 units: a TMEM scratch column group of this thread's lane (warp-collective
          // access: outside the per-window branch), or the global `dc` buffer
          const uint32_t dc_taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(LB_DC_COL0 + i * 64 + cg * 16 + 4 * g);
          uint32_t dct[4] = {0u, 0u, 0u, 0u};
          if (L.dc_tmem && s > 0) tmem_ld4(dc_taddr, dct);
          if (use) {
            const long long idx = lb_cell_index(tile, r, u0, H, ragged);
            const float4 dc4 = L.dc_tmem ? make_float4(__uint_as_float(dct[0]), __uint_as_float(dct[1]), __uint_as_float(dct[2]),
                                                       __uint_as_float(dct[3]))
                                         : *reinterpret_cast<const float4*>(dcp + idx);
            float4 ex4 = ext ? __ldg(reinterpret_cast<const float4*>(dho + (long long)row * H + u0))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
            if (dha != nullptr) {   // this step's hidden state also fed the layer above
              ex4.x += lb_lo(dhv[2 * g]); ex4.y += lb_hi(dhv[2 * g]); ex4.z += lb_lo(dhv[2 * g + 1]); ex4.w += lb_hi(dhv[2 * g + 1]);
            }
            const uint32_t (&gw)[8] = K.g;
            const float dcv[4] = {dc4.x, dc4.y, dc4.z, dc4.w}, ex[4] = {ex4.x, ex4.y, ex4.z, ex4.w};
            uint32_t out[8];
            float dcn[4];
            if (GRU) {
              const float hp[4] = {lb_lo(hpv[2 * g]), lb_hi(hpv[2 * g]), lb_lo(hpv[2 * g + 1]), lb_hi(hpv[2 * g + 1])};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float gr = lb_lo(gw[2 * j]), gz = lb_hi(gw[2 * j]), gn = lb_lo(gw[2 * j + 1]), hn = lb_hi(gw[2 * j + 1]);
                const float dh = __uint_as_float(a[4 * g + j]) + ex[j] + dcv[j];   // dcv: d h_{t+1} z_{t+1}, the direct path
                const float dn = dh * (1.0f - gz) * (1.0f - gn * gn);
                const float dz = dh * (hp[j] - gn) * gz * (1.0f - gz);
                const float dr = dn * hn * gr * (1.0f - gr);
                dcn[j] = dh * gz;
                out[2 * j] = lb_pack(dr, dz);
                out[2 * j + 1] = lb_pack(dn, dn * gr);
              }
            } else {
              const float ct[4] = {K.ct.x, K.ct.y, K.ct.z, K.ct.w}, cp[4] = {K.cp.x, K.cp.y, K.cp.z, K.cp.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float gi = lb_lo(gw[2 * j]), gf = lb_hi(gw[2 * j]), gg = lb_lo(gw[2 * j + 1]), go = lb_hi(gw[2 * j + 1]);
                const float dh = __uint_as_float(a[4 * g + j]) + ex[j];
                const float tc = lb_tanh(ct[j]);
                const float dcs = fmaf(dh * go, 1.0f - tc * tc, dcv[j]);
                const float dai = dcs * gg * gi * (1.0f - gi);
                const float daf = dcs * cp[j] * gf * (1.0f - gf);
                const float dag = dcs * gi * (1.0f - gg * gg);
                const float dao = dh * tc * go * (1.0f - go);
                dcn[j] = dcs * gf;
                out[2 * j] = lb_pack(dai, daf);
                out[2 * j + 1] = lb_pack(dag, dao);
              }
            }
            if (L.dc_tmem) {
#pragma unroll
              for (int j = 0; j < 4; ++j) dct[j] = __float_as_uint(dcn[j]);
            } else {
              *reinterpret_cast<float4*>(dcp + idx) = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]);
            }
            st_global_v8(gp, out);
          } else if (row_ok) {   // a step behind the window's length: no gradient through it
            const uint32_t zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            st_global_v8(gp, zero);
          }
          if (L.dc_tmem) tmem_st4(dc_taddr, dct);   // (unchanged for a window that did not take this step)
        };
        Kept K2;
        load_kept(2, K2);   // in flight while groups 0 and 1 are computed
        cell_backward(0, K0);
        load_kept(3, K0);   // (group 0's registers are free again)
        cell_backward(1, K1);
        cell_backward(2, K2);
        cell_backward(3, K0);
      }
      if (L.dc_tmem) tmem_st_wait();
      // d a_t is read back by TMA (async proxy) in the next step, by every CTA of the cluster
      asm volatile("fence.proxy.async.global;" ::: "memory");
      __threadfence();
      if (stamp && warp == 4) g_lb_stamps[4] += clock64() - t_step;
    }
    __syncwarp();
    lb_cluster_sync();   // d a_t of the cluster's tiles is complete and visible
    if (stamp && warp == 4) g_lb_stamps[5] += clock64() - t_step;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// d W_hh / d W_ih / d bias = sum over the chunks of partial[chunk][4u+g][0..H) / [H..H+F) / [H+F], rows back in
// nn.LSTM's gate-major order (row g*H + u); fixed summation order.
__global__ void lstm_wgrad_reduce_kernel(const float* __restrict__ partial, int chunks, int H, int F, int in_cols,
                                         const float* __restrict__ bias_partial, int bias_blocks,
                                         float* __restrict__ d_w_ih, float* __restrict__ d_w_hh, float* __restrict__ d_bias) {
  const int ld = H + in_cols;
  const long long per_chunk = 4LL * H * ld;
  const long long total = 4LL * H * (H + F + 1);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(e / (H + F + 1)), col = (int)(e % (H + F + 1));   // row = 4u + g
    float sum = 0.0f;
    if (col == H + F && bias_partial != nullptr) {   // no ones column in this layer's input: column sums of d a
      for (int b = 0; b < bias_blocks; ++b) sum += bias_partial[(long long)b * 4 * H + row];
    } else {
      const float* src = partial + (long long)row * ld + col;
      for (int c = 0; c < chunks; ++c) sum += src[(long long)c * per_chunk];
    }
    const int dst_row = (row & 3) * H + (row >> 2);
    if (col < H) d_w_hh[(long long)dst_row * H + col] = sum;
    else if (col < H + F) d_w_ih[(long long)dst_row * F + (col - H)] = sum;
    else d_bias[dst_row] = sum;
  }
}

// Column sums of d a ([rows][4H] bf16) for the bias gradient of the layers above the first: block b adds rows
// b, b + grid, ... (thread = 4 adjacent columns) into bias_partial[b][4H]; the reduction kernel sums the blocks.
constexpr int LB_COLSUM_BLOCKS = 592;
__global__ void __launch_bounds__(256) lstm_colsum_kernel(const __nv_bfloat16* __restrict__ da, long long rows, int cols,
                                                          float* __restrict__ bias_partial) {
  const int c0 = threadIdx.x * 4;
  if (c0 >= cols) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(da + r * cols + c0));
    acc[0] += lb_lo(v.x); acc[1] += lb_hi(v.x); acc[2] += lb_lo(v.y); acc[3] += lb_hi(v.y);
  }
  *reinterpret_cast<float4*>(bias_partial + (long long)blockIdx.x * cols + c0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// Inter-layer dropout of a stacked nn.LSTM (training mode, src/encoders.py:54-65: dropout between the layers):
// out = in * mask with the library's Philox multipliers (site 4, sub = layer), 8 elements per thread.
constexpr int SITE_LSTM_LAYER = 4;
__global__ void lstm_dropout_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long rows,
                                    int cols, DropCfg d, int layer) {
  const int c8n = cols >> 3;
  const long long total = rows * c8n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / c8n;
    const int c8 = (int)(e % c8n);
    float m[8];
    drop8(d, SITE_LSTM_LAYER, layer, row, c8, m);
    const uint4 v = *reinterpret_cast<const uint4*>(in + row * cols + c8 * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = lb_pack(lb_lo(w[j]) * m[2 * j], lb_hi(w[j]) * m[2 * j + 1]);
    *reinterpret_cast<uint4*>(out + row * cols + c8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

template <typename... KArgs, typename... Args>
cudaError_t lb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster,
                      Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = (unsigned)cluster;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int lb_sm_count(int* sms) {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0;
    MSF_CHECK_CUDA(cudaGetDevice(&dev));
    MSF_CHECK_CUDA(cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev));
  }
  *sms = cached;
  return MSF_OK;
}

// The contraction over (t, window) is cut into chunks of whole time steps: enough chunks for two waves of tiles,
// at most 32 (problems per launch) and at most one per step.
void lb_chunking(int steps, int hidden, int sms, int* steps_per_chunk, int* chunks) {
  const int tiles_per_chunk = (4 * hidden / 128) * 2 * ((hidden + 127) / 128);
  int want = (2 * sms + tiles_per_chunk - 1) / tiles_per_chunk;
  if (want > 16) want = 16;
  if (want > steps) want = steps;
  if (want < 1) want = 1;
  const int spc = (steps + want - 1) / want;
  *steps_per_chunk = spc;
  *chunks = (steps + spc - 1) / spc;
}

}  // namespace
}  // namespace msf

extern "C" int msf_lstm_backward_scratch_bytes(int64_t batch, int32_t steps, int32_t hidden, size_t* bytes) {
  using namespace msf;
  MSF_REQUIRE(bytes != nullptr && batch >= 1 && steps >= 1 && hidden >= 64 && hidden <= 256 && hidden % 64 == 0,
              "msf_lstm_backward_scratch_bytes: bad arguments");
  int sms = 0, rc = lb_sm_count(&sms);
  if (rc) return rc;
  int spc, chunks;
  lb_chunking(steps, hidden, sms, &spc, &chunks);
  *bytes = ((size_t)chunks * 4 * hidden * 2 * hidden + (size_t)LB_COLSUM_BLOCKS * 4 * hidden) * sizeof(float);
  return MSF_OK;
}

extern "C" int msf_lstm_backward(const msf_lstm_seq* seqs, int32_t n, int64_t batch, int32_t steps, int32_t hidden,
                                 void* stream) {
  using namespace msf;
  MSF_REQUIRE(seqs != nullptr && n >= 1 && n <= MSF_LSTM_MAX_SEQS, "msf_lstm_backward: 1..%d sequences per call", MSF_LSTM_MAX_SEQS);
  MSF_REQUIRE(batch >= 1 && batch < (1 << 24) && steps >= 1, "msf_lstm_backward: bad batch / steps");
  MSF_REQUIRE(hidden % 64 == 0 && hidden >= 64 && hidden <= 256, "msf_lstm_backward: hidden %d (needs a multiple of 64, <= 256)", hidden);
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 0, rc = lb_sm_count(&sms);
  if (rc) return rc;
  const long long B = batch, H = hidden, N4 = 4LL * hidden;
  LstmBwdLaunch L;
  memset(&L, 0, sizeof(L));
  L.n = n; L.rows = (int)B; L.steps = steps; L.hidden = hidden; L.kb4 = (int)(N4 / 64); L.cs = hidden / 64;
  L.row_tiles = (int)ceil_div(B, 128);
  { const char* e = getenv("MSF_LSTM_DBG"); L.dbg = e ? atoi(e) : 0; }
  MSF_REQUIRE(sms / L.cs >= n, "msf_lstm_backward: %d sequences need %d clusters of %d CTAs", n, n, L.cs);
  int cps = (sms / L.cs) / n;
  if (cps > L.row_tiles) cps = L.row_tiles;
  L.cps = cps;
  L.dc_tmem = (ceil_div(L.row_tiles, cps) <= LB_DC_TILES && !(L.dbg & 4)) ? 1 : 0;
  for (int i = 0; i < n; ++i) {
    const msf_lstm_seq& S = seqs[i];
    MSF_REQUIRE(S.x_bf16 && S.h_all && S.gates && (S.c_all || S.cell_type == 1) && S.w_hh_t && (S.d_h_out || S.d_h_all) &&
                    S.dc && S.partial && S.d_w_ih && S.d_w_hh && S.d_bias,
                "msf_lstm_backward: null pointer in sequence %d", i);
    MSF_REQUIRE(S.cell_type == seqs[0].cell_type && (S.cell_type == 0 || S.cell_type == 1),
                "msf_lstm_backward: cell_type 0 (LSTM) or 1 (GRU), the same for all sequences of a call");
    MSF_REQUIRE(S.in_cols == 0 || S.in_cols == hidden,
                "msf_lstm_backward: in_cols %d (0 for the first layer, hidden for the layers above)", S.in_cols);
    if (S.in_cols != 0) {
      MSF_REQUIRE(S.features == hidden, "msf_lstm_backward: a stacked layer has features == hidden");
    } else {
      MSF_REQUIRE(S.features >= 1 && S.features <= 64, "msf_lstm_backward: features %d (1..64)", S.features);
    }
    if ((rc = tc_encode_map(&L.m[i].da, S.gates, B, N4, N4, steps, B * N4, 64, 128))) return rc;
    if ((rc = tc_encode_map(&L.m[i].wt, S.w_hh_t, H, N4, N4, 1, 0, 64, 64))) return rc;
    L.gates[i] = static_cast<__nv_bfloat16*>(S.gates);
    L.c_all[i] = S.c_all;
    L.h_all[i] = static_cast<const __nv_bfloat16*>(S.h_all);
    L.dc[i] = S.dc;
    L.d_h_out[i] = S.d_h_out;
    L.dh_all[i] = static_cast<const __nv_bfloat16*>(S.d_h_all);
    L.lengths[i] = S.lengths;
  }
  const size_t fixed = 1024 + 8 * LB_NBAR + 64;
  const size_t weights = (size_t)L.kb4 * LB_W_BYTES;
  int stages = (int)((LB_SMEM_LIMIT - fixed - weights) / LB_A_BYTES);
  if (stages > LB_MAX_STAGES) stages = LB_MAX_STAGES;
  MSF_REQUIRE(stages >= 2, "msf_lstm_backward: not enough shared memory for hidden %d", hidden);
  L.stages = stages;
  const size_t smem = fixed + weights + (size_t)stages * LB_A_BYTES;
  const int grid = n * cps * L.cs;
  if (prof_enabled()) prof_begin("LSTM sequence backward", 2.0 * (double)B * H * 4.0 * H * (steps - 1) * n, st);
  auto kernel = seqs[0].cell_type == 1 ? lstm_bwd_kernel<true> : lstm_bwd_kernel<false>;
  MSF_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MSF_CHECK_CUDA(lb_launch(kernel, dim3(grid), dim3(LB_THREADS), smem, st, L.cs, L));
  MSF_LAUNCH_CHECK();
  prof_end(st);
  if (L.dbg & 16) {
    MSF_CHECK_CUDA(cudaStreamSynchronize(st));
    long long v[8];
    MSF_CHECK_CUDA(cudaMemcpyFromSymbol(v, g_lb_stamps, sizeof(v)));
    fprintf(stderr, "lstm_bwd CTA 0, cycles per step: producer done %lld, MMAs issued %lld, first accumulator %lld, last "
                    "accumulator %lld, epilogue done %lld, barrier passed %lld (tiles per cluster %d, stages %d)\n",
            v[0] / steps, v[1] / steps, v[2] / steps, v[3] / steps, v[4] / steps, v[5] / steps,
            (int)ceil_div(L.row_tiles, cps), stages);
    memset(v, 0, sizeof(v));
    MSF_CHECK_CUDA(cudaMemcpyToSymbol(g_lb_stamps, v, sizeof(v)));
  }

  // ---- weight gradients: contraction over (t, window), chunks of whole steps -> fp32 partial sums ----
  int spc, chunks;
  lb_chunking(steps, hidden, sms, &spc, &chunks);
  const int full = steps / spc, rem = steps - full * spc;   // `full` chunks of spc steps (+ one of `rem`)
  const long long rows_c = (long long)spc * B;
  MSF_REQUIRE(rows_c < (1ll << 31), "msf_lstm_backward: chunk of %lld rows too long", rows_c);
  for (int i = 0; i < n; ++i) {
    const msf_lstm_seq& S = seqs[i];
    const bool stacked = S.in_cols != 0;   // input = the layer below's hidden states, no ones column
    const long long IC = stacked ? H : 64, ld_p = H + IC, per_chunk = N4 * ld_p;
    TcBuilder tb(true, 128, no_dropout(), st, "LSTM weight gradients");
    const __nv_bfloat16* da = static_cast<const __nv_bfloat16*>(S.gates);
    const __nv_bfloat16* hh = static_cast<const __nv_bfloat16*>(S.h_all);
    const __nv_bfloat16* xx = static_cast<const __nv_bfloat16*>(S.x_bf16);
    short m_a = -1, m_h = -1, m_x = -1, r_a = -1, r_h = -1, r_x = -1;
    if (full > 0) {
      m_a = (short)tb.add_map(da, rows_c, N4, N4, full, rows_c * N4, 0);
      m_h = (short)tb.add_map(hh, rows_c, H, H, full, rows_c * H, 0);
      m_x = (short)tb.add_map(xx, rows_c, IC, IC, full, rows_c * IC, 0);
    }
    if (rem > 0) {
      const long long done = (long long)full * rows_c;
      r_a = (short)tb.add_map(da + done * N4, (long long)rem * B, N4, N4, 1, 0, 0);
      r_h = (short)tb.add_map(hh + done * H, (long long)rem * B, H, H, 1, 0, 0);
      r_x = (short)tb.add_map(xx + done * IC, (long long)rem * B, IC, IC, 1, 0, 0);
    }
    if (tb.status) return tb.status;
    for (int c = 0; c < chunks; ++c) {
      const bool tail = c >= full;
      for (int which = 0; which < 2; ++which) {   // 0: d W_hh (B operand h_{t-1}), 1: d W_ih | d bias (B operand x_t)
        TcProblem p = tc_blank_problem();
        p.seg[0].a_map = tail ? r_a : m_a;
        p.seg[0].a_z = tail ? 0 : c;
        p.seg[0].b_map = which == 0 ? (tail ? r_h : m_h) : (tail ? r_x : m_x);
        p.seg[0].b_z = tail ? 0 : c;
        p.nseg = 1;
        p.M = (int)N4;
        p.N = which == 0 ? hidden : (int)IC;
        p.K = (int)(tail ? (long long)rem * B : rows_c);
        p.C = S.partial + (long long)c * per_chunk + (which == 0 ? 0 : H);
        p.ldc = ld_p;
        p.c_bf16 = 0;
        p.epi = TC_EPI_STORE;
        if ((rc = tb.add_problem(p))) return rc;
      }
    }
    if ((rc = tb.flush())) return rc;
    float* bias_partial = nullptr;
    if (stacked || S.features == 64) {   // no spare column for the ones of the bias gradient: column sums of d a
      bias_partial = S.partial + (size_t)chunks * per_chunk;
      lstm_colsum_kernel<<<LB_COLSUM_BLOCKS, 256, 0, st>>>(da, (long long)steps * B, (int)N4, bias_partial);
      MSF_LAUNCH_CHECK();
    }
    const long long total = N4 * (H + S.features + 1);
    lstm_wgrad_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(
        S.partial, chunks, hidden, S.features, (int)IC, bias_partial, LB_COLSUM_BLOCKS, S.d_w_ih, S.d_w_hh, S.d_bias);
    MSF_LAUNCH_CHECK();
  }
  return MSF_OK;
}

extern "C" int msf_lstm_dropout(const void* in_bf16, void* out_bf16, int64_t rows, int32_t cols, float p, uint64_t seed,
                                uint64_t offset, int32_t layer, void* stream) {
  using namespace msf;
  MSF_REQUIRE(in_bf16 && out_bf16 && rows >= 0 && cols >= 8 && cols % 8 == 0 && p >= 0.0f && p < 1.0f,
              "msf_lstm_dropout: bad arguments (cols %d must be a multiple of 8)", cols);
  if (rows == 0) return MSF_OK;
  DropCfg d;
  d.seed = seed; d.offset = offset; d.p = p; d.scale = 1.0f / (1.0f - p); d.active = 1; d.state = nullptr;
  long long blocks = ceil_div(rows * (cols / 8), 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  lstm_dropout_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      static_cast<const __nv_bfloat16*>(in_bf16), static_cast<__nv_bfloat16*>(out_bf16), rows, cols, d, layer);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}
