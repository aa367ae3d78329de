// lstm_seq_kernel: the LSTM recurrence of SequenceEncoder (src/encoders.py:54-65,135-166: nn.LSTM(F, H, batch_first),
// one layer, zero initial state, no packing) as ONE persistent launch over all T time steps, for up to
// MSF_LSTM_MAX_SEQS encoders of the same batch / length / hidden size.
//
// The per-step design (lstm.cu: one grouped tcgen05 launch per time step) streams every weight block from the L2
// again for every 128-window tile of every step: 123 MB per step at B = 4096, 54 us per step.  Here
//
//   * a CLUSTER of H/64 CTAs owns one encoder's recurrence for a fixed set of 128-window tiles.  CTA j of the
//     cluster keeps ITS 256 gate columns of W_hh and W_ih (the four gates of hidden units [64j, 64j+64), rows
//     gate-interleaved) RESIDENT in shared memory for the whole sequence: 160 KB at H = 256.  Per step only the
//     A operand moves: [h_{t-1} | x_t] of the cluster's tiles, 80 KB per tile, through a ring of 16 KB k-blocks.
//   * per tile and step: pre = h_{t-1} W_hh^T + x_t W_ih^T as H/64 + 1 k-blocks of tcgen05.mma (M = 128, N = 256)
//     into one of two TMEM accumulators; the epilogue warps (thread = window) apply the cell
//     c_t = f c_{t-1} + i g, h_t = o tanh(c_t) to the four adjacent gate columns of each unit (c fp32 in global
//     memory, L2-resident) and write h_t as bf16 k-block j of the next step's A operand — while the MMAs of the
//     cluster's next tile run into the other accumulator.
//   * h_t of a tile is produced by all CTAs of the cluster and consumed by all of them: one cluster barrier per
//     time step (h travels through global memory / the L2 and comes back by TMA; the writers fence the async proxy
//     before they arrive).  No grid-wide synchronisation, no per-step launch.
//
// Training mode and stacked layers (template SAVE): h_t of every step is kept ([T+1][B][H]), with `gates` set also the
// gate activations (bf16, 256-bit stores: one full sector per thread and group of 4 units) and c_t (fp32) — the tape of
// lstm_bwd.cu; a layer above the first reads the input's share of its pre-activations (z_in, one GEMM over all steps)
// from the gate buffer in place.  Template GRU: the GRU cell on the same machinery (see below).
//
// Roles per CTA: warp 0 lane 0 TMA producer, warp 1 lane 0 MMA issuer, warp 2 TMEM allocator, warps 4..11 epilogue.
// Every mbarrier wait is bounded.
#include <stdlib.h>
#include <string.h>

#include <cuda_fp16.h>

#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace msf {
namespace {

constexpr int LS_EPI_WARPS = 16;                     // 4 per scheduler: the cell arithmetic is MUFU- and latency-bound
constexpr int LS_THREADS = 128 + 32 * LS_EPI_WARPS;
constexpr int LS_MAX_STAGES = 8;
constexpr uint32_t LS_A_BYTES = 128 * 64 * 2;    // one k-block of the A operand: 128 windows x 64
constexpr uint32_t LS_W_BYTES = 256 * 64 * 2;    // one k-block of this CTA's weights: 256 gate columns x 64
constexpr size_t LS_SMEM_LIMIT = 232448;
constexpr int LS_NBAR = 2 * LS_MAX_STAGES + 5;
constexpr uint32_t LS_BIAS_OFF = (8u * LS_NBAR + 16u + 15u) & ~15u;   // behind the barriers and the TMEM slot

struct LstmSeqMaps {
  CUtensorMap x, h[2], whh, wih;
};
struct LstmSeqLaunch {
  LstmSeqMaps m[MSF_LSTM_MAX_SEQS];
  const float* bias[MSF_LSTM_MAX_SEQS];
  void* hbuf[MSF_LSTM_MAX_SEQS][2];
  float* cell[MSF_LSTM_MAX_SEQS];
  float* h_out[MSF_LSTM_MAX_SEQS];
  const int* lengths[MSF_LSTM_MAX_SEQS];   // valid steps per window, or nullptr (all steps)
  // HALL (template): hbuf[.][0] = h_all [T+1][B][H] (m[.].h[0] over it): every hidden state is kept — the input of the
  // next layer and of the backward pass.  Training mode (gates != nullptr): every step's gate activations and cell
  // state are kept as well (lstm_bwd.cu)
  __nv_bfloat16* gates[MSF_LSTM_MAX_SEQS];   // [T][B][4H] bf16, columns gate-interleaved (4u + {i,f,g,o})
  float* c_all[MSF_LSTM_MAX_SEQS];           // [T][B*H] fp32, each step in the cell layout below
  // layers above the first: the input's share of the gate pre-activations, z_t = in_t W_ih^T, computed beforehand by
  // one GEMM over all steps ([T][B][4H] bf16, same column order; may be the `gates` buffer itself: read, then
  // overwritten in place).  Then there is no x / W_ih k-block (nkbx = 0).
  const __nv_bfloat16* zin[MSF_LSTM_MAX_SEQS];
  int nkbx;
  int n, rows, steps, hidden, kbh, cs, cps, row_tiles, stages;
  int dbg;   // MSF_LSTM_DBG: 1 no async-proxy fence, 2 no __threadfence, 8 no cell-state traffic, 16 stamps
  long long h_slice;
};

__device__ __forceinline__ uint32_t ls_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void ls_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// One MUFU per activation (tanh.approx.f32, relative error 2^-11: below the bf16 rounding of h every step) instead of
// the exp + reciprocal pair: the cell arithmetic of a tile is 320 activations per window and bounds the step.
__device__ __forceinline__ float ls_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ls_sigmoid(float x) { return fmaf(0.5f, ls_tanh(0.5f * x), 0.5f); }
// Two activations per MUFU operation: tanh.approx.f16x2 on a packed pair (fp16 in / out: 2^-11 relative, the size of
// the bf16 rounding h gets every step anyway).  The cell state itself stays fp32.
__device__ __forceinline__ float2 ls_tanh2(float lo, float hi) {
  __half2 p = __floats2half2_rn(lo, hi);
  uint32_t r, in = *reinterpret_cast<uint32_t*>(&p);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(in));
  return __half22float2(*reinterpret_cast<__half2*>(&r));
}
// Cell state: fp32, one 128-window tile after the other; inside a full tile unit-group-major ([H/4][128][4]) so that a
// warp's access (32 windows x 4 units) is one contiguous 512-byte run; a ragged last tile stays window-major.
__device__ __forceinline__ long long ls_cell_index(int tile, int r, int u, int H, bool ragged) {
  const long long base = (long long)tile * 128 * H;
  return ragged ? base + (long long)r * H + u : base + ((long long)(u >> 2) * 128 + r) * 4 + (u & 3);
}

// cycles of CTA 0 summed over the steps (dbg & 16): [0] step start -> producer done, [1] -> MMAs issued,
// [2] -> accumulators complete (epilogue warp 4 sees the last tile), [3] -> epilogue done, [4] -> cluster barrier passed
__device__ long long g_ls_stamps[8];

// GRU (nn.GRU, gate order r, z, n; src/encoders.py:66-72): the same kernel with the four accumulator columns of a unit
// holding (r, z, n_x, n_h) — the weight rows are packed so that column n_x only receives the input's share of the
// candidate gate and n_h only the recurrent share (ops.gru_pack_weights) — and the cell
//     r = s(a_r + b_r), z = s(a_z + b_z), n = tanh(a_nx + b_in + r (a_nh + b_hn)), h_t = n + z (h_{t-1} - n)
// with h kept in fp32 in the `cell` buffer.  Training mode keeps (r, z, n, a_nh + b_hn) per unit in the gate buffer.
template <bool SAVE, bool GRU>
__global__ void __launch_bounds__(LS_THREADS, 1) lstm_seq_kernel(const __grid_constant__ LstmSeqLaunch L) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t off0 = smem_u32(smem_raw);
  const uint32_t base = (off0 + 1023u) & ~1023u;
  const int KBH = L.kbh, NKB = L.kbh + L.nkbx, STAGES = L.stages;
  const bool HASX = L.nkbx != 0;
  const uint32_t w_base = base;                                    // NKB weight k-blocks, resident
  const uint32_t ring_base = w_base + (uint32_t)NKB * LS_W_BYTES;  // A k-blocks
  const uint32_t bar_base = ring_base + (uint32_t)STAGES * LS_A_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (LS_MAX_STAGES + s); };
  const uint32_t w_full = bar_base + 8u * (2 * LS_MAX_STAGES);
  auto acc_full = [&](int a) { return bar_base + 8u * (2 * LS_MAX_STAGES + 1 + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (2 * LS_MAX_STAGES + 3 + a); };
  const uint32_t tmem_slot = bar_base + 8u * LS_NBAR;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - off0));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bar_base + LS_BIAS_OFF - off0));   // 256 floats, 16-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ls_ctarank();
  const int cluster_id = (int)blockIdx.x / L.cs;
  const int seq = cluster_id / L.cps, slot = cluster_id % L.cps;   // which encoder, which share of its tiles
  const int my_tiles = (L.row_tiles - slot + L.cps - 1) / L.cps;   // tiles slot, slot + cps, ...
  const LstmSeqMaps& M = L.m[seq];

  if (warp == 0 && lane == 0) {
    if (HASX) tma_prefetch_desc(&M.x);
    tma_prefetch_desc(&M.h[0]);
    if (!SAVE) tma_prefetch_desc(&M.h[1]);
    tma_prefetch_desc(&M.whh);
    if (HASX) tma_prefetch_desc(&M.wih);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(w_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), LS_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 4) {
    const int et = threadIdx.x - 128;
    // gate order i, f, g, o per unit; sigmoid(x) = 0.5 + 0.5 tanh(x / 2): the input and output gates' and the forget
    // gate's pre-activations are halved on the way in (acc * 0.5 + bias / 2), so their bias is stored halved
    // (GRU: r and z are the sigmoid gates, the two halves of the candidate gate are not halved)
    const bool halved = GRU ? (et & 3) < 2 : (et & 3) != 2;
    if (et < 256) bias_s[et] = (L.bias[seq] ? __ldg(L.bias[seq] + rank * 256 + et) : 0.0f) * (halved ? 0.5f : 1.0f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();
  pdl_launch();
  ls_cluster_sync();   // every CTA of the cluster is up before the first step

  // ring position / accumulator use counters advance identically in the three roles
  int stage = 0;
  uint32_t phase = 0, cnt = 0;

  if (warp == 0 && lane == 0) {   // this CTA's 256 gate columns of W_hh and W_ih: resident for the whole sequence
    mbar_expect_tx(w_full, (uint32_t)NKB * LS_W_BYTES);
    for (int kb = 0; kb < KBH; ++kb) tma_load_3d(w_base + kb * LS_W_BYTES, &M.whh, 0, rank * 256, kb, w_full);
    if (HASX) tma_load_3d(w_base + KBH * LS_W_BYTES, &M.wih, 0, rank * 256, 0, w_full);
  }
  if (warp == 1 && lane == 0) {
    mbar_wait(w_full, 0u);
    tc_fence_after();
  }
  const uint32_t idesc = instr_desc(256, false, false);
  const int lq = warp & 3, cg = (warp - 4) >> 2;   // epilogue: TMEM lane quarter, column quarter (64 columns = 16 units)
  const int H = L.hidden;

#pragma unroll 1
  for (int t = 0; t < L.steps; ++t) {
    const int par = t & 1;
    const bool stamp = (L.dbg & 16) && blockIdx.x == 0 && lane == 0;
    const long long t_step = stamp ? clock64() : 0;
    if (warp == 0) {
      // ===== TMA producer: [h_{t-1} | x_t] of every tile of this cluster =====
      if (lane == 0) {
        // per tile: x_t first (it does not depend on the previous step: the first tile's x_t was requested before the
        // step barrier, at the end of the previous step), then the k-blocks of h_{t-1}
        auto load_x = [&](int tt, int i) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), LS_A_BYTES);
          tma_load_3d(ring_base + stage * LS_A_BYTES, &M.x, 0, (slot + i * L.cps) * 128, tt, full_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        };
        if (HASX && t == 0) load_x(0, 0);
        for (int i = 0; i < my_tiles; ++i) {
          const int row0 = (slot + i * L.cps) * 128;
          if (HASX && i > 0) load_x(t, i);
          for (int kb = 0; kb < KBH; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), LS_A_BYTES);
            if (SAVE) tma_load_3d(ring_base + stage * LS_A_BYTES, &M.h[0], kb * 64, row0, t, full_bar(stage));
            else tma_load_3d(ring_base + stage * LS_A_BYTES, &M.h[par], 0, row0, kb, full_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
        if (HASX && t + 1 < L.steps) load_x(t + 1, 0);
        if (stamp) g_ls_stamps[0] += clock64() - t_step;
      }
    } else if (warp == 1) {
      // ===== MMA issuer =====
      if (lane == 0) {
        for (int i = 0; i < my_tiles; ++i, ++cnt) {
          const int acc = (int)(cnt & 1u);
          mbar_wait(acc_empty(acc), ((cnt >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
          for (int kk = 0; kk < NKB; ++kk) {
            const int kb = !HASX ? kk : kk == 0 ? KBH : kk - 1;   // the producer's order: x_t, then the k-blocks of h_{t-1}
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = ring_base + stage * LS_A_BYTES, b_addr = w_base + kb * LS_W_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(d_tmem, smem_desc(a_addr + k * 32, 16, 1024), smem_desc(b_addr + k * 32, 16, 1024), idesc,
                          (kk > 0 || k > 0) ? 1u : 0u);
            tc_commit(empty_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          tc_commit(acc_full(acc));
        }
        if (stamp) g_ls_stamps[1] += clock64() - t_step;
      }
    } else if (warp >= 4) {
      // ===== epilogue: LSTM cell on this CTA's 64 hidden units (16 per warp: 64 accumulator columns) =====
      const long long BH = (long long)L.rows * L.hidden;
      __nv_bfloat16* h_next = SAVE ? reinterpret_cast<__nv_bfloat16*>(L.hbuf[seq][0]) + (long long)(t + 1) * BH
                                   : reinterpret_cast<__nv_bfloat16*>(L.hbuf[seq][par ^ 1]);
      const __nv_bfloat16* h_prev = SAVE ? reinterpret_cast<const __nv_bfloat16*>(L.hbuf[seq][0]) + (long long)t * BH
                                         : reinterpret_cast<const __nv_bfloat16*>(L.hbuf[seq][par]);
      // training mode: c_{t-1} is read from step t-1's slice (zeros at t = 0) and c_t goes to step t's
      const bool tape = SAVE && L.gates[seq] != nullptr;
      // (GRU: the fp32 hidden state stays in place in `cell` also in training mode; the tape holds h as bf16 in h_all)
      float* cellp = (tape && !GRU) ? L.c_all[seq] + (long long)t * BH : L.cell[seq];
      const float* cell_prev = (tape && !GRU) ? cellp - BH : cellp;
      __nv_bfloat16* gates_t = tape ? L.gates[seq] + (long long)t * BH * 4 : nullptr;
      const __nv_bfloat16* zin_t = L.zin[seq] != nullptr ? L.zin[seq] + (long long)t * BH * 4 : nullptr;
      const int* lens = L.lengths[seq];
      float* h32 = L.h_out[seq];
      for (int i = 0; i < my_tiles; ++i, ++cnt) {
        const int acc = (int)(cnt & 1u);
        const int tile = slot + i * L.cps;
        const int r = lq * 32 + lane, row = tile * 128 + r;
        const bool row_ok = row < L.rows;
        const bool ragged = (tile + 1) * 128 > L.rows;
        const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)acc * 256u + (uint32_t)(cg * 64);
        const int ubase = rank * 64 + cg * 16;   // first hidden unit of this thread's 16
        // pack_padded_sequence semantics: a window's state stops after its last valid step (src/encoders.py:140-152)
        const int len = (lens != nullptr && row_ok) ? __ldg(lens + row) : L.steps;
        const bool live = t < len;           // this step still belongs to the window
        const bool last_live = t == len - 1;  // its final hidden state is the one of this step
        const bool cell_io = row_ok && live && !(L.dbg & 8);
        // the previous cell state is fetched before the accumulator is ready
        float4 cprev[4];
#pragma unroll
        for (int g = 0; g < 4; ++g)
          cprev[g] = (cell_io && !(tape && !GRU && t == 0))
                         ? *reinterpret_cast<const float4*>(cell_prev + ls_cell_index(tile, r, ubase + 4 * g, H, ragged))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_wait(acc_full(acc), (cnt >> 1) & 1u);
        tc_fence_after();
        if (stamp && warp == 4 && i == my_tiles - 1) g_ls_stamps[2] += clock64() - t_step;
        uint32_t hp[8];   // this thread's 16 new hidden values as bf16 pairs: one 32-byte sector of the next A operand
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t a[16];
          uint32_t zw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          if (zin_t != nullptr && cell_io)   // the input's share of these 16 pre-activations
            ld_global_v8(zin_t + (long long)row * 4 * H + 4 * (ubase + 4 * g), zw);
          tmem_ld16_issue(taddr + (uint32_t)(16 * g), a);
          tmem_wait16(a);
          if (zin_t != nullptr) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              a[2 * e] = __float_as_uint(__uint_as_float(a[2 * e]) + __uint_as_float(zw[e] << 16));
              a[2 * e + 1] = __float_as_uint(__uint_as_float(a[2 * e + 1]) + __uint_as_float(zw[e] & 0xffff0000u));
            }
          }
          const float cp[4] = {cprev[g].x, cprev[g].y, cprev[g].z, cprev[g].w};
          float cn[4], hn[4], go[4];
          uint32_t gs[8];   // SAVE: the 16 gate activations of these 4 units as bf16 pairs (i,f | g,o per unit)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cg * 64 + 16 * g + 4 * j);
            if (GRU) {   // cp = h_{t-1} (fp32), cn = hn = h_t
              const float2 t_rz = ls_tanh2(fmaf(__uint_as_float(a[4 * j + 0]), 0.5f, b4.x), fmaf(__uint_as_float(a[4 * j + 1]), 0.5f, b4.y));
              const float gr = fmaf(0.5f, t_rz.x, 0.5f), gz = fmaf(0.5f, t_rz.y, 0.5f);
              const float nn = ls_tanh(fmaf(gr, __uint_as_float(a[4 * j + 3]) + b4.w, __uint_as_float(a[4 * j + 2]) + b4.z));
              cn[j] = fmaf(gz, cp[j] - nn, nn);
              hn[j] = cn[j];
              go[j] = 0.0f;
              if (tape) {   // (r, z, n, recurrent share of the candidate gate incl. its bias)
                __nv_bfloat162 s0 = __floats2bfloat162_rn(gr, gz), s1 = __floats2bfloat162_rn(nn, __uint_as_float(a[4 * j + 3]) + b4.w);
                gs[2 * j] = *reinterpret_cast<uint32_t*>(&s0);
                gs[2 * j + 1] = *reinterpret_cast<uint32_t*>(&s1);
              }
              continue;
            }
            const float2 t_if = ls_tanh2(fmaf(__uint_as_float(a[4 * j + 0]), 0.5f, b4.x), fmaf(__uint_as_float(a[4 * j + 1]), 0.5f, b4.y));
            const float2 t_go = ls_tanh2(__uint_as_float(a[4 * j + 2]) + b4.z, fmaf(__uint_as_float(a[4 * j + 3]), 0.5f, b4.w));
            const float gi = fmaf(0.5f, t_if.x, 0.5f), gf = fmaf(0.5f, t_if.y, 0.5f);
            go[j] = fmaf(0.5f, t_go.y, 0.5f);
            cn[j] = fmaf(gf, cp[j], gi * t_go.x);
            if (tape) {
              __nv_bfloat162 s0 = __floats2bfloat162_rn(gi, gf), s1 = __floats2bfloat162_rn(t_go.x, go[j]);
              gs[2 * j] = *reinterpret_cast<uint32_t*>(&s0);
              gs[2 * j + 1] = *reinterpret_cast<uint32_t*>(&s1);
            }
          }
#pragma unroll
          for (int j = 0; j < 4 && !GRU; j += 2) {
            const float2 tc = ls_tanh2(cn[j], cn[j + 1]);
            hn[j] = go[j] * tc.x;
            hn[j + 1] = go[j + 1] * tc.y;
          }
          const int u0 = ubase + 4 * g;
          if (cell_io)
            *reinterpret_cast<float4*>(cellp + ls_cell_index(tile, r, u0, H, ragged)) = make_float4(cn[0], cn[1], cn[2], cn[3]);
          if (tape && cell_io) st_global_v8(gates_t + (long long)row * 4 * H + 4 * u0, gs);
          __nv_bfloat162 p0 = __floats2bfloat162_rn(hn[0], hn[1]), p1 = __floats2bfloat162_rn(hn[2], hn[3]);
          hp[2 * g] = *reinterpret_cast<uint32_t*>(&p0);
          hp[2 * g + 1] = *reinterpret_cast<uint32_t*>(&p1);
          if (last_live && row_ok)
            *reinterpret_cast<float4*>(h32 + (long long)row * H + u0) = make_float4(hn[0], hn[1], hn[2], hn[3]);
        }
        tc_fence_before();
        if (row_ok) {   // k-block (ubase >> 6) = this CTA's rank of [H/64][B][64]
          const long long off = SAVE ? (long long)row * H + ubase
                                     : (long long)(ubase >> 6) * L.h_slice + (long long)row * 64 + (ubase & 63);
          if (live) {
            st_global_v8(h_next + off, hp);
          } else {   // finished window: carry its hidden state forward unchanged
            uint32_t carry[8];
            ld_global_v8(h_prev + off, carry);
            st_global_v8(h_next + off, carry);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(acc));
      }
      // h_t is read back by TMA (async proxy) in the next step, by every CTA of the cluster
      if (!(L.dbg & 1)) asm volatile("fence.proxy.async.global;" ::: "memory");
      if (!(L.dbg & 2)) __threadfence();
      if (stamp && warp == 4) g_ls_stamps[3] += clock64() - t_step;
    }
    __syncwarp();
    ls_cluster_sync();   // h_t of the cluster's tiles is complete and visible
    if (stamp && warp == 4) g_ls_stamps[4] += clock64() - t_step;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <typename... KArgs, typename... Args>
cudaError_t ls_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster,
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

}  // namespace

// true if the persistent kernel takes this shape (hidden 64 / 128 / 192 / 256: a cluster of hidden / 64 CTAs)
bool lstm_seq_eligible(int hidden, int n, long long batch, int sms) {
  if (hidden % 64 != 0 || hidden < 64 || hidden > 256) return false;
  const int cs = hidden / 64;
  return n >= 1 && n <= MSF_LSTM_MAX_SEQS && batch >= 1 && sms / cs >= n && !getenv("MSF_LSTM_STEPS");
}

int lstm_seq_launch(const msf_lstm_seq* seqs, int n, long long batch, int steps, int hidden, cudaStream_t st) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    MSF_CHECK_CUDA(cudaGetDevice(&dev));
    MSF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  MSF_REQUIRE(lstm_seq_eligible(hidden, n, batch, sms), "lstm_seq: shape not supported");
  LstmSeqLaunch L;
  memset(&L, 0, sizeof(L));
  const long long B = batch, N4 = 4LL * hidden, slice = B * 64;
  L.n = n; L.rows = (int)B; L.steps = steps; L.hidden = hidden; L.kbh = hidden / 64; L.cs = hidden / 64;
  L.row_tiles = (int)ceil_div(B, 128);
  L.h_slice = slice;
  { const char* e = getenv("MSF_LSTM_DBG"); L.dbg = e ? atoi(e) : 0; }
  int cps = (sms / L.cs) / n;
  if (cps > L.row_tiles) cps = L.row_tiles;
  L.cps = cps;
  int rc;
  const bool save = seqs[0].h_all != nullptr;   // every hidden state kept ([T+1][B][H]): lower layers, training mode
  const bool zin = seqs[0].z_in != nullptr;     // layers above the first: the input's share comes precomputed
  L.nkbx = zin ? 0 : 1;
  for (int i = 0; i < n; ++i) {
    const msf_lstm_seq& S = seqs[i];
    MSF_REQUIRE(S.w_hh && S.bias && S.h_out, "msf_lstm_forward: null pointer in sequence %d", i);
    MSF_REQUIRE((S.h_all != nullptr) == save && (S.z_in != nullptr) == zin,
                "msf_lstm_forward: the sequences of one call must use the same mode (h_all / z_in)");
    MSF_REQUIRE(S.gates == nullptr || (S.h_all && (S.cell_type == 1 ? S.cell != nullptr : S.c_all != nullptr)),
                "msf_lstm_forward: training mode needs h_all, gates and c_all (GRU: cell) (sequence %d)", i);
    if (zin) {
      L.zin[i] = static_cast<const __nv_bfloat16*>(S.z_in);
      L.m[i].x = L.m[i].wih = CUtensorMap{};
    } else {
      MSF_REQUIRE(S.x_bf16 && S.w_ih, "msf_lstm_forward: null pointer in sequence %d", i);
      if ((rc = tc_encode_map(&L.m[i].x, S.x_bf16, B, 64, 64, steps, slice, 64, 128))) return rc;
      if ((rc = tc_encode_map(&L.m[i].wih, S.w_ih, N4, 64, 64, 1, 0, 64, 256))) return rc;
    }
    if (save) {
      if ((rc = tc_encode_map(&L.m[i].h[0], S.h_all, B, hidden, hidden, steps + 1, B * hidden, 64, 128))) return rc;
      L.m[i].h[1] = L.m[i].h[0];
      L.hbuf[i][0] = S.h_all; L.hbuf[i][1] = S.h_all;
      L.gates[i] = static_cast<__nv_bfloat16*>(S.gates);
      L.c_all[i] = S.c_all;
      MSF_REQUIRE(S.gates != nullptr || S.cell != nullptr, "msf_lstm_forward: null cell buffer in sequence %d", i);
      L.cell[i] = S.cell;
    } else {
      MSF_REQUIRE(S.h_a && S.h_b && S.cell, "msf_lstm_forward: null pointer in sequence %d", i);
      if ((rc = tc_encode_map(&L.m[i].h[0], S.h_a, B, 64, 64, L.kbh, slice, 64, 128))) return rc;
      if ((rc = tc_encode_map(&L.m[i].h[1], S.h_b, B, 64, 64, L.kbh, slice, 64, 128))) return rc;
      L.hbuf[i][0] = S.h_a; L.hbuf[i][1] = S.h_b;
      L.cell[i] = S.cell;
    }
    if ((rc = tc_encode_map(&L.m[i].whh, S.w_hh, N4, 64, 64, L.kbh, N4 * 64, 64, 256))) return rc;
    L.bias[i] = S.bias;
    L.h_out[i] = S.h_out;
    L.lengths[i] = S.lengths;
  }
  const size_t fixed = 1024 + LS_BIAS_OFF + 256 * 4 + 64;
  const size_t weights = (size_t)(L.kbh + L.nkbx) * LS_W_BYTES;
  int stages = (int)((LS_SMEM_LIMIT - fixed - weights) / LS_A_BYTES);
  if (stages > LS_MAX_STAGES) stages = LS_MAX_STAGES;
  MSF_REQUIRE(stages >= 2, "lstm_seq: not enough shared memory for hidden %d", hidden);
  L.stages = stages;
  const size_t smem = fixed + weights + (size_t)stages * LS_A_BYTES;
  const int grid = n * cps * L.cs;
  if (prof_enabled()) prof_begin("LSTM sequence", 2.0 * (double)B * (hidden + 64 * L.nkbx) * 4.0 * hidden * steps * n, st);
  const bool gru = seqs[0].cell_type == 1;
  for (int i = 0; i < n; ++i) {
    MSF_REQUIRE(seqs[i].cell_type == seqs[0].cell_type && (seqs[i].cell_type == 0 || seqs[i].cell_type == 1),
                "msf_lstm_forward: cell_type 0 (LSTM) or 1 (GRU), the same for all sequences of a call");
  }
  auto kernel = gru ? (save ? lstm_seq_kernel<true, true> : lstm_seq_kernel<false, true>)
                    : (save ? lstm_seq_kernel<true, false> : lstm_seq_kernel<false, false>);
  MSF_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MSF_CHECK_CUDA(ls_launch(kernel, dim3(grid), dim3(LS_THREADS), smem, st, L.cs, L));
  MSF_LAUNCH_CHECK();
  prof_end(st);
  if (L.dbg & 16) {
    MSF_CHECK_CUDA(cudaStreamSynchronize(st));
    long long v[8];
    MSF_CHECK_CUDA(cudaMemcpyFromSymbol(v, g_ls_stamps, sizeof(v)));
    fprintf(stderr, "lstm_seq CTA 0, cycles per step: producer done %lld, MMAs issued %lld, accumulators complete %lld, "
                    "epilogue done %lld, barrier passed %lld (tiles per cluster %d, stages %d)\n",
            v[0] / steps, v[1] / steps, v[2] / steps, v[3] / steps, v[4] / steps, (int)ceil_div(L.row_tiles, cps), stages);
    memset(v, 0, sizeof(v));
    MSF_CHECK_CUDA(cudaMemcpyToSymbol(g_ls_stamps, v, sizeof(v)));
  }
  return MSF_OK;
}

}  // namespace msf
