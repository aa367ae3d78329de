// head_kernel: everything of HybridFusion between the aggregated modality tokens and their gradients,
// for one 128-window tile, in ONE kernel (src/fusion.py:410-419,429-479; src/train.py:185-186,310):
//
//   P0  gating scores (warp-shuffle dot products), masked softmax over the M modality tokens with the
//       reference's fallbacks, fused = sum_q w_q * agg_q            -> bf16, 128B-swizzled shared memory
//   G1  Hr   = drop(relu(fused W1^T + b1))      tcgen05.mma, A = that shared-memory block, W1 by TMA
//   G2  z    = Hr W2^T + b2                     (N = 32 tile) -> logits (fp32), softmax -> (conf, pred)
//   E2  CE with label smoothing: row loss, d logits (also summed into d b2), bf16 copy = A of G3
//   G3  dH1  = (dlog W2) * relu'(Hr) * drop     relu/dropout mask kept as 128 bits in registers
//   G4  dfused = dH1 W1
//   P5  backward of P0: d scores ds, dS_q = (w_q dfused + ds_q gw_q) * mask_q / cnt_q
//
// Forward-only mode (inference / the autograd forward) stops after G2.  The intermediate tiles never
// leave the SM on their way to the next GEMM; what the weight-gradient GEMM needs later (fused, Hr,
// dH1, dlog) is written once by TMA stores from the same shared-memory block.
#pragma once

#include <cuda.h>

#include "msf_common.cuh"

namespace msf {

struct HeadLaunch {
  CUtensorMap map_w1, map_w2, map_w2t, map_w1t;   // operand loads (bf16 compute arena)
  CUtensorMap map_agg;                            // aggregated tokens [M][rows][H], box 64 x 32 (P0 / P5 staging)
  CUtensorMap map_fused, map_hr, map_dh1;         // activation stores [rows][H], box 64 x tile_rows
  int train;              // 0: forward only, 1: forward + CE + backward
  int M, H, C, Cp;
  int rows, row_tiles, stages;
  int tile_rows;          // windows per CTA tile: 128, or 32 for small batches (set by head_launch)
  int store_acts;         // forward-only: write fused / Hr (operands of a later backward)
  const __nv_bfloat16* agg;  // [M][rows][H]
  const float* gate_w[MSF_MAX_MODALITIES];
  const float* gate_b[MSF_MAX_MODALITIES];
  const float* mask;      // (rows, M) or nullptr
  float* soft;            // (rows, M) softmax before the mask renormalisation (backward operand)
  float* w;               // (rows, M) fusion weights
  float* w_out;           // optional copy for the caller
  const float* b1;
  const float* b2;
  float* logits;          // (rows, C)
  float* conf;            // optional (forward only): max softmax probability
  long long* pred;        //                          first arg-max
  // ---- train ----
  const long long* labels;
  float smoothing, grad_scale;
  float* row_loss;        // (rows)
  float* loss_out;        // mean of row_loss (fixed-order reduction by the last CTA)
  __nv_bfloat16* dlog;    // (rows, Cp) bf16 d logits: operand of the classifier.3 weight gradient
  float* db2;             // += column sums of d logits
  double* grad_sq;        // optional: cleared here for the weight-gradient GEMM later in the pass (msf_fusion_call.grad_sq)
  __nv_bfloat16* dS;      // [M][rows][H]
  float* ds;              // (rows, M) d loss / d gating score (operand of the gating-layer gradients)
  float inv_cnt[MSF_MAX_MODALITIES];
  DropCfg drop;
};

bool head_eligible(int H, int M, int C);
// Windows per tile head_launch will use for this batch: the box height of the three store maps.
int head_tile_rows(long long rows);
// Fills row_tiles / stages and launches.  Tensor maps must already be encoded (tc_encode_map).
int head_launch(HeadLaunch& L, cudaStream_t stream, const char* label);
// clock64 stamps of the phases of CTA 0 in the last launch (debugging aid)
int head_debug_stamps(long long* out16);

}  // namespace msf
