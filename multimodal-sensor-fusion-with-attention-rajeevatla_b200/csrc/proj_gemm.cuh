// proj_kernel: the input side of HybridFusion.forward (src/fusion.py:364-374) in one launch per step:
//
//   xt_m = bf16(drop0(LN_m(x_m) * mask_m))      (LN_m: the optional LayerNorm train.py puts between encoder and
//                                               fusion; row mean / variance while the fp32 tile is in shared memory)
//                                               read as fp32 rows by the worker warps, written once into the
//                                               128B-swizzled A block (and from there to global by TMA: the
//                                               projection weight gradient needs it later)
//   P_m  = drop1(relu(xt_m Wp_m^T + bp_m))      tcgen05.mma, A = that block, Wp_m by TMA, accumulator in TMEM
//
// replacing the elementwise prep kernel + the grouped-GEMM launch.  One CTA per (128-window tile, modality).
#pragma once

#include <cuda.h>

#include "msf_common.cuh"

namespace msf {

struct ProjZeroRange {
  long long begin, count, stride;
  int batch;
};
constexpr int PROJ_MAX_ZERO = 24;

struct ProjLaunch {
  CUtensorMap map_w[MSF_MAX_MODALITIES];    // Wp_m [H][D_m], box 64 x H
  CUtensorMap map_xt[MSF_MAX_MODALITIES];   // xt_m [rows][D_m], store box 64 x 128
  CUtensorMap map_p;                        // P [M][rows][H], store box 64 x 128
  int M, H, rows, row_tiles, items;
  int w_area, a_area, x_area;               // shared-memory carve-up in bytes (set by proj_launch)
  int D[MSF_MAX_MODALITIES];
  const float* x[MSF_MAX_MODALITIES];       // (rows, D_m) fp32, or bf16 rows when x_bf16
  int x_bf16;
  const float* bias[MSF_MAX_MODALITIES];    // (H)
  int n_active;                             // modalities that get items (0: all M); absent ones of a uniform-mask
  short active[MSF_MAX_MODALITIES];         // inference pass are left out
  const float* mask;                        // (rows, M) or nullptr
  // optional per-modality LayerNorm of the input rows (src/train.py:170-171,267-268), applied in the input phase
  // before mask and dropout: y = (x - mean) * rstd * ln_w + ln_b; both nullptr = none
  const float* ln_w[MSF_MAX_MODALITIES];
  const float* ln_b[MSF_MAX_MODALITIES];
  float ln_eps;
  DropCfg drop;
  // gradient slots to clear at the start of a train pass (first kernel of the step); see fusion_bf16.cu
  ProjZeroRange zero[PROJ_MAX_ZERO];
  int nzero;
  int zero_ctas;            // CTAs at the end of the grid that only clear the ranges below (set by proj_launch)
  float* zero_base;
};

bool proj_eligible(int H, int M, const int* D);
int proj_launch(ProjLaunch& L, cudaStream_t stream, const char* label);

}  // namespace msf
