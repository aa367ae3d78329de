// Chained pair GEMMs of HybridFusion on the tensor cores: for one 128-window tile and one
// "outer" modality the kernel runs, for every partner modality,
//
//     T   = A1[inner] . W1[pair]^T        (tcgen05.mma into TMEM)
//     T'  = epilogue1(T)  -> bf16 -> swizzled shared memory (+ TMA store to global when training)
//     ACC += T' . W2[pair]^T              (A operand read straight from that shared memory)
//
// and finishes with one epilogue over ACC.  Forward (src/fusion.py:383-408 with the q_len = k_len = 1
// attention of src/attention.py:104-140): outer = query q, inner = key k, W1 = value_proj, epilogue1 =
// (+bias) * gate, W2 = out_proj, final = (ACC + sum bias_o + P_q) / count * mask_q -> aggregated[q].
// Backward: outer = key k, inner = query q, A1 = d aggregated[q], W1 = out_proj^T, epilogue1 = * gate,
// W2 = value_proj^T, final = (ACC + dS_k) * relu'(P_k) -> dZ_k.
// The intermediate never makes a round trip through L2/HBM on its way to the second GEMM.
#pragma once

#include <cuda.h>

#include "msf_common.cuh"

namespace msf {

constexpr int CHAIN_MAX_INNER = MSF_MAX_MODALITIES - 1;
constexpr int CHAIN_MAX_PAIRS = MSF_MAX_MODALITIES * (MSF_MAX_MODALITIES - 1);

struct ChainOuter {
  int n;                           // partner modalities with a pair module
  short inner[CHAIN_MAX_INNER];    // z index of A1 for each partner
  short pair[CHAIN_MAX_INNER];     // pair index (z of W1 / W2 / out1, gate slot)
  short sub[CHAIN_MAX_INNER];      // dropout sub-stream of the attention gate (q*M + k)
  short mask_col[CHAIN_MAX_INNER]; // mask column deciding the gate (forward: the key modality)
  // forward with a batch-uniform mask: pairs whose key modality is absent everywhere have gate 0, i.e. they
  // contribute exactly their out_proj bias — no GEMMs, only this list for the bias sum
  short nb;
  short bias_only[CHAIN_MAX_INNER];
};

struct ChainLaunch {
  CUtensorMap map_a1, map_w1, map_w2, map_out1, map_out;
  CUtensorMap map_aux2;   // backward: the P stack (ReLU masks of the final epilogue), box 64 x 128 like map_a1; the
                          // aux tile itself (forward P_q, backward dS_k) is read through map_a1 (same tensor)
  int mode;            // 0 forward, 1 backward
  int M, H, heads, head_dim;
  int head_shift;      // log2(head_dim) when it is a power of two, else -1 (set by chain_launch)
  int rows, row_tiles, items;
  int n_active;                    // outer modalities that get items (0: all M); absent queries of a uniform-mask
  short active[MSF_MAX_MODALITIES];  // inference pass are left out: the head never reads their aggregated token
  int store1;          // write epilogue1's result to global memory (needed by the backward pass)
  int stages;
  int cluster;         // unused (the cluster variant of the kernel is gone); kept so that the launch record keeps its layout
  ChainOuter outer[MSF_MAX_MODALITIES];
  const float* bias1[CHAIN_MAX_PAIRS];   // forward: value_proj bias per pair
  const float* bias2[CHAIN_MAX_PAIRS];   // forward: out_proj bias per pair (summed in the final epilogue)
  float* gate_out;                       // forward: [pairs][rows][heads]
  const float* gate_in;                  // backward
  const float* mask;                     // (rows, M) or nullptr
  const __nv_bfloat16* aux;              // forward: P stack; backward: dS stack      [M][rows][H]
  const __nv_bfloat16* aux2;             // backward: P stack
  float inv_cnt[MSF_MAX_MODALITIES];     // forward: 1 / (1 + present pairs of q)
  float scale;                           // backward: dropout scale of the projection site
  DropCfg drop;
};

bool chain_eligible(int H, int M);
// Box height of map_w1 / map_w2 (rows of a weight block one TMA load brings in) for the kernel chain_launch will
// pick for this shape: H / 2 (map_w1) for chain2_kernel,
// H for chain_kernel.
int chain_w1_box_rows(int H, int M, int heads, long long rows);
int chain_w2_box_rows(int H, int M, int heads, long long rows);
// Fills stages / items / row_tiles and launches.  Tensor maps must already be encoded (tc_encode_map).
int chain_launch(ChainLaunch& L, cudaStream_t stream, const char* label);

}  // namespace msf
