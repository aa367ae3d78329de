// Layout of the bf16 compute arena of the tensor-core HybridFusion path (element offsets).
#pragma once

#include "msf_common.cuh"

namespace msf {

typedef __nv_bfloat16 bf16;

struct ArenaBf16 {
  size_t wp[MSF_MAX_MODALITIES], wpT[MSF_MAX_MODALITIES];  // [H][D_m], [D_m][H]
  size_t wv, wo, wvT, woT;                                 // [pairs][H][H]
  size_t w1, w1T;                                          // [H][H]
  size_t w2;                                               // [C][H]
  size_t w2T;                                              // [H][Cp]  (Cp = C rounded up to 8, zero padded)
  size_t total;
  int Cp;
};

static ArenaBf16 arena_layout(const Layout& L) {
  ArenaBf16 a;
  size_t off = 0;
  auto take = [&](size_t n) {
    const size_t o = off;
    off += align_up(n, 128);  // 256-byte aligned tensors (TMA needs 16)
    return o;
  };
  const size_t H = L.H;
  a.Cp = (int)align_up(L.C, 8);
  for (int m = 0; m < L.M; ++m) a.wp[m] = take(H * L.D[m]);
  for (int m = 0; m < L.M; ++m) a.wpT[m] = take(H * L.D[m]);
  const size_t pairs = L.num_pairs();
  a.wv = take(pairs * H * H);
  a.wo = take(pairs * H * H);
  a.wvT = take(pairs * H * H);
  a.woT = take(pairs * H * H);
  a.w1 = take(H * H);
  a.w1T = take(H * H);
  a.w2 = take((size_t)L.C * H);
  a.w2T = take(H * a.Cp);
  a.total = off;
  return a;
}


}  // namespace msf
