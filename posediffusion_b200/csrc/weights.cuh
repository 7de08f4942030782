// Host-side view of the loaded denoiser weights (both engines).
#pragma once
#include "denoiser.cuh"

namespace pdb {

struct TcLayer {  // row-major [O, K] fp32 operands of the tensor-core engine (LayerNorm folded into QKV / FF1)
  const float *wqkv, *colsum_qkv, *bias_qkv, *wout, *bout, *wff1, *colsum_ff1, *bias_ff1, *wff2, *bff2;
};
struct TcWeights {
  const float *wx, *wz, *w_pivot, *b_first, *tproj, *wlast0, *blast0;
  TcLayer layer[kLayers];
};
struct DenoiserWeights {
  float* arena = nullptr;
  size_t arena_floats = 0;
  DenoiserDev dev = {};
  float* raw = nullptr;   // checkpoint tensors as loaded (row-major), kept for the tensor-core engine
  float* tc_arena = nullptr;
  TcWeights tc = {};
};

}  // namespace pdb
