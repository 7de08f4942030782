// Load-time kernel shared by the tensor-core denoiser engine and the image backbone (not on the hot path).
#pragma once
#include "common.cuh"

namespace pdb {

// LayerNorm folded into the following Linear:  LN(x) W^T + b = rstd (x Wf^T - mean colsum) + biasf  with
// Wf = gamma * W (column-wise), colsum_o = sum_k Wf[o][k], biasf_o = b_o + sum_k beta_k W[o][k].  One block per output row.
static __global__ void fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ gamma,
                               const float* __restrict__ beta, int K, float* __restrict__ Wf, float* __restrict__ colsum,
                               float* __restrict__ biasf) {
  const int o = blockIdx.x;
  float cs = 0.f, bs = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float w = W[(size_t)o * K + k];
    const float wf = w * gamma[k];
    Wf[(size_t)o * K + k] = wf;
    cs += wf;
    bs += w * beta[k];
  }
  __shared__ float red[2][4];
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = cs; red[1][threadIdx.x >> 5] = bs; }
  __syncthreads();
  if (threadIdx.x == 0) {
    colsum[o] = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    biasf[o] = bias[o] + red[1][0] + red[1][1] + red[1][2] + red[1][3];
  }
}

}  // namespace pdb
