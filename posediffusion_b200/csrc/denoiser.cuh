// Denoiser forward + DDPM posterior update as ONE persistent cooperative kernel per launch
// (models/denoiser.py:53-76, util/embedding.py:13-50, models/gaussian_diffuser.py:190-282).
//
// A launch runs diffusion steps t_hi .. t_lo back to back.  Per step the grid walks 43 stages separated by
// group barriers: embed+first, 8 x (LN1+QKV, attention, out-proj+residual, LN2+FF1+ReLU, FF2+residual),
// last0, tail (LayerNorm(128)+ReLU+Linear(128->9) fused with x0 / posterior mean / noise add).
// Loop-invariant work is hoisted: z-projection (385 of the 702 `_first` input columns) once per launch,
// the timestep-embedding MLP and its `_first` block once per weight load (table of 100 rows).
//
// This file is the exact-fp32 engine: CUDA-core FMAs, weights re-laid out k4-major ([K/4][O] float4) so that a
// warp reads 512 contiguous bytes per step while every lane owns one output feature; activations of a token tile
// are staged in shared memory and broadcast.  S = B*N tokens is tiny (20..160 per GPU): every stage is
// latency / weight-bandwidth bound.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "posediff_b200.h"

namespace pdb {

constexpr int kDM = 512;       // d_model            (cfgs/default.yaml:28)
constexpr int kHeads = 4;      // nhead              (:29)
constexpr int kHD = 128;       // head dim
constexpr int kFF = 1024;      // dim_feedforward    (:30)
constexpr int kLayers = 8;     // num_encoder_layers (:31)
constexpr int kHid = 128;      // mlp_hidden_dim     (denoiser.py:29)
constexpr int kZ = 384;
constexpr int kTEmb = 128;
constexpr int kPoseEmb = 189;  // 9 * (2*10 + 1)
constexpr int kPoseEmbPad = 256;  // 189 harmonic-pose columns zero-padded so that K/4 splits over 32 k-slices
constexpr int kFirstIn = 702;
constexpr int kT = PDB_NUM_TIMESTEPS;
constexpr int kDenThreads = 256;
constexpr int kDenWarps = kDenThreads / 32;
constexpr float kLnEps = 1e-5f;

struct LayerWeights {
  const float4* w_qkv;  // packed [512/4][1536]
  const float* b_qkv;
  const float4* w_out;  // [512/4][512]
  const float* b_out;
  const float4* w_ff1;  // [512/4][1024]
  const float* b_ff1;
  const float4* w_ff2;  // [1024/4][512]
  const float* b_ff2;
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};

struct DenoiserDev {  // device pointers into the packed weight arena
  const float4* w_first_x;  // [256/4][512]  (harmonic-pose columns 0..188 of _first.weight, zero padded)
  const float4* w_first_z;  // [384/4][512]  (columns 317..700)
  const float* w_first_pivot;  // [512]      (column 701)
  const float* b_first;        // [512]
  const float* tproj;          // [100][512]  = t_emb(t) @ _first.weight[:,189:317]^T
  LayerWeights layer[kLayers];
  const float4* w_last0;  // [512/4][128]
  const float* b_last0;
  const float *ln_last_g, *ln_last_b;
  const float* w_last3;  // [9][128] row-major (as in the checkpoint)
  const float* b_last3;
  const float* sched;    // [100][8]: a_t, b_t, c1_t, c2_t, sigma_t
};

struct DenoiserRun {
  int batch, frames, tokens;
  int t_hi, t_lo;          // steps t_hi, t_hi-1, ..., t_lo
  int guide_below;         // t < guide_below: leave the posterior mean in x (no noise; the GGS kernel follows)
  int compute_zproj;       // 1: (re)compute the z projection at the start of this launch
  float* x;                // [S,9] sampler state, updated in place
  const float* z;          // [S,384]
  const float* draws;      // [T+1,S,9] or null (no noise added)
  float* trail;            // [T+1,S,9] or null
  float* eps_out;          // [S,9] or null (network output of the LAST step run)
  float* x0_out;           // [S,9] or null
  float* mean_out;         // [S,9] or null
  // workspace
  float *zproj, *h, *qkv, *att, *ff, *u;
  unsigned* bar;           // zero on entry
  long long* dbg_clock;    // may be null: [ctas][8] cycle sums {barrier, tile load + LayerNorm, linear item, attention, tail, steps}
};

inline size_t denoiser_ws_floats(int tokens) {
  return (size_t)tokens * (kDM + kDM + 3 * kDM + kDM + kFF + kHid) + 64;
}

// ---------------------------------------------------------------------------------------------
// token-tile loaders (global -> shared), one warp per row
// ---------------------------------------------------------------------------------------------
// plain copy or LayerNorm of a [rows, K] slab; rows >= valid are zero-filled.  All global loads of the tile are
// issued before any is consumed (ONE L2 round trip per stage); the LayerNorm then runs out of shared memory.
template <int K, int TS>
__device__ __forceinline__ void load_rows(float* __restrict__ Xs, const float* __restrict__ src, int row0, int valid_end,
                                          const float* __restrict__ ln_g, const float* __restrict__ ln_b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int K4 = K / 4;
  constexpr int TOTAL = TS * K4;                    // float4 elements of the tile
  constexpr int PER = (TOTAL + kDenThreads - 1) / kDenThreads;
  float4 v[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int idx = threadIdx.x + i * kDenThreads;
    const int r = idx / K4, c = idx - r * K4;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx < TOTAL && row0 + r < valid_end) v[i] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * K) + c);
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int idx = threadIdx.x + i * kDenThreads;
    if (idx < TOTAL) reinterpret_cast<float4*>(Xs)[idx] = v[i];
  }
  if (ln_g) {
    __syncthreads();
    constexpr int PL = K / 128;  // float4 per lane per row
    for (int r = warp; r < TS; r += kDenWarps) {
      if (row0 + r >= valid_end) continue;
      float4* row = reinterpret_cast<float4*>(Xs + (size_t)r * K);
      float4 x[PL];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < PL; ++i) {
        x[i] = row[lane + 32 * i];
        sum += x[i].x + x[i].y + x[i].z + x[i].w;
      }
      const float mean = warp_sum(sum) * (1.0f / K);
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < PL; ++i) {
        x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
        sq += x[i].x * x[i].x + x[i].y * x[i].y + x[i].z * x[i].z + x[i].w * x[i].w;
      }
      const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / K) + kLnEps);
#pragma unroll
      for (int i = 0; i < PL; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(ln_g) + lane + 32 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(ln_b) + lane + 32 * i);
        x[i].x = x[i].x * rstd * g.x + b.x;
        x[i].y = x[i].y * rstd * g.y + b.y;
        x[i].z = x[i].z * rstd * g.z + b.z;
        x[i].w = x[i].w * rstd * g.w + b.w;
        row[lane + 32 * i] = x[i];
      }
    }
  }
}

// harmonic pose embedding (pytorch3d HarmonicEmbedding, n=10, logspace, append_input; embedding.py:44):
// [sin(x_c 2^k) (c-major) | cos(same) | x | 0 0 0]
__device__ __forceinline__ void load_pose_embed(float* __restrict__ Xs, const float* __restrict__ x, int row0, int rows,
                                                int valid_end) {
  for (int i = threadIdx.x; i < rows * kPoseEmbPad; i += kDenThreads) {
    const int r = i / kPoseEmbPad, col = i - r * kPoseEmbPad;
    const int s = row0 + r;
    float v = 0.f;
    if (s < valid_end && col < kPoseEmb) {
      if (col < 180) {
        const int j = col < 90 ? col : col - 90;
        const int c = j / 10, k = j - c * 10;
        const float arg = __ldcg(x + s * 9 + c) * (float)(1 << k);
        v = col < 90 ? sinf(arg) : cosf(arg);
      } else {
        v = __ldcg(x + s * 9 + (col - 180));
      }
    }
    Xs[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// One linear work item: 32 output features x one token tile.  Xs = staged activations [TS][K].
// ---------------------------------------------------------------------------------------------
enum : int { kEpiNone = 0, kEpiRelu = 1, kEpiSilu = 2 };

// Work item of a linear stage: kFPI = 8 output features x one token tile, so that even the narrow stages
// (512 outputs -> 64 items) spread over many SMs.  Inside the CTA the reduction dimension is cut into 32 k-slices:
// lane = (ks = lane>>3, f = lane&7), slice = warp*4 + ks owns the k4-rows  i*32 + slice.  A warp-load of weights
// touches 4 k4-rows x 128 contiguous bytes; the 4 k-slices of a warp read 64 contiguous bytes of the staged
// activations (broadcast to the 8 feature lanes).
constexpr int kFPI = 8;
constexpr int kSlices = kDenWarps * 4;

template <int K>
struct LinW {
  static constexpr int K4 = K / 4;
  static constexpr int PER = K4 / kSlices;
  static_assert(K4 % kSlices == 0 && PER >= 1 && PER <= 16, "K layout");
  float4 w[PER];
};

// Issue the weight loads of one work item.  Weights never depend on activations, so this is called BEFORE the group
// barrier that precedes the stage: the L2 latency of the weight stream hides behind the barrier.
template <int K>
__device__ __forceinline__ void linear_prefetch(LinW<K>& W, const float4* __restrict__ Wp, int O, int o0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slice = warp * 4 + (lane >> 3);
  const float4* wp = Wp + (size_t)slice * O + o0 + (lane & 7);
#pragma unroll
  for (int i = 0; i < LinW<K>::PER; ++i) W.w[i] = __ldg(wp + (size_t)i * kSlices * O);
}

template <int TS, int K>
__device__ __forceinline__ void linear_item(const LinW<K>& W, const float* __restrict__ Xs, float* __restrict__ red, int o0,
                                            const float* __restrict__ bias, const float* __restrict__ add1, int ld1,
                                            const float* __restrict__ add2, const float* __restrict__ add_row_scaled,
                                            const float* __restrict__ row_scale, float* __restrict__ Y, int ldy, int row0,
                                            int valid_end, int epi) {
  static_assert(TS % 4 == 0 && TS <= 32, "token tile");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int PER = LinW<K>::PER;
  const int ks = lane >> 3, f = lane & 7;
  // epilogue operands of this thread's output (token s_out, feature f_out) are requested before the FMA loop
  const int s_out = threadIdx.x >> 3, f_out = threadIdx.x & 7;
  const bool has_out = s_out < TS && (row0 + s_out) < valid_end;
  float epi_add = 0.f;
  if (has_out) {
    const int row = row0 + s_out, o = o0 + f_out;
    if (bias) epi_add += __ldg(bias + o);
    if (add1) epi_add += __ldcg(add1 + (size_t)row * ld1 + o);
    if (add2) epi_add += __ldg(add2 + o);
    if (add_row_scaled) epi_add += __ldg(add_row_scaled + o) * row_scale[s_out];
  }
  float acc[TS];
#pragma unroll
  for (int s = 0; s < TS; ++s) acc[s] = 0.f;
  const float* xs = Xs + (warp * 4 + ks) * 4;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
#pragma unroll
    for (int s = 0; s < TS; ++s) {
      const float4 xv = *reinterpret_cast<const float4*>(xs + s * K + i * kSlices * 4);
      acc[s] = fmaf(W.w[i].x, xv.x, fmaf(W.w[i].y, xv.y, fmaf(W.w[i].z, xv.z, fmaf(W.w[i].w, xv.w, acc[s]))));
    }
  }
  // reduce over the 4 k-slices of the warp: after two exchange steps lane (ks, f) owns tokens ks*TS/4 .. +TS/4-1
  constexpr int Q = TS / 4;
  {
    const bool hi = ks & 2;
#pragma unroll
    for (int s = 0; s < TS / 2; ++s) {
      const float send = hi ? acc[s] : acc[s + TS / 2];
      const float keep = hi ? acc[s + TS / 2] : acc[s];
      acc[s] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = ks & 1;
#pragma unroll
    for (int s = 0; s < Q; ++s) {
      const float send = hi ? acc[s] : acc[s + Q];
      const float keep = hi ? acc[s + Q] : acc[s];
      acc[s] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
#pragma unroll
  for (int s = 0; s < Q; ++s) red[(warp * TS + ks * Q + s) * kFPI + f] = acc[s];
  __syncthreads();
  if (s_out < TS) {
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < kDenWarps; ++wv) v += red[(wv * TS + s_out) * kFPI + f_out];
    if (has_out) {
      v += epi_add;
      if (epi == kEpiRelu) v = fmaxf(v, 0.f);
      else if (epi == kEpiSilu) v = v / (1.0f + expf(-v));
      Y[(size_t)(row0 + s_out) * ldy + o0 + f_out] = v;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Self-attention for one (sequence, head): N <= 128 keys, head dim 128, fp32.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void attention_item(float* __restrict__ smem, const float* __restrict__ qkv,
                                               float* __restrict__ att, int seq, int head, int chunk, int N) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int KP = kHD + 4;  // padded key rows (float4 aligned): lane = key index reads hit distinct banks
  float* Ks = smem;                        // [N][132]
  float* Vs = Ks + N * KP;                 // [N][128]
  float* Qs = Vs + N * kHD;                // [warps][128]
  float* Ps = Qs + kDenWarps * kHD;        // [warps][128]
  const float* base = qkv + (size_t)seq * N * (3 * kDM) + head * kHD;
  const int i = chunk * kDenWarps + warp;  // this warp's query row
  float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < N) qv = __ldcg(reinterpret_cast<const float4*>(base + (size_t)i * 3 * kDM) + lane);
  for (int e = threadIdx.x; e < N * (kHD / 4); e += kDenThreads) {
    const int j = e / (kHD / 4), d4 = e - j * (kHD / 4);
    const float4 kv = __ldcg(reinterpret_cast<const float4*>(base + (size_t)j * 3 * kDM + kDM) + d4);
    const float4 vv = __ldcg(reinterpret_cast<const float4*>(base + (size_t)j * 3 * kDM + 2 * kDM) + d4);
    *reinterpret_cast<float4*>(Ks + j * KP + d4 * 4) = kv;
    *reinterpret_cast<float4*>(Vs + j * kHD + d4 * 4) = vv;
  }
  const float scaling = 0.08838834764831845f;  // 1/sqrt(128): q is scaled before QK^T (torch MHA)
  *reinterpret_cast<float4*>(Qs + warp * kHD + lane * 4) = make_float4(qv.x * scaling, qv.y * scaling, qv.z * scaling, qv.w * scaling);
  __syncthreads();
  if (i < N) {
    float sc[4];  // up to 128 keys: 4 passes of 32
    float mx = -INFINITY;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int j = p * 32 + lane;
      float dot = -INFINITY;
      if (p * 32 < N && j < N) {
        const float4* kr = reinterpret_cast<const float4*>(Ks + j * KP);
        const float4* qr = reinterpret_cast<const float4*>(Qs + warp * kHD);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll 8
        for (int d = 0; d < kHD / 4; ++d) {
          const float4 a = qr[d], b = kr[d];
          d0 = fmaf(a.x, b.x, d0); d1 = fmaf(a.y, b.y, d1); d2 = fmaf(a.z, b.z, d2); d3 = fmaf(a.w, b.w, d3);
        }
        dot = (d0 + d1) + (d2 + d3);
      }
      sc[p] = dot;
      mx = fmaxf(mx, dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int j = p * 32 + lane;
      const float e = (j < N) ? expf(sc[p] - mx) : 0.f;
      sc[p] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int j = p * 32 + lane;
      if (j < N) Ps[warp * kHD + j] = sc[p] * inv;
    }
    __syncwarp();
    float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < N; ++j) {
      const float pj = Ps[warp * kHD + j];
      const float4 vv = *reinterpret_cast<const float4*>(Vs + j * kHD + lane * 4);
      o4.x = fmaf(pj, vv.x, o4.x); o4.y = fmaf(pj, vv.y, o4.y); o4.z = fmaf(pj, vv.z, o4.z); o4.w = fmaf(pj, vv.w, o4.w);
    }
    *reinterpret_cast<float4*>(att + ((size_t)seq * N + i) * kDM + head * kHD + lane * 4) = o4;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Tail: per token LayerNorm(128) + ReLU + Linear(128->9), then the DDPM arithmetic.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tail_token(const DenoiserDev& W, const DenoiserRun& R, int s, int t, bool last_step) {
  const int lane = threadIdx.x & 31;
  float4 v = __ldcg(reinterpret_cast<const float4*>(R.u + (size_t)s * kHid) + lane);
  const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / kHid);
  v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
  const float rstd = 1.0f / sqrtf(warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w) * (1.0f / kHid) + kLnEps);
  const float4 g = __ldg(reinterpret_cast<const float4*>(W.ln_last_g) + lane);
  const float4 b = __ldg(reinterpret_cast<const float4*>(W.ln_last_b) + lane);
  v.x = fmaxf(v.x * rstd * g.x + b.x, 0.f);
  v.y = fmaxf(v.y * rstd * g.y + b.y, 0.f);
  v.z = fmaxf(v.z * rstd * g.z + b.z, 0.f);
  v.w = fmaxf(v.w * rstd * g.w + b.w, 0.f);
  float mine = 0.f;
#pragma unroll
  for (int c = 0; c < kTargetDim; ++c) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(W.w_last3 + c * kHid) + lane);
    const float dot = warp_sum(w.x * v.x + w.y * v.y + w.z * v.z + w.w * v.w);
    if (lane == c) mine = dot + __ldg(W.b_last3 + c);
  }
  if (lane < kTargetDim) {
    const float* sc = W.sched + t * 8;
    const size_t e = (size_t)s * kTargetDim + lane;
    const float xt = __ldcg(R.x + e);
    const float x0 = sc[0] * xt - sc[1] * mine;             // predict_start_from_noise (:190-194)
    const float mu = sc[2] * x0 + sc[3] * xt;               // q_posterior mean (:201-205)
    float out = mu;
    const int k = (kT - 1) - t;                              // loop iteration index -> draw slot 1 + k
    if (t >= R.guide_below && t > 0 && R.draws) out = mu + sc[4] * __ldg(R.draws + (size_t)(1 + k) * R.tokens * kTargetDim + e);
    R.x[e] = out;
    if (R.trail && t >= R.guide_below) R.trail[(size_t)(1 + k) * R.tokens * kTargetDim + e] = out;
    if (last_step) {
      if (R.eps_out) R.eps_out[e] = mine;
      if (R.x0_out) R.x0_out[e] = x0;
      if (R.mean_out) R.mean_out[e] = mu;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The persistent kernel
// ---------------------------------------------------------------------------------------------
template <int TS>
__global__ void __launch_bounds__(kDenThreads, 1)
denoiser_kernel(const __grid_constant__ DenoiserDev W, const __grid_constant__ DenoiserRun R) {
#ifdef PDB_EMU  // tests/host/cuda_emu.h (CPU emulation of this kernel, test harness only): dynamic shared memory of this CTA
  float* const smem = reinterpret_cast<float*>(emu::g_cta->smem);
#else
  extern __shared__ __align__(16) float smem[];
#endif
  float* Xs = smem;  // [TS][K<=1024] or attention scratch
  __shared__ float pivot[32];
  const int S = R.tokens;
  const int tiles = (S + TS - 1) / TS;
  const int G = gridDim.x;
  unsigned bar_count = 0;
  // Group barrier: release by one thread after the block barrier (cumulative over the CTA's writes), acquire polls.
  long long clk[5] = {0, 0, 0, 0, 0};  // stage timing probe (thread 0, only with R.dbg_clock)
  const bool probe = R.dbg_clock != nullptr && threadIdx.x == 0;
  auto barrier = [&]() {
    ++bar_count;
    long long c0 = 0;
    if (probe) c0 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) {
      red_release_add_u32(R.bar, 1u);
      while (ld_acquire_u32(R.bar) < bar_count * (unsigned)G) {
      }
    }
    __syncthreads();
    if (probe) clk[0] += clock64() - c0;
  };
  // One linear stage: Y = epi(X' W^T + ...), work items = (token tile, 32-feature group).  The weights of this CTA's
  // first item are requested BEFORE the barrier that closes the previous stage (`need_barrier`), the activations after.
  auto linear_stage = [&](auto ktag, const float4* Wp, int O, auto&& load_x, const float* bias, const float* add1, int ld1,
                          const float* add2, const float* add_rs, const float* rs, float* Y, int ldy, int epi,
                          bool need_barrier) {
    constexpr int K = decltype(ktag)::value;
    const int groups = O / kFPI;
    const int n_items = tiles * groups;
    LinW<K> wreg;
    int item = blockIdx.x;
    if (item < n_items) linear_prefetch<K>(wreg, Wp, O, (item % groups) * kFPI);
    if (need_barrier) barrier();
    float* red = Xs + TS * K;
    int staged_tile = -1;  // the activations of a token tile are staged once and reused by this CTA's later items of the stage
    for (; item < n_items; item += G) {
      const int tt = item / groups, fg = item - tt * groups;
      if (item != (int)blockIdx.x) linear_prefetch<K>(wreg, Wp, O, fg * kFPI);
      long long c0 = 0, c1 = 0;
      if (probe) c0 = clock64();
      if (tt != staged_tile) {
        load_x(tt);
        staged_tile = tt;
      }
      __syncthreads();
      if (probe) c1 = clock64();
      linear_item<TS, K>(wreg, Xs, red, fg * kFPI, bias, add1, ld1, add2, add_rs, rs, Y, ldy, tt * TS, S, epi);
      if (probe) {
        clk[1] += c1 - c0;
        clk[2] += clock64() - c1;
      }
    }
  };
  bool pending = false;  // a stage has written global activations that the next stage must wait for
  // ---- loop-invariant: zproj = z @ Wz^T + b_first + pivot * w_pivot   (denoiser.py:62-70) ----
  if (R.compute_zproj) {
    linear_stage(std::integral_constant<int, kZ>{}, W.w_first_z, kDM,
                 [&](int tt) {
                   load_rows<kZ, TS>(Xs, R.z, tt * TS, S, nullptr, nullptr);
                   if (threadIdx.x < TS) pivot[threadIdx.x] = ((tt * TS + threadIdx.x) % R.frames == 0) ? 1.f : 0.f;
                 },
                 W.b_first, nullptr, 0, nullptr, W.w_first_pivot, pivot, R.zproj, kDM, kEpiNone, false);
    pending = true;
  }
  for (int t = R.t_hi; t >= R.t_lo; --t) {
    // ---- embed + first ----
    linear_stage(std::integral_constant<int, kPoseEmbPad>{}, W.w_first_x, kDM,
                 [&](int tt) { load_pose_embed(Xs, R.x, tt * TS, TS, S); }, nullptr, R.zproj, kDM, W.tproj + t * kDM, nullptr,
                 nullptr, R.h, kDM, kEpiNone, pending);
    pending = true;
    for (int l = 0; l < kLayers; ++l) {
      const LayerWeights& L = W.layer[l];
      // LN1 + QKV projection
      linear_stage(std::integral_constant<int, kDM>{}, L.w_qkv, 3 * kDM,
                   [&](int tt) { load_rows<kDM, TS>(Xs, R.h, tt * TS, S, L.ln1_g, L.ln1_b); }, L.b_qkv, nullptr, 0, nullptr,
                   nullptr, nullptr, R.qkv, 3 * kDM, kEpiNone, true);
      // attention per (sequence, head)
      barrier();
      {
        long long c0 = 0;
        if (probe) c0 = clock64();
        const int chunks = (R.frames + kDenWarps - 1) / kDenWarps;
        for (int item = blockIdx.x; item < R.batch * kHeads * chunks; item += G)
          attention_item(Xs, R.qkv, R.att, item / (kHeads * chunks), (item / chunks) % kHeads, item % chunks, R.frames);
        if (probe) clk[3] += clock64() - c0;
      }
      // out-proj + residual (in place on h: each element is read and written by the same thread)
      linear_stage(std::integral_constant<int, kDM>{}, L.w_out, kDM,
                   [&](int tt) { load_rows<kDM, TS>(Xs, R.att, tt * TS, S, nullptr, nullptr); }, L.b_out, R.h, kDM, nullptr,
                   nullptr, nullptr, R.h, kDM, kEpiNone, true);
      // LN2 + FF1 + ReLU
      linear_stage(std::integral_constant<int, kDM>{}, L.w_ff1, kFF,
                   [&](int tt) { load_rows<kDM, TS>(Xs, R.h, tt * TS, S, L.ln2_g, L.ln2_b); }, L.b_ff1, nullptr, 0, nullptr,
                   nullptr, nullptr, R.ff, kFF, kEpiRelu, true);
      // FF2 + residual
      linear_stage(std::integral_constant<int, kFF>{}, L.w_ff2, kDM,
                   [&](int tt) { load_rows<kFF, TS>(Xs, R.ff, tt * TS, S, nullptr, nullptr); }, L.b_ff2, R.h, kDM, nullptr,
                   nullptr, nullptr, R.h, kDM, kEpiNone, true);
    }
    // last0: Linear(512 -> 128)
    linear_stage(std::integral_constant<int, kDM>{}, W.w_last0, kHid,
                 [&](int tt) { load_rows<kDM, TS>(Xs, R.h, tt * TS, S, nullptr, nullptr); }, W.b_last0, nullptr, 0, nullptr,
                 nullptr, nullptr, R.u, kHid, kEpiNone, true);
    // tail: one warp per token
    barrier();
    {
      long long c0 = 0;
      if (probe) c0 = clock64();
      const int warp_global = blockIdx.x * kDenWarps + (threadIdx.x >> 5);
      for (int s = warp_global; s < S; s += G * kDenWarps) tail_token(W, R, s, t, t == R.t_lo);
      if (probe) clk[4] += clock64() - c0;
    }
    pending = true;
  }
  if (probe) {
    long long* c = R.dbg_clock + (size_t)blockIdx.x * 8;
    for (int k = 0; k < 5; ++k) c[k] += clk[k];
    c[5] += R.t_hi - R.t_lo + 1;
  }
}

inline size_t denoiser_smem_bytes(int TS, int frames) {
  size_t lin = (size_t)TS * kFF + (size_t)kDenWarps * TS * kFPI;                      // X tile + reduction
  size_t att = (size_t)frames * (kHD + 4) + (size_t)frames * kHD + 2 * kDenWarps * kHD;  // K, V, Q, P
  return sizeof(float) * (lin > att ? lin : att) + 256;
}

}  // namespace pdb
