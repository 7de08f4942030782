// Post-loop geometry (widened row): pose encoding -> cameras, and the pairwise relative-pose errors the
// reference's evaluation computes from them.  Tiny, latency-bound work: one thread per camera / per pair, no staging.
#include <cmath>

#include "align.cuh"
#include "context.cuh"

using namespace pdb;

namespace {

// pose_encoding_to_camera, "absT_quaR_logFL" (util/camera_transform.py:64-105): T = enc[:3]; R = quaternion_to_matrix(enc[3:7])
// (pytorch3d: real part first, two_s = 2 / |q|^2, no normalisation pass); focal = clamp(exp(enc[7:9] + bias), min, max).
__global__ void pose_to_camera_kernel(const float* __restrict__ pose, int count, float bias, float fmin, float fmax,
                                      float* __restrict__ R, float* __restrict__ T, float* __restrict__ F) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float* e = pose + (size_t)i * 9;
  const float w = e[3], x = e[4], y = e[5], z = e[6];
  const float two_s = 2.0f / (w * w + x * x + y * y + z * z);
  float* r = R + (size_t)i * 9;
  r[0] = 1.f - two_s * (y * y + z * z);
  r[1] = two_s * (x * y - z * w);
  r[2] = two_s * (x * z + y * w);
  r[3] = two_s * (x * y + z * w);
  r[4] = 1.f - two_s * (x * x + z * z);
  r[5] = two_s * (y * z - x * w);
  r[6] = two_s * (x * z - y * w);
  r[7] = two_s * (y * z + x * w);
  r[8] = 1.f - two_s * (x * x + y * y);
  T[(size_t)i * 3 + 0] = e[0];
  T[(size_t)i * 3 + 1] = e[1];
  T[(size_t)i * 3 + 2] = e[2];
  F[(size_t)i * 2 + 0] = fminf(fmaxf(expf(e[7] + bias), fmin), fmax);
  F[(size_t)i * 2 + 1] = fminf(fmaxf(expf(e[8] + bias), fmin), fmax);
}

// relative pose of cameras (i, j) in pytorch3d's row-vector convention (X_view = X_world R + T; util/metric.py:29-41):
// inverse(se3_i) @ se3_j = [[R_i^T R_j, 0], [T_j - T_i R_i^T R_j, 1]]
__device__ __forceinline__ void relative_pose(const float* Ri, const float* Ti, const float* Rj, const float* Tj, float (&R)[9], float (&t)[3]) {
  float u[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) u[k] = -(Ti[0] * Ri[k * 3 + 0] + Ti[1] * Ri[k * 3 + 1] + Ti[2] * Ri[k * 3 + 2]);  // -T_i R_i^T
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) R[a * 3 + c] = Ri[0 * 3 + a] * Rj[0 * 3 + c] + Ri[1 * 3 + a] * Rj[1 * 3 + c] + Ri[2 * 3 + a] * Rj[2 * 3 + c];
#pragma unroll
  for (int c = 0; c < 3; ++c) t[c] = u[0] * Rj[0 * 3 + c] + u[1] * Rj[1 * 3 + c] + u[2] * Rj[2 * 3 + c] + Tj[c];
}

// camera_to_rel_deg (util/metric.py:14-48): for every unordered pair i < j of every sequence (torch.combinations order, sequence
// major) the angle between the relative rotations (pytorch3d so3_relative_angle: acos of (trace - 1) / 2 with the linear
// extrapolation outside +-(1 - 1e-4)) and between the relative translation directions (compare_translation_by_angle), in degrees.
// flag[0] is set when a trace leaves [-1 - 1e-4, 3 + 1e-4] (the reference raises ValueError there).
__global__ void rel_pose_error_kernel(const float* __restrict__ Rp, const float* __restrict__ Tp, const float* __restrict__ Rg,
                                      const float* __restrict__ Tg, int batch, int frames, float* __restrict__ r_deg,
                                      float* __restrict__ t_deg, int* __restrict__ flag) {
  const int pairs = frames * (frames - 1) / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * pairs) return;
  const int b = idx / pairs;
  int p = idx - b * pairs, i = 0;
  while (p >= frames - 1 - i) {  // row i of the strict upper triangle holds frames - 1 - i pairs
    p -= frames - 1 - i;
    ++i;
  }
  const int j = i + 1 + p;
  const size_t ci = (size_t)b * frames + i, cj = (size_t)b * frames + j;
  float Rg_rel[9], tg_rel[3], Rp_rel[9], tp_rel[3];
  relative_pose(Rg + ci * 9, Tg + ci * 3, Rg + cj * 9, Tg + cj * 3, Rg_rel, tg_rel);
  relative_pose(Rp + ci * 9, Tp + ci * 3, Rp + cj * 9, Tp + cj * 3, Rp_rel, tp_rel);
  float trace = 0.f;  // trace(Rg_rel Rp_rel^T)
#pragma unroll
  for (int k = 0; k < 9; ++k) trace = fmaf(Rg_rel[k], Rp_rel[k], trace);
  const float eps = 1e-4f, bound = 1.0f - 1e-4f;
  if (trace < -1.0f - eps || trace > 3.0f + eps) atomicOr(flag, 1);
  const float c = (trace - 1.0f) * 0.5f;
  float ang;
  if (c > bound) {
    ang = acosf(bound) + (c - bound) * (-1.0f / sqrtf(1.0f - bound * bound));
  } else if (c < -bound) {
    ang = acosf(-bound) + (c + bound) * (-1.0f / sqrtf(1.0f - bound * bound));
  } else {
    ang = acosf(c);
  }
  r_deg[idx] = ang * 180.0f / 3.14159265358979323846f;
  // compare_translation_by_angle (util/metric.py:165-180), eps = 1e-15, default_err = 1e6
  const float teps = 1e-15f;
  const float ng = sqrtf(tg_rel[0] * tg_rel[0] + tg_rel[1] * tg_rel[1] + tg_rel[2] * tg_rel[2]) + teps;
  const float np = sqrtf(tp_rel[0] * tp_rel[0] + tp_rel[1] * tp_rel[1] + tp_rel[2] * tp_rel[2]) + teps;
  const float dot = (tp_rel[0] / np) * (tg_rel[0] / ng) + (tp_rel[1] / np) * (tg_rel[1] / ng) + (tp_rel[2] / np) * (tg_rel[2] / ng);
  const float loss = fmaxf(1.0f - dot * dot, teps);
  float err = acosf(sqrtf(1.0f - loss));
  if (isnan(err) || isinf(err)) err = 1e6f;
  t_deg[idx] = err * 180.0f / 3.14159265358979323846f;
}

// corresponding_cameras_alignment, mode "extrinsics": bodies in csrc/align.cuh (shared with the CPU emulation harness)
__global__ void cameras_align_estimate_kernel(const float* __restrict__ Rs, const float* __restrict__ Ts, const float* __restrict__ Rt,
                                              const float* __restrict__ Tt, int count, int estimate_scale, float eps,
                                              float* __restrict__ align) {
  cameras_align_estimate_warp(Rs, Ts, Rt, Tt, count, estimate_scale, eps, align);
}

__global__ void cameras_align_apply_kernel(const float* __restrict__ align, const float* __restrict__ Rs, const float* __restrict__ Ts,
                                           int count, float* __restrict__ Ro, float* __restrict__ To) {
  cameras_align_apply_thread(align, Rs, Ts, count, Ro, To);
}

}  // namespace

extern "C" int pdb_cameras_align(pdb_context* c, const float* R_src_dev, const float* T_src_dev, const float* R_tgt_dev,
                                 const float* T_tgt_dev, int32_t count, int32_t estimate_scale, double eps, float* R_out_dev,
                                 float* T_out_dev, float* align_dev, void* stream) {
  if (!c) return PDB_ERR_INVALID;
  Context* ctx = reinterpret_cast<Context*>(c);
  if (!R_src_dev || !T_src_dev || !R_tgt_dev || !T_tgt_dev || !R_out_dev || !T_out_dev || !align_dev)
    return ctx->fail(PDB_ERR_INVALID, "null argument");
  if (count < 1) return ctx->fail(PDB_ERR_INVALID, "count %d", count);
  PDB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cameras_align_estimate_kernel<<<1, 32, 0, st>>>(R_src_dev, T_src_dev, R_tgt_dev, T_tgt_dev, count, estimate_scale, (float)eps, align_dev);
  PDB_CUDA(ctx, cudaGetLastError());
  cameras_align_apply_kernel<<<(count + 127) / 128, 128, 0, st>>>(align_dev, R_src_dev, T_src_dev, count, R_out_dev, T_out_dev);
  PDB_CUDA(ctx, cudaGetLastError());
  ctx->launches += 2;
  return PDB_OK;
}

extern "C" int pdb_pose_to_camera(pdb_context* c, const float* pose_dev, int32_t count, double log_focal_length_bias,
                                  double min_focal_length, double max_focal_length, float* R_dev, float* T_dev, float* focal_dev,
                                  void* stream) {
  if (!c) return PDB_ERR_INVALID;
  Context* ctx = reinterpret_cast<Context*>(c);
  if (!pose_dev || !R_dev || !T_dev || !focal_dev) return ctx->fail(PDB_ERR_INVALID, "null argument");
  if (count < 1) return ctx->fail(PDB_ERR_INVALID, "count %d", count);
  PDB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pose_to_camera_kernel<<<(count + 127) / 128, 128, 0, st>>>(pose_dev, count, (float)log_focal_length_bias, (float)min_focal_length,
                                                          (float)max_focal_length, R_dev, T_dev, focal_dev);
  PDB_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  return PDB_OK;
}

extern "C" int pdb_rel_pose_error(pdb_context* c, const float* R_pred_dev, const float* T_pred_dev, const float* R_gt_dev,
                                  const float* T_gt_dev, int32_t batch, int32_t frames, float* r_deg_dev, float* t_deg_dev,
                                  int32_t* invalid_dev, void* stream) {
  if (!c) return PDB_ERR_INVALID;
  Context* ctx = reinterpret_cast<Context*>(c);
  if (!R_pred_dev || !T_pred_dev || !R_gt_dev || !T_gt_dev || !r_deg_dev || !t_deg_dev || !invalid_dev)
    return ctx->fail(PDB_ERR_INVALID, "null argument");
  if (batch < 1 || frames < 2) return ctx->fail(PDB_ERR_INVALID, "need batch >= 1 and frames >= 2 (got %d, %d)", batch, frames);
  PDB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = (long long)batch * frames * (frames - 1) / 2;
  if (total > (1ll << 30)) return ctx->fail(PDB_ERR_LIMIT, "%lld pairs", total);
  PDB_CUDA(ctx, cudaMemsetAsync(invalid_dev, 0, sizeof(int32_t), st));
  rel_pose_error_kernel<<<(int)((total + 127) / 128), 128, 0, st>>>(R_pred_dev, T_pred_dev, R_gt_dev, T_gt_dev, batch, frames, r_deg_dev,
                                                                     t_deg_dev, invalid_dev);
  PDB_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  return PDB_OK;
}
