// 7-dof alignment of a set of predicted cameras to target cameras ("Umeyama" step of the reference's demo / test scripts:
// demo.py:126-128 -> pytorch3d.ops.corresponding_cameras_alignment(mode="extrinsics", estimate_scale=True, eps=1e-9)).
//
// pytorch3d is a third-party dependency that is absent from /root/reference and unpinned (install.sh:24); its published
// algorithm (ops/cameras_alignment.py, `_align_camera_extrinsics`) is restated here, in pytorch3d's row-vector convention
// X_view = X_world R + T:
//     M        = mean_i R_src_i R_tgt_i^T                      U, S, V = svd(M)          align_R = V U^T
//     A_i      = R_src_i T_src_i        B_i = R_src_i T_tgt_i   Amu, Bmu = means over the cameras
//     s        = mean((A - Amu) (B - Bmu)) / max(mean((A - Amu)^2), eps)   (1 when estimate_scale is off or there is 1 camera)
//     align_T  = Bmu - s Amu
//     R_i'     = align_R R_src_i        T_i' = align_T R_src_i + s T_src_i
// V U^T is the transposed orthogonal polar factor of M, so it does not depend on the sign / ordering conventions of the SVD
// routine as long as M has full rank (for a rank-deficient M the reference's answer is LAPACK-dependent too).
//
// Host + device: tests/host/geom_host.cu runs the same functions on the CPU against numpy's SVD.
#pragma once
#include "geom.cuh"

namespace pdb {

// out = V U^T for M = U S V^T (row-major 3x3), by one-sided Jacobi: rotate the columns of A = M V until they are orthogonal.
PDB_HD void svd3_v_ut(const float* M, float* out) {
  float A[9], V[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
#pragma unroll
  for (int k = 0; k < 9; ++k) A[k] = M[k];
  for (int sweep = 0; sweep < 12; ++sweep) {
    float off = 0.f;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;  // (0,1), (0,2), (1,2)
      const float alpha = A[p] * A[p] + A[3 + p] * A[3 + p] + A[6 + p] * A[6 + p];
      const float beta = A[q] * A[q] + A[3 + q] * A[3 + q] + A[6 + q] * A[6 + q];
      const float gamma = A[p] * A[q] + A[3 + p] * A[3 + q] + A[6 + p] * A[6 + q];
      const float lim = 1e-7f * sqrtf(alpha * beta);
      if (fabsf(gamma) > lim && gamma != 0.f) {
        off += fabsf(gamma);
        const float zeta = (beta - alpha) / (2.f * gamma);
        const float t = (zeta >= 0.f ? 1.f : -1.f) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
        const float c = 1.f / sqrtf(1.f + t * t), s = c * t;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float ap = A[r * 3 + p], aq = A[r * 3 + q];
          A[r * 3 + p] = c * ap - s * aq;
          A[r * 3 + q] = s * ap + c * aq;
          const float vp = V[r * 3 + p], vq = V[r * 3 + q];
          V[r * 3 + p] = c * vp - s * vq;
          V[r * 3 + q] = s * vp + c * vq;
        }
      }
    }
    if (off == 0.f) break;
  }
  // columns of A are sigma_i u_i; order them by decreasing norm so that a vanishing one (rank-deficient M) comes last
  float n2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) n2[i] = A[i] * A[i] + A[3 + i] * A[3 + i] + A[6 + i] * A[6 + i];
  int o0 = 0, o1 = 1, o2 = 2;
  if (n2[o0] < n2[o1]) { const int t = o0; o0 = o1; o1 = t; }
  if (n2[o1] < n2[o2]) { const int t = o1; o1 = o2; o2 = t; }
  if (n2[o0] < n2[o1]) { const int t = o0; o0 = o1; o1 = t; }
  float U[9];  // column i of U belongs to column i of V
  const float tiny = 1e-12f * n2[o0];
  const int ord[3] = {o0, o1, o2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int i = ord[k];
    if (n2[i] > tiny && n2[i] > 0.f) {
      const float inv = 1.f / sqrtf(n2[i]);
      U[0 + i] = A[0 + i] * inv; U[3 + i] = A[3 + i] * inv; U[6 + i] = A[6 + i] * inv;
    } else if (k == 2) {  // complete the basis: u = u_a x u_b, oriented like v_i relative to (v_a, v_b)
      const int a = ord[0], b = ord[1];
      float ux = U[3 + a] * U[6 + b] - U[6 + a] * U[3 + b];
      float uy = U[6 + a] * U[0 + b] - U[0 + a] * U[6 + b];
      float uz = U[0 + a] * U[3 + b] - U[3 + a] * U[0 + b];
      const float vx = V[3 + a] * V[6 + b] - V[6 + a] * V[3 + b];
      const float vy = V[6 + a] * V[0 + b] - V[0 + a] * V[6 + b];
      const float vz = V[0 + a] * V[3 + b] - V[3 + a] * V[0 + b];
      const float sgn = (vx * V[0 + i] + vy * V[3 + i] + vz * V[6 + i]) < 0.f ? -1.f : 1.f;
      U[0 + i] = sgn * ux; U[3 + i] = sgn * uy; U[6 + i] = sgn * uz;
    } else {  // rank <= 1: no unique answer exists; keep the direction of v_i (M = 0 gives the identity)
      U[0 + i] = V[0 + i]; U[3 + i] = V[3 + i]; U[6 + i] = V[6 + i];
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) out[r * 3 + c] = V[r * 3 + 0] * U[c * 3 + 0] + V[r * 3 + 1] * U[c * 3 + 1] + V[r * 3 + 2] * U[c * 3 + 2];
}

// per-camera terms of the estimate: P = R_src R_tgt^T (9), A = R_src T_src (3), B = R_src T_tgt (3)
PDB_HD void align_camera_terms(const float* Rs, const float* Ts, const float* Rt, const float* Tt, float* P, float* A, float* B) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) P[r * 3 + c] = Rs[r * 3 + 0] * Rt[c * 3 + 0] + Rs[r * 3 + 1] * Rt[c * 3 + 1] + Rs[r * 3 + 2] * Rt[c * 3 + 2];
    A[r] = Rs[r * 3 + 0] * Ts[0] + Rs[r * 3 + 1] * Ts[1] + Rs[r * 3 + 2] * Ts[2];
    B[r] = Rs[r * 3 + 0] * Tt[0] + Rs[r * 3 + 1] * Tt[1] + Rs[r * 3 + 2] * Tt[2];
  }
}

// apply the alignment {align_R[9], align_T[3], s} to one camera
PDB_HD void align_apply_camera(const float* al, const float* Rs, const float* Ts, float* Ro, float* To) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Ro[r * 3 + c] = al[r * 3 + 0] * Rs[0 + c] + al[r * 3 + 1] * Rs[3 + c] + al[r * 3 + 2] * Rs[6 + c];
#pragma unroll
  for (int c = 0; c < 3; ++c) To[c] = al[9] * Rs[0 + c] + al[10] * Rs[3 + c] + al[11] * Rs[6 + c] + al[12] * Ts[c];
}

// ---- kernel bodies (device side; csrc/api_post.cu wraps them in __global__ functions, tests/host/kernels_emu.cpp runs them on
// the CPU emulation of the execution model) ----
// Estimate: ONE warp; lanes stride over the cameras, two passes (means, then centred second moments, as the reference computes
// them), lane 0 finishes with the 3x3 SVD.  align[13] = {align_R (9, row-major), align_T (3), s}.
__device__ __forceinline__ void cameras_align_estimate_warp(const float* __restrict__ Rs, const float* __restrict__ Ts,
                                                            const float* __restrict__ Rt, const float* __restrict__ Tt, int count,
                                                            int estimate_scale, float eps, float* __restrict__ align) {
  const int lane = threadIdx.x;
  float P[9], A[3], B[3], sum[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) sum[k] = 0.f;
  for (int i = lane; i < count; i += 32) {
    align_camera_terms(Rs + (size_t)i * 9, Ts + (size_t)i * 3, Rt + (size_t)i * 9, Tt + (size_t)i * 3, P, A, B);
#pragma unroll
    for (int k = 0; k < 9; ++k) sum[k] += P[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      sum[9 + k] += A[k];
      sum[12 + k] += B[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 15; ++k) sum[k] = warp_sum(sum[k]) / (float)count;  // every lane holds the means
  float scale = 1.f;
  if (estimate_scale && count > 1) {
    float ab = 0.f, aa = 0.f;
    for (int i = lane; i < count; i += 32) {
      align_camera_terms(Rs + (size_t)i * 9, Ts + (size_t)i * 3, Rt + (size_t)i * 9, Tt + (size_t)i * 3, P, A, B);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float ac = A[k] - sum[9 + k], bc = B[k] - sum[12 + k];
        ab = fmaf(ac, bc, ab);
        aa = fmaf(ac, ac, aa);
      }
    }
    ab = warp_sum(ab) / (float)(3 * count);
    aa = warp_sum(aa) / (float)(3 * count);
    scale = ab / fmaxf(aa, eps);  // (Ac * Bc).mean() / (Ac ** 2).mean().clamp(eps)
  }
  if (lane == 0) {
    svd3_v_ut(sum, align);
#pragma unroll
    for (int k = 0; k < 3; ++k) align[9 + k] = sum[12 + k] - scale * sum[9 + k];
    align[12] = scale;
  }
}

// Application: one thread per camera.
__device__ __forceinline__ void cameras_align_apply_thread(const float* __restrict__ align, const float* __restrict__ Rs,
                                                           const float* __restrict__ Ts, int count, float* __restrict__ Ro,
                                                           float* __restrict__ To) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float al[13];
#pragma unroll
  for (int k = 0; k < 13; ++k) al[k] = align[k];
  align_apply_camera(al, Rs + (size_t)i * 9, Ts + (size_t)i * 3, Ro + (size_t)i * 9, To + (size_t)i * 3);
}

}  // namespace pdb
