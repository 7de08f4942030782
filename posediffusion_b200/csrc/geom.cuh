// Closed-form camera geometry of the guided step and its hand-derived adjoint (fp32).
//
// Forward chain restated from the reference:
//   pose_encoding_to_camera            util/camera_transform.py:85-97   (+ pytorch3d quaternion_to_matrix)
//   focal mean over frames             util/geometry_guided_sampling.py:142
//   opencv_from_cameras_projection     pytorch3d (R_cv = D R^T, t_cv = D T, K from NDC focal)
//   get_essential/fundamental_matrix   util/get_fundamental_matrix.py:39-51, F' = F^T (:155)
// Simplification used (exact on SO(3), where quaternion_to_matrix always lands):
//   E = R12 hat(-R12^T t12) = -(A_b R_a^T + R_b A_a^T),  A_n = hat(t_n) R_n,   so
//   M := E^T = -(A_a R_b^T + R_a A_b^T)   and   F' = K^-T M K^-1.
// The adjoint (what torch autograd does for the reference) is written out by hand and checked against
// fp64 autograd in oracle/sampson_f64.py.
#pragma once
#include "common.cuh"

#define PDB_HD __host__ __device__ __forceinline__

namespace pdb {

constexpr float kLogFlBias = 1.8f;  // camera_transform.py:67
constexpr float kFlMin = 0.1f;      // camera_transform.py:68
constexpr float kFlMax = 20.0f;     // camera_transform.py:69

// Per-frame forward terms from one pose row p[9] = (T, q=(w,x,y,z), log-focal).
//   R[9]  : R_cv (row-major) = D * Rp^T, D = diag(-1,-1,1)
//   A[9]  : hat(t_cv) * R_cv
//   fl[2] : clamp(exp(lam + 1.8), 0.1, 20);  inr[2] : 1 where the clamp passes gradient
PDB_HD void frame_forward(const float* p, float* R, float* A, float* fl, float* inr) {
  const float w = p[3], x = p[4], y = p[5], z = p[6];
  const float s2 = 2.0f / (w * w + x * x + y * y + z * z);
  // pytorch3d quaternion_to_matrix (real part first, no normalisation step)
  const float r00 = 1.f - s2 * (y * y + z * z), r01 = s2 * (x * y - z * w), r02 = s2 * (x * z + y * w);
  const float r10 = s2 * (x * y + z * w), r11 = 1.f - s2 * (x * x + z * z), r12 = s2 * (y * z - x * w);
  const float r20 = s2 * (x * z - y * w), r21 = s2 * (y * z + x * w), r22 = 1.f - s2 * (x * x + y * y);
  // R_cv[i][j] = D_i * Rp[j][i]
  R[0] = -r00; R[1] = -r10; R[2] = -r20;
  R[3] = -r01; R[4] = -r11; R[5] = -r21;
  R[6] = r02;  R[7] = r12;  R[8] = r22;
  const float tx = -p[0], ty = -p[1], tz = p[2];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    A[0 + j] = -tz * R[3 + j] + ty * R[6 + j];
    A[3 + j] = tz * R[0 + j] - tx * R[6 + j];
    A[6 + j] = -ty * R[0 + j] + tx * R[3 + j];
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float e = expf(p[7 + k] + kLogFlBias);
    fl[k] = fminf(fmaxf(e, kFlMin), kFlMax);
    inr[k] = (e >= kFlMin && e <= kFlMax) ? 1.f : 0.f;
  }
}

// K^-1 = [[ix,0,kx],[0,iy,ky],[0,0,1]] packed as kin[4] = {ix, iy, kx, ky}.
PDB_HD void pair_M(const float* Ra, const float* Aa, const float* Rb, const float* Ab, float* M) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) acc += Aa[i * 3 + k] * Rb[j * 3 + k] + Ra[i * 3 + k] * Ab[j * 3 + k];
      M[i * 3 + j] = -acc;
    }
}

// P = M K^-1
PDB_HD void times_Kinv(const float* M, const float* kin, float* P) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    P[i * 3 + 0] = M[i * 3 + 0] * kin[0];
    P[i * 3 + 1] = M[i * 3 + 1] * kin[1];
    P[i * 3 + 2] = M[i * 3 + 0] * kin[2] + M[i * 3 + 1] * kin[3] + M[i * 3 + 2];
  }
}

// F' = K^-T (M K^-1); a diagonal pair (same frame twice) has E = 0: the reference's fp32 chain gives exactly 0
// there, hence NaN errors that poison its gradient; we make the zero explicit.
PDB_HD void pair_F(const float* Ra, const float* Aa, const float* Rb, const float* Ab,
                                       const float* kin, bool diagonal, float* F) {
  if (diagonal) {
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = 0.f;
    return;
  }
  float M[9], P[9];
  pair_M(Ra, Aa, Rb, Ab, M);
  times_Kinv(M, kin, P);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    F[0 + j] = kin[0] * P[0 + j];
    F[3 + j] = kin[1] * P[3 + j];
    F[6 + j] = kin[2] * P[0 + j] + kin[3] * P[3 + j] + P[6 + j];
  }
}

// Adjoint of F' wrt the per-frame terms and the intrinsics, for one pair, given G = dL/dF' (3x3).
//   gRa,gAa,gRb,gAb [9] are ACCUMULATED via `add(ptr, value)`; gk[4] += d/d(ix, iy, kx, ky).
template <typename Add>
PDB_HD void pair_adjoint(const float* Ra, const float* Aa, const float* Rb, const float* Ab,
                                             const float* kin, bool diagonal, const float* G, float* gRa, float* gAa,
                                             float* gRb, float* gAb, float* gk, Add add) {
  float M[9];
  if (diagonal) {
#pragma unroll
    for (int i = 0; i < 9; ++i) M[i] = 0.f;
  } else {
    pair_M(Ra, Aa, Rb, Ab, M);
  }
  // H = K^-1 G K^-T
  float KG[9], H[9];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    KG[0 + j] = kin[0] * G[0 + j] + kin[2] * G[6 + j];
    KG[3 + j] = kin[1] * G[3 + j] + kin[3] * G[6 + j];
    KG[6 + j] = G[6 + j];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    H[i * 3 + 0] = KG[i * 3 + 0] * kin[0] + KG[i * 3 + 2] * kin[2];
    H[i * 3 + 1] = KG[i * 3 + 1] * kin[1] + KG[i * 3 + 2] * kin[3];
    H[i * 3 + 2] = KG[i * 3 + 2];
  }
  // dL/dK^-1 = (M K^-1) G^T + (M^T K^-1) G ; only entries (0,0), (1,1), (0,2), (1,2) are free parameters
  float P[9], Q[9], Mt[9];
  times_Kinv(M, kin, P);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Mt[i * 3 + j] = M[j * 3 + i];
  times_Kinv(Mt, kin, Q);
  auto gKi = [&](int i, int j) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc += P[i * 3 + k] * G[j * 3 + k] + Q[i * 3 + k] * G[k * 3 + j];
    return acc;
  };
  add(&gk[0], gKi(0, 0));
  add(&gk[1], gKi(1, 1));
  add(&gk[2], gKi(0, 2));
  add(&gk[3], gKi(1, 2));
  // M = -(A_a R_b^T + R_a A_b^T):  gA_a -= H R_b ; gR_b -= H^T A_a ; gR_a -= H A_b ; gA_b -= H^T R_a
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float hRb = 0.f, htAa = 0.f, hAb = 0.f, htRa = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        hRb += H[i * 3 + k] * Rb[k * 3 + j];
        htAa += H[k * 3 + i] * Aa[k * 3 + j];
        hAb += H[i * 3 + k] * Ab[k * 3 + j];
        htRa += H[k * 3 + i] * Ra[k * 3 + j];
      }
      add(&gAa[i * 3 + j], -hRb);
      add(&gRb[i * 3 + j], -htAa);
      add(&gRa[i * 3 + j], -hAb);
      add(&gAb[i * 3 + j], -htRa);
    }
}

// Adjoint of the per-frame terms wrt the pose row: given gR (wrt R_cv) and gA (wrt A), returns
// gT[3] and gq[4] (focal handled separately, it is shared by all frames).
PDB_HD void frame_adjoint(const float* p, const float* R, const float* gR, const float* gA,
                                              float* gT, float* gq) {
  const float tx = -p[0], ty = -p[1], tz = p[2];
  // gRcv = gR + hat(t)^T gA = gR - hat(t) gA
  float gRcv[9];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    gRcv[0 + j] = gR[0 + j] - (-tz * gA[3 + j] + ty * gA[6 + j]);
    gRcv[3 + j] = gR[3 + j] - (tz * gA[0 + j] - tx * gA[6 + j]);
    gRcv[6 + j] = gR[6 + j] - (-ty * gA[0 + j] + tx * gA[3 + j]);
  }
  // W = gA R^T ;  gt = (W21 - W12, W02 - W20, W10 - W01)
  auto W = [&](int i, int j) { return gA[i * 3 + 0] * R[j * 3 + 0] + gA[i * 3 + 1] * R[j * 3 + 1] + gA[i * 3 + 2] * R[j * 3 + 2]; };
  const float gt0 = W(2, 1) - W(1, 2), gt1 = W(0, 2) - W(2, 0), gt2 = W(1, 0) - W(0, 1);
  gT[0] = -gt0;
  gT[1] = -gt1;
  gT[2] = gt2;
  // gRp[j][i] = D_i gRcv[i][j]
  float g[9];
  g[0] = -gRcv[0]; g[3] = -gRcv[1]; g[6] = -gRcv[2];
  g[1] = -gRcv[3]; g[4] = -gRcv[4]; g[7] = -gRcv[5];
  g[2] = gRcv[6];  g[5] = gRcv[7];  g[8] = gRcv[8];
  const float w = p[3], x = p[4], y = p[5], z = p[6];
  const float s2 = 2.0f / (w * w + x * x + y * y + z * z);
  // Rp = I + s2 * B(q)
  const float gB = g[0] * (-(y * y + z * z)) + g[1] * (x * y - z * w) + g[2] * (x * z + y * w) +
                   g[3] * (x * y + z * w) + g[4] * (-(x * x + z * z)) + g[5] * (y * z - x * w) +
                   g[6] * (x * z - y * w) + g[7] * (y * z + x * w) + g[8] * (-(x * x + y * y));
  const float dw = -z * g[1] + y * g[2] + z * g[3] - x * g[5] - y * g[6] + x * g[7];
  const float dx = y * g[1] + z * g[2] + y * g[3] - 2.f * x * g[4] - w * g[5] + z * g[6] + w * g[7] - 2.f * x * g[8];
  const float dy = -2.f * y * g[0] + x * g[1] + w * g[2] + x * g[3] + z * g[5] - w * g[6] + z * g[7] - 2.f * y * g[8];
  const float dz = -2.f * z * g[0] - w * g[1] + x * g[2] + w * g[3] - 2.f * z * g[4] + y * g[5] + x * g[6] + y * g[7];
  const float c = -s2 * s2 * gB;
  gq[0] = c * w + s2 * dw;
  gq[1] = c * x + s2 * dx;
  gq[2] = c * y + s2 * dy;
  gq[3] = c * z + s2 * dz;
}

// One match of stage 1 (compute_sampson_distance :157-170 forward + the closed-form d err / d F').
// F = F' row-major.  Accumulators: acc[0..8] += G contribution, acc[9] += min(err, smax) (NaN passes, like
// torch.clamp), acc[10] += valid err (only if kWithLoss), acc[11] += 1 per valid match (exact in fp32 below 2^24
// per lane).  `inb` is false only for the padding lanes of a segment's last round (their coordinates are 0).
//
// Arithmetic is arranged for instruction count (this loop is the whole cost of the big configurations):
//   t = top / bottom, err = top * t, a = 2 t, b = a t  (= 2 err / bottom), and validity enters as a 0/1 WEIGHT
//   on `a` (not a select): 0 * (0/0) = NaN poisons the gradient exactly like the reference's autograd does for a
//   diagonal pair, and 0 * finite = 0 drops invalid matches.
PDB_HD float fast_rcp(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // one MUFU.RCP, <= 1 ulp; rcp(0) = inf like the IEEE quotient (denormal x -> inf)
  return r;
#else
  return 1.0f / x;
#endif
}

template <bool kWithLoss, int kStride = 1>
PDB_HD void sampson_match(const float4 pt, const float* F, bool inb, float smax, float* acc) {
  const float u1 = pt.x, v1 = pt.y, u2 = pt.z, v2 = pt.w;
  const float l0 = fmaf(u1, F[0], fmaf(v1, F[3], F[6]));
  const float l1 = fmaf(u1, F[1], fmaf(v1, F[4], F[7]));
  const float l2 = fmaf(u1, F[2], fmaf(v1, F[5], F[8]));
  const float r0 = fmaf(F[0], u2, fmaf(F[1], v2, F[2]));
  const float r1 = fmaf(F[3], u2, fmaf(F[4], v2, F[5]));
  const float top = fmaf(l0, u2, fmaf(l1, v2, l2));
  const float bottom = fmaf(l0, l0, fmaf(l1, l1, fmaf(r0, r0, r1 * r1)));
  const float t = top * fast_rcp(bottom);
  const float err = top * t;
  const bool valid = inb && (err < smax);
  const float wgt = valid ? 1.f : 0.f;
  const float a = wgt * (t + t);
  const float nb = -(a * t);
  const float clamped = (err > smax) ? smax : err;
  acc[9 * kStride] += inb ? clamped : 0.f;
  if (kWithLoss) acc[10 * kStride] += valid ? err : 0.f;
  acc[11 * kStride] += wgt;
  // G_ij += x1_i (a x2_j - b lz_j) - b rz_i x2_j
  const float w0 = fmaf(a, u2, nb * l0), w1 = fmaf(a, v2, nb * l1);
  const float c0 = nb * r0, c1 = nb * r1;
  acc[0 * kStride] = fmaf(c0, u2, fmaf(u1, w0, acc[0 * kStride]));
  acc[1 * kStride] = fmaf(c0, v2, fmaf(u1, w1, acc[1 * kStride]));
  acc[2 * kStride] = fmaf(u1, a, acc[2 * kStride]) + c0;
  acc[3 * kStride] = fmaf(c1, u2, fmaf(v1, w0, acc[3 * kStride]));
  acc[4 * kStride] = fmaf(c1, v2, fmaf(v1, w1, acc[4 * kStride]));
  acc[5 * kStride] = fmaf(v1, a, acc[5 * kStride]) + c1;
  acc[6 * kStride] += w0;
  acc[7 * kStride] += w1;
  acc[8 * kStride] += a;
}

// ---------------------------------------------------------------------------------------------
// "K-folded" formulation used by the kernels: with At_n = K^-T A_n and Rt_n = K^-T R_n (per frame),
//   F'_{ab} = K^-T M K^-1 = -(At_a Rt_b^T + Rt_a At_b^T),
// so every entry of F' is 6 FMAs of per-frame terms and the per-pair adjoint needs no K at all:
//   gAt_a -= G Rt_b ;  gRt_a -= G At_b ;  gRt_b -= G^T At_a ;  gAt_b -= G^T Rt_a.
// The intrinsics' adjoint moves to the (few) frames:  gA = K^-1 gAt, gR = K^-1 gRt, gK^-1 += A gAt^T + R gRt^T.
// ---------------------------------------------------------------------------------------------
PDB_HD void frame_tilde(const float* X, const float* kin, float* Xt) {  // Xt = K^-T X
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    Xt[0 + j] = kin[0] * X[0 + j];
    Xt[3 + j] = kin[1] * X[3 + j];
    Xt[6 + j] = kin[2] * X[0 + j] + kin[3] * X[3 + j] + X[6 + j];
  }
}

PDB_HD float pair_F_entry(const float* At_a, const float* Rt_a, const float* At_b, const float* Rt_b, int i, int j) {
  float acc = At_a[i * 3 + 0] * Rt_b[j * 3 + 0];
  acc = fmaf(At_a[i * 3 + 1], Rt_b[j * 3 + 1], acc);
  acc = fmaf(At_a[i * 3 + 2], Rt_b[j * 3 + 2], acc);
  acc = fmaf(Rt_a[i * 3 + 0], At_b[j * 3 + 0], acc);
  acc = fmaf(Rt_a[i * 3 + 1], At_b[j * 3 + 1], acc);
  acc = fmaf(Rt_a[i * 3 + 2], At_b[j * 3 + 2], acc);
  return -acc;
}

// One output column entry j of the pair adjoint for the frame `self`, given g3 = G[i][0..2] (a side) or
// G[0..2][i] (b side) and the OTHER frame's folded terms:  out_At = -sum_k g3[k] Rt_o[k][j],  out_Rt = -sum_k g3[k] At_o[k][j].
PDB_HD void pair_adjoint_entry(const float* g3, const float* At_o, const float* Rt_o, int j, float* out_At, float* out_Rt) {
  *out_At = -(g3[0] * Rt_o[0 + j] + g3[1] * Rt_o[3 + j] + g3[2] * Rt_o[6 + j]);
  *out_Rt = -(g3[0] * At_o[0 + j] + g3[1] * At_o[3 + j] + g3[2] * At_o[6 + j]);
}

// Per-frame: (gAt, gRt) -> (gA, gR) and the frame's contribution to d/d(ix, iy, kx, ky).
PDB_HD void frame_unfold(const float* A, const float* R, const float* kin, const float* gAt, const float* gRt, float* gA,
                         float* gR, float* gk) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    gA[0 + j] = kin[0] * gAt[0 + j] + kin[2] * gAt[6 + j];
    gA[3 + j] = kin[1] * gAt[3 + j] + kin[3] * gAt[6 + j];
    gA[6 + j] = gAt[6 + j];
    gR[0 + j] = kin[0] * gRt[0 + j] + kin[2] * gRt[6 + j];
    gR[3 + j] = kin[1] * gRt[3 + j] + kin[3] * gRt[6 + j];
    gR[6 + j] = gRt[6 + j];
  }
  auto gKi = [&](int i, int j) {  // sum_k A[i][k] gAt[j][k] + R[i][k] gRt[j][k]
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc += A[i * 3 + k] * gAt[j * 3 + k] + R[i * 3 + k] * gRt[j * 3 + k];
    return acc;
  };
  gk[0] = gKi(0, 0);
  gk[1] = gKi(1, 1);
  gk[2] = gKi(0, 2);
  gk[3] = gKi(1, 2);
}

}  // namespace pdb
