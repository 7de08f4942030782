// TEST HARNESS (not product): runs the __host__ __device__ geometry / Sampson maths of
// posediffusion_b200/csrc/geom.cuh sequentially on the CPU so the closed-form adjoint can be checked
// against the oracle without a GPU.  The parallel decomposition (rounds, warps, barriers) is GPU-only
// and is covered by the -m gpu tests.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../posediffusion_b200/csrc/geom.cuh"

using namespace pdb;

struct HostAdd {
  __host__ __device__ void operator()(float* p, float v) const { *p += v; }
};

extern "C" int geom_host_eval(const float* pose, int N, float height, float width, const float* pts /*[m,4]*/,
                              const int* segs /*[nseg,4] first,count,a,b (first in matches)*/, int nseg, int flags,
                              float smax, float* grad /*[N*9]*/, float* scalars /*[4]*/, float* F_out, float* G_out) {
  std::vector<float> R(N * 9), A(N * 9), fl(N * 2), inr(N * 2), gR(N * 9, 0.f), gA(N * 9, 0.f);
  for (int n = 0; n < N; ++n) frame_forward(pose + n * 9, &R[n * 9], &A[n * 9], &fl[n * 2], &inr[n * 2]);
  const float scale = 0.5f * fminf(height, width), cx = 0.5f * width, cy = 0.5f * height;
  float fx = 0.f, fy = 0.f;
  for (int n = 0; n < N; ++n) { fx += fl[n * 2]; fy += fl[n * 2 + 1]; }
  fx = fx / (float)N * scale;
  fy = fy / (float)N * scale;
  float kin[4] = {1.f / fx, 1.f / fy, -cx / fx, -cy / fy};
  float gk[4] = {0, 0, 0, 0};
  long long nvalid = 0, total = 0;
  float clamp_sum = 0.f, loss_sum = 0.f;
  HostAdd add;
  for (int s = 0; s < nseg; ++s) {
    const int first = segs[s * 4], count = segs[s * 4 + 1], a = segs[s * 4 + 2], b = segs[s * 4 + 3];
    float F[9];
    pair_F(&R[a * 9], &A[a * 9], &R[b * 9], &A[b * 9], kin, a == b, F);
    float acc[16];
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    for (int k = 0; k < count; ++k) {
      const float* p = pts + (size_t)(first + k) * 4;
      sampson_match<true>(make_float4(p[0], p[1], p[2], p[3]), F, true, smax, acc);
    }
    nvalid += (long long)acc[11];
    total += count;
    clamp_sum += acc[9];
    loss_sum += acc[10];
    if (F_out) memcpy(F_out + s * 9, F, sizeof(F));
    if (G_out) memcpy(G_out + s * 9, acc, 9 * sizeof(float));
    pair_adjoint(&R[a * 9], &A[a * 9], &R[b * 9], &A[b * 9], kin, a == b, acc, &gR[a * 9], &gA[a * 9], &gR[b * 9],
                 &gA[b * 9], gk, add);
  }
  const float gfx = (-gk[0] + cx * gk[2]) / (fx * fx), gfy = (-gk[1] + cy * gk[3]) / (fy * fy);
  for (int n = 0; n < N; ++n) {
    float gT[3], gq[4];
    frame_adjoint(pose + n * 9, &R[n * 9], &gR[n * 9], &gA[n * 9], gT, gq);
    for (int k = 0; k < 3; ++k) grad[n * 9 + k] = (flags & 2) ? gT[k] / (float)nvalid : 0.f;
    for (int k = 0; k < 4; ++k) grad[n * 9 + 3 + k] = (flags & 1) ? gq[k] / (float)nvalid : 0.f;
    grad[n * 9 + 7] = (flags & 4) ? gfx * (scale / (float)N) * fl[n * 2] * inr[n * 2] / (float)nvalid : 0.f;
    grad[n * 9 + 8] = (flags & 4) ? gfy * (scale / (float)N) * fl[n * 2 + 1] * inr[n * 2 + 1] / (float)nvalid : 0.f;
  }
  scalars[0] = loss_sum / (float)nvalid;
  scalars[1] = (float)nvalid;
  scalars[2] = clamp_sum / (float)total;
  scalars[3] = 0.f;
  return 0;
}

// The same evaluation through the K-folded per-frame formulation the v2 kernel uses (entry-level functions).
extern "C" int geom_host_eval_folded(const float* pose, int N, float height, float width, const float* pts, const int* segs,
                                     int nseg, int flags, float smax, float* grad, float* scalars) {
  std::vector<float> R(N * 9), A(N * 9), Rt(N * 9), At(N * 9), fl(N * 2), inr(N * 2), gAt(N * 9, 0.f), gRt(N * 9, 0.f);
  for (int n = 0; n < N; ++n) frame_forward(pose + n * 9, &R[n * 9], &A[n * 9], &fl[n * 2], &inr[n * 2]);
  const float scale = 0.5f * fminf(height, width), cx = 0.5f * width, cy = 0.5f * height;
  float fx = 0.f, fy = 0.f;
  for (int n = 0; n < N; ++n) { fx += fl[n * 2]; fy += fl[n * 2 + 1]; }
  fx = fx / (float)N * scale;
  fy = fy / (float)N * scale;
  float kin[4] = {1.f / fx, 1.f / fy, -cx / fx, -cy / fy};
  for (int n = 0; n < N; ++n) { frame_tilde(&A[n * 9], kin, &At[n * 9]); frame_tilde(&R[n * 9], kin, &Rt[n * 9]); }
  long long nvalid = 0, total = 0;
  float clamp_sum = 0.f, loss_sum = 0.f;
  for (int s = 0; s < nseg; ++s) {
    const int first = segs[s * 4], count = segs[s * 4 + 1], a = segs[s * 4 + 2], b = segs[s * 4 + 3];
    float F[9];
    for (int e = 0; e < 9; ++e) F[e] = (a == b) ? 0.f : pair_F_entry(&At[a * 9], &Rt[a * 9], &At[b * 9], &Rt[b * 9], e / 3, e % 3);
    float acc[16];
    for (int k = 0; k < 16; ++k) acc[k] = 0.f;
    for (int k = 0; k < count; ++k) {
      const float* p = pts + (size_t)(first + k) * 4;
      sampson_match<true>(make_float4(p[0], p[1], p[2], p[3]), F, true, smax, acc);
    }
    nvalid += (long long)acc[11];
    total += count;
    clamp_sum += acc[9];
    loss_sum += acc[10];
    for (int side = 0; side < 2; ++side)
      for (int e = 0; e < 9; ++e) {
        const int i = e / 3, j = e % 3, self = side ? b : a, other = side ? a : b;
        float g3[3];
        for (int k = 0; k < 3; ++k) g3[k] = side ? acc[k * 3 + i] : acc[i * 3 + k];
        float oA, oR;
        pair_adjoint_entry(g3, &At[other * 9], &Rt[other * 9], j, &oA, &oR);
        gAt[self * 9 + e] += oA;
        gRt[self * 9 + e] += oR;
      }
  }
  float gk[4] = {0, 0, 0, 0};
  for (int n = 0; n < N; ++n) {
    float gA[9], gR[9], k4[4], gT[3], gq[4];
    frame_unfold(&A[n * 9], &R[n * 9], kin, &gAt[n * 9], &gRt[n * 9], gA, gR, k4);
    for (int k = 0; k < 4; ++k) gk[k] += k4[k];
    frame_adjoint(pose + n * 9, &R[n * 9], gR, gA, gT, gq);
    for (int k = 0; k < 3; ++k) grad[n * 9 + k] = (flags & 2) ? gT[k] / (float)nvalid : 0.f;
    for (int k = 0; k < 4; ++k) grad[n * 9 + 3 + k] = (flags & 1) ? gq[k] / (float)nvalid : 0.f;
  }
  const float gfx = (-gk[0] + cx * gk[2]) / (fx * fx), gfy = (-gk[1] + cy * gk[3]) / (fy * fy);
  for (int n = 0; n < N; ++n) {
    grad[n * 9 + 7] = (flags & 4) ? gfx * (scale / (float)N) * fl[n * 2] * inr[n * 2] / (float)nvalid : 0.f;
    grad[n * 9 + 8] = (flags & 4) ? gfy * (scale / (float)N) * fl[n * 2 + 1] * inr[n * 2 + 1] / (float)nvalid : 0.f;
  }
  scalars[0] = loss_sum / (float)nvalid;
  scalars[1] = (float)nvalid;
  scalars[2] = clamp_sum / (float)total;
  scalars[3] = 0.f;
  return 0;
}
