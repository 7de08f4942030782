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

// ---------------------------------------------------------------------------------------------------------------
// Stage-1 traversal of the GGS kernel (csrc/ggs.cuh), replayed sequentially: CTA / warp partition (the kernel's own
// ggs_cta_range / ggs_warp_range_seg), pair-segment switches, the ring / resident / register-stream walks with their
// fast-path conditions and in-bounds tests, for the plain and the paired stream layouts (csrc/ggs_layout.cuh).
// It checks the INDEXING contract of a layout without a GPU: which row of the stream image is read as which match of
// which segment.  Lanes are a loop; shuffles / barriers / atomics have no counterpart here.
//   segs   [nseg+1][4] {first_round, count, a, b} + sentinel {rounds, 0, 0, 0}
//   F      [nseg][9]   F' of every segment (any values; the test passes random ones)
//   mode   0 = shared-memory resident walk, 1 = bulk-async ring walk, 2 = register-stream walk
//   acc    [nseg][12]  G[9], sum of clamped errors, sum of valid errors, valid count
//   visits [m_total]   how often each match (segment order) was consumed in-bounds
//   fast   [2]         {matches consumed by a packed fast path, matches consumed by a scalar path}
// ---------------------------------------------------------------------------------------------------------------
#include "../../posediffusion_b200/csrc/ggs_layout.cuh"

namespace {
constexpr int kWalkWarps = 16, kWalkUnroll = 4, kWalkMaxSeg = 128;

struct Walk {
  const float4* pts;
  const int4* segs;
  int nseg;
  const float* F;
  float smax;
  bool paired;
  float* acc;
  int* visits;
  long long* fast;
  std::vector<long long> first;  // first match of every segment (segment order)

  int seg_of_round(int r) const {
    int lo = 0, hi = nseg;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (segs[mid].x <= r) lo = mid; else hi = mid;
    }
    return lo;
  }
  // one lane consumes one match row of segment s: `row` = index inside the segment if in bounds
  void consume(int s, const float4 m, bool inb, long long row, bool packed) {
    float a16[16];
    for (int k = 0; k < 16; ++k) a16[k] = 0.f;
    sampson_match<true>(m, F + (size_t)s * 9, inb, smax, a16);
    for (int k = 0; k < 12; ++k) acc[(size_t)s * 12 + k] += a16[k];
    if (inb) {
      visits[first[s] + row] += 1;
      fast[packed ? 0 : 1] += 1;
    }
  }
  // rounds q (and q+1 when paired) of segment s, all 32 lanes; `full` = the kernel took a no-padding fast path
  void round_plain(int s, int q, const float4* src, bool full, bool packed) {
    const int4 sd = segs[s];
    for (int lane = 0; lane < 32; ++lane) {
      const long long row = (long long)(q - sd.x) * 32 + lane;
      consume(s, src[lane], full ? true : row < sd.y, row, packed);
    }
  }
  void unit_paired(int s, int q, const float4* srcX, const float4* srcY, bool full, bool packed) {
    const int4 sd = segs[s];
    for (int lane = 0; lane < 32; ++lane) {
      const long long row_a = (long long)(q - sd.x) * 32 + lane, row_b = (long long)(q + 1 - sd.x) * 32 + lane;
      consume(s, unit_match_a(srcX[lane], srcY[lane]), full ? true : row_a < sd.y, row_a, packed);
      consume(s, unit_match_b(srcX[lane], srcY[lane]), full ? true : row_b < sd.y, row_b, packed);
    }
  }
};
}  // namespace

extern "C" int ggs_host_walk(const float* pts, const int* segs, int nseg, long long rounds, int cpp, int paired, int mode,
                             const float* F, float smax, float* acc, int* visits, long long* fast) {
  Walk w;
  w.pts = reinterpret_cast<const float4*>(pts);
  w.segs = reinterpret_cast<const int4*>(segs);
  w.nseg = nseg;
  w.F = F;
  w.smax = smax;
  w.paired = paired != 0;
  w.acc = acc;
  w.visits = visits;
  w.fast = fast;
  w.first.assign(nseg + 1, 0);
  for (int s = 0; s < nseg; ++s) w.first[s + 1] = w.first[s] + w.segs[s].y;
  if (w.paired && (rounds & 1)) return -2;
  for (int cta = 0; cta < cpp; ++cta) {
    int r_cta0, r_cta1;
    ggs_cta_range(rounds, cta, cpp, w.paired, &r_cta0, &r_cta1);
    if (ggs_rounds_per_cta(rounds, cpp, w.paired) < r_cta1 - r_cta0) return -3;  // the shared-memory cache would overflow
    const bool cta_has_work = r_cta1 > r_cta0;
    const int seg_lo = cta_has_work ? w.seg_of_round(r_cta0) : 0;
    const int seg_hi = cta_has_work ? w.seg_of_round(r_cta1 - 1) : -1;
    const bool single_chunk = (seg_hi - seg_lo + 1) <= kWalkMaxSeg;
    const bool use_ring = mode == 1 && single_chunk;
    const bool resident = mode == 0;
    for (int warp = 0; warp < kWalkWarps; ++warp) {
      int r_w0, r_w1;
      ggs_warp_range_seg(w.segs, seg_lo, seg_hi, r_cta0, r_cta1, warp, kWalkWarps, w.paired, &r_w0, &r_w1);
      if (w.paired && ((r_w0 | r_w1 | r_cta0 | r_cta1) & 1)) return -4;
      const int wseg0 = (r_w1 > r_w0) ? w.seg_of_round(r_w0) : 0;
      for (int cs = seg_lo; cs <= seg_hi; cs += kWalkMaxSeg) {
        const int ce = (cs + kWalkMaxSeg < seg_hi + 1) ? cs + kWalkMaxSeg : seg_hi + 1;
        if (use_ring) {
          const int nr_w = r_w1 - r_w0;
          if (nr_w <= 0) continue;
          const int nch = (nr_w + kWalkUnroll - 1) / kWalkUnroll;
          int s = wseg0;
          int4 sd = w.segs[s];
          int seg_end = w.segs[s + 1].x;
          for (int c = 0; c < nch; ++c) {
            const int q0 = r_w0 + c * kWalkUnroll;
            const int rounds_c = (kWalkUnroll < nr_w - c * kWalkUnroll) ? kWalkUnroll : nr_w - c * kWalkUnroll;
            if (w.paired && (rounds_c & 1)) return -5;  // a bulk copy would split a unit
            const float4* stage = w.pts + (size_t)q0 * 32;  // what the bulk copy brought into the ring stage
            if (q0 + kWalkUnroll <= r_w1 && q0 + kWalkUnroll <= seg_end && (q0 + kWalkUnroll - sd.x) * 32 <= sd.y) {
              if (w.paired) {
                for (int u = 0; u < kWalkUnroll; u += 2) w.unit_paired(s, q0 + u, stage + u * 32, stage + (u + 1) * 32, true, true);
              } else {
                for (int u = 0; u < kWalkUnroll; ++u) w.round_plain(s, q0 + u, stage + u * 32, true, true);
              }
            } else if (w.paired) {
              for (int u = 0; u < kWalkUnroll; u += 2) {
                const int q = q0 + u;
                if (q < r_w1) {
                  if (q >= seg_end) {
                    ++s;
                    sd = w.segs[s];
                    seg_end = w.segs[s + 1].x;
                  }
                  if (u + 1 >= rounds_c) return -6;  // Y would lie outside the copied bytes
                  w.unit_paired(s, q, stage + u * 32, stage + (u + 1) * 32, false, false);
                }
              }
            } else {
              for (int u = 0; u < kWalkUnroll; ++u) {
                const int q = q0 + u;
                if (q < r_w1) {
                  if (q >= seg_end) {
                    ++s;
                    sd = w.segs[s];
                    seg_end = w.segs[s + 1].x;
                  }
                  w.round_plain(s, q, stage + u * 32, false, false);
                }
              }
            }
          }
        } else {
          int s = cs > wseg0 ? cs : wseg0;
          int r = (s < ce) ? (r_w0 > w.segs[s].x ? r_w0 : w.segs[s].x) : r_w1;
          while (s < ce && r < r_w1) {
            const int4 sd = w.segs[s];
            const int r_end = r_w1 < w.segs[s + 1].x ? r_w1 : w.segs[s + 1].x;
            const int seg_first = sd.x, seg_count = sd.y;
            if (w.paired && ((r | r_end) & 1)) return -7;
            if (resident) {
              const float4* base = w.pts + (size_t)r * 32;  // = s_pts + (r - r_cta0) * 32 after the staging copy
              const int nr = r_end - r;
              int n_full = seg_first + seg_count / 32 - r;
              n_full = n_full < nr ? n_full : nr;
              n_full = n_full > 0 ? n_full : 0;
              if (w.paired) {
                const int n_full2 = n_full & ~1;
                for (int q = 0; q < n_full2; q += 2) w.unit_paired(s, r + q, base + q * 32, base + (q + 1) * 32, true, true);
                for (int q = n_full2; q < nr; q += 2) w.unit_paired(s, r + q, base + q * 32, base + (q + 1) * 32, false, false);
              } else {
                int qd = 0;
                for (; qd + 1 < n_full; qd += 2) {
                  w.round_plain(s, r + qd, base + qd * 32, true, true);
                  w.round_plain(s, r + qd + 1, base + (qd + 1) * 32, true, true);
                }
                for (int q = qd; q < nr; ++q) w.round_plain(s, r + q, base + q * 32, false, false);
              }
            } else {
              for (int rr = r; rr < r_end; rr += kWalkUnroll) {
                if (w.paired) {
                  for (int u = 0; u < kWalkUnroll; u += 2)
                    if (rr + u < r_end) {
                      if (rr + u + 1 >= r_end) return -8;  // cur[u+1] would not have been loaded
                      w.unit_paired(s, rr + u, w.pts + (size_t)(rr + u) * 32, w.pts + (size_t)(rr + u + 1) * 32, false, false);
                    }
                } else {
                  for (int u = 0; u < kWalkUnroll; ++u)
                    if (rr + u < r_end) w.round_plain(s, rr + u, w.pts + (size_t)(rr + u) * 32, false, false);
                }
              }
            }
            r = r_end;
            ++s;
          }
        }
      }
    }
  }
  return 0;
}

extern "C" long long layout_float_index_host(long long first_round, long long k, int comp, int paired) {
  return (long long)layout_float_index(first_round, k, comp, paired != 0);
}

// ---------------------------------------------------------------------------------------------------------------
// csrc/align.cuh on the host: the camera alignment exactly as the two kernels of csrc/api_post.cu compute it
// (means, centred second moments, V U^T by one-sided Jacobi, application), sequentially.
// ---------------------------------------------------------------------------------------------------------------
#include "../../posediffusion_b200/csrc/align.cuh"

extern "C" void svd3_v_ut_host(const float* M, float* out) { svd3_v_ut(M, out); }

extern "C" void cameras_align_host(const float* Rs, const float* Ts, const float* Rt, const float* Tt, int count, int estimate_scale,
                                   float eps, float* Ro, float* To, float* align) {
  float P[9], A[3], B[3], sum[15];
  for (int k = 0; k < 15; ++k) sum[k] = 0.f;
  for (int i = 0; i < count; ++i) {
    align_camera_terms(Rs + i * 9, Ts + i * 3, Rt + i * 9, Tt + i * 3, P, A, B);
    for (int k = 0; k < 9; ++k) sum[k] += P[k];
    for (int k = 0; k < 3; ++k) { sum[9 + k] += A[k]; sum[12 + k] += B[k]; }
  }
  for (int k = 0; k < 15; ++k) sum[k] /= (float)count;
  float scale = 1.f;
  if (estimate_scale && count > 1) {
    float ab = 0.f, aa = 0.f;
    for (int i = 0; i < count; ++i) {
      align_camera_terms(Rs + i * 9, Ts + i * 3, Rt + i * 9, Tt + i * 3, P, A, B);
      for (int k = 0; k < 3; ++k) {
        const float ac = A[k] - sum[9 + k], bc = B[k] - sum[12 + k];
        ab = fmaf(ac, bc, ab);
        aa = fmaf(ac, ac, aa);
      }
    }
    ab /= (float)(3 * count);
    aa /= (float)(3 * count);
    scale = ab / fmaxf(aa, eps);
  }
  svd3_v_ut(sum, align);
  for (int k = 0; k < 3; ++k) align[9 + k] = sum[12 + k] - scale * sum[9 + k];
  align[12] = scale;
  for (int i = 0; i < count; ++i) align_apply_camera(align, Rs + i * 9, Ts + i * 3, Ro + i * 9, To + i * 3);
}
