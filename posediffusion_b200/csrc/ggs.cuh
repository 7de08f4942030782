// Geometry-guided sampling on the device: one persistent cooperative kernel runs ALL phases and ALL
// SGD iterations of one `geometry_guided_sampling` call (util/geometry_guided_sampling.py:14-126) for a
// batch of independent sequences, with no host round trip.
//
// Per inner iteration (reference: one compute_sampson_distance forward + autograd backward + clip + SGD step):
//   stage 0  every CTA turns the pose (kept in shared memory, identical in all CTAs of the group) into the
//            per-frame terms and the F' matrices of the pair segments it owns;
//   stage 1  the CTA's warps stream their slice of the packed matches (16 B per match, coalesced float4), and
//            accumulate the error statistics and the per-pair 3x3 gradient G = sum d err / d F' in registers;
//            warp-reduce with 16 shuffles, then shared-memory atomics per segment;
//   stage 2a the per-pair adjoint (G -> per-frame gR, gA, intrinsics) and per-frame adjoint (-> gT, gq) are
//            applied locally; the CTA flushes 7 floats per touched frame + 4 scalars with global atomics;
//   barrier  one group-wide barrier (the only one per iteration; accumulators are triple-buffered);
//   stage 2b every CTA redundantly finishes the step from the summed gradient: early-exit test, gradient
//            mask, norm-relative clip, momentum, update of its own copy of the pose.
#pragma once
#include "geom.cuh"
#include "posediff_b200.h"

namespace pdb {

constexpr int kGgsThreads = 1024;
constexpr int kGgsMaxSeg = 128;   // pair segments handled per chunk by one CTA
constexpr int kGgsUnroll = 4;     // rounds (of 32 matches) in flight per warp
constexpr int kAccTail = 4;       // {g_fx', g_fy', clamp_sum, valid error sum (eval mode)}
constexpr int kSegAcc = 11;       // per-segment shared accumulators: G[9], clamp_sum, valid error sum

struct GgsProblem {
  const float4* pts;    // [rounds*32] (u1,v1,u2,v2), padded per segment to 32-row rounds
  const int4* segs;     // [nseg+1] {first_round, count, frame_a, frame_b}; sentinel {rounds,0,0,0}
  int nseg;
  int rounds;
  long long m_total;
  int frames;
  float height, width;
  float* pose;          // [frames*9] in/out
  float* gacc;          // [3][frames*7 + kAccTail], zero on entry
  int* gcnt;            // [3], zero on entry
  unsigned* bar;        // zero on entry
  pdb_ggs_stats* stats; // may be null
  float* dbg_grad;      // eval mode: [frames*9]
  float* dbg_scalars;   // eval mode: [4]
  float* dbg_F;         // eval mode, may be null: [nseg*9]
  float* dbg_G;         // eval mode, may be null: [nseg*9], zero on entry
};

struct GgsParams {
  int ctas_per_problem;
  int n_phases;
  int iters[PDB_GGS_PHASES];
  int flags[PDB_GGS_PHASES];  // bit0 R, bit1 T, bit2 FL
  float alpha, lr, smax, momentum;
  double min_matches;
};

inline size_t ggs_smem_bytes(int frames) {
  size_t f = 0;
  f += 2 * frames * 9;            // pose, velocity
  f += frames * (9 + 9 + 2 + 2);  // R, A, fl, inr
  f += frames * 18;               // gR, gA accumulators
  f += frames * 9;                // gradient scratch
  f += 32;                        // scalars
  f += kGgsMaxSeg * (9 + kSegAcc);  // F', segment accumulators
  size_t bytes = f * sizeof(float);
  bytes += kGgsMaxSeg * sizeof(int);         // segment valid counts
  bytes += (kGgsMaxSeg + 1) * sizeof(int4);  // segment descriptors
  return bytes + 64;
}

__device__ __forceinline__ void atomic_add_shared(float* p, float v) { atomicAdd(p, v); }

template <bool kEval>
__device__ __forceinline__ void ggs_body(const GgsProblem& pr, const GgsParams& P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kGgsThreads / 32;
  const int cpp = P.ctas_per_problem;
  const int cta = blockIdx.x % cpp;
  const int N = pr.frames, N9 = N * 9;
  const int acc_stride = N * 7 + kAccTail;

  // ---- shared memory carve-up ----
  int4* s_seg = reinterpret_cast<int4*>(smem_raw);
  float* s_pose = reinterpret_cast<float*>(s_seg + kGgsMaxSeg + 1);
  float* s_vel = s_pose + N9;
  float* s_R = s_vel + N9;
  float* s_A = s_R + N9;
  float* s_fl = s_A + N9;
  float* s_inr = s_fl + 2 * N;
  float* s_gR = s_inr + 2 * N;
  float* s_gA = s_gR + N9;
  float* s_grad = s_gA + N9;
  float* s_misc = s_grad + N9;  // [0..3] kin, [4..5] fpx, [6..9] gk, [10] clamp_sum, [16..] tail scalars
  float* s_F = s_misc + 32;
  float* s_sacc = s_F + kGgsMaxSeg * 9;
  int* s_scnt = reinterpret_cast<int*>(s_sacc + kGgsMaxSeg * kSegAcc);
  __shared__ int s_cta_cnt;

  // ---- static work partition: rounds -> CTAs -> warps ----
  const long long R = pr.rounds;
  const int r_cta0 = (int)(R * cta / cpp), r_cta1 = (int)(R * (cta + 1) / cpp);
  const int r_w0 = r_cta0 + (int)((long long)(r_cta1 - r_cta0) * warp / kWarps);
  const int r_w1 = r_cta0 + (int)((long long)(r_cta1 - r_cta0) * (warp + 1) / kWarps);
  auto seg_of_round = [&](int r) {  // last segment whose first_round <= r
    int lo = 0, hi = pr.nseg;       // segs[nseg].x == rounds > r
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (__ldg(&pr.segs[mid].x) <= r) lo = mid; else hi = mid;
    }
    return lo;
  };
  const bool cta_has_work = r_cta1 > r_cta0;
  const int seg_lo = cta_has_work ? seg_of_round(r_cta0) : 0;
  const int seg_hi = cta_has_work ? seg_of_round(r_cta1 - 1) : -1;  // inclusive
  const int wseg0 = (r_w1 > r_w0) ? seg_of_round(r_w0) : 0;

  for (int e = tid; e < N9; e += kGgsThreads) {
    s_pose[e] = pr.pose[e];
    s_vel[e] = 0.f;
  }
  const float scale = 0.5f * fminf(pr.height, pr.width);
  const float cx = 0.5f * pr.width, cy = 0.5f * pr.height;
  unsigned it_global = 0;
  __syncthreads();

  for (int phase = 0; phase < P.n_phases; ++phase) {
    const int flags = P.flags[phase];
    const bool upd_R = flags & 1, upd_T = flags & 2, upd_FL = flags & 4;
    const int iters = P.iters[phase];
    int done = 0, dropped = 0, last_valid = 0;
    float last_logged = __int_as_float(0x7fc00000);
    for (int iter = 0; iter < iters; ++iter) {
      float* acc = pr.gacc + (it_global % 3) * acc_stride;
      int* cnt = pr.gcnt + (it_global % 3);
      // ================= stage 0a: per-frame terms =================
      if (tid < N) {
        frame_forward(s_pose + tid * 9, s_R + tid * 9, s_A + tid * 9, s_fl + tid * 2, s_inr + tid * 2);
      }
      for (int e = tid; e < 2 * N9; e += kGgsThreads) s_gR[e] = 0.f;  // gR and gA are contiguous
      if (tid < 16) s_misc[tid] = 0.f;
      if (tid == 0) s_cta_cnt = 0;
      __syncthreads();
      if (warp == 0) {  // shared focal length: mean over frames (geometry_guided_sampling.py:142)
        float fx = 0.f, fy = 0.f;
        for (int n = lane; n < N; n += 32) {
          fx += s_fl[n * 2];
          fy += s_fl[n * 2 + 1];
        }
        fx = warp_sum(fx) / (float)N * scale;
        fy = warp_sum(fy) / (float)N * scale;
        if (lane == 0) {
          s_misc[0] = 1.f / fx;
          s_misc[1] = 1.f / fy;
          s_misc[2] = -cx / fx;
          s_misc[3] = -cy / fy;
          s_misc[4] = fx;
          s_misc[5] = fy;
        }
      }
      __syncthreads();
      // ================= chunks of <= kGgsMaxSeg pair segments =================
      for (int cs = seg_lo; cs <= seg_hi; cs += kGgsMaxSeg) {
        const int ce = min(cs + kGgsMaxSeg, seg_hi + 1);
        const int nchunk = ce - cs;
        // ---- stage 0b: F' per segment ----
        if (tid <= nchunk) s_seg[tid] = __ldg(&pr.segs[cs + tid]);
        for (int e = tid; e < nchunk * kSegAcc; e += kGgsThreads) s_sacc[e] = 0.f;
        if (tid < nchunk) s_scnt[tid] = 0;
        if (tid < nchunk) {
          const int4 sd = __ldg(&pr.segs[cs + tid]);
          float F[9];
          pair_F(s_R + sd.z * 9, s_A + sd.z * 9, s_R + sd.w * 9, s_A + sd.w * 9, s_misc, sd.z == sd.w, F);
#pragma unroll
          for (int k = 0; k < 9; ++k) s_F[tid * 9 + k] = F[k];
          if (kEval && pr.dbg_F) {  // several CTAs may share a segment: they write identical values
#pragma unroll
            for (int k = 0; k < 9; ++k) pr.dbg_F[(size_t)(cs + tid) * 9 + k] = F[k];
          }
        }
        __syncthreads();
        // ---- stage 1: stream the matches ----
        {
          int s = max(cs, wseg0);
          int r = (s < ce) ? max(r_w0, s_seg[s - cs].x) : r_w1;
          while (s < ce && r < r_w1) {
            const int4 sd = s_seg[s - cs];
            const int r_end = min(r_w1, s_seg[s - cs + 1].x);
            float Fm[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) Fm[k] = s_F[(s - cs) * 9 + k];
            float g[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) g[k] = 0.f;
            int nval = 0;
            const int seg_first = sd.x, seg_count = sd.y;
            for (; r < r_end; r += kGgsUnroll) {
              float4 pt[kGgsUnroll];
#pragma unroll
              for (int u = 0; u < kGgsUnroll; ++u) {
                if (r + u < r_end) pt[u] = ld_stream_f4(pr.pts + (size_t)(r + u) * 32 + lane);
                else pt[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int u = 0; u < kGgsUnroll; ++u) {
                const bool inb = (r + u < r_end) && ((r + u - seg_first) * 32 + lane < seg_count);
                nval += sampson_match<kEval>(pt[u], Fm, inb, P.smax, g);
              }
            }
            // warp reduction: 16 shuffles for the 10 float slots, one redux for the count
            const float tot = warp_reduce16(g, lane);
            const int slot = warp_reduce16_slot(lane);
            if (!(lane & 1) && slot < kSegAcc) atomic_add_shared(&s_sacc[(s - cs) * kSegAcc + slot], tot);
            nval = __reduce_add_sync(0xffffffffu, nval);
            if (lane == 0 && nval) atomicAdd(&s_scnt[s - cs], nval);
            r = r_end;
            ++s;
          }
        }
        __syncthreads();
        // ---- stage 2a: per-pair adjoint into per-frame shared accumulators ----
        if (tid < nchunk) {
          const int4 sd = s_seg[tid];
          float G[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) G[k] = s_sacc[tid * kSegAcc + k];
          auto add = [](float* p, float v) { atomicAdd(p, v); };
          pair_adjoint(s_R + sd.z * 9, s_A + sd.z * 9, s_R + sd.w * 9, s_A + sd.w * 9, s_misc, sd.z == sd.w, G,
                       s_gR + sd.z * 9, s_gA + sd.z * 9, s_gR + sd.w * 9, s_gA + sd.w * 9, s_misc + 6, add);
          atomicAdd(&s_misc[10], s_sacc[tid * kSegAcc + 9]);
          if (kEval) atomicAdd(&s_misc[11], s_sacc[tid * kSegAcc + 10]);
          atomicAdd(&s_cta_cnt, s_scnt[tid]);
          if (kEval && pr.dbg_G) {
#pragma unroll
            for (int k = 0; k < 9; ++k) atomicAdd(&pr.dbg_G[(size_t)(cs + tid) * 9 + k], G[k]);
          }
        }
        __syncthreads();
      }
      // ================= flush this CTA's contribution =================
      if (tid < N) {
        const float* gR = s_gR + tid * 9;
        const float* gA = s_gA + tid * 9;
        bool touched = false;
#pragma unroll
        for (int k = 0; k < 9; ++k) touched |= (gR[k] != 0.f) | (gA[k] != 0.f);
        if (touched) {
          float gT[3], gq[4];
          frame_adjoint(s_pose + tid * 9, s_R + tid * 9, gR, gA, gT, gq);
#pragma unroll
          for (int k = 0; k < 3; ++k) atomicAdd(&acc[tid * 7 + k], gT[k]);
#pragma unroll
          for (int k = 0; k < 4; ++k) atomicAdd(&acc[tid * 7 + 3 + k], gq[k]);
        }
      } else if (tid == kGgsThreads - 32 && cta_has_work) {
        const float fx = s_misc[4], fy = s_misc[5];
        atomicAdd(&acc[N * 7 + 0], (-s_misc[6] + cx * s_misc[8]) / (fx * fx));
        atomicAdd(&acc[N * 7 + 1], (-s_misc[7] + cy * s_misc[9]) / (fy * fy));
        atomicAdd(&acc[N * 7 + 2], s_misc[10]);
        if (kEval) atomicAdd(&acc[N * 7 + 3], s_misc[11]);
        if (s_cta_cnt) atomicAdd(cnt, s_cta_cnt);
      }
      // ================= group barrier =================
      group_barrier(pr.bar, (it_global + 1) * (unsigned)cpp);
      // recycle the accumulator used two iterations from now (nobody reads or writes it at this point)
      if (cta == 0) {
        float* old = pr.gacc + ((it_global + 2) % 3) * acc_stride;
        for (int e = tid; e < acc_stride; e += kGgsThreads) old[e] = 0.f;
        if (tid == 0) pr.gcnt[(it_global + 2) % 3] = 0;
      }
      ++it_global;
      // ================= stage 2b: finish the step (identical in every CTA) =================
      const int n_valid = __ldcg(cnt);
      last_valid = n_valid;
      last_logged = __ldcg(&acc[N * 7 + 2]) / (float)pr.m_total;
      const bool drop = (P.min_matches > 0.0) && ((double)n_valid / (double)N < P.min_matches);
      if (drop && !kEval) {
        dropped = 1;
        break;
      }
      for (int e = tid; e < N9; e += kGgsThreads) {
        const int n = e / 9, c = e - n * 9;
        float gsum;
        if (c < 3) gsum = upd_T ? __ldcg(&acc[n * 7 + c]) : 0.f;
        else if (c < 7) gsum = upd_R ? __ldcg(&acc[n * 7 + c]) : 0.f;
        else gsum = upd_FL ? __ldcg(&acc[N * 7 + (c - 7)]) * (scale / (float)N) * s_fl[n * 2 + (c - 7)] * s_inr[n * 2 + (c - 7)] : 0.f;
        s_grad[e] = gsum / (float)n_valid;
      }
      __syncthreads();
      if (kEval) {
        if (cta == 0) {
          for (int e = tid; e < N9; e += kGgsThreads) pr.dbg_grad[e] = s_grad[e];
          if (tid == 0) {
            pr.dbg_scalars[0] = __ldcg(&acc[N * 7 + 3]) / (float)n_valid;
            pr.dbg_scalars[1] = (float)n_valid;
            pr.dbg_scalars[2] = last_logged;
            pr.dbg_scalars[3] = 0.f;
          }
        }
        break;
      }
      if (warp == 0) {
        float gn2 = 0.f, pn2 = 0.f;
        for (int e = lane; e < N9; e += 32) {
          const float gv = s_grad[e];
          gn2 = fmaf(gv, gv, gn2);
          const float pm = (fabsf(gv) > 0.f) ? s_pose[e] : 0.f;  // grad_mask = grads.abs() > 0 (:117)
          pn2 = fmaf(pm, pm, pn2);
        }
        gn2 = warp_sum(gn2);
        pn2 = warp_sum(pn2);
        if (lane == 0) {
          const float max_norm = P.alpha * sqrtf(pn2) / P.lr;      // :119
          const float coef = max_norm / (sqrtf(gn2) + 1e-6f);      // clip_grad_norm_
          s_misc[16] = (coef > 1.0f) ? 1.0f : coef;                // clamp(max=1), NaN passes through
        }
      }
      __syncthreads();
      {
        const float coef = s_misc[16];
        for (int e = tid; e < N9; e += kGgsThreads) {
          const float gv = s_grad[e] * coef;
          const float v = (done == 0) ? gv : fmaf(P.momentum, s_vel[e], gv);  // SGD momentum buffer, reset per phase
          s_vel[e] = v;
          s_pose[e] = s_pose[e] - P.lr * v;
        }
      }
      ++done;
      __syncthreads();
    }
    if (kEval) break;
    if (cta == 0 && tid == 0 && pr.stats) {
      pr.stats->sampson[phase] = last_logged;
      pr.stats->iters[phase] = done;
      pr.stats->dropped[phase] = dropped;
      pr.stats->n_valid[phase] = last_valid;
    }
    __syncthreads();
  }
  if (!kEval && cta == 0) {
    for (int e = tid; e < N9; e += kGgsThreads) pr.pose[e] = s_pose[e];
  }
}

}  // namespace pdb
