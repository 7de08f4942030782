// TEST HARNESS (not product): runs the GGS kernel body of posediffusion_b200/csrc/ggs.cuh -- the same source nvcc compiles
// for sm_100a -- on the CPU through the execution-model emulation of cuda_emu.h, for any template variant
// (kEval, kPaired) and any of the three streaming modes.  Build:  g++ -O1 -std=c++17 -shared -fPIC (see __graft_entry__.py).
#include "cuda_emu.h"

#include "../../posediffusion_b200/csrc/ggs.cuh"
#include "../../posediffusion_b200/csrc/align.cuh"

// ---------------------------------------------------------------------------------------------------------------
// emulator runtime
// ---------------------------------------------------------------------------------------------------------------
namespace emu {
thread_local Cta* g_cta = nullptr;
thread_local uint3 g_threadIdx = {0, 0, 0}, g_blockIdx = {0, 0, 0};
thread_local dim3 g_blockDim(1, 1, 1), g_gridDim(1, 1, 1);

static void trampoline() {
  Cta* c = g_cta;
  const int me = c->current;
  c->body();
  c->done[me] = 1;
  c->live -= 1;
  // a finished thread no longer takes part in block barriers: release one that is now complete
  if (c->live > 0 && c->block_count == c->live) {
    c->block_count = 0;
    c->block_gen++;
  }
  swapcontext(&c->ctx[me], &c->sched);
}

static void run_cta(int block_index, int grid, int block, const std::function<void()>& body) {
  Cta cta;
  cta.nthreads = block;
  cta.nwarps = (block + kWarp - 1) / kWarp;
  cta.ctx.resize(block);
  cta.stack.resize(block);
  cta.done.assign(block, 0);
  cta.live = block;
  cta.warp_count.assign(cta.nwarps, 0);
  cta.warp_gen.assign(cta.nwarps, 0);
  cta.xch.assign((size_t)cta.nwarps * kWarp, 0);
  cta.body = body;
  cta.smem = static_cast<unsigned char*>(aligned_alloc(1024, kSharedBytes));  // the CTA's dynamic shared memory
  memset(cta.smem, 0xcd, kSharedBytes);  // shared memory is NOT zero-initialised on the device either
  g_cta = &cta;
  g_blockIdx = {(unsigned)block_index, 0, 0};
  g_blockDim = dim3(block, 1, 1);
  g_gridDim = dim3(grid, 1, 1);
  for (int t = 0; t < block; ++t) {
    cta.stack[t] = static_cast<char*>(malloc(kStackBytes));
    getcontext(&cta.ctx[t]);
    cta.ctx[t].uc_stack.ss_sp = cta.stack[t];
    cta.ctx[t].uc_stack.ss_size = kStackBytes;
    cta.ctx[t].uc_link = &cta.sched;
    makecontext(&cta.ctx[t], trampoline, 0);
  }
  while (cta.live > 0) {
    for (int t = 0; t < block; ++t) {
      if (cta.done[t]) continue;
      cta.current = t;
      g_threadIdx = {(unsigned)t, 0, 0};
      swapcontext(&cta.sched, &cta.ctx[t]);
    }
  }
  for (int t = 0; t < block; ++t) free(cta.stack[t]);
  free(cta.smem);
  g_cta = nullptr;
}

void launch(int grid, int block, const std::function<void()>& body) {
  std::vector<std::thread> ctas;
  for (int b = 0; b < grid; ++b) ctas.emplace_back(run_cta, b, grid, block, std::cref(body));
  for (auto& t : ctas) t.join();
}
}  // namespace emu

// ---------------------------------------------------------------------------------------------------------------
// one geometry_guided_sampling / compute_sampson_distance call on the emulated grid
// ---------------------------------------------------------------------------------------------------------------
extern "C" int ggs_emu_run(const float* pts, const int* segs /*[nseg+1][4] incl. sentinel*/, int nseg, int rounds, long long m_total,
                           int frames, float height, float width, float* pose /*[frames*9] in/out*/, int paired, int eval, int cpp,
                           int force_stream, int no_ring, const int* iters, const int* flags, int n_phases, float alpha, float lr,
                           float smax, float momentum, double min_matches, float* dbg_grad, float* dbg_scalars, float* dbg_F,
                           float* dbg_G, pdb_ggs_stats* stats, int* mode_out) {
  using namespace pdb;
  if (frames < 1 || frames > kMaxFrames || cpp < 1 || n_phases < 1 || n_phases > PDB_GGS_PHASES) return -1;
  const size_t acc_floats = 3 * (size_t)(frames * 7 + kAccTail) * kAccPad;
  std::vector<float> gacc(acc_floats, 0.f);
  std::vector<int> gcnt(4, 0);
  std::vector<unsigned> bar(4, 0u);
  GgsProblem pr = {};
  pr.pts = reinterpret_cast<const float4*>(pts);
  pr.segs = reinterpret_cast<const int4*>(segs);
  pr.nseg = nseg;
  pr.rounds = rounds;
  pr.m_total = m_total;
  pr.frames = frames;
  pr.height = height;
  pr.width = width;
  pr.pose = pose;
  pr.gacc = gacc.data();
  pr.gcnt = gcnt.data();
  pr.bar = bar.data();
  pr.stats = stats;
  pr.dbg_grad = dbg_grad;
  pr.dbg_scalars = dbg_scalars;
  pr.dbg_F = dbg_F;
  pr.dbg_G = dbg_G;
  GgsParams P = {};
  P.ctas_per_problem = cpp;
  P.n_phases = n_phases;
  for (int i = 0; i < n_phases; ++i) {
    P.iters[i] = iters[i];
    P.flags[i] = flags[i];
  }
  P.alpha = alpha;
  P.lr = lr;
  P.smax = smax;
  P.momentum = momentum;
  P.min_matches = min_matches;
  // the launch logic of api_core.cu::launch_ggs_chunk: shared-memory-resident slice when it fits, else the bulk-async ring
  const size_t fixed = ggs_smem_fixed_bytes(frames);
  const size_t budget = emu::kSharedBytes > fixed + 1024 ? emu::kSharedBytes - fixed - 1024 : 0;
  const long long rounds_per_cta = ggs_rounds_per_cta(rounds > 0 ? rounds : 1, cpp, paired != 0);
  const bool resident = (size_t)rounds_per_cta * 512 <= budget && !force_stream;
  P.resident_rounds = resident ? (int)rounds_per_cta : 0;
  P.ring = (resident || no_ring) ? 0 : 1;
  if (mode_out) *mode_out = resident ? 0 : (P.ring ? 1 : 2);
  if (fixed + (resident ? (size_t)rounds_per_cta * 512 : (size_t)kRingBytes) > emu::kSharedBytes) return -2;
  const std::function<void()> body = [&]() {
    if (eval) {
      if (paired) ggs_body<true, true>(pr, P); else ggs_body<true, false>(pr, P);
    } else {
      if (paired) ggs_body<false, true>(pr, P); else ggs_body<false, false>(pr, P);
    }
  };
  emu::launch(cpp, kGgsThreads, body);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// camera alignment (csrc/align.cuh): the two kernel bodies with the launch geometry of pdb_cameras_align
// ---------------------------------------------------------------------------------------------------------------
extern "C" void cameras_align_emu(const float* Rs, const float* Ts, const float* Rt, const float* Tt, int count, int estimate_scale, float eps,
                                  float* Ro, float* To, float* align) {
  emu::launch(1, 32, [&]() { pdb::cameras_align_estimate_warp(Rs, Ts, Rt, Tt, count, estimate_scale, eps, align); });
  emu::launch((count + 127) / 128, 128, [&]() { pdb::cameras_align_apply_thread(align, Rs, Ts, count, Ro, To); });
}
