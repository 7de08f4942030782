// Microbenchmark (GPU box only): latency of one grid-wide all-reduce of a small fp32 vector between the persistent CTAs of
// a cooperative launch -- the exchange that closes every inner iteration of the GGS kernel (csrc/ggs.cuh).
//
//   mode 0  round-1 scheme: red.global.add into one 128-byte line per value, release/acquire counter barrier, read-back
//   mode 1  flag-carrying words ("LL": one 64-bit store = {fp32 value, iteration tag}), ONE level: every CTA reads all slots
//   mode 2  LL, TWO levels: groups of S CTAs, the group leader sums its group's slots and publishes a group slot,
//           every CTA then reads the G group slots
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/xchg_probe.cu -o build/xchg_probe
//   build/xchg_probe [words=148] [iters=20000]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ void st_ll(unsigned long long* p, float v, unsigned tag) {
  const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_ll(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct Params {
  int mode, words, iters, S, G, replicas;
  float* acc;                 // mode 0: [3][words*32]
  unsigned* bar;              // mode 0
  unsigned long long* slot1;  // [2][ctas][words]
  unsigned long long* slot2;  // [2][replicas][G][words]
  float* out;                 // [ctas] checksum
  long long* cycles;          // [ctas]
};

__global__ void __launch_bounds__(512, 1) xchg_kernel(const Params P) {
  extern __shared__ float s_sum[];
  const int tid = threadIdx.x, cta = blockIdx.x, C = gridDim.x;
  float mine = 1.0f + 1e-3f * (float)(cta % 7);  // this CTA's contribution per word (changes per iteration below)
  float check = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < P.iters; ++it) {
    const unsigned tag = (unsigned)it + 1u;
    const float contrib = mine + (float)(it & 3);
    if (P.mode == 0) {
      float* acc = P.acc + (size_t)(it % 3) * P.words * 32;
      if (tid < 32)
        for (int e = tid; e < P.words; e += 32) atomicAdd(&acc[e * 32], contrib);
      if (tid < 32) {
        __syncwarp();
        if (tid == 0) {
          red_release_add_u32(P.bar, 1u);
          while (ld_acquire_u32(P.bar) < tag * (unsigned)C) {
          }
        }
        __syncwarp();
        if (cta == 0) {
          float* old = P.acc + (size_t)((it + 2) % 3) * P.words * 32;
          for (int e = tid; e < P.words; e += 32) old[e * 32] = 0.f;
        }
        for (int e = tid; e < P.words; e += 32) s_sum[e] = __ldcg(&acc[e * 32]);
      }
      __syncthreads();
    } else if (P.mode == 1) {
      unsigned long long* mys = P.slot1 + ((size_t)(it & 1) * C + cta) * P.words;
      for (int e = tid; e < P.words; e += blockDim.x) st_ll(mys + e, contrib, tag);
      const unsigned long long* base = P.slot1 + (size_t)(it & 1) * C * P.words;
      for (int e = tid; e < P.words; e += blockDim.x) {
        float sum = 0.f;
        for (int c0 = 0; c0 < C; c0 += 8) {
          unsigned long long w[8];
          bool ok;
          do {
            ok = true;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (c0 + k < C) {
                w[k] = ld_ll(base + (size_t)(c0 + k) * P.words + e);
                ok &= (unsigned)(w[k] >> 32) == tag;
              }
          } while (!ok);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (c0 + k < C) sum += __uint_as_float((unsigned)w[k]);
        }
        s_sum[e] = sum;
      }
      __syncthreads();
    } else {
      const int S = P.S, G = P.G;
      unsigned long long* mys = P.slot1 + ((size_t)(it & 1) * C + cta) * P.words;
      for (int e = tid; e < P.words; e += blockDim.x) st_ll(mys + e, contrib, tag);
      const int g = cta / S;
      if (cta % S == 0) {  // group leader: sum the group's slots, publish
        const int c_lo = g * S, c_hi = min(C, c_lo + S);
        const unsigned long long* base = P.slot1 + (size_t)(it & 1) * C * P.words;
        for (int e = tid; e < P.words; e += blockDim.x) {
          float sum = 0.f;
          for (int c0 = c_lo; c0 < c_hi; c0 += 8) {
            unsigned long long w[8];
            bool ok;
            do {
              ok = true;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (c0 + k < c_hi) {
                  w[k] = ld_ll(base + (size_t)(c0 + k) * P.words + e);
                  ok &= (unsigned)(w[k] >> 32) == tag;
                }
            } while (!ok);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (c0 + k < c_hi) sum += __uint_as_float((unsigned)w[k]);
          }
          for (int r = 0; r < P.replicas; ++r)
            st_ll(P.slot2 + (((size_t)(it & 1) * P.replicas + r) * G + g) * P.words + e, sum, tag);
        }
      }
      const unsigned long long* base2 = P.slot2 + ((size_t)(it & 1) * P.replicas + (cta % P.replicas)) * G * P.words;
      for (int e = tid; e < P.words; e += blockDim.x) {
        float sum = 0.f;
        for (int g0 = 0; g0 < G; g0 += 8) {
          unsigned long long w[8];
          bool ok;
          do {
            ok = true;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (g0 + k < G) {
                w[k] = ld_ll(base2 + (size_t)(g0 + k) * P.words + e);
                ok &= (unsigned)(w[k] >> 32) == tag;
              }
          } while (!ok);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (g0 + k < G) sum += __uint_as_float((unsigned)w[k]);
        }
        s_sum[e] = sum;
      }
      __syncthreads();
    }
    if (tid == 0) {
      check += s_sum[0] + s_sum[P.words - 1];
      mine += 1e-6f * s_sum[0];  // the next contribution depends on this result (no overlap across iterations)
    }
    mine = __shfl_sync(0xffffffffu, mine, 0);
    __syncthreads();
    if (tid >= 32) mine = 0.f;  // only warp 0's value is used; keep the dependency simple
    if (tid == 0) s_sum[P.words] = mine;
    __syncthreads();
    mine = s_sum[P.words];
    __syncthreads();
  }
  if (tid == 0) {
    P.out[cta] = check;
    P.cycles[cta] = clock64() - t0;
  }
}

int main(int argc, char** argv) {
  const int words = argc > 1 ? atoi(argv[1]) : 148;
  const int iters = argc > 2 ? atoi(argv[2]) : 20000;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
  // (a) how many CTAs of the GGS shape (512 threads, ~200 KB dynamic shared memory) can be co-resident as clusters?
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / cs * cs);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (cs > 8) cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, xchg_kernel, &cfg);
    printf("cluster size %2d: max active clusters %d (%d CTAs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  cudaGetLastError();
  Params P = {};
  P.words = words;
  P.iters = iters;
  const int maxC = sms;
  CK(cudaMalloc(&P.acc, sizeof(float) * 3 * words * 32));
  CK(cudaMalloc(&P.bar, 256));
  CK(cudaMalloc(&P.slot1, sizeof(unsigned long long) * 2 * maxC * words));
  CK(cudaMalloc(&P.slot2, sizeof(unsigned long long) * 2 * 8 * 64 * words));
  CK(cudaMalloc(&P.out, sizeof(float) * maxC));
  CK(cudaMalloc(&P.cycles, sizeof(long long) * maxC));
  CK(cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  struct Case { int mode, C, S, rep; };
  std::vector<Case> cases;
  for (int C : {18, 37, 74, 148}) {
    cases.push_back({0, C, 0, 1});
    cases.push_back({1, C, 0, 1});
    for (int S : {4, 8, 12, 16, 24})
      if (S < C) cases.push_back({2, C, S, 1});
  }
  cases.push_back({2, 148, 12, 2});
  cases.push_back({2, 148, 12, 4});
  cases.push_back({2, 148, 8, 4});
  for (const Case& c : cases) {
    P.mode = c.mode;
    P.S = c.S;
    P.G = c.S ? (c.C + c.S - 1) / c.S : 0;
    P.replicas = c.rep;
    CK(cudaMemset(P.acc, 0, sizeof(float) * 3 * words * 32));
    CK(cudaMemset(P.bar, 0, 256));
    CK(cudaMemset(P.slot1, 0, sizeof(unsigned long long) * 2 * maxC * words));
    CK(cudaMemset(P.slot2, 0, sizeof(unsigned long long) * 2 * 8 * 64 * words));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    void* args[] = {&P};
    float best = 1e30f;
    float chk = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaMemset(P.bar, 0, 256));
      CK(cudaMemset(P.acc, 0, sizeof(float) * 3 * words * 32));
      CK(cudaMemset(P.slot1, 0, sizeof(unsigned long long) * 2 * maxC * words));
      CK(cudaMemset(P.slot2, 0, sizeof(unsigned long long) * 2 * 8 * 64 * words));
      CK(cudaEventRecord(a));
      CK(cudaLaunchCooperativeKernel((void*)xchg_kernel, dim3(c.C), dim3(512), args, 64 * 1024, 0));
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      float ms;
      CK(cudaEventElapsedTime(&ms, a, b));
      if (ms < best) best = ms;
      CK(cudaMemcpy(&chk, P.out, sizeof(float), cudaMemcpyDeviceToHost));
    }
    printf("mode %d  ctas %3d  S %2d G %2d rep %d : %8.3f us per all-reduce   (check %.3f)\n", c.mode, c.C, c.S, P.G, c.rep,
           best * 1e3f / iters, chk);
  }
  return 0;
}
