// Tensor-core (tcgen05 + TMA) linear layer: host side (tensor maps, launch) and a test entry point.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "context.cuh"
#include "tc_linear.cuh"

using namespace pdb;

namespace {

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// 2-D fp32 row-major [rows, K] tensor, box = box_rows x 32 floats (128 bytes), 128-byte swizzle, zero fill out of bounds
int make_map(Context* ctx, CUtensorMap* map, const float* ptr, int rows, int K, int box_rows) {
  auto encode = get_encode();
  if (!encode) return ctx->fail(PDB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return ctx->fail(PDB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return PDB_OK;
}

}  // namespace

namespace pdb {

// Y = epilogue(X[S,K] @ W[O,K]^T) on the tensor cores.  K % 32 == 0, O % 64 == 0, all pointers 16-byte aligned.
int enqueue_tc_linear(Context* ctx, const float* X, const float* W, TcEpilogue E, cudaStream_t st) {
  if (E.K % kTcBK || E.O % 64 || E.S < 1) return ctx->fail(PDB_ERR_INVALID, "tc_linear shape (S=%d, O=%d, K=%d)", E.S, E.O, E.K);
  constexpr int BN = 64;
  CUtensorMap mx, mw;
  if (int rc = make_map(ctx, &mx, X, E.S, E.K, kTcBM)) return rc;
  if (int rc = make_map(ctx, &mw, W, E.O, E.K, BN)) return rc;
  const size_t smem = tc_smem_bytes(BN);
  static bool attr = false;
  if (!attr) {
    PDB_CUDA(ctx, cudaFuncSetAttribute(tc_linear_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid(E.O / BN, (E.S + kTcBM - 1) / kTcBM);
  {
    ScopedTimer timer(ctx, st, 1);
    tc_linear_kernel<BN><<<grid, kTcThreads, smem, st>>>(mx, mw, E);
  }
  PDB_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  return PDB_OK;
}

}  // namespace pdb

extern "C" int pdb_debug_tc_linear(pdb_context* c, const float* x_dev, const float* w_dev, const float* bias_dev,
                                   const float* residual_dev, float* y_dev, int32_t S, int32_t O, int32_t K, int32_t relu,
                                   void* stream) {
  if (!c) return PDB_ERR_INVALID;
  Context* ctx = reinterpret_cast<Context*>(c);
  if (!x_dev || !w_dev || !y_dev) return ctx->fail(PDB_ERR_INVALID, "null argument");
  PDB_CUDA(ctx, cudaSetDevice(ctx->device));
  TcEpilogue E = {};
  E.bias = bias_dev;
  E.residual = residual_dev;
  E.ldr = O;
  E.Y = y_dev;
  E.ldy = O;
  E.S = S;
  E.O = O;
  E.K = K;
  E.relu = relu;
  return enqueue_tc_linear(ctx, x_dev, w_dev, E, static_cast<cudaStream_t>(stream));
}
