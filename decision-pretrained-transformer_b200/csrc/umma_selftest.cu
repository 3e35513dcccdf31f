// Bring-up / regression test of the tcgen05 helpers in umma.cuh: D[128,N] = A[128,K] * B[N,K]^T with bf16
// operands staged in shared memory (K64 tiles, 128 B swizzle) and the fp32 accumulator in tensor memory.
#include "common.cuh"
#include "umma.cuh"

namespace dpt {

__global__ void __launch_bounds__(128) umma_selftest_kernel(const float* A, const float* B, float* D, int N, int K) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int kt = (K + 63) / 64;                      // K64 tiles per operand
  unsigned char* a_tiles = smem;                     // kt tiles of 128 rows (16 KB each)
  unsigned char* b_tiles = smem + kt * 16384;        // kt tiles of N rows (N * 128 B each, N % 8 == 0)
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) umma::mbar_init(&bar, 1);
  // stage operands: thread t writes row t of A; rows of B are spread over threads
  for (int c = 0; c < K / 8; ++c) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = A[(size_t)tid * K + 8 * c + i];
    umma::st_chunk(a_tiles + (c >> 3) * 16384, tid, c & 7, v);
  }
  for (int n = tid; n < N; n += 128)
    for (int c = 0; c < K / 8; ++c) {
      float v[8];
      for (int i = 0; i < 8; ++i) v[i] = B[(size_t)n * K + 8 * c + i];
      umma::st_chunk(b_tiles + (c >> 3) * (N * 128), n, c & 7, v);
    }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = umma::make_idesc_bf16(128, N);
    for (int ks = 0; ks < K / 16; ++ks) {
      const int tile = ks >> 2, sub = ks & 3;
      const uint64_t ad = umma::make_desc_k64(umma::smem_u32(a_tiles + tile * 16384) + sub * 32);
      const uint64_t bd = umma::make_desc_k64(umma::smem_u32(b_tiles + tile * (N * 128)) + sub * 32);
      umma::mma_bf16(tbase, ad, bd, idesc, ks > 0);
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    umma::tmem_ld32(umma::tmem_addr(tbase, warp, c0), v);
    for (int i = 0; i < 32 && c0 + i < N; ++i) D[(size_t)tid * N + c0 + i] = v[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 256);
}

}  // namespace dpt

using namespace dpt;

extern "C" int dpt_debug_umma_gemm(const float* A, const float* B, float* D, int N, int K, void* stream) {
  DPT_CHECK_ARG(A && B && D, "dpt_debug_umma_gemm: null pointer");
  DPT_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0, "dpt_debug_umma_gemm: N=%d must be a multiple of 16 in [16,256]", N);
  DPT_CHECK_ARG(K >= 16 && K <= 128 && K % 16 == 0, "dpt_debug_umma_gemm: K=%d must be a multiple of 16 in [16,128]", K);
  const int kt = (K + 63) / 64;
  const size_t smem = (size_t)kt * 16384 + (size_t)kt * N * 128 + 1024;
  DPT_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K);
  DPT_LAUNCH_CHECK();
  return DPT_OK;
}
