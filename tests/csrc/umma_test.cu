// umma_test.cu - single-CTA probe of the tcgen05 GEMM primitives (umma.cuh): loads two fp32 tiles, runs one 3xTF32
// GEMM in each operand-major mode and dumps all 128 TMEM lanes, so the tests can verify descriptors, the format-F
// swizzle and the TMEM row mapping for M = 64 and M = 128 against torch.
#include "umma.cuh"
#include "scgib.h"

namespace scgib {
using namespace umma;

// mode 0: D[M][64] = A[M][64] * B[64][64]^T      (A K-major, B K-major: B stored [n][k])
// mode 1: D[M][64] = A[M][64] * B[64][64]        (A K-major, B MN-major: B stored [k][n])
// mode 2: D[64][64] = A[R][64]^T * B[R][64]      (both MN-major, reduction over the R = M tile rows)
__global__ void __launch_bounds__(128) umma_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                         float* __restrict__ out, int M, int mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int RA = M, RB = (mode == 2) ? M : 64;
  constexpr int TB = tile_bytes(128, 64);
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + TB;
  unsigned char* b_hi = a_lo + TB;
  unsigned char* b_lo = b_hi + TB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&s_tmem, 64);
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  const bool v32 = true;
  for (int i = threadIdx.x; i < RA * 16; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(A + (size_t)(i >> 4) * 64 + (i & 15) * 4);
    if (v32 && mode == 2) {
      store_split4_s(a_hi, a_lo, RA, i >> 4, i & 15, v);
    } else store_split4(a_hi, a_lo, 64, i >> 4, i & 15, v);
  }
  for (int i = threadIdx.x; i < RB * 16; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(B + (size_t)(i >> 4) * 64 + (i & 15) * 4);
    if (v32 && mode != 0) {
      store_split4_s(b_hi, b_lo, RB, i >> 4, i & 15, v);
    } else store_split4(b_hi, b_lo, 64, i >> 4, i & 15, v);
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = s_tmem;
  if (threadIdx.x == 0) {
    Operand a{smem_u32(a_hi), smem_u32(a_lo), mode == 2, (v32 && mode == 2) ? (uint32_t)RA * 128 : (uint32_t)group_bytes(64), kCoreStride};
    Operand b{smem_u32(b_hi), smem_u32(b_lo), mode != 0, (v32 && mode != 0) ? (uint32_t)RB * 128 : (uint32_t)group_bytes(64), kCoreStride};
    const int Mi = (mode == 2) ? 64 : M;
    const int ksteps = (mode == 2) ? M / 8 : 8;
    const uint32_t idesc = idesc_tf32(Mi, 64, mode == 2, mode != 0);
    gemm_3xtf32(tbase, a, b, ksteps, idesc, false);
    mma_commit(&s_bar);
  }
  mbar_wait(&s_bar, 0);
  fence_after_sync();
  for (int c = 0; c < 4; ++c) {
    float v[16];
    tmem_ld16(tmem_addr(tbase, 32 * warp, 16 * c), v);
#pragma unroll
    for (int i = 0; i < 16; ++i) out[(size_t)(32 * warp + lane) * 64 + 16 * c + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 64);
}

}  // namespace scgib

extern "C" SCGIB_API int scgib_debug_umma(const float* A, const float* B, float* out, int32_t M, int32_t mode, void* stream) {
  if (!A || !B || !out) return SCGIB_E_NULL;
  if ((M != 64 && M != 128) || mode < 0 || mode > 2) return SCGIB_E_SHAPE;
  const int smem = 4 * scgib::umma::tile_bytes(128, 64) + 1024;
  cudaFuncSetAttribute(scgib::umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  scgib::umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, out, M, mode);
  return (int)cudaGetLastError();
}
