// umma_probe_bf16.cu - runtime-parametrised probe of the tcgen05 kind::f16 (bf16) operand formats: shared-memory tile
// layouts (128B / 64B swizzle, no-swizzle cores), K-major and MN-major descriptors, M = 64 / 128, N up to 256, the A
// operand in tensor memory (packed bf16x2).  Test infrastructure: the formats used by gin_bf16*.cu were selected with
// this probe on a B200 (tests/gpu_umma_probe_bf16.py).
#include <string.h>
#include "umma.cuh"
#include "scgib.h"

namespace scgib {
using namespace umma;

struct ProbeB {
  int M, N, ksteps;                   // K = 16 per step
  int a_src;                          // 0 = shared memory, 2 = tensor memory (M = 128), 3 = tensor memory, halves swapped
  int a_layout, b_layout;             // 0 = SW128 rows (128 B), 1 = SW64 rows (64 B), 2 = no-swizzle 8x16B cores
  int a_p0, a_p1, b_p0, b_p1;         // layout 0/1: p0 = column-block stride; layout 2: p0 = core stride, p1 = 8-row group stride
  int a_mn, b_mn;
  int a_lbo, a_sbo, a_ltype, a_div, a_adv_lo, a_adv_hi;
  int b_lbo, b_sbo, b_ltype, b_div, b_adv_lo, b_adv_hi;
  int RA, CA, RB, CB;                 // source matrices (fp32, row-major [R][C])
  int reps;
};

__device__ __forceinline__ int probe_off(int layout, int p0, int p1, int r, int c) {
  if (layout == 0) { const int blk = c >> 6, cc = c & 63; return blk * p0 + r * 128 + ((((cc >> 3) ^ (r & 7))) << 4) + (cc & 7) * 2; }
  if (layout == 1) { const int blk = c >> 5, cc = c & 31; return blk * p0 + r * 64 + ((((cc >> 3) ^ ((r >> 1) & 3))) << 4) + (cc & 7) * 2; }
  return (r >> 3) * p1 + (c >> 3) * p0 + (r & 7) * 16 + (c & 7) * 2;
}

__global__ void __launch_bounds__(128) umma_probe_bf16_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                              float* __restrict__ out, ProbeB p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  constexpr int TB = 65536;
  unsigned char* a_t = smem;
  unsigned char* b_t = smem + TB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  for (int i = threadIdx.x; i < 2 * TB / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = s_tmem;
  if (p.a_src >= 2) {   // thread = row = TMEM lane; K elements 2j, 2j+1 packed into column 256 + j
    const int row = threadIdx.x;
    for (int c0 = 0; c0 < p.CA / 2; c0 += 16) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float x0 = row < p.RA ? A[(size_t)row * p.CA + 2 * (c0 + i)] : 0.f;
        const float x1 = row < p.RA ? A[(size_t)row * p.CA + 2 * (c0 + i) + 1] : 0.f;
        v[i] = __uint_as_float(p.a_src == 2 ? pack_bf16x2(x0, x1) : pack_bf16x2(x1, x0));
      }
      tmem_st16(tmem_addr(tbase, 32 * warp, 256 + c0), v);
    }
    tmem_st_wait();
  } else {
    for (int i = threadIdx.x; i < p.RA * p.CA; i += 128) {
      const int r = i / p.CA, c = i % p.CA;
      *reinterpret_cast<unsigned short*>(a_t + probe_off(p.a_layout, p.a_p0, p.a_p1, r, c)) = (unsigned short)(pack_bf16x2(A[i], 0.f) & 0xffffu);
    }
  }
  for (int i = threadIdx.x; i < p.RB * p.CB; i += 128) {
    const int r = i / p.CB, c = i % p.CB;
    *reinterpret_cast<unsigned short*>(b_t + probe_off(p.b_layout, p.b_p0, p.b_p1, r, c)) = (unsigned short)(pack_bf16x2(B[i], 0.f) & 0xffffu);
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  long long t_start = 0;
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_bf16(p.M, p.N, p.a_mn != 0, p.b_mn != 0);
    auto adesc = [&](int s) {
      return desc_base(smem_u32(a_t) + (s / p.a_div) * p.a_adv_hi + (s % p.a_div) * p.a_adv_lo, p.a_lbo, p.a_sbo) | ((uint64_t)p.a_ltype << 61);
    };
    auto bdesc = [&](int s) {
      return desc_base(smem_u32(b_t) + (s / p.b_div) * p.b_adv_hi + (s % p.b_div) * p.b_adv_lo, p.b_lbo, p.b_sbo) | ((uint64_t)p.b_ltype << 61);
    };
    t_start = clock64();
    for (int rep = 0; rep < p.reps; ++rep)
      for (int s = 0; s < p.ksteps; ++s) {
        if (p.a_src >= 2) mma_bf16_ta(tbase, tbase + 256 + 8 * s, bdesc(s), idesc, s > 0 || rep > 0);
        else mma_bf16(tbase, adesc(s), bdesc(s), idesc, s > 0 || rep > 0);
      }
    mma_commit(&s_bar);
  }
  mbar_wait(&s_bar, 0);
  fence_after_sync();
  if (threadIdx.x == 0) out[128 * 256] = (float)(clock64() - t_start);
  for (int c = 0; c < p.N / 16; ++c) {
    float v[16];
    tmem_ld16(tmem_addr(tbase, 32 * warp, 16 * c), v);
#pragma unroll
    for (int i = 0; i < 16; ++i) out[(size_t)(32 * warp + lane) * 256 + 16 * c + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace scgib

// params = 29 int32 in the order of ProbeB; out has 128*256 + 1 floats (the last = SM cycles of the MMA sequence)
extern "C" SCGIB_API int scgib_debug_umma_bf16(const float* A, const float* B, float* out, const int32_t* params, void* stream) {
  if (!A || !B || !out || !params) return SCGIB_E_NULL;
  scgib::ProbeB p;
  static_assert(sizeof(scgib::ProbeB) == 29 * sizeof(int), "ProbeB is 29 ints");
  memcpy(&p, params, sizeof(p));
  if ((p.M != 64 && p.M != 128) || p.N < 8 || p.N > 256 || (p.N & 15) || p.RA < 1 || p.RB < 1 || p.a_div < 1 || p.b_div < 1 ||
      p.ksteps < 1 || p.ksteps > 16 || p.reps < 1 || p.RA * p.CA > 128 * 128 || p.RB * p.CB > 256 * 128)
    return SCGIB_E_SHAPE;
  const int smem = 2 * 65536 + 1024;
  cudaFuncSetAttribute(scgib::umma_probe_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  scgib::umma_probe_bf16_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, out, p);
  return (int)cudaGetLastError();
}
