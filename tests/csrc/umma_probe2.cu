// umma_probe2.cu - runtime-parametrised probe of tcgen05 operand formats (descriptor fields, layout types, A operand
// in tensor memory).  Test infrastructure for the tensor-core kernels: the formats used by gin_tc*.cu were selected
// with this probe on a B200 (tests/gpu_umma_probe2.py).
#include <string.h>
#include "umma.cuh"
#include "scgib.h"

namespace scgib {
using namespace umma;

struct Probe2 {
  int M, N, ksteps, split;            // split: 1 = hi only (plain TF32), 3 = 3xTF32
  int a_fmt, b_fmt;                   // 0 = format G (144-byte cores), 1 = format S (32B-base swizzle), 2 = A in TMEM
  int a_mn, b_mn;                     // operand major bits of the instruction descriptor
  int a_lbo, a_sbo, a_ltype, a_div, a_adv_lo, a_adv_hi;
  int b_lbo, b_sbo, b_ltype, b_div, b_adv_lo, b_adv_hi;
  int RA, RB;                         // rows of the A / B source matrices ([R][64] fp32)
  int reps;                           // timing: the whole MMA sequence is issued `reps` times; cycles -> out[128*64]
  int a_lo_off;                       // byte offset of the A lo tile from the hi tile (0 = default 40960); adjacent tiles allow M-stacking
};

__global__ void __launch_bounds__(128) umma_probe2_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ out, Probe2 p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  constexpr int TB = 40960;   // >= tile_bytes(128, 64) = 36864 and tile_s_bytes(128) = 32768, 1024-aligned
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + (p.a_lo_off > 0 ? p.a_lo_off : TB);
  unsigned char* b_hi = a_lo + TB;
  unsigned char* b_lo = b_hi + TB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&s_tmem, 256);
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = s_tmem;
  if (p.a_fmt == 2) {   // thread = row = TMEM lane; hi at columns 64.., lo at columns 128..
    const int row = threadIdx.x;
    for (int c = 0; c < 4; ++c) {
      float hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float v = row < p.RA ? A[(size_t)row * 64 + 16 * c + i] : 0.f;
        hi[i] = tf32_rna(v); lo[i] = tf32_rna(v - hi[i]);
      }
      tmem_st16(tmem_addr(tbase, 32 * warp, 64 + 16 * c), hi);
      tmem_st16(tmem_addr(tbase, 32 * warp, 128 + 16 * c), lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  } else {
    for (int i = threadIdx.x; i < p.RA * 16; i += 128) {
      const float4 v = *reinterpret_cast<const float4*>(A + (size_t)(i >> 4) * 64 + (i & 15) * 4);
      if (p.a_fmt == 1) store_split4_s(a_hi, a_lo, p.RA, i >> 4, i & 15, v);
      else store_split4(a_hi, a_lo, 64, i >> 4, i & 15, v);
    }
  }
  for (int i = threadIdx.x; i < p.RB * 16; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(B + (size_t)(i >> 4) * 64 + (i & 15) * 4);
    if (p.b_fmt == 1) store_split4_s(b_hi, b_lo, p.RB, i >> 4, i & 15, v);
    else store_split4(b_hi, b_lo, 64, i >> 4, i & 15, v);
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  long long t_start = 0;
  const bool issuer = (p.reps >= 1000) ? (warp == 0 && elect_one()) : (threadIdx.x == 0);   // reps >= 1000: elect.sync issue
  if (issuer) {
    const uint32_t idesc = idesc_tf32(p.M, p.N, p.a_mn != 0, p.b_mn != 0);
    auto adesc = [&](uint32_t base, int s) {
      return desc_base(base + (s / p.a_div) * p.a_adv_hi + (s % p.a_div) * p.a_adv_lo, p.a_lbo, p.a_sbo) | ((uint64_t)p.a_ltype << 61);
    };
    auto bdesc = [&](uint32_t base, int s) {
      return desc_base(base + (s / p.b_div) * p.b_adv_hi + (s % p.b_div) * p.b_adv_lo, p.b_lbo, p.b_sbo) | ((uint64_t)p.b_ltype << 61);
    };
    uint64_t dah[16], dal[16], dbh[16], dbl[16];   // descriptors precomputed: the timed loop only issues
    for (int s = 0; s < p.ksteps; ++s) {
      dbh[s] = bdesc(smem_u32(b_hi), s); dbl[s] = bdesc(smem_u32(b_lo), s);
      dah[s] = adesc(smem_u32(a_hi), s); dal[s] = adesc(smem_u32(a_lo), s);
    }
    t_start = clock64();
    if (p.reps > 1) {   // timing mode: the step-0 MMAs re-issued reps*ksteps times from registers (results are garbage)
      const uint64_t a0h = dah[0], a0l = dal[0], b0h = dbh[0], b0l = dbl[0];
      const uint32_t th = tbase + 64, tl_ = tbase + 128;
      const int n = p.reps * p.ksteps;
      if (p.a_fmt == 2) {
        if (p.split == 3) for (int it = 0; it < n; ++it) { mma_tf32_ta(tbase, tl_, b0h, idesc, true); mma_tf32_ta(tbase, th, b0l, idesc, true); mma_tf32_ta(tbase, th, b0h, idesc, true); }
        else for (int it = 0; it < n; ++it) mma_tf32_ta(tbase, th, b0h, idesc, true);
      } else {
        if (p.split == 3) for (int it = 0; it < n; ++it) { mma_tf32(tbase, a0l, b0h, idesc, true); mma_tf32(tbase, a0h, b0l, idesc, true); mma_tf32(tbase, a0h, b0h, idesc, true); }
        else for (int it = 0; it < n; ++it) mma_tf32(tbase, a0h, b0h, idesc, true);
      }
    } else
    for (int s = 0; s < p.ksteps; ++s) {
      const bool first = (s == 0);
      if (p.a_fmt == 2) {
        const uint32_t ah = tbase + 64 + 8 * s, al = tbase + 128 + 8 * s;
        if (p.split == 3) { mma_tf32_ta(tbase, al, dbh[s], idesc, !first); mma_tf32_ta(tbase, ah, dbl[s], idesc, true); }
        mma_tf32_ta(tbase, ah, dbh[s], idesc, p.split == 3 || !first);
      } else {
        if (p.split == 3) { mma_tf32(tbase, dal[s], dbh[s], idesc, !first); mma_tf32(tbase, dah[s], dbl[s], idesc, true); }
        mma_tf32(tbase, dah[s], dbh[s], idesc, p.split == 3 || !first);
      }
    }
    mma_commit(&s_bar);
  }
  mbar_wait(&s_bar, 0);
  fence_after_sync();
  if (issuer) out[128 * 64] = (float)(clock64() - t_start);
  for (int c = 0; c < 4; ++c) {
    float v[16];
    tmem_ld16(tmem_addr(tbase, 32 * warp, 16 * c), v);
#pragma unroll
    for (int i = 0; i < 16; ++i) out[(size_t)(32 * warp + lane) * 64 + 16 * c + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

}  // namespace scgib

extern "C" SCGIB_API int scgib_debug_umma2(const float* A, const float* B, float* out, const int32_t* params, void* stream) {
  if (!A || !B || !out || !params) return SCGIB_E_NULL;
  scgib::Probe2 p;
  static_assert(sizeof(scgib::Probe2) == 24 * sizeof(int), "Probe2 is 24 ints");
  memcpy(&p, params, sizeof(p));
  if ((p.M != 64 && p.M != 128) || p.N < 8 || p.N > 256 || (p.N > 64 && p.reps < 2) || p.RA < 1 || p.RA > 128 || p.RB < 1 || p.RB > 128 || p.a_div < 1 ||
      p.b_div < 1 || p.ksteps < 1 || p.ksteps > 16 || p.reps < 1)
    return SCGIB_E_SHAPE;
  const int smem = 4 * 40960 + 1024;
  cudaFuncSetAttribute(scgib::umma_probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  scgib::umma_probe2_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, out, p);
  return (int)cudaGetLastError();
}
