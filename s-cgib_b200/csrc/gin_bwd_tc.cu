// gin_bwd_tc.cu - GIN layer backward (part 2) on the 5th-generation tensor cores: a persistent, warp-specialised
// tcgen05 kernel with the same contract as gin_bwd_main_kernel (gin_kernels.cu; reference: autograd of models.py:66-72).
//
//   g_y = rstd * (gamma*g_o - c1 - yhat*c2)                      (BatchNorm backward, applied while loading)
//   G1: g_r = g_y W2        G3: dW2 += g_y^T r     db2 += sum g_y
//   g_u = g_r * [r > 0]
//   G2: g_a = g_u W1        G4: dW1 += g_u^T a     db1 += sum g_u
//
// All four GEMMs run as 3xTF32 tcgen05.mma with fp32 accumulation in tensor memory.  The row tiles g_y, r, g_u, a live
// in shared memory ONCE each (tf32 hi/lo), in the swizzled tile format S of umma.cuh, which the tensor core can read
// both K-major (contraction over the 64 channels: G1, G2) and MN-major (contraction over the 128 tile rows: G3, G4).
// The hi/lo copies of every B operand are adjacent, so one N-stacked MMA yields the hi*hi and hi*lo products at once.
// dW2 / dW1 accumulate in tensor memory over all tiles of the CTA and are written once, as per-CTA partials.
// Shared memory holds two 64 KB tile buffers: X = g_y, then g_u; Y = r, then a (loaded while G1/G3 run).
// Roles: 16 loader warps (coalesced float4 rows -> BN backward -> hi/lo split -> shared memory; next tile prefetched
// into L2), one MMA-issuing warp, 4 epilogue warps (thread = tile row = TMEM lane).
#include <stdlib.h>
#include "kernels.cuh"
#include "umma.cuh"

namespace scgib {
using namespace umma;

// per-tile role timestamps (SCGIB_DBG bit 2048; experiments only, tests/gpu_tc2_trace.py bwd)
__device__ long long g_bwd_trace[160 * 16 * 12];
#define BWD_TRACE(ev, tile) do { if (trace_on && (tile) < 16 && blockIdx.x < 160) g_bwd_trace[((size_t)blockIdx.x * 16 + (tile)) * 12 + (ev)] = clock64(); } while (0)

namespace bwdtc {
constexpr int TM = 128;
constexpr int kEpiWarps = 4, kLoadWarps = 16;
constexpr int kThreadsTotal = (kEpiWarps + 1 + kLoadWarps) * 32;
constexpr int LT = kLoadWarps * 32;                 // loader threads
constexpr int kBuf = 2 * TM * HID * 4;              // one hi + lo tile pair (64 KB)
// TMEM columns: D1 (g_r, 128) | D2 (g_a, 2*KIN) | D3 (dW2, 128, M = 64) | D4 (dW1, 2*KIN, M = 64)
constexpr int kColD1 = 0, kColD2 = 128, kColD3 = 256, kColD4 = 384;
enum { B_FULL1 = 0, B_FULL2, B_GU, B_D1, B_D2, B_COUNT };

template <int KIN>
struct Smem {
  static constexpr int W2B = HID * HID * 4, W1B = KIN * HID * 4;     // one hi (or lo) transposed weight tile
  static constexpr int off_x = 0, off_y = kBuf;
  static constexpr int off_w2 = 2 * kBuf;                            // W2t hi | lo   ([in][out], dense cores)
  static constexpr int off_w1 = off_w2 + 2 * W2B;                    // W1t hi | lo   ([kin][out])
  static constexpr int off_mask = off_w1 + 2 * W1B;                  // uint2 [TM]: r > 0 bits
  static constexpr int off_red = off_mask + TM * 8;                  // float [32][HID] column-sum scratch
  static constexpr int off_bar = off_red + 32 * HID * 4;
  static constexpr int total = off_bar + 128;
};

__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16, off = 16; h >= 1; h >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

template <int KIN>
__global__ void __launch_bounds__(kThreadsTotal, 1)
gin_bwd_tc_kernel(GinBwdMainPair pp) {
  using L = Smem<KIN>;
  const bool second = (int)blockIdx.x >= pp.split;
  const GinBwdMainArgs& p = pp.a[second ? 1 : 0];
  const int bid = second ? (int)blockIdx.x - pp.split : (int)blockIdx.x;          // CTA index / count inside its problem
  const int nblk = second ? (int)gridDim.x - pp.split : pp.split;
  const bool trace_on = pp.trace != 0;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* X = smem + L::off_x;
  unsigned char* Y = smem + L::off_y;
  uint2* s_mask = reinterpret_cast<uint2*>(smem + L::off_mask);
  float* s_red = reinterpret_cast<float*>(smem + L::off_red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::off_bar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::off_bar + B_COUNT * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.V + TM - 1) / TM;
  const int my_tiles = max(0, (n_tiles - bid + nblk - 1) / nblk);
  auto tile_base = [&](int i) { return (bid + i * nblk) * TM; };

  if (threadIdx.x == 0) {
    mbar_init(&bars[B_FULL1], kLoadWarps);
    mbar_init(&bars[B_FULL2], kLoadWarps);
    mbar_init(&bars[B_GU], kEpiWarps * 32);
    mbar_init(&bars[B_D1], 1);
    mbar_init(&bars[B_D2], 1);
  }
  if (warp == kEpiWarps) tmem_alloc(s_tmem, 512);
  // transposed weights, hi/lo split: W2t[in][out] = W2[out][in], W1t[kin][out] = W1[out][kin]  (K-major B operands)
  for (int i = threadIdx.x; i < HID * HID; i += kThreadsTotal) {
    const int o = i / HID, c = i % HID;                       // coalesced read of W2[o][c]
    const float v = __ldg(p.W2 + i), hi = tf32_rna(v), lo = tf32_rna(v - hi);
    const int off = tile_off4(HID, c, o >> 2, 128) + (o & 3) * 4;
    *reinterpret_cast<float*>(smem + L::off_w2 + off) = hi;
    *reinterpret_cast<float*>(smem + L::off_w2 + L::W2B + off) = lo;
  }
  for (int i = threadIdx.x; i < HID * KIN; i += kThreadsTotal) {
    const int o = i / KIN, c = i % KIN;
    const float v = __ldg(p.W1 + i), hi = tf32_rna(v), lo = tf32_rna(v - hi);
    const int off = tile_off4(HID, c, o >> 2, 128) + (o & 3) * 4;
    *reinterpret_cast<float*>(smem + L::off_w1 + off) = hi;
    *reinterpret_cast<float*>(smem + L::off_w1 + L::W1B + off) = lo;
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;

  float4 db2 = make4(0.f);          // loaders: column sums of g_y for channels 4*gl..4*gl+3 over this thread's rows
  float db1[2] = {0.f, 0.f};        // epilogue: column sums of g_u for columns lane, 32 + lane over this warp's rows

  if (warp >= kEpiWarps + 1) {
    // =========================================================================== loaders
    const int pt = (warp - (kEpiWarps + 1)) * 32 + lane;
    constexpr int RPP = LT / 16, NR = TM / RPP;               // 32 rows per pass, 4 passes
    const int gl = pt & 15, gr = pt >> 4;
    const int c = gl * 4;
    const float4 mean = ldg4(p.bn + c), rstd = ldg4(p.bn + HID + c), gamma = ldg4(p.bn + 2 * HID + c);
    const float4 c1 = ldg4(p.cvec + c), c2 = ldg4(p.cvec + HID + c);
    // g_y = ka*g_o - kd*(y - mean) - ke      (ka = rstd*gamma, kd = rstd^2*c2, ke = rstd*c1)
    const float4 ka = make_float4(rstd.x * gamma.x, rstd.y * gamma.y, rstd.z * gamma.z, rstd.w * gamma.w);
    const float4 kd = make_float4(rstd.x * rstd.x * c2.x, rstd.y * rstd.y * c2.y, rstd.z * rstd.z * c2.z, rstd.w * rstd.w * c2.w);
    const float4 ke = make_float4(rstd.x * c1.x, rstd.y * c1.y, rstd.z * c1.z, rstd.w * c1.w);
    constexpr int ALPR = KIN / 4;                             // lanes per `a` row
    for (int i = 0; i < my_tiles; ++i) {
      const int base = tile_base(i);
      if (pt == 0) BWD_TRACE(0, i);
      // ---- phase 1: g_o, y, r rows -> g_y, r (hi/lo) -> X, Y
      float4 go[NR], yy[NR], rr[NR];
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const int v = base + gr + j * RPP;
        const bool ok = v < p.V;
        const size_t o = (size_t)(ok ? v : 0) * HID + c;
        go[j] = ok ? ld4(p.g_o + o) : make4(0.f);
        yy[j] = ok ? ld4_cs(p.y + o) : make4(0.f);
        rr[j] = ok ? ld4_cs(p.r + o) : make4(0.f);
      }
      if (i + 1 < my_tiles) {                                 // next tile -> L2: 256 lines of 128 B per [128][64] tile
        const int nb = tile_base(i + 1), line = pt & 255;
        const size_t off = (size_t)nb * HID + (size_t)line * 32;
        if (off < (size_t)p.V * HID) {
          if (pt < 256) { prefetch_l2(p.g_o + off); prefetch_l2(p.r + off); }
          else {
            prefetch_l2(p.y + off);
            const size_t aoff = (size_t)nb * KIN + (size_t)line * 32;
            if (line * 32 < TM * KIN && aoff < (size_t)p.V * KIN) prefetch_l2(p.a + aoff);
          }
        }
      }
      if (i > 0) mbar_wait(&bars[B_D2], (uint32_t)((i - 1) & 1));   // G2 / G4 of the previous tile have read X and Y
      if (pt == 0) BWD_TRACE(1, i);
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const int row = gr + j * RPP;
        const bool ok = base + row < p.V;
        float4 gy;
        gy.x = ka.x * go[j].x - kd.x * (yy[j].x - mean.x) - ke.x;
        gy.y = ka.y * go[j].y - kd.y * (yy[j].y - mean.y) - ke.y;
        gy.z = ka.z * go[j].z - kd.z * (yy[j].z - mean.z) - ke.z;
        gy.w = ka.w * go[j].w - kd.w * (yy[j].w - mean.w) - ke.w;
        if (!ok) gy = make4(0.f);
        db2 = add4(db2, gy);
        store_split4_s(X, X + kBuf / 2, TM, row, gl, gy);
        store_split4_s(Y, Y + kBuf / 2, TM, row, gl, rr[j]);
        // r > 0 bits of the row: word j holds channel 4*l + j at bit l
        const unsigned b0 = __ballot_sync(0xffffffffu, rr[j].x > 0.f), b1 = __ballot_sync(0xffffffffu, rr[j].y > 0.f);
        const unsigned b2 = __ballot_sync(0xffffffffu, rr[j].z > 0.f), b3 = __ballot_sync(0xffffffffu, rr[j].w > 0.f);
        if (gl == 0) {
          const int sh = lane & 16;
          s_mask[row] = make_uint2(((b0 >> sh) & 0xffffu) | (((b1 >> sh) & 0xffffu) << 16),
                                   ((b2 >> sh) & 0xffffu) | (((b3 >> sh) & 0xffffu) << 16));
        }
      }
      fence_smem_to_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_FULL1]);
      if (pt == 0) BWD_TRACE(2, i);
      // ---- phase 2: a rows -> (after G1 / G3 have read Y) -> Y
      constexpr int ARPP = LT / ALPR, ANR = TM / ARPP;
      const int al = pt % ALPR, ar = pt / ALPR;
      float4 aa[ANR];
#pragma unroll
      for (int j = 0; j < ANR; ++j) {
        const int v = base + ar + j * ARPP;
        aa[j] = v < p.V ? ld4_cs(p.a + (size_t)v * KIN + al * 4) : make4(0.f);
      }
      mbar_wait(&bars[B_D1], (uint32_t)(i & 1));
      if (pt == 0) BWD_TRACE(3, i);
#pragma unroll
      for (int j = 0; j < ANR; ++j) store_split4_s(Y, Y + TM * KIN * 4, TM, ar + j * ARPP, al, aa[j]);
      fence_smem_to_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_FULL2]);
      if (pt == 0) BWD_TRACE(4, i);
    }
  } else if (warp == kEpiWarps) {
    // =========================================================================== MMA issuer
    {   // the whole warp runs the loop (uniform descriptors); one elected lane issues each instruction
      const uint32_t xh = smem_u32(X), xl = xh + kBuf / 2;
      const uint32_t yh = smem_u32(Y);
      const uint32_t yl_a = yh + TM * KIN * 4;                       // lo tile of `a`
      const uint32_t w2 = smem_u32(smem + L::off_w2), w1 = smem_u32(smem + L::off_w1);
      constexpr uint32_t idG1a = idesc_tf32(TM, 2 * HID, false, false), idG1b = idesc_tf32(TM, HID, false, false);
      constexpr uint32_t idG2a = idesc_tf32(TM, 2 * KIN, false, false), idG2b = idesc_tf32(TM, KIN, false, false);
      constexpr uint32_t idG3a = idesc_tf32(64, 2 * HID, true, true), idG3b = idesc_tf32(64, HID, true, true);
      constexpr uint32_t idG4a = idesc_tf32(64, 2 * KIN, true, true), idG4b = idesc_tf32(64, KIN, true, true);
      (void)yl_a;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(&bars[B_FULL1], (uint32_t)(i & 1));
        fence_after_sync();
        if (lane == 0) BWD_TRACE(5, i);
        // G1: g_r = g_y W2          (A = X K-major, B = [W2t_hi | W2t_lo])
#pragma unroll
        for (int k = 0; k < HID / 8; ++k) {
          const uint64_t b = desc_g_dense(w2, HID, k);
          mma_tf32_w(tmem + kColD1, desc_s_kmajor(xh, TM, k), b, idG1a, k > 0);
          mma_tf32_w(tmem + kColD1, desc_s_kmajor(xl, TM, k), b, idG1b, true);
        }
        // G3: dW2 += g_y^T r        (A = X MN-major, B = [r_hi | r_lo] MN-major), accumulated over all tiles
#pragma unroll
        for (int k = 0; k < TM / 8; ++k) {
          const uint64_t b = desc_s_mnmajor(yh, TM, k);
          mma_tf32_w(tmem + kColD3, desc_s_mnmajor(xh, TM, k), b, idG3a, i > 0 || k > 0);
          mma_tf32_w(tmem + kColD3, desc_s_mnmajor(xl, TM, k), b, idG3b, true);
        }
        mma_commit_w(&bars[B_D1]);
        mbar_wait(&bars[B_GU], (uint32_t)(i & 1));
        mbar_wait(&bars[B_FULL2], (uint32_t)(i & 1));
        fence_after_sync();
        if (lane == 0) BWD_TRACE(6, i);
        // G2: g_a = g_u W1          (A = X K-major, B = [W1t_hi | W1t_lo])
#pragma unroll
        for (int k = 0; k < HID / 8; ++k) {
          const uint64_t b = desc_g_dense(w1, HID, k);
          mma_tf32_w(tmem + kColD2, desc_s_kmajor(xh, TM, k), b, idG2a, k > 0);
          mma_tf32_w(tmem + kColD2, desc_s_kmajor(xl, TM, k), b, idG2b, true);
        }
        // G4: dW1 += g_u^T a        (A = X MN-major, B = [a_hi | a_lo] MN-major)
#pragma unroll
        for (int k = 0; k < TM / 8; ++k) {
          const uint64_t b = desc_s_mnmajor(yh, TM, k);
          mma_tf32_w(tmem + kColD4, desc_s_mnmajor(xh, TM, k), b, idG4a, i > 0 || k > 0);
          mma_tf32_w(tmem + kColD4, desc_s_mnmajor(xl, TM, k), b, idG4b, true);
        }
        mma_commit_w(&bars[B_D2]);
      }
    }
  } else {
    // =========================================================================== epilogue (thread = row = TMEM lane)
    const int row = warp * 32 + lane;
    const uint32_t tl = (uint32_t)(warp * 32) << 16;
    for (int i = 0; i < my_tiles; ++i) {
      const int base = tile_base(i);
      const int gv = base + row;
      // ---- epilogue 1: g_u = g_r * [r > 0] -> X (hi/lo), column sums for db1
      mbar_wait(&bars[B_D1], (uint32_t)(i & 1));
      fence_after_sync();
      if (threadIdx.x == 0) BWD_TRACE(7, i);
      const uint2 m = s_mask[row];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float g[32], t2[32];
        tmem_ld16_nowait(tmem + tl + kColD1 + 32 * h, *reinterpret_cast<float (*)[16]>(g));
        tmem_ld16_nowait(tmem + tl + kColD1 + 32 * h + 16, *reinterpret_cast<float (*)[16]>(g + 16));
        tmem_ld16_nowait(tmem + tl + kColD1 + HID + 32 * h, *reinterpret_cast<float (*)[16]>(t2));
        tmem_ld16_nowait(tmem + tl + kColD1 + HID + 32 * h + 16, *reinterpret_cast<float (*)[16]>(t2 + 16));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int cc = 32 * h + j;                          // channel 4*l + q  ->  bit l of 16-bit field q
          const unsigned word = (cc & 2) ? m.y : m.x;
          const bool on = (word >> (((cc & 1) << 4) + (cc >> 2))) & 1u;
          g[j] = on ? g[j] + t2[j] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          store_split4_s(X, X + kBuf / 2, TM, row, 8 * h + q, make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]));
        db1[h] += warp_colsum32(g, lane);                     // rows beyond V hold zeros (zero g_y rows)
      }
      fence_smem_to_async();
      fence_before_sync();
      mbar_arrive(&bars[B_GU]);
      if (threadIdx.x == 0) BWD_TRACE(8, i);
      // ---- epilogue 2: g_a -> global
      mbar_wait(&bars[B_D2], (uint32_t)(i & 1));
      fence_after_sync();
      if (threadIdx.x == 0) BWD_TRACE(9, i);
#pragma unroll
      for (int c0 = 0; c0 < KIN; c0 += 32) {
        float g[32], t2[32];
        tmem_ld16_nowait(tmem + tl + kColD2 + c0, *reinterpret_cast<float (*)[16]>(g));
        tmem_ld16_nowait(tmem + tl + kColD2 + c0 + 16, *reinterpret_cast<float (*)[16]>(g + 16));
        tmem_ld16_nowait(tmem + tl + kColD2 + KIN + c0, *reinterpret_cast<float (*)[16]>(t2));
        tmem_ld16_nowait(tmem + tl + kColD2 + KIN + c0 + 16, *reinterpret_cast<float (*)[16]>(t2 + 16));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] += t2[j];
        if (gv < p.V) {
#pragma unroll
          for (int j = 0; j < 4; ++j) st8(p.g_a + (size_t)gv * KIN + c0 + 8 * j, g + 8 * j);
        }
      }
      fence_before_sync();
      if (threadIdx.x == 0) BWD_TRACE(10, i);
    }
  }
  // ---- every CTA writes its partial gradients (zeros when it had no tile)
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  float* part = p.part + (size_t)blockIdx.x * p.pstride;
  if (warp < kEpiWarps) {
    // dW2 / dW1 from tensor memory: UMMA M = 64 keeps accumulator row o in lane (o/16)*32 + o%16 -> warp o/16, lane o%16
    const uint32_t tl = (uint32_t)(warp * 32) << 16;
    const int o = warp * 16 + (lane & 15);
    const bool act = lane < 16;
#pragma unroll
    for (int c0 = 0; c0 < HID; c0 += 32) {
      float g[32], t2[32];
      if (my_tiles > 0) {
        tmem_ld16_nowait(tmem + tl + kColD3 + c0, *reinterpret_cast<float (*)[16]>(g));
        tmem_ld16_nowait(tmem + tl + kColD3 + c0 + 16, *reinterpret_cast<float (*)[16]>(g + 16));
        tmem_ld16_nowait(tmem + tl + kColD3 + HID + c0, *reinterpret_cast<float (*)[16]>(t2));
        tmem_ld16_nowait(tmem + tl + kColD3 + HID + c0 + 16, *reinterpret_cast<float (*)[16]>(t2 + 16));
        tmem_ld_wait();
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) g[j] = my_tiles > 0 ? g[j] + t2[j] : 0.f;
      if (act) {
#pragma unroll
        for (int j = 0; j < 8; ++j) st4(part + p.off_W2 + (size_t)o * HID + c0 + 4 * j, make_float4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]));
      }
    }
#pragma unroll
    for (int c0 = 0; c0 < KIN; c0 += 32) {
      float g[32], t2[32];
      if (my_tiles > 0) {
        tmem_ld16_nowait(tmem + tl + kColD4 + c0, *reinterpret_cast<float (*)[16]>(g));
        tmem_ld16_nowait(tmem + tl + kColD4 + c0 + 16, *reinterpret_cast<float (*)[16]>(g + 16));
        tmem_ld16_nowait(tmem + tl + kColD4 + KIN + c0, *reinterpret_cast<float (*)[16]>(t2));
        tmem_ld16_nowait(tmem + tl + kColD4 + KIN + c0 + 16, *reinterpret_cast<float (*)[16]>(t2 + 16));
        tmem_ld_wait();
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) g[j] = my_tiles > 0 ? g[j] + t2[j] : 0.f;
      if (act) {
#pragma unroll
        for (int j = 0; j < 8; ++j) st4(part + p.off_W1 + (size_t)o * KIN + c0 + 4 * j, make_float4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]));
      }
    }
    // db1: the four epilogue warps' column sums, fixed order
    s_red[warp * HID + lane] = db1[0];
    s_red[warp * HID + 32 + lane] = db1[1];
  }
  __syncthreads();
  if (threadIdx.x < HID)
    part[p.off_b1 + threadIdx.x] = (s_red[threadIdx.x] + s_red[HID + threadIdx.x]) + (s_red[2 * HID + threadIdx.x] + s_red[3 * HID + threadIdx.x]);
  __syncthreads();
  // db2: loaders' per-thread sums, 32 row groups per channel quad, fixed order
  if (warp >= kEpiWarps + 1) {
    const int pt = (warp - (kEpiWarps + 1)) * 32 + lane;
    st4(s_red + (pt >> 4) * HID + (pt & 15) * 4, db2);
  }
  __syncthreads();
  if (threadIdx.x < HID) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 32; ++g) s += s_red[g * HID + threadIdx.x];
    part[p.off_b2 + threadIdx.x] = s;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem, 512);
}

}  // namespace bwdtc

template <int KIN>
static void launch_bwd_tc(const GinBwdMainPair& pp, int grid, cudaStream_t s) {
  using L = bwdtc::Smem<KIN>;
  static bool once = (cudaFuncSetAttribute(bwdtc::gin_bwd_tc_kernel<KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total), true);
  (void)once;
  bwdtc::gin_bwd_tc_kernel<KIN><<<grid, bwdtc::kThreadsTotal, L::total, s>>>(pp);
}

static int bwd_trace_flag() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SCGIB_DBG"); v = (e && (atoi(e) & 2048)) ? 1 : 0; }
  return v;
}

void launch_gin_bwd_main_tc(const GinBwdMainArgs& a, int kin, int grid, cudaStream_t s) {
  GinBwdMainPair pp;
  pp.trace = 0;
  pp.a[0] = a; pp.a[1] = a;
  pp.split = grid;
  if (kin == DTR) launch_bwd_tc<DTR>(pp, grid, s); else launch_bwd_tc<HID>(pp, grid, s);
}

// the same layer of both encoders in one launch: CTAs [0, split) write the partial gradients of a0, [split, grid) of a1
void launch_gin_bwd_main_tc_pair(const GinBwdMainArgs& a0, const GinBwdMainArgs& a1, int kin, int grid, cudaStream_t s) {
  GinBwdMainPair pp;
  pp.a[0] = a0; pp.a[1] = a1;
  pp.split = pair_split(grid, (a0.V + bwdtc::TM - 1) / bwdtc::TM, (a1.V + bwdtc::TM - 1) / bwdtc::TM);
  pp.trace = bwd_trace_flag();
  if (kin == DTR) launch_bwd_tc<DTR>(pp, grid, s); else launch_bwd_tc<HID>(pp, grid, s);
}

}  // namespace scgib
