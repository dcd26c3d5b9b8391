// gin_tc4.cu - GIN layer forward, fourth-generation tcgen05 kernel: the neighbour AGGREGATION runs on the tensor cores too.
//
// Same contract as gin_fwd_tc3 (gin_tc3.cu; reference models.py:66-72: DGL GINConv 'sum' + MLP + BatchNorm1d statistics),
// KIN = 64 layers (layers >= 1 of both encoders).  Why: ncu and the role timeline of gin_fwd_tc3 show a 128-row tile bound
// by the PRODUCER warps' instruction stream (~26 k warp instructions per tile for the CSR gather out of the shared-memory
// window: index arithmetic replicated in the 16 lanes of a row, BN + ReLU re-applied per gathered neighbour, divergent
// degree loops), with the tensor pipe 11 % busy and HBM at 16 % of its peak.  Molecular batches are LOCAL (every neighbour
// of a row lies within a few rows of it), so the aggregation of a 128-row tile is a small dense product
//     a[128][64] = Adj[128][WIN] . h[WIN][64],   WIN = 128 + 2 * 32 window rows,  Adj = I + A restricted to the window,
// whose left operand is a 0/1 matrix with ~3 non-zeros per row.  The kernel therefore
//   * converts the window ONCE per row: h = relu(BN(y_prev)) (one pass, coalesced 32-byte loads, no indices), split into
//     two fp16 parts hi = fp16(h), lo' = fp16((h - hi) * 2^11) (11 + 11 significand bits = the same 2^-22 as the 3xTF32
//     products of the MLP GEMMs; the 2^11 scale keeps lo' in fp16's normal range), stored as the MN-major B operand
//     (format B of umma.cuh, hi block | lo' block = ONE N = 128 MMA per K step);
//   * builds Adj as an fp16 K-major A operand (48 KB: cleared with 16-byte stores, then one 2-byte store per edge - the only
//     place the CSR indices are touched);
//   * AGG: 12 tcgen05.mma kind::f16 (M 128, N 128, K 16) -> TMEM columns [hi part | lo' part], exact products, fp32 accumulate;
//   * epilogue A: a = hi + 2^-11 lo' (+ the rare neighbours beyond the window, added from global memory), saved for the
//     backward pass, re-split into tf32 hi/lo and written BACK to tensor memory as the A operand of GEMM1 (as r is for GEMM2);
//   * GEMM1 / epilogue 1 / GEMM2 / epilogue 2 as in gin_tc3 (3xTF32, N-stacked B operands, r through tensor memory).
// Roles: 16 epilogue warps = two groups of 8, group g owns TMEM stage g (tiles g, g+2, ...) and runs that tile's three
// epilogue phases in order; 1 MMA warp issuing the two stages' phases interleaved; 8 producer warps.  Per tile the SIMT
// work drops from ~32 k to ~9 k warp instructions.  Values are clamped to the fp16 range (|h| <= 65504) before the split.
#include <stdlib.h>
#include <cuda_fp16.h>
#include "kernels.cuh"
#include "umma.cuh"

namespace scgib {
using namespace umma;

// per-tile role timestamps (SCGIB_DBG bit 1024; experiments only, tests/gpu_tc4_trace.py)
__device__ long long g_tc4_trace[160 * 16 * 16];
#define TC4_TRACE(ev, tile) do { if (trace_on && (tile) < 16 && blockIdx.x < 160) g_tc4_trace[((size_t)blockIdx.x * 16 + (tile)) * 16 + (ev)] = clock64(); } while (0)

namespace tc4 {
constexpr int TM = 128;                        // rows per tile = UMMA M
constexpr int HALO = 32, WIN = TM + 2 * HALO;  // window rows = K of the aggregation product
constexpr int kEpiWarps = 16, kGroupWarps = 8;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kAdjWarps = 2, AT = kAdjWarps * 32;   // adjacency-building warps / threads
constexpr int kConvWarps = 6, CT = kConvWarps * 32; // window-conversion warps / threads
constexpr int kThreads4 = (kEpiWarps + 1 + kAdjWarps + kConvWarps) * 32;
constexpr int kIdxCap = 1024;                  // staged neighbour indices per tile (more edges: read from global)
constexpr int kHBlk = WIN * 128;               // one [WIN rows][64 fp16] block (format B)
constexpr int kHStage = 2 * kHBlk;             // hi | lo'
constexpr int kABlk = TM * 128;                // one [128 rows][64 K columns] block of Adj
constexpr int kABytes = (WIN / 64) * kABlk;
constexpr int W1B = HID * HID * 4;             // one hi (or lo) weight tile
static_assert(WIN % 64 == 0 && WIN % 16 == 0, "window = whole 64-column Adj blocks");

// instruction descriptor: kind::f16 with fp16 A and B, fp32 accumulate
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr uint32_t kIdAgg = idesc_f16(TM, 2 * HID, false, true);
constexpr uint32_t kIdesc = idesc_tf32(TM, HID, false, false);
constexpr uint32_t kIdesc2 = idesc_tf32(TM, 2 * HID, false, false);
// TMEM columns of stage s (base 256 s): [0,128) AGG out (hi part | lo' part) -> a_hi | a_lo -> later D2;  [128,256) D1 -> r_hi | r_lo
constexpr int kColA = 0, kColD1 = 128, kColD2 = 0;

struct Smem {
  static constexpr int off_h = 0;                                    // 2 stages x (hi | lo')
  static constexpr int off_adj = 2 * kHStage;
  static constexpr int off_w1_hi = off_adj + kABytes, off_w1_lo = off_w1_hi + W1B;
  static constexpr int off_w2_hi = off_w1_lo + W1B, off_w2_lo = off_w2_hi + W1B;
  static constexpr int off_f = off_w2_lo + W1B;                      // b1 b2 mean scale beta [HID]
  static constexpr int off_bar = off_f + 5 * HID * 4;                // 16 mbarriers + tmem slot
  static constexpr int off_far = off_bar + 256;                      // int [4][TM]: rows with neighbours beyond the window
  static constexpr int off_ip = off_far + 4 * TM * 4;                // int [3][TM + 4]   (tile being scattered, tile being copied, one spare:
  static constexpr int off_ix = off_ip + 3 * (TM + 4) * 4;           // int [3][kIdxCap]   producer warps run up to a barrier apart)
  static constexpr int total = off_ix + 3 * kIdxCap * 4;
  static_assert(total <= 227 * 1024, "shared memory budget");
  static_assert(off_adj % 1024 == 0 && off_w1_hi % 1024 == 0 && kHBlk % 1024 == 0, "swizzled blocks are 1024-byte aligned");
};

enum { B_FULL = 0, B_AGG = 2, B_A = 4, B_D1 = 6, B_R = 8, B_D2 = 10, B_E2 = 12, B_FULLA = 14, B_COUNT = 16 };

__device__ __forceinline__ void mma_f16_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  if (elect_one()) mma_bf16(d_tmem, a_desc, b_desc, idesc, accumulate);      // the same kind::f16 instruction; formats are in idesc
}
// bounded mbarrier wait (SCGIB_DBG bit 4096 builds only) is not needed in production: plain try_wait loop of umma.cuh

// (a, b) -> packed fp16 pairs hi = fp16(v), lo = fp16((v - hi) * 2048); element a in the low half
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  float ha, hb;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(ha), "=f"(hb) : "r"(hi));
  const float ea = (a - ha) * 2048.f, eb = (b - hb) * 2048.f;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(eb), "f"(ea));
}

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
__device__ __forceinline__ void ld8nc(const float* p, float* v) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]),
               "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}

__global__ void __launch_bounds__(kThreads4, 1)
gin_fwd_tc4_kernel(GinFwdPair pp) {
  using L = Smem;
  constexpr int KIN = HID;
  const bool second = (int)blockIdx.x >= pp.split;
  const GinFwdArgs& p = pp.a[second ? 1 : 0];
  const int bid = second ? (int)blockIdx.x - pp.split : (int)blockIdx.x;          // CTA index / count inside its problem
  const int nblk = second ? (int)gridDim.x - pp.split : pp.split;
  extern __shared__ __align__(1024) unsigned char smem[];
  float* s_b1 = reinterpret_cast<float*>(smem + L::off_f);
  float* s_b2 = s_b1 + HID;
  float* s_mean = s_b2 + HID;                                 // BN of the producing layer: h = max((y - mean) * scale + beta, 0)
  float* s_scale = s_mean + HID;
  float* s_beta = s_scale + HID;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::off_bar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::off_bar + B_COUNT * 8);
  int* s_far = reinterpret_cast<int*>(smem + L::off_far);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.V + TM - 1) / TM;
  const int my_tiles = (n_tiles - bid + nblk - 1) / nblk;     // tiles bid + i*nblk
  const bool rev = p.reverse != 0;
  const bool has_bn = p.bn_in != nullptr;
  const bool trace_on = (p.dbg & 1024) != 0;
  auto tile_base = [&](int i) { return (bid + (rev ? my_tiles - 1 - i : i) * nblk) * TM; };

  // ---- one-time setup: barriers, TMEM, weights (natural [out][in] = K-major B operand, dense cores), biases
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars[B_FULL + s], kConvWarps);
      mbar_init(&bars[B_FULLA + s], kAdjWarps);
      mbar_init(&bars[B_AGG + s], 1);
      mbar_init(&bars[B_A + s], kGroupWarps * 32);
      mbar_init(&bars[B_D1 + s], 1);
      mbar_init(&bars[B_R + s], kGroupWarps * 32);
      mbar_init(&bars[B_D2 + s], 1);
      mbar_init(&bars[B_E2 + s], kGroupWarps * 32);
    }
  }
  if (warp == kMmaWarp) tmem_alloc(s_tmem, 512);
  for (int i = threadIdx.x; i < HID * (KIN / 4); i += kThreads4) {
    const int o = i / (KIN / 4), c4 = i % (KIN / 4);
    store_split4(smem + L::off_w1_hi, smem + L::off_w1_lo, KIN, o, c4, ldg4(p.W1 + (size_t)o * KIN + c4 * 4), 128);
    store_split4(smem + L::off_w2_hi, smem + L::off_w2_lo, HID, o, c4, ldg4(p.W2 + (size_t)o * HID + c4 * 4), 128);
  }
  if (threadIdx.x < HID) { s_b1[threadIdx.x] = p.b1[threadIdx.x]; s_b2[threadIdx.x] = p.b2[threadIdx.x]; }
  pdl_sync();      // everything above reads parameters only; from here on: the previous kernel's outputs (bn_in, activations)
  if (threadIdx.x < HID) {
    const int c = threadIdx.x;
    if (has_bn) { s_mean[c] = p.bn_in[c]; s_scale[c] = p.bn_in[HID + c] * p.bn_in[2 * HID + c]; s_beta[c] = p.bn_in[3 * HID + c]; }
    else { s_mean[c] = 0.f; s_scale[c] = 1.f; s_beta[c] = 0.f; }
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;

  double run_n = 0.0, run_mean = 0.0, run_m2 = 0.0;           // epilogue warps: column c0 + lane over this warp's rows

  if (warp > kMmaWarp + kAdjWarps) {
    // =========================================================================== converters: window rows -> h tiles
    // (no indices here: coalesced 32-byte loads of 8 channels, relu(BN(.)), fp16 hi / lo' split, 16-byte swizzled stores)
    const int pt = (warp - (kMmaWarp + 1 + kAdjWarps)) * 32 + lane;
    auto prefetch_window = [&](int i) {                        // 192 rows x 256 B = 384 lines of 128 B
      const long long ws = (long long)tile_base(i) - HALO;
      for (int l = pt; l < WIN * 2; l += CT) {
        const long long g = ws + (l >> 1);
        if (g >= 0 && g < p.V) prefetch_l2(p.in + (size_t)g * KIN + (l & 1) * 32);
      }
    };
    const int c8 = pt & 7, w0 = pt >> 3;                       // this thread's 16-byte chunk (8 channels) and first window row
    constexpr int RPP = CT / 8, UPT = WIN / RPP, NB = 2, UPB = UPT / NB;   // rows per pass, units per thread, batches
    static_assert(UPT * RPP == WIN && UPB * NB == UPT, "window rows divide over the converter threads");
    float mean8[8], sc8[8], be8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { mean8[j] = s_mean[c8 * 8 + j]; sc8[j] = s_scale[c8 * 8 + j]; be8[j] = s_beta[c8 * 8 + j]; }
    if (my_tiles > 0) prefetch_window(0);
    for (int i = 0; i < my_tiles; ++i) {
      const int s = i & 1, use = i >> 1;
      const int ws = tile_base(i) - HALO;
      if (pt == 0) TC4_TRACE(0, i);
      if (i + 1 < my_tiles) prefetch_window(i + 1);            // one tile ahead, towards L2
      // the stage's h tiles are free once AGG of its previous tile has read them
      if (use > 0) mbar_wait(&bars[B_AGG + s], (uint32_t)((use - 1) & 1));
      if (pt == 0) TC4_TRACE(1, i);
      unsigned char* hhi = smem + L::off_h + s * kHStage;
      unsigned char* hlo = hhi + kHBlk;
#pragma unroll
      for (int bt = 0; bt < NB; ++bt) {
        float v[UPB][8];
#pragma unroll
        for (int j = 0; j < UPB; ++j) {
          const int w = w0 + RPP * (UPB * bt + j), g = ws + w;
          if (g >= 0 && g < p.V) ld8nc(p.in + (size_t)g * KIN + c8 * 8, v[j]);
          else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[j][q] = 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < UPB; ++j) {
          const int w = w0 + RPP * (UPB * bt + j), g = ws + w;
          const bool ok = g >= 0 && g < p.V;
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float h = fmaf(v[j][q] - mean8[q], sc8[q], be8[q]);
            h = has_bn ? fmaxf(h, 0.f) : fmaxf(h, -65504.f);
            v[j][q] = ok ? fminf(h, 65504.f) : 0.f;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) split_f16x2(v[j][2 * q], v[j][2 * q + 1], hi[q], lo[q]);
          const int off = tile_b_off(w, c8);
          *reinterpret_cast<uint4*>(hhi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(hlo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_smem_to_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_FULL + s]);
      if (pt == 0) TC4_TRACE(2, i);
    }
  } else if (warp > kMmaWarp) {
    // =========================================================================== adjacency builders (kAdjWarps warps)
    // Adj (single buffer): free once AGG of the previous tile has read it; cleared with 16-byte stores, then the diagonal and
    // one 2-byte entry per edge - the only place the CSR indices are touched.  Runs concurrently with the converters.
    const int pt = (warp - (kMmaWarp + 1)) * 32 + lane;
    int* s_ip = reinterpret_cast<int*>(smem + L::off_ip);
    int* s_ix = reinterpret_cast<int*>(smem + L::off_ix);
    auto adj_sync = [&]() { asm volatile("bar.sync 1, %0;" :: "n"(AT) : "memory"); };
    // indptr slice and neighbour rows of tile i -> index buffer i % 3 (cp.async)
    auto stage_indices = [&](int i, int e_begin, int e_end) {
      const int buf = i % 3, base = tile_base(i);
      for (int r = pt; r <= TM; r += AT) cp_async4(&s_ip[buf * (TM + 4) + r], p.indptr + min(base + r, p.V));
      const int n = min(e_end - e_begin, kIdxCap);
      for (int e = pt; e < n; e += AT) cp_async4(&s_ix[buf * kIdxCap + e], p.indices + e_begin + e);
    };
    auto bounds = [&](int i, int& e_begin, int& e_end) {
      const int base = tile_base(i);
      e_begin = __ldg(p.indptr + base); e_end = __ldg(p.indptr + min(base + TM, p.V));
    };
    int nb_begin = 0, nb_end = 0;
    if (my_tiles > 0) { bounds(0, nb_begin, nb_end); stage_indices(0, nb_begin, nb_end); }
    cp_async_commit();
    if (my_tiles > 1) bounds(1, nb_begin, nb_end);
    unsigned char* adj = smem + L::off_adj;
    for (int i = 0; i < my_tiles; ++i) {
      const int s = i & 1, buf = i % 3;
      const int base = tile_base(i), ws = base - HALO;
      // one tile ahead: indices of tile i+1 (its edge range was read an iteration ago), edge range of tile i+2
      if (i + 1 < my_tiles) stage_indices(i + 1, nb_begin, nb_end);
      cp_async_commit();
      if (i + 2 < my_tiles) bounds(i + 2, nb_begin, nb_end);
      if (i > 0) mbar_wait(&bars[B_AGG + ((i - 1) & 1)], (uint32_t)(((i - 1) >> 1) & 1));
      if (pt == 0) TC4_TRACE(3, i);
#pragma unroll 8
      for (int j = 0; j < kABytes / 16 / AT; ++j) *reinterpret_cast<uint4*>(adj + (size_t)(pt + j * AT) * 16) = make_uint4(0u, 0u, 0u, 0u);
      asm volatile("cp.async.wait_group 1;" ::: "memory");      // this thread's copies of tile i's indices have landed
      adj_sync();                                               // Adj cleared, every thread's index copies visible
      const int* ip = s_ip + buf * (TM + 4);
      const int* ix = s_ix + buf * kIdxCap;
      const int e_begin = ip[0];
#pragma unroll
      for (int rr = 0; rr < TM / AT; ++rr) {
        const int r = pt + rr * AT;
        int far = 0;
        if (base + r < p.V) {
          auto bump = [&](int w) {                              // Adj[r][w] += 1 (this thread owns row r: plain read-modify-write)
            unsigned short* a = reinterpret_cast<unsigned short*>(adj + (w >> 6) * kABlk + tile_b_off(r, (w & 63) >> 3) + (w & 7) * 2);
            const unsigned short cur = *a;
            *a = cur == 0 ? (unsigned short)0x3C00 : __half_as_ushort(__hadd(__ushort_as_half(cur), __ushort_as_half((unsigned short)0x3C00)));
          };
          bump(r + HALO);
          const int e0 = ip[r] - e_begin, e1 = ip[r + 1] - e_begin;
          for (int e = e0; e < e1; ++e) {
            const int u = (e < kIdxCap) ? ix[e] : __ldg(p.indices + e_begin + e);
            const int w = u - ws;
            if ((unsigned)w < (unsigned)WIN) bump(w); else ++far;
          }
        }
        s_far[(i & 3) * TM + r] = far;
      }
      fence_smem_to_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_FULLA + s]);
      if (pt == 0) TC4_TRACE(4, i);
    }
    cp_async_wait_all();
  } else if (warp == kMmaWarp) {
    // =========================================================================== MMA issuer (converged warp, elected lane)
    const uint32_t w1h = smem_u32(smem + L::off_w1_hi), w2h = smem_u32(smem + L::off_w2_hi);
    const uint32_t adj = smem_u32(smem + L::off_adj);
    auto agg = [&](int i) {
      const int s = i & 1, use = i >> 1;
      mbar_wait(&bars[B_FULL + s], (uint32_t)(use & 1));
      mbar_wait(&bars[B_FULLA + s], (uint32_t)(use & 1));
      if (use > 0) mbar_wait(&bars[B_E2 + s], (uint32_t)((use - 1) & 1));     // epilogue 2 of the stage's previous tile has read D2
      fence_after_sync();
      if (lane == 0) TC4_TRACE(5, i);
      const uint32_t hb = smem_u32(smem + L::off_h + s * kHStage);
      const uint32_t d = tmem + s * 256 + kColA;
#pragma unroll
      for (int k = 0; k < WIN / 16; ++k)
        mma_f16_w(d, desc_b_kmajor(adj + (uint32_t)(k >> 2) * kABlk, k & 3), desc_b_mnmajor(hb, kHBlk, k), kIdAgg, k > 0);
      mma_commit_w(&bars[B_AGG + s]);
    };
    auto gemm1 = [&](int i) {
      const int s = i & 1, use = i >> 1;
      mbar_wait(&bars[B_A + s], (uint32_t)(use & 1));
      fence_after_sync();
      if (lane == 0) TC4_TRACE(6, i);
      const uint32_t d = tmem + s * 256 + kColD1;
      const uint32_t ah = tmem + s * 256 + kColA, al = ah + HID;
#pragma unroll
      for (int k = 0; k < KIN / 8; ++k) {
        const uint64_t dbh = desc_g_dense(w1h, KIN, k);     // hi tile; the N = 128 view continues into the lo tile
        mma_tf32_ta_w(d, ah + 8 * k, dbh, kIdesc2, k > 0);
        mma_tf32_ta_w(d, al + 8 * k, dbh, kIdesc, true);
      }
      mma_commit_w(&bars[B_D1 + s]);
    };
    auto gemm2 = [&](int i) {
      const int s = i & 1, use = i >> 1;
      mbar_wait(&bars[B_R + s], (uint32_t)(use & 1));
      fence_after_sync();
      if (lane == 0) TC4_TRACE(7, i);
      const uint32_t d = tmem + s * 256 + kColD2;
      const uint32_t rh = tmem + s * 256 + kColD1, rl = rh + HID;
#pragma unroll
      for (int k = 0; k < HID / 8; ++k) {
        const uint64_t dbh = desc_g_dense(w2h, HID, k);
        mma_tf32_ta_w(d, rh + 8 * k, dbh, kIdesc2, k > 0);
        mma_tf32_ta_w(d, rl + 8 * k, dbh, kIdesc, true);
      }
      mma_commit_w(&bars[B_D2 + s]);
    };
    // tiles in pairs (one per TMEM stage), phases interleaved: each MMA phase of one stage overlaps an epilogue phase of the other
    for (int i0 = 0; i0 < my_tiles; i0 += 2) {
      const bool two = i0 + 1 < my_tiles;
      agg(i0); if (two) agg(i0 + 1);
      gemm1(i0); if (two) gemm1(i0 + 1);
      gemm2(i0); if (two) gemm2(i0 + 1);
    }
  } else {
    // =========================================================================== epilogue
    // group g = warp / 8 owns TMEM stage g; warp: TMEM lane quarter q = warp & 3 (rows 32 q ..), column half (warp >> 2) & 1
    const int grp = warp >> 3, q = warp & 3, half = (warp >> 2) & 1;
    const int row = q * 32 + lane;
    const int c0 = half * 32;
    const uint32_t tl = (uint32_t)(q * 32) << 16;
    const int s = grp;
    const uint32_t t0 = tmem + s * 256 + tl;
    for (int i = grp; i < my_tiles; i += 2) {
      const int use = i >> 1;
      const int base = tile_base(i);
      const int gv = base + row;
      const bool valid = gv < p.V;
      // ---- epilogue A: a = hi part + 2^-11 lo' part (+ far neighbours) -> global (saved) and tensor memory (tf32 hi | lo)
      mbar_wait(&bars[B_AGG + s], (uint32_t)(use & 1));
      // (s_far: written by the adjacency warps before their release-arrive on B_FULLA, which the MMA warp acquired before it
      //  issued AGG; this thread acquired AGG's commit above.  No wait on B_FULLA here: the adjacency warps may already be two
      //  tiles ahead, and a parity wait on a barrier that has advanced two phases never returns.)
      fence_after_sync();
      if ((threadIdx.x & 255) == 0) TC4_TRACE(8, i);
      const int far = s_far[(i & 3) * TM + row];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[16], v2[16], hi[16], lo[16];
        tmem_ld16_nowait(t0 + kColA + c0 + 16 * c, v);
        tmem_ld16_nowait(t0 + kColA + HID + c0 + 16 * c, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(v2[j], 1.f / 2048.f, v[j]);
        if (far > 0 && valid) {                                 // neighbours beyond the window (graphs larger than the halo): rare
          const int ws = base - HALO;
          const int e0 = __ldg(p.indptr + gv), e1 = __ldg(p.indptr + gv + 1);
          for (int e = e0; e < e1; ++e) {
            const int u = __ldg(p.indices + e);
            if ((unsigned)(u - ws) < (unsigned)WIN) continue;
            const float* src = p.in + (size_t)u * KIN + c0 + 16 * c;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int ch = c0 + 16 * c + j;
              float h = fmaf(__ldg(src + j) - s_mean[ch], s_scale[ch], s_beta[ch]);
              h = has_bn ? fmaxf(h, 0.f) : h;
              v[j] += h;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) { hi[j] = tf32_rna(v[j]); lo[j] = tf32_rna(v[j] - hi[j]); }
        tmem_st16(t0 + kColA + c0 + 16 * c, hi);            // in place: these columns have just been read by this thread
        tmem_st16(t0 + kColA + HID + c0 + 16 * c, lo);
        if (p.a_out && valid) {
          st8_cs(p.a_out + (size_t)gv * KIN + c0 + 16 * c, v);
          st8_cs(p.a_out + (size_t)gv * KIN + c0 + 16 * c + 8, v + 8);
        }
      }
      tmem_st_wait();
      fence_before_sync();
      mbar_arrive(&bars[B_A + s]);
      if ((threadIdx.x & 255) == 0) TC4_TRACE(9, i);
      // ---- epilogue 1: r = relu(u + b1) -> global (saved) and tensor memory
      mbar_wait(&bars[B_D1 + s], (uint32_t)(use & 1));
      fence_after_sync();
      if ((threadIdx.x & 255) == 0) TC4_TRACE(10, i);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[16], v2[16], hi[16], lo[16];
        tmem_ld16_nowait(t0 + kColD1 + c0 + 16 * c, v);
        tmem_ld16_nowait(t0 + kColD1 + HID + c0 + 16 * c, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] = fmaxf((v[j] + v2[j]) + s_b1[c0 + 16 * c + j], 0.f);
          hi[j] = tf32_rna(v[j]);
          lo[j] = tf32_rna(v[j] - hi[j]);
        }
        tmem_st16(t0 + kColD1 + c0 + 16 * c, hi);
        tmem_st16(t0 + kColD1 + HID + c0 + 16 * c, lo);
        if (p.r_out && valid) {
          st8_cs(p.r_out + (size_t)gv * HID + c0 + 16 * c, v);
          st8_cs(p.r_out + (size_t)gv * HID + c0 + 16 * c + 8, v + 8);
        }
      }
      tmem_st_wait();
      fence_before_sync();
      mbar_arrive(&bars[B_R + s]);
      if ((threadIdx.x & 255) == 0) TC4_TRACE(11, i);
      // ---- epilogue 2: y -> global, statistics of this warp's rows
      const int cnt = max(0, min(32, p.V - (base + q * 32)));   // valid rows of this warp (warp-uniform)
      mbar_wait(&bars[B_D2 + s], (uint32_t)(use & 1));
      fence_after_sync();
      if ((threadIdx.x & 255) == 0) TC4_TRACE(12, i);
      {
        float y[32], t[32];
        tmem_ld16_nowait(t0 + kColD2 + c0, *reinterpret_cast<float (*)[16]>(y));
        tmem_ld16_nowait(t0 + kColD2 + c0 + 16, *reinterpret_cast<float (*)[16]>(y + 16));
        tmem_ld16_nowait(t0 + kColD2 + HID + c0, *reinterpret_cast<float (*)[16]>(t));
        tmem_ld16_nowait(t0 + kColD2 + HID + c0 + 16, *reinterpret_cast<float (*)[16]>(t + 16));
        tmem_ld_wait();
        fence_before_sync();
        mbar_arrive(&bars[B_E2 + s]);                         // D2 (and the a columns under it) may be overwritten by the stage's next AGG
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = (y[j] + t[j]) + s_b2[c0 + j];
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) st8(p.y_out + (size_t)gv * HID + c0 + 8 * j, y + 8 * j);
        }
        if (cnt > 0) {     // column mean, then centred M2 (two transpose-reduces; lane l <-> column c0 + l); Chan update in fp64
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = valid ? y[j] : 0.f;
          const float mu = warp_colsum32(t, lane) / (float)cnt;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = y[j] - __shfl_sync(0xffffffffu, mu, j);
            t[j] = valid ? d * d : 0.f;
          }
          const float m2 = warp_colsum32(t, lane);
          const double nb = (double)cnt, nt = run_n + nb, dl = (double)mu - run_mean;
          run_m2 += (double)m2 + dl * dl * run_n * nb / nt;
          run_mean += dl * nb / nt;
          run_n = nt;
        }
      }
      if ((threadIdx.x & 255) == 0) TC4_TRACE(13, i);
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (warp == kMmaWarp) tmem_dealloc(tmem, 512);
  // ---- CTA partial (n, mean, M2) per column: Chan combine of the 8 warps (2 groups x 4 row quarters) of the column's half
  double* s_stat = reinterpret_cast<double*>(smem);           // [16 warps][3][32] (the stages are dead by now)
  if (warp < kEpiWarps) {
    s_stat[(warp * 3 + 0) * 32 + lane] = run_n;
    s_stat[(warp * 3 + 1) * 32 + lane] = run_mean;
    s_stat[(warp * 3 + 2) * 32 + lane] = run_m2;
  }
  __syncthreads();
  if (threadIdx.x < HID) {
    const int c = threadIdx.x, half = c >> 5, l = c & 31;
    double n = 0.0, mean = 0.0, m2 = 0.0;
    for (int g = 0; g < 2; ++g)
      for (int q = 0; q < 4; ++q) {
        const int w = g * 8 + half * 4 + q;
        const double nb = s_stat[(w * 3 + 0) * 32 + l];
        if (nb > 0.0) {
          const double mb = s_stat[(w * 3 + 1) * 32 + l], qb = s_stat[(w * 3 + 2) * 32 + l];
          const double nt = n + nb, dl = mb - mean;
          m2 += qb + dl * dl * n * nb / nt;
          mean += dl * nb / nt;
          n = nt;
        }
      }
    double* part = reinterpret_cast<double*>(p.part) + (size_t)bid * 3 * HID;
    part[c] = n; part[HID + c] = mean; part[2 * HID + c] = m2;
  }
  if (!last_cta_arrives(p.counter, (unsigned)nblk)) return;
  // ---- batch statistics: the last CTA of the problem combines the per-CTA partials in fp64 (as gin_fwd_tc3)
  {
    constexpr int SEGS = 4, BATCH = 8;
    const int c = threadIdx.x & (HID - 1), seg = threadIdx.x >> 6;
    const double* part = reinterpret_cast<const double*>(p.part);
    double* s_comb = reinterpret_cast<double*>(smem) + 16 * 3 * 32;        // [SEGS][3][HID]
    if (seg < SEGS) {
      double n = 0.0, mean = 0.0, m2 = 0.0;
      for (int b0 = seg; b0 < nblk; b0 += SEGS * BATCH) {
        double pn[BATCH], pm[BATCH], pq[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
          const int b = b0 + k * SEGS;
          const bool ok = b < nblk;
          pn[k] = ok ? __ldcg(part + (size_t)b * 3 * HID + c) : 0.0;
          pm[k] = ok ? __ldcg(part + (size_t)b * 3 * HID + HID + c) : 0.0;
          pq[k] = ok ? __ldcg(part + (size_t)b * 3 * HID + 2 * HID + c) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
          const double nt = n + pn[k], dl = pm[k] - mean, w = pn[k] / fmax(nt, 1.0);
          m2 += pq[k] + dl * dl * n * w;
          mean += dl * w;
          n = nt;
        }
      }
      s_comb[(seg * 3 + 0) * HID + c] = n; s_comb[(seg * 3 + 1) * HID + c] = mean; s_comb[(seg * 3 + 2) * HID + c] = m2;
    }
    __syncthreads();
    if (threadIdx.x < HID) {
      double n = 0.0, mean = 0.0, m2 = 0.0;
#pragma unroll
      for (int w4 = 0; w4 < SEGS; ++w4) {
        const double nb = s_comb[(w4 * 3 + 0) * HID + c], mb = s_comb[(w4 * 3 + 1) * HID + c], qb = s_comb[(w4 * 3 + 2) * HID + c];
        const double nt = n + nb, dl = mb - mean, w = nb / fmax(nt, 1.0);
        m2 += qb + dl * dl * n * w;
        mean += dl * w;
        n = nt;
      }
      const double var = m2 / (double)p.V;
      p.bn_out[c] = (float)mean;
      p.bn_out[HID + c] = (float)(1.0 / sqrt(var + (double)kBnEps));
      if (p.gamma) { p.bn_out[2 * HID + c] = p.gamma[c]; p.bn_out[3 * HID + c] = p.beta[c]; }
      if (p.running) {
        const double unb = p.V > 1 ? var * (double)p.V / (double)(p.V - 1) : var;
        p.running[c] = 0.9f * p.running[c] + 0.1f * (float)mean;
        p.running[HID + c] = 0.9f * p.running[HID + c] + 0.1f * (float)unb;
      }
    }
  }
}

}  // namespace tc4

static void launch_tc4(const GinFwdPair& pp, int grid, cudaStream_t s) {
  static bool once = (cudaFuncSetAttribute(tc4::gin_fwd_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc4::Smem::total), true);
  (void)once;
  launch_k((tc4::gin_fwd_tc4_kernel), dim3(grid), dim3(tc4::kThreads4), tc4::Smem::total, s, pp);
}

// KIN = 64 layers only (layer 0 of an encoder, KIN = 32 and the ego row map, stays on gin_fwd_tc3)
void launch_gin_fwd_tc4(const GinFwdArgs& a, cudaStream_t s) {
  GinFwdPair pp;
  pp.a[0] = a; pp.a[1] = a;
  const int grid = min((a.V + tc4::TM - 1) / tc4::TM, num_sms());
  pp.split = grid;
  launch_tc4(pp, grid, s);
}
static int dbg_mask4() {
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("SCGIB_DBG"); dbg = e ? atoi(e) : 0; }
  return dbg;
}
void launch_gin_fwd_tc4_pair(const GinFwdArgs& a0, const GinFwdArgs& a1, cudaStream_t s) {
  GinFwdPair pp;
  pp.a[0] = a0; pp.a[1] = a1;
  pp.a[0].dbg = pp.a[1].dbg = dbg_mask4();
  const int t0 = (a0.V + tc4::TM - 1) / tc4::TM, t1 = (a1.V + tc4::TM - 1) / tc4::TM;
  const int grid = min(t0 + t1, num_sms());
  pp.split = pair_split(grid, t0, t1);
  launch_tc4(pp, grid, s);
}

}  // namespace scgib
extern "C" __attribute__((visibility("default"))) int scgib_debug_tc4_trace(long long* host_out, int n) {
  return (int)cudaMemcpyFromSymbol(host_out, scgib::g_tc4_trace, (size_t)n * sizeof(long long));
}
