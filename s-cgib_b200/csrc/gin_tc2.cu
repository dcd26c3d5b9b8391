// gin_tc2.cu - GIN layer forward as a warp-specialised, persistent tcgen05 kernel (one CTA per SM).
//
// Same contract as gin_fwd_kernel (gin_kernels.cu; reference models.py:66-72: DGL GINConv 'sum' + MLP + BatchNorm1d
// statistics).  Roles inside the CTA:
//   producer warps (8 or 16): CSR gather-aggregate of a 128-row tile (float4 lanes, BN+ReLU of the previous layer
//       applied on load), `a` -> global (saved) and, split into tf32 hi/lo, -> a shared-memory stage in the swizzled
//       tile format S (umma.cuh).  The gather is latency-bound (degree ~2, 256-byte rows), so it is organised for
//       memory-level parallelism: the tile's indptr slice and its (contiguous) indices range - with the ego row map
//       already applied - are staged in shared memory ONE TILE AHEAD, and every thread has the self row and the first
//       two neighbour rows of all its rows in flight at once (no dependent index -> row chain in the tile's own loop);
//   MMA warp (one thread): GEMM1 u = a W1^T (A from shared memory, K-major view of the stage), GEMM2 y = r W2^T with
//       the A operand in TENSOR MEMORY - both as 3xTF32 with fp32 accumulation in TMEM;
//   epilogue warps (4, thread = tile row = TMEM lane): r = relu(u + b1) -> global (saved) and hi/lo back into TMEM
//       (r never touches shared memory), y = acc + b2 -> global, column statistics by a warp transpose-reduce.
// Two stages (shared-memory `a` tiles and TMEM column sets) keep gather, tensor pipe and epilogue of different tiles
// in flight at once; all hand-offs are mbarriers (no __syncthreads in the main loop).
// Batch statistics: every warp keeps a running Chan (n, mean, M2) of its rows, the CTA publishes one partial, the last
// CTA combines them in fp64 in a fixed order (deterministic).
#include <stdlib.h>
#include "kernels.cuh"
#include "umma.cuh"

namespace scgib {
using namespace umma;

// per-tile role timestamps (SCGIB_DBG bit 1024; experiments only): [cta][tile < 16][event < 12] SM clock values
__device__ long long g_tc2_trace[160 * 16 * 12];   // also written by gin_tc3.cu
#define TC2_TRACE(ev, tile) do { if ((p.dbg & 1024) && (tile) < 16 && blockIdx.x < 160) g_tc2_trace[((size_t)blockIdx.x * 16 + (tile)) * 12 + (ev)] = clock64(); } while (0)

namespace tc2 {
constexpr int TM = 128;                       // rows per tile = UMMA M
constexpr int kEpiWarps = 4;
constexpr int kIdxCap = 1024;                 // staged neighbour indices per tile (more edges: read from global)
constexpr int kStageBytes = 2 * TM * HID * 4; // hi + lo tile of one stage (format S)
// 3xTF32 with the B operands stacked along N: the hi and lo weight tiles are adjacent in shared memory, so ONE N = 128
// MMA computes A_hi W_hi^T (columns 0..63) and A_hi W_lo^T (columns 64..127); a second N = 64 MMA adds A_lo W_hi^T
// into columns 0..63 and the epilogue sums the two halves.  (A tcgen05.mma with N <= 64 costs ~54 cycles whatever N
// is, N = 128 costs 64: measured with tests/gpu_umma_time.py.)
constexpr uint32_t kIdesc = idesc_tf32(TM, HID, false, false);
constexpr uint32_t kIdesc2 = idesc_tf32(TM, 2 * HID, false, false);
// TMEM columns of stage s (base + 256 s): D1 (128 columns: hi*hi | hi*lo halves) | D2 (128).  Epilogue 1 overwrites D1
// IN PLACE with r_hi (columns 0..63) and r_lo (64..127) - every thread has read both halves of its row chunk before it
// stores - so GEMM1 of the next user of the stage depends only on the in-order tensor pipe, not on epilogue 2.
constexpr int kColD1 = 0, kColRhi = 0, kColRlo = 64, kColD2 = 128;

template <int KIN>
struct Smem {
  static constexpr int W1B = HID * KIN * 4, W2B = HID * HID * 4;
  static constexpr int off_stage = 0;
  static constexpr int off_w1_hi = 2 * kStageBytes, off_w1_lo = off_w1_hi + W1B;
  static constexpr int off_w2_hi = off_w1_lo + W1B, off_w2_lo = off_w2_hi + W2B;
  static constexpr int off_f = off_w2_lo + W2B;                       // b1[64] b2[64]
  static constexpr int off_stat = off_f + 2 * HID * 4;                // double [4 warps][3][64]
  static constexpr int off_dred = off_stat + kEpiWarps * 3 * HID * 8; // double [4][64]
  static constexpr int off_bar = off_dred + 4 * HID * 8;              // 12 mbarriers + tmem slot
  static constexpr int off_ip = off_bar + 128;                        // int [2][TM + 4]   indptr slice of the tile
  static constexpr int off_self = off_ip + 2 * (TM + 4) * 4;          // int [2][TM]       mapped self row
  static constexpr int off_ix = off_self + 2 * TM * 4;                // int [2][kIdxCap]  mapped neighbour rows
  static constexpr int total = off_ix + 2 * kIdxCap * 4;
};

// barrier indices (each x2 stages)
enum { B_FULL_A = 0, B_EMPTY_A = 2, B_D1 = 4, B_R = 6, B_D2 = 8, B_DFREE = 10, B_COUNT = 12 };

// column sums of a [32 rows (lanes)][32 columns (registers)] block by a transpose-reduce: lane l returns column l
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

template <int KIN, int PW>
__global__ void __launch_bounds__((kEpiWarps + 1 + PW) * 32, 1)
gin_fwd_tc2_kernel(GinFwdPair pp) {
  using L = Smem<KIN>;
  const bool second = (int)blockIdx.x >= pp.split;
  const GinFwdArgs& p = pp.a[second ? 1 : 0];
  const int bid = second ? (int)blockIdx.x - pp.split : (int)blockIdx.x;          // CTA index / count inside its problem
  const int nblk = second ? (int)gridDim.x - pp.split : pp.split;
  extern __shared__ __align__(1024) unsigned char smem[];
  float* s_b1 = reinterpret_cast<float*>(smem + L::off_f);
  float* s_b2 = s_b1 + HID;
  double* s_stat = reinterpret_cast<double*>(smem + L::off_stat);
  double* s_dred = reinterpret_cast<double*>(smem + L::off_dred);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::off_bar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::off_bar + B_COUNT * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.V + TM - 1) / TM;
  const int my_tiles = (n_tiles - bid + nblk - 1) / nblk;   // tiles bid + i*nblk

  // ---- one-time setup: barriers, TMEM, weights (natural [out][in] = K-major B operand, dense cores), biases
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars[B_FULL_A + s], PW);
      mbar_init(&bars[B_EMPTY_A + s], 1);
      mbar_init(&bars[B_D1 + s], 1);
      mbar_init(&bars[B_R + s], kEpiWarps * 32);
      mbar_init(&bars[B_D2 + s], 1);
      mbar_init(&bars[B_DFREE + s], kEpiWarps * 32);
    }
  }
  if (warp == kEpiWarps) tmem_alloc(s_tmem, 512);
  if (!(p.dbg & 512))
  for (int i = threadIdx.x; i < HID * (KIN / 4); i += blockDim.x) {
    const int o = i / (KIN / 4), c4 = i % (KIN / 4);
    store_split4(smem + L::off_w1_hi, smem + L::off_w1_lo, KIN, o, c4, ldg4(p.W1 + (size_t)o * KIN + c4 * 4), 128);
  }
  if (!(p.dbg & 512))
  for (int i = threadIdx.x; i < HID * (HID / 4); i += blockDim.x) {
    const int o = i / (HID / 4), c4 = i % (HID / 4);
    store_split4(smem + L::off_w2_hi, smem + L::off_w2_lo, HID, o, c4, ldg4(p.W2 + (size_t)o * HID + c4 * 4), 128);
  }
  if (threadIdx.x < HID) { s_b1[threadIdx.x] = p.b1[threadIdx.x]; s_b2[threadIdx.x] = p.b2[threadIdx.x]; }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;

  if (warp >= kEpiWarps + 1) {
    // =========================================================================== producers
    constexpr int PT = PW * 32;                                // producer threads
    constexpr int LPR = KIN / 4, RPP = PT / LPR, NR = TM / RPP;
    const int pt = (warp - (kEpiWarps + 1)) * 32 + lane;
    const int gl = pt % LPR, gr = pt / LPR;
    int* s_ip = reinterpret_cast<int*>(smem + L::off_ip);
    int* s_self = reinterpret_cast<int*>(smem + L::off_self);
    int* s_ix = reinterpret_cast<int*>(smem + L::off_ix);
    Bn4 bn;
    const bool has_bn = (p.bn_in != nullptr);
    if (has_bn) bn.load(p.bn_in, gl * 4);
    auto act = [&](float4 h) { return has_bn ? bn.act(h) : h; };
    auto tile_base = [&](int i) { return (bid + i * nblk) * TM; };
    auto prod_sync = [&]() { asm volatile("bar.sync 1, %0;" :: "n"(PT) : "memory"); };
    // indptr slice, mapped self rows and (unmapped) neighbour rows of a tile -> index buffer `buf`, asynchronously
    // (cp.async: no register staging, nothing waits until the end of the iteration)
    auto stage_indices = [&](int buf, int base, int e_begin, int e_end) {
      for (int r = pt; r <= TM; r += PT) cp_async4(&s_ip[buf * (TM + 4) + r], p.indptr + min(base + r, p.V));
      if (p.row_map)
        for (int r = pt; r < TM; r += PT) cp_async4(&s_self[buf * TM + r], p.row_map + min(base + r, p.V - 1));
      const int n = min(e_end - e_begin, kIdxCap);
      for (int e = pt; e < n; e += PT) cp_async4(&s_ix[buf * kIdxCap + e], p.indices + e_begin + e);
      cp_async_commit();
    };
    int nb_begin = 0, nb_end = 0;                              // indices range of the NEXT tile (loaded one tile ahead)
    if (my_tiles > 0) {
      const int b0 = tile_base(0);
      stage_indices(0, b0, __ldg(p.indptr + b0), __ldg(p.indptr + min(b0 + TM, p.V)));
      if (my_tiles > 1) { const int b1 = tile_base(1); nb_begin = __ldg(p.indptr + b1); nb_end = __ldg(p.indptr + min(b1 + TM, p.V)); }
    }
    cp_async_wait_all();
    prod_sync();
    for (int i = 0; i < my_tiles; ++i) {
      const int s = i & 1, use = i >> 1;
      const int base = tile_base(i);
      const int* ip = s_ip + s * (TM + 4);
      const int* ix = s_ix + s * kIdxCap;
      const int e_begin = ip[0];
      if (pt == 0) TC2_TRACE(0, i);
      auto nbr = [&](int e) {                                  // mapped neighbour row of tile-relative edge e
        const int u = (e < kIdxCap) ? ix[e] : __ldg(p.indices + e_begin + e);
        return p.row_map ? __ldg(p.row_map + u) : u;
      };
      unsigned char* hi = smem + L::off_stage + s * kStageBytes;
      unsigned char* lo = hi + TM * KIN * 4;
      // The thread's NR rows are processed in batches of NB: registers hold 3 NB rows in flight (no spills).
      constexpr int NB = NR > 4 ? 4 : NR;
#pragma unroll 1
      for (int jb = 0; jb < NR; jb += NB) {
        // ---- [A] self row + first two neighbour rows of every row of the batch: all loads in flight together
        int deg[NB], eoff[NB];
        float4 h0[NB], h1[NB], h2[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int r = gr + (jb + j) * RPP, v = base + r;
          const bool ok = v < p.V && !(p.dbg & 4);
          const int p0 = ip[r], p1 = ip[r + 1];
          deg[j] = ok ? p1 - p0 : 0;
          eoff[j] = p0 - e_begin;
          const int sv = p.row_map ? s_self[s * TM + r] : v;
          h0[j] = ok ? ld4(p.in + (size_t)sv * KIN + gl * 4) : make4(0.f);
          h1[j] = deg[j] > 0 ? ld4(p.in + (size_t)nbr(eoff[j]) * KIN + gl * 4) : make4(0.f);
          h2[j] = deg[j] > 1 ? ld4(p.in + (size_t)nbr(eoff[j] + 1) * KIN + gl * 4) : make4(0.f);
        }
        if (pt == 0 && jb == 0) TC2_TRACE(1, i);
        // ---- [B] stage the next tile's indices while those loads are in flight; fetch the range of the tile after it
        if (jb == 0 && i + 1 < my_tiles) {
          stage_indices((i + 1) & 1, tile_base(i + 1), nb_begin, nb_end);
          if (i + 2 < my_tiles) { const int b2 = tile_base(i + 2); nb_begin = __ldg(p.indptr + b2); nb_end = __ldg(p.indptr + min(b2 + TM, p.V)); }
        }
        // ---- [C] a_v = f(h_v) + sum_u f(h_u), neighbours in CSR order (same summation order as gin_fwd_kernel)
        float4 agg[NB];
        int maxd = 0;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          agg[j] = (base + gr + (jb + j) * RPP < p.V) ? act(h0[j]) : make4(0.f);
          if (deg[j] > 0) agg[j] = add4(agg[j], act(h1[j]));
          if (deg[j] > 1) agg[j] = add4(agg[j], act(h2[j]));
          maxd = max(maxd, deg[j]);
        }
        for (int d = 2; d < maxd; d += 2) {                      // rows with more than two neighbours
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            h1[j] = deg[j] > d ? ld4(p.in + (size_t)nbr(eoff[j] + d) * KIN + gl * 4) : make4(0.f);
            h2[j] = deg[j] > d + 1 ? ld4(p.in + (size_t)nbr(eoff[j] + d + 1) * KIN + gl * 4) : make4(0.f);
          }
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            if (deg[j] > d) agg[j] = add4(agg[j], act(h1[j]));
            if (deg[j] > d + 1) agg[j] = add4(agg[j], act(h2[j]));
          }
        }
        if (pt == 0 && jb == 0) TC2_TRACE(2, i);
        if (jb == 0 && use > 0) mbar_wait(&bars[B_EMPTY_A + s], (uint32_t)((use - 1) & 1));   // GEMM1 of the previous user is done
        if (pt == 0 && jb == 0) TC2_TRACE(3, i);
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int r = gr + (jb + j) * RPP;
          if (p.a_out && base + r < p.V && !(p.dbg & 8)) st4_cs(p.a_out + (size_t)(base + r) * KIN + gl * 4, agg[j]);
          if (!(p.dbg & 16)) store_split4_s(hi, lo, TM, r, gl, agg[j]);
        }
      }
      if (!(p.dbg & 128)) fence_smem_to_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_FULL_A + s]);
      if (pt == 0) TC2_TRACE(4, i);
      cp_async_wait_all();
      prod_sync();   // index buffer (i+1)&1 is complete, buffer i&1 may be overwritten by the next iteration
    }
  } else if (warp == kEpiWarps) {
    // =========================================================================== MMA issuer
    {   // the whole warp runs the loop (uniform descriptors); one elected lane issues each instruction
      const uint32_t w1h = smem_u32(smem + L::off_w1_hi), w2h = smem_u32(smem + L::off_w2_hi);
      static_assert(L::off_w1_lo == L::off_w1_hi + L::W1B && L::off_w2_lo == L::off_w2_hi + L::W2B, "hi/lo weight tiles must be adjacent");
      auto gemm1 = [&](int i) {
        const int s = i & 1, use = i >> 1;
        mbar_wait(&bars[B_FULL_A + s], (uint32_t)(use & 1));
        fence_after_sync();
        if (lane == 0) TC2_TRACE(5, i);
        const uint32_t ah = smem_u32(smem + L::off_stage + s * kStageBytes), al = ah + TM * KIN * 4;
        const uint32_t d = tmem + s * 256 + kColD1;
#pragma unroll
        for (int k = 0; k < KIN / 8; ++k) {
          if (p.dbg & 64) break;
          const uint64_t dah = desc_s_kmajor(ah, TM, k), dal = desc_s_kmajor(al, TM, k);
          const uint64_t dbh = desc_g_dense(w1h, KIN, k);   // hi tile; the N = 128 view continues into the lo tile
          mma_tf32_w(d, dah, dbh, kIdesc2, k > 0);
          mma_tf32_w(d, dal, dbh, kIdesc, true);
        }
        mma_commit_w(&bars[B_D1 + s]);
        mma_commit_w(&bars[B_EMPTY_A + s]);
      };
      auto gemm2 = [&](int i) {
        const int s = i & 1, use = i >> 1;
        mbar_wait(&bars[B_R + s], (uint32_t)(use & 1));
        fence_after_sync();
        if (lane == 0) TC2_TRACE(6, i);
        const uint32_t d = tmem + s * 256 + kColD2;
        const uint32_t rh = tmem + s * 256 + kColRhi, rl = tmem + s * 256 + kColRlo;
#pragma unroll
        for (int k = 0; k < HID / 8; ++k) {
          if (p.dbg & 64) break;
          const uint64_t dbh = desc_g_dense(w2h, HID, k);
          mma_tf32_ta_w(d, rh + 8 * k, dbh, kIdesc2, k > 0);
          mma_tf32_ta_w(d, rl + 8 * k, dbh, kIdesc, true);
        }
        mma_commit_w(&bars[B_D2 + s]);
      };
      if (my_tiles > 0) gemm1(0);
      if (my_tiles > 1) gemm1(1);
      for (int i = 0; i < my_tiles; ++i) {
        gemm2(i);
        if (i + 2 < my_tiles) gemm1(i + 2);
      }
    }
  } else {
    // =========================================================================== epilogue (thread = row = TMEM lane)
    const int row = warp * 32 + lane;
    const uint32_t tl = (uint32_t)(warp * 32) << 16;
    double run_n = 0.0, run_mean[2] = {0.0, 0.0}, run_m2[2] = {0.0, 0.0};   // columns lane, 32 + lane
    auto epi1 = [&](int i) {
      const int s = i & 1, use = i >> 1;
      const int gv = (bid + i * nblk) * TM + row;
      mbar_wait(&bars[B_D1 + s], (uint32_t)(use & 1));
      fence_after_sync();
      if (threadIdx.x == 0) TC2_TRACE(7, i);
      const uint32_t t0 = tmem + s * 256 + tl;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (p.dbg & 32) break;
        float v[16], v2[16], hi[16], lo[16];
        tmem_ld16_nowait(t0 + kColD1 + 16 * c, v);
        tmem_ld16_nowait(t0 + kColD1 + HID + 16 * c, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] = fmaxf((v[j] + v2[j]) + s_b1[16 * c + j], 0.f);
          hi[j] = tf32_rna(v[j]);
          lo[j] = tf32_rna(v[j] - hi[j]);
        }
        tmem_st16(t0 + kColRhi + 16 * c, hi);
        tmem_st16(t0 + kColRlo + 16 * c, lo);
        if (p.r_out && gv < p.V && !(p.dbg & 1)) {
          st8_cs(p.r_out + (size_t)gv * HID + 16 * c, v);
          st8_cs(p.r_out + (size_t)gv * HID + 16 * c + 8, v + 8);
        }
      }
      tmem_st_wait();
      fence_before_sync();
      mbar_arrive(&bars[B_R + s]);
      if (threadIdx.x == 0) TC2_TRACE(8, i);
    };
    auto epi2 = [&](int i) {
      const int s = i & 1, use = i >> 1;
      const int base = (bid + i * nblk) * TM;
      const int gv = base + row;
      const bool valid = gv < p.V;
      const int cnt = max(0, min(32, p.V - (base + warp * 32)));   // valid rows of this warp (warp-uniform)
      mbar_wait(&bars[B_D2 + s], (uint32_t)(use & 1));
      fence_after_sync();
      if (threadIdx.x == 0) TC2_TRACE(9, i);
      const uint32_t t0 = tmem + s * 256 + tl + kColD2;
      float mean_h[2] = {0.f, 0.f}, m2_h[2] = {0.f, 0.f};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (p.dbg & 32) break;
        float y[32], t[32];
        tmem_ld16_nowait(t0 + 32 * h, *reinterpret_cast<float (*)[16]>(y));
        tmem_ld16_nowait(t0 + 32 * h + 16, *reinterpret_cast<float (*)[16]>(y + 16));
        tmem_ld16_nowait(t0 + HID + 32 * h, *reinterpret_cast<float (*)[16]>(t));
        tmem_ld16_nowait(t0 + HID + 32 * h + 16, *reinterpret_cast<float (*)[16]>(t + 16));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = (y[j] + t[j]) + s_b2[32 * h + j];
        if (valid && !(p.dbg & 1)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) st8(p.y_out + (size_t)gv * HID + 32 * h + 8 * j, y + 8 * j);
        }
        // statistics of this warp's rows: column mean, then centred M2 (two transpose-reduces; lane l <-> column 32h + l)
        if (cnt > 0 && !(p.dbg & 2)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = valid ? y[j] : 0.f;
          const float mu = warp_colsum32(t, lane) / (float)cnt;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = y[j] - __shfl_sync(0xffffffffu, mu, j);
            t[j] = valid ? d * d : 0.f;
          }
          mean_h[h] = mu;
          m2_h[h] = warp_colsum32(t, lane);
        }
      }
      if (cnt > 0) {   // Chan update of the running (n, mean, M2)
        const double nb = (double)cnt, nt = run_n + nb;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double dl = (double)mean_h[h] - run_mean[h];
          run_m2[h] += (double)m2_h[h] + dl * dl * run_n * nb / nt;
          run_mean[h] += dl * nb / nt;
        }
        run_n = nt;
      }
      fence_before_sync();   // TMEM reads of this stage are complete before GEMM2 of its next user (ordered through B_R)
      if (threadIdx.x == 0) TC2_TRACE(10, i);
    };
    if (my_tiles > 0) epi1(0);
    for (int i = 0; i < my_tiles; ++i) {
      if (i + 1 < my_tiles) epi1(i + 1);
      epi2(i);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      s_stat[(warp * 3 + 0) * HID + 32 * h + lane] = run_n;
      s_stat[(warp * 3 + 1) * HID + 32 * h + lane] = run_mean[h];
      s_stat[(warp * 3 + 2) * HID + 32 * h + lane] = run_m2[h];
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (warp == kEpiWarps) tmem_dealloc(tmem, 512);
  // ---- CTA partial (n, mean, M2) per column: Chan combine of the 4 epilogue warps in fp64
  if (threadIdx.x < HID) {
    const int c = threadIdx.x;
    double n = 0.0, mean = 0.0, m2 = 0.0;
    for (int w = 0; w < kEpiWarps; ++w) {
      const double nb = s_stat[(w * 3 + 0) * HID + c];
      if (nb > 0.0) {
        const double mb = s_stat[(w * 3 + 1) * HID + c], qb = s_stat[(w * 3 + 2) * HID + c];
        const double nt = n + nb, dl = mb - mean;
        m2 += qb + dl * dl * n * nb / nt;
        mean += dl * nb / nt;
        n = nt;
      }
    }
    double* part = reinterpret_cast<double*>(p.part) + (size_t)bid * 3 * HID;
    part[c] = n; part[HID + c] = mean; part[2 * HID + c] = m2;
  }
  if ((p.dbg & 256) || !last_cta_arrives(p.counter, (unsigned)nblk)) return;
  // ---- batch statistics: the last CTA combines the per-CTA partials in fp64.  Thread (c, seg) first LOADS its partials
  //      (independent loads, one L2 latency), then Chan-combines them in a fixed order; 8 segments, then a serial 8-way.
  {
    constexpr int SEGS = 4, BATCH = 8;                       // 4 segments x 64 columns = 256 threads (every variant has them)
    const int c = threadIdx.x & (HID - 1), seg = threadIdx.x >> 6;
    const double* part = reinterpret_cast<const double*>(p.part);
    double* s_comb = reinterpret_cast<double*>(smem);        // [SEGS][3][HID] (the stages are dead by now)
    if (seg < SEGS) {
      double n = 0.0, mean = 0.0, m2 = 0.0;
      for (int b0 = seg; b0 < nblk; b0 += SEGS * BATCH) {
        double pn[BATCH], pm[BATCH], pq[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {                    // independent loads first ...
          const int b = b0 + k * SEGS;
          const bool ok = b < nblk;
          pn[k] = ok ? __ldcg(part + (size_t)b * 3 * HID + c) : 0.0;
          pm[k] = ok ? __ldcg(part + (size_t)b * 3 * HID + HID + c) : 0.0;
          pq[k] = ok ? __ldcg(part + (size_t)b * 3 * HID + 2 * HID + c) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {                    // ... then the ordered Chan combine
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
      for (int w8 = 0; w8 < SEGS; ++w8) {
        const double nb = s_comb[(w8 * 3 + 0) * HID + c], mb = s_comb[(w8 * 3 + 1) * HID + c], qb = s_comb[(w8 * 3 + 2) * HID + c];
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
  (void)s_dred;
}

}  // namespace tc2

template <int KIN, int PW>
static void launch_tc2(const GinFwdPair& pp, int grid, cudaStream_t s) {
  using L = tc2::Smem<KIN>;
  static bool once = (cudaFuncSetAttribute(tc2::gin_fwd_tc2_kernel<KIN, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total), true);
  (void)once;
  constexpr int threads = (tc2::kEpiWarps + 1 + PW) * 32;
  tc2::gin_fwd_tc2_kernel<KIN, PW><<<grid, threads, L::total, s>>>(pp);
}

static int dbg_mask() {
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("SCGIB_DBG"); dbg = e ? atoi(e) : 0; }
  return dbg;
}

static void launch_tc2_any(GinFwdPair& pp, int grid, int kin, int variant, cudaStream_t s) {
  pp.a[0].dbg = pp.a[1].dbg = dbg_mask();
  if (kin == DTR) { if (variant == 2) launch_tc2<DTR, 16>(pp, grid, s); else launch_tc2<DTR, 8>(pp, grid, s); }
  else { if (variant == 2) launch_tc2<HID, 16>(pp, grid, s); else launch_tc2<HID, 8>(pp, grid, s); }
}

// variant 1: 8 producer warps, variant 2: 16 producer warps
void launch_gin_fwd_tc2(const GinFwdArgs& a, int kin, int variant, cudaStream_t s) {
  GinFwdPair pp;
  pp.a[0] = a; pp.a[1] = a;
  const int grid = min((a.V + tc2::TM - 1) / tc2::TM, num_sms());
  pp.split = grid;                                   // single problem: every CTA works on a[0]
  launch_tc2_any(pp, grid, kin, variant, s);
}

// the same layer of both encoders in one launch (separate part / counter / bn_out buffers per problem)
void launch_gin_fwd_tc2_pair(const GinFwdArgs& a0, const GinFwdArgs& a1, int kin, int variant, cudaStream_t s) {
  GinFwdPair pp;
  pp.a[0] = a0; pp.a[1] = a1;
  const int t0 = (a0.V + tc2::TM - 1) / tc2::TM, t1 = (a1.V + tc2::TM - 1) / tc2::TM;
  const int grid = min(t0 + t1, num_sms());
  pp.split = pair_split(grid, t0, t1);
  launch_tc2_any(pp, grid, kin, variant, s);
}

}  // namespace scgib
