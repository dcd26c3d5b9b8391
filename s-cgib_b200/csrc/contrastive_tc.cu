// contrastive_tc.cu - the B x B similarity blocks of the contrastive loss on the 5th-generation tensor cores.
//
// reference: sim / batched_semi_loss, models.py:606-629:  refl = exp(z1 z1^T), between = exp(z1 z2^T), row sums.
// Forward: CTA = 128 rows of z1 (UMMA M = 128) x a strided set of 64-row column blocks.  Both operands are K-major
// (contraction over the 64 features), so the normalised rows - pre-split into tf32 hi/lo by normalize_kernel - are
// copied with cp.async straight into the no-swizzle core-matrix layout: no register staging at all.  Pipeline per
// CTA: cp.async of block t+2, tcgen05.mma of block t+1 (3xTF32, accumulators in the other half of TMEM) and the
// exp / row-sum epilogue of block t (tcgen05.ld) overlap.
#include "kernels.cuh"
#include "umma.cuh"
#include "side_jobs.cuh"

namespace scgib {
using namespace umma;

constexpr int CI = 128;          // rows of z1 per CTA
constexpr int CJ = 64;           // columns (rows of z1 / z2) per block
constexpr int kCore = 128;       // dense cores: tiles are written by cp.async only
constexpr int ZI_BYTES = tile_bytes(CI, HID, kCore);   // 32768
constexpr int ZJ_BYTES = tile_bytes(CJ, HID, kCore);   // 16384
constexpr uint32_t kIdescC = idesc_tf32(CI, CJ, false, false);

struct ConTcLayout {
  static constexpr int off_zi = 0;                          // hi, lo
  static constexpr int off_zj = 2 * ZI_BYTES;               // [2 buffers][z1h, z1l, z2h, z2l]
  static constexpr int off_rs = off_zj + 2 * 4 * ZJ_BYTES;  // float [2][128]
  static constexpr int off_bar = off_rs + 2 * CI * 4;
  static constexpr int total = off_bar + 32;
};

// copy rows [base, base+R) of a [B][64] matrix into a K-major core-matrix tile (rows >= B zero-filled)
template <int R>
__device__ __forceinline__ void cp_async_tile_g(unsigned char* dst, const float* __restrict__ src, int base, int B) {
  for (int i = threadIdx.x; i < R * 16; i += kThreads) {
    const int r = i >> 4, c4 = i & 15;
    const bool ok = base + r < B;
    cp_async16(dst + tile_off4(HID, r, c4, kCore), src + (size_t)(ok ? base + r : 0) * HID + c4 * 4, ok);
  }
}

// Grid: CTAs [0, iblocks * jsplit) are the similarity CTAs (row block b % iblocks, column split b / iblocks); the CTAs
// after them run the side jobs of ConFwdSides (recon_reduce, then compressor_ema); the CTA that finishes last runs
// loss_finalize when sd.finalize is set.
__device__ __forceinline__ void con_fwd_finish(const ConFwdSides& sd) {
  if (!sd.finalize) return;
  if (!last_cta_arrives(sd.counter)) return;
  loss_finalize_body<kThreads>(sd.fin);
}

__global__ void __launch_bounds__(kThreads, 1)
contrastive_fwd_tc_kernel(ContrastiveFwdArgs p, ConFwdSides sd, int iblocks) {
  pdl_sync();
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nmain = iblocks * p.jsplit;
  if ((int)blockIdx.x >= nmain) {
    const int side = (int)blockIdx.x - nmain;
    if (side < sd.n_reduce) recon_reduce_body(sd.rpart, sd.rgrid, sd.G, sd.edge, HID, side, sd.n_reduce);
    else compressor_ema_body<HID, kThreads>(sd.cstat, p.B, sd.running);
    con_fwd_finish(sd);
    return;
  }
  const int bx = (int)blockIdx.x % iblocks, by = (int)blockIdx.x / iblocks, gy = p.jsplit;
  unsigned char* zi_hi = smem + ConTcLayout::off_zi;
  unsigned char* zi_lo = zi_hi + ZI_BYTES;
  unsigned char* zj = smem + ConTcLayout::off_zj;
  float* s_rs = reinterpret_cast<float*>(smem + ConTcLayout::off_rs);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + ConTcLayout::off_bar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + ConTcLayout::off_bar + 16);
  const float* z1h = p.zsplit;
  const float* z1l = z1h + (size_t)p.B * HID;
  const float* z2h = z1l + (size_t)p.B * HID;
  const float* z2l = z2h + (size_t)p.B * HID;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ibase = bx * CI;
  const int jblocks = (p.B + CJ - 1) / CJ;
  const int nblk = (jblocks - by + gy - 1) / gy;   // blocks of this CTA
  auto jb_of = [&](int t) { return by + t * gy; };
  auto load_j = [&](int t) {
    unsigned char* buf = zj + (t & 1) * 4 * ZJ_BYTES;
    const int jbase = jb_of(t) * CJ;
    cp_async_tile_g<CJ>(buf, z1h, jbase, p.B);
    cp_async_tile_g<CJ>(buf + ZJ_BYTES, z1l, jbase, p.B);
    cp_async_tile_g<CJ>(buf + 2 * ZJ_BYTES, z2h, jbase, p.B);
    cp_async_tile_g<CJ>(buf + 3 * ZJ_BYTES, z2l, jbase, p.B);
  };

  if (warp == 0) tmem_alloc(s_tmem, 256);
  if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
  cp_async_tile_g<CI>(zi_hi, z1h, ibase, p.B);
  cp_async_tile_g<CI>(zi_lo, z1l, ibase, p.B);
  if (nblk > 0) load_j(0);
  cp_async_commit();
  if (nblk > 1) load_j(1);
  cp_async_commit();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;
  const Operand opI{smem_u32(zi_hi), smem_u32(zi_lo), false, (uint32_t)group_bytes(HID, kCore), kCore};
  auto issue = [&](int t) {   // one thread: D1[set] = zi z1_j^T, D2[set] = zi z2_j^T
    const uint32_t buf = smem_u32(zj + (t & 1) * 4 * ZJ_BYTES);
    const Operand o1{buf, buf + ZJ_BYTES, false, (uint32_t)group_bytes(HID, kCore), kCore};
    const Operand o2{buf + 2 * ZJ_BYTES, buf + 3 * ZJ_BYTES, false, (uint32_t)group_bytes(HID, kCore), kCore};
    const uint32_t d = tmem + (t & 1) * 128;
    gemm_3xtf32(d, opI, o1, HID / 8, kIdescC, false);
    gemm_3xtf32(d + 64, opI, o2, HID / 8, kIdescC, false);
    mma_commit(&s_bar[t & 1]);
  };
  // first block
  asm volatile("cp.async.wait_group 1;" ::: "memory");
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (threadIdx.x == 0 && nblk > 0) issue(0);

  const int row = 32 * (warp & 3) + lane;            // UMMA M = 128: accumulator row = TMEM lane
  const int hcol = (warp >> 2) * 32;
  const int gi = ibase + row;
  const uint32_t tl = (uint32_t)(32 * (warp & 3)) << 16;
  float rs = 0.f;
  for (int t = 0; t < nblk; ++t) {
    // A. block t+1 has landed -> issue its MMAs into the other TMEM half (its previous content was consumed in t-1)
    cp_async_wait_all();
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (threadIdx.x == 0 && t + 1 < nblk) issue(t + 1);
    // B. block t's MMAs are done -> its smem buffer is free
    mbar_wait(&s_bar[t & 1], (uint32_t)((t >> 1) & 1));
    fence_after_sync();
    // C. prefetch block t+2 into the freed buffer
    if (t + 2 < nblk) load_j(t + 2);
    cp_async_commit();
    // D. epilogue of block t: exp and masked row sums
    const int jbase = jb_of(t) * CJ;
    const uint32_t d = tmem + (t & 1) * 128 + tl + hcol;
#pragma unroll
    for (int part = 0; part < 2; ++part) {           // part 0: refl (z1 z1^T, diagonal excluded); part 1: between
#pragma unroll
      for (int c16 = 0; c16 < 2; ++c16) {
        float v[16];
        tmem_ld16(d + part * 64 + c16 * 16, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int gj = jbase + hcol + c16 * 16 + j;
          const bool use = (gj < p.B) && (part == 1 || gj != gi);
          rs += use ? __expf(v[j]) : 0.f;   // |s| <= 1: ex2.approx is good to ~2e-7 relative
        }
      }
    }
    fence_before_sync();
  }
  s_rs[(warp >> 2) * CI + row] = rs;
  __syncthreads();
  if (threadIdx.x < CI && ibase + threadIdx.x < p.B)
    p.rowsum[(size_t)by * p.B + ibase + threadIdx.x] = s_rs[threadIdx.x] + s_rs[CI + threadIdx.x];
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
  con_fwd_finish(sd);
}

// ------------------------------------------------------------------------------------------------
// backward (same math as contrastive_bwd_kernel, loss_kernels.cu):
//   blockIdx.z == 0 (rows i of z1):  g1_i = sum_{j != i} e^{z1_i.z1_j} (1/D_i + 1/D_j) z1_j + sum_j e^{z1_i.z2_j}/D_i z2_j
//   blockIdx.z == 1 (rows j of z2):  g2_j = sum_i e^{z1_i.z2_j}/D_i z1_i
// Per 128-row block and 64-row column block: S = Zrow Zcol^T (3xTF32, K-major operands) -> TMEM; the 256 threads turn
// S into the weighted probabilities P (exp, 1/D weights, masks), split P into tf32 hi/lo and write it BACK to tensor
// memory, where it is the A operand of the second GEMM  acc += P Zcol  (B = the same shared-memory column tile, now
// read MN-major; its hi and lo copies are adjacent, so one N = 128 MMA yields P_hi Z_hi and P_hi Z_lo).  P never touches
// shared memory and the B x B matrices never exist in HBM.  The similarity GEMMs of block t+1 are issued before the
// epilogue of block t (double-buffered S columns), column tiles are double-buffered with cp.async.
// ------------------------------------------------------------------------------------------------
constexpr int ZS_BYTES = tile_s_bytes(CJ);                 // 16384: one [64][64] tile in format S
constexpr uint32_t kIdSim = idesc_tf32(CI, CJ, false, false);
constexpr uint32_t kIdPVa = idesc_tf32(CI, 2 * HID, false, true), kIdPVb = idesc_tf32(CI, HID, false, true);
// TMEM: S set s at 128 s (S1 | S2), P_hi 256, P_lo 320, acc 384 (128 columns, N-stacked)
constexpr int kColP = 256, kColAcc = 384;

struct ConBwdTcLayout {
  static constexpr int off_zi = 0;                            // hi, lo (dense cores, K-major A)
  static constexpr int off_zj = 2 * ZI_BYTES;                 // [2 stages][z1 hi, z1 lo, z2 hi, z2 lo] format S
  static constexpr int off_di = off_zj + 2 * 4 * ZS_BYTES;    // float [128]
  static constexpr int off_dj = off_di + CI * 4;              // float [2][64]
  static constexpr int off_bar = off_dj + 2 * CJ * 4;         // sim[2], pv
  static constexpr int total = off_bar + 64;
};

template <int R>
__device__ __forceinline__ void cp_async_tile_s(unsigned char* dst, const float* __restrict__ src, int base, int B) {
  for (int i = threadIdx.x; i < R * 16; i += kThreads) {
    const int r = i >> 4, c4 = i & 15;
    const bool ok = base + r < B;
    cp_async16(dst + tile_s_off4(R, r, c4), src + (size_t)(ok ? base + r : 0) * HID + c4 * 4, ok);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
contrastive_bwd_tc_kernel(ContrastiveBwdArgs p, const float* __restrict__ zsplit, ConBwdSides sd, int iblocks) {
  pdl_sync();
  using L = ConBwdTcLayout;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nmain = iblocks * p.jsplit;
  if ((int)blockIdx.x >= nmain) {          // side CTAs: the adjacency-reconstruction backward (independent of this kernel's work)
    recon_bwd_body<HID>(sd.recon, smem, (int)blockIdx.x - nmain, sd.n_recon);
    return;
  }
  const int bx = (int)blockIdx.x % iblocks, by = (int)blockIdx.x / iblocks, gy = p.jsplit;
  unsigned char* zi_hi = smem + L::off_zi;
  unsigned char* zi_lo = zi_hi + ZI_BYTES;
  unsigned char* zj = smem + L::off_zj;
  float* s_di = reinterpret_cast<float*>(smem + L::off_di);
  float* s_dj = reinterpret_cast<float*>(smem + L::off_dj);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::off_bar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::off_bar + 32);
  const size_t n = (size_t)p.B * HID;
  const float *z1h = zsplit, *z1l = zsplit + n, *z2h = zsplit + 2 * n, *z2l = zsplit + 3 * n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ibase = bx * CI;
  const int jblocks = (p.B + CJ - 1) / CJ;
  const int nblk = (jblocks - by + gy - 1) / gy;
  auto jb_of = [&](int t) { return (by + t * gy) * CJ; };

  if (warp == 0) tmem_alloc(s_tmem, 512);
  if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_init(&s_bar[2], 1); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;
  const uint32_t zih = smem_u32(zi_hi), zil = smem_u32(zi_lo);
  const int row = 32 * (warp & 3) + lane;              // TMEM lane = row of the 128-row block
  const int hcol = (warp >> 2) * 32;                   // this thread's half of the 64 columns
  const int gi = ibase + row;
  const uint32_t tl = (uint32_t)(32 * (warp & 3)) << 16;
  int sim_n[2] = {0, 0};                               // completed waits per S set (mbarrier phase bookkeeping)
  int pv_issued = 0, pv_waited = 0;                    // PV GEMM groups committed / observed complete (block-uniform)
  auto wait_pv = [&]() {                               // every committed PV GEMM has completed (P and its column tile are free)
    while (pv_waited < pv_issued) { mbar_wait(&s_bar[2], (uint32_t)(pv_waited & 1)); ++pv_waited; }
    fence_after_sync();
  };

  // Both gradient halves in one CTA (balanced work, single wave): mode 0 = rows of z1 (g1), mode 1 = rows of z2 (g2)
  for (int mode = 0; mode < 2; ++mode) {
    const bool mode1 = (mode == 1);
    auto load_j = [&](int t) {
      unsigned char* buf = zj + (t & 1) * 4 * ZS_BYTES;
      const int jbase = jb_of(t);
      cp_async_tile_s<CJ>(buf, z1h, jbase, p.B);
      cp_async_tile_s<CJ>(buf + ZS_BYTES, z1l, jbase, p.B);
      if (!mode1) {
        cp_async_tile_s<CJ>(buf + 2 * ZS_BYTES, z2h, jbase, p.B);
        cp_async_tile_s<CJ>(buf + 3 * ZS_BYTES, z2l, jbase, p.B);
      }
      if (threadIdx.x < CJ) s_dj[(t & 1) * CJ + threadIdx.x] = (jbase + threadIdx.x < p.B) ? 1.f / __ldg(p.D + jbase + threadIdx.x) : 0.f;
    };
    // similarity GEMMs of block t into S set t&1 (whole warp 0 runs this; one elected lane issues)
    auto issue_sim = [&](int t) {
      const uint32_t buf = smem_u32(zj + (t & 1) * 4 * ZS_BYTES);
      const uint32_t d = tmem + (t & 1) * 128;
      const int nsim = mode1 ? 1 : 2;
      for (int q = 0; q < nsim; ++q) {
        const uint32_t bh = buf + q * 2 * ZS_BYTES, bl = bh + ZS_BYTES;
#pragma unroll
        for (int k = 0; k < HID / 8; ++k) {
          const uint64_t ah = desc_g_dense(zih, HID, k), al = desc_g_dense(zil, HID, k);
          const uint64_t dbh = desc_s_kmajor(bh, CJ, k), dbl = desc_s_kmajor(bl, CJ, k);
          mma_tf32_w(d + q * 64, al, dbh, kIdSim, k > 0);
          mma_tf32_w(d + q * 64, ah, dbl, kIdSim, true);
          mma_tf32_w(d + q * 64, ah, dbh, kIdSim, true);
        }
      }
      mma_commit_w(&s_bar[t & 1]);
    };
    // acc += P Zcol for source q (0: z1 tile, 1: z2 tile) of block t; P (hi | lo) is in tensor memory
    auto issue_pv = [&](int t, int q, bool first) {
      const uint32_t bh = smem_u32(zj + (t & 1) * 4 * ZS_BYTES) + q * 2 * ZS_BYTES;
#pragma unroll
      for (int k = 0; k < CJ / 8; ++k) {
        const uint64_t b = desc_s_mnmajor(bh, CJ, k);
        mma_tf32_ta_w(tmem + kColAcc, tmem + kColP + 8 * k, b, kIdPVa, !(first && k == 0));
        mma_tf32_ta_w(tmem + kColAcc, tmem + kColP + 64 + 8 * k, b, kIdPVb, true);
      }
      mma_commit_w(&s_bar[2]);
    };
    // ---- prologue of the mode: row tile (all earlier MMAs have completed: see the end of the loop body) + two column blocks
    cp_async_tile_g<CI>(zi_hi, mode1 ? z2h : z1h, ibase, p.B);
    cp_async_tile_g<CI>(zi_lo, mode1 ? z2l : z1l, ibase, p.B);
    if (mode == 0 && threadIdx.x < CI) s_di[threadIdx.x] = (ibase + threadIdx.x < p.B) ? 1.f / __ldg(p.D + ibase + threadIdx.x) : 0.f;
    if (nblk > 0) load_j(0);
    cp_async_commit();
    if (nblk > 1) load_j(1);
    cp_async_commit();
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 0 && nblk > 0) issue_sim(0);
    const float di = s_di[row];
    int pv_in_mode = 0;
    // one weighted-probability block: S (32 columns of this thread) -> P hi/lo in tensor memory
    auto make_p = [&](uint32_t s_addr, int jbase, const float* dj, int kind) {   // kind 0: refl (z1 z1), 1: between, 2: mode 1
      float v[32];
      tmem_ld16_nowait(s_addr, *reinterpret_cast<float (*)[16]>(v));
      tmem_ld16_nowait(s_addr + 16, *reinterpret_cast<float (*)[16]>(v + 16));
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int gj = jbase + hcol + j;
        float w;
        if (kind == 0) w = (gj != gi) ? di + dj[hcol + j] : 0.f;
        else if (kind == 1) w = di;
        else w = dj[hcol + j];
        v[j] = (gj < p.B && gi < p.B) ? __expf(v[j]) * w : 0.f;   // |s| <= 1: ex2.approx is good to ~2e-7 relative
      }
      wait_pv();                                         // the previous P has been consumed before it is overwritten
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { hi[j] = tf32_rna(v[16 * c + j]); lo[j] = tf32_rna(v[16 * c + j] - hi[j]); }
        tmem_st16(tmem + tl + kColP + hcol + 16 * c, hi);
        tmem_st16(tmem + tl + kColP + 64 + hcol + 16 * c, lo);
      }
      tmem_st_wait();
    };
    for (int t = 0; t < nblk; ++t) {
      // A. block t+1 has landed (its similarity GEMMs, into the other S set, are issued after the first PV GEMM below)
      cp_async_wait_all();
      fence_smem_to_async();
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      // B. similarities of block t are in tensor memory
      mbar_wait(&s_bar[t & 1], (uint32_t)(sim_n[t & 1] & 1));
      ++sim_n[t & 1];
      fence_after_sync();
      const int jbase = jb_of(t);
      const float* dj = s_dj + (t & 1) * CJ;
      const uint32_t sset = tmem + tl + (t & 1) * 128 + hcol;
      const int nsrc = mode1 ? 1 : 2;
      for (int q = 0; q < nsrc; ++q) {
        make_p(sset + q * 64, jbase, dj, mode1 ? 2 : q);
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        if (warp == 0) {
          issue_pv(t, q, pv_in_mode == 0);
          // the similarity GEMMs of block t+1 queue BEHIND the first PV GEMM of block t (the tensor pipe is in order):
          // the next make_p waits for a short PV, and its exponentials overlap the long similarity GEMMs
          if (q == 0 && t + 1 < nblk) issue_sim(t + 1);
        }
        ++pv_issued;
        ++pv_in_mode;
      }
      // C. stage t&1 is free once the PV GEMMs of block t have completed: prefetch block t+2 into it
      if (t + 2 < nblk) { wait_pv(); load_j(t + 2); }
      cp_async_commit();
    }
    // ---- result rows of this mode: acc = columns 0..63 + columns 64..127
    float* out = (mode1 ? p.g2p : p.g1p) + (size_t)by * p.B * HID;
    if (nblk > 0) {
      wait_pv();
      float a0[32], a1[32];
      tmem_ld16_nowait(tmem + tl + kColAcc + hcol, *reinterpret_cast<float (*)[16]>(a0));
      tmem_ld16_nowait(tmem + tl + kColAcc + hcol + 16, *reinterpret_cast<float (*)[16]>(a0 + 16));
      tmem_ld16_nowait(tmem + tl + kColAcc + 64 + hcol, *reinterpret_cast<float (*)[16]>(a1));
      tmem_ld16_nowait(tmem + tl + kColAcc + 64 + hcol + 16, *reinterpret_cast<float (*)[16]>(a1 + 16));
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) a0[j] += a1[j];
      if (gi < p.B) {
#pragma unroll
        for (int j = 0; j < 4; ++j) st8(out + (size_t)gi * HID + hcol + 8 * j, a0 + 8 * j);
      }
    } else if (gi < p.B) {
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) st8(out + (size_t)gi * HID + hcol + 8 * j, z);
    }
    // every MMA of this mode has completed (wait_pv covers the similarity GEMMs too: in-order completion) and every
    // thread has read its accumulator rows before the next mode overwrites tiles and tensor memory
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static_assert(ConBwdTcLayout::total >= (int)sizeof(ReconBwdSmem<HID>), "recon_bwd side CTAs use the contrastive kernel's shared memory");

void launch_contrastive_fwd_tc_sides(const ContrastiveFwdArgs& a, const ConFwdSides& sides, cudaStream_t s) {
  static bool once = (cudaFuncSetAttribute(contrastive_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           ConTcLayout::total), true);
  (void)once;
  const int iblocks = (a.B + CI - 1) / CI;
  const int grid = iblocks * a.jsplit + sides.n_reduce + sides.n_ema;
  launch_k((contrastive_fwd_tc_kernel), dim3(grid), dim3(kThreads), ConTcLayout::total, s, a, sides, iblocks);
}
void launch_contrastive_fwd_tc(const ContrastiveFwdArgs& a, cudaStream_t s) { launch_contrastive_fwd_tc_sides(a, ConFwdSides{}, s); }

void launch_contrastive_bwd_tc_sides(const ContrastiveBwdArgs& a, const float* zsplit, const ConBwdSides& sides, cudaStream_t s) {
  static bool once = (cudaFuncSetAttribute(contrastive_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           ConBwdTcLayout::total), true);
  (void)once;
  const int iblocks = (a.B + CI - 1) / CI;
  const int grid = iblocks * a.jsplit + sides.n_recon;
  launch_k((contrastive_bwd_tc_kernel), dim3(grid), dim3(kThreads), ConBwdTcLayout::total, s, a, zsplit, sides, iblocks);
}
void launch_contrastive_bwd_tc(const ContrastiveBwdArgs& a, const float* zsplit, cudaStream_t s) {
  launch_contrastive_bwd_tc_sides(a, zsplit, ConBwdSides{}, s);
}

}  // namespace scgib
