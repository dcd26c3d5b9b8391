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

__global__ void __launch_bounds__(kThreads, 1)
contrastive_fwd_tc_kernel(ContrastiveFwdArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
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
  const int ibase = blockIdx.x * CI;
  const int jblocks = (p.B + CJ - 1) / CJ;
  const int nblk = (jblocks - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;   // blocks of this CTA
  auto jb_of = [&](int t) { return (int)blockIdx.y + t * (int)gridDim.y; };
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
          rs += use ? expf(v[j]) : 0.f;
        }
      }
    }
    fence_before_sync();
  }
  s_rs[(warp >> 2) * CI + row] = rs;
  __syncthreads();
  if (threadIdx.x < CI && ibase + threadIdx.x < p.B)
    p.rowsum[(size_t)blockIdx.y * p.B + ibase + threadIdx.x] = s_rs[threadIdx.x] + s_rs[CI + threadIdx.x];
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

void launch_contrastive_fwd_tc(const ContrastiveFwdArgs& a, cudaStream_t s) {
  static bool once = (cudaFuncSetAttribute(contrastive_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           ConTcLayout::total), true);
  (void)once;
  dim3 grid((a.B + CI - 1) / CI, a.jsplit);
  contrastive_fwd_tc_kernel<<<grid, kThreads, ConTcLayout::total, s>>>(a);
}

}  // namespace scgib
