// gin_tc.cu - GIN layer forward on the 5th-generation tensor cores (tcgen05 + tensor memory).
//
// Same contract as gin_fwd_kernel (gin_kernels.cu; reference models.py:66-72) - the two 64-wide GEMMs of the GIN MLP
// run as 3xTF32 tcgen05.mma (fp32 parity, see umma.cuh) with the accumulators in TMEM instead of FP32 FFMA register
// tiles.  Per 64-row tile:
//   gather-aggregate (CSR, float4 lanes, BN+ReLU of the previous layer on load) -> hi/lo split into a K-major smem tile
//   -> MMA1 (u = a W1^T) -> TMEM -> +b1, ReLU -> r (saved) + hi/lo split into smem -> MMA2 (y = r W2^T) -> TMEM -> +b2
//   -> y (saved), per-tile (mean, M2) -> last CTA finalises the batch statistics.
// 2 CTAs per SM (105 KB smem, 128 TMEM columns each) so one CTA's gather overlaps the other's MMA / epilogue.
#include "kernels.cuh"
#include "umma.cuh"

namespace scgib {
using namespace umma;

constexpr int TR = 64;                                  // rows per tile (UMMA M = 64)
constexpr int kWCore = 128;                             // dense cores for the weight tiles (written once)
constexpr uint32_t kIdesc = idesc_tf32(TR, HID, false, false);

template <int KIN>
struct FwdTcLayout {
  static constexpr int ACT = tile_bytes(TR, HID);                 // 18432: a tile [64][KIN] or r tile [64][64]
  static constexpr int W1B = tile_bytes(HID, KIN, kWCore);        // [64 n][KIN k]
  static constexpr int W2B = tile_bytes(HID, HID, kWCore);
  static constexpr int YS = TR * (HID + 1) * 4;                   // plain y tile for the statistics (aliases act_lo.. no: own)
  static constexpr int off_act_hi = 0, off_act_lo = ACT, off_w1_hi = 2 * ACT, off_w1_lo = off_w1_hi + W1B,
                       off_w2_hi = off_w1_lo + W1B, off_w2_lo = off_w2_hi + W2B, off_f = off_w2_lo + W2B;
  // float scratch: b1[64] b2[64] mean[64] red[4][64] | doubles dred[4*64] | bar | tmem slot
  static constexpr int off_dred = off_f + (3 * 64 + 4 * 64) * 4;
  static constexpr int off_bar = off_dred + 4 * 64 * 8;
  static constexpr int total = off_bar + 16;
  static_assert(YS <= 2 * ACT, "y tile must fit in the activation tiles");
};

template <int KIN>
__global__ void __launch_bounds__(kThreads, 2)
gin_fwd_tc_kernel(GinFwdArgs p) {
  using L = FwdTcLayout<KIN>;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* act_hi = smem + L::off_act_hi;
  unsigned char* act_lo = smem + L::off_act_lo;
  unsigned char* w1_hi = smem + L::off_w1_hi;
  unsigned char* w1_lo = smem + L::off_w1_lo;
  unsigned char* w2_hi = smem + L::off_w2_hi;
  unsigned char* w2_lo = smem + L::off_w2_lo;
  float* s_b1 = reinterpret_cast<float*>(smem + L::off_f);
  float* s_b2 = s_b1 + 64;
  float* s_mean = s_b2 + 64;
  float* s_red = s_mean + 64;                                   // [4][64]
  double* s_dred = reinterpret_cast<double*>(smem + L::off_dred);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::off_bar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L::off_bar + 8);
  float* ys = reinterpret_cast<float*>(act_hi);                 // [64][65] plain tile, valid after MMA2 has completed

  constexpr int LPR = KIN / 4, RPP = kThreads / LPR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) tmem_alloc(s_tmem, 128);
  if (threadIdx.x == 0) mbar_init(s_bar, 1);
  // weights: natural [out][in] = K-major B operand; hi/lo split once per CTA
  for (int i = threadIdx.x; i < HID * (KIN / 4); i += kThreads) {
    const int o = i / (KIN / 4), c4 = i % (KIN / 4);
    store_split4(w1_hi, w1_lo, KIN, o, c4, ldg4(p.W1 + (size_t)o * KIN + c4 * 4), kWCore);
  }
  for (int i = threadIdx.x; i < HID * (HID / 4); i += kThreads) {
    const int o = i / (HID / 4), c4 = i % (HID / 4);
    store_split4(w2_hi, w2_lo, HID, o, c4, ldg4(p.W2 + (size_t)o * HID + c4 * 4), kWCore);
  }
  if (threadIdx.x < HID) { s_b1[threadIdx.x] = p.b1[threadIdx.x]; s_b2[threadIdx.x] = p.b2[threadIdx.x]; }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;
  const uint32_t d1 = tmem, d2 = tmem + 64;
  const Operand opA1{smem_u32(act_hi), smem_u32(act_lo), false, (uint32_t)group_bytes(KIN), kCoreStride};
  const Operand opA2{smem_u32(act_hi), smem_u32(act_lo), false, (uint32_t)group_bytes(HID), kCoreStride};
  const Operand opW1{smem_u32(w1_hi), smem_u32(w1_lo), false, (uint32_t)group_bytes(KIN, kWCore), kWCore};
  const Operand opW2{smem_u32(w2_hi), smem_u32(w2_lo), false, (uint32_t)group_bytes(HID, kWCore), kWCore};

  const int n_tiles = (p.V + TR - 1) / TR;
  const int gl = threadIdx.x % LPR, gr = threadIdx.x / LPR;
  Bn4 bn;
  const bool has_bn = (p.bn_in != nullptr);
  if (has_bn) bn.load(p.bn_in, gl * 4);
  // epilogue mapping (UMMA M = 64): accumulator row i lives in TMEM lane (i/16)*32 + i%16
  const int q = warp & 3, hcol = (warp >> 2) * 32;
  const int erow = 16 * q + (lane & 15);
  const bool eactive = lane < 16;
  const uint32_t tlane = (uint32_t)(32 * q) << 16;
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int base = tile * TR;
    // ---- gather-aggregate into the K-major A tile (hi/lo); all row passes of the thread interleaved
    {
      constexpr int NR = TR / RPP;
      int vv[NR];
      float4 agg[NR];
#pragma unroll
      for (int j = 0; j < NR; ++j) vv[j] = base + gr + j * RPP;
      gather_aggregate<KIN, NR>(p.in, p.row_map, p.indptr, p.indices, p.V, vv, gl, has_bn ? &bn : nullptr, agg);
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        if (p.a_out && vv[j] < p.V) st4(p.a_out + (size_t)vv[j] * KIN + gl * 4, agg[j]);
        store_split4(act_hi, act_lo, KIN, gr + j * RPP, gl, agg[j]);
      }
    }
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (threadIdx.x == 0) {
      gemm_3xtf32(d1, opA1, opW1, KIN / 8, kIdesc, false);
      mma_commit(s_bar);
    }
    mbar_wait(s_bar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---- epilogue 1: r = relu(u + b1) -> global (saved) and the A tile of GEMM2
    {
      float v[32];
      float (&v0)[16] = *reinterpret_cast<float (*)[16]>(v);
      float (&v1)[16] = *reinterpret_cast<float (*)[16]>(v + 16);
      tmem_ld16(tmem + tlane + hcol, v0);          // d1 = columns 0..63
      tmem_ld16(tmem + tlane + hcol + 16, v1);
      if (eactive) {
        const int gv = base + erow;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = hcol + 4 * j;
          const float4 rv = relu4(make_float4(v[4 * j] + s_b1[c], v[4 * j + 1] + s_b1[c + 1], v[4 * j + 2] + s_b1[c + 2],
                                              v[4 * j + 3] + s_b1[c + 3]));
          store_split4(act_hi, act_lo, HID, erow, c >> 2, rv);
          if (p.r_out && gv < p.V) st4(p.r_out + (size_t)gv * HID + c, rv);
        }
      }
    }
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (threadIdx.x == 0) {
      gemm_3xtf32(d2, opA2, opW2, HID / 8, kIdesc, false);
      mma_commit(s_bar);
    }
    mbar_wait(s_bar, phase);
    phase ^= 1;
    fence_after_sync();
    // ---- epilogue 2: y = acc + b2 -> global; plain copy into smem for the column statistics
    {
      float v[32];
      float (&v0)[16] = *reinterpret_cast<float (*)[16]>(v);
      float (&v1)[16] = *reinterpret_cast<float (*)[16]>(v + 16);
      tmem_ld16(d2 + tlane + hcol, v0);
      tmem_ld16(d2 + tlane + hcol + 16, v1);
      if (eactive) {
        const int gv = base + erow;
        const bool valid = gv < p.V;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] += s_b2[hcol + j];
          ys[erow * (HID + 1) + hcol + j] = valid ? v[j] : 0.f;
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st4(p.y_out + (size_t)gv * HID + hcol + 4 * j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        }
      }
    }
    fence_before_sync();
    __syncthreads();
    // ---- per-tile column mean and centred second moment (two passes over the smem copy)
    const int cnt = min(TR, p.V - base);
    const int c = threadIdx.x & 63, sg = threadIdx.x >> 6;      // 4 row segments of 16
    {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) s += ys[(sg * 16 + r) * (HID + 1) + c];
      s_red[sg * 64 + c] = s;
    }
    __syncthreads();
    if (threadIdx.x < 64) s_mean[c] = (s_red[c] + s_red[64 + c] + s_red[128 + c] + s_red[192 + c]) / (float)cnt;
    __syncthreads();
    {
      const float mu = s_mean[c];
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int row = sg * 16 + r;
        const float d = ys[row * (HID + 1) + c] - mu;
        s = (row < cnt) ? fmaf(d, d, s) : s;
      }
      s_red[sg * 64 + c] = s;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      p.part[(size_t)tile * 2 * HID + c] = s_mean[c];
      p.part[(size_t)tile * 2 * HID + HID + c] = s_red[c] + s_red[64 + c] + s_red[128 + c] + s_red[192 + c];
    }
    fence_after_sync();   // next tile's generic smem writes / MMAs are ordered after this tile's TMEM reads
  }
  // ---- batch statistics: last CTA combines the per-tile (n, mean, M2) in fp64, fixed order
  const bool last = last_cta_arrives(p.counter);
  if (last) {
    const int c = threadIdx.x & (HID - 1), seg = threadIdx.x >> 6;
    double s = 0.0;
    for (int t = seg; t < n_tiles; t += 4) {
      const double n = (double)min(TR, p.V - t * TR);
      s += n * (double)__ldcg(p.part + (size_t)t * 2 * HID + c);
    }
    s_dred[seg * HID + c] = s;
    __syncthreads();
    const double mean = (s_dred[c] + s_dred[HID + c] + s_dred[2 * HID + c] + s_dred[3 * HID + c]) / (double)p.V;
    __syncthreads();
    double qq = 0.0;
    for (int t = seg; t < n_tiles; t += 4) {
      const double n = (double)min(TR, p.V - t * TR);
      const double d = (double)__ldcg(p.part + (size_t)t * 2 * HID + c) - mean;
      qq += (double)__ldcg(p.part + (size_t)t * 2 * HID + HID + c) + n * d * d;
    }
    s_dred[seg * HID + c] = qq;
    __syncthreads();
    if (threadIdx.x < HID) {
      const double var = (s_dred[c] + s_dred[HID + c] + s_dred[2 * HID + c] + s_dred[3 * HID + c]) / (double)p.V;
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
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int gin_fwd_tc_tiles(int V) { return (V + TR - 1) / TR; }

void launch_gin_fwd_tc(const GinFwdArgs& a, int kin, cudaStream_t s) {
  const int grid = min(gin_fwd_tc_tiles(a.V), 2 * num_sms());
  if (kin == DTR) {
    static bool once = (cudaFuncSetAttribute(gin_fwd_tc_kernel<DTR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             FwdTcLayout<DTR>::total), true);
    (void)once;
    gin_fwd_tc_kernel<DTR><<<grid, kThreads, FwdTcLayout<DTR>::total, s>>>(a);
  } else {
    static bool once = (cudaFuncSetAttribute(gin_fwd_tc_kernel<HID>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             FwdTcLayout<HID>::total), true);
    (void)once;
    gin_fwd_tc_kernel<HID><<<grid, kThreads, FwdTcLayout<HID>::total, s>>>(a);
  }
}

}  // namespace scgib
