// loss_kernels.cu - adjacency-reconstruction loss (Gram identity), the InfoNCE-style contrastive loss
// (flash-style: the B x B similarity matrices are never materialised), loss finalisation, the deterministic
// reduction of per-CTA parameter-gradient partials and the fused Adam step.
//
// Reference call sites replaced (paths relative to the reference tree):
//   loss_recon_adj (dense N x N)        models.py:762-768    ||ZZ^T - A||_F^2 / N  ==  (||Z^T Z||_F^2 - 2 sum_E z_i.z_j + |E|) / N
//   sim / batched_semi_loss             models.py:606-629
//   KL_Loss = mean(KL_tensor)           models.py:679
//   loss = KL + recon + contrastive     exp_pretraining.py:321
//   torch.optim.Adam(lr, wd=5e-5)       exp_pretraining.py:86, 112, 323
#include "kernels.cuh"
#include "side_jobs.cuh"

namespace scgib {

constexpr int GT = kReconTile;
constexpr int CT = 64;   // contrastive tile

// ------------------------------------------------------------------------------------------------
// recon forward: per-CTA partial Gram matrix + partial edge-dot sum
// ------------------------------------------------------------------------------------------------
template <int HID> struct ReconFwdSmem { float z[GT * (HID + 4)]; float red[kThreads / 32]; };

template <int HID>
__global__ void __launch_bounds__(kThreads, HID == 64 ? 2 : 1)
recon_fwd_kernel(ReconFwdArgs p) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ReconFwdSmem<HID>& sm = *reinterpret_cast<ReconFwdSmem<HID>*>(smem_raw);
  constexpr int GLD = HID + 4, LPR = HID / 4, RPP = kThreads / LPR;
  using T = TNMap<HID, HID>;
  float G[T::TO][T::TJ];
#pragma unroll
  for (int i = 0; i < T::TO; ++i)
#pragma unroll
    for (int j = 0; j < T::TJ; ++j) G[i][j] = 0.f;
  float ed = 0.f;
  const int l = threadIdx.x % LPR, hw = threadIdx.x / LPR;
  const int n_tiles = (p.N + GT - 1) / GT;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int base = tile * GT;
    __syncthreads();
    load_row_tile<GT, HID>(sm.z, GLD, p.Z, base, p.N);
    __syncthreads();
    gemm_tn<HID, HID>(sm.z, GLD, sm.z, GLD, GT, G);
    // per-edge dots sum_v z_v . (A Z)_v: the neighbour sums of NR rows per lane group are gathered together (every stage of
    // the indptr -> indices -> row chain issued for all of them at once; a row-at-a-time loop was the kernel's serial part)
    constexpr int NR = 4;
    for (int rb = hw; rb < GT; rb += RPP * NR) {
      int vv[NR];
      float4 nb[NR];
#pragma unroll
      for (int j = 0; j < NR; ++j) vv[j] = (rb + j * RPP < GT) ? base + rb + j * RPP : p.N;
      gather_aggregate<HID, NR, false>(p.Z, nullptr, p.indptr, p.indices, p.N, vv, l, nullptr, nb);
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        if (vv[j] < p.N) {
          const float4 zv = ld4(sm.z + (rb + j * RPP) * GLD + l * 4);
          ed += zv.x * nb[j].x + zv.y * nb[j].y + zv.z * nb[j].z + zv.w * nb[j].w;
        }
      }
    }
  }
  float* part = p.part + (size_t)blockIdx.x * (HID * HID + 4);
#pragma unroll
  for (int i = 0; i < T::TO; ++i)
#pragma unroll
    for (int j = 0; j < T::TJ; ++j) part[(T::o0() + i) * HID + T::j0() + j] = G[i][j];
  ed = warp_sum(ed);
  if ((threadIdx.x & 31) == 0) sm.red[threadIdx.x >> 5] = ed;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) s += sm.red[w];
    part[HID * HID] = s;
  }
}
template <int H>
static void launch_recon_fwd_t(const ReconFwdArgs& a, int grid, cudaStream_t s) {
  static bool once = (cudaFuncSetAttribute(recon_fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ReconFwdSmem<H>)), true);
  (void)once;
  launch_k((recon_fwd_kernel<H>), dim3(grid), dim3(kThreads), sizeof(ReconFwdSmem<H>), s, a);
}
void launch_recon_fwd(const ReconFwdArgs& a, int hidden, int grid, cudaStream_t s) {
  if (hidden == 64) launch_recon_fwd_t<64>(a, grid, s); else launch_recon_fwd_t<128>(a, grid, s);
}

__global__ void __launch_bounds__(kThreads)
recon_reduce_kernel(const float* __restrict__ part, int grid, float* __restrict__ G, float* __restrict__ edge_sum, int HID) {
  pdl_sync();
  recon_reduce_body(part, grid, G, edge_sum, HID, (int)blockIdx.x, (int)gridDim.x);
}
void launch_recon_reduce(const float* part, int grid, float* G, float* edge_sum, int hidden, cudaStream_t s) {
  launch_k((recon_reduce_kernel), dim3((hidden * hidden + 1 + kThreads - 1) / kThreads), dim3(kThreads), 0, s, part, grid, G, edge_sum, hidden);
}

// recon backward: gZ = scale * (4/N) * (Z G - A Z)
template <int HID>
__global__ void __launch_bounds__(kThreads, HID == 64 ? 2 : 1)
recon_bwd_kernel(ReconBwdArgs p) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  recon_bwd_body<HID>(p, smem_raw, (int)blockIdx.x, (int)gridDim.x);
}
template <int H>
static void launch_recon_bwd_t(const ReconBwdArgs& a, cudaStream_t s) {
  static bool once = (cudaFuncSetAttribute(recon_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ReconBwdSmem<H>)), true);
  (void)once;
  const int grid = min((a.N + GT - 1) / GT, (H == 64 ? 2 : 1) * num_sms());
  launch_k((recon_bwd_kernel<H>), dim3(grid), dim3(kThreads), sizeof(ReconBwdSmem<H>), s, a);
}
void launch_recon_bwd(const ReconBwdArgs& a, int hidden, cudaStream_t s) {
  if (hidden == 64) launch_recon_bwd_t<64>(a, s); else launch_recon_bwd_t<128>(a, s);
}

// ------------------------------------------------------------------------------------------------
// contrastive loss
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }

__device__ __forceinline__ float tf32_round(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// z1 = core / max(||core||, 1e-12), z2 = readout / max(||readout||, 1e-12), diag = z1 . z2   (warp per row)
// (HID = 128: the same kernel run on channel pairs 2*lane and 64 + 2*lane; norms and the diagonal sum both halves)
template <int HID>
__global__ void __launch_bounds__(kThreads)
normalize_kernel(NormalizeArgs p) {
  pdl_sync();
  const int lane = threadIdx.x & 31, c = 2 * lane;
  if (HID == 128) {
    for (int i = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); i < p.B; i += gridDim.x * (kThreads / 32)) {
      const float2 a0 = ld2(p.core + (size_t)i * HID + c), a1 = ld2(p.core + (size_t)i * HID + 64 + c);
      const float2 b0 = ld2(p.readout + (size_t)i * HID + c), b1 = ld2(p.readout + (size_t)i * HID + 64 + c);
      const float na = fmaxf(sqrtf(warp_sum((a0.x * a0.x + a0.y * a0.y) + (a1.x * a1.x + a1.y * a1.y))), 1e-12f);
      const float nb = fmaxf(sqrtf(warp_sum((b0.x * b0.x + b0.y * b0.y) + (b1.x * b1.x + b1.y * b1.y))), 1e-12f);
      const float2 za0 = make_float2(a0.x / na, a0.y / na), za1 = make_float2(a1.x / na, a1.y / na);
      const float2 zb0 = make_float2(b0.x / nb, b0.y / nb), zb1 = make_float2(b1.x / nb, b1.y / nb);
      st2(p.z1 + (size_t)i * HID + c, za0); st2(p.z1 + (size_t)i * HID + 64 + c, za1);
      st2(p.z2 + (size_t)i * HID + c, zb0); st2(p.z2 + (size_t)i * HID + 64 + c, zb1);
      const float d = warp_sum((za0.x * zb0.x + za0.y * zb0.y) + (za1.x * zb1.x + za1.y * zb1.y));
      if (lane == 0) { p.n1[i] = na; p.n2[i] = nb; p.diag[i] = d; }
    }
    return;
  }
  for (int i = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); i < p.B; i += gridDim.x * (kThreads / 32)) {
    const float2 a = ld2(p.core + (size_t)i * HID + c), b = ld2(p.readout + (size_t)i * HID + c);
    const float na = fmaxf(sqrtf(warp_sum(a.x * a.x + a.y * a.y)), 1e-12f);
    const float nb = fmaxf(sqrtf(warp_sum(b.x * b.x + b.y * b.y)), 1e-12f);
    const float2 za = make_float2(a.x / na, a.y / na), zb = make_float2(b.x / nb, b.y / nb);
    st2(p.z1 + (size_t)i * HID + c, za);
    st2(p.z2 + (size_t)i * HID + c, zb);
    if (p.zsplit) {   // fp16 hi / lo parts [4][B][HID] halves (z1 hi, z1 lo, z2 hi, z2 lo) for the tensor-core kernels
      uint32_t* zs = reinterpret_cast<uint32_t*>(p.zsplit);
      const size_t n = (size_t)p.B * HID / 2, o = ((size_t)i * HID + c) / 2;
      uint32_t ah, al, bh, bl;
      split_f16x2_plain(za.x, za.y, ah, al);
      split_f16x2_plain(zb.x, zb.y, bh, bl);
      zs[o] = ah; zs[n + o] = al; zs[2 * n + o] = bh; zs[3 * n + o] = bl;
    }
    const float d = warp_sum(za.x * zb.x + za.y * zb.y);
    if (lane == 0) { p.n1[i] = na; p.n2[i] = nb; p.diag[i] = d; }
  }
}
void launch_normalize(const NormalizeArgs& a, int hidden, cudaStream_t s) {
  const int grid = min((a.B + 7) / 8, 8 * num_sms());
  if (hidden == 64) launch_k((normalize_kernel<64>), dim3(grid), dim3(kThreads), 0, s, a);
  else launch_k((normalize_kernel<128>), dim3(grid), dim3(kThreads), 0, s, a);
}

int contrastive_jsplit(int B) {
  // The tcgen05 kernels (contrastive_tc.cu) run one CTA per SM on 128-row blocks: split the column range so that
  // (row blocks) x (splits) fills the SMs in a single wave.  The FFMA kernels (64-row blocks, 2 CTAs per SM) accept the same value.
  const int iblocks = (B + 127) / 128;
  int js = (num_sms() - kConSideCtas) / iblocks;     // kConSideCtas SMs stay free for the side jobs of those launches
  const int jblocks = (B + CT - 1) / CT;
  if (js > jblocks) js = jblocks;
  if (js < 1) js = 1;
  return js;
}

// S[i][j] = a_i . b_j for a 64x64 tile pair; thread (ti,tj): rows ti*4+ii, columns tj+16*jj (conflict-free float4 reads)
template <int HID>
__device__ __forceinline__ void sim_tile(const float* __restrict__ As, const float* __restrict__ Bs, float (&s)[4][4]) {
  constexpr int GLD = HID + 4;
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 2
  for (int k = 0; k < HID; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = ld4(As + (ti * 4 + i) * GLD + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = ld4(Bs + (tj + 16 * j) * GLD + k);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        s[i][j] = fmaf(a[i].x, b[j].x, fmaf(a[i].y, b[j].y, fmaf(a[i].z, b[j].z, fmaf(a[i].w, b[j].w, s[i][j]))));
  }
}

template <int HID> struct ConFwdSmem { float zi[CT * (HID + 4)]; float zj1[CT * (HID + 4)]; float zj2[CT * (HID + 4)]; };

template <int HID>
__global__ void __launch_bounds__(kThreads, 2)
contrastive_fwd_kernel(ContrastiveFwdArgs p) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ConFwdSmem<HID>& sm = *reinterpret_cast<ConFwdSmem<HID>*>(smem_raw);
  constexpr int GLD = HID + 4;
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const int ibase = blockIdx.x * CT;
  const int jblocks = (p.B + CT - 1) / CT;
  load_row_tile<CT, HID>(sm.zi, GLD, p.z1, ibase, p.B);
  float rs[4] = {0.f, 0.f, 0.f, 0.f};
  for (int jb = blockIdx.y; jb < jblocks; jb += gridDim.y) {
    const int jbase = jb * CT;
    __syncthreads();
    load_row_tile<CT, HID>(sm.zj1, GLD, p.z1, jbase, p.B);
    load_row_tile<CT, HID>(sm.zj2, GLD, p.z2, jbase, p.B);
    __syncthreads();
    float s[4][4];
    sim_tile<HID>(sm.zi, sm.zj1, s);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gi = ibase + ti * 4 + i, gj = jbase + tj + 16 * j;
        if (gj < p.B && gj != gi) rs[i] += expf(s[i][j]);       // refl_sim.sum(1) - refl_sim.diag()
      }
    sim_tile<HID>(sm.zi, sm.zj2, s);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gj = jbase + tj + 16 * j;
        if (gj < p.B) rs[i] += expf(s[i][j]);                    // between_sim.sum(1)
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = rs[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int gi = ibase + ti * 4 + i;
    if (tj == 0 && gi < p.B) p.rowsum[(size_t)blockIdx.y * p.B + gi] = v;
  }
}
template <int H>
static void launch_contrastive_fwd_t(const ContrastiveFwdArgs& a, cudaStream_t s) {
  static bool once = (cudaFuncSetAttribute(contrastive_fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConFwdSmem<H>)), true);
  (void)once;
  dim3 grid((a.B + CT - 1) / CT, a.jsplit);
  launch_k((contrastive_fwd_kernel<H>), dim3(grid), dim3(kThreads), sizeof(ConFwdSmem<H>), s, a);
}
void launch_contrastive_fwd(const ContrastiveFwdArgs& a, int hidden, cudaStream_t s) {
  if (hidden == 64) launch_contrastive_fwd_t<64>(a, s); else launch_contrastive_fwd_t<128>(a, s);
}

// backward.  blockIdx.z == 0: rows of z1 (g1);  blockIdx.z == 1: rows of z2 (g2).
//   g1_i = sum_{j!=i} e^{s_r(i,j)} (1/D_i + 1/D_j) z1_j + sum_j e^{s_b(i,j)}/D_i z2_j
//   g2_j = sum_i e^{s_b(i,j)}/D_i z1_i
// (the 1/B factor and the -z2_i / -z1_j terms are applied in the finalise kernel)
template <int HID> struct ConBwdSmem { float zi[CT * (HID + 4)]; float zj1[CT * (HID + 4)]; float zj2[CT * (HID + 4)]; float P[CT * (CT + 4)]; float Di[CT]; float Dj[CT]; };

template <int HID>
__global__ void __launch_bounds__(kThreads, HID == 64 ? 2 : 1)
contrastive_bwd_kernel(ContrastiveBwdArgs p) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ConBwdSmem<HID>& sm = *reinterpret_cast<ConBwdSmem<HID>*>(smem_raw);
  constexpr int GLD = HID + 4, PLD = CT + 4;
  using M = NNMap<CT, HID>;
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const bool mode1 = (blockIdx.z == 1);
  const int ibase = blockIdx.x * CT;
  const int jblocks = (p.B + CT - 1) / CT;
  load_row_tile<CT, HID>(sm.zi, GLD, mode1 ? p.z2 : p.z1, ibase, p.B);
  if (threadIdx.x < CT) sm.Di[threadIdx.x] = (ibase + threadIdx.x < p.B) ? 1.f / p.D[ibase + threadIdx.x] : 0.f;
  float acc[M::TM][4];
#pragma unroll
  for (int m = 0; m < M::TM; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
  for (int jb = blockIdx.y; jb < jblocks; jb += gridDim.y) {
    const int jbase = jb * CT;
    __syncthreads();
    load_row_tile<CT, HID>(sm.zj1, GLD, p.z1, jbase, p.B);
    if (!mode1) load_row_tile<CT, HID>(sm.zj2, GLD, p.z2, jbase, p.B);
    if (threadIdx.x < CT) sm.Dj[threadIdx.x] = (jbase + threadIdx.x < p.B) ? 1.f / p.D[jbase + threadIdx.x] : 0.f;
    __syncthreads();
    float s[4][4];
    sim_tile<HID>(sm.zi, sm.zj1, s);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int li = ti * 4 + i, lj = tj + 16 * j;
        const int gi = ibase + li, gj = jbase + lj;
        float w;
        if (mode1) w = sm.Dj[lj];                                   // rows are z2_j', columns z1_i: weight 1/D_i of the column
        else w = (gj != gi) ? sm.Di[li] + sm.Dj[lj] : 0.f;
        sm.P[li * PLD + lj] = (gj < p.B && gi < p.B) ? expf(s[i][j]) * w : 0.f;
      }
    __syncthreads();
    gemm_nn<CT, CT, HID>(sm.P, PLD, sm.zj1, GLD, acc);
    if (!mode1) {
      sim_tile<HID>(sm.zi, sm.zj2, s);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int li = ti * 4 + i, lj = tj + 16 * j;
          sm.P[li * PLD + lj] = (jbase + lj < p.B && ibase + li < p.B) ? expf(s[i][j]) * sm.Di[li] : 0.f;
        }
      __syncthreads();
      gemm_nn<CT, CT, HID>(sm.P, PLD, sm.zj2, GLD, acc);
    }
  }
  float* out = (mode1 ? p.g2p : p.g1p) + (size_t)blockIdx.y * p.B * HID;
  const int c0 = M::col0(), r0 = M::row0();
#pragma unroll
  for (int m = 0; m < M::TM; ++m) {
    const int gi = ibase + r0 + m;
    if (gi < p.B) st4(out + (size_t)gi * HID + c0, make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]));
  }
}
template <int H>
static void launch_contrastive_bwd_t(const ContrastiveBwdArgs& a, cudaStream_t s) {
  static bool once = (cudaFuncSetAttribute(contrastive_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ConBwdSmem<H>)), true);
  (void)once;
  dim3 grid((a.B + CT - 1) / CT, a.jsplit, 2);
  launch_k((contrastive_bwd_kernel<H>), dim3(grid), dim3(kThreads), sizeof(ConBwdSmem<H>), s, a);
}
void launch_contrastive_bwd(const ContrastiveBwdArgs& a, int hidden, cudaStream_t s) {
  if (hidden == 64) launch_contrastive_bwd_t<64>(a, s); else launch_contrastive_bwd_t<128>(a, s);
}

// g wrt the un-normalised readouts: (g - z_hat (z_hat . g)) / max(||z||, 1e-12)     (warp per row)
template <int HID>
__global__ void __launch_bounds__(kThreads)
contrastive_bwd_finalize_kernel(ContrastiveBwdFinArgs p) {
  pdl_sync();
  constexpr int NH = HID / 64;                  // a lane owns channel pairs 2*lane + 64*h
  const int lane = threadIdx.x & 31, c = 2 * lane;
  const float k = p.scale / (float)p.B;
  for (int i = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); i < p.B; i += gridDim.x * (kThreads / 32)) {
    float2 g1[NH], g2[NH], z1[NH], z2[NH];
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      g1[h] = make_float2(0.f, 0.f); g2[h] = make_float2(0.f, 0.f);
      for (int js = 0; js < p.jsplit; ++js) {
        const float2 a = ld2(p.g1p + ((size_t)js * p.B + i) * HID + 64 * h + c), b = ld2(p.g2p + ((size_t)js * p.B + i) * HID + 64 * h + c);
        g1[h].x += a.x; g1[h].y += a.y; g2[h].x += b.x; g2[h].y += b.y;
      }
      z1[h] = ld2(p.z1 + (size_t)i * HID + 64 * h + c); z2[h] = ld2(p.z2 + (size_t)i * HID + 64 * h + c);
      g1[h].x = k * (g1[h].x - z2[h].x); g1[h].y = k * (g1[h].y - z2[h].y);
      g2[h].x = k * (g2[h].x - z1[h].x); g2[h].y = k * (g2[h].y - z1[h].y);
      const float p1 = z1[h].x * g1[h].x + z1[h].y * g1[h].y, p2 = z2[h].x * g2[h].x + z2[h].y * g2[h].y;
      t1 = h == 0 ? p1 : t1 + p1; t2 = h == 0 ? p2 : t2 + p2;
    }
    const float d1 = warp_sum(t1), d2 = warp_sum(t2);
    const float n1 = p.n1[i], n2 = p.n2[i];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      st2(p.g_core + (size_t)i * HID + 64 * h + c, make_float2((g1[h].x - z1[h].x * d1) / n1, (g1[h].y - z1[h].y * d1) / n1));
      st2(p.g_readout + (size_t)i * HID + 64 * h + c, make_float2((g2[h].x - z2[h].x * d2) / n2, (g2[h].y - z2[h].y * d2) / n2));
    }
  }
}
void launch_contrastive_bwd_finalize(const ContrastiveBwdFinArgs& a, int hidden, cudaStream_t s) {
  const int grid = min((a.B + 7) / 8, 8 * num_sms());
  if (hidden == 64) launch_k((contrastive_bwd_finalize_kernel<64>), dim3(grid), dim3(kThreads), 0, s, a);
  else launch_k((contrastive_bwd_finalize_kernel<128>), dim3(grid), dim3(kThreads), 0, s, a);
}

// ------------------------------------------------------------------------------------------------
// losses = {KL, contrastive, recon, total}; also the contrastive denominators D_i (saved for backward)
// ------------------------------------------------------------------------------------------------
constexpr int kFin = 1024;     // one CTA; latency-bound (B rows x jsplit dependent loads each): as many threads as a CTA allows
__global__ void __launch_bounds__(kFin)
loss_finalize_kernel(LossFinalizeArgs p) {
  pdl_sync();
  loss_finalize_body<kFin>(p);
}
void launch_loss_finalize(const LossFinalizeArgs& a, cudaStream_t s) { launch_k((loss_finalize_kernel), dim3(1), dim3(kFin), 0, s, a); }

// ------------------------------------------------------------------------------------------------
// grads[off+i] = sum_c part[c*pstride + off + i]   (fixed order => run-to-run bit-stable)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
reduce_partials_kernel(const float* __restrict__ part, int64_t pstride, int nparts, ReduceRanges r, float* __restrict__ grads) {
  pdl_sync();
  // 4 consecutive lanes share one output element: each sums a contiguous quarter of the partial rows in order, the
  // quarters are combined as (q0 + q1) + (q2 + q3) - a fixed order, with 4x the loads in flight
  const int64_t off = r.off[blockIdx.y], len = r.len[blockIdx.y];
  const int c0 = r.c0[blockIdx.y], c1 = min(r.c1[blockIdx.y], nparts);
  const int sub = threadIdx.x & 3;
  const int chunk = (c1 - c0 + 3) / 4;
  const int a = min(c0 + sub * chunk, c1), b = min(a + chunk, c1);
  const int64_t len_pad = (len + 63) / 64 * 64;          // whole warps stay in the loop for the shuffles
  for (int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 2; i < len_pad; i += ((int64_t)gridDim.x * kThreads) >> 2) {
    double s = 0.0;
    if (i < len) {
#pragma unroll 8
      for (int c = a; c < b; ++c) s += (double)part[(size_t)c * pstride + off + i];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (sub == 0 && i < len) {
      const int ilv = r.ilv[blockIdx.y];
      grads[ilv > 0 ? r.dst[blockIdx.y] + (i / ilv) * 2 * ilv + i % ilv : off + i] = (float)s;
    }
  }
}
void launch_reduce_partials(const float* part, int64_t pstride, int nparts, const ReduceRanges& r, float* grads,
                            cudaStream_t s) {
  if (r.n == 0) return;
  dim3 grid(128, r.n);  // largest range is 8192 floats x 4 lanes = 128 CTAs x 256
  launch_k((reduce_partials_kernel), dim3(grid), dim3(kThreads), 0, s, part, pstride, nparts, r, grads);
}

// ------------------------------------------------------------------------------------------------
// Adam (L2-in-gradient weight decay), same update order as torch.optim.Adam's single-tensor path
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float lr, float b1, float b2, float eps, float wd, float gscale, float step_size, float bc2_sqrt) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i] * gscale);
    const float mi = fmaf(1.f - b1, gi - m[i], m[i]);          // lerp(m, g, 1-b1)
    const float vi = fmaf(1.f - b2, gi * gi, b2 * v[i]);
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}
void launch_adam(float* params, const float* grads, float* m, float* v, int64_t n, int64_t step, float lr, float b1,
                 float b2, float eps, float wd, float gscale, cudaStream_t s) {
  const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
  int64_t gb = (n + kThreads - 1) / kThreads;
  const int grid = (int)(gb < 4 * (int64_t)num_sms() ? gb : 4 * (int64_t)num_sms());
  adam_kernel<<<grid, kThreads, 0, s>>>(params, grads, m, v, n, lr, b1, b2, eps, wd, gscale, (float)(lr / bc1),
                                        (float)sqrt(bc2));
}

}  // namespace scgib
