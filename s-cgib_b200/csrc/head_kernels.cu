// head_kernels.cu - everything between the two GIN encoders and the losses:
//   ego sum-pooling + attention logit, compressor linear, the per-graph information-bottleneck gate with
//   the core/graph readouts, KL and the core-candidate softmax fused in one segment kernel, and the head MLP.
//
// Reference call sites replaced (paths relative to the reference tree):
//   dgl.sum_nodes                         models.py:716, 725, 733, 684
//   compress / compression (Python loop)  models.py:595-604, 631-660
//   attention (Python loop)               models.py:738-749
//   self.MLP(interaction_map)             models.py:569-572, 676
#include "kernels.cuh"

namespace scgib {

constexpr int GT = 128;
constexpr int GLD = HID + 4;

// ------------------------------------------------------------------------------------------------
// segment sums (half-warp per segment, float4 lanes)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
segment_sum_kernel(const float* __restrict__ in, const int32_t* __restrict__ seg_ptr, int S, const float* __restrict__ bn,
                   float* __restrict__ out) {
  const int l = threadIdx.x & 15;
  Bn4 b;
  if (bn) b.load(bn, l * 4);
  for (int s = blockIdx.x * 16 + (threadIdx.x >> 4); s < S; s += gridDim.x * 16) {
    const int r0 = __ldg(seg_ptr + s), r1 = __ldg(seg_ptr + s + 1);
    float4 acc = make4(0.f);
    for (int r = r0; r < r1; ++r) {
      const float4 y = ld4(in + (size_t)r * HID + l * 4);
      acc = add4(acc, bn ? b.act(y) : y);
    }
    st4(out + (size_t)s * HID + l * 4, acc);
  }
}
void launch_segment_sum(const float* in, const int32_t* seg_ptr, int S, const float* bn, float* out, cudaStream_t s) {
  const int grid = min((S + 15) / 16, 16 * num_sms());
  segment_sum_kernel<<<grid, kThreads, 0, s>>>(in, seg_ptr, S, bn, out);
}

// C_v = sum over the ego-net of v of relu(BN(y)) ; logit_v = w_cand . C_v
template <bool BF>
__global__ void __launch_bounds__(kThreads)
ego_pool_fwd_kernel(EgoPoolFwdArgs p) {
  const int l = threadIdx.x & 15;
  Bn4 b;
  b.load(p.bn, l * 4);
  const float4 w = ldg4(p.w_cand + l * 4);
  for (int v = blockIdx.x * 16 + (threadIdx.x >> 4); v < p.N; v += gridDim.x * 16) {
    const int r0 = __ldg(p.ego_ptr + v), r1 = __ldg(p.ego_ptr + v + 1);
    float4 acc = make4(0.f);
#pragma unroll 4
    for (int r = r0; r < r1; ++r) acc = add4(acc, b.act(ld4a<BF>(p.y, (size_t)r * HID + l * 4)));
    st4(p.C + (size_t)v * HID + l * 4, acc);
    float d = acc.x * w.x + acc.y * w.y + acc.z * w.z + acc.w * w.w;
    const unsigned hmask = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;   // the two half-warps may diverge at the tail
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(hmask, d, o);
    if (l == 0) p.logit[v] = d;
  }
}
void launch_ego_pool_fwd(const EgoPoolFwdArgs& a, cudaStream_t s) {
  const int grid = min((a.N + 15) / 16, 16 * num_sms());
  if (a.y_bf16) ego_pool_fwd_kernel<true><<<grid, kThreads, 0, s>>>(a);
  else ego_pool_fwd_kernel<false><<<grid, kThreads, 0, s>>>(a);
}

// ------------------------------------------------------------------------------------------------
// H = relu(BN(y_last)) ; q = H Wc1^T + bc1
// ------------------------------------------------------------------------------------------------
struct GateLinFwdSmem { float tile[GT * GLD]; float w[HID * HID]; };

template <bool BF>
__global__ void __launch_bounds__(kThreads, 2)
gate_lin_fwd_kernel(GateLinFwdArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GateLinFwdSmem& sm = *reinterpret_cast<GateLinFwdSmem*>(smem_raw);
  using M = NNMap<GT, HID>;
  load_matrix<HID>(sm.w, HID, p.Wc1t, HID);
  const int n_tiles = (p.N + GT - 1) / GT;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int base = tile * GT;
    __syncthreads();
    for (int i = threadIdx.x; i < GT * (HID / 4); i += kThreads) {
      const int r = i / (HID / 4), c = (i % (HID / 4)) * 4;
      const int v = base + r;
      float4 h = make4(0.f);
      if (v < p.N) {
        Bn4 b;
        b.load(p.bn, c);
        h = b.act(ld4a<BF>(p.y, (size_t)v * HID + c));
        st4(p.H + (size_t)v * HID + c, h);
      }
      st4(sm.tile + r * GLD + c, h);
    }
    __syncthreads();
    float acc[M::TM][4];
    const int c0 = M::col0(), r0 = M::row0();
    const float4 bias = ldg4(p.bc1 + c0);
#pragma unroll
    for (int m = 0; m < M::TM; ++m) { acc[m][0] = bias.x; acc[m][1] = bias.y; acc[m][2] = bias.z; acc[m][3] = bias.w; }
    gemm_nn<GT, HID, HID>(sm.tile, GLD, sm.w, HID, acc);
#pragma unroll
    for (int m = 0; m < M::TM; ++m) {
      const int v = base + r0 + m;
      if (v < p.N) st4(p.q + (size_t)v * HID + c0, make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]));
    }
  }
}
void launch_gate_lin_fwd(const GateLinFwdArgs& a, cudaStream_t s) {
  static bool once = (cudaFuncSetAttribute(gate_lin_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GateLinFwdSmem)),
                      cudaFuncSetAttribute(gate_lin_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GateLinFwdSmem)), true);
  (void)once;
  const int grid = min((a.N + GT - 1) / GT, 2 * num_sms());
  if (a.y_bf16) gate_lin_fwd_kernel<true><<<grid, kThreads, sizeof(GateLinFwdSmem), s>>>(a);
  else gate_lin_fwd_kernel<false><<<grid, kThreads, sizeof(GateLinFwdSmem), s>>>(a);
}

// gH += g_q Wc1 ; dWc1 += g_q^T H ; dbc1 += sum g_q
struct GateLinBwdSmem { float gq[2][GT * GLD]; float h[2][GT * GLD]; float w[HID * HID]; };   // two tile buffers: cp.async prefetch

__global__ void __launch_bounds__(kThreads, 1)
gate_lin_bwd_kernel(GateLinBwdArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GateLinBwdSmem& sm = *reinterpret_cast<GateLinBwdSmem*>(smem_raw);
  using M = NNMap<GT, HID>;
  using T = TNMap<HID, HID>;
  load_matrix<HID>(sm.w, HID, p.Wc1, HID);
  float dW[T::TO][T::TJ];
#pragma unroll
  for (int i = 0; i < T::TO; ++i)
#pragma unroll
    for (int j = 0; j < T::TJ; ++j) dW[i][j] = 0.f;
  float dbias = 0.f;
  const int n_tiles = (p.N + GT - 1) / GT;
  // the next tile's g_q / H rows are copied (cp.async, no register staging) while this tile's GEMMs run
  if ((int)blockIdx.x < n_tiles) {
    cp_async_row_tile<GT, HID>(sm.gq[0], GLD, p.g_q, blockIdx.x * GT, p.N);
    cp_async_row_tile<GT, HID>(sm.h[0], GLD, p.H, blockIdx.x * GT, p.N);
  }
  cp_async_commit();
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int base = tile * GT;
    const float* gq = sm.gq[it & 1];
    const float* hh = sm.h[it & 1];
    cp_async_wait_all();
    __syncthreads();                      // tile `it` has landed; every thread is done with the other buffer (tile it-1)
    if (tile + (int)gridDim.x < n_tiles) {
      cp_async_row_tile<GT, HID>(sm.gq[(it + 1) & 1], GLD, p.g_q, (tile + gridDim.x) * GT, p.N);
      cp_async_row_tile<GT, HID>(sm.h[(it + 1) & 1], GLD, p.H, (tile + gridDim.x) * GT, p.N);
    }
    cp_async_commit();
    float acc[M::TM][4];
#pragma unroll
    for (int m = 0; m < M::TM; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
    gemm_nn<GT, HID, HID>(gq, GLD, sm.w, HID, acc);
    const int c0 = M::col0(), r0 = M::row0();
#pragma unroll
    for (int m = 0; m < M::TM; ++m) {
      const int v = base + r0 + m;
      if (v < p.N) {
        float* dst = p.gH + (size_t)v * HID + c0;
        const float4 old = ld4(dst);
        st4(dst, make_float4(old.x + acc[m][0], old.y + acc[m][1], old.z + acc[m][2], old.w + acc[m][3]));
      }
    }
    gemm_tn<HID, HID>(gq, GLD, hh, GLD, GT, dW);
    if (threadIdx.x < HID) {
      float s = 0.f;
#pragma unroll 8
      for (int r = 0; r < GT; ++r) s += gq[r * GLD + threadIdx.x];
      dbias += s;
    }
  }
  cp_async_wait_all();
  float* part = p.part + (size_t)blockIdx.x * p.pstride;
#pragma unroll
  for (int i = 0; i < T::TO; ++i)
#pragma unroll
    for (int j = 0; j < T::TJ; ++j) part[p.off_W + (T::o0() + i) * HID + T::j0() + j] = dW[i][j];
  if (threadIdx.x < HID) part[p.off_b + threadIdx.x] = dbias;
}
void launch_gate_lin_bwd(const GateLinBwdArgs& a, int grid, cudaStream_t s) {
  static bool once = (cudaFuncSetAttribute(gate_lin_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(GateLinBwdSmem)), true);
  (void)once;
  gate_lin_bwd_kernel<<<grid, kThreads, sizeof(GateLinBwdSmem), s>>>(a);
}

// ------------------------------------------------------------------------------------------------
// per-graph gate (warp per graph, lane owns channels 2l, 2l+1)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }

constexpr float kKlEps = 0.0000001f;   // models.py:632

template <int RB>      // rows in flight per warp in the gate pass (4 for molecule-sized graphs, 8 for ~150-node graphs)
__global__ void __launch_bounds__(kThreads)
graph_gate_fwd_kernel(GraphGateFwdArgs p) {
  const int lane = threadIdx.x & 31;
  const int c = 2 * lane;
  const float2 gam = ld2(p.gamma_c + c), bet = ld2(p.beta_c + c), w2 = ld2(p.wc2 + c);
  const float bc2 = __ldg(p.bc2);
  const int warps = gridDim.x * (kThreads / 32);
  for (int g = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); g < p.B; g += warps) {
    const int v0 = __ldg(p.graph_ptr + g), v1 = __ldg(p.graph_ptr + g + 1);
    const float n = (float)(v1 - v0);
    // pass 1: means
    float2 sH = make_float2(0.f, 0.f), sQ = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int v = v0; v < v1; ++v) {
      const float2 h = ld2(p.H + (size_t)v * HID + c), q = ld2(p.q + (size_t)v * HID + c);
      sH.x += h.x; sH.y += h.y; sQ.x += q.x; sQ.y += q.y;
    }
    const float2 muH = make_float2(sH.x / n, sH.y / n), muQ = make_float2(sQ.x / n, sQ.y / n);
    // pass 2: centred second moments
    float2 vH = make_float2(0.f, 0.f), vQ = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int v = v0; v < v1; ++v) {
      const float2 h = ld2(p.H + (size_t)v * HID + c), q = ld2(p.q + (size_t)v * HID + c);
      float d;
      d = h.x - muH.x; vH.x = fmaf(d, d, vH.x); d = h.y - muH.y; vH.y = fmaf(d, d, vH.y);
      d = q.x - muQ.x; vQ.x = fmaf(d, d, vQ.x); d = q.y - muQ.y; vQ.y = fmaf(d, d, vQ.y);
    }
    const float2 sd = make_float2(sqrtf(vH.x / (n - 1.f)), sqrtf(vH.y / (n - 1.f)));         // torch.std_mean: unbiased
    float2 rstd = make_float2(1.f / sqrtf(vQ.x / n + kBnEps), 1.f / sqrtf(vQ.y / n + kBnEps));  // BN: biased
    float2 muQe = muQ;
    if (p.eval_running) {       // model.eval(): nn.BatchNorm1d normalises with its running statistics
      const float2 rm = ld2(p.eval_running + c), rv = ld2(p.eval_running + HID + c);
      muQe = rm;
      rstd = make_float2(1.f / sqrtf(rv.x + kBnEps), 1.f / sqrtf(rv.y + kBnEps));
    }
    st2(p.readout + (size_t)g * HID + c, sH);
    float* gs = p.gstat + (size_t)g * 4 * HID;
    st2(gs + c, muH); st2(gs + HID + c, sd); st2(gs + 2 * HID + c, muQe); st2(gs + 3 * HID + c, rstd);
    if (p.cstat) {
      st2(p.cstat + (size_t)g * 2 * HID + c, muQ);
      st2(p.cstat + (size_t)g * 2 * HID + HID + c, make_float2(vQ.x / (n - 1.f), vQ.y / (n - 1.f)));
    }
    // pass 3: gate, noisy features, core readout, KL of the last graph
    const bool last = (g == p.B - 1);
    float2 core = make_float2(0.f, 0.f), kl1 = make_float2(0.f, 0.f), kl2 = make_float2(0.f, 0.f);
    const float2 isd = make_float2(1.f / (sd.x + kKlEps), 1.f / (sd.y + kKlEps));
    // latency-bound loop: the loads and shuffle chains of RB rows overlap
    for (int vb = v0; vb < v1; vb += RB) {
      float2 h[RB], q[RB], fu[RB];
      float u[RB], pv[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int v = min(vb + i, v1 - 1);
        h[i] = ld2(p.H + (size_t)v * HID + c); q[i] = ld2(p.q + (size_t)v * HID + c);
        fu[i] = ld2(p.feat_u + (size_t)v * HID + c); u[i] = __ldg(p.gate_u + v);
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const float ox = fmaf((q[i].x - muQe.x) * rstd.x, gam.x, bet.x), oy = fmaf((q[i].y - muQe.y) * rstd.y, gam.y, bet.y);
        pv[i] = fmaxf(ox, 0.f) * w2.x + fmaxf(oy, 0.f) * w2.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < RB; ++i) pv[i] += __shfl_xor_sync(0xffffffffu, pv[i], o);
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int v = vb + i;
        if (v >= v1) break;
        const float eps = __fadd_rn(__fmul_rn(-0.9998f, u[i]), 0.9999f);       // (bias-(1-bias))*u + (1-bias), bias=1e-4
        const float gi = logf(eps) - logf(1.f - eps);
        const float lam = 1.f / (1.f + expf(-(gi + (pv[i] + bc2))));
        const float ln = 1.f - lam;
        const float2 m = make_float2(lam * h[i].x + ln * muH.x, lam * h[i].y + ln * muH.y);
        const float2 sg = make_float2(ln * sd.x, ln * sd.y);
        const float2 z = make_float2(m.x + fu[i].x * sg.x, m.y + fu[i].y * sg.y);
        st2(p.noisy + (size_t)v * HID + c, z);
        core.x += z.x; core.y += z.y;
        if (lane == 0) p.lam[v] = lam;
        if (last) {
          float t;
          t = sg.x * isd.x; kl1.x = fmaf(0.5f * t, t, kl1.x); t = sg.y * isd.y; kl1.y = fmaf(0.5f * t, t, kl1.y);
          t = (m.x - muH.x) * isd.x; kl2.x = fmaf(t, t, kl2.x); t = (m.y - muH.y) * isd.y; kl2.y = fmaf(t, t, kl2.y);
        }
      }
    }
    st2(p.core + (size_t)g * HID + c, core);
    if (last) {
      const float tot = warp_sum(kl1.x + kl1.y + n * (kl2.x + kl2.y));
      if (lane == 0) p.kl[0] = tot / ((float)HID * n);
    }
    // attention softmax over the graph's nodes (core half and bias cancel: SURVEY F14)
    float mx = -INFINITY;
    for (int v = v0 + lane; v < v1; v += 32) mx = fmaxf(mx, __ldg(p.logit + v));
    mx = warp_max(mx);
    float se = 0.f;
    for (int v = v0 + lane; v < v1; v += 32) se += expf(__ldg(p.logit + v) - mx);
    se = warp_sum(se);
    for (int v = v0 + lane; v < v1; v += 32) p.alpha[v] = expf(__ldg(p.logit + v) - mx) / se;
  }
}
void launch_graph_gate_fwd(const GraphGateFwdArgs& a, cudaStream_t s) {
  const int grid = min((a.B + 7) / 8, 8 * num_sms());
  if (a.N / max(a.B, 1) >= 48) graph_gate_fwd_kernel<8><<<grid, kThreads, 0, s>>>(a);
  else graph_gate_fwd_kernel<4><<<grid, kThreads, 0, s>>>(a);
}

// running stats of the compressor BN after B sequential per-graph updates (closed form, fixed order)
// r_B = 0.9^B r_0 + sum_g 0.1 * 0.9^(B-1-g) stat_g ; graphs older than kEmaWindow contribute < 0.9^768 ~ 1e-35.
constexpr int kEmaWindow = 768;
constexpr int kEmaSeg = 8;
__global__ void __launch_bounds__(2 * HID * kEmaSeg)
compressor_ema_kernel(const float* __restrict__ cstat, int B, float* __restrict__ running) {
  __shared__ double s_part[kEmaSeg][2 * HID];
  const int j = threadIdx.x & (2 * HID - 1);  // 0..63 mean, 64..127 var
  const int seg = threadIdx.x / (2 * HID);
  const int g0 = B > kEmaWindow ? B - kEmaWindow : 0;
  // newest graph first: weight 0.1 * 0.9^k for age k = B-1-g, advanced by a constant factor (two pow() per thread
  // instead of one per term: the fp64 pow dominated this kernel)
  double acc = 0.0;
  double w = 0.1 * pow(0.9, (double)seg);
  const double step = pow(0.9, (double)kEmaSeg);
  for (int g = B - 1 - seg; g >= g0; g -= kEmaSeg) {
    acc += w * (double)cstat[(size_t)g * 2 * HID + j];
    w *= step;
  }
  s_part[seg][j] = acc;
  __syncthreads();
  if (seg == 0) {
    double r = pow(0.9, (double)B) * (double)running[j];
#pragma unroll
    for (int k = 0; k < kEmaSeg; ++k) r += s_part[k][j];
    running[j] = (float)r;
  }
}
void launch_compressor_ema(const float* cstat, int B, float* running, cudaStream_t s) {
  compressor_ema_kernel<<<1, 2 * HID * kEmaSeg, 0, s>>>(cstat, B, running);
}

__global__ void __launch_bounds__(kThreads)
graph_gate_bwd_kernel(GraphGateBwdArgs p) {
  __shared__ __align__(16) float s_red[(kThreads / 32) * 5 * HID];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = 2 * lane;
  const float2 gam = ld2(p.gamma_c + c), bet = ld2(p.beta_c + c), w2 = ld2(p.wc2 + c), wc = ld2(p.w_cand + c);
  float2 a_dg = make_float2(0.f, 0.f), a_db = a_dg, a_dw2 = a_dg, a_dwc = a_dg;
  float a_dbc2 = 0.f;
  const int warps = gridDim.x * (kThreads / 32);
  for (int g = blockIdx.x * (kThreads / 32) + warp; g < p.B; g += warps) {
    const int v0 = __ldg(p.graph_ptr + g), v1 = __ldg(p.graph_ptr + g + 1);
    const float n = (float)(v1 - v0);
    const float* gs = p.gstat + (size_t)g * 4 * HID;
    const float2 muH = ld2(gs + c), sd = ld2(gs + HID + c), muQ = ld2(gs + 2 * HID + c), rstd = ld2(gs + 3 * HID + c);
    const float2 gcore = ld2(p.g_core + (size_t)g * HID + c), gread = ld2(p.g_readout + (size_t)g * HID + c);
    const bool last = (g == p.B - 1) && (p.kl_scale != 0.f);
    const float2 isd = make_float2(1.f / (sd.x + kKlEps), 1.f / (sd.y + kKlEps));
    // ---- attention pass A: S = sum_v alpha_v (gT_v . C_v)
    float S = 0.f;
#pragma unroll 4
    for (int v = v0; v < v1; ++v) {
      const float2 gT = ld2(p.gI2 + (size_t)v * p.gI_stride + c), C = ld2(p.C + (size_t)v * HID + c);
      S += __ldg(p.alpha + v) * warp_sum(gT.x * C.x + gT.y * C.y);
    }
    // ---- pass B: attention backward, gate backward (scalar part), per-graph BN sums.  The loop is latency-bound (one
    //      warp per graph, two dependent shuffle reductions per row): the loads of RB rows are issued together and the
    //      rows' shuffle chains interleave; sums are still accumulated in row order.
    float2 m1 = make_float2(0.f, 0.f), m2 = make_float2(0.f, 0.f);
    constexpr int RB = 4;
    for (int vb = v0; vb < v1; vb += RB) {
      float2 gT[RB], C[RB], gz0[RB], h[RB], fu[RB], q[RB];
      float al[RB], lam[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int v = min(vb + i, v1 - 1);
        gT[i] = ld2(p.gI2 + (size_t)v * p.gI_stride + c); C[i] = ld2(p.C + (size_t)v * HID + c);
        gz0[i] = ld2(p.gI + (size_t)v * p.gI_stride + c); h[i] = ld2(p.H + (size_t)v * HID + c);
        fu[i] = ld2(p.feat_u + (size_t)v * HID + c); q[i] = ld2(p.q + (size_t)v * HID + c);
        al[i] = __ldg(p.alpha + v); lam[i] = __ldg(p.lam + v);
      }
      float dot[RB], part[RB];
      float2 gh[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        dot[i] = gT[i].x * C[i].x + gT[i].y * C[i].y;
        const float2 gz = make_float2(gz0[i].x + gcore.x, gz0[i].y + gcore.y);
        part[i] = gz.x * (h[i].x - muH.x - fu[i].x * sd.x) + gz.y * (h[i].y - muH.y - fu[i].y * sd.y);
        gh[i] = make_float2(fmaf(lam[i], gz.x, gread.x), fmaf(lam[i], gz.y, gread.y));
        if (last) {
          const float k = p.kl_scale / ((float)HID * n);
          const float2 r1 = make_float2(sd.x * isd.x, sd.y * isd.y);                             // sigma/(sigma+e)
          const float2 r2 = make_float2((h[i].x - muH.x) * isd.x, (h[i].y - muH.y) * isd.y);     // (H-mu)/(sigma+e)
          part[i] += k * (-(1.f - lam[i]) * (r1.x * r1.x + r1.y * r1.y) + 2.f * n * lam[i] * (r2.x * r2.x + r2.y * r2.y));
          gh[i].x += k * 2.f * n * lam[i] * lam[i] * r2.x * isd.x;
          gh[i].y += k * 2.f * n * lam[i] * lam[i] * r2.y * isd.y;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          dot[i] += __shfl_xor_sync(0xffffffffu, dot[i], o);
          part[i] += __shfl_xor_sync(0xffffffffu, part[i], o);
        }
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int v = vb + i;
        if (v >= v1) break;
        const float dl = al[i] * (dot[i] - S);
        st2(p.gC + (size_t)v * HID + c, make_float2(fmaf(dl, wc.x, al[i] * gT[i].x), fmaf(dl, wc.y, al[i] * gT[i].y)));
        a_dwc.x = fmaf(dl, C[i].x, a_dwc.x); a_dwc.y = fmaf(dl, C[i].y, a_dwc.y);
        st2(p.gH + (size_t)v * HID + c, gh[i]);
        const float gpv = part[i] * lam[i] * (1.f - lam[i]);
        if (lane == 0) p.gp[v] = gpv;
        a_dbc2 += gpv;
        const float2 qh = make_float2((q[i].x - muQ.x) * rstd.x, (q[i].y - muQ.y) * rstd.y);
        const float ox = fmaf(qh.x, gam.x, bet.x), oy = fmaf(qh.y, gam.y, bet.y);
        a_dw2.x = fmaf(gpv, fmaxf(ox, 0.f), a_dw2.x); a_dw2.y = fmaf(gpv, fmaxf(oy, 0.f), a_dw2.y);
        const float gox = ox > 0.f ? gpv * w2.x : 0.f, goy = oy > 0.f ? gpv * w2.y : 0.f;
        a_db.x += gox; a_db.y += goy;
        a_dg.x = fmaf(gox, qh.x, a_dg.x); a_dg.y = fmaf(goy, qh.y, a_dg.y);
        m1.x = fmaf(gam.x, gox, m1.x); m1.y = fmaf(gam.y, goy, m1.y);
        m2.x = fmaf(gam.x * gox, qh.x, m2.x); m2.y = fmaf(gam.y * goy, qh.y, m2.y);
        // pass C needs gpv of this row again: keep it in the q slot's x lane is not possible (per-lane data) -> re-read gp
      }
    }
    m1.x /= n; m1.y /= n; m2.x /= n; m2.y /= n;
    __syncwarp();
    // ---- pass C: per-graph BN backward -> g_q
    for (int vb = v0; vb < v1; vb += RB) {
      float gpv[RB];
      float2 q[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int v = min(vb + i, v1 - 1);
        gpv[i] = __ldcg(p.gp + v);
        q[i] = ld2(p.q + (size_t)v * HID + c);
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int v = vb + i;
        if (v >= v1) break;
        const float2 qh = make_float2((q[i].x - muQ.x) * rstd.x, (q[i].y - muQ.y) * rstd.y);
        const float ox = fmaf(qh.x, gam.x, bet.x), oy = fmaf(qh.y, gam.y, bet.y);
        const float gox = ox > 0.f ? gpv[i] * w2.x : 0.f, goy = oy > 0.f ? gpv[i] * w2.y : 0.f;
        st2(p.g_q + (size_t)v * HID + c, make_float2(rstd.x * (gam.x * gox - m1.x - qh.x * m2.x),
                                                      rstd.y * (gam.y * goy - m1.y - qh.y * m2.y)));
      }
    }
  }
  // ---- parameter-gradient partials: warp -> CTA -> last CTA
  float* r = s_red + warp * 5 * HID;
  st2(r + c, a_dg); st2(r + HID + c, a_db); st2(r + 2 * HID + c, a_dw2); st2(r + 3 * HID + c, a_dwc);
  if (lane == 0) r[4 * HID] = a_dbc2;
  __syncthreads();
  for (int j = threadIdx.x; j < 4 * HID + 1; j += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += s_red[w * 5 * HID + j];
    p.part[(size_t)blockIdx.x * 5 * HID + j] = s;
  }
  if (!last_cta_arrives(p.counter)) return;
  for (int j = threadIdx.x; j < 4 * HID + 1; j += kThreads) {
    // four interleaved partial sums: 32 independent loads in flight instead of a chain of ~300 L2 round trips
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const int nb = (int)gridDim.x;
    const float* src = p.part + j;
    int b = 0;
#pragma unroll 8
    for (; b + 3 < nb; b += 4) {
      s0 += (double)__ldcg(src + (size_t)b * 5 * HID);
      s1 += (double)__ldcg(src + (size_t)(b + 1) * 5 * HID);
      s2 += (double)__ldcg(src + (size_t)(b + 2) * 5 * HID);
      s3 += (double)__ldcg(src + (size_t)(b + 3) * 5 * HID);
    }
    for (; b < nb; ++b) s0 += (double)__ldcg(src + (size_t)b * 5 * HID);
    const float f = (float)((s0 + s1) + (s2 + s3));
    if (j < HID) p.d_gamma_c[j] = f;
    else if (j < 2 * HID) p.d_beta_c[j - HID] = f;
    else if (j < 3 * HID) p.d_wc2[j - 2 * HID] = f;
    else if (j < 4 * HID) { p.d_attn_w[HID + j - 3 * HID] = f; p.d_attn_w[j - 3 * HID] = 0.f; }  // core half: exactly 0 (F14)
    else { p.d_bc2[0] = f; p.d_attn_b[0] = 0.f; }
  }
}
void launch_graph_gate_bwd(const GraphGateBwdArgs& a, cudaStream_t s) {
  const int grid = min((a.B + 7) / 8, 4 * num_sms());   // latency-bound warp-per-graph loops: as many resident warps as fit
  graph_gate_bwd_kernel<<<grid, kThreads, 0, s>>>(a);
}

// ------------------------------------------------------------------------------------------------
// head MLP forward: Z = Wm2 relu(Wm1 [noisy || alpha*C] + bm1) + bm2
// ------------------------------------------------------------------------------------------------
constexpr int HT = 64;                 // rows per tile
constexpr int HLD = 2 * HID + 4;

struct HeadFwdSmem { float tile[HT * HLD]; float w1t[2 * HID * HID]; float w2t[HID * HID]; };

__device__ __forceinline__ void head_load_tile(float* tile, const float* __restrict__ noisy, const float* __restrict__ C,
                                               const float* __restrict__ alpha, int base, int N, float* imap,
                                               float* aC = nullptr) {
  const int l = threadIdx.x & 31;
  for (int r = threadIdx.x >> 5; r < HT; r += kThreads / 32) {
    const int v = base + r;
    float4 val = make4(0.f);
    if (v < N) {
      if (l < 16) {
        val = ld4(noisy + (size_t)v * HID + l * 4);
      } else {
        const float al = __ldg(alpha + v);
        const float4 cc = ld4(C + (size_t)v * HID + (l - 16) * 4);
        val = make_float4(al * cc.x, al * cc.y, al * cc.z, al * cc.w);
        if (aC) st4(aC + (size_t)v * HID + (l - 16) * 4, val);
      }
      if (imap) st4(imap + (size_t)v * 2 * HID + l * 4, val);
    }
    st4(tile + r * HLD + l * 4, val);
  }
}

__global__ void __launch_bounds__(kThreads, 2)
head_fwd_kernel(HeadFwdArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HeadFwdSmem& sm = *reinterpret_cast<HeadFwdSmem*>(smem_raw);
  using M = NNMap<HT, HID>;
  load_matrix<HID>(sm.w1t, HID, p.W1t, 2 * HID);
  load_matrix<HID>(sm.w2t, HID, p.W2t, HID);
  const int n_tiles = (p.N + HT - 1) / HT;
  const int c0 = M::col0(), r0 = M::row0();
  const float4 b1 = ldg4(p.b1 + c0), b2 = ldg4(p.b2 + c0);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int base = tile * HT;
    __syncthreads();
    head_load_tile(sm.tile, p.noisy, p.C, p.alpha, base, p.N, p.imap, p.aC);
    __syncthreads();
    float acc[M::TM][4];
#pragma unroll
    for (int m = 0; m < M::TM; ++m) { acc[m][0] = b1.x; acc[m][1] = b1.y; acc[m][2] = b1.z; acc[m][3] = b1.w; }
    gemm_nn<HT, 2 * HID, HID>(sm.tile, HLD, sm.w1t, HID, acc);
    __syncthreads();
#pragma unroll
    for (int m = 0; m < M::TM; ++m) {
      const float4 rv = relu4(make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]));
      st4(sm.tile + (r0 + m) * GLD + c0, rv);
      const int v = base + r0 + m;
      if (v < p.N) st4(p.r + (size_t)v * HID + c0, rv);
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < M::TM; ++m) { acc[m][0] = b2.x; acc[m][1] = b2.y; acc[m][2] = b2.z; acc[m][3] = b2.w; }
    gemm_nn<HT, HID, HID>(sm.tile, GLD, sm.w2t, HID, acc);
#pragma unroll
    for (int m = 0; m < M::TM; ++m) {
      const int v = base + r0 + m;
      if (v < p.N) st4(p.Z + (size_t)v * HID + c0, make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]));
    }
  }
}
void launch_head_fwd(const HeadFwdArgs& a, cudaStream_t s) {
  static bool once = (cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(HeadFwdSmem)), true);
  (void)once;
  const int grid = min((a.N + HT - 1) / HT, 2 * num_sms());
  head_fwd_kernel<<<grid, kThreads, sizeof(HeadFwdSmem), s>>>(a);
}

// head MLP backward (persistent): gI = ((gZ W2) * [r>0]) W1 ; dW2 += gZ^T r ; dW1 += g_u^T I ; biases
struct HeadBwdSmem {
  float gy[HT * GLD]; float gu[HT * GLD]; float r[HT * GLD]; float a[HT * HLD];
  float w2[HID * HID]; float w1[HID * 2 * HID];
};

__global__ void __launch_bounds__(kThreads, 1)
head_bwd_kernel(HeadBwdArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HeadBwdSmem& sm = *reinterpret_cast<HeadBwdSmem*>(smem_raw);
  using M2 = NNMap<HT, HID>;
  using M1 = NNMap<HT, 2 * HID>;
  using T2 = TNMap<HID, HID>;
  using T1 = TNMap<HID, 2 * HID>;
  load_matrix<HID>(sm.w2, HID, p.W2, HID);
  load_matrix<2 * HID>(sm.w1, 2 * HID, p.W1, HID);
  float dW2[T2::TO][T2::TJ], dW1[T1::TO][T1::TJ];
#pragma unroll
  for (int i = 0; i < T2::TO; ++i)
#pragma unroll
    for (int j = 0; j < T2::TJ; ++j) dW2[i][j] = 0.f;
#pragma unroll
  for (int i = 0; i < T1::TO; ++i)
#pragma unroll
    for (int j = 0; j < T1::TJ; ++j) dW1[i][j] = 0.f;
  float dbias = 0.f;
  const int n_tiles = (p.N + HT - 1) / HT;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int base = tile * HT;
    __syncthreads();
    load_row_tile<HT, HID>(sm.gy, GLD, p.gZ, base, p.N);
    load_row_tile<HT, HID>(sm.r, GLD, p.r, base, p.N);
    head_load_tile(sm.a, p.noisy, p.C, p.alpha, base, p.N, nullptr);
    __syncthreads();
    {
      float acc[M2::TM][4];
#pragma unroll
      for (int m = 0; m < M2::TM; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
      gemm_nn<HT, HID, HID>(sm.gy, GLD, sm.w2, HID, acc);
      const int c0 = M2::col0(), r0 = M2::row0();
#pragma unroll
      for (int m = 0; m < M2::TM; ++m) {
        const float4 rr = ld4(sm.r + (r0 + m) * GLD + c0);
        st4(sm.gu + (r0 + m) * GLD + c0,
            make_float4(rr.x > 0.f ? acc[m][0] : 0.f, rr.y > 0.f ? acc[m][1] : 0.f,
                        rr.z > 0.f ? acc[m][2] : 0.f, rr.w > 0.f ? acc[m][3] : 0.f));
      }
    }
    __syncthreads();
    {
      float acc[M1::TM][4];
#pragma unroll
      for (int m = 0; m < M1::TM; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
      gemm_nn<HT, HID, 2 * HID>(sm.gu, GLD, sm.w1, 2 * HID, acc);
      const int c0 = M1::col0(), r0 = M1::row0();
#pragma unroll
      for (int m = 0; m < M1::TM; ++m) {
        const int v = base + r0 + m;
        if (v < p.N) st4(p.gI + (size_t)v * 2 * HID + c0, make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]));
      }
    }
    gemm_tn<HID, HID>(sm.gy, GLD, sm.r, GLD, HT, dW2);
    gemm_tn<HID, 2 * HID>(sm.gu, GLD, sm.a, HLD, HT, dW1);
    if (threadIdx.x < 2 * HID) {
      const float* src = (threadIdx.x < HID) ? sm.gy : sm.gu;
      const int c = threadIdx.x & (HID - 1);
      float s = 0.f;
#pragma unroll 8
      for (int r = 0; r < HT; ++r) s += src[r * GLD + c];
      dbias += s;
    }
  }
  float* part = p.part + (size_t)blockIdx.x * p.pstride;
#pragma unroll
  for (int i = 0; i < T2::TO; ++i)
#pragma unroll
    for (int j = 0; j < T2::TJ; ++j) part[p.off_W2 + (T2::o0() + i) * HID + T2::j0() + j] = dW2[i][j];
#pragma unroll
  for (int i = 0; i < T1::TO; ++i)
#pragma unroll
    for (int j = 0; j < T1::TJ; ++j) part[p.off_W1 + (T1::o0() + i) * 2 * HID + T1::j0() + j] = dW1[i][j];
  if (threadIdx.x < HID) part[p.off_b2 + threadIdx.x] = dbias;
  else if (threadIdx.x < 2 * HID) part[p.off_b1 + threadIdx.x - HID] = dbias;
}
__global__ void __launch_bounds__(kThreads)
head_bwd_prep_kernel(const float* __restrict__ W1, float* __restrict__ W1a, float* __restrict__ W1b, float* __restrict__ bn,
                     float* __restrict__ cvec) {
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < HID * 2 * HID; i += gridDim.x * kThreads) {
    const int o = i / (2 * HID), k = i % (2 * HID);
    const float w = W1[i];
    if (k < HID) W1a[o * HID + k] = w; else W1b[o * HID + k - HID] = w;
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < 4 * HID; i += kThreads) bn[i] = (i >= HID && i < 3 * HID) ? 1.f : 0.f;   // mean 0, rstd 1, gamma 1, beta 0
    for (int i = threadIdx.x; i < 2 * HID; i += kThreads) cvec[i] = 0.f;
  }
}
void launch_head_bwd_prep(const float* W1, float* W1a, float* W1b, float* bn, float* cvec, cudaStream_t s) {
  head_bwd_prep_kernel<<<8, kThreads, 0, s>>>(W1, W1a, W1b, bn, cvec);
}

__global__ void __launch_bounds__(kThreads) head_dw1_interleave_kernel(float* __restrict__ dW1) {
  __shared__ float s_w[2 * HID * HID];
  for (int i = threadIdx.x; i < 2 * HID * HID; i += kThreads) s_w[i] = dW1[i];
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * HID * HID; i += kThreads) {
    const int o = i / (2 * HID), k = i % (2 * HID);
    dW1[i] = k < HID ? s_w[o * HID + k] : s_w[HID * HID + o * HID + k - HID];
  }
}
void launch_head_dw1_interleave(float* dW1, cudaStream_t s) { head_dw1_interleave_kernel<<<1, kThreads, 0, s>>>(dW1); }

void launch_head_bwd(const HeadBwdArgs& a, int grid, cudaStream_t s) {
  static bool once = (cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(HeadBwdSmem)), true);
  (void)once;
  head_bwd_kernel<<<grid, kThreads, sizeof(HeadBwdSmem), s>>>(a);
}

}  // namespace scgib
