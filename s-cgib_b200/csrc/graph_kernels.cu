// graph_kernels.cu - k-hop ego-network extraction on the GPU (warp-per-seed frontier expansion in
// shared memory) + the exclusive scans that size the flattened ego batch.
//
// Replaces, per batch, what the reference does offline with one Python call per node:
//   dgl.khop_in_subgraph(g, v, k)[0]   (reference exp_pretraining.py:271, exp_pcqm4mv2.py:422,425)
//   dgl.batch(chain(batch_subgraphs))  (reference exp_pretraining.py:308-309)
// Node lists are bit-exact with DGL's: ascending parent ids, containing the seed; the induced CSR
// keeps neighbours ascending.
#include "common.cuh"
#include "../../include/scgib.h"

namespace scgib {

constexpr int kEgoWarps = 8;  // seeds per CTA
constexpr int CAP = SCGIB_EGO_CAP;

// Build the k-hop ball of seed v (unsorted BFS order) in `ball`, then rank-sort it into `sorted`.
// Returns the (warp-uniform) ball size, clamped to CAP; sets *status on overflow.
__device__ __forceinline__ int build_sorted_ball(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                 int v, int k, int* ball, int* sorted, int lane, int32_t* status) {
  int m = 1;
  if (lane == 0) ball[0] = v;
  __syncwarp();
  int fb = 0, fe = 1;
  bool overflow = false;
  for (int hop = 0; hop < k && fb < fe && !overflow; ++hop) {
    for (int f = fb; f < fe && !overflow; ++f) {
      const int u = ball[f];
      const int e0 = __ldg(indptr + u), e1 = __ldg(indptr + u + 1);
      for (int eb = e0; eb < e1; eb += 32) {
        const int e = eb + lane;
        const int w = (e < e1) ? __ldg(indices + e) : -1;
        bool isnew = (w >= 0);
        if (isnew) {
          for (int i = 0; i < m; ++i)
            if (ball[i] == w) { isnew = false; break; }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, isnew);
        const int pos = m + __popc(mask & ((1u << lane) - 1u));
        if (isnew && pos < CAP) ball[pos] = w;
        m += __popc(mask);
        if (m > CAP) { m = CAP; overflow = true; }
        __syncwarp();
        if (overflow) break;
      }
    }
    fb = fe;
    fe = m;
  }
  if (overflow && lane == 0) atomicExch(status, 1);
  // rank sort (all ids distinct)
  for (int i = lane; i < m; i += 32) {
    const int val = ball[i];
    int rank = 0;
    for (int j = 0; j < m; ++j) rank += (ball[j] < val);
    sorted[rank] = val;
  }
  __syncwarp();
  return m;
}

__device__ __forceinline__ int find_sorted(const int* sorted, int m, int w) {
  int lo = 0, hi = m;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sorted[mid] < w) lo = mid + 1; else hi = mid;
  }
  return (lo < m && sorted[lo] == w) ? lo : -1;
}

__global__ void __launch_bounds__(kEgoWarps * 32)
ego_count_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int N, int k,
                 int32_t* __restrict__ cnt_nodes, int32_t* __restrict__ cnt_edges, int32_t* status) {
  __shared__ int s_ball[kEgoWarps][CAP];
  __shared__ int s_sorted[kEgoWarps][CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int v = blockIdx.x * kEgoWarps + warp; v < N; v += gridDim.x * kEgoWarps) {
    const int m = build_sorted_ball(indptr, indices, v, k, s_ball[warp], s_sorted[warp], lane, status);
    int cnt = 0;
    for (int a = lane; a < m; a += 32) {
      const int u = s_sorted[warp][a];
      const int e0 = __ldg(indptr + u), e1 = __ldg(indptr + u + 1);
      for (int e = e0; e < e1; ++e) cnt += (find_sorted(s_sorted[warp], m, __ldg(indices + e)) >= 0);
    }
    cnt = (int)warp_sum((float)cnt);  // counts are small (< 2^24): exact in fp32
    if (lane == 0) { cnt_nodes[v] = m; cnt_edges[v] = cnt; }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kEgoWarps * 32)
ego_fill_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int N, int k,
                const int32_t* __restrict__ ego_ptr, const int32_t* __restrict__ ego_eptr,
                int32_t* __restrict__ ego_nodes, int32_t* __restrict__ ego_seed,
                int32_t* __restrict__ sub_indptr, int32_t* __restrict__ sub_indices) {
  __shared__ int s_ball[kEgoWarps][CAP];
  __shared__ int s_sorted[kEgoWarps][CAP];
  __shared__ int32_t s_status;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int v = blockIdx.x * kEgoWarps + warp; v < N; v += gridDim.x * kEgoWarps) {
    const int m = build_sorted_ball(indptr, indices, v, k, s_ball[warp], s_sorted[warp], lane, &s_status);
    const int row_base = __ldg(ego_ptr + v);
    int edge_off = __ldg(ego_eptr + v);
    for (int a0 = 0; a0 < m; a0 += 32) {
      const int a = a0 + lane;
      int deg = 0, u = -1, e0 = 0, e1 = 0;
      if (a < m) {
        u = s_sorted[warp][a];
        e0 = __ldg(indptr + u); e1 = __ldg(indptr + u + 1);
        for (int e = e0; e < e1; ++e) deg += (find_sorted(s_sorted[warp], m, __ldg(indices + e)) >= 0);
      }
      // warp exclusive scan of deg
      int incl = deg;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (a < m) {
        int off = edge_off + incl - deg;
        ego_nodes[row_base + a] = u;
        ego_seed[row_base + a] = v;
        sub_indptr[row_base + a] = off;
        for (int e = e0; e < e1; ++e) {
          const int p = find_sorted(s_sorted[warp], m, __ldg(indices + e));
          if (p >= 0) sub_indices[off++] = row_base + p;
        }
      }
      edge_off += total;
    }
    if (v == N - 1 && lane == 0) sub_indptr[row_base + m] = edge_off;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of two int32 arrays at once (3 phases, deterministic); out has n+1 entries
// ------------------------------------------------------------------------------------------------
constexpr int kScanItems = 4;
constexpr int kScanBlock = kThreads * kScanItems;

__device__ __forceinline__ int2 block_scan_incl(int2 v, int2* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int tx = __shfl_up_sync(0xffffffffu, v.x, o);
    const int ty = __shfl_up_sync(0xffffffffu, v.y, o);
    if (lane >= o) { v.x += tx; v.y += ty; }
  }
  if (lane == 31) s_warp[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int2 w = (lane < kThreads / 32) ? s_warp[lane] : make_int2(0, 0);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int tx = __shfl_up_sync(0xffffffffu, w.x, o);
      const int ty = __shfl_up_sync(0xffffffffu, w.y, o);
      if (lane >= o) { w.x += tx; w.y += ty; }
    }
    if (lane < kThreads / 32) s_warp[lane] = w;
  }
  __syncthreads();
  if (warp > 0) { v.x += s_warp[warp - 1].x; v.y += s_warp[warp - 1].y; }
  return v;
}

__global__ void __launch_bounds__(kThreads)
scan_block_sums_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int n, int2* __restrict__ bsum) {
  __shared__ int2 s_warp[kThreads / 32];
  const int base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
  int2 t = make_int2(0, 0);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) { t.x += a[base + i]; t.y += b[base + i]; }
  const int2 incl = block_scan_incl(t, s_warp);
  if (threadIdx.x == kThreads - 1) bsum[blockIdx.x] = incl;
}

__global__ void __launch_bounds__(kThreads)
scan_spine_kernel(int2* __restrict__ bsum, int nb) {  // single CTA: exclusive scan of block sums, in place
  __shared__ int2 s_warp[kThreads / 32];
  __shared__ int2 s_carry;
  if (threadIdx.x == 0) s_carry = make_int2(0, 0);
  __syncthreads();
  for (int c = 0; c < nb; c += kThreads) {
    const int i = c + threadIdx.x;
    const int2 v = (i < nb) ? bsum[i] : make_int2(0, 0);
    const int2 incl = block_scan_incl(v, s_warp);
    const int2 carry = s_carry;
    if (i < nb) bsum[i] = make_int2(carry.x + incl.x - v.x, carry.y + incl.y - v.y);
    __syncthreads();
    if (threadIdx.x == kThreads - 1) s_carry = make_int2(carry.x + incl.x, carry.y + incl.y);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads)
scan_apply_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int n, const int2* __restrict__ bsum,
                  int32_t* __restrict__ oa, int32_t* __restrict__ ob) {
  __shared__ int2 s_warp[kThreads / 32];
  const int base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
  int va[kScanItems], vb[kScanItems];
  int2 t = make_int2(0, 0);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    va[i] = (base + i < n) ? a[base + i] : 0;
    vb[i] = (base + i < n) ? b[base + i] : 0;
    t.x += va[i]; t.y += vb[i];
  }
  const int2 incl = block_scan_incl(t, s_warp);
  const int2 off = bsum[blockIdx.x];
  int ra = off.x + incl.x - t.x, rb = off.y + incl.y - t.y;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) { oa[base + i] = ra; ob[base + i] = rb; }
    ra += va[i]; rb += vb[i];
    if (base + i == n - 1) { oa[n] = ra; ob[n] = rb; }
  }
}

// ------------------------------------------------------------------------------------------------
// On-device batch assembly (replaces DataLoader + MoleculeDataset.collate + dgl.batch, molecules.py:349-362,
// exp_pretraining.py:283): the packed dataset shard (all molecules as one CSR) stays resident in HBM and a batch is
// the list of B molecule ids.  dgl.batch semantics: molecules in list order, node / edge ids offset.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
batch_count_kernel(const int32_t* __restrict__ mol_ptr, const int32_t* __restrict__ ds_indptr,
                   const int32_t* __restrict__ ids, int B, int32_t* __restrict__ cnt_nodes, int32_t* __restrict__ cnt_edges) {
  const int b = blockIdx.x * kThreads + threadIdx.x;
  if (b >= B) return;
  const int m = __ldg(ids + b);
  const int n0 = __ldg(mol_ptr + m), n1 = __ldg(mol_ptr + m + 1);
  cnt_nodes[b] = n1 - n0;
  cnt_edges[b] = __ldg(ds_indptr + n1) - __ldg(ds_indptr + n0);
}

__global__ void __launch_bounds__(kThreads)
batch_fill_kernel(const int32_t* __restrict__ mol_ptr, const int32_t* __restrict__ ds_indptr,
                  const int32_t* __restrict__ ds_indices, const float* __restrict__ ds_x, int F,
                  const int32_t* __restrict__ ids, int B, const int32_t* __restrict__ graph_ptr,
                  const int32_t* __restrict__ edge_ptr, int32_t* __restrict__ indptr, int32_t* __restrict__ indices,
                  float* __restrict__ x) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (kThreads / 32);
  for (int b = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); b < B; b += warps) {
    const int m = __ldg(ids + b);
    const int s0 = __ldg(mol_ptr + m), n = __ldg(mol_ptr + m + 1) - s0;     // source node range
    const int d0 = __ldg(graph_ptr + b);                                    // destination node offset
    const int se0 = __ldg(ds_indptr + s0), de0 = __ldg(edge_ptr + b);
    const int ne = __ldg(ds_indptr + s0 + n) - se0;
    for (int j = lane; j < n; j += 32) indptr[d0 + j] = __ldg(ds_indptr + s0 + j) - se0 + de0;
    if (b == B - 1 && lane == 0) indptr[d0 + n] = de0 + ne;
    for (int e = lane; e < ne; e += 32) indices[de0 + e] = __ldg(ds_indices + se0 + e) - s0 + d0;
    const float* xs = ds_x + (size_t)s0 * F;
    float* xd = x + (size_t)d0 * F;
    for (int i = lane; i < n * F; i += 32) xd[i] = __ldg(xs + i);
  }
}

// ------------------------------------------------------------------------------------------------
// Input validation (failure detection before a step is launched): the conditions the kernels rely on.
// status[0] = 0 ok, else the smallest violated code; status[1] = an offending node / graph id.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
batch_validate_kernel(const int32_t* __restrict__ graph_ptr, const int32_t* __restrict__ indptr,
                      const int32_t* __restrict__ indices, int B, int N, int E, int32_t* status) {
  auto fail = [&](int code, int where) {
    const int old = atomicCAS(status, 0, code);
    if (old == 0 || code < old) { atomicMin(status, code); status[1] = where; }
  };
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i == 0) {
    if (graph_ptr[0] != 0 || graph_ptr[B] != N) fail(1, 0);            // graph_ptr does not cover [0, N)
    if (indptr[0] != 0 || indptr[N] != E) fail(2, 0);                  // indptr does not cover [0, E)
  }
  if (i < B) {
    const int n = graph_ptr[i + 1] - graph_ptr[i];
    if (n < 2) fail(3, i);                                              // per-graph BatchNorm / unbiased std need n >= 2 (models.py:642-647)
  }
  if (i < N) {
    const int e0 = indptr[i], e1 = indptr[i + 1];
    if (e1 < e0 || e0 < 0 || e1 > E) { fail(4, i); return; }            // indptr not monotone
    int lo = 0, hi = B;                                                 // the graph of node i
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (graph_ptr[mid] <= i) lo = mid; else hi = mid; }
    const int g0 = graph_ptr[lo], g1 = graph_ptr[lo + 1];
    int prev = -1;
    for (int e = e0; e < e1; ++e) {
      const int u = indices[e];
      if (u < g0 || u >= g1) { fail(5, i); break; }                     // neighbour outside the node's own graph
      if (u <= prev) { fail(6, i); break; }                             // neighbours not strictly ascending (to_bidirected order)
      if (u == i) { fail(7, i); break; }                                // self loop
      prev = u;
    }
  }
}

}  // namespace scgib

using namespace scgib;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" SCGIB_API size_t scgib_ego_workspace_bytes(int32_t N) {
  const size_t nb = ((size_t)N + kScanBlock - 1) / kScanBlock;
  return align_up((size_t)2 * N * sizeof(int32_t), 256) + align_up((nb + 1) * sizeof(int2), 256) + 256;
}

extern "C" SCGIB_API int scgib_ego_count(const int32_t* indptr, const int32_t* indices, int32_t N, int32_t k,
                               int32_t* ego_ptr, int32_t* ego_eptr, int32_t* status,
                               void* workspace, size_t workspace_bytes, void* stream_) {
  if (!indptr || !indices || !ego_ptr || !ego_eptr || !status || !workspace) return SCGIB_E_NULL;
  if (N < 1 || k < 1) return SCGIB_E_RANGE;
  if (workspace_bytes < scgib_ego_workspace_bytes(N)) return SCGIB_E_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  int32_t* cnt_nodes = (int32_t*)workspace;
  int32_t* cnt_edges = cnt_nodes + N;
  int2* bsum = (int2*)((char*)workspace + align_up((size_t)2 * N * sizeof(int32_t), 256));
  cudaMemsetAsync(status, 0, sizeof(int32_t), stream);
  const int grid = (N + kEgoWarps - 1) / kEgoWarps;
  ego_count_kernel<<<grid, kEgoWarps * 32, 0, stream>>>(indptr, indices, N, k, cnt_nodes, cnt_edges, status);
  const int nb = (N + kScanBlock - 1) / kScanBlock;
  scan_block_sums_kernel<<<nb, kThreads, 0, stream>>>(cnt_nodes, cnt_edges, N, bsum);
  scan_spine_kernel<<<1, kThreads, 0, stream>>>(bsum, nb);
  scan_apply_kernel<<<nb, kThreads, 0, stream>>>(cnt_nodes, cnt_edges, N, bsum, ego_ptr, ego_eptr);
  return (int)cudaGetLastError();
}

extern "C" SCGIB_API int scgib_ego_fill(const int32_t* indptr, const int32_t* indices, int32_t N, int32_t k,
                              const int32_t* ego_ptr, const int32_t* ego_eptr,
                              int32_t* ego_nodes, int32_t* ego_seed, int32_t* sub_indptr, int32_t* sub_indices,
                              void* stream_) {
  if (!indptr || !indices || !ego_ptr || !ego_eptr || !ego_nodes || !ego_seed || !sub_indptr) return SCGIB_E_NULL;
  if (N < 1 || k < 1) return SCGIB_E_RANGE;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int grid = (N + kEgoWarps - 1) / kEgoWarps;
  ego_fill_kernel<<<grid, kEgoWarps * 32, 0, stream>>>(indptr, indices, N, k, ego_ptr, ego_eptr, ego_nodes, ego_seed,
                                                       sub_indptr, sub_indices);
  return (int)cudaGetLastError();
}

extern "C" SCGIB_API size_t scgib_batch_workspace_bytes(int32_t B) { return scgib_ego_workspace_bytes(B); }

extern "C" SCGIB_API int scgib_batch_assemble_count(const int32_t* mol_ptr, const int32_t* ds_indptr, const int32_t* ids,
                                          int32_t B, int32_t* graph_ptr, int32_t* edge_ptr, void* workspace,
                                          size_t workspace_bytes, void* stream_) {
  if (!mol_ptr || !ds_indptr || !ids || !graph_ptr || !edge_ptr || !workspace) return SCGIB_E_NULL;
  if (B < 1) return SCGIB_E_RANGE;
  if (workspace_bytes < scgib_batch_workspace_bytes(B)) return SCGIB_E_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  int32_t* cnt_nodes = (int32_t*)workspace;
  int32_t* cnt_edges = cnt_nodes + B;
  int2* bsum = (int2*)((char*)workspace + align_up((size_t)2 * B * sizeof(int32_t), 256));
  batch_count_kernel<<<(B + kThreads - 1) / kThreads, kThreads, 0, stream>>>(mol_ptr, ds_indptr, ids, B, cnt_nodes, cnt_edges);
  const int nb = (B + kScanBlock - 1) / kScanBlock;
  scan_block_sums_kernel<<<nb, kThreads, 0, stream>>>(cnt_nodes, cnt_edges, B, bsum);
  scan_spine_kernel<<<1, kThreads, 0, stream>>>(bsum, nb);
  scan_apply_kernel<<<nb, kThreads, 0, stream>>>(cnt_nodes, cnt_edges, B, bsum, graph_ptr, edge_ptr);
  return (int)cudaGetLastError();
}

extern "C" SCGIB_API int scgib_batch_assemble_fill(const int32_t* mol_ptr, const int32_t* ds_indptr, const int32_t* ds_indices,
                                         const float* ds_x, int32_t F, const int32_t* ids, int32_t B,
                                         const int32_t* graph_ptr, const int32_t* edge_ptr, int32_t* indptr,
                                         int32_t* indices, float* x, void* stream_) {
  if (!mol_ptr || !ds_indptr || !ds_x || !ids || !graph_ptr || !edge_ptr || !indptr || !x) return SCGIB_E_NULL;
  if (B < 1 || F < 1) return SCGIB_E_RANGE;
  const int grid = (B + kThreads / 32 - 1) / (kThreads / 32);
  batch_fill_kernel<<<grid, kThreads, 0, (cudaStream_t)stream_>>>(mol_ptr, ds_indptr, ds_indices, ds_x, F, ids, B, graph_ptr,
                                                                edge_ptr, indptr, indices, x);
  return (int)cudaGetLastError();
}

extern "C" SCGIB_API int scgib_batch_validate(const int32_t* graph_ptr, const int32_t* indptr, const int32_t* indices, int32_t B,
                                    int32_t N, int32_t E, int32_t* status, void* stream_) {
  if (!graph_ptr || !indptr || !status || (E > 0 && !indices)) return SCGIB_E_NULL;
  if (B < 1 || N < 1 || E < 0) return SCGIB_E_RANGE;
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), stream);
  const int n = N > B ? N : B;
  batch_validate_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, stream>>>(graph_ptr, indptr, indices, B, N, E, status);
  return (int)cudaGetLastError();
}
