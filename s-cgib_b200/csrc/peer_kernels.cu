// peer_kernels.cu - gradient all-reduce FUSED with Adam over NVLink peer memory (one kernel, no NCCL call).
//
// Replaces, under data parallelism, `loss.backward()`'s implicit gradient + `optimizer.step()` of the reference loop
// (exp_pretraining.py:321-323) for one process per GPU: every rank keeps its flat gradient buffer in memory that its
// peers have mapped (CUDA IPC over NVLink 5 / NVSwitch).  The kernel
//   1. publishes "my gradients of step s are complete" to every peer (system-scope release store into the peer's flags),
//   2. waits until every peer has published step s (acquire loads of its own flag array),
//   3. reads all `world` gradient buffers with peer loads, sums them IN RANK ORDER (every rank forms the identical
//      sum, so the replicas stay bit-identical without a broadcast), scales by 1/world and applies Adam.
// The 323 KB gradient never makes a second pass through HBM and there is one launch instead of a NCCL all-reduce plus
// an optimiser kernel.  Gradients are double-buffered by step parity: a rank overwrites buffer s & 1 in backward(s + 2),
// which is stream-ordered after its own step-(s + 1) kernel, which waited for every peer's step-(s + 1) flag, which a
// peer only sets after its step-s kernel (the last reader of buffer s & 1) has completed.
#include <string.h>
#include "kernels.cuh"
#include "../../include/scgib.h"

namespace scgib {

constexpr int kMaxWorld = 16;
struct PeerPtrs {
  const float* grads[kMaxWorld];     // rank r's gradient buffer of this step's parity
  unsigned int* flags[kMaxWorld];    // rank r's flag array [world]: flags[r][q] = last step published by rank q
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld4_volatile(const float* p) {   // peer data: never served from a stale L1 line
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(kThreads)
allreduce_adam_kernel(PeerPtrs pp, int rank, int world, unsigned int seq, float* __restrict__ p, float* __restrict__ m,
                      float* __restrict__ v, int64_t n4, float b1, float b2, float eps, float wd, float step_size,
                      float bc2_sqrt) {
  // 1. publish (the backward kernels that wrote my gradients precede this kernel in stream order)
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(pp.flags[threadIdx.x] + rank, seq);
  }
  // 2. wait for every rank (flags only grow; a rank that is ahead has seq + 1)
  if (threadIdx.x < world) {
    const unsigned int* f = pp.flags[rank] + threadIdx.x;
    unsigned long long t0 = 0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(ld_acquire_sys(f) - seq) < 0) {
      __nanosleep(64);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) __trap();     // a peer died: fail the launch (20 s) instead of hanging the GPU
    }
  }
  __syncthreads();
  // 3. rank-ordered sum of the peers' gradients + Adam (L2-in-gradient weight decay, torch.optim.Adam update order)
  const float inv = 1.f / (float)world;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
    float4 g = ld4_volatile(pp.grads[0] + 4 * i);
    for (int r = 1; r < world; ++r) {
      const float4 t = ld4_volatile(pp.grads[r] + 4 * i);
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    const float4 pi = ld4(p + 4 * i), mi = ld4(m + 4 * i), vi = ld4(v + 4 * i);
    float4 po, mo, vo;
#define SCGIB_ADAM1(c)                                                         \
    {                                                                          \
      const float gi = fmaf(wd, pi.c, g.c * inv);                              \
      mo.c = fmaf(1.f - b1, gi - mi.c, mi.c);                                  \
      vo.c = fmaf(1.f - b2, gi * gi, b2 * vi.c);                               \
      po.c = pi.c - step_size * (mo.c / (sqrtf(vo.c) / bc2_sqrt + eps));       \
    }
    SCGIB_ADAM1(x) SCGIB_ADAM1(y) SCGIB_ADAM1(z) SCGIB_ADAM1(w)
#undef SCGIB_ADAM1
    st4(p + 4 * i, po); st4(m + 4 * i, mo); st4(v + 4 * i, vo);
  }
}

}  // namespace scgib

using namespace scgib;

extern "C" SCGIB_API int scgib_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64) {
  if (!dev_ptr || !handle64 || bytes == 0) return SCGIB_E_NULL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  cudaError_t e = cudaMalloc(dev_ptr, bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(*dev_ptr, 0, bytes);
  if (e != cudaSuccess) return (int)e;
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, *dev_ptr);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle64, &h, 64);
  return (int)cudaDeviceSynchronize();
}

extern "C" SCGIB_API int scgib_peer_open(const unsigned char* handle64, void** dev_ptr) {
  if (!dev_ptr || !handle64) return SCGIB_E_NULL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  return (int)cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" SCGIB_API int scgib_peer_close(void* dev_ptr) { return dev_ptr ? (int)cudaIpcCloseMemHandle(dev_ptr) : SCGIB_E_NULL; }
extern "C" SCGIB_API int scgib_peer_free(void* dev_ptr) { return dev_ptr ? (int)cudaFree(dev_ptr) : SCGIB_E_NULL; }

extern "C" SCGIB_API int scgib_allreduce_adam_f32(float* params, float* exp_avg, float* exp_avg_sq, int64_t n,
                                        const void* const* peer_grads, const void* const* peer_flags, int32_t rank,
                                        int32_t world, uint32_t seq, int64_t step, float lr, float beta1, float beta2,
                                        float eps, float weight_decay, void* stream) {
  if (!params || !exp_avg || !exp_avg_sq || !peer_grads || !peer_flags) return SCGIB_E_NULL;
  if (n < 4 || (n & 3) || step < 1 || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || seq == 0) return SCGIB_E_RANGE;
  if (((uintptr_t)params & 15u) || ((uintptr_t)exp_avg & 15u) || ((uintptr_t)exp_avg_sq & 15u)) return SCGIB_E_ALIGN;
  PeerPtrs pp;
  for (int r = 0; r < world; ++r) {
    if (!peer_grads[r] || !peer_flags[r]) return SCGIB_E_NULL;
    pp.grads[r] = (const float*)peer_grads[r];
    pp.flags[r] = (unsigned int*)peer_flags[r];
  }
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const int64_t n4 = n / 4;
  const int grid = (int)((n4 + kThreads - 1) / kThreads < 2 * (int64_t)num_sms() ? (n4 + kThreads - 1) / kThreads : 2 * (int64_t)num_sms());
  allreduce_adam_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(pp, rank, world, seq, params, exp_avg, exp_avg_sq, n4,
                                                                     beta1, beta2, eps, weight_decay, (float)(lr / bc1),
                                                                     (float)sqrt(bc2));
  return (int)cudaGetLastError();
}
