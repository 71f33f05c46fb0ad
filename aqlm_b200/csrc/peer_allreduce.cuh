// One-shot all-reduce of the sharded matvec's fp32 partials over NVLink PEER MEMORY, fused with the scale+bias epilogue.
//
// New work (the reference has no multi-GPU hot path).  BASELINE configs[4] asks for one all-reduce of the
// partial output vector per linear; at 32-112 KiB that collective is pure latency (NCCL: ~10-20 us, several times the
// shard's compute).  Here every rank owns a buffer that all peers have mapped (cudaIpc): in ONE kernel a rank
//   A. pushes its partial vector into slot [set][my_rank] of every peer's buffer with 16-byte P2P stores,
//   B. publishes flag[my_rank][slice] = step on every peer with a system-scope release store (per CTA slice),
//   C. waits until the W flags of that slice in its OWN buffer reach `step` (system-scope acquire),
//   D. adds the W partial vectors in rank order (deterministic), applies scale + bias, writes the output.
// Two buffer sets alternate by step parity: a rank can be at most one step ahead of any peer (it needs the peer's
// flag of step s before it can finish s), so set (s+1)&1 is never still being read when it is overwritten.
#pragma once

#include "common.cuh"

namespace aqlm_b200 {

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerMaxCtas = 16;     // CTAs of the stand-alone exchange kernel below
constexpr int kPeerFlagStride = 256; // flags per source rank: one per CTA of the FUSED GEMV+exchange kernel (grid = SM count)
constexpr int kPeerFlagBytes = kPeerMaxWorld * kPeerFlagStride * 4;  // flag[src rank][cta]
constexpr int kPeerThreads = 512;

struct PeerParams {
  uint8_t* peer_base[kPeerMaxWorld];  // peer r's shared buffer as mapped in THIS process (peer_base[rank] = own)
  const float* local;                 // [n] this rank's partials
  const void* scales;                 // [out]
  const void* bias;                   // [out] or null
  void* y;                            // [batch, out] T
  unsigned int* step;                 // local device counter (steps completed)
  unsigned int* tickets;              // [2] local, zero on entry, left zero
  long long max_elems;                // floats per (set, source rank) slot
  int n;                              // batch * out (multiple of 4)
  int out_features;
  int rank, world;
};

__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ float* peer_slot(uint8_t* base, long long max_elems, int world, int set, int src) {
  return reinterpret_cast<float*>(base + kPeerFlagBytes) + ((long long)set * world + src) * max_elems;
}

template <typename T>
__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_epilogue_kernel(const PeerParams p) {
  griddep_launch_dependents();
  griddep_wait();  // the partials come from the GEMV launched just before
  const unsigned int s = *p.step + 1u;  // the step this call completes (only this kernel's last CTA writes *step)
  const int set = (int)(s & 1u);
  const int n4 = p.n >> 2;
  // each CTA owns one contiguous slice of the vector end to end: push it, signal it, wait for it, reduce it
  const int per = (n4 + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(n4, i0 + per);

  // ---- A: push my slice into every rank's slot [set][my rank] (own copy included) ----
  const float4* loc = reinterpret_cast<const float4*>(p.local);
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const float4 v = loc[i];
#pragma unroll 1
    for (int r = 0; r < p.world; ++r)
      reinterpret_cast<float4*>(peer_slot(p.peer_base[r], p.max_elems, p.world, set, p.rank))[i] = v;
  }
  // ---- B: CTA barrier, then thread r publishes flag[my rank][slice] = s on rank r with a system-scope RELEASE store
  //      (cumulative over the whole CTA's stores through the barrier); C: thread r polls flag[r][slice] in MY buffer ----
  __syncthreads();
  if ((int)threadIdx.x < p.world) {
    const int r = threadIdx.x;
    st_release_sys_u32(reinterpret_cast<unsigned int*>(p.peer_base[r]) + p.rank * kPeerFlagStride + blockIdx.x, s);
    const unsigned int* f = reinterpret_cast<const unsigned int*>(p.peer_base[p.rank]) + r * kPeerFlagStride + blockIdx.x;
    while ((int)(ld_acquire_sys_u32(f) - s) < 0) {
    }
  }
  __syncthreads();
  // ---- D: fixed-order sum over source ranks + scale + bias ----
  T* y = reinterpret_cast<T*>(p.y);
  const T* sc = reinterpret_cast<const T*>(p.scales);
  const T* bi = reinterpret_cast<const T*>(p.bias);
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < p.world; ++r) {
      const float4 v = ld_cg_f4(reinterpret_cast<const float4*>(peer_slot(p.peer_base[p.rank], p.max_elems, p.world, set, r)) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const int e = i << 2;
    const int o = e % p.out_features;  // out_features % 4 == 0, so the 4 elements share a batch row
    const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float sv = DT<T>::to_float(sc[o + u]);
      const float bv = bi ? DT<T>::to_float(bi[o + u]) : 0.f;
      y[e + u] = DT<T>::from_float(fmaf(a[u], sv, bv));
    }
  }
  // ---- step bookkeeping (off the critical path): the last CTA to finish advances the local step counter ----
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int old = atomicAdd(p.tickets + 1, 1u);
    if (old == gridDim.x - 1) {
      p.tickets[1] = 0u;
      *p.step = s;
    }
  }
}

}  // namespace aqlm_b200
