// Microbenchmark round 2: gather paths that do NOT go through the LSU tag stage, and the conflict-free
// shared-memory layouts the Kx8 LUT kernel relies on.  One variant per process (argv[1]) so that a faulting
// experimental instruction cannot poison the others.
//
//   bulk16     per-code `cp.async.bulk` (TMA unit, 16 B) global -> smem, mbarrier completion, LDS readback
//   gather4_b1 TMA `tile::gather4` with a {8,1} box        gather4_b4  same with a {8,4} box
//   ldgsts     per-code `cp.async.cg` 16 B (LDGSTS) -> smem
//   mix        half of the codes via LDG, half via cp.async.bulk (are the two paths additive?)
//   lds_rep8   256-entry 16-B table replicated 8x, lane%8 picks the replica (conflict-free LDS.128)
//   lut32      4-byte LUT lookups, layout [code][32 banks], bank = lane (conflict-free LDS.32)
//   lut32_rand 4-byte LUT lookups at random banks (conflicts)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/gather_microbench2 tools/gather_microbench2.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t code_of(const uint4& c, int e) {
  const uint32_t w[4] = {c.x, c.y, c.z, c.w};
  return (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
}
__device__ __forceinline__ uint32_t byte_of(const uint4& c, int e) {
  const uint32_t w[4] = {c.x, c.y, c.z, c.w};
  return (w[e >> 2] >> ((e & 3) * 8)) & 0xffu;
}
#define ACC(v) { acc.x ^= (v).x; acc.y += (v).y; acc.z ^= (v).z; acc.w += (v).w; }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

constexpr int TH = 256;

// ---- bulk16: every code -> one 16-byte cp.async.bulk into this thread's smem slots ------------------
template <bool MIX>
__global__ void __launch_bounds__(TH) k_bulk(const uint4* __restrict__ codes, size_t nchunks, const uint4* __restrict__ table, uint4* out) {
  extern __shared__ __align__(128) uint4 slots[];  // [TH][8]
  __shared__ __align__(8) uint64_t bar_mem;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bar_mem);
  if (threadIdx.x == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint4 acc = make_uint4(0, 0, 0, 0);
  uint32_t parity = 0;
  const uint32_t my = (uint32_t)__cvta_generic_to_shared(slots + threadIdx.x * 8);
  constexpr int NB = MIX ? 4 : 8;  // codes per thread that go through the bulk path
  for (size_t c0 = (size_t)blockIdx.x * TH; c0 < nchunks; c0 += (size_t)gridDim.x * TH) {
    const size_t c = c0 + threadIdx.x;
    const bool live = c < nchunks;
    uint4 cw = live ? ld_stream(codes + c) : make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
      const size_t n_live = (nchunks - c0) < (size_t)TH ? (nchunks - c0) : (size_t)TH;
      mbar_expect_tx(bar, (uint32_t)(n_live * NB * 16));
    }
    __syncthreads();  // expect_tx before any complete_tx; also: everyone is done reading the previous round
    if (live) {
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        const uint4* src = table + code_of(cw, e);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                     ::"r"(my + e * 16), "l"(src), "r"(bar) : "memory");
      }
    }
    uint4 v[8];
    if (MIX && live) {
#pragma unroll
      for (int e = 4; e < 8; ++e) {
        const uint4* p = table + code_of(cw, e);
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[e].x), "=r"(v[e].y), "=r"(v[e].z), "=r"(v[e].w) : "l"(p));
      }
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    if (live) {
#pragma unroll
      for (int e = 0; e < NB; ++e) v[e] = slots[threadIdx.x * 8 + e];
#pragma unroll
      for (int e = 0; e < 8; ++e) ACC(v[e]);
    }
  }
  if (acc.x == 0x12345 && acc.y == 0x777) out[0] = acc;
}

// ---- gather4: 2 TMA gather4 instructions per thread per chunk (4 rows of 16 B each) -----------------
__global__ void __launch_bounds__(TH) k_gather4(const uint4* __restrict__ codes, size_t nchunks, const __grid_constant__ CUtensorMap tmap, uint4* out) {
  extern __shared__ __align__(128) uint4 slots[];  // [TH][2][8]: each gather4 gets a 128-byte slot (64 B used)
  __shared__ __align__(8) uint64_t bar_mem;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bar_mem);
  if (threadIdx.x == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint4 acc = make_uint4(0, 0, 0, 0);
  uint32_t parity = 0;
  const uint32_t my = (uint32_t)__cvta_generic_to_shared(slots + threadIdx.x * 16);
  for (size_t c0 = (size_t)blockIdx.x * TH; c0 < nchunks; c0 += (size_t)gridDim.x * TH) {
    const size_t c = c0 + threadIdx.x;
    const bool live = c < nchunks;
    uint4 cw = live ? ld_stream(codes + c) : make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
      const size_t n_live = (nchunks - c0) < (size_t)TH ? (nchunks - c0) : (size_t)TH;
      mbar_expect_tx(bar, (uint32_t)(n_live * 8 * 16));
    }
    __syncthreads();
    if (live) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r0 = code_of(cw, 4 * h), r1 = code_of(cw, 4 * h + 1), r2 = code_of(cw, 4 * h + 2), r3 = code_of(cw, 4 * h + 3);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(my + h * 128), "l"(&tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
      }
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    if (live) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 4; ++e) { uint4 v = slots[threadIdx.x * 16 + h * 8 + e]; ACC(v); }
    }
  }
  if (acc.x == 0x12345 && acc.y == 0x777) out[0] = acc;
}

// ---- mix_g4: per chunk, NG4 gather4 instructions (4 codes each) via TMA and the other codes via LDG -----
template <int NG4>
__global__ void __launch_bounds__(TH) k_mix_g4(const uint4* __restrict__ codes, size_t nchunks, const uint4* __restrict__ table,
                                               const __grid_constant__ CUtensorMap tmap, uint4* out) {
  extern __shared__ __align__(128) uint4 slots[];  // [TH][8] : one 128-byte slot per thread, first 64 B used
  __shared__ __align__(8) uint64_t bar_mem;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bar_mem);
  if (threadIdx.x == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint4 acc = make_uint4(0, 0, 0, 0);
  uint32_t parity = 0;
  const uint32_t my = (uint32_t)__cvta_generic_to_shared(slots + threadIdx.x * 8);
  // each thread handles 2 chunks per round: chunk A fully via LDG (8 codes), chunk B: 4*NG4 codes via gather4, rest LDG
  for (size_t c0 = (size_t)blockIdx.x * TH * 2; c0 < nchunks; c0 += (size_t)gridDim.x * TH * 2) {
    const size_t ca = c0 + threadIdx.x, cb = c0 + TH + threadIdx.x;
    const bool la = ca < nchunks, lb = cb < nchunks;
    uint4 cwa = la ? ld_stream(codes + ca) : make_uint4(0, 0, 0, 0);
    uint4 cwb = lb ? ld_stream(codes + cb) : make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
      size_t n_live = (nchunks > c0 + TH) ? ((nchunks - c0 - TH) < (size_t)TH ? (nchunks - c0 - TH) : (size_t)TH) : 0;
      mbar_expect_tx(bar, (uint32_t)(n_live * NG4 * 64));
    }
    __syncthreads();
    if (lb) {
#pragma unroll
      for (int h = 0; h < NG4; ++h) {
        const int r0 = code_of(cwb, 4 * h), r1 = code_of(cwb, 4 * h + 1), r2 = code_of(cwb, 4 * h + 2), r3 = code_of(cwb, 4 * h + 3);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(my + h * 64), "l"(&tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
      }
    }
    uint4 v[16];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint4* pp = table + code_of(cwa, e);
      asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[e].x), "=r"(v[e].y), "=r"(v[e].z), "=r"(v[e].w) : "l"(pp));
    }
#pragma unroll
    for (int e = 4 * NG4; e < 8; ++e) {
      const uint4* pp = table + code_of(cwb, e);
      asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[8 + e].x), "=r"(v[8 + e].y), "=r"(v[8 + e].z), "=r"(v[8 + e].w) : "l"(pp));
    }
    mbar_wait(bar, parity);
    parity ^= 1;
#pragma unroll
    for (int e = 0; e < 4 * NG4; ++e) v[8 + e] = slots[threadIdx.x * 8 + e];
    if (la) {
#pragma unroll
      for (int e = 0; e < 8; ++e) ACC(v[e]);
    }
    if (lb) {
#pragma unroll
      for (int e = 0; e < 8; ++e) ACC(v[8 + e]);
    }
  }
  if (acc.x == 0x12345 && acc.y == 0x777) out[0] = acc;
}

// ---- mix_tex: half of the codes via LDG, half via the texture unit -----------------------------------
__global__ void __launch_bounds__(TH) k_mix_tex(const uint4* __restrict__ codes, size_t nchunks, const uint4* __restrict__ table,
                                                cudaTextureObject_t tex, uint4* out) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t c = (size_t)blockIdx.x * TH + threadIdx.x; c < nchunks; c += (size_t)gridDim.x * TH) {
    uint4 cw = ld_stream(codes + c);
    uint4 v[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint4* pp = table + code_of(cw, e);
      asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[e].x), "=r"(v[e].y), "=r"(v[e].z), "=r"(v[e].w) : "l"(pp));
    }
#pragma unroll
    for (int e = 4; e < 8; ++e) v[e] = tex1Dfetch<uint4>(tex, (int)code_of(cw, e));
#pragma unroll
    for (int e = 0; e < 8; ++e) ACC(v[e]);
  }
  if (acc.x == 0x12345 && acc.y == 0x777) out[0] = acc;
}

// ---- ldgsts: per-code cp.async 16 B ------------------------------------------------------------------
__global__ void __launch_bounds__(TH) k_ldgsts(const uint4* __restrict__ codes, size_t nchunks, const uint4* __restrict__ table, uint4* out) {
  extern __shared__ __align__(128) uint4 slots[];  // [TH][8]
  uint4 acc = make_uint4(0, 0, 0, 0);
  const uint32_t my = (uint32_t)__cvta_generic_to_shared(slots + threadIdx.x * 8);
  for (size_t c = (size_t)blockIdx.x * TH + threadIdx.x; c < nchunks; c += (size_t)gridDim.x * TH) {
    uint4 cw = ld_stream(codes + c);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint4* src = table + code_of(cw, e);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(my + e * 16), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int e = 0; e < 8; ++e) { uint4 v = slots[threadIdx.x * 8 + e]; ACC(v); }
  }
  if (acc.x == 0x12345 && acc.y == 0x777) out[0] = acc;
}

// ---- lds_rep8: 256-entry table x 8 replicas, lane%8 selects its replica ------------------------------
__global__ void __launch_bounds__(1024) k_lds_rep8(const uint4* __restrict__ codes, size_t nchunks, const uint4* __restrict__ table, uint4* out) {
  extern __shared__ __align__(128) uint4 stab[];  // [256][8]
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) stab[i] = table[i >> 3];
  __syncthreads();
  const int rep = threadIdx.x & 7;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += (size_t)gridDim.x * blockDim.x) {
    uint4 cw = ld_stream(codes + c);
#pragma unroll
    for (int e = 0; e < 16; ++e) { uint4 v = stab[byte_of(cw, e) * 8 + rep]; ACC(v); }
  }
  if (acc.x == 0x12345 && acc.y == 0x777) out[0] = acc;
}

// ---- lut32: 4-byte lookups; BANKED: word address = code*32 + lane (conflict-free) --------------------
template <bool BANKED>
__global__ void __launch_bounds__(1024) k_lut32(const uint4* __restrict__ codes, size_t nchunks, const uint4* __restrict__ table, uint4* out) {
  extern __shared__ __align__(128) float lut[];  // 256*32 floats = 32 KB
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) lut[i] = (float)(i & 255);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float a0 = 0.f, a1 = 0.f;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += (size_t)gridDim.x * blockDim.x) {
    uint4 cw = ld_stream(codes + c);
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
      const uint32_t b0 = byte_of(cw, e), b1 = byte_of(cw, e + 1);
      if (BANKED) { a0 += lut[b0 * 32 + lane]; a1 += lut[b1 * 32 + lane]; }
      else { a0 += lut[(b0 * 32 + b1) & 8191]; a1 += lut[(b1 * 32 + b0) & 8191]; }
    }
  }
  if (a0 + a1 == 12345.f) out[0] = make_uint4(1, 2, 3, 4);
}

template <typename F>
static float time_ms(F launch, int iters) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) launch(i);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < iters; ++i) launch(i);
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / iters;
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const char* variant = argc > 1 ? argv[1] : "bulk16";
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t ncodes = (size_t)14336 * 512;
  const size_t nchunks = ncodes / 8;
  const int NBUF = 12;
  uint4* codes;
  CK(cudaMalloc(&codes, NBUF * nchunks * sizeof(uint4)));
  {
    std::vector<uint32_t> h(NBUF * nchunks * 4);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (auto& w : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; w = (uint32_t)(s >> 16); }
    CK(cudaMemcpy(codes, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  }
  uint4* table;
  CK(cudaMalloc(&table, 65536 * sizeof(uint4)));
  CK(cudaMemset(table, 1, 65536 * sizeof(uint4)));
  uint4* out;
  CK(cudaMalloc(&out, 64));
  const int iters = 24;
#define BUF(i) (codes + (size_t)((i) % NBUF) * nchunks)
  auto report = [&](const char* name, int cps, int th, float ms, double units_per_code) {
    const double g = ncodes * units_per_code / (ms * 1e-3) / 1e9;
    printf("{\"variant\": \"%s\", \"ctas_per_sm\": %d, \"threads\": %d, \"ms\": %.4f, \"Gops_s\": %.1f, \"ops_per_clk_per_sm_at_1.9GHz\": %.3f}\n",
           name, cps, th, ms, g, g / sms / 1.9);
    fflush(stdout);
  };

  if (!strcmp(variant, "bulk16") || !strcmp(variant, "mix")) {
    const bool mix = !strcmp(variant, "mix");
    for (int cps : {2, 4, 6}) {
      const size_t smem = TH * 8 * 16;
      float ms;
      if (mix) {
        CK(cudaFuncSetAttribute(k_bulk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ms = time_ms([&](int i) { k_bulk<true><<<sms * cps, TH, smem>>>(BUF(i), nchunks, table, out); }, iters);
      } else {
        CK(cudaFuncSetAttribute(k_bulk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ms = time_ms([&](int i) { k_bulk<false><<<sms * cps, TH, smem>>>(BUF(i), nchunks, table, out); }, iters);
      }
      CK(cudaDeviceSynchronize());
      report(variant, cps, TH, ms, 1.0);
    }
  } else if (!strcmp(variant, "mix_tex")) {
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = table;
    rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
    rd.res.linear.sizeInBytes = 65536 * sizeof(uint4);
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    for (int cps : {4, 8}) {
      float ms = time_ms([&](int i) { k_mix_tex<<<sms * cps, TH>>>(BUF(i), nchunks, table, tex, out); }, iters);
      report(variant, cps, TH, ms, 1.0);
    }
  } else if (!strncmp(variant, "mix_g4", 6)) {
    encode_fn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    CUtensorMap tmap;
    cuuint64_t dims[2] = {8, 65536};
    cuuint64_t strides[1] = {16};
    cuuint32_t box[2] = {8, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("{\"variant\": \"%s\", \"error\": \"cuTensorMapEncodeTiled=%d\"}\n", variant, (int)r); return 0; }
    const bool two = !strcmp(variant, "mix_g4x2");
    for (int cps : {2, 4, 6}) {
      const size_t smem = TH * 8 * 16;
      float ms;
      if (two) {
        CK(cudaFuncSetAttribute(k_mix_g4<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ms = time_ms([&](int i) { k_mix_g4<2><<<sms * cps, TH, smem>>>(BUF(i), nchunks, table, tmap, out); }, iters);
      } else {
        CK(cudaFuncSetAttribute(k_mix_g4<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ms = time_ms([&](int i) { k_mix_g4<1><<<sms * cps, TH, smem>>>(BUF(i), nchunks, table, tmap, out); }, iters);
      }
      CK(cudaDeviceSynchronize());
      report(variant, cps, TH, ms, 1.0);
    }
  } else if (!strncmp(variant, "gather4", 7)) {
    const int box_rows = !strcmp(variant, "gather4_b4") ? 4 : 1;
    encode_fn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    CUtensorMap tmap;
    cuuint64_t dims[2] = {8, 65536};
    cuuint64_t strides[1] = {16};
    cuuint32_t box[2] = {8, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("{\"variant\": \"%s\", \"error\": \"cuTensorMapEncodeTiled=%d\"}\n", variant, (int)r); return 0; }
    for (int cps : {2, 3}) {
      const size_t smem = TH * 16 * 16;
      CK(cudaFuncSetAttribute(k_gather4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      float ms = time_ms([&](int i) { k_gather4<<<sms * cps, TH, smem>>>(BUF(i), nchunks, tmap, out); }, iters);
      CK(cudaDeviceSynchronize());
      report(variant, cps, TH, ms, 1.0);
    }
  } else if (!strcmp(variant, "ldgsts")) {
    for (int cps : {2, 4, 6}) {
      const size_t smem = TH * 8 * 16;
      CK(cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      float ms = time_ms([&](int i) { k_ldgsts<<<sms * cps, TH, smem>>>(BUF(i), nchunks, table, out); }, iters);
      report(variant, cps, TH, ms, 1.0);
    }
  } else if (!strcmp(variant, "lds_rep8")) {
    for (int cps : {1, 2}) {
      float ms = time_ms([&](int i) { k_lds_rep8<<<sms * cps, 1024, 32768>>>(BUF(i), nchunks, table, out); }, iters);
      report(variant, cps, 1024, ms, 2.0);  // 16 one-byte codes per chunk = 2 lookups per 2-byte code slot
    }
  } else if (!strcmp(variant, "lut32") || !strcmp(variant, "lut32_rand")) {
    const bool banked = !strcmp(variant, "lut32");
    for (int cps : {1, 2}) {
      float ms = banked ? time_ms([&](int i) { k_lut32<true><<<sms * cps, 1024, 32768>>>(BUF(i), nchunks, table, out); }, iters)
                        : time_ms([&](int i) { k_lut32<false><<<sms * cps, 1024, 32768>>>(BUF(i), nchunks, table, out); }, iters);
      report(variant, cps, 1024, ms, 2.0);
    }
  } else {
    printf("{\"error\": \"unknown variant %s\"}\n", variant);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
