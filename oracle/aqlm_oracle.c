/*
 * CPU ORACLE (C restatement) for the AQLM quantized-linear hot path.
 * TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product (aqlm_b200) never does.
 *
 * Parity pinning: validated against outputs of the reference itself (tests/golden/ .npz files, produced by
 * tests/golden/make_golden.py importing /root/reference/inference_lib/src/aqlm) through
 * tests/test_oracle_golden.py.
 *
 * Restated algorithms (paths relative to /root/reference):
 *   - dequantize + linear:  inference_lib/src/aqlm/utils.py:43-70 (_dequantize_weight) and
 *                           inference_lib/src/aqlm/inference_kernels/dequantization.py:9-21
 *   - code unpacking:       inference_lib/src/aqlm/utils.py:29-31 (signed storage -> unsigned code)
 *   - LUT GEMV:             inference_lib/src/aqlm/inference_kernels/numba_kernel.py:37-48
 *                           (the CPU kernel benchmark/matmul_benchmark_cpu.py:100-111 times), made
 *                           race-free: the reference does `output_vec[i] +=` inside numba.prange
 *                           (numba_kernel.py:43-46), which loses updates with >1 thread.
 *
 * Only out_group_size == 1 is implemented here (every published scheme; the numpy oracle covers the
 * general case).  All floating inputs are float32 arrays: the caller widens fp16/bf16 exactly.
 *
 * Build: make -C oracle   (gcc -O3 -pthread -shared -fPIC; see oracle/Makefile)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ---- minimal pthread parallel-for (no OpenMP runtime in the image) ---- */
typedef void (*range_fn)(int64_t begin, int64_t end, int tid, void* ctx);
typedef struct { range_fn fn; int64_t begin, end; int tid; void* ctx; } task_t;
static void* task_main(void* p) { task_t* t = (task_t*)p; t->fn(t->begin, t->end, t->tid, t->ctx); return NULL; }
static void parallel_for(int64_t n, int nthreads, range_fn fn, void* ctx) {
  if (nthreads > n) nthreads = (int)(n > 0 ? n : 1);
  if (nthreads <= 1) { fn(0, n, 0, ctx); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  task_t* ts = (task_t*)malloc(sizeof(task_t) * nthreads);
  for (int t = 0; t < nthreads; ++t) {
    ts[t].fn = fn; ts[t].ctx = ctx; ts[t].tid = t;
    ts[t].begin = n * t / nthreads; ts[t].end = n * (t + 1) / nthreads;
    pthread_create(&th[t], NULL, task_main, &ts[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th); free(ts);
}

/* utils.py:29-31: data.to(int64) % 2**nbits  (two's complement bit pattern == unsigned code) */
static inline uint32_t unpack_code(const void* codes, int itemsize, int64_t idx, int nbits) {
  uint32_t mask = (nbits >= 32) ? 0xffffffffu : ((1u << nbits) - 1u);
  if (itemsize == 1) return ((const uint8_t*)codes)[idx] & mask;
  if (itemsize == 2) return ((const uint16_t*)codes)[idx] & mask;
  return ((const uint32_t*)codes)[idx] & mask;
}

int aqlm_oracle_num_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

typedef struct {
  const void* codes; int code_itemsize; const float* codebooks; const float* scales; const float* bias;
  const float* x; int64_t batch; float* out; /* W or y */
  int64_t out_features, in_groups; int num_codebooks, nbits, g;
} dq_ctx;

/* one dequantized weight row: utils.py:60-68 (embedding_bag sum over codebooks, then * scales) */
static inline void dequant_row(const dq_ctx* c, int64_t o, float* wrow) {
  const int64_t cb_size = (int64_t)1 << c->nbits;
  const int g = c->g;
  const float s = c->scales ? c->scales[o] : 1.0f;
  for (int64_t j = 0; j < c->in_groups; ++j) {
    float acc[64];
    for (int e = 0; e < g; ++e) acc[e] = 0.0f;
    for (int k = 0; k < c->num_codebooks; ++k) {
      uint32_t code = unpack_code(c->codes, c->code_itemsize, (o * c->in_groups + j) * c->num_codebooks + k, c->nbits);
      const float* v = c->codebooks + ((int64_t)k * cb_size + code) * g;
      for (int e = 0; e < g; ++e) acc[e] += v[e];
    }
    for (int e = 0; e < g; ++e) wrow[j * g + e] = acc[e] * s;
  }
}

static void dequant_rows(int64_t begin, int64_t end, int tid, void* p) {
  (void)tid;
  const dq_ctx* c = (const dq_ctx*)p;
  for (int64_t o = begin; o < end; ++o) dequant_row(c, o, c->out + o * c->in_groups * c->g);
}

/* utils.py:43-70.  W[o, j*g + e] = scales[o] * sum_c codebooks[c, code[o,j,c], 0, e].
 * codes [out, in_groups, K] packed; codebooks [K, 2^nbits, 1, g] f32; scales [out] f32 or NULL. */
int aqlm_oracle_dequantize_weight(const void* codes, int code_itemsize, const float* codebooks,
                                  const float* scales, float* W, int64_t out_features, int64_t in_groups,
                                  int num_codebooks, int nbits, int g, int nthreads) {
  if (code_itemsize != 1 && code_itemsize != 2 && code_itemsize != 4) return -1;
  if (g > 64) return -2;
  if (nthreads <= 0) nthreads = aqlm_oracle_num_threads();
  dq_ctx c = {codes, code_itemsize, codebooks, scales, NULL, NULL, 0, W, out_features, in_groups, num_codebooks, nbits, g};
  parallel_for(out_features, nthreads, dequant_rows, &c);
  return 0;
}

static void gemm_rows(int64_t begin, int64_t end, int tid, void* p) {
  (void)tid;
  const dq_ctx* c = (const dq_ctx*)p;
  const int64_t in_features = c->in_groups * c->g;
  float* wrow = (float*)malloc(sizeof(float) * (size_t)in_features);
  for (int64_t o = begin; o < end; ++o) {
    dequant_row(c, o, wrow);
    for (int64_t b = 0; b < c->batch; ++b) { /* F.linear, dequantization.py:21 */
      const float* xb = c->x + b * in_features;
      double sum = 0.0;
      for (int64_t k = 0; k < in_features; ++k) sum += (double)wrow[k] * (double)xb[k];
      float r = (float)sum;
      if (c->bias) r += c->bias[o];
      c->out[b * c->out_features + o] = r;
    }
  }
  free(wrow);
}

/* dequantization.py:9-21: y = F.linear(x, dequantize(codes), bias).  x [batch, in], y [batch, out].
 * One weight row is materialised at a time (never the whole W) -- same arithmetic, less memory. */
int aqlm_oracle_dequantize_gemm(const float* x, int64_t batch, const void* codes, int code_itemsize,
                                const float* codebooks, const float* scales, const float* bias, float* y,
                                int64_t out_features, int64_t in_groups, int num_codebooks, int nbits, int g,
                                int nthreads) {
  if (code_itemsize != 1 && code_itemsize != 2 && code_itemsize != 4) return -1;
  if (g > 64) return -2;
  if (nthreads <= 0) nthreads = aqlm_oracle_num_threads();
  dq_ctx c = {codes, code_itemsize, codebooks, scales, bias, x, batch, y, out_features, in_groups, num_codebooks, nbits, g};
  parallel_for(out_features, nthreads, gemm_rows, &c);
  return 0;
}

typedef struct {
  const float* x; const uint8_t* codes_alt; const float* codebooks; float* lut; float* partial;
  int64_t out_features, in_groups; int num_codebooks, g;
} lut_ctx;

static void lut_groups(int64_t begin, int64_t end, int tid, void* p) {
  const lut_ctx* c = (const lut_ctx*)p;
  const int cb_size = 256, K = c->num_codebooks, g = c->g;
  float* out = c->partial + (int64_t)tid * c->out_features;
  for (int64_t j = begin; j < end; ++j) {
    float* lj = c->lut + j * K * cb_size;
    for (int k = 0; k < K; ++k)            /* numba_kernel.py:39-40: lut = x_groups @ codebooks^T */
      for (int e2 = 0; e2 < cb_size; ++e2) {
        const float* v = c->codebooks + ((int64_t)k * cb_size + e2) * g;
        float d = 0.0f;
        for (int e = 0; e < g; ++e) d += c->x[j * g + e] * v[e];
        lj[k * cb_size + e2] = d;
      }
    const uint8_t* cj = c->codes_alt + j * c->out_features * K;
    for (int64_t i = 0; i < c->out_features; ++i) { /* numba_kernel.py:43-46 */
      float acc = 0.0f;
      for (int k = 0; k < K; ++k) acc += lj[k * cb_size + cj[i * K + k]];
      out[i] += acc;
    }
  }
}

/* numba_kernel.py:37-48 on the permuted layout codes_alt [in_groups, out, K] uint8 (inference.py:78-83):
 *   lut[j,c,k] = x_j . codebooks[c,k,:]     y[i] = scales[i] * sum_j sum_c lut[j,c,codes_alt[j,i,c]]
 * Threads split the in-groups (like numba.prange, numba_kernel.py:43) but each accumulates into a
 * private output vector that is reduced at the end -- the race-free version of the same loop nest. */
int aqlm_oracle_lut_gemv(const float* x, const uint8_t* codes_alt, const float* codebooks, const float* scales,
                         float* y, int64_t out_features, int64_t in_groups, int num_codebooks, int g,
                         int nthreads) {
  const int cb_size = 256;
  if (nthreads <= 0) nthreads = aqlm_oracle_num_threads();
  if (nthreads > in_groups) nthreads = (int)in_groups;
  float* lut = (float*)malloc(sizeof(float) * (size_t)in_groups * num_codebooks * cb_size);
  float* partial = (float*)calloc((size_t)nthreads * out_features, sizeof(float));
  if (!lut || !partial) { free(lut); free(partial); return -3; }
  lut_ctx c = {x, codes_alt, codebooks, lut, partial, out_features, in_groups, num_codebooks, g};
  parallel_for(in_groups, nthreads, lut_groups, &c);
  for (int64_t i = 0; i < out_features; ++i) {
    float acc = 0.0f;
    for (int t = 0; t < nthreads; ++t) acc += partial[(int64_t)t * out_features + i];
    y[i] = acc * scales[i]; /* numba_kernel.py:47 */
  }
  free(lut);
  free(partial);
  return 0;
}
