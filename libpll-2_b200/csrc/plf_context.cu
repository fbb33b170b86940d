/*
 * plf_context.cu -- device context, memory plumbing and the P-matrix kernel.
 */
#include "plf_backend.h"
#include "plf_device.cuh"
#include "plf_internal.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread unsigned long long g_launches = 0;
void plf_count_launch(void) { ++g_launches; }
void plf_count_launches(unsigned long long n) { g_launches += n; }
extern "C" unsigned long long plf_kernel_launches(void) { return g_launches; }

void plf_set_error(plf_ctx * ctx, const char * fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
  va_end(ap);
}

extern "C" const char * plf_last_error(const plf_ctx_t * ctx) { return ctx->err; }
extern "C" int plf_ctx_device(const plf_ctx_t * ctx) { return ctx->device; }
extern "C" void * plf_ctx_stream(const plf_ctx_t * ctx) { return (void *)ctx->stream; }

extern "C" int plf_device_count(char * err, size_t errlen)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess)
  {
    if (err) snprintf(err, errlen, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return 0;
  }
  return n;
}

extern "C" int plf_ctx_create(int device, int managed, plf_ctx_t ** out, char * err, size_t errlen)
{
  *out = NULL;
  int n = plf_device_count(err, errlen);
  if (n <= 0)
  {
    if (err && !err[0]) snprintf(err, errlen, "no CUDA device is visible (there is no CPU fallback)");
    return 0;
  }
  if (device < 0 || device >= n)
  {
    if (err) snprintf(err, errlen, "CUDA device %d requested, %d visible", device, n);
    return 0;
  }
  plf_ctx * ctx = (plf_ctx *)calloc(1, sizeof(plf_ctx));
  if (!ctx) return 0;
  ctx->device = device;
  ctx->managed = managed;
  ctx->dna_stream = -1;
  ctx->dna_level_max_sites = -1;
  ctx->dna_flow = -1;
  ctx->graph_mode = -1;
  ctx->aa_stream = -1;
  ctx->aam_log2r[0] = ctx->aam_log2r[1] = -1;
  {
    const char * v = getenv("PLF_AA_FAST");
    ctx->aa_fast = !(v && v[0] == '0');
    v = getenv("PLF_AA_MMA");
    ctx->aa_mma = !(v && v[0] == '0');
    v = getenv("PLF_EDGE_FAST");
    ctx->edge_fast = !(v && v[0] == '0');
    v = getenv("PLL_CUDA_GUARD");
    ctx->guard = (v && v[0] && v[0] != '0');
  }
  cudaError_t e = cudaSetDevice(device);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess && prop.major < 10)
  {
    if (err)
      snprintf(err, errlen, "device %d (%s, sm_%d%d) is not Blackwell: this library is built for sm_100a only",
               device, prop.name, prop.major, prop.minor);
    free(ctx);
    return 0;
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess && !managed)
  {
    /* stream-ordered pool private to this context: site-repeat reallocation
     * (CLVs, scalers resized to the class count) never synchronises and freed
     * blocks are recycled without a trip to the driver */
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    e = cudaMemPoolCreate(&ctx->pool, &props);
    if (e == cudaSuccess)
    {
      unsigned long long keep = ~0ull;
      e = cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_result, 4 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->d_ticket, 64);
  if (e == cudaSuccess) e = cudaMemset(ctx->d_ticket, 0, 64);
  if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_result, 4 * sizeof(double));
  if (e != cudaSuccess)
  {
    if (err) snprintf(err, errlen, "CUDA context setup failed: %s", cudaGetErrorString(e));
    free(ctx);
    return 0;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  snprintf(ctx->name, sizeof(ctx->name), "%s sm_%d%d, %d SMs, %.1f GB", prop.name, prop.major, prop.minor,
           prop.multiProcessorCount, (double)prop.totalGlobalMem / 1e9);
  *out = ctx;
  return 1;
}

extern "C" void plf_device_description(const plf_ctx_t * ctx, char * buf, size_t len)
{
  snprintf(buf, len, "%s", ctx->name);
}

extern "C" void plf_ctx_destroy(plf_ctx_t * ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  plf_graph_cache_destroy(ctx);
  free(ctx->guard_recs);
  cudaFree(ctx->ws_ops.ptr);
  cudaFree(ctx->ws_once.ptr);
  cudaFree(ctx->ws_flow.ptr);
  cudaFree(ctx->ws_small.ptr);
  cudaFree(ctx->ws_partial.ptr);
  cudaFree(ctx->d_result);
  cudaFree(ctx->d_ticket);
  cudaFree(ctx->ws_edge.ptr);
  cudaFreeHost(ctx->h_result);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  cudaStreamDestroy(ctx->stream);
  free(ctx);
}

void * plf_ws_reserve(plf_ctx * ctx, plf_ws * ws, size_t bytes)
{
  if (ws->bytes >= bytes && ws->ptr) return ws->ptr;
  if (ws->ptr)
  {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ws->ptr);
    ws->ptr = NULL;
    ws->bytes = 0;
  }
  size_t want = bytes < 4096 ? 4096 : bytes + bytes / 2;
  cudaError_t e = cudaMalloc(&ws->ptr, want);
  if (e != cudaSuccess)
  {
    plf_set_error(ctx, "workspace cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    ws->ptr = NULL;
    return NULL;
  }
  ws->bytes = want;
  return ws->ptr;
}

/* ---- guarded allocations ($PLL_CUDA_GUARD=1) -------------------------------------------------------- *
 * compute-sanitizer is not available on every pool.  In guard mode every device buffer handed out by        *
 * plf_alloc sits between two PLF_GUARD-byte bands of 0xA5; plf_check_guards() reads all bands back and      *
 * counts the allocations whose bands were written to.  Catches out-of-bounds WRITES of the kernels (bulk    *
 * stores, scaler / identifier / pair-list stores, table scatters) on the small and odd shapes of the        *
 * parity suite; the 16 bytes of by-design slack behind a buffer lie inside the allocation, not the band.   */
#define PLF_GUARD 256

static void guard_register(plf_ctx * ctx, void * user, size_t bytes)
{
  if (ctx->guard_count == ctx->guard_cap)
  {
    const size_t cap = ctx->guard_cap ? 2 * ctx->guard_cap : 1024;
    plf_guard_rec * g = (plf_guard_rec *)realloc(ctx->guard_recs, cap * sizeof(plf_guard_rec));
    if (!g) return;
    ctx->guard_recs = g;
    ctx->guard_cap = cap;
  }
  ctx->guard_recs[ctx->guard_count].user = user;
  ctx->guard_recs[ctx->guard_count].bytes = bytes;
  ++ctx->guard_count;
}

extern "C" void * plf_alloc(plf_ctx_t * ctx, size_t bytes, int zero)
{
  void * p = NULL;
  /* every buffer carries >= 16 bytes of slack: the bulk-copy kernels round the
   * last tile of 1- and 4-byte arrays (tip codes, scalers, weights, invariant
   * flags) up to the 16-byte copy granule */
  bytes = (bytes + 16 + 255) & ~(size_t)255;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return NULL;
  const size_t total = bytes + (ctx->guard ? 2 * PLF_GUARD : 0);
  cudaError_t e;
  if (ctx->managed)
  {
    e = cudaMallocManaged(&p, total, cudaMemAttachGlobal);
    if (e == cudaSuccess) cudaMemPrefetchAsync(p, total, ctx->device, ctx->stream);
  }
  else
    e = cudaMallocFromPoolAsync(&p, total, ctx->pool, ctx->stream);
  if (e != cudaSuccess)
  {
    cudaGetLastError();
    plf_set_error(ctx, "device allocation of %zu bytes failed: %s", total, cudaGetErrorString(e));
    return NULL;
  }
  if (ctx->guard)
  {
    cudaMemsetAsync(p, 0xA5, PLF_GUARD, ctx->stream);
    cudaMemsetAsync((char *)p + PLF_GUARD + bytes, 0xA5, PLF_GUARD, ctx->stream);
    p = (char *)p + PLF_GUARD;
    guard_register(ctx, p, bytes);
  }
  if (zero) cudaMemsetAsync(p, 0, bytes, ctx->stream);
  return p;
}

/* number of live allocations whose guard bands were written to (-1: guard mode is off); synchronises */
extern "C" int plf_check_guards(plf_ctx_t * ctx)
{
  if (!ctx->guard) return -1;
  if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -2;
  int damaged = 0;
  unsigned char band[2 * PLF_GUARD];
  for (size_t i = 0; i < ctx->guard_count; ++i)
  {
    const plf_guard_rec & g = ctx->guard_recs[i];
    if (cudaMemcpy(band, (char *)g.user - PLF_GUARD, PLF_GUARD, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(band + PLF_GUARD, (char *)g.user + g.bytes, PLF_GUARD, cudaMemcpyDeviceToHost) != cudaSuccess)
      return -2;
    int bad = 0;
    for (int k = 0; k < 2 * PLF_GUARD && !bad; ++k) bad = band[k] != 0xA5;
    if (bad)
    {
      ++damaged;
      plf_set_error(ctx, "guard band of the %zu-byte allocation at %p was written to", g.bytes, g.user);
    }
  }
  return damaged;
}

/* Grows the stream-ordered pool by `bytes` (capped at 40 % of the free device memory) and hands the block
 * straight back: it stays cached in the pool (release threshold = max), so the many small allocations
 * that follow are carved out of it without a trip to the driver. */
extern "C" int plf_pool_reserve(plf_ctx_t * ctx, size_t bytes)
{
  if (!ctx->pool || !bytes) return 1;
  size_t free_b = 0, total_b = 0;
  if (cudaSetDevice(ctx->device) != cudaSuccess || cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 1;
  if (bytes > free_b / 5 * 2) bytes = free_b / 5 * 2;
  void * p = NULL;
  if (cudaMallocFromPoolAsync(&p, bytes, ctx->pool, ctx->stream) != cudaSuccess)
  {
    cudaGetLastError(); /* a warm-up only: failure is not an error */
    return 1;
  }
  cudaFreeAsync(p, ctx->stream);
  return 1;
}

extern "C" void plf_free(plf_ctx_t * ctx, void * p)
{
  if (!p) return;
  cudaSetDevice(ctx->device);
  if (ctx->guard)
  {
    /* forget the record (a damaged band would have been reported by the last plf_check_guards) */
    for (size_t i = 0; i < ctx->guard_count; ++i)
      if (ctx->guard_recs[i].user == p)
      {
        ctx->guard_recs[i] = ctx->guard_recs[--ctx->guard_count];
        break;
      }
    p = (char *)p - PLF_GUARD;
  }
  if (ctx->pool)
  {
    /* stream-ordered: every kernel queued so far that uses `p` finishes first */
    cudaFreeAsync(p, ctx->stream);
    return;
  }
  cudaStreamSynchronize(ctx->stream);
  cudaFree(p);
}

extern "C" int plf_upload(plf_ctx_t * ctx, void * dst, const void * src, size_t bytes)
{
  if (!bytes) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  PLF_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  /* the source may be a caller buffer that is reused right away: pageable
   * copies are staged by the driver before returning, pinned ones are not */
  PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 1;
}

extern "C" int plf_upload_async(plf_ctx_t * ctx, void * dst, const void * src, size_t bytes)
{
  if (!bytes) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  PLF_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return 1;
}

extern "C" int plf_download(plf_ctx_t * ctx, void * dst, const void * src, size_t bytes)
{
  if (!bytes) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  PLF_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 1;
}

/* page-locked host staging (truly asynchronous H2D copies) */
extern "C" void * plf_pinned_alloc(plf_ctx_t * ctx, size_t bytes)
{
  void * p = NULL;
  if (cudaSetDevice(ctx->device) != cudaSuccess || cudaMallocHost(&p, bytes) != cudaSuccess)
  {
    cudaGetLastError();
    plf_set_error(ctx, "pinned host allocation of %zu bytes failed", bytes);
    return NULL;
  }
  return p;
}

extern "C" void plf_pinned_free(plf_ctx_t * ctx, void * p)
{
  if (!p) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaFreeHost(p);
}

extern "C" int plf_memset0(plf_ctx_t * ctx, void * dst, size_t bytes)
{
  if (!bytes) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  PLF_CHECK(ctx, cudaMemsetAsync(dst, 0, bytes, ctx->stream));
  return 1;
}

extern "C" int plf_sync(plf_ctx_t * ctx)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 1;
}

/* ------------------------------------------------------------------------ *
 *  P-matrices: one block per (matrix, rate category).                       *
 *  Reference: pll_core_update_pmatrix (src/core_pmatrix.c:24), 4x4 AVX       *
 *  (src/core_pmatrix_avx.c:42) and 20x20 AVX2 (src/core_pmatrix_avx2.c:49)   *
 *  evaluation orders, see oracle/plf_oracle.c:orc_update_pmatrix.            *
 * ------------------------------------------------------------------------ */
__global__ void k_pmatrix(double * __restrict__ pbase, const double * __restrict__ model,
                          const unsigned int * __restrict__ matrix_indices,
                          const double * __restrict__ branch_lengths,
                          const double * __restrict__ expd_host, int st, int sp, int R)
{
  extern __shared__ double sm[];
  double * e = sm;            /* [sp]      expm1 values       */
  double * tmp = sm + sp;     /* [st][st]  Vinv * diag(e)     */
  const int i = blockIdx.x, n = blockIdx.y;
  const double t = branch_lengths[i];
  double * pmat = pbase + ((size_t)matrix_indices[i] * R + n) * st * sp;
  const double rate = model[n];
  const double pinv = model[2 * R + n];
  const double * evals = model + 3 * R + (size_t)R * sp + (size_t)n * sp;
  const double * evecs = model + 3 * R + (size_t)2 * R * sp + (size_t)n * st * sp;
  const double * ievecs = evecs + (size_t)R * st * sp;
  const bool fixed = (st == 4 || st == 20);

  if (!(t > 0.0))
  {
    /* identity (core_pmatrix.c:243-248; the 4/20 kernels also clear padding) */
    const int cols = fixed ? sp : st;
    for (int x = threadIdx.x; x < st * cols; x += blockDim.x)
    {
      const int j = x / cols, k = x % cols;
      pmat[j * sp + k] = (j == k) ? 1.0 : 0.0;
    }
    return;
  }
  for (int j = threadIdx.x; j < st; j += blockDim.x)
  {
    if (expd_host)
      e[j] = expd_host[((size_t)i * R + n) * st + j];
    else
    {
      double x = (evals[j] * rate) * t;
      if (pinv > 1e-8) x = x / (1.0 - pinv);
      e[j] = expm1(x);
    }
  }
  __syncthreads();
  for (int x = threadIdx.x; x < st * st; x += blockDim.x)
  {
    const int j = x / st, m = x % st;
    tmp[x] = (st == 20) ? e[m] * ievecs[j * sp + m] : ievecs[j * sp + m] * e[m];
  }
  __syncthreads();
  for (int x = threadIdx.x; x < st * st; x += blockDim.x)
  {
    const int j = x / st, k = x % st;
    const double * tj = tmp + j * st;
    double v;
    if (st == 4)
    {
      const double p0 = tj[0] * evecs[0 * 4 + k], p1 = tj[1] * evecs[1 * 4 + k];
      const double p2 = tj[2] * evecs[2 * 4 + k], p3 = tj[3] * evecs[3 * 4 + k];
      v = ((p0 + p1) + (p2 + p3)) + ((j == k) ? 1.0 : 0.0);
    }
    else if (st == 20)
    {
      double a[4];
#pragma unroll
      for (int l = 0; l < 4; ++l)
      {
        a[l] = tj[l] * evecs[l * 20 + k];
        for (int q = 1; q < 5; ++q) a[l] = fma(tj[l + 4 * q], evecs[(l + 4 * q) * 20 + k], a[l]);
      }
      v = (a[0] + a[1]) + (a[2] + a[3]);
      if (j == k) v += 1.0;
    }
    else
    {
      v = (j == k) ? 1.0 : 0.0;
      for (int m = 0; m < st; ++m) v += tj[m] * evecs[m * sp + k];
    }
    pmat[j * sp + k] = v;
  }
}

extern "C" int plf_update_pmatrices(plf_ctx_t * ctx, const plf_shape_t * sh, const double * d_model,
                                    double * d_pmatrix_base, const unsigned int * h_matrix_indices,
                                    const double * h_branch_lengths, unsigned int count,
                                    const double * h_expd)
{
  if (!count) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  const unsigned int R = sh->rate_cats, st = sh->states, sp = sh->states_padded;
  const size_t nb_idx = ((size_t)count * sizeof(unsigned int) + 15) & ~(size_t)15;
  const size_t nb_bl = (size_t)count * sizeof(double);
  const size_t nb_ex = h_expd ? (size_t)count * R * st * sizeof(double) : 0;
  char * ws = (char *)plf_ws_reserve(ctx, &ctx->ws_small, nb_idx + nb_bl + nb_ex);
  if (!ws) return 0;
  {
    /* one host-to-device copy instead of three: on narrow alignments the call is its enqueue cost
     * (pageable source: staged by the driver before the call returns) */
    char * h = (char *)malloc(nb_idx + nb_bl + nb_ex);
    if (!h)
    {
      plf_set_error(ctx, "P-matrix arguments: out of host memory (%zu B)", nb_idx + nb_bl + nb_ex);
      return 0;
    }
    memcpy(h, h_matrix_indices, count * sizeof(unsigned int));
    memcpy(h + nb_idx, h_branch_lengths, nb_bl);
    if (h_expd) memcpy(h + nb_idx + nb_bl, h_expd, nb_ex);
    const cudaError_t e = cudaMemcpyAsync(ws, h, nb_idx + nb_bl + nb_ex, cudaMemcpyHostToDevice, ctx->stream);
    free(h);
    PLF_CHECK(ctx, e);
  }
  dim3 grid(count, R);
  const int threads = st * st >= 256 ? 256 : (st * st >= 64 ? 128 : 32);
  const size_t smem = ((size_t)sp + (size_t)st * st) * sizeof(double);
  k_pmatrix<<<grid, threads, smem, ctx->stream>>>(d_pmatrix_base, d_model, (const unsigned int *)ws,
                                                 (const double *)(ws + nb_idx),
                                                 h_expd ? (const double *)(ws + nb_idx + nb_bl) : NULL, (int)st,
                                                 (int)sp, (int)R);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}
