/*
 * plf_compress.cu -- site pattern compression on the device.
 *
 * Replaces pll_compress_site_patterns / pll_compress_site_patterns_msa
 * (reference src/compress.c:171-410): alignment columns are encoded with the
 * state map, sorted lexicographically (taxon 0 is the most significant
 * character), identical columns are merged and counted.  The reference sorts
 * column pointers with a randomised multikey quicksort on one core; the result
 * -- unique columns in ascending order, their weights, and the site -> pattern
 * map -- does not depend on the algorithm, so it is reproduced bit for bit by
 *   1. k_encode        characters -> state codes, first illegal character found
 *                      in the reference's (sequence, position) scan order;
 *   2. one stable LSD radix pass per taxon, last taxon first, over the site
 *      permutation only (k_hist / k_scan_hist / k_scatter: 256 bins, per-block
 *      histograms, stable ranks by warp match + per-digit running counters);
 *   3. k_heads         first column of each run of equal columns;
 *      exclusive scan  pattern index of every sorted position;
 *   4. k_emit          weights, site -> pattern map and the unique columns,
 *                      decoded back to characters.
 * Work is O(sites x taxa) bytes moved per pass over 4-byte indices; the
 * alignment itself is read through the permutation, never moved.
 */
#include "plf_backend.h"
#include "plf_device.cuh"
#include "plf_internal.h"

#define CMP_THREADS 256
#define CMP_CHUNK 4096 /* elements per block of a radix pass */

/* codes compare as signed chars in the reference (src/compress.c:55,66) */
__device__ __forceinline__ unsigned int sort_digit(unsigned char code) { return (unsigned int)(code ^ 0x80u); }

__global__ void k_encode(unsigned char * __restrict__ data, const unsigned char * __restrict__ charmap,
                         unsigned long long total, unsigned long long * __restrict__ first_bad)
{
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * blockDim.x)
  {
    const unsigned char c = charmap[data[i]];
    if (!c) atomicMin(first_bad, i);
    data[i] = c;
  }
}

__global__ void k_iota(unsigned int * __restrict__ perm, unsigned int n)
{
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) perm[i] = i;
}

/* hist[digit * nblocks + block] = number of elements of this block's chunk with that digit */
__global__ void __launch_bounds__(CMP_THREADS)
k_hist(const unsigned char * __restrict__ row, const unsigned int * __restrict__ perm, unsigned int n,
       unsigned int * __restrict__ hist)
{
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const unsigned int base = blockIdx.x * CMP_CHUNK;
  for (unsigned int i = base + threadIdx.x; i < base + CMP_CHUNK && i < n; i += CMP_THREADS)
    atomicAdd(&h[sort_digit(row[perm[i]])], 1u);
  __syncthreads();
  hist[threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

/* exclusive scan of `len` counters in place by one block (digit-major, block-minor order) */
__global__ void __launch_bounds__(1024)
k_scan_hist(unsigned int * __restrict__ v, unsigned int len)
{
  __shared__ unsigned int buf[1024];
  __shared__ unsigned int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (unsigned int base = 0; base < len; base += 1024)
  {
    const unsigned int i = base + threadIdx.x;
    const unsigned int x = i < len ? v[i] : 0;
    buf[threadIdx.x] = x;
    __syncthreads();
    for (unsigned int o = 1; o < 1024; o <<= 1)
    {
      const unsigned int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < len) v[i] = carry + buf[threadIdx.x] - x;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
}

/* stable scatter of this block's chunk: out[offset[digit][block] + rank among equal digits before it] */
__global__ void __launch_bounds__(CMP_THREADS)
k_scatter(const unsigned char * __restrict__ row, const unsigned int * __restrict__ perm, unsigned int n,
          const unsigned int * __restrict__ offsets, unsigned int * __restrict__ out)
{
  __shared__ unsigned int running[256];
  __shared__ unsigned int wcount[CMP_THREADS / 32][256];
  const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  running[threadIdx.x] = offsets[threadIdx.x * gridDim.x + blockIdx.x];
  const unsigned int base = blockIdx.x * CMP_CHUNK;
  for (unsigned int tile = 0; tile < CMP_CHUNK; tile += CMP_THREADS)
  {
    for (unsigned int w = 0; w < CMP_THREADS / 32; ++w) wcount[w][threadIdx.x] = 0;
    __syncthreads();
    const unsigned int i = base + tile + threadIdx.x;
    const bool valid = i < n;
    unsigned int p = 0, d = 0, rank = 0;
    const unsigned int vmask = __ballot_sync(0xffffffffu, valid);
    if (valid)
    {
      p = perm[i];
      d = sort_digit(row[p]);
      const unsigned int same = __match_any_sync(vmask, d);
      rank = __popc(same & ((1u << lane) - 1u));
      if (rank == 0) wcount[warp][d] = __popc(same);
    }
    __syncthreads();
    if (valid)
    {
      unsigned int before = running[d];
      for (unsigned int w = 0; w < warp; ++w) before += wcount[w][d];
      out[before + rank] = p;
    }
    __syncthreads();
    unsigned int add = 0;
    for (unsigned int w = 0; w < CMP_THREADS / 32; ++w) add += wcount[w][threadIdx.x];
    running[threadIdx.x] += add;
    __syncthreads();
    if (base + tile + CMP_THREADS >= n) break;
  }
}

/* head[i] = 1 when sorted column i differs from sorted column i-1 */
__global__ void k_heads(const unsigned char * __restrict__ data, const unsigned int * __restrict__ perm, unsigned int n,
                        unsigned int count, unsigned int * __restrict__ head)
{
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    unsigned int differs = (i == 0);
    if (i)
    {
      const unsigned int a = perm[i], b = perm[i - 1];
      for (unsigned int t = 0; t < count && !differs; ++t)
        differs = data[(size_t)t * n + a] != data[(size_t)t * n + b];
    }
    head[i] = differs;
  }
}

/* block sums -> scan -> inclusive positions: three small kernels over n flags */
__global__ void __launch_bounds__(CMP_THREADS)
k_block_sums(const unsigned int * __restrict__ flag, unsigned int n, unsigned int * __restrict__ sums)
{
  __shared__ unsigned int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  const unsigned int base = blockIdx.x * CMP_CHUNK;
  unsigned int mine = 0;
  for (unsigned int i = base + threadIdx.x; i < base + CMP_CHUNK && i < n; i += CMP_THREADS) mine += flag[i];
  atomicAdd(&s, mine);
  __syncthreads();
  if (threadIdx.x == 0) sums[blockIdx.x] = s;
}

/* ref[i] = (number of heads at positions <= i) - 1: pattern index of sorted position i */
__global__ void __launch_bounds__(CMP_THREADS)
k_pattern_index(const unsigned int * __restrict__ head, unsigned int n, const unsigned int * __restrict__ block_offset,
                unsigned int * __restrict__ ref)
{
  __shared__ unsigned int buf[CMP_THREADS];
  __shared__ unsigned int carry;
  if (threadIdx.x == 0) carry = block_offset[blockIdx.x];
  __syncthreads();
  const unsigned int base = blockIdx.x * CMP_CHUNK;
  for (unsigned int tile = 0; tile < CMP_CHUNK && base + tile < n; tile += CMP_THREADS)
  {
    const unsigned int i = base + tile + threadIdx.x;
    const unsigned int x = i < n ? head[i] : 0;
    buf[threadIdx.x] = x;
    __syncthreads();
    for (unsigned int o = 1; o < CMP_THREADS; o <<= 1)
    {
      const unsigned int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) ref[i] = carry + buf[threadIdx.x] - 1;
    __syncthreads();
    if (threadIdx.x == CMP_THREADS - 1) carry += buf[CMP_THREADS - 1];
    __syncthreads();
  }
}

__global__ void k_emit(const unsigned char * __restrict__ data, const unsigned int * __restrict__ perm,
                       const unsigned int * __restrict__ head, const unsigned int * __restrict__ ref, unsigned int n,
                       unsigned int count, const unsigned char * __restrict__ inv_charmap,
                       unsigned int * __restrict__ weight, unsigned int * __restrict__ site_pattern,
                       unsigned char * __restrict__ out /* [count][n] */)
{
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const unsigned int r = ref[i], p = perm[i];
    atomicAdd(&weight[r], 1u); /* integer: order-independent */
    site_pattern[p] = r;
    if (head[i])
      for (unsigned int t = 0; t < count; ++t) out[(size_t)t * n + r] = inv_charmap[data[(size_t)t * n + p]];
  }
}

/* h_rows: `count` host strings of `n` characters.  On success the unique columns are written back
 * into the first *compressed characters of every string (caller adds the terminating zero),
 * h_weight[0 .. *compressed) and, when given, h_site_pattern[0 .. n) are filled.
 * Returns 1, 0 on a CUDA failure, -1 when a character is not in the map (*bad_seq, *bad_pos, 0-based). */
extern "C" int plf_compress_patterns(plf_ctx_t * ctx, char ** h_rows, unsigned int count, unsigned int n,
                                     const unsigned char * h_charmap, const unsigned char * h_inv_charmap,
                                     unsigned int * h_weight, unsigned int * h_site_pattern, unsigned int * compressed,
                                     unsigned int * bad_seq, unsigned int * bad_pos)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  const unsigned long long total = (unsigned long long)count * n;
  const unsigned int nblocks = (n + CMP_CHUNK - 1) / CMP_CHUNK;
  unsigned char * d_data = (unsigned char *)plf_alloc(ctx, total, 0);
  unsigned char * d_out = (unsigned char *)plf_alloc(ctx, total, 0);
  unsigned char * d_maps = (unsigned char *)plf_alloc(ctx, 512, 0);
  unsigned int * d_perm = (unsigned int *)plf_alloc(ctx, (size_t)n * 4, 0);
  unsigned int * d_perm2 = (unsigned int *)plf_alloc(ctx, (size_t)n * 4, 0);
  unsigned int * d_hist = (unsigned int *)plf_alloc(ctx, (size_t)256 * nblocks * 4, 0);
  unsigned int * d_head = (unsigned int *)plf_alloc(ctx, (size_t)n * 4, 0);
  unsigned int * d_ref = (unsigned int *)plf_alloc(ctx, (size_t)n * 4, 0);
  unsigned int * d_weight = (unsigned int *)plf_alloc(ctx, (size_t)n * 4, 1);
  unsigned int * d_site = (unsigned int *)plf_alloc(ctx, (size_t)n * 4, 0);
  unsigned int * d_sums = (unsigned int *)plf_alloc(ctx, (size_t)(nblocks + 1) * 4, 0);
  unsigned long long * d_bad = (unsigned long long *)plf_alloc(ctx, 8, 0);
  int rc = 0;
  unsigned long long h_bad = ~0ull;
  unsigned int last_ref = 0;
  const unsigned int wide = (unsigned int)ctx->sm_count * 8;
  const unsigned int gb = (unsigned int)((total + CMP_THREADS - 1) / CMP_THREADS) < wide
                              ? (unsigned int)((total + CMP_THREADS - 1) / CMP_THREADS)
                              : wide;
  const unsigned int gn = (n + CMP_THREADS - 1) / CMP_THREADS < wide ? (n + CMP_THREADS - 1) / CMP_THREADS : wide;
  if (!d_data || !d_out || !d_maps || !d_perm || !d_perm2 || !d_hist || !d_head || !d_ref || !d_weight || !d_site ||
      !d_sums || !d_bad)
    goto done;
  for (unsigned int t = 0; t < count; ++t)
    if (!plf_upload(ctx, d_data + (size_t)t * n, h_rows[t], n)) goto done;
  if (!plf_upload(ctx, d_maps, h_charmap, 256) || !plf_upload(ctx, d_maps + 256, h_inv_charmap, 256) ||
      !plf_upload(ctx, d_bad, &h_bad, 8))
    goto done;
  k_encode<<<gb ? gb : 1, CMP_THREADS, 0, ctx->stream>>>(d_data, d_maps, total, d_bad);
  plf_count_launch();
  if (!plf_download(ctx, &h_bad, d_bad, 8)) goto done;
  if (h_bad != ~0ull)
  {
    *bad_seq = (unsigned int)(h_bad / n);
    *bad_pos = (unsigned int)(h_bad % n);
    rc = -1;
    goto done;
  }
  k_iota<<<gn ? gn : 1, CMP_THREADS, 0, ctx->stream>>>(d_perm, n);
  plf_count_launch();
  for (unsigned int t = count; t-- > 0;)
  {
    const unsigned char * row = d_data + (size_t)t * n;
    k_hist<<<nblocks, CMP_THREADS, 0, ctx->stream>>>(row, d_perm, n, d_hist);
    k_scan_hist<<<1, 1024, 0, ctx->stream>>>(d_hist, 256 * nblocks);
    k_scatter<<<nblocks, CMP_THREADS, 0, ctx->stream>>>(row, d_perm, n, d_hist, d_perm2);
    for (int i = 0; i < 3; ++i) plf_count_launch();
    unsigned int * tmp = d_perm;
    d_perm = d_perm2;
    d_perm2 = tmp;
  }
  k_heads<<<gn ? gn : 1, CMP_THREADS, 0, ctx->stream>>>(d_data, d_perm, n, count, d_head);
  k_block_sums<<<nblocks, CMP_THREADS, 0, ctx->stream>>>(d_head, n, d_sums);
  k_scan_hist<<<1, 1024, 0, ctx->stream>>>(d_sums, nblocks);
  k_pattern_index<<<nblocks, CMP_THREADS, 0, ctx->stream>>>(d_head, n, d_sums, d_ref);
  k_emit<<<gn ? gn : 1, CMP_THREADS, 0, ctx->stream>>>(d_data, d_perm, d_head, d_ref, n, count, d_maps + 256, d_weight,
                                                       d_site, d_out);
  for (int i = 0; i < 5; ++i) plf_count_launch();
  if (cudaGetLastError() != cudaSuccess) goto done;
  if (!plf_download(ctx, &last_ref, d_ref + (n - 1), 4)) goto done;
  *compressed = last_ref + 1;
  for (unsigned int t = 0; t < count; ++t)
    if (!plf_download(ctx, h_rows[t], d_out + (size_t)t * n, *compressed)) goto done;
  if (!plf_download(ctx, h_weight, d_weight, (size_t)*compressed * 4)) goto done;
  if (h_site_pattern && !plf_download(ctx, h_site_pattern, d_site, (size_t)n * 4)) goto done;
  rc = 1;
done:
  if (rc == 0 && !ctx->err[0]) plf_set_error(ctx, "pattern compression failed: %s", cudaGetErrorString(cudaGetLastError()));
  plf_free(ctx, d_data);
  plf_free(ctx, d_out);
  plf_free(ctx, d_maps);
  plf_free(ctx, d_perm);
  plf_free(ctx, d_perm2);
  plf_free(ctx, d_hist);
  plf_free(ctx, d_head);
  plf_free(ctx, d_ref);
  plf_free(ctx, d_weight);
  plf_free(ctx, d_site);
  plf_free(ctx, d_sums);
  plf_free(ctx, d_bad);
  return rc;
}
