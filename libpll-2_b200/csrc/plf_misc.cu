/*
 * plf_misc.cu -- small device passes around the hot path:
 *   - tip CLVs from state masks            (reference set_tipclv, src/pll.c:959-1024)
 *   - invariant-site detection             (src/models.c:651-752)
 *   - site-repeat class identifiers        (src/repeats.c:299-382), bit-exact
 */
#include "plf_backend.h"
#include "plf_device.cuh"
#include "plf_internal.h"

/* ---- tip CLVs ------------------------------------------------------------ */
__global__ void k_tip_clv(double * __restrict__ clv, const unsigned char * __restrict__ seq,
                          const plf_state_t * __restrict__ map, const unsigned int * __restrict__ id_site,
                          unsigned int entries, int st, int sp, int R)
{
  const size_t total = (size_t)entries * R * sp;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x)
  {
    const unsigned int n = (unsigned int)(x / ((size_t)R * sp));
    const int j = (int)(x % sp);
    const unsigned int site = id_site ? id_site[n] : n;
    const plf_state_t m = map[seq[site]];
    clv[x] = (j < st && ((m >> j) & 1ull)) ? 1.0 : 0.0;
  }
}

extern "C" int plf_tip_clv_from_states(plf_ctx_t * ctx, const plf_shape_t * sh, double * d_clv,
                                       const unsigned char * d_seq, const unsigned long long * d_map,
                                       const unsigned int * d_id_site, unsigned int entries)
{
  if (!entries) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  const size_t total = (size_t)entries * sh->rate_cats * sh->states_padded;
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 16) blocks = (size_t)ctx->sm_count * 16;
  k_tip_clv<<<(unsigned int)blocks, 256, 0, ctx->stream>>>(d_clv, d_seq, d_map, d_id_site, entries,
                                                          (int)sh->states, (int)sh->states_padded,
                                                          (int)sh->rate_cats);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* ---- pattern-tip codes from raw characters (src/pll.c:875-957) -------------- *
 * out[s] = low byte of lut[seq[s]]; bit 8 of a lut entry marks a character the caller's map does not know:  *
 * the smallest such site goes to *first_bad.  16 characters per thread and step.                            */
__global__ void k_tip_map(const unsigned char * __restrict__ seq, const unsigned short * __restrict__ lut,
                          unsigned int sites, unsigned char * __restrict__ out, unsigned int * __restrict__ first_bad)
{
  __shared__ unsigned short s_lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const unsigned int chunks = (sites + 15) >> 4; /* both buffers carry >= 16 bytes of slack */
  unsigned int bad = 0xFFFFFFFFu;
  for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += gridDim.x * blockDim.x)
  {
    const uint4 in = reinterpret_cast<const uint4 *>(seq)[c];
    const unsigned int w[4] = {in.x, in.y, in.z, in.w};
    unsigned int o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
      unsigned int acc = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b)
      {
        const unsigned int v = s_lut[(w[k] >> (8 * b)) & 255u];
        const unsigned int site = c * 16 + k * 4 + b;
        if ((v & 0x100u) && site < sites && site < bad) bad = site;
        acc |= (v & 255u) << (8 * b);
      }
      o[k] = acc;
    }
    reinterpret_cast<uint4 *>(out)[c] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  if (bad != 0xFFFFFFFFu) atomicMin(first_bad, bad);
}

extern "C" int plf_tip_map(plf_ctx_t * ctx, const unsigned char * d_seq, const unsigned short * d_lut,
                           unsigned int sites, unsigned char * d_out, unsigned int * d_first_bad)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  PLF_CHECK(ctx, cudaMemsetAsync(d_first_bad, 0xFF, sizeof(unsigned int), ctx->stream));
  unsigned int blocks = ((sites + 15) / 16 + 255) / 256;
  if (blocks > (unsigned int)ctx->sm_count * 8) blocks = (unsigned int)ctx->sm_count * 8;
  if (blocks < 1) blocks = 1;
  k_tip_map<<<blocks, 256, 0, ctx->stream>>>(d_seq, d_lut, sites, d_out, d_first_bad);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* ---- invariant sites ------------------------------------------------------ */
__global__ void k_invariant(int * __restrict__ out, unsigned int sites, unsigned int tips,
                            const unsigned char * const * __restrict__ tipchars,
                            const double * const * __restrict__ tipclv,
                            const unsigned int * const * __restrict__ tip_site_id,
                            const plf_state_t * __restrict__ tipmap, int st, int sp, int R)
{
  for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < sites; s += gridDim.x * blockDim.x)
  {
    plf_state_t acc = (st >= 64) ? ~0ull : ((1ull << st) - 1ull);
    for (unsigned int t = 0; t < tips && acc; ++t)
    {
      plf_state_t m = 0;
      if (tipchars)
      {
        const unsigned int c = tipchars[t][s];
        m = (st == 4) ? (plf_state_t)c : tipmap[c];
      }
      else
      {
        const unsigned int id = (tip_site_id && tip_site_id[t]) ? tip_site_id[t][s] : s;
        const double * c = tipclv[t] + (size_t)id * R * sp;
        for (int k = 0; k < st; ++k) m |= ((plf_state_t)c[k]) << k;
      }
      acc &= m;
    }
    out[s] = (acc == 0 || __popcll(acc) > 1) ? -1 : (__ffsll((long long)acc) - 1);
  }
}

extern "C" int plf_invariant_sites(plf_ctx_t * ctx, const plf_shape_t * sh, unsigned int sites, unsigned int tips,
                                   const unsigned char * const * d_tipchars, const double * const * d_tipclv,
                                   const unsigned int * const * d_tip_site_id,
                                   const unsigned long long * d_tipmap, int * d_out)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  unsigned int blocks = (sites + 255) / 256;
  if (blocks > (unsigned int)ctx->sm_count * 8) blocks = (unsigned int)ctx->sm_count * 8;
  k_invariant<<<blocks, 256, 0, ctx->stream>>>(d_out, sites, tips, d_tipchars, d_tipclv, d_tip_site_id, d_tipmap,
                                              (int)sh->states, (int)sh->states_padded, (int)sh->rate_cats);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* ---- site-repeat identifiers ----------------------------------------------- *
 * The reference numbers the classes of a parent node by order of FIRST
 * OCCURRENCE while scanning sites 0..S-1 of the key
 *     key[s] = id_left[s] + id_right[s] * ids_left
 * through a dense lookup table (repeats.c:334-347).  Deterministic parallel
 * equivalent:
 *   1. lookup[key[s]] = min over s (atomicMin)        -> first site of each class
 *   2. first[s] = (lookup[key[s]] == s); exclusive scan of first[] -> rank
 *   3. site_id[s] = rank[lookup[key[s]]]; id_site[rank[s]] = s where first[s]
 *   4. lookup[key[s]] = EMPTY again
 * The scan is a three-kernel block scan over 1024-site tiles.                  */

#define SCAN_TILE 1024
/* idr == NULL: the key is idl[s] itself (tip class codes) */
#define REP_KEY(s) (idr ? idl[(s)] + idr[(s)] * ids_left : idl[(s)])
/* blockIdx.y selects the job: independent nodes of one traversal level are
 * numbered by the same six launches, each in its own slice of the lookup pool */
#define REP_JOB                                                     \
  const plf_rep_job_t jb = jobs[blockIdx.y];                        \
  const unsigned int * __restrict__ idl = jb.site_id_left;          \
  const unsigned int * __restrict__ idr = jb.site_id_right;         \
  const unsigned int ids_left = jb.ids_left;                        \
  unsigned int * __restrict__ lookup = lookup_pool + jb.lookup_offset

__global__ void k_rep_min(const plf_rep_job_t * __restrict__ jobs, unsigned int sites,
                          unsigned int * __restrict__ lookup_pool)
{
  REP_JOB;
  for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < sites; s += gridDim.x * blockDim.x)
    atomicMin(&lookup[REP_KEY(s)], s);
}

/* per-tile count of first occurrences */
__global__ void k_rep_count(const plf_rep_job_t * __restrict__ jobs, unsigned int sites,
                            unsigned int * __restrict__ lookup_pool, unsigned int * __restrict__ tile_count_all,
                            unsigned int ntiles)
{
  REP_JOB;
  unsigned int * tile_count = tile_count_all + (size_t)blockIdx.y * ntiles;
  __shared__ unsigned int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const unsigned int base = blockIdx.x * SCAN_TILE;
  unsigned int mine = 0;
  for (unsigned int s = base + threadIdx.x; s < base + SCAN_TILE && s < sites; s += blockDim.x)
    mine += (lookup[REP_KEY(s)] == s);
  atomicAdd(&cnt, mine); /* integer: order-independent */
  __syncthreads();
  if (threadIdx.x == 0) tile_count[blockIdx.x] = cnt;
}

/* exclusive scan of one job's tile counts by one block; total to totals[job] */
__global__ void k_rep_scan_tiles(unsigned int * __restrict__ tile_count_all, unsigned int ntiles,
                                 unsigned int * __restrict__ totals)
{
  unsigned int * tile_count = tile_count_all + (size_t)blockIdx.y * ntiles;
  __shared__ unsigned int buf[1024];
  __shared__ unsigned int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (unsigned int base = 0; base < ntiles; base += 1024)
  {
    const unsigned int i = base + threadIdx.x;
    const unsigned int v = i < ntiles ? tile_count[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (unsigned int o = 1; o < 1024; o <<= 1)
    {
      unsigned int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < ntiles) tile_count[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.y] = carry;
}

/* rank of each first occurrence: rank_of_site[s] for first occurrences, and
 * id_site[rank] = s */
__global__ void k_rep_rank(const plf_rep_job_t * __restrict__ jobs, unsigned int sites,
                           unsigned int * __restrict__ lookup_pool, const unsigned int * __restrict__ tile_count_all,
                           unsigned int ntiles, unsigned int * __restrict__ rank_all)
{
  REP_JOB;
  const unsigned int * tile_offset = tile_count_all + (size_t)blockIdx.y * ntiles;
  unsigned int * rank_of_site = rank_all + (size_t)blockIdx.y * sites;
  __shared__ unsigned int buf[SCAN_TILE];
  const unsigned int s = blockIdx.x * SCAN_TILE + threadIdx.x;
  const unsigned int f = (s < sites) ? (lookup[REP_KEY(s)] == s) : 0u;
  buf[threadIdx.x] = f;
  __syncthreads();
  for (unsigned int o = 1; o < SCAN_TILE; o <<= 1)
  {
    unsigned int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
    __syncthreads();
    buf[threadIdx.x] += t;
    __syncthreads();
  }
  if (f)
  {
    const unsigned int r = tile_offset[blockIdx.x] + buf[threadIdx.x] - 1;
    rank_of_site[s] = r;
    jb.id_site_parent[r] = s;
  }
}

/* site_id[s] = rank of the class of s; the lookup entries go back to EMPTY */
__global__ void k_rep_assign(const plf_rep_job_t * __restrict__ jobs, unsigned int sites,
                             unsigned int * __restrict__ lookup_pool, const unsigned int * __restrict__ rank_all)
{
  REP_JOB;
  const unsigned int * rank_of_site = rank_all + (size_t)blockIdx.y * sites;
  for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < sites; s += gridDim.x * blockDim.x)
    jb.site_id_parent[s] = rank_of_site[lookup[REP_KEY(s)]];
}

__global__ void k_rep_clean(const plf_rep_job_t * __restrict__ jobs, unsigned int sites,
                            unsigned int * __restrict__ lookup_pool)
{
  REP_JOB;
  for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < sites; s += gridDim.x * blockDim.x)
    lookup[REP_KEY(s)] = 0xFFFFFFFFu;
}

extern "C" size_t plf_repeats_batch_workspace(unsigned int sites, unsigned int njobs)
{
  const size_t ntiles = ((size_t)sites + SCAN_TILE - 1) / SCAN_TILE;
  return (size_t)njobs * (sizeof(plf_rep_job_t) + (ntiles + 1 + sites) * sizeof(unsigned int)) + 64;
}

/* class identifiers of `njobs` independent parent nodes: six launches and ONE
 * host synchronisation for the whole batch; class counts to h_ids[njobs] */
extern "C" int plf_repeats_ids_batch(plf_ctx_t * ctx, unsigned int sites, const plf_rep_job_t * h_jobs,
                                     unsigned int njobs, unsigned int * d_lookup_pool, unsigned int * h_ids)
{
  if (!njobs) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  const unsigned int ntiles = (sites + SCAN_TILE - 1) / SCAN_TILE;
  /* workspace: jobs [njobs] | totals [njobs] | tile counts [njobs][ntiles] | rank_of_site [njobs][sites] */
  const size_t jobs_bytes = ((size_t)njobs * sizeof(plf_rep_job_t) + 15) & ~(size_t)15;
  char * ws = (char *)plf_ws_reserve(ctx, &ctx->ws_partial,
                                     jobs_bytes + ((size_t)njobs * (1 + (size_t)ntiles + sites)) * sizeof(unsigned int));
  if (!ws) return 0;
  plf_rep_job_t * d_jobs = (plf_rep_job_t *)ws;
  unsigned int * totals = (unsigned int *)(ws + jobs_bytes);
  unsigned int * tile = totals + njobs;
  unsigned int * rank = tile + (size_t)njobs * ntiles;
  PLF_CHECK(ctx, cudaMemcpyAsync(d_jobs, h_jobs, (size_t)njobs * sizeof(plf_rep_job_t), cudaMemcpyHostToDevice,
                                 ctx->stream));
  unsigned int blocks = (sites + 255) / 256;
  const unsigned int cap = (unsigned int)ctx->sm_count * 8;
  const unsigned int share = cap / njobs > 4 ? cap / njobs : 4; /* one wave over the whole batch */
  if (blocks > share) blocks = share;
  const dim3 gs(blocks, njobs), gt(ntiles, njobs);
  k_rep_min<<<gs, 256, 0, ctx->stream>>>(d_jobs, sites, d_lookup_pool);
  k_rep_count<<<gt, 256, 0, ctx->stream>>>(d_jobs, sites, d_lookup_pool, tile, ntiles);
  k_rep_scan_tiles<<<dim3(1, njobs), 1024, 0, ctx->stream>>>(tile, ntiles, totals);
  k_rep_rank<<<gt, SCAN_TILE, 0, ctx->stream>>>(d_jobs, sites, d_lookup_pool, tile, ntiles, rank);
  k_rep_assign<<<gs, 256, 0, ctx->stream>>>(d_jobs, sites, d_lookup_pool, rank);
  k_rep_clean<<<gs, 256, 0, ctx->stream>>>(d_jobs, sites, d_lookup_pool);
  for (int i = 0; i < 6; ++i) plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  PLF_CHECK(ctx, cudaMemcpyAsync(h_ids, totals, (size_t)njobs * sizeof(unsigned int), cudaMemcpyDeviceToHost,
                                 ctx->stream));
  PLF_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return 1;
}

extern "C" int plf_repeats_ids(plf_ctx_t * ctx, unsigned int sites, const unsigned int * d_idl,
                               unsigned int ids_left, const unsigned int * d_idr, unsigned int * d_site_id,
                               unsigned int * d_id_site, unsigned int * d_lookup, unsigned int * h_ids)
{
  plf_rep_job_t job;
  job.site_id_left = d_idl;
  job.site_id_right = d_idr;
  job.site_id_parent = d_site_id;
  job.id_site_parent = d_id_site;
  job.ids_left = ids_left;
  job.lookup_offset = 0;
  return plf_repeats_ids_batch(ctx, sites, &job, 1, d_lookup, h_ids);
}

/* ---- identifiers of a whole operation list, no host round trip per level ------------------------- *
 * Same numbering as above (first occurrence order), with                                              *
 *   - the enable rule of pll_default_enable_repeats (src/repeats.c:100-110) evaluated by every block   *
 *     of a job from the children's class counts, which the previous level left in d_node_ids;          *
 *   - 64-bit lookup entries (tag << 32 | first site): a pass uses a tag smaller than every tag used    *
 *     before, so atomicMin replaces whatever an earlier pass left and nothing has to be cleaned.       */
struct RidCtx
{
  bool enabled;
  unsigned int ids_left;
  const unsigned int * idl;
  const unsigned int * idr;
  unsigned long long * lookup;
};

__device__ __forceinline__ RidCtx rid_ctx(const plf_rid_job_t & jb, unsigned int sites, unsigned int lookup_size,
                                          const unsigned int * __restrict__ node_ids,
                                          unsigned long long * __restrict__ pool)
{

  RidCtx c;
  const unsigned long long nl = node_ids[jb.left], nr = node_ids[jb.right];
  const unsigned long long pairs = nl * nr;
  c.enabled = pairs && pairs < (unsigned long long)lookup_size && nl <= sites / 2 && nr <= sites / 2 &&
              pairs <= jb.lookup_entries;
  c.ids_left = (unsigned int)nl;
  c.idl = jb.site_id_left;
  c.idr = jb.site_id_right;
  c.lookup = pool + jb.lookup_offset;
  return c;
}
#define RID_KEY(c, s) ((c).idl[(s)] + (c).idr[(s)] * (c).ids_left)

/* work items of the per-site passes: (job, chunk of RID_CHUNK sites); a persistent grid strides over them, so
 * a level of hundreds of small jobs and a level of one large job fill the chip alike */
#define RID_CHUNK 4096u
#define RID_THREADS 256

__global__ void __launch_bounds__(RID_THREADS)
k_rid_min(const plf_rid_job_t * __restrict__ jobs, unsigned int njobs, unsigned int sites, unsigned int lookup_size,
          const unsigned int * __restrict__ node_ids, unsigned long long * __restrict__ pool,
          const unsigned int * __restrict__ tag_base, unsigned int tag_offset)
{
  const unsigned int chunks = (sites + RID_CHUNK - 1) / RID_CHUNK;
  /* the tag lives in device memory so that a captured graph of the whole identifier update can be replayed:
   * k_rid_advance lowers the base once per update, every pass of the update adds its own fixed offset */
  const unsigned long long hi = (unsigned long long)(*tag_base + tag_offset) << 32;
  for (unsigned int w = blockIdx.x; w < njobs * chunks; w += gridDim.x)
  {
    const RidCtx c = rid_ctx(jobs[w / chunks], sites, lookup_size, node_ids, pool);
    if (!c.enabled) continue;
    const unsigned int lo = (w % chunks) * RID_CHUNK, end = min(lo + RID_CHUNK, sites);
    /* four sites in flight per thread: identifiers of both children, then the lookup entry.  Sites are visited
     * in increasing order, so most of them find their class already claimed by an earlier site of this pass:
     * the plain L2 read then spares the atomic (and the serialisation on popular keys) */
    for (unsigned int s0 = lo + threadIdx.x; s0 < end; s0 += 4 * RID_THREADS)
    {
      unsigned long long * e[4];
      unsigned long long seen[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        const unsigned int s = s0 + u * RID_THREADS;
        e[u] = s < end ? &c.lookup[RID_KEY(c, s)] : nullptr;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) seen[u] = e[u] ? __ldcg(e[u]) : 0ull;
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        const unsigned long long mine = hi | (s0 + u * RID_THREADS);
        if (e[u] && seen[u] > mine) atomicMin(e[u], mine);
      }
    }
  }
}

__global__ void k_rid_count(const plf_rid_job_t * __restrict__ jobs, unsigned int sites, unsigned int lookup_size,
                            const unsigned int * __restrict__ node_ids, unsigned long long * __restrict__ pool,
                            unsigned int * __restrict__ tile_count_all, unsigned int ntiles)
{
  const RidCtx c = rid_ctx(jobs[blockIdx.y], sites, lookup_size, node_ids, pool);
  if (!c.enabled) return;
  __shared__ unsigned int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const unsigned int base = blockIdx.x * SCAN_TILE;
  unsigned int mine = 0;
  for (unsigned int s = base + threadIdx.x; s < base + SCAN_TILE && s < sites; s += blockDim.x)
    mine += ((unsigned int)c.lookup[RID_KEY(c, s)] == s);
  atomicAdd(&cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0) tile_count_all[(size_t)blockIdx.y * ntiles + blockIdx.x] = cnt;
}

/* exclusive scan of a job's tile counts; the class count goes to raw_ids[job] and, with the reference's
 * "no compression => 0" rule (src/repeats.c:366-372), to node_ids[parent] */
__global__ void k_rid_scan(const plf_rid_job_t * __restrict__ jobs, unsigned int sites, unsigned int lookup_size,
                           unsigned int * __restrict__ node_ids, unsigned long long * __restrict__ pool,
                           unsigned int * __restrict__ tile_count_all, unsigned int ntiles,
                           unsigned int * __restrict__ raw_ids)
{
  const plf_rid_job_t jb = jobs[blockIdx.y];
  const RidCtx c = rid_ctx(jb, sites, lookup_size, node_ids, pool);
  __shared__ unsigned int buf[1024];
  __shared__ unsigned int carry;
  if (!c.enabled)
  {
    if (threadIdx.x == 0)
    {
      raw_ids[blockIdx.y] = 0;
      node_ids[jb.parent] = 0;
    }
    return;
  }
  unsigned int * tile_count = tile_count_all + (size_t)blockIdx.y * ntiles;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (unsigned int base = 0; base < ntiles; base += 1024)
  {
    const unsigned int i = base + threadIdx.x;
    const unsigned int v = i < ntiles ? tile_count[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (unsigned int o = 1; o < 1024; o <<= 1)
    {
      unsigned int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < ntiles) tile_count[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0)
  {
    raw_ids[blockIdx.y] = carry;
    /* the children's counts were read by every block of this job's earlier kernels; later levels read this */
    node_ids[jb.parent] = carry >= sites ? 0u : carry;
  }
}

/* NOTE: k_rid_rank / k_rid_assign run after k_rid_scan has overwritten node_ids[parent]; a job never has its
 * own parent as a child (the host falls back to the per-op path for lists that recycle buffers), so the
 * enable rule still sees the children's counts.
 * Rank of every first occurrence within its 1024-site tile by ballots (one bit per site), plus the tile's
 * offset: the class number.  It goes to id_site_parent[rank] = site and to rank_pool[key], a table with the
 * lookup pool's index space, from which k_rid_assign reads every site's class with ONE gather. */
__global__ void __launch_bounds__(RID_THREADS)
k_rid_rank(const plf_rid_job_t * __restrict__ jobs, unsigned int sites, unsigned int lookup_size,
           const unsigned int * __restrict__ node_ids, unsigned long long * __restrict__ pool,
           const unsigned int * __restrict__ tile_count_all, unsigned int ntiles, unsigned int * __restrict__ rank_pool)
{
  const plf_rid_job_t jb = jobs[blockIdx.y];
  const RidCtx c = rid_ctx(jb, sites, lookup_size, node_ids, pool);
  if (!c.enabled) return;
  const unsigned int * tile_offset = tile_count_all + (size_t)blockIdx.y * ntiles;
  unsigned int * rank_of_key = rank_pool + jb.lookup_offset;
  /* a tile is 4 rows of 256 sites; thread t owns site t of every row (four independent load chains);
   * counts[row * 8 + warp] in site order */
  __shared__ unsigned int counts[32];
  const unsigned int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned int base = blockIdx.x * SCAN_TILE;
  unsigned int key[4], bal[4];
  bool f[4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
  {
    const unsigned int s = base + u * RID_THREADS + threadIdx.x;
    key[u] = s < sites ? RID_KEY(c, s) : 0u;
  }
  unsigned long long ent[4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
  {
    const unsigned int s = base + u * RID_THREADS + threadIdx.x;
    ent[u] = s < sites ? c.lookup[key[u]] : ~0ull;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
  {
    const unsigned int s = base + u * RID_THREADS + threadIdx.x;
    f[u] = s < sites && (unsigned int)ent[u] == s;
    bal[u] = __ballot_sync(0xffffffffu, f[u]);
    if (lane == 0) counts[u * 8 + w] = __popc(bal[u]);
  }
  __syncthreads();
  if (w == 0)
  {
    /* exclusive scan of the 32 counts */
    unsigned int v = counts[lane], x = v;
    for (int o = 1; o < 32; o <<= 1)
    {
      const unsigned int t = __shfl_up_sync(0xffffffffu, x, o);
      if ((int)lane >= o) x += t;
    }
    counts[lane] = x - v;
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (f[u])
    {
      const unsigned int r = tile_offset[blockIdx.x] + counts[u * 8 + w] + __popc(bal[u] & ((1u << lane) - 1u));
      rank_of_key[key[u]] = r;
      jb.id_site_parent[r] = base + u * RID_THREADS + threadIdx.x;
    }
}

__global__ void __launch_bounds__(RID_THREADS)
k_rid_assign(const plf_rid_job_t * __restrict__ jobs, unsigned int njobs, unsigned int sites, unsigned int lookup_size,
             const unsigned int * __restrict__ node_ids, unsigned long long * __restrict__ pool,
             const unsigned int * __restrict__ rank_pool)
{
  const unsigned int chunks = (sites + RID_CHUNK - 1) / RID_CHUNK;
  for (unsigned int w = blockIdx.x; w < njobs * chunks; w += gridDim.x)
  {
    const plf_rid_job_t jb = jobs[w / chunks];
    const RidCtx c = rid_ctx(jb, sites, lookup_size, node_ids, pool);
    if (!c.enabled) continue;
    const unsigned int * rank_of_key = rank_pool + jb.lookup_offset;
    const unsigned int lo = (w % chunks) * RID_CHUNK, end = min(lo + RID_CHUNK, sites);
    for (unsigned int s0 = lo + threadIdx.x; s0 < end; s0 += 4 * RID_THREADS)
    {
      unsigned int key[4], r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        const unsigned int s = s0 + u * RID_THREADS;
        key[u] = s < end ? RID_KEY(c, s) : 0xFFFFFFFFu;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = key[u] != 0xFFFFFFFFu ? rank_of_key[key[u]] : 0u;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (key[u] != 0xFFFFFFFFu) jb.site_id_parent[s0 + u * RID_THREADS] = r[u];
    }
  }
}

__global__ void k_rid_advance(unsigned int * tag_base, unsigned int n) { *tag_base -= n; }

extern "C" int plf_repeats_advance_tags(plf_ctx_t * ctx, unsigned int * d_tag_base, unsigned int n)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  k_rid_advance<<<1, 1, 0, ctx->stream>>>(d_tag_base, n);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* stream capture of a sequence of launches queued through this backend, for callers that repeat it */
extern "C" int plf_capture_begin(plf_ctx_t * ctx)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  PLF_CHECK(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  return 1;
}
extern "C" int plf_capture_end(plf_ctx_t * ctx, void ** exec_out)
{
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  *exec_out = nullptr;
  const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
  if (e != cudaSuccess || !graph)
  {
    cudaGetLastError();
    plf_set_error(ctx, "stream capture failed: %s", cudaGetErrorString(e));
    return 0;
  }
  const cudaError_t e2 = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e2 != cudaSuccess)
  {
    cudaGetLastError();
    plf_set_error(ctx, "graph instantiation failed: %s", cudaGetErrorString(e2));
    return 0;
  }
  *exec_out = exec;
  return 1;
}
extern "C" int plf_capture_abort(plf_ctx_t * ctx)
{
  cudaGraph_t graph = nullptr;
  cudaStreamEndCapture(ctx->stream, &graph);
  if (graph) cudaGraphDestroy(graph);
  cudaGetLastError();
  return 1;
}
extern "C" int plf_graph_replay(plf_ctx_t * ctx, void * exec, unsigned long long launches)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  PLF_CHECK(ctx, cudaGraphLaunch((cudaGraphExec_t)exec, ctx->stream));
  plf_count_launches(launches);
  return 1;
}
extern "C" void plf_graph_free(void * exec)
{
  if (exec) cudaGraphExecDestroy((cudaGraphExec_t)exec);
}

extern "C" size_t plf_repeats_pass_workspace(unsigned int sites, unsigned int njobs)
{
  const size_t ntiles = ((size_t)sites + SCAN_TILE - 1) / SCAN_TILE;
  return (size_t)njobs * ntiles * sizeof(unsigned int) + 64;
}

extern "C" int plf_repeats_pass(plf_ctx_t * ctx, unsigned int sites, unsigned int lookup_buffer_size,
                                const plf_rid_job_t * d_jobs, unsigned int first_job, unsigned int njobs,
                                unsigned long long * d_lookup_pool, unsigned int * d_rank_pool,
                                const unsigned int * d_tag_base, unsigned int tag_offset,
                                unsigned int * d_node_ids, unsigned int * d_raw_ids, void * d_scratch)
{
  if (!njobs) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  const unsigned int ntiles = (sites + SCAN_TILE - 1) / SCAN_TILE;
  unsigned int * tile = (unsigned int *)d_scratch;
  const plf_rid_job_t * jobs = d_jobs + first_job;
  const unsigned int chunks = (sites + RID_CHUNK - 1) / RID_CHUNK;
  unsigned long long items = (unsigned long long)njobs * chunks;
  const unsigned int cap = (unsigned int)ctx->sm_count * 8; /* resident CTAs of 256 threads */
  const unsigned int grid = (unsigned int)(items < cap ? items : cap);
  const dim3 gt(ntiles, njobs);
  k_rid_min<<<grid, RID_THREADS, 0, ctx->stream>>>(jobs, njobs, sites, lookup_buffer_size, d_node_ids, d_lookup_pool,
                                                  d_tag_base, tag_offset);
  k_rid_count<<<gt, 256, 0, ctx->stream>>>(jobs, sites, lookup_buffer_size, d_node_ids, d_lookup_pool, tile, ntiles);
  k_rid_scan<<<dim3(1, njobs), 1024, 0, ctx->stream>>>(jobs, sites, lookup_buffer_size, d_node_ids, d_lookup_pool, tile,
                                                       ntiles, d_raw_ids + first_job);
  k_rid_rank<<<gt, RID_THREADS, 0, ctx->stream>>>(jobs, sites, lookup_buffer_size, d_node_ids, d_lookup_pool, tile, ntiles,
                                                 d_rank_pool);
  k_rid_assign<<<grid, RID_THREADS, 0, ctx->stream>>>(jobs, njobs, sites, lookup_buffer_size, d_node_ids, d_lookup_pool,
                                                     d_rank_pool);
  for (int i = 0; i < 5; ++i) plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* pair lists of gathering ops (see plf_backend.h): one launch for the whole batch */
__global__ void k_rep_pairs(const plf_pair_job_t * __restrict__ jobs)
{
  const plf_pair_job_t jb = jobs[blockIdx.y];
  for (unsigned int n = blockIdx.x * blockDim.x + threadIdx.x; n < jb.entries; n += gridDim.x * blockDim.x)
  {
    const unsigned int site = jb.parent_id_site ? jb.parent_id_site[n] : n;
    uint2 v;
    v.x = jb.left_site_id ? jb.left_site_id[site] : site;
    v.y = jb.right_site_id ? jb.right_site_id[site] : site;
    reinterpret_cast<uint2 *>(jb.out)[n] = v;
  }
}

extern "C" int plf_repeats_pairs(plf_ctx_t * ctx, const plf_pair_job_t * h_jobs, unsigned int njobs)
{
  if (!njobs) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  plf_pair_job_t * d_jobs = (plf_pair_job_t *)plf_ws_reserve(ctx, &ctx->ws_partial, (size_t)njobs * sizeof(plf_pair_job_t));
  if (!d_jobs) return 0;
  PLF_CHECK(ctx, cudaMemcpyAsync(d_jobs, h_jobs, (size_t)njobs * sizeof(plf_pair_job_t), cudaMemcpyHostToDevice, ctx->stream));
  unsigned int max_entries = 0;
  for (unsigned int i = 0; i < njobs; ++i)
    if (h_jobs[i].entries > max_entries) max_entries = h_jobs[i].entries;
  for (unsigned int first = 0; first < njobs; first += PLF_MAX_RUN_OPS)
  {
    const unsigned int n = njobs - first < PLF_MAX_RUN_OPS ? njobs - first : PLF_MAX_RUN_OPS;
    unsigned int blocks = (max_entries + 255) / 256;
    const unsigned int share = ((unsigned int)ctx->sm_count * 8) / n > 2 ? ((unsigned int)ctx->sm_count * 8) / n : 2;
    if (blocks > share) blocks = share;
    if (blocks < 1) blocks = 1;
    k_rep_pairs<<<dim3(blocks, n), 256, 0, ctx->stream>>>(d_jobs + first);
    plf_count_launch();
  }
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

/* keys[s] = class code of the tip character at site s (repeats.c:204-216) */
__global__ void k_tip_keys(const unsigned char * __restrict__ seq, const unsigned char * __restrict__ charmap,
                           unsigned int sites, unsigned int * __restrict__ keys)
{
  for (unsigned int s = blockIdx.x * blockDim.x + threadIdx.x; s < sites; s += gridDim.x * blockDim.x)
    keys[s] = charmap[seq[s]];
}

extern "C" int plf_tip_keys(plf_ctx_t * ctx, const unsigned char * d_seq, const unsigned char * d_charmap,
                            unsigned int sites, unsigned int * d_keys)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  unsigned int blocks = (sites + 255) / 256;
  if (blocks > (unsigned int)ctx->sm_count * 8) blocks = (unsigned int)ctx->sm_count * 8;
  k_tip_keys<<<blocks, 256, 0, ctx->stream>>>(d_seq, d_charmap, sites, d_keys);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}

extern "C" int plf_copy_d2d(plf_ctx_t * ctx, void * dst, const void * src, size_t bytes)
{
  if (!bytes) return 1;
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  PLF_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return 1;
}

extern "C" int plf_fill_u32(plf_ctx_t * ctx, unsigned int * d, unsigned int value, size_t n)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  if (value == 0 || value == 0xFFFFFFFFu)
  {
    PLF_CHECK(ctx, cudaMemsetAsync(d, value ? 0xFF : 0, n * sizeof(unsigned int), ctx->stream));
    return 1;
  }
  plf_set_error(ctx, "plf_fill_u32: only 0 and 0xFFFFFFFF are supported");
  return 0;
}

/* ---- marginal ancestral state probabilities ------------------------------------- *
 * anc[site][j] = sum_r clv[site][r][j] pi_r[j] w_r, normalised over j                *
 * (src/likelihood.c:733-760); one thread per site                                    */
__global__ void k_ancestral(const double * __restrict__ clv, const double * __restrict__ model, unsigned int sites,
                            int R, int st, int sp, double * __restrict__ out)
{
  const double * weights = model + R;
  const double * freqs = model + 3 * R;
  for (unsigned int n = blockIdx.x * blockDim.x + threadIdx.x; n < sites; n += gridDim.x * blockDim.x)
  {
    const double * c = clv + (size_t)n * R * sp;
    double * a = out + (size_t)n * st;
    double sum = 0;
    for (int j = 0; j < st; ++j)
    {
      double v = 0;
      for (int r = 0; r < R; ++r) v += c[(size_t)r * sp + j] * freqs[(size_t)r * sp + j] * weights[r];
      a[j] = v;
      sum += v;
    }
    for (int j = 0; j < st; ++j) a[j] /= sum;
  }
}

extern "C" int plf_ancestral(plf_ctx_t * ctx, const plf_shape_t * sh, const double * d_clv, const double * d_model,
                             unsigned int sites, double * d_out)
{
  PLF_CHECK(ctx, cudaSetDevice(ctx->device));
  unsigned int blocks = (sites + 127) / 128;
  const unsigned int cap = (unsigned int)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_ancestral<<<blocks, 128, 0, ctx->stream>>>(d_clv, d_model, sites, (int)sh->rate_cats, (int)sh->states,
                                              (int)sh->states_padded, d_out);
  plf_count_launch();
  PLF_CHECK(ctx, cudaGetLastError());
  return 1;
}
