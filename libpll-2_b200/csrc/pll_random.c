/*
 * pll_random.c -- re-entrant pseudo-random numbers with the semantics of glibc's random_r family, which the
 * reference carries as src/random.c (pll_random_r :236, pll_srandom_r :144, pll_initstate_r :205,
 * pll_setstate_r :300, pll_random_create/getint/destroy :395-423) so that a seed gives the same sequence on
 * every platform.  Stepwise addition (pll_parsimony.c) shuffles the taxa with it.
 *
 * Written from the published algorithm: an additive feedback generator x[i] = x[i - deg] + x[i - deg + sep]
 * over a table of `deg` 32-bit words (deg/sep = 7/3, 15/1, 31/3, 63/1 for state buffers of 32/64/128/256
 * bytes), output = x >> 1; an 8-byte buffer degenerates to the linear congruential generator
 * x = 1103515245 x + 12345 mod 2^31.  The table is seeded with the Park-Miller "minimal standard" generator
 * (16807 x mod 2^31 - 1, by Schrage's method) and the first 10 * deg outputs are discarded.  The word before
 * the table remembers generator type and rear position (5 * rear + type) across setstate calls.
 */
#include <stdlib.h>
#include <string.h>

#include "pll_b200.h"

#define RNG_TYPES 5
static const int rng_min_bytes[RNG_TYPES] = {8, 32, 64, 128, 256};
static const int rng_degree[RNG_TYPES] = {0, 7, 15, 31, 63};
static const int rng_sep[RNG_TYPES] = {0, 3, 1, 3, 1};

/* what the word in front of the table should hold for the generator currently in buf */
static void rng_stash(struct pll_random_data * buf)
{
  if (!buf->state) return;
  buf->state[-1] = buf->rand_type == 0 ? 0 : RNG_TYPES * (int)(buf->rptr - buf->state) + buf->rand_type;
}

PLL_EXPORT int pll_random_r(struct pll_random_data * buf, int * result)
{
  if (!buf || !result) return -1;
  if (buf->rand_type == 0)
  {
    const unsigned int x = ((unsigned int)buf->state[0] * 1103515245u + 12345u) & 0x7fffffffu;
    buf->state[0] = (int)x;
    *result = (int)x;
    return 0;
  }
  {
    int * f = buf->fptr, * r = buf->rptr;
    const unsigned int x = (unsigned int)*f + (unsigned int)*r;
    *f = (int)x;
    *result = (int)(x >> 1);
    if (++f >= buf->end_ptr)
    {
      f = buf->state;
      ++r;
    }
    else if (++r >= buf->end_ptr)
      r = buf->state;
    buf->fptr = f;
    buf->rptr = r;
  }
  return 0;
}

PLL_EXPORT int pll_srandom_r(unsigned int seed, struct pll_random_data * buf)
{
  int i, deg, word;
  if (!buf || buf->rand_type < 0 || buf->rand_type >= RNG_TYPES) return -1;
  if (!seed) seed = 1;
  buf->state[0] = (int)seed;
  if (buf->rand_type == 0) return 0;
  deg = buf->rand_deg;
  word = (int)seed;
  for (i = 1; i < deg; ++i)
  {
    /* 16807 * word mod (2^31 - 1) without overflow */
    const long hi = word / 127773, lo = word % 127773;
    long next = 16807 * lo - 2836 * hi;
    if (next < 0) next += 2147483647;
    word = (int)next;
    buf->state[i] = word;
  }
  buf->fptr = buf->state + buf->rand_sep;
  buf->rptr = buf->state;
  for (i = 0; i < 10 * deg; ++i)
  {
    int discard;
    pll_random_r(buf, &discard);
  }
  return 0;
}

PLL_EXPORT int pll_initstate_r(unsigned int seed, char * arg_state, size_t n, struct pll_random_data * buf)
{
  int type;
  if (!buf || !arg_state) return -1;
  rng_stash(buf);
  for (type = RNG_TYPES - 1; type >= 0 && n < (size_t)rng_min_bytes[type]; --type)
    ;
  if (type < 0) return -1;
  buf->rand_type = type;
  buf->rand_deg = rng_degree[type];
  buf->rand_sep = rng_sep[type];
  buf->state = (int *)arg_state + 1;
  buf->end_ptr = buf->state + buf->rand_deg;
  pll_srandom_r(seed, buf);
  rng_stash(buf);
  return 0;
}

PLL_EXPORT int pll_setstate_r(char * arg_state, struct pll_random_data * buf)
{
  int * table;
  int type;
  if (!buf || !arg_state) return -1;
  rng_stash(buf);
  table = (int *)arg_state + 1;
  type = table[-1] % RNG_TYPES;
  if (type < 0 || type >= RNG_TYPES) return -1;
  buf->rand_type = type;
  buf->rand_deg = rng_degree[type];
  buf->rand_sep = rng_sep[type];
  if (type != 0)
  {
    const int rear = table[-1] / RNG_TYPES;
    buf->rptr = table + rear;
    buf->fptr = table + (rear + buf->rand_sep) % buf->rand_deg;
  }
  buf->state = table;
  buf->end_ptr = table + buf->rand_deg;
  return 0;
}

PLL_EXPORT pll_random_state * pll_random_create(unsigned int seed)
{
  pll_random_state * rs = (pll_random_state *)calloc(1, sizeof(pll_random_state));
  if (!rs) return NULL;
  rs->state_buf = (char *)calloc(128, 1);
  if (!rs->state_buf)
  {
    free(rs);
    return NULL;
  }
  pll_initstate_r(seed, rs->state_buf, 128, &rs->rdata);
  pll_srandom_r(seed, &rs->rdata);
  return rs;
}

/* 0 <= r < maxval */
PLL_EXPORT int pll_random_getint(pll_random_state * rstate, int maxval)
{
  int r;
  pll_random_r(&rstate->rdata, &r);
  return r % maxval;
}

PLL_EXPORT void pll_random_destroy(pll_random_state * rstate)
{
  if (!rstate) return;
  free(rstate->state_buf);
  free(rstate);
}
