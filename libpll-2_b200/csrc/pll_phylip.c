/*
 * pll_phylip.c -- PHYLIP alignment reader (sequential and interleaved), the other input format either side of
 * the likelihood path (the FASTA reader is pll_fasta.c).  Host only.
 *
 * Reference: src/phylip.c (pll_phylip_open :284, _rewind :351, _close :374, _parse_interleaved :382,
 * _parse_sequential :570, pll_phylip_load :714).  Same handle layout, results, error codes and messages;
 * written around one helper that appends the data characters of a text line to a sequence.
 *
 * Format as the reference reads it:
 *   header      two positive integers (taxa, sites); anything but blanks after them is rejected
 *   sequential  per taxon: label (up to the first blank), then data on the rest of the line and on as many
 *               following lines as it takes to reach `sites` characters
 *   interleaved first block: per taxon one label and the first non-empty run of data; every following
 *               non-empty line continues the next taxon in turn; all taxa of a block carry the same number of
 *               characters
 * Characters are classified by the caller's table (0 stripped and counted, 1 data, 2 fatal, 3 ignored).
 * The handle's `lineno` stays at 1 after opening, as in the reference (its messages quote it).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pll_b200.h"

/* src/maps.c:181-205: the same classes as pll_map_fasta */
const unsigned int pll_map_phylip[256] = {
    [0 ... 8] = 2,    [9 ... 13] = 3,    [14 ... 31] = 2,   ['-'] = 1,         ['.'] = 1,
    ['0' ... '9'] = 1, ['?'] = 1,        ['A' ... 'Z'] = 1, ['a' ... 'z'] = 1,
};

static void phy_error(int code, const char * msg)
{
  pll_errno = code;
  snprintf(pll_errmsg, 200, "%s", msg);
}

static int phy_blank(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }

static int phy_grow(pll_phylip_t * fd, size_t capacity)
{
  char * bigger;
  if (capacity <= fd->line_maxsize) return 1;
  bigger = (char *)malloc(capacity);
  if (!bigger)
  {
    phy_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return 0;
  }
  if (fd->line) memcpy(bigger, fd->line, fd->line_size);
  free(fd->line);
  fd->line = bigger;
  fd->line_maxsize = capacity;
  return 1;
}

/* next text line without its newline, of any length; NULL (and fd->line released) at end of file */
static char * phy_nextline(pll_phylip_t * fd)
{
  fd->line_size = 0;
  while (fgets(fd->buffer, PLL_LINEALLOC, fd->fp))
  {
    const size_t n = strlen(fd->buffer);
    if (fd->line_size + n + 1 > fd->line_maxsize && !phy_grow(fd, fd->line_maxsize + n + PLL_LINEALLOC)) return NULL;
    memcpy(fd->line + fd->line_size, fd->buffer, n);
    fd->line_size += n;
    if (n && fd->buffer[n - 1] == '\n')
    {
      fd->line[fd->line_size - 1] = 0;
      return fd->line;
    }
  }
  if (!fd->line_size)
  {
    free(fd->line);
    fd->line = NULL;
    fd->line_maxsize = 0;
    return NULL;
  }
  fd->line[fd->line_size] = 0;
  return fd->line;
}

/* appends the data characters of `text` to sequence `seqno` from position `at`; count appended or -1 */
static int phy_take(pll_phylip_t * fd, pll_msa_t * msa, const char * text, int seqno, int at)
{
  char * dst = msa->sequence[seqno] + at;
  int n = 0;
  char c;
  while ((c = *text++))
    switch (fd->chrstatus[(unsigned char)c])
    {
      case 0:
        fd->stripped_count++;
        fd->stripped[(unsigned char)c]++;
        break;
      case 1:
        if (at + n >= msa->length)
        {
          pll_errno = PLL_ERROR_PHYLIP_LONGSEQ;
          snprintf(pll_errmsg, 200, "Sequence %d (%.100s) longer than expected", seqno + 1, msa->label[seqno]);
          return -1;
        }
        dst[n++] = c;
        break;
      case 2:
        if ((unsigned char)c >= 32)
        {
          pll_errno = PLL_ERROR_PHYLIP_ILLEGALCHAR;
          snprintf(pll_errmsg, 200, "illegal character '%c' on line %ld in the fasta file", c, fd->lineno);
        }
        else
        {
          pll_errno = PLL_ERROR_PHYLIP_UNPRINTABLECHAR;
          snprintf(pll_errmsg, 200, "illegal unprintable character %#.2x (hexadecimal) on line %ld in the fasta file",
                   c, fd->lineno);
        }
        return -1;
      default: /* 3: ignored */
        break;
    }
  return n;
}

/* header line -> an empty alignment of the announced shape */
static pll_msa_t * phy_begin(pll_phylip_t * fd)
{
  int taxa = 0, sites = 0, used = 0, i;
  const char * p = fd->line;
  pll_msa_t * msa;
  if (!p || sscanf(p, "%d%n", &taxa, &used) < 1 || !used || taxa <= 0)
  {
    phy_error(PLL_ERROR_PHYLIP_SYNTAX, "Invalid number of sequences in header");
    return NULL;
  }
  p += used;
  used = 0;
  if (sscanf(p, "%d%n", &sites, &used) < 1 || !used || sites <= 0)
  {
    phy_error(PLL_ERROR_PHYLIP_SYNTAX, "Invalid sequence length in header");
    return NULL;
  }
  p += used;
  while (*p && phy_blank(*p)) ++p;
  if (*p) return NULL; /* trailing options are not understood (src/phylip.c:222-240 ends in failure) */

  msa = (pll_msa_t *)calloc(1, sizeof(pll_msa_t));
  if (msa)
  {
    msa->count = taxa;
    msa->length = sites;
    msa->sequence = (char **)calloc((size_t)taxa, sizeof(char *));
    msa->label = (char **)calloc((size_t)taxa, sizeof(char *));
  }
  for (i = 0; msa && msa->sequence && msa->label && i < taxa; ++i)
  {
    msa->sequence[i] = (char *)malloc((size_t)sites + 1);
    if (!msa->sequence[i]) break;
    msa->sequence[i][sites] = 0;
  }
  if (!msa || !msa->sequence || !msa->label || i < taxa || !phy_grow(fd, (size_t)sites + 300))
  {
    pll_msa_destroy(msa);
    phy_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  return msa;
}

/* the label at the start of `p` (up to the first blank) becomes label[seqno]; returns what follows it */
static char * phy_label(pll_msa_t * msa, int seqno, char * p)
{
  size_t n = 0;
  while (p[n] && !phy_blank(p[n])) ++n;
  msa->label[seqno] = (char *)malloc(n + 1);
  if (!msa->label[seqno])
  {
    phy_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  memcpy(msa->label[seqno], p, n);
  msa->label[seqno][n] = 0;
  return p + n;
}

/* next line that is not blank, positioned on its first non-blank character; NULL at end of file */
static char * phy_next_content(pll_phylip_t * fd)
{
  char * p;
  while ((p = phy_nextline(fd)))
  {
    while (*p && phy_blank(*p)) ++p;
    if (*p) return p;
  }
  return NULL;
}

static pll_msa_t * phy_fail(pll_msa_t * msa, int code, const char * fmt, int a, int b, int c)
{
  if (fmt)
  {
    pll_errno = code;
    snprintf(pll_errmsg, 200, fmt, a, b, c);
  }
  pll_msa_destroy(msa);
  return NULL;
}

/* One run of an interleaved block: data of `text` and, while that is empty, of the following lines.
 * 1 = taken (all runs of a block must be equally long: *run_len), 0 = end of file, -1 = error. */
static int phy_run(pll_phylip_t * fd, pll_msa_t * msa, char * text, int seqno, int at, int * run_len)
{
  while (text)
  {
    const int n = phy_take(fd, msa, text, seqno, at);
    if (n < 0) return -1;
    if (n)
    {
      if (!*run_len)
        *run_len = n;
      else if (*run_len != n)
      {
        pll_errno = PLL_ERROR_PHYLIP_NONALIGNED;
        snprintf(pll_errmsg, 200, "Sequence %d (%.100s) data out of alignment", seqno + 1, msa->label[seqno]);
        return -1;
      }
      return 1;
    }
    text = phy_nextline(fd);
  }
  return 0;
}

PLL_EXPORT pll_msa_t * pll_phylip_parse_interleaved(pll_phylip_t * fd)
{
  pll_msa_t * msa = phy_begin(fd);
  int seqno = 0, run_len = 0, done = 0, block = 2, r = 1;
  char * p;
  if (!msa) return NULL;

  /* first block: labels and first runs */
  while (seqno < msa->count && (p = phy_next_content(fd)))
  {
    if (!(p = phy_label(msa, seqno, p))) return phy_fail(msa, 0, NULL, 0, 0, 0);
    r = phy_run(fd, msa, p, seqno, 0, &run_len);
    if (r <= 0) break;
    ++seqno;
  }
  if (r < 0) return phy_fail(msa, 0, NULL, 0, 0, 0);
  if (seqno != msa->count)
    return phy_fail(msa, PLL_ERROR_PHYLIP_SYNTAX, "Found %d sequence(s) but expected %d", seqno, msa->count, 0);
  done = run_len;

  /* every further non-empty line continues the next taxon in turn */
  seqno = 0;
  run_len = 0;
  while ((r = phy_run(fd, msa, phy_nextline(fd), seqno, done, &run_len)) > 0)
  {
    if (++seqno == msa->count)
    {
      seqno = 0;
      done += run_len;
      run_len = 0;
      ++block;
    }
  }
  if (r < 0) return phy_fail(msa, 0, NULL, 0, 0, 0);
  if (seqno)
    return phy_fail(msa, PLL_ERROR_PHYLIP_SYNTAX, "Found %d sequences in block %d but expected %d", seqno, block,
                    msa->count);
  if (done != msa->length)
    return phy_fail(msa, PLL_ERROR_PHYLIP_SYNTAX, "Sequence length is %d but expected %d", done, msa->length, 0);
  return msa;
}

PLL_EXPORT pll_msa_t * pll_phylip_parse_sequential(pll_phylip_t * fd)
{
  pll_msa_t * msa = phy_begin(fd);
  int seqno = 0;
  char * p;
  if (!msa) return NULL;
  while ((p = phy_next_content(fd)))
  {
    int have = 0;
    if (seqno == msa->count)
      return phy_fail(msa, PLL_ERROR_PHYLIP_SYNTAX, "Found at least %d sequences but expected %d", seqno + 1,
                      msa->count, 0);
    if (!(p = phy_label(msa, seqno, p))) return phy_fail(msa, 0, NULL, 0, 0, 0);
    for (;;)
    {
      const int n = phy_take(fd, msa, p, seqno, have);
      if (n < 0) return phy_fail(msa, 0, NULL, 0, 0, 0);
      have += n;
      if (have == msa->length) break;
      if (!(p = phy_nextline(fd)))
      {
        pll_errno = PLL_ERROR_PHYLIP_SYNTAX;
        snprintf(pll_errmsg, 200, "Sequence %d (%.100s) has %d characters but expected %d", seqno + 1,
                 msa->label[seqno], have, msa->length);
        return phy_fail(msa, 0, NULL, 0, 0, 0);
      }
    }
    ++seqno;
  }
  if (seqno != msa->count)
    return phy_fail(msa, PLL_ERROR_PHYLIP_SYNTAX, "Found %d sequence(s) but expected %d", seqno, msa->count, 0);
  return msa;
}

static void phy_reset_counts(pll_phylip_t * fd)
{
  fd->stripped_count = 0;
  memset(fd->stripped, 0, sizeof(fd->stripped));
}

PLL_EXPORT pll_phylip_t * pll_phylip_open(const char * filename, const unsigned int * map)
{
  pll_phylip_t * fd = (pll_phylip_t *)calloc(1, sizeof(pll_phylip_t));
  if (!fd)
  {
    phy_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  fd->no = -1;
  fd->chrstatus = map;
  fd->fp = fopen(filename, "r");
  if (!fd->fp)
  {
    pll_errno = PLL_ERROR_FILE_OPEN;
    snprintf(pll_errmsg, 200, "Unable to open file (%s)", filename);
    free(fd);
    return NULL;
  }
  if (fseek(fd->fp, 0, SEEK_END))
  {
    pll_errno = PLL_ERROR_FILE_SEEK;
    snprintf(pll_errmsg, 200, "Unable to seek in file (%s)", filename);
    fclose(fd->fp);
    free(fd);
    return NULL;
  }
  fd->filesize = ftell(fd->fp);
  rewind(fd->fp);
  phy_reset_counts(fd);
  if (!phy_nextline(fd)) /* the header line is cached in the handle */
  {
    free(fd->line);
    fclose(fd->fp);
    free(fd);
    return NULL;
  }
  fd->lineno = 1;
  return fd;
}

PLL_EXPORT int pll_phylip_rewind(pll_phylip_t * fd)
{
  rewind(fd->fp);
  phy_reset_counts(fd);
  if (!phy_nextline(fd))
  {
    phy_error(PLL_ERROR_FILE_SEEK, "Unable to rewind and cache data");
    return PLL_FAILURE;
  }
  fd->lineno = 1;
  fd->no = -1;
  return PLL_SUCCESS;
}

PLL_EXPORT void pll_phylip_close(pll_phylip_t * fd)
{
  if (!fd) return;
  fclose(fd->fp);
  free(fd->line);
  free(fd);
}

PLL_EXPORT pll_msa_t * pll_phylip_load(const char * fname, pll_bool_t interleaved)
{
  pll_phylip_t * fd = pll_phylip_open(fname, pll_map_generic);
  pll_msa_t * msa;
  if (!fd) return NULL;
  msa = interleaved ? pll_phylip_parse_interleaved(fd) : pll_phylip_parse_sequential(fd);
  pll_phylip_close(fd);
  return msa;
}
