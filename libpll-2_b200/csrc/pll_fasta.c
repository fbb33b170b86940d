/*
 * pll_fasta.c -- FASTA reader (the data format in front of pll_set_tip_states /
 * pll_compress_site_patterns).
 *
 * Same API, struct layout and error codes as the reference (src/fasta.c:40-417,
 * src/pll.h:358-370, 864-887): pll_fasta_open / _getnext / _rewind / _close /
 * _getfilesize / _getfilepos, pll_fasta_load, pll_msa_destroy, and the character
 * class tables pll_map_fasta / pll_map_generic (0 = stripped and counted, 1 = legal,
 * 2 = fatal, 3 = silently stripped).  One line of look-ahead lives in the handle
 * (fd->line), records are assembled in buffers that grow geometrically.  Host only.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pll_b200.h"

/* src/maps.c:207-232: letters, digits, '-', '.', '?' are sequence characters; control characters
 * are fatal except TAB/LF/VT/FF/CR, which are skipped; everything else is stripped and counted */
const unsigned int pll_map_fasta[256] = {
    [0 ... 8] = 2,    [9 ... 13] = 3,    [14 ... 31] = 2,   ['-'] = 1,         ['.'] = 1,
    ['0' ... '9'] = 1, ['?'] = 1,        ['A' ... 'Z'] = 1, ['a' ... 'z'] = 1,
};

/* src/maps.c:234-262: any printable non-blank character is data */
const unsigned int pll_map_generic[256] = {
    [0 ... 8] = 2, [9 ... 13] = 3, [14 ... 31] = 2, [32] = 3, [33 ... 126] = 1, [127] = 2, [128 ... 254] = 1, [255] = 2,
};

static void fasta_error(int code, const char * msg)
{
  pll_errno = code;
  snprintf(pll_errmsg, 200, "%s", msg);
}

/* caches the next line (empty string at end of file) */
static void next_line(pll_fasta_t * fd)
{
  fd->line[0] = 0;
  if (!fgets(fd->line, PLL_LINEALLOC, fd->fp)) fd->line[0] = 0;
}

static int prime(pll_fasta_t * fd, const char * what)
{
  int i;
  rewind(fd->fp);
  fd->stripped_count = 0;
  for (i = 0; i < 256; ++i) fd->stripped[i] = 0;
  next_line(fd);
  if (!fd->line[0])
  {
    fasta_error(PLL_ERROR_FILE_SEEK, what);
    return 0;
  }
  fd->lineno = 1;
  return 1;
}

PLL_EXPORT pll_fasta_t * pll_fasta_open(const char * filename, const unsigned int * map)
{
  pll_fasta_t * fd = (pll_fasta_t *)calloc(1, sizeof(pll_fasta_t));
  char msg[200];
  if (!fd)
  {
    fasta_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  fd->no = -1;
  fd->chrstatus = map;
  fd->fp = fopen(filename, "r");
  if (!fd->fp)
  {
    snprintf(msg, sizeof(msg), "Unable to open file (%s)", filename);
    fasta_error(PLL_ERROR_FILE_OPEN, msg);
    free(fd);
    return NULL;
  }
  if (fseek(fd->fp, 0, SEEK_END))
  {
    snprintf(msg, sizeof(msg), "Unable to seek in file (%s)", filename);
    fasta_error(PLL_ERROR_FILE_SEEK, msg);
    fclose(fd->fp);
    free(fd);
    return NULL;
  }
  fd->filesize = ftell(fd->fp);
  snprintf(msg, sizeof(msg), "Unable to read file (%s)", filename);
  if (!prime(fd, msg))
  {
    fclose(fd->fp);
    free(fd);
    return NULL;
  }
  return fd;
}

PLL_EXPORT int pll_fasta_rewind(pll_fasta_t * fd)
{
  return prime(fd, "Unable to rewind and cache data") ? PLL_SUCCESS : PLL_FAILURE;
}

PLL_EXPORT void pll_fasta_close(pll_fasta_t * fd)
{
  fclose(fd->fp);
  free(fd);
}

PLL_EXPORT long pll_fasta_getfilesize(const pll_fasta_t * fd) { return fd->filesize; }
PLL_EXPORT long pll_fasta_getfilepos(pll_fasta_t * fd) { return ftell(fd->fp); }

/* one record: header without '>' and line end, sequence with the character classes applied.
 * *head and *seq are malloc'ed and owned by the caller on success (src/fasta.c:130-316) */
PLL_EXPORT int pll_fasta_getnext(pll_fasta_t * fd, char ** head, long * head_len, char ** seq, long * seq_len,
                                 long * seqno)
{
  size_t cap = 4096, n = 0, hl;
  char * s, * h;
  const char * eol;
  *head_len = 0;
  *seq_len = 0;
  if (!fd->line[0])
  {
    pll_errno = PLL_ERROR_FILE_EOF;
    snprintf(pll_errmsg, 200, "End of file\n");
    return PLL_FAILURE;
  }
  if (fd->line[0] != '>')
  {
    fasta_error(PLL_ERROR_FASTA_INVALIDHEADER, "Illegal header line in query fasta file");
    return PLL_FAILURE;
  }
  /* header: up to the first CR if the line has one, else up to the LF */
  eol = strchr(fd->line + 1, '\r');
  if (!eol) eol = strchr(fd->line + 1, '\n');
  hl = eol ? (size_t)(eol - (fd->line + 1)) : strlen(fd->line + 1);
  h = (char *)malloc(hl + 1 > 4096 ? hl + 1 : 4096);
  s = (char *)malloc(cap);
  if (!h || !s)
  {
    free(h);
    free(s);
    fasta_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return PLL_FAILURE;
  }
  memcpy(h, fd->line + 1, hl);
  h[hl] = 0;
  next_line(fd);
  fd->lineno++;
  while (fd->line[0] && fd->line[0] != '>')
  {
    const unsigned char * p;
    for (p = (const unsigned char *)fd->line; *p; ++p)
    {
      switch (fd->chrstatus[*p])
      {
        case 0:
          fd->stripped_count++;
          fd->stripped[*p]++;
          break;
        case 1:
          if (n + 2 > cap)
          {
            char * grown = (char *)realloc(s, cap * 2);
            if (!grown)
            {
              free(h);
              free(s);
              fasta_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
              return PLL_FAILURE;
            }
            s = grown;
            cap *= 2;
          }
          s[n++] = (char)*p;
          break;
        case 2:
          if (*p >= 32 && *p < 128)
          {
            pll_errno = PLL_ERROR_FASTA_ILLEGALCHAR;
            snprintf(pll_errmsg, 200, "illegal character '%c' on line %ld in the fasta file", *p, fd->lineno);
          }
          else
          {
            pll_errno = PLL_ERROR_FASTA_UNPRINTABLECHAR;
            snprintf(pll_errmsg, 200, "illegal unprintable character %#.2x (hexadecimal) on line %ld in the fasta file",
                     (signed char)*p, fd->lineno);
          }
          free(h);
          free(s);
          return PLL_FAILURE;
        default: /* 3: silently stripped */
          break;
      }
    }
    next_line(fd);
    fd->lineno++;
  }
  s[n] = 0;
  *head = h;
  *head_len = (long)hl;
  *seq = s;
  *seq_len = (long)n;
  *seqno = ++fd->no;
  return PLL_SUCCESS;
}

/* src/phylip.c (pll_msa_destroy): labels and sequences are owned by the msa */
PLL_EXPORT void pll_msa_destroy(pll_msa_t * msa)
{
  int i;
  if (!msa) return;
  for (i = 0; i < msa->count; ++i)
  {
    if (msa->label) free(msa->label[i]);
    if (msa->sequence) free(msa->sequence[i]);
  }
  free(msa->label);
  free(msa->sequence);
  free(msa);
}

/* whole file into an alignment; all sequences must have one length (src/fasta.c:328-417) */
PLL_EXPORT pll_msa_t * pll_fasta_load(const char * fname)
{
  pll_fasta_t * fp = pll_fasta_open(fname, pll_map_generic);
  pll_msa_t * msa;
  char * seq = NULL, * hdr = NULL;
  long seqlen, hdrlen, seqno;
  size_t cap = 0;
  int i = 0;
  if (!fp) return NULL;
  msa = (pll_msa_t *)calloc(1, sizeof(pll_msa_t));
  if (!msa)
  {
    pll_fasta_close(fp);
    fasta_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
    return NULL;
  }
  msa->length = -1;
  while (pll_fasta_getnext(fp, &hdr, &hdrlen, &seq, &seqlen, &seqno))
  {
    if (msa->length == -1)
      msa->length = (int)seqlen;
    else if (msa->length != seqlen)
    {
      free(hdr);
      free(seq);
      msa->count = i;
      pll_errno = PLL_ERROR_FASTA_NONALIGNED;
      snprintf(pll_errmsg, 200,
               "FASTA file does not contain equal size sequences: sequence %d has length of %ld (expected: %d)", i,
               seqlen, msa->length);
      pll_msa_destroy(msa);
      pll_fasta_close(fp);
      return NULL;
    }
    if ((size_t)i >= cap)
    {
      const size_t want = cap ? cap * 2 : 64;
      char ** labels = (char **)realloc(msa->label, want * sizeof(char *));
      char ** seqs = labels ? (char **)realloc(msa->sequence, want * sizeof(char *)) : NULL;
      if (labels) msa->label = labels;
      if (seqs) msa->sequence = seqs;
      if (!labels || !seqs)
      {
        free(hdr);
        free(seq);
        msa->count = i;
        pll_msa_destroy(msa);
        pll_fasta_close(fp);
        fasta_error(PLL_ERROR_MEM_ALLOC, "Unable to allocate enough memory.");
        return NULL;
      }
      cap = want;
    }
    msa->label[i] = hdr;
    msa->sequence[i] = seq;
    ++i;
  }
  msa->count = i;
  pll_fasta_close(fp);
  if (pll_errno != PLL_ERROR_FILE_EOF)
  {
    pll_msa_destroy(msa);
    return NULL;
  }
  pll_errno = 0;
  return msa;
}
