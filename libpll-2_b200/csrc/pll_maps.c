/*
 * pll_maps.c -- character -> state-mask tables callers pass to
 * pll_set_tip_states (same symbols and values as the reference's src/maps.c:26-180;
 * the values are the IUPAC nucleotide / amino-acid ambiguity codes).
 */
#include "pll_b200.h"

#define BOTH(u, l, v) [u] = (v), [l] = (v)

const pll_state_t pll_map_bin[256] = {
  ['0'] = 1, ['1'] = 2, ['-'] = 3, ['.'] = 3, ['?'] = 3,
};

const pll_state_t pll_map_nt[256] = {
  BOTH('A', 'a', 1),  BOTH('C', 'c', 2),  BOTH('G', 'g', 4),  BOTH('T', 't', 8),
  BOTH('U', 'u', 8),  BOTH('M', 'm', 3),  BOTH('R', 'r', 5),  BOTH('S', 's', 6),
  BOTH('V', 'v', 7),  BOTH('W', 'w', 9),  BOTH('Y', 'y', 10), BOTH('H', 'h', 11),
  BOTH('K', 'k', 12), BOTH('D', 'd', 13), BOTH('B', 'b', 14), BOTH('N', 'n', 15),
  BOTH('O', 'o', 15), BOTH('X', 'x', 15), ['-'] = 15, ['.'] = 15, ['?'] = 15,
};

#define AA(i) (1ull << (i))
const pll_state_t pll_map_aa[256] = {
  BOTH('A', 'a', AA(0)),  BOTH('R', 'r', AA(1)),  BOTH('N', 'n', AA(2)),  BOTH('D', 'd', AA(3)),
  BOTH('C', 'c', AA(4)),  BOTH('Q', 'q', AA(5)),  BOTH('E', 'e', AA(6)),  BOTH('G', 'g', AA(7)),
  BOTH('H', 'h', AA(8)),  BOTH('I', 'i', AA(9)),  BOTH('L', 'l', AA(10)), BOTH('K', 'k', AA(11)),
  BOTH('M', 'm', AA(12)), BOTH('F', 'f', AA(13)), BOTH('P', 'p', AA(14)), BOTH('S', 's', AA(15)),
  BOTH('T', 't', AA(16)), BOTH('W', 'w', AA(17)), BOTH('Y', 'y', AA(18)), BOTH('V', 'v', AA(19)),
  BOTH('B', 'b', AA(2) | AA(3)),   /* N or D */
  BOTH('Z', 'z', AA(5) | AA(6)),   /* Q or E */
  BOTH('J', 'j', AA(9) | AA(10)),  /* I or L */
  BOTH('X', 'x', 0xFFFFF), ['*'] = 0xFFFFF, ['-'] = 0xFFFFF, ['.'] = 0xFFFFF, ['?'] = 0xFFFFF,
};
