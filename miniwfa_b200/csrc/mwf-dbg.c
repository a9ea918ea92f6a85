/* mwf-dbg.c -- CIGAR self-checks (reference mwf-dbg.c:6-31) */
#include <assert.h>
#include <stdio.h>
#include "miniwfa.h"

int32_t mwf_cigar2score(const mwf_opt_t *opt, int32_t n_cigar, const uint32_t *cigar, int32_t *tl, int32_t *ql)
{
	int32_t i, score = 0, t_used = 0, q_used = 0;
	for (i = 0; i < n_cigar; ++i) {
		const int32_t op = cigar[i] & 0xf, len = (int32_t)(cigar[i] >> 4);
		switch (op) {
		case 1: case 2: { /* a gap is charged by the cheaper of the two affine pieces */
			const int32_t p1 = opt->o1 + len * opt->e1, p2 = opt->o2 + len * opt->e2;
			score += p1 < p2 ? p1 : p2;
			if (op == 1) q_used += len; else t_used += len;
			break;
		}
		case 8: score += len * opt->x; /* fall through */
		case 0: case 7: t_used += len, q_used += len; break;
		default: break;
		}
	}
	if (tl) *tl = t_used;
	if (ql) *ql = q_used;
	return score;
}

void mwf_assert_cigar(const mwf_opt_t *opt, int32_t n_cigar, const uint32_t *cigar, int32_t tl0, int32_t ql0, int32_t s0)
{
	int32_t tl, ql;
	const int32_t s = mwf_cigar2score(opt, n_cigar, cigar, &tl, &ql);
	assert(tl == tl0);
	assert(ql == ql0);
	if (s > s0) fprintf(stderr, "[%s] s0=%d, s=%d\n", __func__, s0, s);
}
