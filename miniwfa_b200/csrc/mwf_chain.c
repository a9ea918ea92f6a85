/* placeholder replaced below in this round: see mwf_chain.c */
#include <stdio.h>
#include <stdlib.h>
#include "miniwfa.h"
void mwf_wfa_chain(void *km, const mwf_opt_t *opt, int32_t tl, const char *ts, int32_t ql, const char *qs, mwf_rst_t *r)
{
	(void)km; (void)opt; (void)tl; (void)ts; (void)ql; (void)qs; (void)r;
	fprintf(stderr, "[miniwfa_b200] mwf_wfa_chain: not built yet\n");
	abort();
}
