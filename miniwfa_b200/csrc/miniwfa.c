/*
 * miniwfa.c -- host C driver behind the miniwfa API (reference miniwfa.c:11-18, 603-615, 898-908).
 *
 * The reference's per-score loop (mwf_wfa_core / mwf_wfa_seg) does not exist on the host
 * any more: mwf_wfa_exact() submits the pair to the CUDA engine through the extern "C"
 * shim in mwf_b200.h and only allocates the result from the caller's arena.
 */
#include <string.h>
#include "miniwfa.h"
#include "mwf_b200.h"

void mwf_opt_init(mwf_opt_t *opt) /* same defaults as reference miniwfa.c:11-18 */
{
	memset(opt, 0, sizeof(*opt));
	opt->x = 4;
	opt->o1 = 4, opt->e1 = 2;
	opt->o2 = 15, opt->e2 = 1;
	opt->kmer = 13, opt->max_occ = 2, opt->min_len = 30;
}

void mwf_wfa_exact(void *km, const mwf_opt_t *opt, int32_t tl, const char *ts, int32_t ql, const char *qs, mwf_rst_t *r)
{
	mwf_wfa_exact_batch(km, opt, 1, &tl, &ts, &ql, &qs, r);
}

void mwf_wfa_auto(void *km, const mwf_opt_t *opt0, int32_t tl, const char *ts, int32_t ql, const char *qs, mwf_rst_t *r)
{
	mwf_opt_t opt = *opt0; /* reference miniwfa.c:898-908 */
	opt.step = 0, opt.max_iter = 100000000;
	mwf_wfa_exact(km, &opt, tl, ts, ql, qs, r);
	if (r->s < 0) {
		if (opt.flag & MWF_F_CIGAR) opt.step = 5000;
		opt.max_iter = -1;
		mwf_wfa_chain(km, &opt, tl, ts, ql, qs, r);
	}
}
