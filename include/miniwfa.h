/*
 * miniwfa.h -- public C API of the B200-native wavefront aligner.
 *
 * Drop-in for lh3/miniwfa's miniwfa.h: the same two structs (identical field
 * order and sizes: mwf_opt_t 56 bytes, mwf_rst_t 24 bytes on LP64), the same
 * flag values and the same entry points (reference miniwfa.h:32-51, 62, 83-89).
 * The per-score hot path (wf_next + wf_extend + traceback bytes) runs in
 * sm_100a CUDA kernels behind the extern "C" shim declared in mwf_b200.h.
 */
#ifndef MINIWFA_H
#define MINIWFA_H

#include <stdint.h>

#define MWF_F_CIGAR      0x1      /* produce a CIGAR */
#define MWF_F_NO_KALLOC  0x2      /* scratch from libc instead of child arenas */
#define MWF_F_DEBUG      0x10000  /* print the traceback end state to stderr */

typedef struct {
	int32_t flag;               /* MWF_F_* bits */
	int32_t x, o1, e1, o2, e2;  /* mismatch; gap open/extend of the two affine pieces */
	int32_t step;               /* >0: low-memory mode, checkpoint every `step` scores */
	int32_t max_s;              /* >0: give up (r->s = -1) once the score exceeds this */
	int64_t max_iter;           /* >0: give up once this many wavefront cells were computed */
	int32_t max_occ, kmer, min_len; /* chaining heuristic */
} mwf_opt_t;

typedef struct {
	int32_t s;        /* penalty of the optimal alignment, or -1 if stopped */
	int32_t n_cigar;  /* number of CIGAR operations */
	int64_t n_iter;   /* wavefront cells computed */
	uint32_t *cigar;  /* len<<4|op (7 '=', 8 'X', 1 'I', 2 'D'); owned by the caller's km */
} mwf_rst_t;

#ifdef __cplusplus
extern "C" {
#endif

/* defaults: x=4, o1=4, e1=2, o2=15, e2=1, kmer=13, max_occ=2, min_len=30 */
void mwf_opt_init(mwf_opt_t *opt);

/*
 * Global alignment of target ts[0..tl) against query qs[0..ql); arbitrary bytes, compared
 * for exact equality.  km is a kalloc arena or NULL for malloc; r is fully overwritten and
 * r->cigar (if any) is allocated from km.
 *   mwf_wfa_exact: optimal; step>0 selects the two-pass low-memory mode.
 *   mwf_wfa_chain: k-mer chaining heuristic, gaps closed with mwf_wfa_exact.
 *   mwf_wfa_auto : exact with a 1e8-cell budget, then chain with step=5000.
 */
void mwf_wfa_exact(void *km, const mwf_opt_t *opt, int32_t tl, const char *ts, int32_t ql, const char *qs, mwf_rst_t *r);
void mwf_wfa_chain(void *km, const mwf_opt_t *opt, int32_t tl, const char *ts, int32_t ql, const char *qs, mwf_rst_t *r);
void mwf_wfa_auto(void *km, const mwf_opt_t *opt, int32_t tl, const char *ts, int32_t ql, const char *qs, mwf_rst_t *r);

/* debugging helpers (reference mwf-dbg.c) */
int32_t mwf_cigar2score(const mwf_opt_t *opt, int32_t n_cigar, const uint32_t *cigar, int32_t *tl, int32_t *ql);
void mwf_assert_cigar(const mwf_opt_t *opt, int32_t n_cigar, const uint32_t *cigar, int32_t tl0, int32_t ql0, int32_t s0);

#ifdef __cplusplus
}
#endif

#endif
