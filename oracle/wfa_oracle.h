/*
 * oracle/wfa_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the exact dual-affine wavefront aligner of lh3/miniwfa
 * (reference @ 66770a3).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product library
 * (miniwfa_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  liboracle.so is checked against (a) the golden
 * vectors in tests/golden (JSON files), which were produced by the unmodified
 * reference compiled into oracle/_ref (tests/golden/make_golden.py), and
 * (b) oracle/_ref itself, differentially, whenever that build is present.
 *
 * The option / result structs have the same layout as the reference's
 * mwf_opt_t / mwf_rst_t (miniwfa.h:36-51) so one ctypes definition serves all
 * three libraries (reference, oracle, product).
 */
#ifndef WFA_ORACLE_H
#define WFA_ORACLE_H

#include <stdint.h>

#define ORC_F_CIGAR 0x1 /* miniwfa.h:32 */

typedef struct {
	int32_t flag;
	int32_t x, o1, e1, o2, e2;
	int32_t step;
	int32_t max_s;
	int64_t max_iter;
	int32_t max_occ, kmer, min_len; /* unused by the exact path */
} orc_opt_t;

typedef struct {
	int32_t s;
	int32_t n_cigar;
	int64_t n_iter;
	uint32_t *cigar; /* malloc'ed; release with orc_free() */
} orc_rst_t;

#ifdef __cplusplus
extern "C" {
#endif

void orc_opt_init(orc_opt_t *opt);                          /* miniwfa.c:11-18 */
void orc_wfa_exact(const orc_opt_t *opt, int32_t tl, const char *ts,
                   int32_t ql, const char *qs, orc_rst_t *r); /* miniwfa.c:603-615 */
void orc_wfa_auto_exact_leg(const orc_opt_t *opt, int32_t tl, const char *ts,
                   int32_t ql, const char *qs, orc_rst_t *r); /* miniwfa.c:898-902 (exact leg only) */
int32_t orc_cigar2score(const orc_opt_t *opt, int32_t n_cigar, const uint32_t *cigar,
                   int32_t *tl, int32_t *ql);               /* mwf-dbg.c:6-22 */
void orc_free(void *p);

/* low-memory pass 1 alone: returns malloc'ed (s,d) int32 pairs; *n_seg = number of checkpoints */
int32_t *orc_wfa_checkpoints(const orc_opt_t *opt, int32_t tl, const char *ts,
                   int32_t ql, const char *qs, int32_t *n_seg); /* miniwfa.c:551-601 */

#ifdef __cplusplus
}
#endif
#endif
