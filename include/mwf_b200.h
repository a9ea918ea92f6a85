/*
 * mwf_b200.h -- the C-ABI boundary between the host C driver (miniwfa.c) and the
 * sm_100a CUDA engine (wfa_engine.cu).  Plain pointers and sizes only.
 *
 * What each entry point replaces in the reference (/root/reference/miniwfa.c @ 66770a3):
 *
 *   mwf_b200_batch_create   wf_stripe_init (:103-121), the km_init2 scratch arenas (:388-389)
 *                           and wf_tb_add's per-score kcalloc (:33-44): ring, traceback and
 *                           snapshot storage become HBM workspaces sized once per batch.
 *   mwf_b200_batch_upload   wf_pad_str (:182-209): both sequences of every pair are staged
 *                           once, 16-byte aligned with zeroed slack, into HBM.  No sentinel
 *                           bytes are needed (the kernel clamps the match run to the matrix),
 *                           so inputs using all 256 byte values are accepted.
 *   mwf_b200_batch_run      (tile engine: wfa_tile.cuh; streaming kernels: wfa_engine.cu)
 *                           the whole of mwf_wfa_core's score loop (:397-426) -- the extend
 *                           loop + wf_extend1_padded (:400-411, :212-226), wf_next_basic /
 *                           wf_next_prep / wf_next_score / wf_next_tb (:243-327),
 *                           wf_stripe_shrink (:144-171), the checkpoint band collapse
 *                           (:413-416) -- plus, when opt.step > 0, mwf_wfa_seg's pass 1
 *                           (wf_next_seg :495-526, wf_snapshot :451-483, wf_traceback_seg
 *                           :528-549) and finally wf_traceback (:329-377), all on the device.
 *   mwf_b200_batch_fetch    the tail of mwf_wfa_core (:427-434): fills mwf_rst_t and
 *                           allocates r->cigar from the caller's km.
 *   mwf_wfa_exact_batch     NEW (the reference has no batch call; its CLI loops, main.c:67):
 *                           n independent pairs in one submission.  mwf_wfa_exact() is this
 *                           with n = 1.
 *
 * Errors: like the reference (assert/abort, kalloc.c:32-36), any CUDA failure or an
 * exhausted device workspace prints a message to stderr and aborts.  There is no CPU
 * fallback: without a usable GPU the library aborts with "no CUDA device".
 */
#ifndef MWF_B200_H
#define MWF_B200_H

#include <stdint.h>
#include <stddef.h>
#include "miniwfa.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mwf_b200_batch mwf_b200_batch_t;

/* kernel families; AUTO picks per batch:
 *   TILE : temporally blocked -- ring tiles resident in shared memory for up to 64 scores (high-memory and score-only modes)
 *   CTA  : streaming, one CTA per pair (small pairs, low-memory mode)
 *   GRID : streaming, the whole grid on one pair (few large pairs in low-memory mode) */
#define MWF_B200_KERNEL_AUTO 0
#define MWF_B200_KERNEL_CTA  1
#define MWF_B200_KERNEL_GRID 2
#define MWF_B200_KERNEL_TILE 3

int  mwf_b200_device_count(void);
/* device used by batches created afterwards on any thread; default: $MWF_B200_DEVICE, else $LOCAL_RANK, else 0 */
void mwf_b200_set_device(int dev);
int  mwf_b200_get_device(void);
/* devices one mwf_wfa_exact_batch() call is spread over: 1 = the current device only; n > 1 = n devices starting at the current
 * one; 0 (default) = every visible device ($MWF_B200_DEVICES caps them) unless a device was pinned by mwf_b200_set_device(),
 * $MWF_B200_DEVICE or $LOCAL_RANK (one process per GPU), and only for batches worth it (sum of squared lengths >= 1e11).  Pairs are
 * dealt out by cost; no data moves between devices. */
void mwf_b200_set_devices(int n);
/* force a kernel family (tests/bench); default: $MWF_B200_KERNEL ("cta"/"grid"/"tile"), else AUTO */
void mwf_b200_set_kernel(int kernel);
/* device and pinned-host workspaces are cached across batches; this frees every cached buffer */
void mwf_b200_release_cache(void);
/* threads per CTA for subsequently created batches (0 = default) */
void mwf_b200_set_block_threads(int threads);

mwf_b200_batch_t *mwf_b200_batch_create(const mwf_opt_t *opt, int32_t n_pairs, const int32_t *tl, const int32_t *ql);
/* run on this CUDA stream (a cudaStream_t) instead of the batch's own */
void mwf_b200_batch_set_stream(mwf_b200_batch_t *b, void *cuda_stream);
void mwf_b200_batch_upload(mwf_b200_batch_t *b, const char *const *ts, const char *const *qs); /* host -> pinned -> HBM, async */
void mwf_b200_batch_run(mwf_b200_batch_t *b);     /* run the alignment kernels on the batch's stream.  The tile engine runs a pass as one
                                                      persistent kernel and reads its outcome back, so this returns when the batch is
                                                      (nearly) done; the streaming kernels are only enqueued */
void mwf_b200_batch_wait(mwf_b200_batch_t *b);    /* block until the stream is idle; aborts on a device-side error */
void mwf_b200_batch_fetch(mwf_b200_batch_t *b, void *km, mwf_rst_t *r); /* HBM -> host; r[0..n_pairs) */
void mwf_b200_batch_destroy(mwf_b200_batch_t *b);

/* measurements of the last mwf_b200_batch_run (valid after _wait) */
double  mwf_b200_batch_kernel_ms(const mwf_b200_batch_t *b); /* CUDA-event time over the alignment kernels */
int64_t mwf_b200_batch_launches(const mwf_b200_batch_t *b);  /* kernels launched by the last run */
int     mwf_b200_batch_kernel_used(const mwf_b200_batch_t *b); /* MWF_B200_KERNEL_CTA, _GRID or _TILE */
int64_t mwf_b200_batch_h2d_bytes(const mwf_b200_batch_t *b);
int64_t mwf_b200_batch_d2h_bytes(const mwf_b200_batch_t *b);

/* n independent exact alignments; r[i] as mwf_wfa_exact would fill it */
void mwf_wfa_exact_batch(void *km, const mwf_opt_t *opt, int32_t n_pairs,
                         const int32_t *tl, const char *const *ts,
                         const int32_t *ql, const char *const *qs, mwf_rst_t *r);

/* The k-mer front end of mwf_wfa_chain on the device (kmer_front.cuh); sequences are host buffers, as everywhere in this API.
 *
 * mwf_b200_kmer_hits replaces mg_fc_kmer x 2 + radix_sort_mwf64 + the match loop + radix_sort_mwf64 + the word swap of
 * mg_chain (miniwfa.c:737-770): every (target position, query position) pair of a shared k-mer that has at most max_occ
 * copies in either sequence, as query << 32 | target (positions of the k-mer's last base), in ascending (target, query)
 * order -- the array mg_lis_64 runs on.  Returns their number; *hits is pinned host memory owned by the library
 * (NULL when there are none), to be released with mwf_b200_kmer_free. */
int64_t mwf_b200_kmer_hits(int32_t tl, const char *ts, int32_t ql, const char *qs, int32_t k, int32_t max_occ, uint64_t **hits);
void    mwf_b200_kmer_free(uint64_t *hits);
/* The counting part of mwf_ksim (miniwfa.c:786-812): k-mers in s1, k-mers in s2, sum over distinct k-mers of
 * min(copies in s1, copies in s2).  The caller does the two divisions. */
void    mwf_b200_kmer_shared(int32_t l1, const char *s1, int32_t l2, const char *s2, int32_t k, int64_t *n1, int64_t *n2, int64_t *shared);
/* host scratch from the library's cache of pinned buffers (resident memory: no page faults on reuse); NULL is never returned */
void   *mwf_b200_host_scratch(size_t bytes);
void    mwf_b200_host_scratch_free(void *p);
int64_t mwf_b200_kmer_launches(void); /* kernels launched by the two calls above since the library was loaded */

#ifdef __cplusplus
}
#endif
#endif
