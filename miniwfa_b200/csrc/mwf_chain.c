/*
 * mwf_chain.c -- mwf_wfa_chain(): the k-mer chaining heuristic of the reference (miniwfa.c:620-896), host C.
 *
 * Same result as the reference (same anchors, same score, same CIGAR words): unique k-mer matches (k = opt->kmer,
 * at most opt->max_occ copies per sequence, :718-767) -> longest strictly increasing chain (:678-697, :769-783) ->
 * co-diagonal runs shorter than opt->min_len dropped (:829-848) -> the gaps between anchors filled (:861-891).
 *
 * What is different is *how* the gaps are filled.  The reference calls mwf_wfa_exact() once per gap, one after the
 * other (:877).  Here every gap that needs an exact alignment is collected first and the whole set goes to the GPU
 * engine as ONE batch (mwf_wfa_exact_batch, mwf_b200.h): many small independent alignments are exactly what the
 * engine is built for.  The gaps are independent, so the per-gap results -- and therefore the
 * concatenated CIGAR and the summed score -- are the same.
 *
 * Like the reference, this function does not touch r->n_iter (:850-896 never writes it).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include <stdio.h>
#include <time.h>
#include "miniwfa.h"
#include "mwf_b200.h"
#include "kalloc.h"

static double now_ms(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return 1e3 * (double)t.tv_sec + 1e-6 * (double)t.tv_nsec;
}

/* ---- CIGAR assembly (reference :46-62, :816-827) ---- */

typedef struct { uint32_t *w; int32_t n, cap; } cigbuf_t;

static void cig_reserve(void *km, cigbuf_t *c, int32_t extra)
{
	if (c->n + extra > c->cap) {
		int32_t cap = c->cap + (c->cap >> 1) + 4;
		if (cap < c->n + extra) cap = c->n + extra;
		c->w = (uint32_t*)krealloc(km, c->w, sizeof(uint32_t) * (size_t)cap);
		c->cap = cap;
	}
}

static void cig_add(void *km, cigbuf_t *c, int op, int32_t len) /* an operation equal to the last one is merged into it */
{
	if (c->n > 0 && (int)(c->w[c->n - 1] & 0xf) == op) c->w[c->n - 1] += (uint32_t)len << 4;
	else {
		cig_reserve(km, c, 1);
		c->w[c->n++] = (uint32_t)len << 4 | (uint32_t)op;
	}
}

static void cig_cat(void *km, cigbuf_t *c, int32_t n, const uint32_t *w) /* only the first word can merge */
{
	if (n <= 0) return;
	cig_add(km, c, (int)(w[0] & 0xf), (int32_t)(w[0] >> 4));
	cig_reserve(km, c, n - 1);
	memcpy(c->w + c->n, w + 1, sizeof(uint32_t) * (size_t)(n - 1));
	c->n += n - 1;
}

/* ---- k-mers (reference :699-730) ---- */

static inline int base_code(unsigned char ch) /* A/C/G/T(U) in either case, or the raw codes 0..3; anything else breaks a k-mer */
{
	switch (ch) {
	case 0: case 'A': case 'a': return 0;
	case 1: case 'C': case 'c': return 1;
	case 2: case 'G': case 'g': return 2;
	case 3: case 'T': case 't': case 'U': case 'u': return 3;
	default: return 4;
	}
}

/* every k-mer of seq as (k-mer << 1 | which) << 32 | position of its last base */
static int32_t list_kmers(int32_t len, const char *seq, int which, int k, uint64_t *out)
{
	const uint64_t mask = (1ULL << 2 * k) - 1;
	uint64_t word = 0;
	int32_t i, run = 0, n = 0;
	for (i = 0; i < len; ++i) {
		const int c = base_code((unsigned char)seq[i]);
		if (c > 3) { run = 0, word = 0; continue; }
		word = (word << 2 | (uint64_t)c) & mask;
		if (++run >= k) out[n++] = (word << 1 | (uint64_t)which) << 32 | (uint32_t)i;
	}
	return n;
}

/* ascending sort of distinct 64-bit keys: byte-wise LSD radix sort, passes whose byte is constant are skipped */
static void sort64(void *km, uint64_t *a, size_t n)
{
	uint64_t *tmp, *src = a, *dst;
	int pass;
	if (n < 2) return;
	tmp = (uint64_t*)kmalloc(km, sizeof(uint64_t) * n);
	dst = tmp;
	for (pass = 0; pass < 8; ++pass) {
		size_t cnt[256], i, sum = 0;
		const int sh = pass * 8;
		memset(cnt, 0, sizeof(cnt));
		for (i = 0; i < n; ++i) ++cnt[src[i] >> sh & 0xff];
		if (cnt[src[0] >> sh & 0xff] == n) continue;
		for (i = 0; i < 256; ++i) { const size_t c = cnt[i]; cnt[i] = sum; sum += c; }
		for (i = 0; i < n; ++i) dst[cnt[src[i] >> sh & 0xff]++] = src[i];
		{ uint64_t *t = src; src = dst; dst = t; }
	}
	if (src != a) memcpy(a, src, sizeof(uint64_t) * n);
	kfree(km, tmp);
}

/* longest strictly increasing subsequence of v[0..n): returns its *n_out values in a kmalloc'ed array (:678-697, :772-774).
 * tailv[l] / tail[l] = value / index of the smallest last element of an increasing chain of length l (tailv is strictly
 * increasing in l).  An element x goes behind the longest chain whose last value is below it, lo = max { l : tailv[l] < x }.
 * On similar sequences that chain is nearly always among the longest few, so the search first gallops down from the top; a
 * stray match (~10 % of the matches of a 3 %-divergent pair) lands anywhere in an array of tens of MB, and bisecting that costs
 * a dozen cache misses -- those go through samp[b] = tailv[64 b], which stays in cache, and then one 64-entry block.
 * Same lo as plain bisection, so the same chain as the reference's mg_lis_64. */
#define LIS_BLK 64
#define LIS_AHEAD 16 /* a power of two */
#define LIS_WS_MIN 65536

/* the largest index in [lo, hi) whose value is below x, where a[lo] counts as below x whatever it holds (lo may be the unused
 * slot 0); a[] increasing on (lo, hi).  Branch-free bisection: the outcome of each comparison is data, not control. */
static inline int32_t lis_last_below(const uint64_t *a, int32_t lo, int32_t hi, uint64_t x)
{
	int32_t n = hi - lo;
	while (n > 1) {
		const int32_t half = n >> 1;
		lo = a[lo + half] < x ? lo + half : lo;
		n -= half;
	}
	return lo;
}

static uint64_t *longest_increasing(void *km, int32_t n, const uint64_t *v, int32_t *n_out)
{
	int32_t *tail, *prev, i, len = 0, at, guess[LIS_AHEAD];
	uint64_t *tailv, *samp, *out;
	const size_t n_samp = (size_t)n / LIS_BLK + 2;
	/* Large inputs take the ~16 bytes per match of scratch from the library's cache of pinned host buffers: memory that is
	 * already resident, where a fresh kmalloc of tens of MB is paid for in page faults on every call. */
	void *ws = n >= LIS_WS_MIN ? mwf_b200_host_scratch(sizeof(uint64_t) * ((size_t)n + 1 + n_samp) + sizeof(int32_t) * (2 * (size_t)n + 2)) : 0;
	*n_out = 0;
	if (n <= 0) return 0;
	if (ws) {
		tailv = (uint64_t*)ws, samp = tailv + n + 1;
		tail = (int32_t*)(samp + n_samp), prev = tail + n + 1;
	} else {
		tail = (int32_t*)kmalloc(km, sizeof(int32_t) * ((size_t)n + 1));
		tailv = (uint64_t*)kmalloc(km, sizeof(uint64_t) * ((size_t)n + 1));
		samp = (uint64_t*)kmalloc(km, sizeof(uint64_t) * n_samp);
		prev = (int32_t*)kmalloc(km, sizeof(int32_t) * (size_t)n);
	}
	for (i = 0; i < LIS_AHEAD; ++i) guess[i] = -1;
	for (i = 0; i < n; ++i) {
		const uint64_t x = v[i];
		const int32_t g = guess[i & (LIS_AHEAD - 1)];
		int32_t lo = len;
		guess[i & (LIS_AHEAD - 1)] = -1;
		if (i + LIS_AHEAD < n && len >= 4 * LIS_BLK) { /* a stray match a few elements ahead: find and fetch the block it will land in */
			const uint64_t y = v[i + LIS_AHEAD];
			if (tailv[len - LIS_BLK] >= y) {
				const int32_t bl = lis_last_below(samp, 0, len / LIS_BLK + 1, y);
				guess[i & (LIS_AHEAD - 1)] = bl; /* (i + LIS_AHEAD) & (LIS_AHEAD - 1) is the same slot */
				__builtin_prefetch(&tailv[bl * LIS_BLK + LIS_BLK / 4]), __builtin_prefetch(&tailv[bl * LIS_BLK + 3 * LIS_BLK / 4]);
				__builtin_prefetch(&tail[bl * LIS_BLK + LIS_BLK / 4]), __builtin_prefetch(&tail[bl * LIS_BLK + 3 * LIS_BLK / 4]);
			}
		}
		if (len > 0 && tailv[len] >= x) {
			int32_t hi = len, step = 1; /* invariant: tailv[hi] >= x */
			lo = hi - 1;
			while (lo > 0 && tailv[lo] >= x && step < LIS_BLK) hi = lo, step <<= 1, lo = hi - step;
			if (lo > 0 && tailv[lo] >= x) { /* far below the top: the block first (samp[b] = tailv[LIS_BLK b], b = 1 .. lo / LIS_BLK) */
				const int32_t nb = lo / LIS_BLK;
				int32_t bl;
				hi = lo;
				if (g >= 0 && g <= nb && (g == 0 || samp[g] < x) && (g == nb || samp[g + 1] >= x)) bl = g; /* found while prefetching, still right */
				else bl = lis_last_below(samp, 0, nb + 1, x);
				lo = bl * LIS_BLK;
				if (lo + LIS_BLK < hi) hi = lo + LIS_BLK; /* = LIS_BLK (bl + 1), and samp[bl + 1] >= x */
			}
			if (lo < 0) lo = 0;
			lo = lis_last_below(tailv, lo, hi, x);
		}
		prev[i] = lo > 0 ? tail[lo] : -1;
		tail[lo + 1] = i, tailv[lo + 1] = x;
		if ((lo + 1) % LIS_BLK == 0) samp[(lo + 1) / LIS_BLK] = x;
		if (lo + 1 > len) len = lo + 1;
	}
	out = (uint64_t*)kmalloc(km, sizeof(uint64_t) * (size_t)len);
	for (i = len - 1, at = tail[len]; i >= 0; --i) out[i] = v[at], at = prev[at];
	if (ws) mwf_b200_host_scratch_free(ws);
	else kfree(km, prev), kfree(km, samp), kfree(km, tailv), kfree(km, tail);
	*n_out = len;
	return out;
}

/* Where the k-mer lists are built, sorted and matched: on the device (kmer_front.cuh) from FRONT_MIN_BASES bases up, on the
 * host below that (a few launches and two synchronisations cost more than sorting a few thousand keys here).
 * MWF_B200_CHAIN_FRONT=gpu|host forces one of the two (tests). */
#define FRONT_MIN_BASES 4096

static int front_on_device(int64_t bases)
{
	const char *v = getenv("MWF_B200_CHAIN_FRONT");
	if (v && !strcmp(v, "gpu")) return 1;
	if (v && !strcmp(v, "host")) return 0;
	return bases >= FRONT_MIN_BASES;
}

/* k-mer matches as (query position << 32 | target position), in ascending (target, query) order (:737-770), on the host */
static uint64_t *host_hits(void *km, int32_t tl, const char *ts, int32_t ql, const char *qs, int k, int max_occ, int32_t *n_out)
{
	uint64_t *km_list, *hit = 0;
	int32_t n_km, n_hit = 0, cap_hit = 0, i, g0;
	km_list = (uint64_t*)kmalloc(km, sizeof(uint64_t) * ((size_t)tl + ql));
	n_km = list_kmers(tl, ts, 0, k, km_list);
	n_km += list_kmers(ql, qs, 1, k, km_list + n_km);
	sort64(km, km_list, (size_t)n_km);
	for (g0 = 0, i = 1; i <= n_km; ++i) { /* groups of equal k-mers; inside a group the target copies come first */
		int32_t split, s, t;
		if (i < n_km && km_list[i] >> 33 == km_list[g0] >> 33) continue;
		for (split = g0; split < i && (km_list[split] >> 32 & 1) == 0; ++split) {}
		if (split > g0 && split < i && split - g0 <= max_occ && i - split <= max_occ)
			for (s = g0; s < split; ++s)
				for (t = split; t < i; ++t) {
					if (n_hit == cap_hit) {
						cap_hit = cap_hit ? cap_hit + (cap_hit >> 1) : 256;
						hit = (uint64_t*)krealloc(km, hit, sizeof(uint64_t) * (size_t)cap_hit);
					}
					hit[n_hit++] = km_list[s] << 32 | (uint32_t)km_list[t];
				}
		g0 = i;
	}
	kfree(km, km_list);
	sort64(km, hit, (size_t)n_hit); /* by target position, then query position */
	for (i = 0; i < n_hit; ++i) hit[i] = hit[i] >> 32 | hit[i] << 32; /* compare on (query, target) */
	*n_out = n_hit;
	return hit;
}

/* anchors (target position << 32 | query position, last base of the k-mer) of the best co-linear chain (:732-784) */
static uint64_t *chain_anchors(void *km, int32_t tl, const char *ts, int32_t ql, const char *qs, int k, int max_occ, int32_t *n_out)
{
	const int on_device = front_on_device((int64_t)tl + ql), timing = getenv("MWF_B200_CHAIN_TIMING") != 0;
	const double t0 = timing ? now_ms() : 0;
	double t1 = 0;
	uint64_t *hit = 0, *anchors;
	int32_t n_hit = 0, i, n_pick = 0;
	*n_out = 0;
	if (tl < k || ql < k) return 0;
	assert(k >= 2 && k <= 15);
	if (on_device) {
		const int64_t n = mwf_b200_kmer_hits(tl, ts, ql, qs, k, max_occ, &hit);
		assert(n <= INT32_MAX); /* the reference counts matches in an int32_t (:734) */
		n_hit = (int32_t)n;
	} else hit = host_hits(km, tl, ts, ql, qs, k, max_occ, &n_hit);
	if (timing) t1 = now_ms();
	anchors = longest_increasing(km, n_hit, hit, &n_pick);
	if (anchors == 0) anchors = (uint64_t*)kmalloc(km, sizeof(uint64_t));
	for (i = 0; i < n_pick; ++i) anchors[i] = anchors[i] >> 32 | anchors[i] << 32; /* back to target << 32 | query (:781-782) */
	if (on_device) mwf_b200_kmer_free(hit);
	else kfree(km, hit);
	*n_out = n_pick;
	if (timing) fprintf(stderr, "[mwf_chain] %d k-mer matches in %.2f ms (%s), longest increasing chain %.2f ms\n", n_hit, t1 - t0, on_device ? "device" : "host", now_ms() - t1);
	return anchors;
}

/* drop runs of co-diagonal anchors that span fewer than min_len bases (:829-848) */
static int32_t drop_short_runs(int32_t n, uint64_t *a, int32_t tl, int32_t ql, int k, int32_t min_len)
{
	int32_t i, j, m, run_start = -1, run_len = 0, ox = 0, oy = 0, px = 0;
	for (i = 0; i <= n; ++i) {
		const int32_t x = i == n ? tl : (int32_t)(a[i] >> 32) + 1, y = i == n ? ql : (int32_t)(uint32_t)a[i] + 1;
		if (x - ox != y - oy) { /* leaves the diagonal of the run's first anchor */
			if (run_len < min_len)
				for (j = run_start > 0 ? run_start : 0; j < i; ++j) a[j] = 0;
			ox = x, oy = y, run_start = i, run_len = k;
		} else run_len += x - px;
		px = x;
	}
	for (i = m = 0; i < n; ++i)
		if (a[i] != 0) a[m++] = a[i];
	return m;
}

/* fraction of shared k-mers, the larger of the two directions (:786-812) */
static double kmer_similarity(void *km, int32_t l1, const char *s1, int32_t l2, const char *s2, int k)
{
	uint64_t *a;
	int32_t n, i, g0, n1 = 0, n2 = 0, t1 = 0, t2 = 0;
	double p1, p2;
	if (l1 < k || l2 < k) return 0;
	if (front_on_device((int64_t)l1 + l2)) {
		int64_t c1, c2, shared;
		const double t0 = getenv("MWF_B200_CHAIN_TIMING") ? now_ms() : 0;
		mwf_b200_kmer_shared(l1, s1, l2, s2, k, &c1, &c2, &shared);
		if (t0 > 0) fprintf(stderr, "[mwf_chain] shared k-mers of a %d x %d gap in %.2f ms (device)\n", l1, l2, now_ms() - t0);
		p1 = (double)shared / (double)c1, p2 = (double)shared / (double)c2;
		return p1 > p2 ? p1 : p2;
	}
	a = (uint64_t*)kmalloc(km, sizeof(uint64_t) * ((size_t)l1 + l2));
	n = list_kmers(l1, s1, 0, k, a);
	n += list_kmers(l2, s2, 1, k, a + n);
	sort64(km, a, (size_t)n);
	for (g0 = 0, i = 1; i <= n; ++i) {
		int32_t split, c1, c2, shared;
		if (i < n && a[i] >> 33 == a[g0] >> 33) continue;
		for (split = g0; split < i && (a[split] >> 32 & 1) == 0; ++split) {}
		c1 = split - g0, c2 = i - split, shared = c1 < c2 ? c1 : c2;
		n1 += c1, n2 += c2;
		if (c1 > 0 && c2 > 0) t1 += shared, t2 += shared;
		g0 = i;
	}
	kfree(km, a);
	p1 = (double)t1 / n1, p2 = (double)t2 / n2;
	return p1 > p2 ? p1 : p2;
}

static int32_t one_gap_cost(const mwf_opt_t *opt, int32_t len)
{
	const int32_t a = opt->o2 + len * opt->e2, b = opt->o1 + len * opt->e1;
	return a < b ? a : b;
}

enum { SEG_MATCH, SEG_TWO_GAPS, SEG_EXACT, SEG_DEL, SEG_INS, SEG_NONE };

typedef struct { int32_t kind, x0, y0, x1, y1, job; } seg_t;

void mwf_wfa_chain(void *km, const mwf_opt_t *opt, int32_t tl, const char *ts, int32_t ql, const char *qs, mwf_rst_t *r)
{
	void *km_tmp = !(opt->flag & MWF_F_NO_KALLOC) ? km_init2(km, 0) : 0; /* scratch arena, as the reference's km_wfa (:857) */
	const int want_cigar = !!(opt->flag & MWF_F_CIGAR);
	const int timing = getenv("MWF_B200_CHAIN_TIMING") != 0; /* phase times to stderr */
	const double t0 = timing ? now_ms() : 0;
	double t1 = 0, t2 = 0, t3 = 0;
	int32_t n_a, i, x0 = 0, y0 = 0, n_job = 0, score = 0;
	uint64_t *a = chain_anchors(km_tmp, tl, ts, ql, qs, opt->kmer, opt->max_occ, &n_a);
	seg_t *seg = 0;
	int32_t n_seg = 0, cap_seg = 0;
	cigbuf_t c = { 0, 0, 0 };

	if (timing) t1 = now_ms();

	n_a = drop_short_runs(n_a, a, tl, ql, opt->kmer, opt->min_len);
	for (i = 0; i <= n_a; ++i) { /* classify what lies between consecutive anchors (:861-889) */
		const int32_t x1 = i == n_a ? tl : (int32_t)(a[i] >> 32) + 1, y1 = i == n_a ? ql : (int32_t)(uint32_t)a[i] + 1;
		int32_t kind, job = -1;
		if (i < n_a && x1 - x0 == y1 - y0 && x1 - x0 <= opt->kmer) kind = SEG_MATCH; /* inside overlapping k-mer matches */
		else if (x0 < x1 && y0 < y1) {
			if (x1 - x0 >= 10000 && y1 - y0 >= 10000 && kmer_similarity(km, x1 - x0, ts + x0, y1 - y0, qs + y0, opt->kmer) < 0.02)
				kind = SEG_TWO_GAPS; /* two long unrelated stretches: a deletion and an insertion (:869-874) */
			else kind = SEG_EXACT, job = n_job++;
		} else if (x0 < x1) kind = SEG_DEL;
		else if (y0 < y1) kind = SEG_INS;
		else kind = SEG_NONE;
		/* most anchors are the next base of the same match (millions on a Mb-scale pair): runs of SEG_MATCH become one
		 * segment -- their CIGAR operations would be merged by cig_add anyway -- and empty segments are not stored */
		if (kind == SEG_MATCH && n_seg > 0 && seg[n_seg - 1].kind == SEG_MATCH) seg[n_seg - 1].x1 = x1, seg[n_seg - 1].y1 = y1;
		else if (kind != SEG_NONE) {
			seg_t *g;
			if (n_seg == cap_seg) {
				cap_seg = cap_seg ? cap_seg + (cap_seg >> 1) : 1024;
				seg = (seg_t*)krealloc(km_tmp, seg, sizeof(seg_t) * (size_t)cap_seg);
			}
			g = &seg[n_seg++];
			g->kind = kind, g->x0 = x0, g->y0 = y0, g->x1 = x1, g->y1 = y1, g->job = job;
		}
		x0 = x1, y0 = y1;
	}
	{ /* all exact gap fills in one submission (the reference loops over mwf_wfa_exact, :877) */
		int32_t *jtl = (int32_t*)kmalloc(km_tmp, sizeof(int32_t) * (size_t)(n_job + 1)), *jql = (int32_t*)kmalloc(km_tmp, sizeof(int32_t) * (size_t)(n_job + 1));
		const char **jts = (const char**)kmalloc(km_tmp, sizeof(char*) * (size_t)(n_job + 1)), **jqs = (const char**)kmalloc(km_tmp, sizeof(char*) * (size_t)(n_job + 1));
		mwf_rst_t *jr = (mwf_rst_t*)kcalloc(km_tmp, (size_t)(n_job + 1), sizeof(mwf_rst_t));
		for (i = 0; i < n_seg; ++i)
			if (seg[i].kind == SEG_EXACT) {
				const int32_t j = seg[i].job;
				jtl[j] = seg[i].x1 - seg[i].x0, jts[j] = ts + seg[i].x0;
				jql[j] = seg[i].y1 - seg[i].y0, jqs[j] = qs + seg[i].y0;
			}
		if (timing) t2 = now_ms();
		if (n_job > 0) mwf_wfa_exact_batch(km_tmp, opt, n_job, jtl, jts, jql, jqs, jr);
		if (timing) t3 = now_ms();
		for (i = 0; i < n_seg; ++i) { /* concatenate in order */
			const seg_t *g = &seg[i];
			const int32_t dx = g->x1 - g->x0, dy = g->y1 - g->y0;
			switch (g->kind) {
			case SEG_MATCH:
				if (want_cigar) cig_add(km, &c, 7, dx);
				break;
			case SEG_TWO_GAPS:
				if (want_cigar) cig_add(km, &c, 2, dx), cig_add(km, &c, 1, dy);
				score += opt->o2 * 2 + opt->e2 * (dx + dy);
				break;
			case SEG_EXACT:
				if (want_cigar) cig_cat(km, &c, jr[g->job].n_cigar, jr[g->job].cigar);
				score += jr[g->job].s;
				kfree(km_tmp, jr[g->job].cigar);
				break;
			case SEG_DEL: /* the reference records these two even in score-only mode (:883, :886) */
				cig_add(km, &c, 2, dx);
				score += one_gap_cost(opt, dx);
				break;
			case SEG_INS:
				cig_add(km, &c, 1, dy);
				score += one_gap_cost(opt, dy);
				break;
			default: break;
			}
		}
		kfree(km_tmp, jr); kfree(km_tmp, (void*)jqs); kfree(km_tmp, (void*)jts); kfree(km_tmp, jql); kfree(km_tmp, jtl);
	}
	kfree(km_tmp, seg);
	kfree(km_tmp, a);
	if (km_tmp) km_destroy(km_tmp);
	r->s = score;
	r->n_cigar = c.n;
	r->cigar = (uint32_t*)krelocate(km, c.w, sizeof(uint32_t) * (size_t)c.n);
	if (timing)
		fprintf(stderr, "[mwf_chain] %d x %d: anchors %.2f ms (%d), classify %.2f ms, %d gap fills %.2f ms, assemble %.2f ms\n",
		        tl, ql, t1 - t0, n_a, t2 - t1, n_job, t3 - t2, now_ms() - t3);
}
