/*
 * oracle/wfa_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see wfa_oracle.h).
 *
 * Scalar C restatement of the reference's exact dual-affine WFA
 * (/root/reference/miniwfa.c @ 66770a3).  Parity: PINNED against the golden
 * vectors in tests/golden/ and against oracle/_ref (the unmodified reference).
 *
 * Deliberately written differently from the reference so that it is an
 * independent check rather than a copy:
 *   - every wavefront ("front") owns full-width arrays and an explicit [lo,hi];
 *     a read outside that range yields NEG through rd(), instead of the
 *     reference's NEG_INF padding cells (miniwfa.c:96-99);
 *   - match extension is a bounded byte loop, not the padded 8-byte XOR/ctz
 *     probe (miniwfa.c:182-226); both compute the same longest common prefix
 *     because the reference's pad bytes can never match;
 *   - no arena allocator: plain malloc/free.
 * Every function cites the reference lines it restates.
 */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "wfa_oracle.h"

#define NEG (-0x40000000) /* WF_NEG_INF, miniwfa.c:67 */

/* array order = snapshot flatten order H,E1,F1,E2,F2 (miniwfa.c:466-470) */
enum { A_H = 0, A_E1 = 1, A_F1 = 2, A_E2 = 3, A_F2 = 4 };

typedef struct {
	int32_t lo, hi;
	int32_t *v[5];
} front_t;

typedef struct {
	int32_t n, top, s, lo, hi; /* wf_stripe_t, miniwfa.c:74-77 */
	int32_t off;               /* v[a][d + off] */
	front_t *f;
} ring_t;

static inline int32_t imax(int32_t a, int32_t b) { return a >= b ? a : b; }

static inline int32_t rd(const front_t *f, int a, int32_t d, int32_t off)
{
	return (d < f->lo || d > f->hi) ? NEG : f->v[a][d + off];
}

void orc_free(void *p) { free(p); }

void orc_opt_init(orc_opt_t *opt) /* miniwfa.c:11-18 */
{
	memset(opt, 0, sizeof(*opt));
	opt->x = 4;
	opt->o1 = 4, opt->e1 = 2;
	opt->o2 = 15, opt->e2 = 1;
	opt->kmer = 13, opt->max_occ = 2, opt->min_len = 30;
}

static int32_t max_penalty(const orc_opt_t *o) /* miniwfa.c:390-392 */
{
	int32_t m = o->x;
	m = imax(m, o->o1 + o->e1);
	m = imax(m, o->o2 + o->e2);
	return m;
}

/* wf_stripe_init, miniwfa.c:103-121: max_pen+1 fronts [0,0] holding NEG; the newest holds H[0] = -1 */
static ring_t *ring_new(int32_t max_pen, int32_t tl, int32_t ql)
{
	ring_t *w = (ring_t*)calloc(1, sizeof(ring_t));
	int32_t i, a;
	size_t cap = (size_t)tl + ql + 1;
	w->n = max_pen + 1;
	w->off = tl;
	w->f = (front_t*)calloc(w->n, sizeof(front_t));
	for (i = 0; i < w->n; ++i)
		for (a = 0; a < 5; ++a) {
			w->f[i].v[a] = (int32_t*)malloc(cap * sizeof(int32_t));
			w->f[i].v[a][w->off] = NEG;
		}
	w->top = 0, w->s = 0, w->lo = w->hi = 0;
	w->f[w->top].v[A_H][w->off] = -1;
	return w;
}

static void ring_free(ring_t *w)
{
	int32_t i, a;
	for (i = 0; i < w->n; ++i)
		for (a = 0; a < 5; ++a) free(w->f[i].v[a]);
	free(w->f); free(w);
}

/* wf_stripe_add, miniwfa.c:79-101 (without the pad fill: rd() range-checks instead) */
static front_t *ring_push(ring_t *w, int32_t lo, int32_t hi)
{
	front_t *f;
	w->s++;
	w->top = (w->top + 1) % w->n;
	f = &w->f[w->top];
	f->lo = lo, f->hi = hi;
	return f;
}

/* wf_stripe_get, miniwfa.c:132-137: the front of score (s - back) */
static const front_t *ring_back(const ring_t *w, int32_t back)
{
	return &w->f[((w->top - back) % w->n + w->n) % w->n];
}

static int on_matrix(int32_t d, int32_t k, int32_t tl, int32_t ql) /* good_diag, miniwfa.c:139-142 */
{
	return k >= -1 && k < tl && d + k >= -1 && d + k < ql;
}

/* does diagonal d hold an on-matrix value in any front of the ring?  (miniwfa.c:148-154) */
static int diag_alive(const ring_t *w, int32_t d, int32_t tl, int32_t ql)
{
	int32_t j, a;
	for (j = 0; j < w->n; ++j) {
		const front_t *f = &w->f[j];
		if (d < f->lo || d > f->hi) continue;
		for (a = 0; a < 5; ++a)
			if (on_matrix(d, f->v[a][d + w->off], tl, ql)) return 1;
	}
	return 0;
}

/* wf_stripe_shrink, miniwfa.c:144-171 */
static void ring_shrink(ring_t *w, int32_t tl, int32_t ql)
{
	int32_t d;
	for (d = w->lo; d <= w->hi; ++d)
		if (diag_alive(w, d, tl, ql)) break;
	assert(d <= w->hi);
	w->lo = d;
	for (d = w->hi; d >= w->lo; --d)
		if (diag_alive(w, d, tl, ql)) break;
	assert(d >= w->lo);
	w->hi = d;
}

/* edge rule, miniwfa.c:325-326 (and :524-525): keep the new edge iff any state reaches >= -1 there */
static void edge_rule(ring_t *band_owner, const front_t *f, int32_t off)
{
	int32_t a, lo_ok = 0, hi_ok = 0;
	for (a = 0; a < 5; ++a) {
		if (f->v[a][f->lo + off] >= -1) lo_ok = 1;
		if (f->v[a][f->hi + off] >= -1) hi_ok = 1;
	}
	if (lo_ok) band_owner->lo = f->lo;
	if (hi_ok) band_owner->hi = f->hi;
}

/*
 * One score step: wf_next_prep + wf_next_tb (miniwfa.c:243-259, 281-308).  wf_next_score
 * (:261-279) is the same arithmetic without the byte, so one routine serves both.
 * tb (may be NULL) is indexed by d - lo.
 */
static front_t *front_next(const orc_opt_t *o, ring_t *w, int32_t lo, int32_t hi, uint8_t *tb)
{
	front_t *nf = ring_push(w, lo, hi);
	const front_t *fx = ring_back(w, o->x);
	const front_t *fo1 = ring_back(w, o->o1 + o->e1), *fo2 = ring_back(w, o->o2 + o->e2);
	const front_t *fe1 = ring_back(w, o->e1), *fe2 = ring_back(w, o->e2);
	int32_t d, off = w->off;
	for (d = lo; d <= hi; ++d) {
		int32_t open1_i = rd(fo1, A_H, d - 1, off), ext1_i = rd(fe1, A_E1, d - 1, off);
		int32_t open2_i = rd(fo2, A_H, d - 1, off), ext2_i = rd(fe2, A_E2, d - 1, off);
		int32_t open1_d = rd(fo1, A_H, d + 1, off), ext1_d = rd(fe1, A_F1, d + 1, off);
		int32_t open2_d = rd(fo2, A_H, d + 1, off), ext2_d = rd(fe2, A_F2, d + 1, off);
		int32_t e1 = imax(open1_i, ext1_i), e2 = imax(open2_i, ext2_i);
		int32_t f1 = imax(open1_d, ext1_d) + 1, f2 = imax(open2_d, ext2_d) + 1;
		int32_t e = imax(e1, e2), f = imax(f1, f2), g = imax(e, f);
		int32_t hx = rd(fx, A_H, d, off) + 1;
		uint8_t bits = 0, state;
		if (open1_i < ext1_i) bits |= 0x08; /* E1 extended rather than opened */
		if (open1_d < ext1_d) bits |= 0x10; /* F1 */
		if (open2_i < ext2_i) bits |= 0x20; /* E2 */
		if (open2_d < ext2_d) bits |= 0x40; /* F2 */
		if (hx >= g) state = 0;             /* mismatch wins ties (miniwfa.c:304) */
		else if (e >= f) state = e1 >= e2 ? 1 : 3;
		else state = f1 >= f2 ? 2 : 4;
		nf->v[A_E1][d + off] = e1, nf->v[A_E2][d + off] = e2;
		nf->v[A_F1][d + off] = f1, nf->v[A_F2][d + off] = f2;
		nf->v[A_H][d + off] = imax(hx, g);
		if (tb) tb[d - lo] = bits | state;
	}
	return nf;
}

/* second loop of wf_next_seg, miniwfa.c:503-523: replay the recorded choices on the provenance ring */
static front_t *front_replay(const orc_opt_t *o, ring_t *sf, int32_t lo, int32_t hi, const uint8_t *tb)
{
	front_t *nf = ring_push(sf, lo, hi);
	const front_t *fx = ring_back(sf, o->x);
	const front_t *fo1 = ring_back(sf, o->o1 + o->e1), *fo2 = ring_back(sf, o->o2 + o->e2);
	const front_t *fe1 = ring_back(sf, o->e1), *fe2 = ring_back(sf, o->e2);
	int32_t d, off = sf->off;
	for (d = lo; d <= hi; ++d) {
		uint8_t x = tb[d - lo];
		int32_t e1 = (x & 0x08) ? rd(fe1, A_E1, d - 1, off) : rd(fo1, A_H, d - 1, off);
		int32_t f1 = (x & 0x10) ? rd(fe1, A_F1, d + 1, off) : rd(fo1, A_H, d + 1, off);
		int32_t e2 = (x & 0x20) ? rd(fe2, A_E2, d - 1, off) : rd(fo2, A_H, d - 1, off);
		int32_t f2 = (x & 0x40) ? rd(fe2, A_F2, d + 1, off) : rd(fo2, A_H, d + 1, off);
		int32_t h;
		switch (x & 7) {
			case 1: h = e1; break;
			case 2: h = f1; break;
			case 3: h = e2; break;
			case 4: h = f2; break;
			default: h = rd(fx, A_H, d, off);
		}
		nf->v[A_E1][d + off] = e1, nf->v[A_F1][d + off] = f1;
		nf->v[A_E2][d + off] = e2, nf->v[A_F2][d + off] = f2;
		nf->v[A_H][d + off] = h;
	}
	return nf;
}

/* longest common prefix of ts[k+1..] and qs[d+k+1..]; equals wf_extend1_padded (miniwfa.c:212-226) */
static int32_t extend(int32_t tl, const char *ts, int32_t ql, const char *qs, int32_t k, int32_t d)
{
	while (k + 1 < tl && d + k + 1 < ql && ts[k + 1] == qs[d + k + 1]) ++k;
	return k;
}

/* traceback rows, one per score >= 1 (miniwfa.c:23-44) */
typedef struct { int32_t lo, hi; uint8_t *x; } tbrow_t;
typedef struct { int32_t n, m; tbrow_t *a; } tbrows_t;

static uint8_t *tb_new_row(tbrows_t *t, int32_t lo, int32_t hi)
{
	tbrow_t *p;
	if (t->n == t->m) {
		t->m = t->m ? t->m * 2 : 64;
		t->a = (tbrow_t*)realloc(t->a, t->m * sizeof(tbrow_t));
	}
	p = &t->a[t->n++];
	p->lo = lo, p->hi = hi;
	p->x = (uint8_t*)calloc((size_t)hi - lo + 1, 1);
	return p->x;
}

typedef struct { int32_t n, m; uint32_t *a; } cig_t;

static void cig_add(cig_t *c, uint32_t op, uint32_t len) /* wf_cigar_push1, miniwfa.c:51-62 */
{
	if (c->n > 0 && (c->a[c->n - 1] & 0xf) == op) { c->a[c->n - 1] += len << 4; return; }
	if (c->n == c->m) {
		c->m = c->m ? c->m * 2 : 16;
		c->a = (uint32_t*)realloc(c->a, c->m * sizeof(uint32_t));
	}
	c->a[c->n++] = len << 4 | op;
}

/* wf_traceback, miniwfa.c:329-377 */
static uint32_t *traceback(const orc_opt_t *o, const tbrows_t *t, int32_t tl, const char *ts,
                           int32_t ql, const char *qs, int32_t last, int32_t *n_cigar)
{
	static const uint32_t op_of[5] = { 8, 1, 2, 1, 2 }; /* X, I, D, I, D */
	cig_t c = { 0, 0, 0 };
	int32_t i = ql - 1, k = tl - 1, row = t->n - 1, a, b;
	const int32_t open_cost[5] = { o->x, o->o1 + o->e1, o->o1 + o->e1, o->o2 + o->e2, o->o2 + o->e2 };
	const int32_t ext_cost[5]  = { o->x, o->e1, o->e1, o->e2, o->e2 };
	while (i >= 0 && k >= 0) {
		int32_t state, ext, x;
		if (last == 0) { /* a run of matches may precede the next edit (:335-341) */
			int32_t run = 0;
			while (i >= 0 && k >= 0 && qs[i] == ts[k]) --i, --k, ++run;
			if (run > 0) cig_add(&c, 7, run);
			if (i < 0 || k < 0) break;
		}
		assert(row >= 0);
		assert(i - k >= t->a[row].lo && i - k <= t->a[row].hi);
		x = t->a[row].x[i - k - t->a[row].lo];
		state = last == 0 ? (x & 7) : last;
		assert(state >= 0 && state <= 4);
		ext = state > 0 ? (x >> (state + 2)) & 1 : 0;
		cig_add(&c, op_of[state], 1);
		if (state == 0) --i, --k;
		else if (state == 1 || state == 3) --i;
		else --k;
		row -= ext ? ext_cost[state] : open_cost[state];
		last = (state > 0 && ext) ? state : 0;
	}
	if (i >= 0) cig_add(&c, 1, i + 1);       /* leading insertion (:368) */
	else if (k >= 0) cig_add(&c, 2, k + 1);  /* leading deletion  (:369) */
	for (a = 0, b = c.n - 1; a < b; ++a, --b) { uint32_t t_ = c.a[a]; c.a[a] = c.a[b]; c.a[b] = t_; }
	*n_cigar = c.n;
	return c.a;
}

/* mwf_wfa_core, miniwfa.c:380-435.  seg = (s,d) pairs from pass 1 or NULL. */
static void core(const orc_opt_t *o, int32_t tl, const char *ts, int32_t ql, const char *qs,
                 int32_t n_seg, const int32_t *seg, orc_rst_t *r)
{
	int32_t is_tb = !!(o->flag & ORC_F_CIGAR), last_state = 0, stopped = 0, sid = 0, i;
	ring_t *w = ring_new(max_penalty(o), tl, ql);
	tbrows_t t = { 0, 0, 0 };
	memset(r, 0, sizeof(*r));
	for (;;) {
		front_t *p = &w->f[w->top];
		int32_t d, lo, hi, done = 0;
		for (d = p->lo; d <= p->hi; ++d) { /* extend loop, miniwfa.c:400-411 */
			int32_t k0 = p->v[A_H][d + w->off], k;
			if (k0 < -1 || d + k0 < -1 || k0 >= tl || d + k0 >= ql) continue;
			k = extend(tl, ts, ql, qs, k0, d);
			if (k == tl - 1 && d + k == ql - 1) {
				/* the reference reads tb.a[-1] when tl==ql==0 with CIGAR on (UB); we define last_state=0 there */
				if (k == k0 && is_tb && t.n > 0)
					last_state = t.a[t.n - 1].x[d - t.a[t.n - 1].lo] & 7;
				done = 1;
				break;
			}
			p->v[A_H][d + w->off] = k;
		}
		if (done) break;
		if (is_tb && seg && sid < n_seg && seg[2 * sid] == w->s) { /* band collapse, :413-416 */
			assert(seg[2 * sid + 1] >= w->lo && seg[2 * sid + 1] <= w->hi);
			w->lo = w->hi = seg[2 * sid + 1];
			++sid;
		}
		lo = w->lo > -tl ? w->lo - 1 : -tl;
		hi = w->hi < ql ? w->hi + 1 : ql;
		p = front_next(o, w, lo, hi, is_tb ? tb_new_row(&t, lo, hi) : 0);
		edge_rule(w, p, w->off);
		if ((w->s & 0xff) == 0) ring_shrink(w, tl, ql);
		r->n_iter += hi - lo + 1;
		if ((o->max_iter > 0 && r->n_iter > o->max_iter) || (o->max_s > 0 && w->s > o->max_s)) {
			stopped = 1;
			break;
		}
	}
	r->s = stopped ? -1 : w->s;
	if (is_tb && !stopped)
		r->cigar = traceback(o, &t, tl, ts, ql, qs, last_state, &r->n_cigar);
	for (i = 0; i < t.n; ++i) free(t.a[i].x);
	free(t.a);
	ring_free(w);
}

/* snapshots of the provenance ring, miniwfa.c:440-493 */
typedef struct { int32_t n, n_intv, max_s; int32_t *x; int32_t *ilo, *icnt; } snap_t;

static void snapshot(ring_t *sf, snap_t *ss) /* wf_snapshot1, miniwfa.c:451-474 */
{
	int32_t j, t = 0, d, a;
	ss->n = 0, ss->max_s = sf->s, ss->n_intv = sf->n;
	for (j = 0; j < sf->n; ++j) ss->n += 5 * (sf->f[j].hi - sf->f[j].lo + 1);
	ss->x = (int32_t*)malloc((size_t)ss->n * sizeof(int32_t));
	ss->ilo = (int32_t*)malloc(sf->n * sizeof(int32_t));
	ss->icnt = (int32_t*)malloc(sf->n * sizeof(int32_t));
	for (j = 0; j < sf->n; ++j) { /* oldest front first */
		front_t *f = &sf->f[(sf->top + 1 + j) % sf->n];
		ss->ilo[j] = f->lo, ss->icnt[j] = 5 * (f->hi - f->lo + 1);
		for (d = f->lo; d <= f->hi; ++d)
			for (a = 0; a < 5; ++a) { /* remember the old label, relabel the cell with its own index */
				ss->x[t] = f->v[a][d + sf->off];
				f->v[a][d + sf->off] = t++;
			}
	}
	assert(t == ss->n);
}

/* mwf_wfa_seg + wf_traceback_seg, miniwfa.c:528-601.  Returns (s,d) pairs. */
int32_t *orc_wfa_checkpoints(const orc_opt_t *o, int32_t tl, const char *ts, int32_t ql, const char *qs, int32_t *n_seg_)
{
	int32_t max_pen = max_penalty(o), last = NEG, n_snap = 0, m_snap = 0, j, *seg;
	ring_t *w = ring_new(max_pen, tl, ql), *sf = ring_new(max_pen, tl, ql);
	uint8_t *xbuf = (uint8_t*)calloc((size_t)tl + ql + 1, 1);
	snap_t *snaps = 0;
	for (;;) {
		front_t *p = &w->f[w->top], *q;
		int32_t d, lo, hi, done = 0;
		for (d = p->lo; d <= p->hi; ++d) { /* miniwfa.c:572-581 */
			int32_t k0 = p->v[A_H][d + w->off], k;
			if (k0 < -1 || d + k0 < -1 || k0 >= tl || d + k0 >= ql) continue;
			k = extend(tl, ts, ql, qs, k0, d);
			if (k == tl - 1 && d + k == ql - 1) {
				last = sf->f[sf->top].v[A_H][d + sf->off];
				done = 1;
				break;
			}
			p->v[A_H][d + w->off] = k;
		}
		if (done) break;
		lo = w->lo > -tl ? w->lo - 1 : -tl;
		hi = w->hi < ql ? w->hi + 1 : ql;
		if ((w->s + 1) % o->step == 0) { /* miniwfa.c:585-586 */
			if (n_snap == m_snap) {
				m_snap = m_snap ? m_snap * 2 : 8;
				snaps = (snap_t*)realloc(snaps, m_snap * sizeof(snap_t));
			}
			snapshot(sf, &snaps[n_snap++]);
		}
		front_next(o, w, lo, hi, xbuf);
		q = front_replay(o, sf, lo, hi, xbuf);
		/* NB (reference quirk, miniwfa.c:524-525): in pass 1 the edge rule tests the PROVENANCE
		 * values, which are >= 0 for every cell that descends from a snapshotted cell, so after
		 * the first snapshot the band widens on almost every step.  Not observable in the result. */
		edge_rule(w, q, sf->off);
		if ((w->s & 0xff) == 0) ring_shrink(w, tl, ql);
	}
	/* wf_traceback_seg, miniwfa.c:528-549: chase labels backwards through the snapshots */
	seg = (int32_t*)malloc((n_snap ? n_snap : 1) * 2 * sizeof(int32_t));
	for (j = n_snap - 1; j >= 0; --j) {
		snap_t *p = &snaps[j];
		int32_t k, m = 0;
		for (k = 0; k < p->n_intv; ++k) {
			if (last >= m && last < m + p->icnt[k]) break;
			m += p->icnt[k];
		}
		assert(k < p->n_intv);
		seg[2 * j] = p->max_s - (p->n_intv - k - 1);
		seg[2 * j + 1] = p->ilo[k] + (last - m) / 5;
		last = p->x[last];
	}
	assert(last == -1);
	for (j = 0; j < n_snap; ++j) { free(snaps[j].x); free(snaps[j].ilo); free(snaps[j].icnt); }
	free(snaps); free(xbuf);
	ring_free(w); ring_free(sf);
	*n_seg_ = n_snap;
	return seg;
}

void orc_wfa_exact(const orc_opt_t *o, int32_t tl, const char *ts, int32_t ql, const char *qs, orc_rst_t *r)
{
	int32_t n_seg = 0, *seg = 0;
	if (o->step > 0) seg = orc_wfa_checkpoints(o, tl, ts, ql, qs, &n_seg); /* miniwfa.c:610-611 */
	core(o, tl, ts, ql, qs, n_seg, n_seg > 0 ? seg : 0, r);
	free(seg);
}

void orc_wfa_auto_exact_leg(const orc_opt_t *o0, int32_t tl, const char *ts, int32_t ql, const char *qs, orc_rst_t *r)
{
	orc_opt_t o = *o0; /* miniwfa.c:900-902; r->s < 0 means the reference would fall back to chaining */
	o.step = 0, o.max_iter = 100000000;
	orc_wfa_exact(&o, tl, ts, ql, qs, r);
}

int32_t orc_cigar2score(const orc_opt_t *o, int32_t n_cigar, const uint32_t *cigar, int32_t *tl, int32_t *ql)
{
	int32_t i, s = 0, t = 0, q = 0; /* mwf-dbg.c:6-22 */
	for (i = 0; i < n_cigar; ++i) {
		int32_t op = cigar[i] & 0xf, len = (int32_t)(cigar[i] >> 4);
		if (op == 1 || op == 2) {
			int32_t g1 = o->o1 + len * o->e1, g2 = o->o2 + len * o->e2;
			s += g1 < g2 ? g1 : g2;
			if (op == 1) q += len; else t += len;
		} else if (op == 8) s += len * o->x, t += len, q += len;
		else if (op == 0 || op == 7) t += len, q += len;
	}
	if (tl) *tl = t;
	if (ql) *ql = q;
	return s;
}
