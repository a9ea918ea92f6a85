/*
 * main.c -- test-mwf: the command-line harness of the reference (main.c:19-92) over this library.
 *
 * Same options (-c -p INT -u -t -l INT -f INT -a -e -K -d, reference main.c:29-39), same pairing of the i-th
 * records of the two files (main.c:67), same PAF-like output line and CIGAR (main.c:73-80), and mwf_assert_cigar
 * on every CIGAR (main.c:72).  Different on purpose: in exact mode the pairs are submitted in batches
 * (mwf_wfa_exact_batch), because one submission of many pairs keeps every SM busy, and a reader thread parses the next
 * chunk of both files (256 Mbases or 65 536 pairs; MWF_CLI_CHUNK_BASES / MWF_CLI_CHUNK_PAIRS) while the GPU works on the
 * current one; the "T" lines on stderr therefore report wall-clock seconds of a batch divided by its number of pairs.
 * Input: FASTA or 4-line FASTQ, plain or gzip (zlib), one sequence per record, multi-line FASTA accepted.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <time.h>
#include <zlib.h>
#include <pthread.h>
#include "miniwfa.h"
#include "mwf_b200.h"

typedef struct { char *name, *seq; int32_t len; } rec_t;

static void *xrealloc(void *p, size_t n)
{
	void *q = realloc(p, n);
	if (q == 0 && n) { fprintf(stderr, "ERROR: out of memory\n"); exit(1); }
	return q;
}

static int read_line(gzFile fp, char **buf, size_t *cap) /* one line without its end-of-line bytes; -1 at end of file */
{
	size_t len = 0;
	for (;;) {
		if (*cap - len < 2) *cap = *cap ? *cap * 2 : 65536, *buf = (char*)xrealloc(*buf, *cap);
		if (gzgets(fp, *buf + len, (int)(*cap - len < 0x40000000 ? *cap - len : 0x40000000)) == 0) {
			if (len == 0) return -1;
			break;
		}
		len += strlen(*buf + len);
		if (len && (*buf)[len - 1] == '\n') break;
	}
	while (len && ((*buf)[len - 1] == '\n' || (*buf)[len - 1] == '\r')) (*buf)[--len] = 0;
	return (int)len;
}

/* incremental FASTA / FASTQ reader: one record per call */
typedef struct {
	gzFile fp;
	char *line;
	size_t cap;
	int n, pending, fastq; /* pending: `line` holds a header line that has not been consumed yet */
} reader_t;

static void reader_open(reader_t *r, const char *path)
{
	memset(r, 0, sizeof(*r));
	r->fp = strcmp(path, "-") ? gzopen(path, "r") : gzdopen(0, "r");
	if (r->fp == 0) { fprintf(stderr, "ERROR: cannot open %s\n", path); exit(1); }
	gzbuffer(r->fp, 1 << 20);
}

static void reader_close(reader_t *r) { free(r->line); gzclose(r->fp); }

static int is_header(const reader_t *r, int n, int have)
{
	return n > 0 && (r->line[0] == '>' || (r->line[0] == '@' && (!have || r->fastq)));
}

static int next_record(reader_t *r, rec_t *out) /* 1 = a record, 0 = end of file */
{
	char *seq = 0, *sp;
	size_t scap = 0, slen = 0, qual_left = 0;
	int in_qual = 0;
	if (!r->pending) { /* skip to the first header */
		while ((r->n = read_line(r->fp, &r->line, &r->cap)) >= 0 && !is_header(r, r->n, 0)) {}
		if (r->n < 0) return 0;
	}
	r->pending = 0;
	r->fastq = r->line[0] == '@';
	for (sp = r->line + 1; *sp && *sp != ' ' && *sp != '\t'; ++sp) {}
	*sp = 0;
	out->name = strdup(r->line + 1);
	while ((r->n = read_line(r->fp, &r->line, &r->cap)) >= 0) {
		const int n = r->n;
		if (in_qual) { /* FASTQ quality: as many characters as bases */
			qual_left = qual_left > (size_t)n ? qual_left - (size_t)n : 0;
			if (qual_left == 0) in_qual = 0;
			continue;
		}
		if (is_header(r, n, 1)) { r->pending = 1; break; }
		if (r->fastq && n > 0 && r->line[0] == '+') in_qual = slen > 0, qual_left = slen;
		else {
			if (slen + n + 1 > scap) scap = (slen + n + 1) * 2, seq = (char*)xrealloc(seq, scap);
			memcpy(seq + slen, r->line, n), slen += n;
		}
	}
	out->seq = (char*)xrealloc(seq, slen + 1), out->seq[slen] = 0, out->len = (int32_t)slen;
	return 1;
}

/*
 * Ingest pipeline: a reader thread parses the i-th records of both files into chunks of pairs (bounded by
 * bases and by pairs) while the main thread aligns the previous chunk on the GPU and prints it; at most two chunks wait in the
 * queue, so host memory stays bounded on files of any size.  Inside a chunk the library stages the sequences through pinned
 * memory (mwf_b200_batch_upload).
 */
typedef struct chunk_s { rec_t *t, *q; int n; struct chunk_s *next; } chunk_t;

typedef struct {
	const char *path[2];
	long long max_bases;
	int max_pairs;
	pthread_mutex_t mu;
	pthread_cond_t cv;
	chunk_t *head, *tail;
	int n_queued, done;
} ingest_t;

static void *ingest_main(void *arg)
{
	ingest_t *g = (ingest_t*)arg;
	reader_t r1, r2;
	int eof = 0;
	reader_open(&r1, g->path[0]);
	reader_open(&r2, g->path[1]);
	while (!eof) {
		chunk_t *c = (chunk_t*)calloc(1, sizeof(chunk_t));
		long long bases = 0;
		int cap = 0;
		while (c->n < g->max_pairs && bases < g->max_bases) {
			rec_t a, b;
			if (!next_record(&r1, &a)) { eof = 1; break; }
			if (!next_record(&r2, &b)) { free(a.name); free(a.seq); eof = 1; break; }
			if (c->n == cap) {
				cap = cap ? cap * 2 : 64;
				c->t = (rec_t*)xrealloc(c->t, sizeof(rec_t) * cap), c->q = (rec_t*)xrealloc(c->q, sizeof(rec_t) * cap);
			}
			c->t[c->n] = a, c->q[c->n] = b, ++c->n;
			bases += (long long)a.len + b.len;
		}
		pthread_mutex_lock(&g->mu);
		while (g->n_queued >= 2) pthread_cond_wait(&g->cv, &g->mu);
		if (c->n > 0) {
			if (g->tail) g->tail->next = c; else g->head = c;
			g->tail = c, ++g->n_queued;
		} else free(c);
		if (eof) g->done = 1;
		pthread_cond_broadcast(&g->cv);
		pthread_mutex_unlock(&g->mu);
	}
	reader_close(&r1);
	reader_close(&r2);
	return 0;
}

static chunk_t *ingest_pop(ingest_t *g) /* NULL after the last chunk */
{
	chunk_t *c;
	pthread_mutex_lock(&g->mu);
	while (g->head == 0 && !g->done) pthread_cond_wait(&g->cv, &g->mu);
	c = g->head;
	if (c) {
		g->head = c->next;
		if (g->head == 0) g->tail = 0;
		--g->n_queued;
		pthread_cond_broadcast(&g->cv);
	}
	pthread_mutex_unlock(&g->mu);
	return c;
}

static double wall(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void print_pair(const mwf_opt_t *opt, const rec_t *t, const rec_t *q, const mwf_rst_t *r)
{
	if (opt->flag & MWF_F_CIGAR) mwf_assert_cigar(opt, r->n_cigar, r->cigar, t->len, q->len, r->s);
	printf("%s\t%ld\t0\t%ld\t+\t%s\t%ld\t0\t%ld\t%d", t->name, (long)t->len, (long)t->len, q->name, (long)q->len, (long)q->len, r->s);
	if (opt->flag & MWF_F_CIGAR) {
		int32_t i;
		putchar('\t');
		for (i = 0; i < r->n_cigar; ++i) printf("%d%c", r->cigar[i] >> 4, "MIDNSHP=XBid"[r->cigar[i] & 0xf]);
	}
	putchar('\n');
	fflush(stdout);
}

int main(int argc, char *argv[])
{
	mwf_opt_t opt;
	ingest_t ing;
	chunk_t *ck;
	pthread_t tid;
	mwf_rst_t *rst;
	int c, mode = 0, n, i;
	double t0, dt;

	mwf_opt_init(&opt);
	while ((c = getopt(argc, argv, "cKdep:autl:f:")) >= 0) {
		switch (c) {
		case 'K': opt.flag |= MWF_F_NO_KALLOC; break;
		case 'c': opt.flag |= MWF_F_CIGAR; break;
		case 'd': opt.flag |= MWF_F_DEBUG; break;
		case 'p': opt.flag |= MWF_F_CIGAR, opt.step = atoi(optarg); break;
		case 'a': opt.o2 = opt.o1, opt.e2 = opt.e1; break;
		case 'e': opt.x = 1, opt.o1 = opt.o2 = 0, opt.e1 = opt.e2 = 1; break;
		case 'l': opt.min_len = atoi(optarg); break;
		case 'f': opt.max_occ = atoi(optarg); break;
		case 'u': mode = 1; break; /* chaining heuristic */
		case 't': mode = 2; break; /* exact within a 10^8-cell budget, else chaining */
		default: fprintf(stderr, "ERROR: unknown option\n"); return 1;
		}
	}
	if (argc - optind < 2) {
		fprintf(stderr, "Usage: test-mwf [options] <in1.fa> <in2.fa>\n");
		fprintf(stderr, "Options:\n");
		fprintf(stderr, "  -c       generate CIGAR\n");
		fprintf(stderr, "  -p INT   step size (force -c; 0 to disable) [%d]\n", opt.step);
		fprintf(stderr, "  -u       apply the chaining heuristic\n");
		fprintf(stderr, "  -t       automatically choose between the exact and the chaining mode\n");
		fprintf(stderr, "  -l INT   min gapless length for chain filtering [%d]\n", opt.min_len);
		fprintf(stderr, "  -f INT   max k-mer occurrence [%d]\n", opt.max_occ);
		fprintf(stderr, "  -a       mimic affine gap\n");
		fprintf(stderr, "  -e       mimic edit distance\n");
		fprintf(stderr, "  -K       disable the kalloc allocator\n");
		return 1;
	}
	memset(&ing, 0, sizeof(ing));
	ing.path[0] = argv[optind], ing.path[1] = argv[optind + 1];
	ing.max_bases = getenv("MWF_CLI_CHUNK_BASES") ? atoll(getenv("MWF_CLI_CHUNK_BASES")) : 256LL << 20;
	ing.max_pairs = getenv("MWF_CLI_CHUNK_PAIRS") ? atoi(getenv("MWF_CLI_CHUNK_PAIRS")) : 65536;
	if (ing.max_bases < 1) ing.max_bases = 1;
	if (ing.max_pairs < 1) ing.max_pairs = 1;
	pthread_mutex_init(&ing.mu, 0);
	pthread_cond_init(&ing.cv, 0);
	if (pthread_create(&tid, 0, ingest_main, &ing) != 0) { fprintf(stderr, "ERROR: cannot start the reader thread\n"); return 1; }
	while ((ck = ingest_pop(&ing)) != 0) {
		n = ck->n;
		rst = (mwf_rst_t*)calloc(n, sizeof(mwf_rst_t));
		t0 = wall();
		if (mode == 0) { /* every pair of the chunk in one submission */
			int32_t *tl = (int32_t*)malloc(sizeof(int32_t) * n), *ql = (int32_t*)malloc(sizeof(int32_t) * n);
			const char **ts = (const char**)malloc(sizeof(char*) * n), **qs = (const char**)malloc(sizeof(char*) * n);
			for (i = 0; i < n; ++i) tl[i] = ck->t[i].len, ts[i] = ck->t[i].seq, ql[i] = ck->q[i].len, qs[i] = ck->q[i].seq;
			mwf_wfa_exact_batch(0, &opt, n, tl, ts, ql, qs, rst);
			free(tl); free(ql); free((void*)ts); free((void*)qs);
		} else {
			for (i = 0; i < n; ++i) {
				if (mode == 1) mwf_wfa_chain(0, &opt, ck->t[i].len, ck->t[i].seq, ck->q[i].len, ck->q[i].seq, &rst[i]);
				else mwf_wfa_auto(0, &opt, ck->t[i].len, ck->t[i].seq, ck->q[i].len, ck->q[i].seq, &rst[i]);
			}
		}
		dt = wall() - t0;
		for (i = 0; i < n; ++i) {
			print_pair(&opt, &ck->t[i], &ck->q[i], &rst[i]);
			free(rst[i].cigar);
			fprintf(stderr, "T\t%s\t%s\t%.3f\n", ck->t[i].name, ck->q[i].name, dt / n);
			free(ck->t[i].name); free(ck->t[i].seq); free(ck->q[i].name); free(ck->q[i].seq);
		}
		free(ck->t); free(ck->q); free(ck); free(rst);
	}
	pthread_join(tid, 0);
	return 0;
}
