/*
 * main.c -- test-mwf: the command-line harness of the reference (main.c:19-92) over this library.
 *
 * Same options (-c -p INT -u -t -l INT -f INT -a -e -K -d, reference main.c:29-39), same pairing of the i-th
 * records of the two files (main.c:67), same PAF-like output line and CIGAR (main.c:73-80), and mwf_assert_cigar
 * on every CIGAR (main.c:72).  Different on purpose: in exact mode all pairs of the two files are read first and
 * submitted as ONE batch (mwf_wfa_exact_batch), because one submission keeps every SM busy; the "T" lines on
 * stderr therefore report wall-clock seconds of the whole batch divided by the number of pairs.
 * Input: FASTA or 4-line FASTQ, plain or gzip (zlib), one sequence per record, multi-line FASTA accepted.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <time.h>
#include <zlib.h>
#include "miniwfa.h"
#include "mwf_b200.h"

typedef struct { char *name, *seq; int32_t len; } rec_t;
typedef struct { rec_t *a; int n, cap; } recs_t;

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

static void read_records(const char *path, recs_t *out)
{
	gzFile fp = strcmp(path, "-") ? gzopen(path, "r") : gzdopen(0, "r");
	char *line = 0, *seq = 0;
	size_t cap = 0, scap = 0, slen = 0;
	int n, fastq = 0, have = 0, in_qual = 0;
	size_t qual_left = 0;
	if (fp == 0) { fprintf(stderr, "ERROR: cannot open %s\n", path); exit(1); }
	out->a = 0, out->n = out->cap = 0;
	while ((n = read_line(fp, &line, &cap)) >= 0) {
		if (in_qual) { /* FASTQ quality: as many characters as bases */
			qual_left = qual_left > (size_t)n ? qual_left - (size_t)n : 0;
			if (qual_left == 0) in_qual = 0;
			continue;
		}
		if (n > 0 && (line[0] == '>' || (line[0] == '@' && (!have || fastq)))) {
			char *sp;
			if (have) { out->a[out->n - 1].seq = (char*)xrealloc(seq, slen + 1), out->a[out->n - 1].seq[slen] = 0, out->a[out->n - 1].len = (int32_t)slen; }
			fastq = line[0] == '@';
			if (out->n == out->cap) out->cap = out->cap ? out->cap * 2 : 16, out->a = (rec_t*)xrealloc(out->a, sizeof(rec_t) * out->cap);
			for (sp = line + 1; *sp && *sp != ' ' && *sp != '\t'; ++sp) {}
			*sp = 0;
			out->a[out->n].name = strdup(line + 1), out->a[out->n].seq = 0, out->a[out->n].len = 0;
			++out->n, have = 1;
			seq = 0, scap = slen = 0;
		} else if (have && fastq && n > 0 && line[0] == '+') {
			in_qual = slen > 0, qual_left = slen;
		} else if (have) {
			if (slen + n + 1 > scap) scap = (slen + n + 1) * 2, seq = (char*)xrealloc(seq, scap);
			memcpy(seq + slen, line, n), slen += n;
		}
	}
	if (have) { out->a[out->n - 1].seq = (char*)xrealloc(seq, slen + 1), out->a[out->n - 1].seq[slen] = 0, out->a[out->n - 1].len = (int32_t)slen; }
	free(line);
	gzclose(fp);
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
	recs_t f1, f2;
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
	read_records(argv[optind], &f1);
	read_records(argv[optind + 1], &f2);
	n = f1.n < f2.n ? f1.n : f2.n;
	rst = (mwf_rst_t*)calloc(n > 0 ? n : 1, sizeof(mwf_rst_t));
	t0 = wall();
	if (mode == 0) { /* every pair of the two files in one submission */
		int32_t *tl = (int32_t*)malloc(sizeof(int32_t) * (n + 1)), *ql = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
		const char **ts = (const char**)malloc(sizeof(char*) * (n + 1)), **qs = (const char**)malloc(sizeof(char*) * (n + 1));
		for (i = 0; i < n; ++i) tl[i] = f1.a[i].len, ts[i] = f1.a[i].seq, ql[i] = f2.a[i].len, qs[i] = f2.a[i].seq;
		mwf_wfa_exact_batch(0, &opt, n, tl, ts, ql, qs, rst);
		free(tl); free(ql); free((void*)ts); free((void*)qs);
	} else {
		for (i = 0; i < n; ++i) {
			if (mode == 1) mwf_wfa_chain(0, &opt, f1.a[i].len, f1.a[i].seq, f2.a[i].len, f2.a[i].seq, &rst[i]);
			else mwf_wfa_auto(0, &opt, f1.a[i].len, f1.a[i].seq, f2.a[i].len, f2.a[i].seq, &rst[i]);
		}
	}
	dt = wall() - t0;
	for (i = 0; i < n; ++i) {
		print_pair(&opt, &f1.a[i], &f2.a[i], &rst[i]);
		free(rst[i].cigar);
		fprintf(stderr, "T\t%s\t%s\t%.3f\n", f1.a[i].name, f2.a[i].name, dt / n);
	}
	for (i = 0; i < f1.n; ++i) free(f1.a[i].name), free(f1.a[i].seq);
	for (i = 0; i < f2.n; ++i) free(f2.a[i].name), free(f2.a[i].seq);
	free(f1.a); free(f2.a); free(rst);
	return 0;
}
