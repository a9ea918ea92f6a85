/* Dev tool (host only): times longest_increasing() of csrc/mwf_chain.c -- and an alternative longest_increasing2() from lis2.c when
 * built with -DLIS2 -- on the k-mer matches of a synthetic 5 Mb / 3 % pair (t.bin / q.bin written by run.sh).  Minimum of 7 runs. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include CHAIN_C
void mwf_wfa_exact_batch(void *km, const mwf_opt_t *opt, int32_t n, const int32_t *tl, const char *const *ts, const int32_t *ql, const char *const *qs, mwf_rst_t *r) { abort(); }
int64_t mwf_b200_kmer_hits(int32_t tl, const char *ts, int32_t ql, const char *qs, int32_t k, int32_t max_occ, uint64_t **hits) { abort(); }
void mwf_b200_kmer_free(uint64_t *hits) { abort(); }
void *mwf_b200_host_scratch(size_t bytes) { static void *p; static size_t cap; if (bytes > cap) { free(p); p = malloc(bytes); memset(p, 0, bytes); cap = bytes; } return p; }
void mwf_b200_host_scratch_free(void *p) { }
void mwf_b200_kmer_shared(int32_t l1, const char *s1, int32_t l2, const char *s2, int32_t k, int64_t *n1, int64_t *n2, int64_t *shared) { abort(); }
static char *slurp(const char *fn, int32_t *len) { FILE *f = fopen(fn, "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); char *b = malloc(n + 64); if (fread(b, 1, n, f) != (size_t)n) abort(); fclose(f); *len = (int32_t)n; return b; }
#ifdef LIS2
#include "lis2.c"
#endif
int main(int argc, char **argv)
{
	int32_t n_hit, rep, n_out = 0, tl, ql, i;
	uint64_t *hit, *o;
	FILE *f = fopen("hits.bin", "rb");
	double best = 1e30, best2 = 1e30;
	unsigned long long h = 0, h2 = 0;
	if (f) { fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); hit = malloc(n); if (fread(hit, 1, n, f) != (size_t)n) abort(); fclose(f); n_hit = n / 8; }
	else {
		char *t = slurp("t.bin", &tl), *q = slurp("q.bin", &ql);
		hit = host_hits(0, tl, t, ql, q, 13, 2, &n_hit);
		f = fopen("hits.bin", "wb"); fwrite(hit, 8, n_hit, f); fclose(f);
	}
	for (rep = 0; rep < 7; ++rep) {
		double t0 = now_ms();
		o = longest_increasing(0, n_hit, hit, &n_out);
		t0 = now_ms() - t0;
		if (t0 < best) best = t0;
		for (h = 0, i = 0; i < n_out; ++i) h = h * 1000003u + o[i];
		free(o);
#ifdef LIS2
		t0 = now_ms();
		o = longest_increasing2(0, n_hit, hit, &n_out);
		t0 = now_ms() - t0;
		if (t0 < best2) best2 = t0;
		for (h2 = 0, i = 0; i < n_out; ++i) h2 = h2 * 1000003u + o[i];
		free(o);
#endif
	}
	printf("%s: %d matches -> chain of %d: %.2f ms (hash %llx)", argv[1], n_hit, n_out, best, h);
#ifdef LIS2
	printf("; alternative %.2f ms (hash %llx%s)", best2, h2, h2 == h ? ", same" : ", DIFFERENT");
#endif
	printf("\n");
	return 0;
}
