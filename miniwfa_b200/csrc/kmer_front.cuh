/*
 * kmer_front.cuh -- the k-mer front end of mwf_wfa_chain on the device.
 *
 * Reference (miniwfa.c @ 66770a3): mg_fc_kmer (:718-730) lists every k-mer of both sequences as
 * (k-mer << 1 | which) << 32 | position-of-last-base, radix_sort_mwf64 sorts the list, a scan over groups of equal k-mers
 * emits target x query position pairs for k-mers with at most max_occ copies on either side (:748-764), a second sort
 * orders them by (target, query) position (:768); mwf_ksim (:786-812) runs the same list + sort and sums min(copies in
 * the target, copies in the query) over the groups.  On a 5 Mb pair that is 3 s of one host core -- all of
 * mwf_wfa_chain's time once its gap fills run as one GPU batch.  Here:
 *
 *   kmer_list_kernel   one thread per 4 positions, base codes staged in shared memory; an invalid position (fewer than
 *                      k valid bases behind it) writes an all-ones key, so the list keeps one slot per base and needs
 *                      no compaction -- the sort moves those keys to the end.
 *   sort 1             cub::DeviceRadixSort (LSD, stable) on the (k-mer, which) bits only: positions are already
 *                      ascending inside each sequence, so stability gives the reference's full 64-bit order.
 *   kmer_group_kernel  one thread per sorted key; a thread on the first key of a group finds the end of the target copies
 *                      and of the group by galloping search and, depending on the mode, counts the group's hits, writes
 *                      them (slots from a block scan + one atomic per block) or adds min(m1, m2) to the similarity sum.
 *   sort 2             the hits, on all significant bits of (target << 32 | query).
 *   kmer_swap_kernel   (target, query) -> (query, target), the order mg_lis_64 compares in (:769-770).
 *
 * The sorts are library code (CUB, part of the CUDA toolkit); the hot path of north_star does not include them.  The
 * longest increasing subsequence stays on the host (mwf_chain.c): it is a sequential scan.
 *
 * Included by wfa_engine.cu (shares its workspace cache, CUDA_OK and die()).
 */
#include <cub/device/device_radix_sort.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/block/block_reduce.cuh>

#define KMER_NONE 0xFFFFFFFFFFFFFFFFULL
enum { KMER_TILE = 1024, KMER_THREADS = 256, KMER_PER_THREAD = KMER_TILE / KMER_THREADS };
enum { KG_COUNT = 0, KG_FILL = 1, KG_SIM = 2 };

/* seq_nt4_table of the reference (:699-716): A/C/G/T/U in either case and the raw codes 0..3; anything else breaks a k-mer */
__device__ __forceinline__ int kmer_base_code(unsigned int ch)
{
	if (ch < 4) return (int)ch;
	switch (ch & 0xDFu) { /* clears the lower-case bit: only 'A' and 'a' give 'A', and so on */
	case 'A': return 0;
	case 'C': return 1;
	case 'G': return 2;
	case 'T': case 'U': return 3;
	default: return 4;
	}
}

/* keys[i] for every position i of one sequence; counts[which] += number of valid k-mers */
__global__ void __launch_bounds__(KMER_THREADS) kmer_list_kernel(const unsigned char *__restrict__ seq, long long len, int which, int k,
                                                                 unsigned long long *__restrict__ keys, unsigned long long *counts)
{
	__shared__ unsigned char code[KMER_TILE + 16];
	const long long tile0 = (long long)blockIdx.x * KMER_TILE;
	const int back = k - 1; /* <= 14 */
	for (int j = threadIdx.x; j < KMER_TILE + back; j += KMER_THREADS) {
		const long long p = tile0 - back + j;
		code[j] = (p >= 0 && p < len) ? (unsigned char)kmer_base_code(seq[p]) : (unsigned char)4;
	}
	__syncthreads();
	const unsigned long long mask = (1ULL << 2 * k) - 1;
	const int j0 = threadIdx.x * KMER_PER_THREAD; /* first position of this thread, relative to the tile */
	unsigned long long word = 0;
	int run = 0, n_valid = 0;
	for (int j = 0; j < back + KMER_PER_THREAD; ++j) {
		const int c = code[j0 + j];
		if (c < 4) word = (word << 2 | (unsigned long long)c) & mask, ++run;
		else word = 0, run = 0;
		if (j >= back) {
			const long long p = tile0 + j0 + (j - back);
			if (p < len) {
				const bool ok = run >= k;
				keys[p] = ok ? ((word << 1 | (unsigned long long)which) << 32 | (unsigned long long)p) : KMER_NONE;
				n_valid += ok;
			}
		}
	}
	typedef cub::BlockReduce<int, KMER_THREADS> Reduce;
	__shared__ typename Reduce::TempStorage tmp;
	const int total = Reduce(tmp).Sum(n_valid);
	if (threadIdx.x == 0 && total) atomicAdd(&counts[which], (unsigned long long)total);
}

/* first index >= from whose key >> shift differs from val (keys sorted; n if there is none) */
__device__ __forceinline__ long long kmer_run_end(const unsigned long long *__restrict__ keys, long long from, long long n, int shift, unsigned long long val)
{
	if (from >= n || keys[from] >> shift != val) return from;
	long long lo = from, step = 1; /* keys[lo] matches */
	while (lo + step < n && keys[lo + step] >> shift == val) lo += step, step <<= 1;
	long long hi = lo + step < n ? lo + step : n; /* keys[hi] does not match, or hi == n */
	while (hi - lo > 1) {
		const long long mid = lo + ((hi - lo) >> 1);
		if (keys[mid] >> shift == val) lo = mid; else hi = mid;
	}
	return hi;
}

/* out[0]: number of hits (KG_COUNT) / next free slot (KG_FILL) / sum of min(m1, m2) (KG_SIM) */
template<int MODE>
__global__ void __launch_bounds__(KMER_THREADS) kmer_group_kernel(const unsigned long long *__restrict__ keys, long long n, long long max_occ,
                                                                  unsigned long long *__restrict__ hits, unsigned long long *out)
{
	const long long i = (long long)blockIdx.x * KMER_THREADS + threadIdx.x;
	long long m1 = 0, m2 = 0, first = 0, split = 0;
	unsigned long long cnt = 0;
	if (i < n) {
		const unsigned long long key = keys[i], grp = key >> 33;
		if (i == 0 || keys[i - 1] >> 33 != grp) { /* first key of its group: the target copies come first (which = 0) */
			first = i;
			split = kmer_run_end(keys, i, n, 32, grp << 1);
			m1 = split - i;
			if (MODE == KG_SIM || (m1 > 0 && m1 <= max_occ)) {
				const long long end = kmer_run_end(keys, split, n, 33, grp);
				m2 = end - split;
			}
			if (MODE == KG_SIM) cnt = (unsigned long long)(m1 < m2 ? m1 : m2);
			else if (m1 > 0 && m2 > 0 && m1 <= max_occ && m2 <= max_occ) cnt = (unsigned long long)(m1 * m2);
		}
	}
	if (MODE == KG_FILL) {
		typedef cub::BlockScan<unsigned long long, KMER_THREADS> Scan;
		__shared__ typename Scan::TempStorage tmp;
		__shared__ unsigned long long base;
		unsigned long long off, total;
		Scan(tmp).ExclusiveSum(cnt, off, total);
		if (threadIdx.x == 0) base = total ? atomicAdd(out, total) : 0;
		__syncthreads();
		if (cnt) {
			unsigned long long *dst = hits + base + off;
			for (long long s = 0; s < m1; ++s) {
				const unsigned long long tpos = keys[first + s] << 32;
				for (long long t = 0; t < m2; ++t) *dst++ = tpos | (keys[split + t] & 0xFFFFFFFFULL);
			}
		}
	} else {
		typedef cub::BlockReduce<unsigned long long, KMER_THREADS> Reduce;
		__shared__ typename Reduce::TempStorage tmp;
		const unsigned long long total = Reduce(tmp).Sum(cnt);
		if (threadIdx.x == 0 && total) atomicAdd(out, total);
	}
}

__global__ void kmer_swap_kernel(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out, long long n)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { const unsigned long long v = in[i]; out[i] = v >> 32 | v << 32; }
}

/* ---- host side ---- */

struct KmerList {
	int dev;
	cudaStream_t st;
	unsigned char *d_seq;
	unsigned long long *d_keys[2], *d_cnt, *h_cnt; /* cnt: [0] target k-mers, [1] query k-mers, [2] group result */
	unsigned long long *sorted;
	long long n_all, n_valid;
	int launches;
};

static int bits_for(long long v) { int b = 0; while (b < 63 && (1LL << b) <= v) ++b; return b; } /* v < 2^b */

static void kmer_sort(KmerList *L, unsigned long long *a, unsigned long long *b, long long n, int begin_bit, int end_bit, unsigned long long **result)
{
	cub::DoubleBuffer<unsigned long long> buf(a, b);
	size_t tmp_bytes = 0;
	CUDA_OK(cub::DeviceRadixSort::SortKeys((void*)0, tmp_bytes, buf, n, begin_bit, end_bit, L->st));
	void *tmp = 0;
	ws_alloc(&tmp, tmp_bytes, false, L->dev);
	CUDA_OK(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, buf, n, begin_bit, end_bit, L->st));
	CUDA_OK(cudaStreamSynchronize(L->st)); /* tmp goes back to the cache only after the sort is done with it */
	ws_free(tmp);
	L->launches += 4 + (end_bit - begin_bit + 7) / 8; /* histogram + scan + one pass per digit, roughly */
	*result = buf.Current();
}

/* upload both sequences, list their k-mers and sort them; L->sorted[0..n_valid) are the valid keys */
static void kmer_list_sorted(KmerList *L, long long l1, const char *s1, long long l2, const char *s2, int k)
{
	if (mwf_b200_device_count() <= 0) die("no CUDA device: the k-mer front end of mwf_wfa_chain has no CPU fallback");
	if (k < 2 || k > 15) die("k-mer length out of range (2..15)");
	memset(L, 0, sizeof(*L));
	L->dev = mwf_b200_get_device();
	CUDA_OK(cudaSetDevice(L->dev));
	CUDA_OK(cudaStreamCreateWithFlags(&L->st, cudaStreamNonBlocking));
	L->n_all = l1 + l2;
	ws_dev(&L->d_seq, (size_t)L->n_all + 16, L->dev);
	ws_dev(&L->d_keys[0], sizeof(unsigned long long) * (size_t)L->n_all, L->dev);
	ws_dev(&L->d_keys[1], sizeof(unsigned long long) * (size_t)L->n_all, L->dev);
	ws_dev(&L->d_cnt, 64, L->dev);
	ws_host(&L->h_cnt, 64);
	CUDA_OK(cudaMemsetAsync(L->d_cnt, 0, 64, L->st));
	CUDA_OK(cudaMemcpyAsync(L->d_seq, s1, (size_t)l1, cudaMemcpyHostToDevice, L->st));
	CUDA_OK(cudaMemcpyAsync(L->d_seq + l1, s2, (size_t)l2, cudaMemcpyHostToDevice, L->st));
	kmer_list_kernel<<<(unsigned)((l1 + KMER_TILE - 1) / KMER_TILE), KMER_THREADS, 0, L->st>>>(L->d_seq, l1, 0, k, L->d_keys[0], L->d_cnt);
	kmer_list_kernel<<<(unsigned)((l2 + KMER_TILE - 1) / KMER_TILE), KMER_THREADS, 0, L->st>>>(L->d_seq + l1, l2, 1, k, L->d_keys[0] + l1, L->d_cnt);
	CUDA_OK(cudaGetLastError());
	CUDA_OK(cudaMemcpyAsync(L->h_cnt, L->d_cnt, 16, cudaMemcpyDeviceToHost, L->st));
	L->launches += 2;
	/* bit 32 + 2k + 1 is the one that tells a real key (0) from KMER_NONE (1) */
	kmer_sort(L, L->d_keys[0], L->d_keys[1], L->n_all, 32, 32 + 2 * k + 2, &L->sorted);
	L->n_valid = (long long)(L->h_cnt[0] + L->h_cnt[1]);
}

static void kmer_list_release(KmerList *L)
{
	ws_free(L->d_seq); ws_free(L->d_keys[0]); ws_free(L->d_keys[1]); ws_free(L->d_cnt); ws_free(L->h_cnt);
	CUDA_OK(cudaStreamDestroy(L->st));
}

static std::atomic<long long> g_kmer_launches(0);

extern "C" int64_t mwf_b200_kmer_hits(int32_t tl, const char *ts, int32_t ql, const char *qs, int32_t k, int32_t max_occ, uint64_t **hits)
{
	KmerList L;
	*hits = 0;
	if (tl < k || ql < k) return 0;
	kmer_list_sorted(&L, tl, ts, ql, qs, k);
	long long n_hit = 0;
	if (L.n_valid > 0) {
		const unsigned blocks = (unsigned)((L.n_valid + KMER_THREADS - 1) / KMER_THREADS);
		kmer_group_kernel<KG_COUNT><<<blocks, KMER_THREADS, 0, L.st>>>(L.sorted, L.n_valid, max_occ, 0, L.d_cnt + 2);
		CUDA_OK(cudaGetLastError());
		CUDA_OK(cudaMemcpyAsync(L.h_cnt + 2, L.d_cnt + 2, 8, cudaMemcpyDeviceToHost, L.st));
		CUDA_OK(cudaStreamSynchronize(L.st));
		n_hit = (long long)L.h_cnt[2];
		L.launches += 1;
		if (n_hit > 0) {
			unsigned long long *d_hit[2], *h_hit = 0, *d_sorted;
			ws_dev(&d_hit[0], sizeof(unsigned long long) * (size_t)n_hit, L.dev);
			ws_dev(&d_hit[1], sizeof(unsigned long long) * (size_t)n_hit, L.dev);
			ws_host(&h_hit, sizeof(unsigned long long) * (size_t)n_hit);
			CUDA_OK(cudaMemsetAsync(L.d_cnt + 3, 0, 8, L.st));
			kmer_group_kernel<KG_FILL><<<blocks, KMER_THREADS, 0, L.st>>>(L.sorted, L.n_valid, max_occ, d_hit[0], L.d_cnt + 3);
			CUDA_OK(cudaGetLastError());
			kmer_sort(&L, d_hit[0], d_hit[1], n_hit, 0, 32 + bits_for(tl), &d_sorted);
			unsigned long long *d_other = d_sorted == d_hit[0] ? d_hit[1] : d_hit[0];
			kmer_swap_kernel<<<(unsigned)((n_hit + 255) / 256), 256, 0, L.st>>>(d_sorted, d_other, n_hit);
			CUDA_OK(cudaGetLastError());
			CUDA_OK(cudaMemcpyAsync(h_hit, d_other, sizeof(unsigned long long) * (size_t)n_hit, cudaMemcpyDeviceToHost, L.st));
			CUDA_OK(cudaStreamSynchronize(L.st));
			L.launches += 2;
			ws_free(d_hit[0]); ws_free(d_hit[1]);
			*hits = (uint64_t*)h_hit;
		}
	}
	g_kmer_launches += L.launches;
	kmer_list_release(&L);
	return n_hit;
}

extern "C" void mwf_b200_kmer_free(uint64_t *hits) { ws_free(hits); }

extern "C" void *mwf_b200_host_scratch(size_t bytes)
{
	void *p = 0;
	if (mwf_b200_device_count() <= 0) die("no CUDA device: this library has no CPU fallback");
	CUDA_OK(cudaSetDevice(mwf_b200_get_device()));
	ws_alloc(&p, bytes, true, 0);
	return p;
}

extern "C" void mwf_b200_host_scratch_free(void *p) { ws_free(p); }

extern "C" void mwf_b200_kmer_shared(int32_t l1, const char *s1, int32_t l2, const char *s2, int32_t k, int64_t *n1, int64_t *n2, int64_t *shared)
{
	KmerList L;
	*n1 = *n2 = *shared = 0;
	if (l1 < k || l2 < k) return;
	kmer_list_sorted(&L, l1, s1, l2, s2, k);
	*n1 = (int64_t)L.h_cnt[0], *n2 = (int64_t)L.h_cnt[1];
	if (L.n_valid > 0) {
		const unsigned blocks = (unsigned)((L.n_valid + KMER_THREADS - 1) / KMER_THREADS);
		kmer_group_kernel<KG_SIM><<<blocks, KMER_THREADS, 0, L.st>>>(L.sorted, L.n_valid, 0, 0, L.d_cnt + 2);
		CUDA_OK(cudaGetLastError());
		CUDA_OK(cudaMemcpyAsync(L.h_cnt + 2, L.d_cnt + 2, 8, cudaMemcpyDeviceToHost, L.st));
		CUDA_OK(cudaStreamSynchronize(L.st));
		*shared = (int64_t)L.h_cnt[2];
		L.launches += 1;
	}
	g_kmer_launches += L.launches;
	kmer_list_release(&L);
}

extern "C" int64_t mwf_b200_kmer_launches(void) { return g_kmer_launches; }
