/* the block of 64 piles an element lands in: bl = the largest b in [0, nb] with samp[b] < x (b = 0 counts as below), found through
 * samp2[c] = samp[64 c] -- a few hundred entries that stay in L1 -- and then 64 entries of samp */
static inline int32_t lis_block(const uint64_t *samp2, const uint64_t *samp, int32_t nb, uint64_t x)
{
	const int32_t c = lis_last_below(samp2, 0, nb / LIS_BLK + 1, x);
	const int32_t lo = c * LIS_BLK, hi = lo + LIS_BLK < nb + 1 ? lo + LIS_BLK : nb + 1;
	return lis_last_below(samp, lo, hi, x);
}

static uint64_t *longest_increasing2(void *km, int32_t n, const uint64_t *v, int32_t *n_out)
{
	int32_t *tail, *prev, i, len = 0, at, guess[LIS_AHEAD];
	uint64_t *tailv, *samp, *samp2, *out;
	const size_t n_samp = (size_t)n / LIS_BLK + 2, n_samp2 = n_samp / LIS_BLK + 2;
	/* Large inputs take the ~16 bytes per match of scratch from the library's cache of pinned host buffers: memory that is
	 * already resident, where a fresh kmalloc of tens of MB is paid for in page faults on every call. */
	void *ws = n >= LIS_WS_MIN ? mwf_b200_host_scratch(sizeof(uint64_t) * ((size_t)n + 1 + n_samp + n_samp2) + sizeof(int32_t) * (2 * (size_t)n + 2)) : 0;
	*n_out = 0;
	if (n <= 0) return 0;
	if (ws) {
		tailv = (uint64_t*)ws, samp = tailv + n + 1, samp2 = samp + n_samp;
		tail = (int32_t*)(samp2 + n_samp2), prev = tail + n + 1;
	} else {
		tail = (int32_t*)kmalloc(km, sizeof(int32_t) * ((size_t)n + 1));
		tailv = (uint64_t*)kmalloc(km, sizeof(uint64_t) * ((size_t)n + 1));
		samp = (uint64_t*)kmalloc(km, sizeof(uint64_t) * n_samp);
		samp2 = (uint64_t*)kmalloc(km, sizeof(uint64_t) * n_samp2);
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
				const int32_t bl = lis_block(samp2, samp, len / LIS_BLK, y);
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
				else bl = lis_block(samp2, samp, nb, x);
				lo = bl * LIS_BLK;
				if (lo + LIS_BLK < hi) hi = lo + LIS_BLK; /* = LIS_BLK (bl + 1), and samp[bl + 1] >= x */
			}
			if (lo < 0) lo = 0;
			lo = lis_last_below(tailv, lo, hi, x);
		}
		prev[i] = lo > 0 ? tail[lo] : -1;
		tail[lo + 1] = i, tailv[lo + 1] = x;
		if ((lo + 1) % LIS_BLK == 0) {
			const int32_t sb = (lo + 1) / LIS_BLK;
			samp[sb] = x;
			if (sb % LIS_BLK == 0) samp2[sb / LIS_BLK] = x;
		}
		if (lo + 1 > len) len = lo + 1;
	}
	out = (uint64_t*)kmalloc(km, sizeof(uint64_t) * (size_t)len);
	for (i = len - 1, at = tail[len]; i >= 0; --i) out[i] = v[at], at = prev[at];
	if (ws) mwf_b200_host_scratch_free(ws);
	else kfree(km, prev), kfree(km, samp2), kfree(km, samp), kfree(km, tailv), kfree(km, tail);
	*n_out = len;
	return out;
}

