/*
 * wfa_tile.cuh -- the temporally blocked ("tile") engine: wf_next + wf_extend for T scores at a
 * time with the wavefront ring resident in shared memory.  Included by wfa_engine.cu.
 *
 * Reference (/root/reference/miniwfa.c @ 66770a3): the per-score loop of mwf_wfa_core (:397-426):
 * extend loop + wf_extend1_padded (:400-411, :212-226), wf_next_basic/prep/score/tb (:243-327),
 * wf_stripe_shrink (:144-171), the stop tests (:421-425), the checkpoint collapse (:413-416).
 *
 * Why: one score step reads 7 and writes 5 int32 per diagonal; with the ring in HBM that is 48-64 B per
 * cell and one barrier per score.  Here a *time block* advances a pair by Tb <= T scores:
 *   - the diagonals the block can touch, [wflo-Tb-nring, wfhi+Tb+nring], are cut into tiles;
 *   - a CTA bulk-copies (cp.async.bulk, mbarrier) the tile plus a halo of HL >= Tb diagonals on each side
 *     -- all R = nring + 2(e1+1) + 2(e2+1) live ring rows (27 for the default penalties, not the
 *     reference's 85) -- from the pair's state buffer into shared memory, runs Tb fused next+extend
 *     steps with one __syncthreads each, and bulk-copies the useful columns to the other state buffer;
 *   - halo columns go stale by one diagonal per step and are never stored.
 * HBM traffic drops to ~ (W+U)/(U*Tb) * 4R bytes per cell (3-7 B) plus 1 B/cell of traceback.
 *
 * Exactness of the band (n_iter, max_iter stop): the first and the last tile of a block replay the
 * reference's lo/hi rule (:325-326, :417-418) step by step, force every cell outside [lo_t, hi_t] to
 * NEG_INF (what the reference's pads hold, :96-99) and log lo_t, hi_t; wfa_plan_kernel then replays the
 * block score by score (n_iter, stop tests, termination), trims the band at multiples of 256 from
 * per-diagonal "alive" words the tiles accumulate over the last nring scores, and cuts the next block.
 *
 * Kernels: wfa_tile_init_kernel (score 0), wfa_plan_kernel (one CTA per pair, between blocks),
 * wfa_tile_kernel<MODE, CPT> (persistent CTAs pulling (pair, tile) items; CPT = cells per thread: 4 when a launch has enough
 * tiles to fill the GPU -- with a split-phase step barrier in interior tiles -- and 1 on the same 512-wide tile otherwise),
 * wfa_pack_kernel (two- / four-bit copies of the sequences for the match-run probes), wfa_tile_traceback_kernel;
 * low-memory mode: wfa_tile_checkpoint_kernel; traceback beyond s^2 bytes of HBM: snapshots taken by the tiles, then
 * wfa_tile_segstart_kernel, wfa_tile_trace_seg_kernel / wfa_tile_ckpt_seg_kernel per segment.
 */
#ifndef WFA_TILE_CUH
#define WFA_TILE_CUH

#define TILE_TMAX 64
#define TILE_NRING_MAX 64
#define TILE_EDEPTH_MAX 8

enum { TS_RUN = 0, TS_DONE = 1, TS_STOPPED = 2, TS_ARENA = 3, TS_SHRINK = 4, TS_SEGEND = 5, TS_IDLE = 6 };

struct TileCtl { /* per pair, lives in HBM for the whole run */
	int status, s, wflo, wfhi, cur, last, sid, copied;
	int tiles_done;        /* persistent scheduling: tiles of the block in flight that have finished */
	int shrink_s;          /* score of the last trim / band collapse that took diagonals away (slices older than that still hold them) */
	long long n_iter;
	/* the block in flight */
	int Tb, A4, total4, n_tiles;
	int done_t, done_last, fin_lo, fin_hi;
	long long row_base, row_size; /* traceback rows of the block: byte (row t, index i) at row_base + (t-1)*row_size + i */
	long long snap_off;           /* >= 0: the tiles of this block also save the state they load (a snapshot at score s) */
	int snap_rowsize;
	int nat_ok;            /* every slice the block in flight reads is newer than shrink_s: nothing lies outside the band (tile_fast2_block<.., true>) */
	int lo_log[TILE_TMAX], hi_log[TILE_TMAX];
};

#define TILE_CTL_HEAD 28 /* ints before lo_log */

struct TileCounters { unsigned int n_items, next; };

/* persistent scheduling (wfa_tile_persist_kernel): one queue of work items per pass, fed by the planner that runs inside the
 * tile kernel.  An item is (pair, tile) -- or (pair, -1): plan the pair's next block -- or (-1, .): retire.  Consumers take
 * tickets from `head`; the item of ticket t is published by storing its word (q_word) into q_items[t & mask]. */
struct PersistCtl {
	unsigned int head, tail;
	int n_inflight;    /* pairs with a block in flight (or a plan item queued); 0 => the retire items go out */
	int stop_req;      /* a planner asked for the other tile geometry: pairs park after their block in flight */
	int switch_to;
	int total_tiles;   /* sum of n_tiles over the blocks in flight */
	int scratch;        /* target of the release that orders a tile's alive words */
	int n_start, n_cut; /* pairs in flight when the launch began; blocks cut since (the total above is meaningful once every pair has been cut) */
};

/* segmented traceback: a snapshot of the ring state every snap_P scores lets the traceback bytes be
 * recomputed one segment of snap_P scores at a time, so a CIGAR never needs s^2 bytes at once */
struct SnapDir { int s, wflo, wfhi, A4, rowsize, pad; long long off, n_iter; };
struct TraceState { int fwd_status, s_final, i, k, row, last, cur_op, n_out; unsigned int cur_len; int pad; long long n_iter; };

struct TParams {
	Pen pen;
	int is_tb, max_s;
	long long max_iter;
	int n_pairs, pair0;        /* this wave: pairs order[pair0 .. pair0+n_pairs) */
	const int *order;
	const PairDesc *pairs;
	PairOut *outs;
	const uint8_t *seq;
	const uint32_t *seqp;      /* packed copies (wfa_pack_kernel), word offset = raw byte offset / 8 */
	const uint2 *seqp2;        /* the same words as overlapping pairs {word i, word i + 1}: one 8-byte load per probe (tile_cells_fast2) */
	const int *packed;         /* [pair index]: bits per code (2 or 4), 0 when the pair stays on raw bytes */
	uint32_t *cigar;
	TileCtl *ctl;              /* [n_pairs] */
	int32_t *state;            /* [n_pairs][2][R][pitch] */
	int32_t *alive;            /* [n_pairs][pitch] */
	int pitch, R, W, HL, T;
	int2 *items;
	TileCounters *cnt;         /* [2] */
	int *n_running;
	int *err;                  /* bits 1 << TS_ARENA, 1 << TS_SHRINK raised by the planner */
	uint8_t *arena;
	long long arena_cap;
	unsigned long long *arena_used;
	long long *rowtab;         /* [n_pairs][rowtab_stride] */
	long long rowtab_stride;
	int *seg;                  /* checkpoints (s,d) per pair [n_pairs][seg_stride] (low-memory mode), or null */
	int *n_seg;                /* [n_pairs] */
	int seg_stride, seg_use, step; /* seg_use: this pass collapses the band at the checkpoints (pass 2, miniwfa.c:413-416) */
	int s_limit;               /* no alignment of the batch can cost more (all-gap bound): a guard against endless runs */
	int fast;                  /* interior tiles of the 4-cells-per-thread geometry keep the gap rows in registers (tile_cells_fast) */
	int fast_edge;             /* tiles at the band's edges / on the terminal diagonal run the register-resident step too (tile_fast2_block<.., true>) */
	/* persistent scheduling */
	PersistCtl *pq;
	unsigned long long *q_items; /* [q_mask + 1] entries, zeroed before a pass (q_word) */
	unsigned int q_mask, q_bits;  /* q_mask + 1 = 1 << q_bits */
	int vmod;                  /* > 0: slot v stands for pair slot v % vmod (several traceback segments of a pair recomputed at once) */
	int geom_id, n_geom, many; /* this launch's geometry (0 latency, 1 throughput); switch above `many` / below many / 2 tiles in flight */
	/* segmented traceback */
	int snap_P, snapdir_stride, snap_take; /* snap_take: save a snapshot whenever s is a multiple of snap_P (a multiple of 256) */
	int32_t *snap_arena;
	long long snap_cap;           /* int32 words */
	unsigned long long *snap_used;
	SnapDir *snapdir;             /* [n_pairs][snapdir_stride] */
	int *n_snap;                  /* [n_pairs] */
	int *s_stop;                  /* [n_pairs] or null: a pass ends when a pair reaches this score */
	TraceState *trace;            /* [n_pairs] */
	/* byte offsets of the rows a score touches inside a tile, by score modulo the ring depths (wf_next_prep, :252-257) */
	int4 tabH[TILE_NRING_MAX];  /* [s % nring]  = {H[s-x], H[s-o1-e1], H[s-o2-e2], H[s]} */
	int4 tabE1[TILE_EDEPTH_MAX]; /* [s % (e1+1)] = {E1[s-e1], F1[s-e1], E1[s], F1[s]} */
	int4 tabE2[TILE_EDEPTH_MAX]; /* [s % (e2+1)] = {E2[s-e2], F2[s-e2], E2[s], F2[s]} */
};

/* the pair slot behind a (possibly virtual) slot: per-pair arrays -- order, rowtab, snapshots, trace state -- are indexed by it */
__device__ __forceinline__ int pair_slot(const TParams &P, int slot) { return P.vmod > 0 ? slot % P.vmod : slot; }

/* index of diagonal 0 in a state row: independent of the tile geometry, so that the geometry may change between launches */
__device__ __forceinline__ int tile_doff(const TParams &P, int tl) { return tl + P.pen.nring + TILE_TMAX + 8; }

/* Row index inside a state buffer / the shared-memory tile: H of score s in row s % nring; then E1, F1 (depth e1+1 each) and
 * E2, F2 (depth e2+1 each).  The byte offsets a score needs are tabulated on the host (tabH / tabE1 / tabE2 in TParams). */

/* ------------------------------------------------------------------------------------------ */
/* bulk-copy (TMA, non-tensor form) + mbarrier helpers                                          */
/* ------------------------------------------------------------------------------------------ */

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
	             :: "l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/* ------------------------------------------------------------------------------------------ */
/* score 0 (wf_stripe_init, miniwfa.c:103-121, and the first extend)                            */
/* ------------------------------------------------------------------------------------------ */

__global__ void wfa_tile_init_kernel(const TParams P, int slot0)
{
	const int slot = slot0 + blockIdx.x, pi = P.order[P.pair0 + pair_slot(P, slot)];
	const PairDesc pd = P.pairs[pi];
	const int n = P.pen.nring, doff = tile_doff(P, pd.tl);
	int32_t *st = P.state + (size_t)slot * 2 * P.R * P.pitch;
	const int span = 2 * (n + 2) + 1; /* every slice reads as NEG_INF around diagonal 0 */
	for (int t = threadIdx.x; t < P.R * span; t += blockDim.x) {
		const int row = t / span, c = t % span;
		st[(size_t)row * P.pitch + doff - (n + 2) + c] = NEG_INF;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		const uint32_t *T = reinterpret_cast<const uint32_t*>(P.seq + pd.t_off), *Q = reinterpret_cast<const uint32_t*>(P.seq + pd.q_off);
		const int k = extend_run(T, Q, -1, 0, min(pd.tl - 1, pd.ql - 1));
		st[doff] = k; /* H of score 0 lives in slot 0 */
		TileCtl *c = P.ctl + slot;
		c->s = 0, c->wflo = c->wfhi = 0, c->cur = 0, c->last = 0, c->sid = 0, c->copied = 0, c->n_iter = 0, c->shrink_s = -0x10000, c->nat_ok = 0;
		c->Tb = 0, c->n_tiles = 0, c->done_t = 0x7fffffff, c->done_last = 0, c->snap_off = -1, c->tiles_done = 0;
		c->status = (k == pd.tl - 1 && k == pd.ql - 1) ? TS_DONE : TS_RUN;
		if (c->status == TS_DONE) {
			PairOut o;
			o.s = 0, o.n_cigar = 0, o.n_iter = 0, o.cigar_pos = pd.cigar_off + pd.cigar_cap, o.status = ST_OK, o.pad_ = 0;
			o.end_s = 0, o.end_i = -1, o.end_k = -1, o.pad2_ = 0;
			P.outs[pi] = o;
		} else atomicAdd(P.n_running, 1);
	}
}

/* ------------------------------------------------------------------------------------------ */
/* between blocks: replay the finished block, trim, cut the next one                            */
/* ------------------------------------------------------------------------------------------ */

/* queue helpers of the persistent scheduler */
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int *p)
{
	unsigned int v;
	asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned int *p, unsigned int v)
{
	asm volatile("st.volatile.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
/* A queue entry is ONE 64-bit word: lap tag (20 bits, never 0) | slot + 1 (22 bits) | tile + 1 (22 bits), so that publishing is a
 * single store and taking an item a single load -- no payload / sequence-word pair and no second round trip to L2. */
__device__ __forceinline__ unsigned long long q_word(const TParams &P, unsigned int ticket, int slot, int tile)
{
	const unsigned long long tag = (ticket >> P.q_bits) % 0xfffffu + 1u;
	return tag << 44 | (unsigned long long)(unsigned int)(slot + 1) << 22 | (unsigned long long)(unsigned int)(tile + 1);
}
/* publish item `ticket` (a release, not __threadfence(): that one also drops the SM's L1) */
__device__ __forceinline__ void q_publish(const TParams &P, unsigned int ticket, int slot, int tile)
{
	asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(P.q_items + (ticket & P.q_mask)), "l"(q_word(P, ticket, slot, tile)) : "memory");
}

/* wait for the item of a ticket taken from `head` (one thread per CTA).  No acquire fence: it would drop the L1 lines of all four
 * CTAs of the SM (the sequence windows) on every item; the producer released the word after everything the item refers to, and
 * everything other SMs write inside this kernel is read past L1 (ld.cg, bulk copies).  Not inlined: the tile kernel sits at its
 * register limit. */
__device__ __noinline__ int2 q_take(const TParams &P, const unsigned int ticket)
{
	const unsigned long long *q = P.q_items + (ticket & P.q_mask);
	const unsigned int tag = (ticket >> P.q_bits) % 0xfffffu + 1u;
	unsigned int ns = 32, spins = 0;
	unsigned long long w;
	for (;;) {
		asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(q) : "memory");
		if ((unsigned int)(w >> 44) == tag) break;
		__nanosleep(ns);
		if (ns < 256) ns <<= 1;
		if (++spins == 0x30000000u) __trap(); /* minutes without an item: a lost hand-over must end as a launch failure, not as a device that hangs for good */
	}
	return make_int2((int)((unsigned int)(w >> 22) & 0x3fffffu) - 1, (int)((unsigned int)w & 0x3fffffu) - 1);
}

/* The planner of one pair: replay the block that has just finished, trim, cut the next block and hand out its tiles.
 * PERSIST = false: wfa_plan_kernel, one CTA per pair between two launches of the tile kernel, tiles appended to the item list
 * of launch `it`.  PERSIST = true: called inside wfa_tile_persist_kernel by the CTA that finished the last tile of the pair's
 * block (or took the pair's plan item); tiles go to the queue; `retire` is set when the last pair in flight has ended or
 * parked, and the caller then publishes the retire items.  All threads of the CTA must call it. */
template<bool PERSIST>
__device__ __noinline__ void plan_pair(const TParams &P, int slot, int it, int *retire)
{
	__shared__ int sh[4];
	__shared__ long long sh_row[2];
	__shared__ int sh_emit[4]; /* items base, n_tiles, first score of the rows, number of rows */
	static_assert(offsetof(TileCtl, lo_log) == 4 * TILE_CTL_HEAD, "TileCtl: the scalar fields come first");
	/* scratch: inside the persistent kernel the CTA's tile rows are idle while it plans (the kernel's static shared memory decides
	 * whether a fourth CTA fits an SM); the one-CTA-per-pair plan kernel has its own */
	__shared__ __align__(8) int scratch_s[PERSIST ? 2 : TILE_TMAX + TILE_CTL_HEAD + 4];
	extern __shared__ __align__(128) int32_t smem_tile[];
	int *const head = PERSIST ? reinterpret_cast<int*>(smem_tile) : scratch_s;
	int *const widths = head + TILE_CTL_HEAD + 4;
	const int pslot = pair_slot(P, slot), pi = P.order[P.pair0 + pslot];
	TileCtl *c = P.ctl + slot;
	/* TileCtl is written by other SMs inside the persistent kernel: read from L2, and all of it at once -- its scalar fields, the
	 * per-score bounds the tiles logged and the stop request come in with ONE round trip (a block of a single large pair waits for
	 * this function; field by field it was ten dependent round trips, ~7 us).  CL(f) reads the copy; the few fields thread 0 changes
	 * and reads again go through CSET. */
	TileCtl *hw = reinterpret_cast<TileCtl*>(head);
#define CL(f) (const_cast<const TileCtl*>(hw)->f)
#define CSET(f, v) do { c->f = (v); hw->f = (v); } while (0)
	__syncthreads(); /* (the shared words may still be read by the previous call) */
	for (int t = threadIdx.x; t <= TILE_TMAX + TILE_CTL_HEAD; t += blockDim.x) { /* (CTAs of 64 threads and more) */
		if (t < TILE_TMAX) widths[t] = __ldcg(&c->hi_log[t]) - __ldcg(&c->lo_log[t]) + 1;
		else if (t < TILE_TMAX + TILE_CTL_HEAD) head[t - TILE_TMAX] = __ldcg(reinterpret_cast<const int*>(c) + (t - TILE_TMAX));
		else head[TILE_CTL_HEAD] = PERSIST ? *(volatile int*)&P.pq->stop_req : 0;
	}
	__syncthreads();
	if (CL(status) != TS_RUN) return;
	const PairDesc pd = P.pairs[pi];
	const int tl = pd.tl, ql = pd.ql, n = P.pen.nring, doff = tile_doff(P, tl);
	int status = TS_RUN, s = CL(s), wflo = CL(wflo), wfhi = CL(wfhi);
	const int Tb_done = CL(Tb);
	if (Tb_done > 0) { /* replay, in the order of miniwfa.c:419-426 */
		if (threadIdx.x == 0) {
			long long n_iter = CL(n_iter);
			const int Tb = Tb_done, s0 = s, done_t = CL(done_t);
			int last = 0;
			for (int t = 1; t <= Tb; ++t) {
				n_iter += widths[t - 1];
				s = s0 + t;
				if ((P.max_iter > 0 && n_iter > P.max_iter) || (P.max_s > 0 && s > P.max_s)) { status = TS_STOPPED; break; }
				if (done_t == t) { status = TS_DONE; last = CL(done_last); break; }
			}
			CSET(n_iter, n_iter); c->s = s, c->last = last; /* (the other threads may still be reading s from the copy) */
			c->wflo = CL(fin_lo), c->wfhi = CL(fin_hi), c->cur = CL(cur) ^ 1;
			sh[0] = status, sh[1] = s;
		}
		__syncthreads();
		status = sh[0], s = sh[1], wflo = CL(fin_lo), wfhi = CL(fin_hi);
		__syncthreads();
		if (status == TS_RUN && (s & 0xff) == 0) { /* wf_stripe_shrink (:144-171) from the tiles' alive words */
			const int32_t *alive = P.alive + (size_t)slot * P.pitch + doff;
			const int tag = s | 1, lane = threadIdx.x & 31;
			if (threadIdx.x < 32) {
				int nl = wfhi + 1;
				for (int base = wflo; base <= wfhi; base += 32) {
					const int d = base + lane;
					const unsigned m = __ballot_sync(0xffffffffu, d <= wfhi && __ldcg(alive + d) == tag);
					if (m) { nl = base + __ffs(m) - 1; break; }
				}
				int nh = nl - 1;
				for (int base = wfhi; base >= nl; base -= 32) {
					const int d = base - lane;
					const unsigned m = __ballot_sync(0xffffffffu, d >= nl && __ldcg(alive + d) == tag);
					if (m) { nh = base - (__ffs(m) - 1); break; }
				}
				if (lane == 0) sh[2] = nl, sh[3] = nh;
			}
			__syncthreads();
			const int nl = sh[2], nh = sh[3];
			if (nl > wfhi || nh < nl) status = TS_SHRINK; /* the reference asserts (:157, :169) */
			else {
				if (threadIdx.x == 0 && (nl != wflo || nh != wfhi)) CSET(shrink_s, s);
				wflo = nl, wfhi = nh;
			}
		}
	}
	if (threadIdx.x == 0) sh_emit[1] = 0, sh_emit[3] = 0;
	__syncthreads();
	if (threadIdx.x == 0) {
	bool park = false;
	if (status == TS_RUN && s > P.s_limit) status = TS_SHRINK; /* cannot happen: the all-gap alignment costs less */
	if (status == TS_RUN && P.s_stop && s >= P.s_stop[slot]) status = TS_SEGEND; /* end of a traceback segment */
	if (PERSIST && status == TS_RUN && head[TILE_CTL_HEAD]) { /* the other tile geometry takes over: the next launch cuts the block */
		park = true;
		c->wflo = wflo, c->wfhi = wfhi;
		atomicSub(&P.pq->total_tiles, CL(n_tiles));
		c->Tb = 0, c->n_tiles = 0;
	}
	if (status == TS_RUN && !park) { /* cut the next block */
		int sid = CL(sid), Tb, copy_only = 0;
		const int n_seg = P.seg_use ? P.n_seg[slot] : 0;
		const int *seg = P.seg + (size_t)slot * P.seg_stride;
		if (sid < n_seg && seg[2 * sid] == s) { /* band collapse (:413-416) */
			if (!CL(copied)) copy_only = 1, c->copied = 1; /* first bring both state buffers to the same contents: the narrow blocks
			                                                  that follow store only their own columns, but still read the wide slices */
			else {
				wflo = wfhi = seg[2 * sid + 1];
				++sid, c->copied = 0; CSET(shrink_s, s);
			}
		}
		Tb = min(P.T, ((s | 0xff) + 1) - s);
		if (sid < n_seg && seg[2 * sid] > s) Tb = min(Tb, seg[2 * sid] - s);
		if (copy_only) Tb = 0;
		if (P.max_s > 0) Tb = min(Tb, P.max_s + 1 - s);
		if (P.s_stop) Tb = min(Tb, P.s_stop[slot] - s);
		if (P.is_tb) Tb = min(Tb, (int)P.rowtab_stride - 1 - s); /* no score lies beyond the all-gap alignment */
		Tb = copy_only ? 0 : max(Tb, 1);
		const int lo_sup = max(wflo - Tb, -tl) - n, hi_sup = min(wfhi + Tb, ql) + n;
		const int A4 = (lo_sup + doff) & ~3, Bx = (hi_sup + doff) | 3;
		const int total4 = (Bx - A4 + 1) >> 2, umax = P.W - 2 * P.HL;
		const int n_tiles = (total4 * 4 + umax - 1) / umax;
		if (P.is_tb && Tb > 0) { /* wf_tb_add (:33-44): one row per score, here as wide as the block's superset */
			const long long rowsize = Bx - A4 + 1;
			const unsigned long long base = atomicAdd(P.arena_used, (unsigned long long)(rowsize * Tb));
			if ((long long)base + rowsize * Tb > P.arena_cap || s + Tb >= P.rowtab_stride) status = TS_ARENA;
			else {
				sh_row[0] = (long long)base - A4, sh_row[1] = rowsize, sh_emit[2] = s, sh_emit[3] = Tb; /* the table is filled below, by all threads */
				c->row_base = (long long)base - A4, c->row_size = rowsize;
			}
		}
		c->snap_off = -1;
		if (status == TS_RUN && P.snap_take && s > 0 && s % P.snap_P == 0 && Tb > 0) { /* the tiles save what they load */
			const int k = s / P.snap_P - 1, rowsize = Bx - A4 + 1;
			const long long words = (long long)P.R * rowsize;
			const unsigned long long off = atomicAdd(P.snap_used, (unsigned long long)words);
			if (k >= P.snapdir_stride || (long long)off + words > P.snap_cap) status = TS_ARENA;
			else {
				SnapDir d;
				d.s = s, d.wflo = wflo, d.wfhi = wfhi, d.A4 = A4, d.rowsize = rowsize, d.pad = 0, d.off = (long long)off, d.n_iter = CL(n_iter);
				P.snapdir[(size_t)slot * P.snapdir_stride + k] = d;
				P.n_snap[slot] = k + 1;
				c->snap_off = (long long)off, c->snap_rowsize = rowsize;
			}
		}
		if (status == TS_RUN) {
			const int old_tiles = CL(n_tiles);
			c->wflo = wflo, c->wfhi = wfhi, c->sid = sid;
			c->Tb = Tb, c->A4 = A4, c->total4 = total4, c->n_tiles = n_tiles;
			c->done_t = 0x7fffffff, c->done_last = 0, c->tiles_done = 0;
			c->nat_ok = s - CL(shrink_s) >= n;
			if (PERSIST) {
				/* (issued back to back, used afterwards: each is a round trip to L2 on the critical path of a single pair's block) */
				const unsigned int base_t = atomicAdd(&P.pq->tail, (unsigned int)n_tiles);
				const int tot0 = atomicAdd(&P.pq->total_tiles, n_tiles - old_tiles);
				int n_cut = 0, n_start = 0;
				if (P.n_geom > 1) { n_cut = atomicAdd(&P.pq->n_cut, 1) + 1; n_start = *(volatile int*)&P.pq->n_start; }
				sh_emit[0] = (int)base_t, sh_emit[1] = n_tiles;
				const int tot = tot0 + n_tiles - old_tiles;
				if (P.n_geom > 1 && n_cut >= n_start && ((P.geom_id == 0 && tot >= P.many) || (P.geom_id == 1 && tot < P.many / 2)) && !head[TILE_CTL_HEAD]) { /* (a request raised since the copy was taken is raised again, to the same geometry) */
					P.pq->switch_to = P.geom_id ^ 1;
					__threadfence();
					*(volatile int*)&P.pq->stop_req = 1;
				}
			} else sh_emit[0] = (int)atomicAdd(&P.cnt[it & 1].n_items, (unsigned int)n_tiles), sh_emit[1] = n_tiles;
		}
	}
	if (status != TS_RUN) {
		c->status = status;
		if (status == TS_ARENA || status == TS_SHRINK) atomicOr(P.err, 1 << status);
		atomicSub(P.n_running, 1);
		if (PERSIST) atomicSub(&P.pq->total_tiles, CL(n_tiles));
		if (status != TS_SEGEND) { /* (a segment's end leaves the result record to the pass that reaches the end) */
			PairOut o;
			o.s = status == TS_DONE ? s : -1;
			o.n_cigar = 0, o.n_iter = CL(n_iter), o.cigar_pos = pd.cigar_off + pd.cigar_cap;
			o.status = status == TS_DONE ? ST_OK : status == TS_STOPPED ? ST_STOPPED : status == TS_ARENA ? ST_ARENA : ST_SHRINK;
			o.pad_ = 0;
			o.end_s = 0, o.end_i = 0, o.end_k = 0, o.pad2_ = 0;
			P.outs[pi] = o;
		}
	}
	if (PERSIST && (status != TS_RUN || park)) { /* this pair has no block in flight any more */
		__threadfence();
		if (atomicSub(&P.pq->n_inflight, 1) == 1) *retire = 1;
	}
	} /* thread 0 */
	__syncthreads();
	if (PERSIST) { /* the block's tiles: ONE release per thread, the rest of its words relaxed (a MEMBAR each would cost a large pair
	                * with thousands of tiles per block tens of microseconds in which the whole GPU waits) */
		const unsigned int base = (unsigned int)sh_emit[0];
		const int n_emit = sh_emit[1];
		bool first = true;
		for (int j = threadIdx.x; j < n_emit; j += blockDim.x) {
			unsigned long long *q = P.q_items + ((base + j) & P.q_mask);
			const unsigned long long w = q_word(P, base + j, slot, j);
			if (first) asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(q), "l"(w) : "memory");
			else asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(q), "l"(w) : "memory"); /* (after the release in program order) */
			first = false;
		}
	} else {
		for (int j = threadIdx.x; j < sh_emit[1]; j += blockDim.x) P.items[sh_emit[0] + j] = make_int2(slot, j);
	}
	if (threadIdx.x < sh_emit[3]) /* wf_tb_add (:33-44): where the traceback row of every score of the block starts */
		P.rowtab[(size_t)pslot * P.rowtab_stride + sh_emit[2] + 1 + threadIdx.x] = sh_row[0] + (long long)threadIdx.x * sh_row[1];
}
#undef CL
#undef CSET

__global__ void __launch_bounds__(128) wfa_plan_kernel(const __grid_constant__ TParams P, int it)
{
	if (blockIdx.x == 0 && threadIdx.x == 0) { P.cnt[(it + 1) & 1].n_items = 0; P.cnt[(it + 1) & 1].next = 0; }
	plan_pair<false>(P, blockIdx.x, it, 0);
}

/* ------------------------------------------------------------------------------------------ */
/* the tile kernel                                                                             */
/* ------------------------------------------------------------------------------------------ */

/* shared-memory accesses with explicit 32-bit shared addresses (one add per row, no generic-address arithmetic);
 * N = 1, 2 or 4 consecutive int32 */
template<int N> __device__ __forceinline__ void ldsv(uint32_t a, int (&v)[N]);
template<> __device__ __forceinline__ void ldsv<4>(uint32_t a, int (&v)[4])
{
	asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(a) : "memory");
}
template<> __device__ __forceinline__ void ldsv<2>(uint32_t a, int (&v)[2])
{
	asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(a) : "memory");
}
template<> __device__ __forceinline__ void ldsv<1>(uint32_t a, int (&v)[1])
{
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v[0]) : "r"(a) : "memory");
}
template<int N> __device__ __forceinline__ void stsv(uint32_t a, const int (&v)[N]);
template<> __device__ __forceinline__ void stsv<4>(uint32_t a, const int (&v)[4])
{
	asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
template<> __device__ __forceinline__ void stsv<2>(uint32_t a, const int (&v)[2])
{
	asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(a), "r"(v[0]), "r"(v[1]) : "memory");
}
template<> __device__ __forceinline__ void stsv<1>(uint32_t a, const int (&v)[1])
{
	asm volatile("st.shared.b32 [%0], %1;" :: "r"(a), "r"(v[0]) : "memory");
}
__device__ __forceinline__ int lds1_if(uint32_t a, bool p) /* NEG_INF unless p; the load is predicated, not branched around */
{
	int v = NEG_INF;
	asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.shared.b32 %0, [%1]; }" : "+r"(v) : "r"(a), "r"((int)p) : "memory");
	return v;
}

/*
 * The two sequences of a pair as the match-run probe sees them (wf_extend1_padded, miniwfa.c:212-226): 32-bit words holding
 * either 4 bytes (raw) or, when the pair uses at most 4 distinct byte values, 16 two-bit codes (wfa_pack2_kernel).  Equal
 * codes <=> equal bytes, so the run lengths are the same; the packed form touches 4x fewer cache lines per probe (the kernel is
 * bound by the L1 data pipe) and one probe covers 16 bases, so the "longer than one probe" path is left to the optimal path.
 * The three shift amounts make one code path serve both forms.
 */
struct SeqView {
	const uint32_t *T, *Q;
	int s_idx, s_amt, s_adv; /* position -> word index (>> 2 | 4), funnel amount (<< 3 | 1), differing bit -> position (>> 3 | 1) */
	__device__ __forceinline__ uint32_t word(const uint32_t *__restrict__ w, int pos) const
	{
		const uint32_t *p = w + (pos >> s_idx);
		return __funnelshift_r(__ldg(p), __ldg(p + 1), pos << s_amt); /* the funnel shift takes its amount modulo 32 */
	}
	__device__ __forceinline__ uint32_t probe(int tp, int qp) const { return word(T, tp) ^ word(Q, qp); }
	__device__ __forceinline__ int advance(uint32_t x) const { return __clz(__brev(x)) >> s_adv; } /* all positions of the word when x == 0 */
};

/* continue a match run: everything up to k is known to match, clamped to kmax.  Four words of each sequence per round (16 bytes
 * or 64 packed bases), all ten loads in flight together: on the optimal path a run is ~1/divergence bases long and this loop is
 * the serial part of a score step. */
__device__ __noinline__ int tile_extend_more(const SeqView sv, int k, int d, int kmax)
{
	const int ppw = 32 >> sv.s_adv; /* positions per word */
	while (k < kmax) {
		const int tp = k + 1, qp = d + k + 1;
		const uint32_t *tw = sv.T + (tp >> sv.s_idx), *qw = sv.Q + (qp >> sv.s_idx);
		uint32_t a[5], b[5];
#pragma unroll
		for (int i = 0; i < 5; ++i) a[i] = __ldg(tw + i), b[i] = __ldg(qw + i);
		int adv = 4 * ppw;
#pragma unroll
		for (int i = 3; i >= 0; --i) {
			const uint32_t x = __funnelshift_r(a[i], a[i + 1], tp << sv.s_amt) ^ __funnelshift_r(b[i], b[i + 1], qp << sv.s_amt);
			if (x) adv = i * ppw + ((__ffs(x) - 1) >> sv.s_adv);
		}
		k += adv;
		if (adv < 4 * ppw) break;
	}
	return min(k, kmax);
}

/* what one score step leaves in registers for the caller: the new cells of this thread */
template<int CPT> struct CellOut { int H[CPT], E1[CPT], F1[CPT], E2[CPT], F2[CPT]; uint32_t tb; };

/* one score step for the CPT consecutive diagonals of this thread; sb = shared address of the tile + 4 * CPT * tid.
 * qh = {Hx, Ho1, Ho2, nH}, q1 = {pE1, pF1, nE1, nF1}, q2 = {pE2, pF2, nE2, nF2}: byte offsets of the rows (wf_next_prep, :252-257).
 * The d-1 / d+1 neighbours come from registers inside the thread, from warp shuffles across threads, and from two
 * scalar shared loads across warps.
 * EDGE = false: the whole warp lies strictly inside the band and does not hold the terminal diagonal,
 * so no masking, no edge rule, no termination test. */
template<int MODE, bool EDGE, int CPT>
__device__ __forceinline__ int tile_cells(uint32_t sb, const int4 &qh, const int4 &q1, const int4 &q2, int d0, int lo_t, int hi_t, int dfin, int tl,
                                          const int (&kmin)[CPT], const int (&kspan)[CPT],
                                          const SeqView &sv,
                                          bool no_left, bool no_right, bool useful, CellOut<CPT> &o)
{
	const int lane = threadIdx.x & 31;
	int ho1[CPT], pe1[CPT], pf1[CPT], ho2[CPT], pe2[CPT], pf2[CPT], hx[CPT];
	ldsv<CPT>(sb + qh.y, ho1); ldsv<CPT>(sb + q1.x, pe1); ldsv<CPT>(sb + q1.y, pf1); ldsv<CPT>(sb + qh.z, ho2);
	ldsv<CPT>(sb + q2.x, pe2); ldsv<CPT>(sb + q2.y, pf2); ldsv<CPT>(sb + qh.x, hx);
	int A1[CPT + 2], A2[CPT + 2], C1[CPT + 2], C2[CPT + 2], bA1[CPT + 2], bA2[CPT + 2], bC1[CPT + 2], bC2[CPT + 2];
#pragma unroll
	for (int j = 0; j < CPT; ++j) {
		A1[j + 1] = max(ho1[j], pe1[j]), A2[j + 1] = max(ho2[j], pe2[j]);
		C1[j + 1] = max(ho1[j], pf1[j]), C2[j + 1] = max(ho2[j], pf2[j]); /* the +1 of F is applied below */
		if (MODE != MODE_SCORE) bA1[j + 1] = ho1[j] < pe1[j], bA2[j + 1] = ho2[j] < pe2[j], bC1[j + 1] = ho1[j] < pf1[j], bC2[j + 1] = ho2[j] < pf2[j];
	}
	A1[0] = __shfl_up_sync(0xffffffffu, A1[CPT], 1);
	A2[0] = __shfl_up_sync(0xffffffffu, A2[CPT], 1);
	C1[CPT + 1] = __shfl_down_sync(0xffffffffu, C1[1], 1);
	C2[CPT + 1] = __shfl_down_sync(0xffffffffu, C2[1], 1);
	if (MODE != MODE_SCORE) {
		const int bl = __shfl_up_sync(0xffffffffu, bA1[CPT] | bA2[CPT] << 1, 1);
		const int br = __shfl_down_sync(0xffffffffu, bC1[1] | bC2[1] << 1, 1);
		bA1[0] = bl & 1, bA2[0] = bl >> 1, bC1[CPT + 1] = br & 1, bC2[CPT + 1] = br >> 1;
	}
	{ /* lanes 0 and 31: the neighbour belongs to another warp (or to nobody: stale halo) -- four predicated scalar loads, no branch */
		const bool left = lane == 0, edge_lane = (left && !no_left) || (lane == 31 && !no_right);
		const uint32_t nb = sb + (left ? -4 : 4 * CPT);
		const int o1 = lds1_if(nb + qh.y, edge_lane), o2 = lds1_if(nb + qh.z, edge_lane);
		const int x1 = lds1_if(nb + (left ? q1.x : q1.y), edge_lane), x2 = lds1_if(nb + (left ? q2.x : q2.y), edge_lane);
		const int m1 = max(o1, x1), m2 = max(o2, x2);
		if (lane == 0) A1[0] = m1, A2[0] = m2;
		if (lane == 31) C1[CPT + 1] = m1, C2[CPT + 1] = m2;
		if (MODE != MODE_SCORE) {
			if (lane == 0) bA1[0] = o1 < x1, bA2[0] = o2 < x2;
			if (lane == 31) bC1[CPT + 1] = o1 < x1, bC2[CPT + 1] = o2 < x2;
		}
	}
	int st[CPT], h0[CPT];
	bool ext[CPT];
	uint32_t tbw = 0;
	int myfl = 0;
#pragma unroll
	for (int j = 0; j < CPT; ++j) {
		int E1 = A1[j], E2 = A2[j], F1 = C1[j + 2] + 1, F2 = C2[j + 2] + 1;
		const int e = max(E1, E2), f = max(F1, F2), gmx = max(e, f), hxp = hx[j] + 1;
		int H = max(hxp, gmx);
		st[j] = 0;
		if (MODE != MODE_SCORE) { /* the 7-bit pack, miniwfa.c:290-306 */
			const int z = hxp >= gmx ? 0 : (e >= f ? (E1 >= E2 ? 1 : 3) : (F1 >= F2 ? 2 : 4));
			st[j] = z;
			tbw |= (uint32_t)(z | bA1[j] << 3 | bC1[j + 2] << 4 | bA2[j] << 5 | bC2[j + 2] << 6) << (8 * j);
		}
		if (EDGE) {
			const int d = d0 + j;
			if (d < lo_t || d > hi_t) H = E1 = E2 = F1 = F2 = NEG_INF; /* outside the slice: what the reference's pads hold */
			else if (H >= -1 || E1 >= -1 || F1 >= -1 || E2 >= -1 || F2 >= -1) { /* edge rule, :325-326 */
				if (d == lo_t) myfl |= FL_LO;
				if (d == hi_t) myfl |= FL_HI;
			}
		}
		o.E1[j] = E1, o.E2[j] = E2, o.F1[j] = F1, o.F2[j] = F2, h0[j] = H;
		ext[j] = (unsigned)(H - kmin[j]) <= (unsigned)kspan[j]; /* on the matrix (:402) */
	}
	/* wf_extend (:400-411): first probe of the match run, 4 bytes per sequence, all loads in flight together */
	uint32_t px[CPT];
#pragma unroll
	for (int j = 0; j < CPT; ++j) {
		const int tp = ext[j] ? h0[j] + 1 : 0, qp = ext[j] ? d0 + j + h0[j] + 1 : 0;
		px[j] = sv.probe(tp, qp);
	}
	stsv<CPT>(sb + q1.z, o.E1); stsv<CPT>(sb + q1.w, o.F1); stsv<CPT>(sb + q2.z, o.E2); stsv<CPT>(sb + q2.w, o.F2);
	bool more = false;
	bool unres[CPT];
#pragma unroll
	for (int j = 0; j < CPT; ++j) {
		const int kmax = kmin[j] + kspan[j];
		const int adv = sv.advance(px[j]); /* positions up to the first difference; the whole word when all match */
		const int k = min(h0[j] + adv, kmax);
		unres[j] = ext[j] && px[j] == 0 && k < kmax;
		more |= unres[j];
		o.H[j] = ext[j] ? k : h0[j];
	}
	if (more) { /* rare: a run longer than the first probe */
#pragma unroll
		for (int j = 0; j < CPT; ++j)
			if (unres[j]) o.H[j] = tile_extend_more(sv, o.H[j], d0 + j, kmin[j] + kspan[j]);
	}
	if (EDGE && useful && dfin >= d0 && dfin < d0 + CPT) { /* end of both sequences, :405-409 */
#pragma unroll
		for (int j = 0; j < CPT; ++j)
			if (d0 + j == dfin && ext[j] && o.H[j] == tl - 1) {
				myfl |= FL_DONE;
				if (MODE == MODE_TB && o.H[j] == h0[j]) myfl |= st[j] << FL_LAST_SHIFT;
			}
	}
	stsv<CPT>(sb + qh.w, o.H);
	o.tb = tbw;
	return myfl;
}

/* split-phase step barrier: a warp arrives when its stores of a score are done and waits only where it first needs another
 * warp's cells -- the d-1 / d+1 neighbours of its outermost diagonals -- so most of the next score overlaps the wait */
__device__ __forceinline__ void step_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

/* the interior step (EDGE = false) for 4 cells per thread, reordered around the split-phase barrier:
 *   loads of the own columns, cells 1 and 2 of every thread and their sequence probes  -- need nothing from other warps
 *   wait(previous score of all warps)
 *   neighbour columns of lanes 0 / 31, cells 0 and 3, then all stores                  -- stores only now: the ring slots they
 *                                                                                         overwrite were still being read
 * Same arithmetic as tile_cells<MODE, false, 4>. */
template<int MODE>
__device__ __forceinline__ void tile_cells_overlap(uint32_t sb, const int4 &qh, const int4 &q1, const int4 &q2, int d0,
                                                   const int (&kmin)[4], const int (&kspan)[4],
                                                   const SeqView &sv, bool reuse_e2,
                                                   bool no_left, bool no_right, uint64_t *stepbar, bool wait, uint32_t parity, CellOut<4> &o)
{
	const int lane = threadIdx.x & 31;
	int ho1[4], pe1[4], pf1[4], ho2[4], pe2[4], pf2[4], hx[4];
	ldsv<4>(sb + qh.y, ho1); ldsv<4>(sb + q1.x, pe1); ldsv<4>(sb + q1.y, pf1); ldsv<4>(sb + qh.z, ho2);
	if (reuse_e2) { /* e2 == 1: E2 / F2 of the previous score are this thread's own cells of the previous step, still in registers */
#pragma unroll
		for (int j = 0; j < 4; ++j) pe2[j] = o.E2[j], pf2[j] = o.F2[j];
	} else { ldsv<4>(sb + q2.x, pe2); ldsv<4>(sb + q2.y, pf2); }
	ldsv<4>(sb + qh.x, hx);
	int A1[6], A2[6], C1[6], C2[6], bA1[6], bA2[6], bC1[6], bC2[6];
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		A1[j + 1] = max(ho1[j], pe1[j]), A2[j + 1] = max(ho2[j], pe2[j]);
		C1[j + 1] = max(ho1[j], pf1[j]), C2[j + 1] = max(ho2[j], pf2[j]);
		if (MODE != MODE_SCORE) bA1[j + 1] = ho1[j] < pe1[j], bA2[j + 1] = ho2[j] < pe2[j], bC1[j + 1] = ho1[j] < pf1[j], bC2[j + 1] = ho2[j] < pf2[j];
	}
	A1[0] = __shfl_up_sync(0xffffffffu, A1[4], 1);
	A2[0] = __shfl_up_sync(0xffffffffu, A2[4], 1);
	C1[5] = __shfl_down_sync(0xffffffffu, C1[1], 1);
	C2[5] = __shfl_down_sync(0xffffffffu, C2[1], 1);
	if (MODE != MODE_SCORE) {
		const int bl = __shfl_up_sync(0xffffffffu, bA1[4] | bA2[4] << 1, 1);
		const int br = __shfl_down_sync(0xffffffffu, bC1[1] | bC2[1] << 1, 1);
		bA1[0] = bl & 1, bA2[0] = bl >> 1, bC1[5] = br & 1, bC2[5] = br >> 1;
	}
	int h0[4];
	bool ext[4];
	uint32_t px[4], tbw = 0;
#define TILE_CELL(j) do { \
		const int E1 = A1[j], E2 = A2[j], F1 = C1[j + 2] + 1, F2 = C2[j + 2] + 1; \
		const int e = max(E1, E2), f = max(F1, F2), gmx = max(e, f), hxp = hx[j] + 1; \
		const int H = max(hxp, gmx); \
		if (MODE != MODE_SCORE) { \
			const int z = hxp >= gmx ? 0 : (e >= f ? (E1 >= E2 ? 1 : 3) : (F1 >= F2 ? 2 : 4)); \
			tbw |= (uint32_t)(z | bA1[j] << 3 | bC1[j + 2] << 4 | bA2[j] << 5 | bC2[j + 2] << 6) << (8 * j); \
		} \
		o.E1[j] = E1, o.E2[j] = E2, o.F1[j] = F1, o.F2[j] = F2, h0[j] = H; \
		ext[j] = (unsigned)(H - kmin[j]) <= (unsigned)kspan[j]; \
		const int tp = ext[j] ? H + 1 : 0, qp = ext[j] ? d0 + j + H + 1 : 0; \
		px[j] = sv.probe(tp, qp); \
	} while (0)
	TILE_CELL(1);
	TILE_CELL(2);
	if (wait) mbar_wait(stepbar, parity); /* every warp has finished the previous score */
	{
		const bool left = lane == 0, edge_lane = (left && !no_left) || (lane == 31 && !no_right);
		const uint32_t nb = sb + (left ? -4 : 16);
		const int o1 = lds1_if(nb + qh.y, edge_lane), o2 = lds1_if(nb + qh.z, edge_lane);
		const int x1 = lds1_if(nb + (left ? q1.x : q1.y), edge_lane), x2 = lds1_if(nb + (left ? q2.x : q2.y), edge_lane);
		const int m1 = max(o1, x1), m2 = max(o2, x2);
		if (lane == 0) A1[0] = m1, A2[0] = m2;
		if (lane == 31) C1[5] = m1, C2[5] = m2;
		if (MODE != MODE_SCORE) {
			if (lane == 0) bA1[0] = o1 < x1, bA2[0] = o2 < x2;
			if (lane == 31) bC1[5] = o1 < x1, bC2[5] = o2 < x2;
		}
	}
	TILE_CELL(0);
	TILE_CELL(3);
#undef TILE_CELL
	stsv<4>(sb + q1.z, o.E1); stsv<4>(sb + q1.w, o.F1);
	stsv<4>(sb + q2.z, o.E2); stsv<4>(sb + q2.w, o.F2);
	bool more = false;
	bool unres[4];
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int kmax = kmin[j] + kspan[j];
		const int adv = sv.advance(px[j]);
		const int k = min(h0[j] + adv, kmax);
		unres[j] = ext[j] && px[j] == 0 && k < kmax;
		more |= unres[j];
		o.H[j] = ext[j] ? k : h0[j];
	}
	if (more) {
#pragma unroll
		for (int j = 0; j < 4; ++j)
			if (unres[j]) o.H[j] = tile_extend_more(sv, o.H[j], d0 + j, kmin[j] + kspan[j]);
	}
	stsv<4>(sb + qh.w, o.H);
	o.tb = tbw;
}

/*
 * The interior step with the gap rows in registers ("fast" path: 4 cells per thread, e1 <= 2 and e2 <= 2).
 *
 * E1/F1 of score s-e1 and E2/F2 of score s-e2 at a thread's own diagonals are what the same thread computed e1 / e2 steps
 * earlier, so they stay in registers (two sets when the depth is 2; the step loop is unrolled by two and alternates them).
 * Shared memory keeps the H ring, plus of the gap rows only what another warp reads -- the last cell of lane 31 (E rows) and the
 * first cell of lane 0 (F rows) -- until the last e1 / e2 steps of the block, whose gap rows are stored whole for the bulk copy
 * back to the state buffer.  Per step a thread issues 3 row loads and 1 row store instead of 5-7 and 5, and the row offsets of
 * the step come from a small per-block table in shared memory (3 broadcast loads) instead of indexed kernel-parameter reads.
 * The integer ALU pipe is what bounds this kernel (two cycles per warp instruction): the probe positions are kept pre-multiplied
 * by the code width, so that one shift gives the word index and the funnel shift takes its amount as it is; the first differing
 * code is popc(~x & (x - 1)), which needs no test for x == 0.
 * Same arithmetic as tile_cells<MODE, false, 4> (wf_next_score / wf_next_tb, miniwfa.c:261-308; first probe of
 * wf_extend1_padded, :212-226).
 */
struct StepTab { int4 h, e; }; /* {Hx, Ho1, Ho2, nH}, {pE1, pE2, nE1, nE2}: byte offsets of the rows of one step; an F row lies a fixed distance after its E row */

struct FastCtx {
	uint32_t sb;              /* shared address of this thread's 4 cells in row 0 */
	uint32_t nb, nb1, nb2;    /* lanes 0 / 31: the neighbour cell in an H row, in an E1 (lane 0) or F1 (lane 31) row, in an E2 or F2 row */
	uint32_t bs1, bs2;        /* lanes 0 / 31: where this thread's outer gap cell goes -- lane 31: E row, cell 3; lane 0: F row, cell 0 */
	uint32_t f1off, f2off;    /* byte distance from an E1 row to the F1 row of the same score, E2 to F2 */
	bool left, right, edge_lane, bnd_lane; /* lane 0; lane 31; the neighbour cell exists; lane 0 or 31 */
	const uint32_t *seqw;     /* the words both sequences are probed in (packed codes, or the raw bytes) */
	uint32_t tbits, dq0;      /* bit offset of the target in seqw; (bit offset of the query) - tbits + d0 x code width */
	uint32_t cb, cb2, cb3;    /* code width in bits: 2, 4 or 8; twice, three times that */
	int lcb;                  /* log2(cb) */
	int d0, tl, ql, tlm1, qlm1d0; /* tl - 1; ql - 1 - d0 */
};

/* number of trailing zero bits; 0xffffffff for 0.  BREV + FLO run on the XU pipe, not on the integer ALU pipe that bounds this
 * kernel (written in C, the compiler turns every form of this into popc(~x & (x - 1)) plus a test for zero: four ALU instructions) */
__device__ __forceinline__ uint32_t ctz32_sat(uint32_t x)
{
	uint32_t r;
	asm("{ .reg .b32 t; brev.b32 t, %1; bfind.shiftamt.u32 %0, t; }" : "=r"(r) : "r"(x));
	return r;
}

template<int MODE>
__device__ __forceinline__ void tile_cells_fast(const FastCtx &c, const StepTab *tab, const SeqView &sv,
                                                int (&pe1)[4], int (&pf1)[4], int (&pe2)[4], int (&pf2)[4],
                                                bool full1, bool full2, uint64_t *stepbar, bool wait, uint32_t parity,
                                                int (&Hn)[4], uint32_t &tb_out)
{
	const int4 qh = tab->h, qe = tab->e;
	int ho1[4], ho2[4], hx[4];
	ldsv<4>(c.sb + qh.y, ho1); ldsv<4>(c.sb + qh.z, ho2); ldsv<4>(c.sb + qh.x, hx);
	int A1[6], A2[6], C1[6], C2[6], bA1[6], bA2[6], bC1[6], bC2[6];
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		A1[j + 1] = max(ho1[j], pe1[j]), A2[j + 1] = max(ho2[j], pe2[j]);
		C1[j + 1] = max(ho1[j], pf1[j]), C2[j + 1] = max(ho2[j], pf2[j]);
		if (MODE != MODE_SCORE) bA1[j + 1] = ho1[j] < pe1[j], bA2[j + 1] = ho2[j] < pe2[j], bC1[j + 1] = ho1[j] < pf1[j], bC2[j + 1] = ho2[j] < pf2[j];
	}
	A1[0] = __shfl_up_sync(0xffffffffu, A1[4], 1);
	A2[0] = __shfl_up_sync(0xffffffffu, A2[4], 1);
	C1[5] = __shfl_down_sync(0xffffffffu, C1[1], 1);
	C2[5] = __shfl_down_sync(0xffffffffu, C2[1], 1);
	if (MODE != MODE_SCORE) {
		const int bl = __shfl_up_sync(0xffffffffu, bA1[4] | bA2[4] << 1, 1);
		const int br = __shfl_down_sync(0xffffffffu, bC1[1] | bC2[1] << 1, 1);
		bA1[0] = bl & 1, bA2[0] = bl >> 1, bC1[5] = br & 1, bC2[5] = br >> 1;
	}
	int h0[4];
	uint32_t tz[4];
	bool ext[4];
	uint32_t tbw = 0;
	/* wf_next (:261-308) for cell j, then the first probe of its match run (:212-226): valid (:402) iff 0 <= k+1 <= tl and
	 * 0 <= d+k+1 <= ql; a cell off the matrix probes position 0 of T and d of Q, which is readable and never used */
#define FAST_CELL(j) do { \
		const int E1 = A1[j], E2 = A2[j], F1 = C1[j + 2] + 1, F2 = C2[j + 2] + 1; \
		const int e = max(E1, E2), f = max(F1, F2), gmx = max(e, f), hxp = hx[j] + 1; \
		const int H = max(hxp, gmx); \
		if (MODE != MODE_SCORE) { \
			const int z = hxp >= gmx ? 0 : (e >= f ? (E1 >= E2 ? 1 : 3) : (F1 >= F2 ? 2 : 4)); \
			tbw |= (uint32_t)(z | bA1[j] << 3 | bC1[j + 2] << 4 | bA2[j] << 5 | bC2[j + 2] << 6) << (8 * j); \
		} \
		pe1[j] = E1, pe2[j] = E2, pf1[j] = F1, pf2[j] = F2, h0[j] = H; \
		const int tp = H + 1, qp = tp + c.d0 + j; \
		ext[j] = (unsigned)tp <= (unsigned)c.tl && (unsigned)qp <= (unsigned)c.ql; \
		const uint32_t tpb = (uint32_t)(ext[j] ? tp : 0) * c.cb + c.tbits; /* bit positions in seqw */ \
		const uint32_t qpb = tpb + c.dq0 + (j == 0 ? 0u : j == 1 ? c.cb : j == 2 ? c.cb2 : c.cb3); \
		const uint32_t *tw = c.seqw + (tpb >> 5), *qw = c.seqw + (qpb >> 5); \
		tz[j] = ctz32_sat(__funnelshift_r(__ldg(tw), __ldg(tw + 1), tpb) ^ __funnelshift_r(__ldg(qw), __ldg(qw + 1), qpb)); \
	} while (0)
	FAST_CELL(1);
	FAST_CELL(2);
	if (wait) mbar_wait(stepbar, parity); /* every warp has finished the previous score */
	{
		const int o1 = lds1_if(c.nb + qh.y, c.edge_lane), o2 = lds1_if(c.nb + qh.z, c.edge_lane);
		const int x1 = lds1_if(c.nb1 + qe.x, c.edge_lane), x2 = lds1_if(c.nb2 + qe.y, c.edge_lane);
		const int m1 = max(o1, x1), m2 = max(o2, x2);
		if (c.left) A1[0] = m1, A2[0] = m2;
		if (c.right) C1[5] = m1, C2[5] = m2;
		if (MODE != MODE_SCORE) {
			if (c.left) bA1[0] = o1 < x1, bA2[0] = o2 < x2;
			if (c.right) bC1[5] = o1 < x1, bC2[5] = o2 < x2;
		}
	}
	FAST_CELL(0);
	FAST_CELL(3);
#undef FAST_CELL
	/* gap rows: whole in the last steps of the block (they go back to the state buffer), else only the cell another warp reads */
	if (full1) { stsv<4>(c.sb + qe.z, pe1); stsv<4>(c.sb + qe.z + c.f1off, pf1); }
	else {
		const int v = c.right ? pe1[3] : pf1[0];
		asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.b32 [%0], %1; }" :: "r"(c.bs1 + qe.z), "r"(v), "r"((int)c.bnd_lane) : "memory");
	}
	if (full2) { stsv<4>(c.sb + qe.w, pe2); stsv<4>(c.sb + qe.w + c.f2off, pf2); }
	else {
		const int v = c.right ? pe2[3] : pf2[0];
		asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.b32 [%0], %1; }" :: "r"(c.bs2 + qe.w), "r"(v), "r"((int)c.bnd_lane) : "memory");
	}
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int k = min(h0[j] + (int)(tz[j] >> c.lcb), min(c.tlm1, c.qlm1d0 - j)); /* clamped to the matrix, as the reference's sentinels do */
		Hn[j] = ext[j] ? k : h0[j];
	}
	if ((tz[0] | tz[1] | tz[2] | tz[3]) & 32) { /* rare: some probe matched in all positions (its tz is 0xffffffff, its k above is void) */
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int kmax = min(c.tlm1, c.qlm1d0 - j);
			if (ext[j] && tz[j] == 0xffffffffu) {
				const int k = min(h0[j] + (32 >> c.lcb), kmax);
				Hn[j] = k < kmax ? tile_extend_more(sv, k, c.d0 + j, kmax) : k;
			}
		}
	}
	stsv<4>(c.sb + qh.w, Hn);
	tb_out = tbw;
}

__device__ __forceinline__ bool on_matrix_u(int d, int k, int tl, int ql)
{
	return (unsigned)(k + 1) <= (unsigned)tl && (unsigned)(d + k + 1) <= (unsigned)ql;
}

template<int CPT>
__device__ __forceinline__ int alive_cells(int d0, int tl, int ql, const CellOut<CPT> &o)
{
	int bits = 0;
#pragma unroll
	for (int j = 0; j < CPT; ++j) {
		const int d = d0 + j;
		if (on_matrix_u(d, o.H[j], tl, ql) || on_matrix_u(d, o.E1[j], tl, ql) || on_matrix_u(d, o.F1[j], tl, ql) ||
		    on_matrix_u(d, o.E2[j], tl, ql) || on_matrix_u(d, o.F2[j], tl, ql)) bits |= 1 << j;
	}
	return bits;
}

__device__ __forceinline__ int alive_cells4(int d0, int tl, int ql, const int (&H)[4], const int (&E1)[4], const int (&F1)[4], const int (&E2)[4], const int (&F2)[4])
{
	int bits = 0;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int d = d0 + j;
		if (on_matrix_u(d, H[j], tl, ql) || on_matrix_u(d, E1[j], tl, ql) || on_matrix_u(d, F1[j], tl, ql) ||
		    on_matrix_u(d, E2[j], tl, ql) || on_matrix_u(d, F2[j], tl, ql)) bits |= 1 << j;
	}
	return bits;
}

template<int CPT> __device__ __forceinline__ void store_tb(uint8_t *p, uint32_t w) /* CPT traceback bytes, streaming store */
{
	if (CPT == 4) __stcs(reinterpret_cast<uint32_t*>(p), w);
	else if (CPT == 2) __stcs(reinterpret_cast<unsigned short*>(p), (unsigned short)w);
	else __stcs(reinterpret_cast<unsigned char*>(p), (unsigned char)w);
}

/* Tb steps of an interior tile on the fast path; E1 / E2 = the gap-extension penalties (1 or 2) = how many steps back the gap
 * rows are read.  Returns the alive bits of this thread's cells (wf_stripe_shrink's input). */
template<int MODE, int E1, int E2>
__device__ __forceinline__ int tile_fast_block(const FastCtx &c, const StepTab *tab, const SeqView &sv, int d0, int Tb, int t_alive, int tl, int ql,
                                               bool useful, uint8_t *tbp, long long tb_pitch, uint64_t *stepbar, uint32_t &step_phase)
{
	const int lane = threadIdx.x & 31;
	int e1a[4], f1a[4], e1b[4], f1b[4], e2a[4], f2a[4], e2b[4], f2b[4], Hn[4];
	int alive_bits = 0;
	uint32_t tbw;
	/* the gap rows the first steps read: scores s0-e+1 .. s0, out of the loaded state */
	ldsv<4>(c.sb + tab[0].e.x, e1a); ldsv<4>(c.sb + tab[0].e.x + c.f1off, f1a);
	ldsv<4>(c.sb + tab[0].e.y, e2a); ldsv<4>(c.sb + tab[0].e.y + c.f2off, f2a);
	if (E1 == 2) { ldsv<4>(c.sb + tab[1].e.x, e1b); ldsv<4>(c.sb + tab[1].e.x + c.f1off, f1b); }
	if (E2 == 2) { ldsv<4>(c.sb + tab[1].e.y, e2b); ldsv<4>(c.sb + tab[1].e.y + c.f2off, f2b); }
#define FAST_STEP(X1, Y1, X2, Y2) do { \
		tile_cells_fast<MODE>(c, tab + (t - 1), sv, X1, Y1, X2, Y2, t > Tb - E1, t > Tb - E2, stepbar, t > 1, step_phase & 1, Hn, tbw); \
		if (t > 1) ++step_phase; \
		if (MODE == MODE_TB) { if (useful) store_tb<4>(tbp, tbw); tbp += tb_pitch; } \
		if (t > t_alive) alive_bits |= alive_cells4(d0, tl, ql, Hn, X1, Y1, X2, Y2); \
		if (t < Tb) { __syncwarp(); if (lane == 0) step_arrive(stepbar); } \
	} while (0)
	for (int t = 1;;) {
		FAST_STEP(e1a, f1a, e2a, f2a);
		if (++t > Tb) break;
		FAST_STEP((E1 == 2 ? e1b : e1a), (E1 == 2 ? f1b : f1a), (E2 == 2 ? e2b : e2a), (E2 == 2 ? f2b : f2a));
		if (++t > Tb) break;
	}
#undef FAST_STEP
	return alive_bits;
}

/*
 * The same step for pairs on two-bit codes with e1 = 2, e2 = 1 (the default penalties on DNA: the batch workloads), with the
 * selects of the probe replaced by clamps:
 *   - the probe position of the target, in bits of the packed buffer, is umin(2 H + c1, position tl): a cell at NEG_INF (2 H wraps to
 *     a huge unsigned value) or past the end reads position tl -- readable, never used;
 *   - twice the bases left to the end of the cell's diagonal, lim = kend_j - position, is <= 0 for exactly the cells the reference
 *     skips (miniwfa.c:402: NEG_INF, k >= tl, d + k >= ql; a cell before the start of its diagonal cannot arise),
 *     so the run length is min.relu(first differing bit, lim) >> 1: one instruction clamps to the matrix and zeroes the skipped cells;
 *   - the query is read at (target position) + 2 d + const: for skipped cells that is anywhere in [start of T, end of Q + tl], inside
 *     the sequence buffer (mwf_b200_batch_create leaves max_len of slack at its end).
 * The gap rows of the last scores go back to shared memory once, after the loop, instead of under a test in every step.
 */
template<int CPT> struct Fast2Ctx {
	uint32_t sb;              /* shared address of this thread's cells in row 0 */
	uint32_t nbh;             /* lanes 0 / 31: the neighbour's cell in row 0 of the H ring */
	uint32_t xr, xw;          /* lanes 0 / 31: exchange records read / written in steps of odd t (the other buffer is XCH_BUF bytes on) */
	bool left, right, edge_lane, bnd_lane;
	const uint2 *seqw;        /* the packed buffer, as overlapping pairs of words */
	uint32_t c1, tend, dq0;   /* bit position of T[H + 1] = cb H + c1 (cb = bits per code); of T[tl]; (bit position of Q[0]) - (of T[0]) + cb d0 */
	int kend[CPT];            /* cb kmax_j + c1 */
	int d0;
};

/* Gap cells across warp boundaries.  The d-1 neighbour of a warp's first cell and the d+1 neighbour of its last belong to other
 * warps, whose gap rows live in registers.  Lane 31 (E side) and lane 0 (F side) of every warp therefore publish, in every step,
 * the two gap values their neighbour needs in the NEXT step -- E1 / F1 of the previous score (e1 = 2: the set not written in this
 * step) and E2 / F2 of this score (e2 = 1) -- as one 8-byte record, double-buffered by the parity of the step; the reader gets
 * both with one load.  Records: [buffer][warp + 1][side], side 0 = what lane 0 of `warp` reads, side 1 = what its lane 31 reads. */
#define TILE_MAX_WARPS 16
#define XCH_BUF ((TILE_MAX_WARPS + 2) * 16)

/* LCB = log2(bits per code): 1 for two-bit codes (DNA), 2 for four-bit codes (DNA with N, soft-masked, IUPAC).  CPT = 4 is the
 * throughput geometry (cells 1 and 2 of a thread need nothing from other warps and go before the step barrier's wait); CPT = 2
 * and 1 put more warps on a tile, for single large pairs. */
template<int MODE, int CPT, int LCB, bool EDGE>
__device__ __forceinline__ void tile_cells_fast2(const Fast2Ctx<CPT> &c, const int4 qh, const uint32_t xoff_r, const uint32_t xoff_w, const SeqView &sv,
                                                 int (&pe1)[CPT], int (&pf1)[CPT], const int (&oe1)[CPT], const int (&of1)[CPT], int (&pe2)[CPT], int (&pf2)[CPT],
                                                 uint64_t *stepbar, bool wait, uint32_t parity, int (&Hn)[CPT], uint32_t &tb_out,
                                                 const int jfin, const int tlm1, uint32_t &inval_bits, int &done_z)
{
	constexpr uint32_t CB = 1u << LCB;
	int ho1[CPT], ho2[CPT], hx[CPT];
	ldsv<CPT>(c.sb + qh.y, ho1); ldsv<CPT>(c.sb + qh.z, ho2); ldsv<CPT>(c.sb + qh.x, hx);
	int A1[CPT + 2], A2[CPT + 2], C1[CPT + 2], C2[CPT + 2];
	/* traceback bits (wf_next_tb, miniwfa.c:290-298): "the extension beat the opening", one bit per gap state of the cell a gap
	 * value comes FROM.  cw = the four bits of this thread's own cells, already at the positions they take in the traceback byte
	 * of the neighbouring diagonal -- byte j of cw: bit 3 = (ho1 < pe1), bit 5 = (ho2 < pe2) of cell j (read by cell j + 1 as its
	 * E1 / E2 bits), bit 4 = (ho1 < pf1), bit 6 = (ho2 < pf2) (read by cell j - 1 as its F1 / F2 bits).  Each bit is the sign of
	 * a difference, shifted in with one funnel shift (no compare + select pair per bit). */
	uint32_t cw = 0;
#pragma unroll
	for (int j = 0; j < CPT; ++j) {
		A1[j + 1] = max(ho1[j], pe1[j]), A2[j + 1] = max(ho2[j], pe2[j]);
		C1[j + 1] = max(ho1[j], pf1[j]), C2[j + 1] = max(ho2[j], pf2[j]);
	}
	if (MODE != MODE_SCORE) {
#pragma unroll
		for (int j = CPT - 1; j >= 0; --j) { /* most significant byte first; every difference fits 32 bits (|values| <= 2^30 + small) */
			if (j != CPT - 1) cw <<= 4; /* bits 2..0 of byte j + 1 and bit 7 of byte j stay 0 */
			cw = __funnelshift_l((uint32_t)(ho2[j] - pf2[j]), cw, 1);
			cw = __funnelshift_l((uint32_t)(ho2[j] - pe2[j]), cw, 1);
			cw = __funnelshift_l((uint32_t)(ho1[j] - pf1[j]), cw, 1);
			cw = __funnelshift_l((uint32_t)(ho1[j] - pe1[j]), cw, 1);
		}
		cw <<= 3;
	}
	A1[0] = __shfl_up_sync(0xffffffffu, A1[CPT], 1);
	A2[0] = __shfl_up_sync(0xffffffffu, A2[CPT], 1);
	C1[CPT + 1] = __shfl_down_sync(0xffffffffu, C1[1], 1);
	C2[CPT + 1] = __shfl_down_sync(0xffffffffu, C2[1], 1);
	uint32_t cl = 0, cr = 0; /* cw of the threads owning diagonals d0 - 1 (its last byte matters) and d0 + CPT (its first byte) */
	if (MODE != MODE_SCORE) {
		cl = __shfl_up_sync(0xffffffffu, cw, 1);
		cr = __shfl_down_sync(0xffffffffu, cw, 1);
	}
	int h0[CPT], lim[CPT];
	uint32_t tz[CPT];
	uint32_t tbw = 0;
#define FAST2_CELL(j) do { \
		const int E1 = A1[j], E2 = A2[j], F1 = C1[j + 2] + 1, F2 = C2[j + 2] + 1; \
		const int e = max(E1, E2), f = max(F1, F2), gmx = max(e, f), hxp = hx[j] + 1; \
		const int H = max(hxp, gmx); \
		if (MODE != MODE_SCORE) { \
			const int z = hxp >= gmx ? 0 : (e >= f ? (E1 >= E2 ? 1 : 3) : (F1 >= F2 ? 2 : 4)); \
			tbw |= (uint32_t)z << (8 * (j)); \
		} \
		pe1[j] = E1, pe2[j] = E2, pf1[j] = F1, pf2[j] = F2, h0[j] = H; \
		const uint32_t tpb = min(((uint32_t)H << LCB) + c.c1, c.tend); \
		const uint32_t qpb = tpb + c.dq0 + CB * j; \
		const uint2 tv = __ldg(c.seqw + (tpb >> 5)), qv = __ldg(c.seqw + (qpb >> 5)); \
		tz[j] = ctz32_sat(__funnelshift_r(tv.x, tv.y, tpb) ^ __funnelshift_r(qv.x, qv.y, qpb)); \
		lim[j] = c.kend[j] - (int)tpb; \
		if (EDGE && LCB > 1 && H < -1) lim[j] = 0; /* (cells outside the band: 4 H wraps to a position inside T; interior tiles hold no such cell) */ \
	} while (0)
	if (CPT == 4) { FAST2_CELL(1); FAST2_CELL(2); }
	if (wait) mbar_wait(stepbar, parity); /* every warp has finished the previous score */
	{
		const int o1 = lds1_if(c.nbh + qh.y, c.edge_lane), o2 = lds1_if(c.nbh + qh.z, c.edge_lane);
		int x1 = NEG_INF, x2 = NEG_INF;
		asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q ld.shared.v2.b32 {%0,%1}, [%2]; }" : "+r"(x1), "+r"(x2) : "r"(c.xr + xoff_r), "r"((int)c.edge_lane) : "memory");
		const int m1 = max(o1, x1), m2 = max(o2, x2);
		if (c.left) A1[0] = m1, A2[0] = m2;
		if (c.right) C1[CPT + 1] = m1, C2[CPT + 1] = m2;
		if (MODE != MODE_SCORE) { /* the neighbour warp's bits: E1 / E2 at bits 3 / 5 of the last byte, F1 / F2 at bits 4 / 6 of the first */
			const uint32_t v = ((uint32_t)(o1 - x1) >> 31) | ((uint32_t)(o2 - x2) >> 31) << 2;
			if (c.left) cl = v << (3 + 8 * (CPT - 1));
			if (c.right) cr = v << 4;
		}
	}
	if (MODE != MODE_SCORE) /* E bits from the diagonal below (bytes shifted up by one), F bits from the diagonal above */
		tbw |= (((cw << 8) | ((cl >> (8 * (CPT - 1))) & 0xffu)) & 0x28282828u) | (((cw >> 8) | ((cr & 0xffu) << (8 * (CPT - 1)))) & 0x50505050u);
	if (CPT == 4) { FAST2_CELL(0); FAST2_CELL((CPT - 1)); }
	else {
#pragma unroll
		for (int j = 0; j < CPT; ++j) FAST2_CELL(j);
	}
#undef FAST2_CELL
	{ /* what the neighbour warp needs in the next step */
		const int v1 = c.right ? oe1[CPT - 1] : of1[0], v2 = c.right ? pe2[CPT - 1] : pf2[0];
		asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q st.shared.v2.b32 [%0], {%1,%2}; }" :: "r"(c.xw + xoff_w), "r"(v1), "r"(v2), "r"((int)c.bnd_lane) : "memory");
	}
	uint32_t any = 0;
#pragma unroll
	for (int j = 0; j < CPT; ++j) {
		int m;
		asm("min.relu.s32 %0, %1, %2;" : "=r"(m) : "r"((int)tz[j]), "r"(lim[j])); /* tz = 0xffffffff (all codes of the word equal) gives 0 here */
		Hn[j] = h0[j] + (m >> LCB);
		any |= tz[j];
	}
	if (any & 32) { /* rare: some probe matched in all its positions */
#pragma unroll
		for (int j = 0; j < CPT; ++j)
			if (tz[j] == 0xffffffffu && lim[j] > 0) {
				const int kmax = h0[j] + (lim[j] >> LCB), k = min(h0[j] + (32 >> LCB), kmax);
				Hn[j] = k < kmax ? tile_extend_more(sv, k, c.d0 + j, kmax) : k;
			}
	}
	stsv<CPT>(c.sb + qh.w, Hn);
	tb_out = tbw;
	if (EDGE) { /* tiles at the band's edges / holding the terminal diagonal (tile_fast2_block) */
		uint32_t iv = 0; /* bit j: cell j holds no value (H is the largest of the five, so H < -1 says it for all of them: the edge rule, :325-326) */
#pragma unroll
		for (int j = CPT - 1; j >= 0; --j) iv = __funnelshift_l((uint32_t)(Hn[j] + 1), iv, 1);
		inval_bits = iv;
		done_z = -1;
		if (jfin >= 0) { /* the end of both sequences (:405-409); the state that got there when no match run was added (wf_traceback's start) */
#pragma unroll
			for (int j = 0; j < CPT; ++j)
				if (j == jfin && Hn[j] == tlm1) done_z = (MODE != MODE_SCORE && Hn[j] == h0[j]) ? (int)(tbw >> (8 * j) & 7u) : 0;
		}
	}
}

/* the cells of a thread that wf_stripe_shrink keeps: any of the five values on the matrix.  H is the largest of the five, so the
 * other four are looked at only where H itself is off the matrix */
template<int CPT>
__device__ __forceinline__ int alive_cells_h(int d0, int tl, int ql, const int (&H)[CPT], const int (&E1)[CPT], const int (&F1)[CPT], const int (&E2)[CPT], const int (&F2)[CPT])
{
	int bits = 0;
#pragma unroll
	for (int j = 0; j < CPT; ++j)
		if (on_matrix_u(d0 + j, H[j], tl, ql)) bits |= 1 << j;
	if (bits != (1 << CPT) - 1) {
#pragma unroll
		for (int j = 0; j < CPT; ++j) {
			const int d = d0 + j;
			if (on_matrix_u(d, E1[j], tl, ql) || on_matrix_u(d, F1[j], tl, ql) || on_matrix_u(d, E2[j], tl, ql) || on_matrix_u(d, F2[j], tl, ql)) bits |= 1 << j;
		}
	}
	return bits;
}

/* EDGE = true: the same steps for a tile at an edge of the band and / or holding the terminal diagonal.  The caller has set every
 * cell outside the band [wflo0, wfhi0] of the block's first score to NEG_INF in all rows of the tile (what the reference's pads hold),
 * and no diagonal of the tile lies outside [-tl, ql].  Then the recurrence itself grows the band the way the reference does: the only
 * cell outside the band that can get a value in a step is the one next to it (its sources are cells of the band), and the reference
 * takes exactly that cell in when it holds a value (:325-326, :417-418).  A cell that gets its first value says so with one bit per step
 * and side in sc[12..15]; the terminal diagonal reports the first step in which it reaches the end of both sequences. */
template<int MODE, int CPT, int LCB, bool EDGE>
__device__ __forceinline__ int tile_fast2_block(const Fast2Ctx<CPT> &c, const StepTab *tab, const SeqView &sv, uint32_t f1off, uint32_t f2off, int d0, int Tb, int t_alive,
                                                int tl, int ql, bool useful, uint8_t *tbp, long long tb_pitch, uint64_t *stepbar, uint32_t &step_phase,
                                                const uint32_t outl, const uint32_t outr, const int jfin, int *sc, int &done_t, int &done_last)
{
	const int lane = threadIdx.x & 31;
	int e1a[CPT], f1a[CPT], e1b[CPT], f1b[CPT], e2a[CPT], f2a[CPT], Hn[CPT];
	int alive_bits = 0;
	uint32_t tbw, ever = 0, iv = 0;
	int dz = -1;
	ldsv<CPT>(c.sb + tab[0].e.x, e1a); ldsv<CPT>(c.sb + tab[0].e.x + f1off, f1a); /* E1 / F1 of scores s0 - 1 and s0 */
	ldsv<CPT>(c.sb + tab[1].e.x, e1b); ldsv<CPT>(c.sb + tab[1].e.x + f1off, f1b);
	ldsv<CPT>(c.sb + tab[0].e.y, e2a); ldsv<CPT>(c.sb + tab[0].e.y + f2off, f2a); /* E2 / F2 of score s0 */
	{ /* what the neighbour warps need in step 1: E1 / F1 of score s0 - 1, E2 / F2 of score s0 */
		const int v1 = c.right ? e1a[CPT - 1] : f1a[0], v2 = c.right ? e2a[CPT - 1] : f2a[0];
		asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q st.shared.v2.b32 [%0], {%1,%2}; }" :: "r"(c.xw + XCH_BUF), "r"(v1), "r"(v2), "r"((int)c.bnd_lane) : "memory");
	}
	__syncthreads();
#define FAST2_STEP(X1, Y1, O1, P1, XR, XW) do { \
		tile_cells_fast2<MODE, CPT, LCB, EDGE>(c, tab[t - 1].h, XR, XW, sv, X1, Y1, O1, P1, e2a, f2a, stepbar, t > 1, step_phase & 1, Hn, tbw, jfin, tl - 1, iv, dz); \
		if (t > 1) ++step_phase; \
		if (MODE == MODE_TB) { if (useful) store_tb<CPT>(tbp, tbw); tbp += tb_pitch; } \
		if (t > t_alive) alive_bits |= alive_cells_h<CPT>(d0, tl, ql, Hn, X1, Y1, e2a, f2a); \
		if (EDGE) { \
			const uint32_t nw = ~iv & (outl | outr) & ~ever; \
			if (nw) { \
				ever |= nw; \
				if (nw & outl) atomicOr(&sc[12 + ((t - 1) >> 5)], 1 << ((t - 1) & 31)); \
				if (nw & outr) atomicOr(&sc[14 + ((t - 1) >> 5)], 1 << ((t - 1) & 31)); \
			} \
			if (dz >= 0 && done_t == 0x7fffffff) done_t = t, done_last = dz; \
		} \
		if (t < Tb) { __syncwarp(); if (lane == 0) step_arrive(stepbar); } \
	} while (0)
	int t = 1;
	for (;;) { /* odd t: reads the records of buffer 1, writes buffer 0; even t: the other way round */
		FAST2_STEP(e1a, f1a, e1b, f1b, XCH_BUF, 0);
		if (++t > Tb) break;
		FAST2_STEP(e1b, f1b, e1a, f1a, 0, XCH_BUF);
		if (++t > Tb) break;
	}
#undef FAST2_STEP
	/* the gap rows the next block loads: E1 / F1 of the last two scores, E2 / F2 of the last one */
	const bool odd = Tb & 1; /* the last step wrote set a */
	if (odd) {
		stsv<CPT>(c.sb + tab[Tb - 1].e.z, e1a); stsv<CPT>(c.sb + tab[Tb - 1].e.z + f1off, f1a);
		if (Tb > 1) { stsv<CPT>(c.sb + tab[Tb - 2].e.z, e1b); stsv<CPT>(c.sb + tab[Tb - 2].e.z + f1off, f1b); }
	} else {
		stsv<CPT>(c.sb + tab[Tb - 1].e.z, e1b); stsv<CPT>(c.sb + tab[Tb - 1].e.z + f1off, f1b);
		stsv<CPT>(c.sb + tab[Tb - 2].e.z, e1a); stsv<CPT>(c.sb + tab[Tb - 2].e.z + f1off, f1a);
	}
	stsv<CPT>(c.sb + tab[Tb - 1].e.w, e2a); stsv<CPT>(c.sb + tab[Tb - 1].e.w + f2off, f2a);
	return alive_bits;
}

/* threads per CTA and CTAs per SM the kernel is compiled for: 4 cells per thread keeps the instruction count per cell lowest
 * (batches); 2 and 1 cells per thread put 2x / 4x the threads on a tile, which shortens the dependent chain of one score
 * step when there are too few tiles to fill the GPU (single large pairs) */
#define TILE_MAX_THREADS(CPT) ((CPT) == 2 ? 256 : 512)
#define TILE_MIN_CTAS(CPT) ((CPT) == 4 ? 1 : 2)

/* the persistent CTA state of the tile kernels: shared-memory layout and the phases of the two mbarriers */
struct TileSmem {
	int32_t *rows;        /* [R][W] */
	int *sc;              /* [0..2] flags, [3..7] item, [8..11] the two mbarriers, [12..13] ticket taken ahead */
	uint64_t *bar, *stepbar;
	StepTab *steptab;     /* [max(T, 2)] row offsets of the steps of the block in flight */
	uint32_t xch;         /* [2][TILE_MAX_WARPS + 2] records of 16 bytes: gap cells across warp boundaries */
	uint32_t sb;          /* shared address of this thread's cells in row 0 */
	uint32_t phase, step_phase;
};

#ifdef MWF_PHASE_PROF /* development build only: clock cycles of thread 0 per phase of the persistent kernel, summed over CTAs */
__device__ unsigned long long g_phase[16];
#define PH_DECL __shared__ long long ph_acc[16]; __shared__ long long ph_last; long long *const ph_p = ph_acc, *const ph_l = &ph_last;
#define PH(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); ph_p[k] += now_ - *ph_l; *ph_l = now_; } } while (0)
#define PH_ARGS , long long *ph_p, long long *ph_l
#define PH_PASS , ph_p, ph_l
#else
#define PH(k) do {} while (0)
#define PH_ARGS
#define PH_PASS
#endif

template<int CPT>
__device__ __forceinline__ void tile_smem_setup(const TParams &P, int32_t *smem_tile, TileSmem &S)
{
	const int tid = threadIdx.x;
	S.rows = smem_tile;
	S.sc = S.rows + (size_t)P.R * P.W;
	S.bar = reinterpret_cast<uint64_t*>(S.sc + 8), S.stepbar = reinterpret_cast<uint64_t*>(S.sc + 10);
	S.steptab = reinterpret_cast<StepTab*>(S.sc + 16);
	S.xch = smem_u32(S.steptab + max(P.T, 2));
	S.sb = smem_u32(S.rows) + 4 * CPT * tid;
	S.phase = 0, S.step_phase = 0;
	if (tid == 0) { S.sc[12] = 0; mbar_init(S.bar, 1); mbar_init(S.stepbar, blockDim.x >> 5); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
}

/* One work item: tile `tile` of the block in flight of pair `slot` -- load, Tb fused next+extend steps, store.  On return the
 * bulk stores of the tile are committed (not yet complete); returns the number of tiles of the block. */
template<int MODE, int CPT>
__device__ __forceinline__ int tile_item(const TParams &P, TileSmem &S, const int slot, const int tile, bool &wrote_alive PH_ARGS)
{
	const int W = P.W, R = P.R, HL = P.HL, pitch = P.pitch;
	int32_t *rows = S.rows;
	int *sc = S.sc;
	uint64_t *bar = S.bar, *stepbar = S.stepbar;
	StepTab *steptab = S.steptab;
	const uint32_t xch = S.xch, sb = S.sb;
	uint32_t &phase = S.phase, &step_phase = S.step_phase;
	const int tid = threadIdx.x, lane = tid & 31, NT = blockDim.x;
	const int n = P.pen.nring, d1 = P.pen.e1 + 1, d2 = P.pen.e2 + 1;
	const bool no_left = tid == 0, no_right = tid == NT - 1;
	TileCtl *ctl = P.ctl + slot;
	const int pi = P.order[P.pair0 + pair_slot(P, slot)];
	const PairDesc pd = P.pairs[pi];
	const int tl = pd.tl, ql = pd.ql, doff = tile_doff(P, tl), dfin = ql - tl;
	/* (fields of TileCtl and the alive words are written by other SMs inside the persistent kernel: read from L2, never through L1) */
	const int Tb = __ldcg(&ctl->Tb), s0 = __ldcg(&ctl->s), n_tiles = __ldcg(&ctl->n_tiles), cur = __ldcg(&ctl->cur);
	const int total4 = __ldcg(&ctl->total4), A4 = __ldcg(&ctl->A4);
	const int u0 = (int)((long long)tile * total4 / n_tiles), u1 = (int)((long long)(tile + 1) * total4 / n_tiles);
	const int ustart = A4 + 4 * u0, ulen = 4 * (u1 - u0), idx0 = ustart - HL;
	const bool left_edge = tile == 0, right_edge = tile == n_tiles - 1;
	int wflo_c = __ldcg(&ctl->wflo), wfhi_c = __ldcg(&ctl->wfhi);
	int32_t *st_in = P.state + ((size_t)slot * 2 + cur) * R * pitch;
	int32_t *st_out = P.state + ((size_t)slot * 2 + (cur ^ 1)) * R * pitch;
	/* ---- load the tile: R rows of W int32, one bulk copy per row ---- */
	PH(11);
	if (tid < 32) {
		fence_async_smem();
		if (tid == 0) mbar_expect_tx(bar, (uint32_t)(R * W * 4));
		__syncwarp();
		for (int r = tid; r < R; r += 32)
			bulk_g2s(rows + (size_t)r * W, st_in + (size_t)r * pitch + idx0, (uint32_t)(W * 4), bar);
	}
	PH(12);
	if (tid < 3) sc[tid] = 0;
	if (P.fast && tid < max(Tb, 2)) { /* step tid+1 of the block works on score s0 + tid + 1 */
		const int s = s0 + tid + 1;
		const int4 h = P.tabH[s % n], a = P.tabE1[s % d1], b2 = P.tabE2[s % d2];
		StepTab e;
		e.h = h, e.e = make_int4(a.x, b2.x, a.z, b2.z);
		steptab[tid] = e;
	}
	SeqView sv;
	const int code_bits = P.packed ? P.packed[pi] : 0;
	if (code_bits == 2) /* two-bit codes, 16 per word */
		sv.T = P.seqp + (pd.t_off >> 3), sv.Q = P.seqp + (pd.q_off >> 3), sv.s_idx = 4, sv.s_amt = 1, sv.s_adv = 1;
	else if (code_bits == 4) /* four-bit codes, 8 per word */
		sv.T = P.seqp + (pd.t_off >> 3), sv.Q = P.seqp + (pd.q_off >> 3), sv.s_idx = 3, sv.s_amt = 2, sv.s_adv = 2;
	else sv.T = reinterpret_cast<const uint32_t*>(P.seq + pd.t_off), sv.Q = reinterpret_cast<const uint32_t*>(P.seq + pd.q_off), sv.s_idx = 2, sv.s_amt = 3, sv.s_adv = 3;
	const int c = CPT * tid, d0 = idx0 + c - doff;
	const bool useful = c >= HL && c < HL + ulen;
	const bool special = left_edge || right_edge || (dfin >= idx0 - doff && dfin < idx0 - doff + W); /* flags matter */
	uint8_t *tbp = 0; /* traceback bytes of this thread's diagonals at the block's first score (wf_tb_add, :33-44) */
	long long tb_pitch = 0;
	if (MODE == MODE_TB) tbp = P.arena + __ldcg(&ctl->row_base) + (idx0 + c), tb_pitch = __ldcg(&ctl->row_size);
	int kmin[CPT], kspan[CPT];
#pragma unroll
	for (int j = 0; j < CPT; ++j) { /* H is on the matrix iff kmin <= H <= kmin + kspan (:402) */
		const int d = d0 + j, lo = max(-1, -1 - d), hi = min(tl - 1, ql - 1 - d);
		if (hi >= lo) kmin[j] = lo, kspan[j] = hi - lo;
		else kmin[j] = 0x3fffffff, kspan[j] = 0;
	}
	const int wd_lo = d0 - CPT * lane, wd_hi = wd_lo + 32 * CPT - 1;
	const int bnd = (s0 | 0xff) + 1; /* next band trim */
	const int t_alive = bnd - n - s0; /* steps t > t_alive feed wf_stripe_shrink (:144-171) */
	int alive_bits = 0;
	int hs = s0 % n, e1s = s0 % d1, e2s = s0 % d2;
	PH(1);
	mbar_wait(bar, phase);
	phase ^= 1;
	__syncthreads();
	PH(2);
	const long long snap_off = __ldcg(&ctl->snap_off);
	if (snap_off >= 0) { /* snapshot for the segmented traceback: the state at score s0, useful columns of every row */
		if (tid < 32) {
			int32_t *dst = P.snap_arena + snap_off + (ustart - A4);
			const int rs = __ldcg(&ctl->snap_rowsize);
			fence_async_smem();
			for (int r = tid; r < R; r += 32) bulk_s2g(dst + (size_t)r * rs, rows + (size_t)r * W + HL, (uint32_t)(ulen * 4));
			bulk_commit();
			bulk_wait_read();
		}
		__syncthreads();
	}
	/* ---- Tb fused next+extend steps ---- */
	CellOut<CPT> o;
	bool stepped = false;
	const bool fast2_ok = Tb > 0 && P.fast >= 2 && code_bits != 0 && P.pen.e1 == 2 && P.pen.e2 == 1; /* packed codes, default gap extensions */
	const bool fast_edge = special && fast2_ok && P.fast_edge && __ldcg(&ctl->nat_ok) && idx0 - doff >= -tl && idx0 - doff + W - 1 <= ql; /* (no diagonal of the tile off the matrix) */
	if ((!special && fast2_ok) || fast_edge) { /* the register-resident step */
		Fast2Ctx<CPT> c;
		const uint32_t rb = 4u * W, cb = (uint32_t)code_bits;
		const int warp = tid >> 5;
		c.sb = sb, c.left = lane == 0, c.right = lane == 31, c.bnd_lane = lane == 0 || lane == 31;
		c.edge_lane = (lane == 0 && !no_left) || (lane == 31 && !no_right);
		c.nbh = sb + (lane == 0 ? -4 : 4 * CPT);
		c.xr = xch + 16u * (warp + 1) + (lane == 0 ? 0u : 8u);                    /* side 0: from lane 31 of the warp before; side 1: from lane 0 of the next */
		c.xw = xch + (lane == 0 ? 16u * warp + 8u : 16u * (warp + 2));           /* lane 0 writes side 1 of the warp before, lane 31 side 0 of the next */
		c.seqw = P.seqp2;
		const uint32_t tbits = (uint32_t)(sv.T - P.seqp) << 5, qbits = (uint32_t)(sv.Q - P.seqp) << 5;
		c.c1 = tbits + cb, c.tend = tbits + cb * (uint32_t)tl, c.dq0 = qbits - tbits + cb * (uint32_t)d0;
#pragma unroll
		for (int j = 0; j < CPT; ++j) c.kend[j] = (int)cb * min(tl - 1, ql - 1 - (d0 + j)) + (int)c.c1;
		c.d0 = d0;
		int done_t = 0x7fffffff, done_last = 0;
		if (!fast_edge) {
			if (code_bits == 2) alive_bits = tile_fast2_block<MODE, CPT, 1, false>(c, steptab, sv, d1 * rb, d2 * rb, d0, Tb, t_alive, tl, ql, useful, tbp, tb_pitch, stepbar, step_phase, 0u, 0u, -1, sc, done_t, done_last);
			else alive_bits = tile_fast2_block<MODE, CPT, 2, false>(c, steptab, sv, d1 * rb, d2 * rb, d0, Tb, t_alive, tl, ql, useful, tbp, tb_pitch, stepbar, step_phase, 0u, 0u, -1, sc, done_t, done_last);
			__syncthreads();
		} else {
			/* what lies outside the band of the block's first score reads as NEG_INF in every row (the reference's pads; stale or
			 * trimmed cells otherwise) */
			uint32_t outl = 0, outr = 0;
#pragma unroll
			for (int j = 0; j < CPT; ++j) {
				if (left_edge && d0 + j < wflo_c) outl |= 1u << j;
				if (right_edge && d0 + j > wfhi_c) outr |= 1u << j;
			}
			if (outl | outr) {
				const uint32_t out = outl | outr;
				for (int r = 0; r < R; ++r) {
#pragma unroll
					for (int j = 0; j < CPT; ++j)
						if (out >> j & 1) asm volatile("st.shared.b32 [%0], %1;" :: "r"(sb + (uint32_t)r * rb + 4u * j), "r"(NEG_INF) : "memory");
				}
			}
			if (tid < 4) sc[12 + tid] = 0;
			if (!useful) outl = outr = 0; /* (halo cells are some other tile's, or nobody's) */
			const int jfin = useful && dfin >= d0 && dfin < d0 + CPT ? dfin - d0 : -1;
			if (code_bits == 2) alive_bits = tile_fast2_block<MODE, CPT, 1, true>(c, steptab, sv, d1 * rb, d2 * rb, d0, Tb, t_alive, tl, ql, useful, tbp, tb_pitch, stepbar, step_phase, outl, outr, jfin, sc, done_t, done_last);
			else alive_bits = tile_fast2_block<MODE, CPT, 2, true>(c, steptab, sv, d1 * rb, d2 * rb, d0, Tb, t_alive, tl, ql, useful, tbp, tb_pitch, stepbar, step_phase, outl, outr, jfin, sc, done_t, done_last);
			__syncthreads();
			/* the slices' bounds, score by score, for the planner's replay (:417-418): bit t - 1 of a mask = the band took a cell in step t */
			const unsigned long long ml = (unsigned long long)(unsigned int)sc[12] | (unsigned long long)(unsigned int)sc[13] << 32;
			const unsigned long long mr = (unsigned long long)(unsigned int)sc[14] | (unsigned long long)(unsigned int)sc[15] << 32;
			if (tid < Tb) {
				const unsigned long long before = (1ull << tid) - 1;
				if (left_edge) ctl->lo_log[tid] = max(wflo_c - __popcll(ml & before) - 1, -tl);
				if (right_edge) ctl->hi_log[tid] = min(wfhi_c + __popcll(mr & before) + 1, ql);
			}
			wflo_c -= __popcll(ml), wfhi_c += __popcll(mr);
			if (done_t != 0x7fffffff && __ldcg(&ctl->done_t) == 0x7fffffff) { ctl->done_t = done_t; ctl->done_last = done_last; } /* (one thread holds the terminal diagonal) */
			__syncthreads(); /* sc[12..] is the caller's again */
		}
		stepped = true;
	}
	if constexpr (CPT == 4) {
		if (stepped) {
		} else if (!special && Tb > 0 && P.fast) { /* interior tile, throughput geometry, gap rows in registers */
			FastCtx c;
			const uint32_t rb = 4u * W;
			c.sb = sb, c.left = lane == 0, c.right = lane == 31, c.bnd_lane = lane == 0 || lane == 31;
			c.edge_lane = (lane == 0 && !no_left) || (lane == 31 && !no_right);
			c.f1off = d1 * rb, c.f2off = d2 * rb; /* rows: H [n], E1 [d1], F1 [d1], E2 [d2], F2 [d2] */
			c.nb = sb + (lane == 0 ? -4 : 16);
			c.nb1 = c.nb + (lane == 0 ? 0u : c.f1off), c.nb2 = c.nb + (lane == 0 ? 0u : c.f2off);
			c.bs1 = lane == 31 ? sb + 12 : sb + c.f1off, c.bs2 = lane == 31 ? sb + 12 : sb + c.f2off;
			c.lcb = sv.s_amt, c.cb = 1u << sv.s_amt, c.cb2 = 2u << sv.s_amt, c.cb3 = 3u << sv.s_amt;
			c.seqw = code_bits ? P.seqp : reinterpret_cast<const uint32_t*>(P.seq);
			c.tbits = (uint32_t)(sv.T - c.seqw) << 5;
			c.dq0 = ((uint32_t)(sv.Q - c.seqw) << 5) - c.tbits + ((uint32_t)d0 << sv.s_amt);
			c.d0 = d0;
			c.tl = tl, c.ql = ql, c.tlm1 = tl - 1, c.qlm1d0 = ql - 1 - d0;
			const int e1 = P.pen.e1, e2 = P.pen.e2;
			if (e1 == 2 && e2 == 1) alive_bits = tile_fast_block<MODE, 2, 1>(c, steptab, sv, d0, Tb, t_alive, tl, ql, useful, tbp, tb_pitch, stepbar, step_phase);
			else if (e1 == 2) alive_bits = tile_fast_block<MODE, 2, 2>(c, steptab, sv, d0, Tb, t_alive, tl, ql, useful, tbp, tb_pitch, stepbar, step_phase);
			else if (e2 == 1) alive_bits = tile_fast_block<MODE, 1, 1>(c, steptab, sv, d0, Tb, t_alive, tl, ql, useful, tbp, tb_pitch, stepbar, step_phase);
			else alive_bits = tile_fast_block<MODE, 1, 2>(c, steptab, sv, d0, Tb, t_alive, tl, ql, useful, tbp, tb_pitch, stepbar, step_phase);
			__syncthreads();
			stepped = true;
		} else if (!special) { /* interior tile, throughput geometry: split-phase step barrier */
			for (int t = 1; t <= Tb; ++t) {
				hs = hs + 1 == n ? 0 : hs + 1, e1s = e1s + 1 == d1 ? 0 : e1s + 1, e2s = e2s + 1 == d2 ? 0 : e2s + 1;
				const int4 qh = P.tabH[hs], q1 = P.tabE1[e1s], q2 = P.tabE2[e2s];
				tile_cells_overlap<MODE>(sb, qh, q1, q2, d0, kmin, kspan, sv, t > 1 && P.pen.e2 == 1, no_left, no_right, stepbar, t > 1, step_phase & 1, o);
				if (t > 1) ++step_phase;
				if (MODE == MODE_TB) { if (useful) store_tb<CPT>(tbp, o.tb); tbp += tb_pitch; }
				if (t > t_alive) alive_bits |= alive_cells<CPT>(d0, tl, ql, o);
				if (t < Tb) { __syncwarp(); if (lane == 0) step_arrive(stepbar); }
			}
			__syncthreads();
			stepped = true;
		}
	}
	if (stepped) {
	} else if (!special) {
		for (int t = 1; t <= Tb; ++t) {
			hs = hs + 1 == n ? 0 : hs + 1, e1s = e1s + 1 == d1 ? 0 : e1s + 1, e2s = e2s + 1 == d2 ? 0 : e2s + 1;
			const int4 qh = P.tabH[hs], q1 = P.tabE1[e1s], q2 = P.tabE2[e2s];
			tile_cells<MODE, false, CPT>(sb, qh, q1, q2, d0, 0, 0, dfin, tl, kmin, kspan, sv, no_left, no_right, useful, o);
			if (MODE == MODE_TB) { if (useful) store_tb<CPT>(tbp, o.tb); tbp += tb_pitch; }
			if (t > t_alive) alive_bits |= alive_cells<CPT>(d0, tl, ql, o);
			__syncthreads();
		}
	} else {
		for (int t = 1; t <= Tb; ++t) {
			hs = hs + 1 == n ? 0 : hs + 1, e1s = e1s + 1 == d1 ? 0 : e1s + 1, e2s = e2s + 1 == d2 ? 0 : e2s + 1;
			const int4 qh = P.tabH[hs], q1 = P.tabE1[e1s], q2 = P.tabE2[e2s];
			const int lo_t = left_edge ? max(wflo_c - 1, -tl) : -0x3fffffff;   /* :417-418 */
			const int hi_t = right_edge ? min(wfhi_c + 1, ql) : 0x3fffffff;
			const bool edge = wd_lo <= lo_t || wd_hi >= hi_t || (dfin >= wd_lo && dfin <= wd_hi);
			if (edge) {
				const int myfl = tile_cells<MODE, true, CPT>(sb, qh, q1, q2, d0, lo_t, hi_t, dfin, tl, kmin, kspan, sv, no_left, no_right, useful, o);
				if (myfl) atomicOr(&sc[t % 3], myfl);
			} else tile_cells<MODE, false, CPT>(sb, qh, q1, q2, d0, lo_t, hi_t, dfin, tl, kmin, kspan, sv, no_left, no_right, useful, o);
			if (MODE == MODE_TB) { if (useful) store_tb<CPT>(tbp, o.tb); tbp += tb_pitch; }
			if (t > t_alive) alive_bits |= alive_cells<CPT>(d0, tl, ql, o);
			if (tid == 0) {
				sc[(t + 1) % 3] = 0;
				if (left_edge) ctl->lo_log[t - 1] = lo_t;
				if (right_edge) ctl->hi_log[t - 1] = hi_t;
			}
			__syncthreads();
			const int fl = sc[t % 3];
			if (fl & FL_LO) wflo_c = lo_t;
			if (fl & FL_HI) wfhi_c = hi_t;
			if ((fl & FL_DONE) && tid == 0 && __ldcg(&ctl->done_t) == 0x7fffffff) { ctl->done_t = t; ctl->done_last = fl >> FL_LAST_SHIFT; }
		}
	}
#ifdef MWF_PHASE_PROF
	if (special) { PH(7); if (threadIdx.x == 0) ph_p[15] += 1; } else { PH(3); if (threadIdx.x == 0) ph_p[14] += 1; }
#endif
	/* ---- store the useful columns of every row into the other state buffer ---- */
	if (tid < 32) {
		fence_async_smem();
		PH(8);
		for (int r = tid; r < R; r += 32)
			bulk_s2g(st_out + (size_t)r * pitch + ustart, rows + (size_t)r * W + HL, (uint32_t)(ulen * 4));
		PH(9);
		bulk_commit();
		PH(10);
	}
	if (tid == 0) {
		if (left_edge) ctl->fin_lo = wflo_c;
		if (right_edge) ctl->fin_hi = wfhi_c;
	}
	wrote_alive = Tb > t_alive;
	if (Tb > t_alive && useful) { /* alive words: tag = score of the coming trim | alive bit */
		int32_t *ap = P.alive + (size_t)slot * pitch + idx0 + c;
#pragma unroll
		for (int j = 0; j < CPT; ++j) {
			const int a = __ldcg(ap + j);
			__stcg(ap + j, ((a & ~1) == bnd ? a : bnd) | (alive_bits >> j & 1));
		}
	}
	return n_tiles;
}

/* shared memory: rows[R][W] int32 | ctl ints [16] (flags, item, mbarriers) | step table | exchange records */
template<int MODE, int CPT>
__global__ void __launch_bounds__(TILE_MAX_THREADS(CPT), TILE_MIN_CTAS(CPT)) wfa_tile_kernel(const __grid_constant__ TParams P, int it)
{
	extern __shared__ __align__(128) int32_t smem_tile[];
	TileSmem S;
	tile_smem_setup<CPT>(P, smem_tile, S);
	const int tid = threadIdx.x;
	const unsigned int n_items = P.cnt[it & 1].n_items;
	for (;;) {
		__syncthreads();
		if (tid == 0) S.sc[3] = (int)atomicAdd(&P.cnt[it & 1].next, 1u);
		__syncthreads();
		const unsigned int item = (unsigned int)S.sc[3];
		if (item >= n_items) break;
		const int2 it2 = P.items[item];
		bool wrote_alive;
#ifdef MWF_PHASE_PROF
		PH_DECL
#endif
		tile_item<MODE, CPT>(P, S, it2.x, it2.y, wrote_alive PH_PASS);
		if (tid < 32) bulk_wait_read(); /* the rows may be overwritten by the next item's load */
	}
	if (tid < 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); /* stores complete before the CTA retires */
}

/*
 * The persistent form: ONE launch per pass (per tile geometry) instead of one plan + one tile launch per block of scores.
 * CTAs take tickets from a queue; the item of a ticket is a tile of some pair's block in flight, a pair to plan, or "retire".
 * The CTA that finishes the last tile of a block (a counter in the pair's TileCtl) runs the planner for that pair on the spot
 * -- replay, trim, cut, publish the next block's tiles -- so pairs advance independently: no launch boundary, no tail in which
 * most SMs wait for the last tiles of a launch, and a single pair pays one counter + one planner call per block instead of two
 * launches.  Ordering: a tile's bulk stores are complete (wait_group 0) and fenced before its CTA counts it done; the planner
 * publishes an item with a fence between payload and sequence word; the consumer fences after seeing the sequence word (which
 * also drops the SM's L1 lines) and crosses to the async proxy before its bulk loads.
 */
template<int MODE, int CPT>
__global__ void __launch_bounds__(TILE_MAX_THREADS(CPT), TILE_MIN_CTAS(CPT)) wfa_tile_persist_kernel(const __grid_constant__ TParams P)
{
	extern __shared__ __align__(128) int32_t smem_tile[];
	TileSmem S;
	tile_smem_setup<CPT>(P, smem_tile, S);
	const int tid = threadIdx.x;
#ifdef MWF_PHASE_PROF
	PH_DECL
	if (tid == 0) { for (int k = 0; k < 16; ++k) ph_acc[k] = 0; ph_last = clock64(); }
#endif
	for (;;) {
		__syncthreads();
		if (tid == 0) { /* (the ticket was taken while the previous tile's stores drained, when there was one) */
			const unsigned int ticket = S.sc[12] ? (unsigned int)S.sc[13] : atomicAdd(&P.pq->head, 1u);
			const int2 v = q_take(P, ticket);
			S.sc[3] = v.x, S.sc[4] = v.y, S.sc[5] = 0, S.sc[12] = 0;
		}
		__syncthreads();
		const int slot = S.sc[3], tile = S.sc[4];
		PH(0);
		if (slot < 0) break;
		bool plan = tile < 0;
		if (!plan) {
			asm volatile("fence.proxy.async;" ::: "memory"); /* the state rows were written through the async proxy of other SMs */
			bool wrote_alive;
			const int n_tiles = tile_item<MODE, CPT>(P, S, slot, tile, wrote_alive PH_PASS);
			PH(4);
			unsigned int next_ticket = 0;
			if (tid == 0) next_ticket = atomicAdd(&P.pq->head, 1u); /* its round trip overlaps the waits below; the item is awaited only after this CTA's duties (count, plan) */
			if (tid < 32) { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); asm volatile("fence.proxy.async;" ::: "memory"); }
			PH(13);
			if (wrote_alive) /* one block in eight: every thread has stored alive words.  A release on a scratch word orders them
			                  * (MEMBAR.ALL.GPU) without the L1 invalidation __threadfence() would add */
				asm volatile("red.release.gpu.global.add.s32 [%0], 0;" :: "l"(&P.pq->scratch) : "memory");
			__syncthreads();
			if (tid == 0) { /* release: the rows, the band logs and (through the barrier) what the other threads stored */
				int prev;
				asm volatile("atom.add.release.gpu.global.s32 %0, [%1], 1;" : "=r"(prev) : "l"(&P.ctl[slot].tiles_done) : "memory");
				S.sc[5] = prev == n_tiles - 1;
				S.sc[12] = 1, S.sc[13] = (int)next_ticket;
			}
			__syncthreads();
			plan = S.sc[5] != 0;
			PH(5);
		}
		if (plan) {
			__syncthreads();
			if (tid == 0) S.sc[6] = 0;
			plan_pair<true>(P, slot, 0, S.sc + 6);
			__syncthreads();
			if (S.sc[6]) { /* nothing in flight any more: one retire item per CTA of the grid */
				if (tid == 0) S.sc[7] = (int)atomicAdd(&P.pq->tail, gridDim.x);
				__syncthreads();
				for (unsigned int j = tid; j < gridDim.x; j += blockDim.x) q_publish(P, (unsigned int)S.sc[7] + j, -1, -1);
			}
			PH(6);
		}
	}
#ifdef MWF_PHASE_PROF
	if (tid == 0) for (int k = 0; k < 16; ++k) atomicAdd(&g_phase[k], (unsigned long long)ph_acc[k]);
#endif
}

/* start of a persistent pass (or of its continuation in the other tile geometry): one plan item per running pair */
__global__ void wfa_tile_persist_begin_kernel(const TParams P, int grid)
{
	__shared__ int n_run;
	if (threadIdx.x == 0) n_run = 0;
	__syncthreads();
	for (int slot = threadIdx.x; slot < P.n_pairs; slot += blockDim.x)
		if (P.ctl[slot].status == TS_RUN) q_publish(P, (unsigned int)atomicAdd(&n_run, 1), slot, -1);
	__syncthreads();
	if (threadIdx.x == 0) {
		P.pq->head = 0, P.pq->tail = (unsigned int)(n_run ? n_run : grid), P.pq->n_inflight = n_run, P.pq->stop_req = 0, P.pq->switch_to = P.geom_id, P.pq->n_start = n_run, P.pq->n_cut = 0;
		int tot = 0;
		for (int slot = 0; slot < P.n_pairs; ++slot) if (P.ctl[slot].status == TS_RUN) tot += P.ctl[slot].n_tiles;
		P.pq->total_tiles = tot;
	}
	if (n_run == 0) /* nothing to do: retire the grid at once */
		for (int j = threadIdx.x; j < grid; j += blockDim.x) q_publish(P, (unsigned int)j, -1, -1);
}

typedef void (*tile_kernel_fn)(const TParams, int);
typedef void (*tile_persist_fn)(const TParams);

static tile_kernel_fn tile_kernel_for(bool tb, int cpt)
{
	if (tb) return cpt == 4 ? wfa_tile_kernel<MODE_TB, 4> : cpt == 2 ? wfa_tile_kernel<MODE_TB, 2> : wfa_tile_kernel<MODE_TB, 1>;
	return cpt == 4 ? wfa_tile_kernel<MODE_SCORE, 4> : cpt == 2 ? wfa_tile_kernel<MODE_SCORE, 2> : wfa_tile_kernel<MODE_SCORE, 1>;
}

static tile_persist_fn tile_persist_for(bool tb, int cpt)
{
	if (tb) return cpt == 4 ? wfa_tile_persist_kernel<MODE_TB, 4> : cpt == 2 ? wfa_tile_persist_kernel<MODE_TB, 2> : wfa_tile_persist_kernel<MODE_TB, 1>;
	return cpt == 4 ? wfa_tile_persist_kernel<MODE_SCORE, 4> : cpt == 2 ? wfa_tile_persist_kernel<MODE_SCORE, 2> : wfa_tile_persist_kernel<MODE_SCORE, 1>;
}

/*
 * Checkpoints of low-memory mode without the provenance stripe.  The reference's pass 1 (mwf_wfa_seg, miniwfa.c:551-601)
 * carries, for every cell, the index of its ancestor in the last snapshot and chains the snapshots backwards
 * (wf_traceback_seg, :528-549): checkpoint k is the (score, diagonal) of the last cell of the optimal path computed at or
 * before snapshot score s_k = (k+1)*step - 1.  wf_next_seg replays exactly the choice bits of wf_next_tb (:502-523), so the
 * same cells are met by walking the traceback bytes of an unbanded high-memory pass backwards (wf_traceback's moves,
 * :343-366) and noting where the walk crosses each s_k -- down to score 0, through a leading gap as well.
 * One warp per pair; no CIGAR is produced here (pass 2 does that on its own, banded, bytes).
 */
__global__ void wfa_tile_checkpoint_kernel(const TParams P)
{
	const int slot = blockIdx.x, pi = P.order[P.pair0 + slot], lane = threadIdx.x & 31;
	const TileCtl *c = P.ctl + slot;
	if (c->status != TS_DONE) { if (lane == 0) P.n_seg[slot] = 0; return; }
	const PairDesc pd = P.pairs[pi];
	const uint8_t *T8 = P.seq + pd.t_off, *Q8 = P.seq + pd.q_off;
	const long long *rowtab = P.rowtab + (size_t)slot * P.rowtab_stride;
	const int doff = tile_doff(P, pd.tl), step = P.step;
	const Pen pen = P.pen;
	int *seg = P.seg + (size_t)slot * P.seg_stride;
	int i = pd.ql - 1, k = pd.tl - 1, row = c->s, last = c->last;
	const int n_seg = row / step; /* snapshots are taken at scores m*step - 1 < final score (:585) */
	int ks = n_seg - 1, snap_s = n_seg * step - 1;
	if (lane == 0) P.n_seg[slot] = n_seg;
	while (row > 0 && ks >= 0) {
		if (last == 0) { /* greedy backward matches (:335-341): no score, no diagonal change */
			for (;;) {
				const int ii = i - lane, kk = k - lane;
				const bool same = ii >= 0 && kk >= 0 && __ldg(Q8 + ii) == __ldg(T8 + kk);
				const unsigned m = __ballot_sync(0xffffffffu, !same);
				if (m) { const int cnt = __ffs(m) - 1; i -= cnt, k -= cnt; break; }
				i -= 32, k -= 32;
			}
		}
		const int x = __ldcg(P.arena + rowtab[row] + (i - k + doff));
		const int state = last == 0 ? (x & 7) : last;
		const int ext = state > 0 ? (x >> (state + 2)) & 1 : 0;
		if (state == 0) { --i, --k; row -= pen.x; }
		else if (state == 1) { --i; row -= ext ? pen.e1 : pen.oe1; }
		else if (state == 3) { --i; row -= ext ? pen.e2 : pen.oe2; }
		else if (state == 2) { --k; row -= ext ? pen.e1 : pen.oe1; }
		else { --k; row -= ext ? pen.e2 : pen.oe2; }
		last = (state > 0 && ext) ? state : 0;
		while (ks >= 0 && row <= snap_s) { /* the cell just reached is the newest one at or below this snapshot */
			if (lane == 0) seg[2 * ks] = row, seg[2 * ks + 1] = i - k;
			--ks, snap_s -= step;
		}
	}
}

/* wf_traceback (miniwfa.c:329-377) for the pairs of a wave: one warp per pair */
__global__ void wfa_tile_traceback_kernel(const TParams P)
{
	const int slot = blockIdx.x, pi = P.order[P.pair0 + slot];
	const TileCtl *c = P.ctl + slot;
	if (c->status != TS_DONE) return;
	const PairDesc pd = P.pairs[pi];
	Job J;
	J.tl = pd.tl, J.ql = pd.ql, J.doff = tile_doff(P, pd.tl);
	J.T8 = P.seq + pd.t_off, J.Q8 = P.seq + pd.q_off;
	J.arena = P.arena, J.arena_cap = P.arena_cap;
	J.rowtab = P.rowtab + (size_t)slot * P.rowtab_stride;
	int end_state[3];
	__shared__ long long rtw[TB_ROWWIN];
	__shared__ TbCone cone[2];
	__shared__ __align__(16) TbSeqWin sw;
	const int n_cigar = traceback_warp(J, P.pen, c->s, c->last, P.cigar + pd.cigar_off + pd.cigar_cap, end_state, rtw, cone, &sw, cone + 1);
	if (threadIdx.x == 0) {
		P.outs[pi].end_s = end_state[0], P.outs[pi].end_i = end_state[1], P.outs[pi].end_k = end_state[2];
		P.outs[pi].n_cigar = n_cigar;
		P.outs[pi].cigar_pos = pd.cigar_off + pd.cigar_cap - n_cigar;
	}
}

/* ------------------------------------------------------------------------------------------ */
/* segmented traceback: forward pass with snapshots, then segments recomputed from the end      */
/* ------------------------------------------------------------------------------------------ */

/* after the forward (score-only) pass: remember how every pair ended and where its traceback starts */
__global__ void wfa_tile_trace_begin_kernel(const TParams P)
{
	const int slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= P.n_pairs) return;
	const TileCtl *c = P.ctl + slot;
	const PairDesc pd = P.pairs[P.order[P.pair0 + slot]];
	TraceState t;
	t.fwd_status = c->status, t.s_final = c->s, t.i = pd.ql - 1, t.k = pd.tl - 1, t.row = c->s, t.last = 0;
	t.cur_op = -1, t.n_out = 0, t.cur_len = 0, t.pad = 0, t.n_iter = c->n_iter;
	P.trace[slot] = t;
	if (P.step > 0 && P.n_seg) P.n_seg[slot] = c->status == TS_DONE ? c->s / P.step : 0; /* snapshots at m*step - 1 < final score (:585) */
}

/* start segment j of every pair that has one: segment 0 starts at score 0 (wfa_tile_init_kernel has run), segment j > 0 at
 * snapshot j-1, whose rows go back into state buffer 0.  grid = (n_pairs, copy CTAs) */
__global__ void wfa_tile_segstart_kernel(const TParams P, int j_first)
{
	/* slot v of a group of segments: segment j_first - v / vmod of pair slot v % vmod (one segment per slot when vmod = 0) */
	const int slot = blockIdx.x, ps = pair_slot(P, slot), j = j_first - (P.vmod > 0 ? slot / P.vmod : 0), pi = P.order[P.pair0 + ps];
	TileCtl *c = P.ctl + slot;
	const TraceState *ts = P.trace + ps;
	const int n_snap = P.n_snap[ps];
	const bool active = ts->fwd_status == TS_DONE && j <= n_snap && j >= 0;
	if (j == 0) {
		if (blockIdx.y == 0 && threadIdx.x == 0) {
			if (!active && c->status == TS_RUN) { c->status = TS_IDLE; atomicSub(P.n_running, 1); }
			P.s_stop[slot] = n_snap > 0 ? P.snap_P : ts->s_final;
		}
		return;
	}
	if (!active) { if (blockIdx.y == 0 && threadIdx.x == 0) c->status = TS_IDLE; return; }
	const SnapDir d = P.snapdir[(size_t)ps * P.snapdir_stride + (j - 1)];
	int32_t *st = P.state + (size_t)slot * 2 * P.R * P.pitch;
	const int32_t *src = P.snap_arena + d.off;
	const int n4 = d.rowsize >> 2;
	for (int r = 0; r < P.R; ++r) {
		const int4 *s4 = reinterpret_cast<const int4*>(src + (size_t)r * d.rowsize);
		int4 *d4 = reinterpret_cast<int4*>(st + (size_t)r * P.pitch + d.A4);
		for (int x = blockIdx.y * blockDim.x + threadIdx.x; x < n4; x += gridDim.y * blockDim.x) d4[x] = s4[x];
	}
	if (blockIdx.y == 0 && threadIdx.x == 0) {
		c->s = d.s, c->wflo = d.wflo, c->wfhi = d.wfhi, c->cur = 0, c->last = 0, c->sid = 0, c->copied = 0, c->n_iter = d.n_iter, c->shrink_s = d.s, c->nat_ok = 0;
		c->Tb = 0, c->n_tiles = 0, c->done_t = 0x7fffffff, c->done_last = 0, c->snap_off = -1, c->tiles_done = 0;
		c->status = TS_RUN;
		atomicAdd(P.n_running, 1);
		P.s_stop[slot] = j == n_snap ? ts->s_final : (j + 1) * P.snap_P;
		(void)pi;
	}
}

/* wf_traceback (miniwfa.c:329-377) over the traceback bytes of segment j only: the walk stops when it needs a row at or
 * below the segment's first score and resumes there in the next (earlier) segment.  One warp per pair. */
__global__ void wfa_tile_trace_seg_kernel(const TParams P, int j, int vslot0)
{
	const int slot = blockIdx.x, pi = P.order[P.pair0 + slot], lane = threadIdx.x & 31;
	TraceState *tsp = P.trace + slot;
	const int n_snap = P.n_snap[slot];
	if (tsp->fwd_status != TS_DONE || j > n_snap) return;
	const TileCtl *c = P.ctl + vslot0 + slot; /* the (virtual) slot that recomputed segment j */
	const PairDesc pd = P.pairs[pi];
	const uint8_t *T8 = P.seq + pd.t_off, *Q8 = P.seq + pd.q_off;
	const long long *rowtab = P.rowtab + (size_t)slot * P.rowtab_stride;
	const int doff = tile_doff(P, pd.tl), s_lo = j * P.snap_P;
	const Pen pen = P.pen;
	TraceState t = *tsp;
	if (j == n_snap) t.last = c->last; /* the pass that reached the end knows the state the path ends in (:405-409) */
	int i = t.i, k = t.k, row = t.row, last = t.last, cur_op = t.cur_op, n_out = t.n_out;
	uint32_t cur_len = t.cur_len;
	uint32_t *wp = P.cigar + pd.cigar_off + pd.cigar_cap - n_out;
	__shared__ TbCone cone; /* the traceback bytes around the walk (wfa_engine.cu) */
	__shared__ __align__(16) TbSeqWin sw; /* and the sequences under the match runs */
	int qwb = 0x7fffffff, twb = 0x7fffffff;
	TbConePos cp;
	cp.top = -1, cp.lo = 0, cp.c0 = 0;
	auto rowtab_at = [&](int r_) -> long long { return rowtab[r_]; };
#define CIG_PUSH(op_, len_) do { \
		if ((op_) == cur_op) cur_len += (len_); \
		else { if (cur_op >= 0) { --wp; if (lane == 0) *wp = cur_len << 4 | (uint32_t)cur_op; ++n_out; } cur_op = (op_), cur_len = (len_); } \
	} while (0)
	while (i >= 0 && k >= 0) {
		if (last == 0) { /* greedy backward matches, :335-341 */
			int run = 0;
			for (;;) {
				const int ii = i - lane, kk = k - lane;
				if (qwb > max(0, i - 31)) qwb = tbseq_fill(sw.q, Q8, i);
				if (twb > max(0, k - 31)) twb = tbseq_fill(sw.t, T8, k);
				const bool same = ii >= 0 && kk >= 0 && sw.q[ii - qwb] == sw.t[kk - twb];
				const unsigned m = __ballot_sync(0xffffffffu, !same);
				if (m) { const int cnt = __ffs(m) - 1; run += cnt, i -= cnt, k -= cnt; break; }
				run += 32, i -= 32, k -= 32;
			}
			if (run > 0) CIG_PUSH(7, (uint32_t)run);
			if (i < 0 || k < 0) break;
		}
		if (row <= s_lo && j > 0) break; /* this row's bytes belong to an earlier segment */
		const long long col = (long long)(i - k + doff);
		if (!tbcone_has(cp, row, col)) tbcone_fill(&cone, cp, P.arena, P.arena_cap, rowtab_at, row, col);
		const int x = tbcone_get(&cone, cp, row, col);
		const int state = last == 0 ? (x & 7) : last;
		const int ext = state > 0 ? (x >> (state + 2)) & 1 : 0;
		if (state == 0) { CIG_PUSH(8, 1u); --i, --k; row -= pen.x; }
		else if (state == 1) { CIG_PUSH(1, 1u); --i; row -= ext ? pen.e1 : pen.oe1; }
		else if (state == 3) { CIG_PUSH(1, 1u); --i; row -= ext ? pen.e2 : pen.oe2; }
		else if (state == 2) { CIG_PUSH(2, 1u); --k; row -= ext ? pen.e1 : pen.oe1; }
		else { CIG_PUSH(2, 1u); --k; row -= ext ? pen.e2 : pen.oe2; }
		last = (state > 0 && ext) ? state : 0;
	}
	if (j == 0 || i < 0 || k < 0) { /* the walk is over: leading gap (:368-369), flush, result record */
		const int end_i = i, end_k = k;
		if (i >= 0) CIG_PUSH(1, (uint32_t)(i + 1));
		else if (k >= 0) CIG_PUSH(2, (uint32_t)(k + 1));
		if (cur_op >= 0) { --wp; if (lane == 0) *wp = cur_len << 4 | (uint32_t)cur_op; ++n_out; }
		if (lane == 0) {
			PairOut o;
			o.s = t.s_final, o.n_cigar = n_out, o.n_iter = t.n_iter, o.cigar_pos = pd.cigar_off + pd.cigar_cap - n_out, o.status = ST_OK, o.pad_ = 0;
			o.end_s = row, o.end_i = end_i, o.end_k = end_k, o.pad2_ = 0;
			P.outs[pi] = o;
			tsp->fwd_status = TS_IDLE; /* nothing left to do in the remaining segments */
		}
		return;
	}
#undef CIG_PUSH
	if (lane == 0) {
		t.i = i, t.k = k, t.row = row, t.last = last, t.cur_op = cur_op, t.n_out = n_out, t.cur_len = cur_len;
		*tsp = t;
	}
}


/* the checkpoint walk of wfa_tile_checkpoint_kernel over the traceback bytes of segment j only (low-memory requests whose
 * unbanded pass does not fit the arena as s^2 bytes): stops when it needs a row at or below the segment's first score */
__global__ void wfa_tile_ckpt_seg_kernel(const TParams P, int j)
{
	const int slot = blockIdx.x, pi = P.order[P.pair0 + slot], lane = threadIdx.x & 31;
	TraceState *tsp = P.trace + slot;
	const int n_snap = P.n_snap[slot];
	if (tsp->fwd_status != TS_DONE || j > n_snap) return;
	const TileCtl *c = P.ctl + slot;
	const PairDesc pd = P.pairs[pi];
	const uint8_t *T8 = P.seq + pd.t_off, *Q8 = P.seq + pd.q_off;
	const long long *rowtab = P.rowtab + (size_t)slot * P.rowtab_stride;
	const int doff = tile_doff(P, pd.tl), s_lo = j * P.snap_P, step = P.step;
	const Pen pen = P.pen;
	int *seg = P.seg + (size_t)slot * P.seg_stride;
	TraceState t = *tsp;
	if (j == n_snap) t.last = c->last;
	int i = t.i, k = t.k, row = t.row, last = t.last;
	int ks = row / step - 1, snap_s = (ks + 1) * step - 1; /* checkpoints at or above `row` are already known */
	while (row > s_lo && ks >= 0) {
		if (last == 0) {
			for (;;) {
				const int ii = i - lane, kk = k - lane;
				const bool same = ii >= 0 && kk >= 0 && __ldg(Q8 + ii) == __ldg(T8 + kk);
				const unsigned m = __ballot_sync(0xffffffffu, !same);
				if (m) { const int cnt = __ffs(m) - 1; i -= cnt, k -= cnt; break; }
				i -= 32, k -= 32;
			}
		}
		const int x = __ldcg(P.arena + rowtab[row] + (i - k + doff));
		const int state = last == 0 ? (x & 7) : last;
		const int ext = state > 0 ? (x >> (state + 2)) & 1 : 0;
		if (state == 0) { --i, --k; row -= pen.x; }
		else if (state == 1) { --i; row -= ext ? pen.e1 : pen.oe1; }
		else if (state == 3) { --i; row -= ext ? pen.e2 : pen.oe2; }
		else if (state == 2) { --k; row -= ext ? pen.e1 : pen.oe1; }
		else { --k; row -= ext ? pen.e2 : pen.oe2; }
		last = (state > 0 && ext) ? state : 0;
		while (ks >= 0 && row <= snap_s) {
			if (lane == 0) seg[2 * ks] = row, seg[2 * ks + 1] = i - k;
			--ks, snap_s -= step;
		}
	}
	if (lane == 0) {
		t.i = i, t.k = k, t.row = row, t.last = last;
		if (ks < 0) t.fwd_status = TS_IDLE; /* every checkpoint found: the earlier segments need no recompute */
		*tsp = t;
	}
}


/* Packing of the pairs whose two sequences together use at most 4 (DNA) or at most 16 (DNA with N, soft-masked, IUPAC) distinct
 * byte values: a presence bitmap of the pair, codes = rank of the byte among the values present (so equal codes <=> equal bytes),
 * 2 or 4 bits per code, 16 or 8 codes per 32-bit word, position p in bits c(p % (32/c)).  One CTA of 256 threads per pair.
 * Bytes past the end of a sequence become arbitrary codes; the match run is clamped to the matrix anyway. */
__global__ void __launch_bounds__(256) wfa_pack_kernel(const uint8_t *__restrict__ seq, const PairDesc *__restrict__ pairs, uint32_t *__restrict__ seqp, uint2 *__restrict__ seqp2, int *__restrict__ packed)
{
	__shared__ unsigned int bm[8];
	__shared__ unsigned char lut[256];
	__shared__ int sh_bits;
	const int pi = blockIdx.x, tid = threadIdx.x;
	const PairDesc pd = pairs[pi];
	if (tid < 8) bm[tid] = 0;
	__syncthreads();
	unsigned int loc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
	for (int which = 0; which < 2; ++which) {
		const uint8_t *p = seq + (which ? pd.q_off : pd.t_off);
		const int len = which ? pd.ql : pd.tl;
		for (long long i = (long long)tid * 16; i < len; i += 256 * 16) {
			const uint4 v = *reinterpret_cast<const uint4*>(p + i);
			const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
			for (int b = 0; b < 16; ++b)
				if (i + b < len) { const unsigned int c = w[b >> 2] >> (8 * (b & 3)) & 0xffu; loc[c >> 5] |= 1u << (c & 31); }
		}
	}
#pragma unroll
	for (int w = 0; w < 8; ++w) if (loc[w]) atomicOr(&bm[w], loc[w]);
	__syncthreads();
	if (tid == 0) {
		int cnt = 0;
		for (int w = 0; w < 8; ++w) cnt += __popc(bm[w]);
		sh_bits = cnt <= 4 ? 2 : cnt <= 16 ? 4 : 0;
		packed[pi] = sh_bits;
	}
	{
		int r = __popc(bm[tid >> 5] & ((1u << (tid & 31)) - 1u));
		for (int w = 0; w < (tid >> 5); ++w) r += __popc(bm[w]);
		lut[tid] = (unsigned char)(r & 15);
	}
	__syncthreads();
	const int bits = sh_bits;
	if (bits == 0) return;
	const int cpw = 32 / bits; /* codes per word */
	for (int which = 0; which < 2; ++which) {
		const long long off = which ? pd.q_off : pd.t_off;
		const uint8_t *p = seq + off;
		const int n_words = ((which ? pd.ql : pd.tl) + cpw - 1) / cpw + 5; /* the probe reads one word ahead, the long-run loop four */
		uint32_t *out = seqp + (off >> 3);
		for (int wi = tid; wi < n_words; wi += 256) {
			const uint8_t *src = p + (long long)wi * cpw;
			uint32_t code = 0;
			if (bits == 2) {
				const uint4 v = *reinterpret_cast<const uint4*>(src);
				const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
				for (int b = 0; b < 16; ++b) code |= (uint32_t)(lut[w[b >> 2] >> (8 * (b & 3)) & 0xffu] & 3) << (2 * b);
			} else {
				const uint2 v = *reinterpret_cast<const uint2*>(src);
				const uint32_t w[2] = { v.x, v.y };
#pragma unroll
				for (int b = 0; b < 8; ++b) code |= (uint32_t)lut[w[b >> 2] >> (8 * (b & 3)) & 0xffu] << (4 * b);
			}
			out[wi] = code;
		}
	}
	if (seqp2 == 0) return;
	__syncthreads(); /* the words above are visible to the whole CTA: now the overlapping pairs */
	for (int which = 0; which < 2; ++which) {
		const long long off = which ? pd.q_off : pd.t_off;
		const int n_words = ((which ? pd.ql : pd.tl) + cpw - 1) / cpw + 4;
		const uint32_t *in = seqp + (off >> 3);
		uint2 *out = seqp2 + (off >> 3);
		for (int wi = tid; wi < n_words; wi += 256) out[wi] = make_uint2(in[wi], in[wi + 1]);
	}
}

#endif
