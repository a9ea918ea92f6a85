/*
 * wfa_engine.cu -- sm_100a engine for the exact dual-affine wavefront alignment hot path.
 *
 * What runs here (reference: /root/reference/miniwfa.c @ 66770a3):
 *   step kernel   = wf_next_score / wf_next_tb (:261-308) fused with the extend loop and
 *                   wf_extend1_padded (:400-411, :212-226), the edge rule (:325-326), the
 *                   termination test (:405-409), wf_stripe_shrink (:144-171) and the stop
 *                   tests (:421-425); in low-memory pass 1 also wf_next_seg's replay (:503-523)
 *                   and wf_snapshot1 (:451-474);
 *   after the loop: wf_traceback_seg (:528-549) and wf_traceback (:329-377), on the device.
 *
 * Two kernel families share the same templated step:
 *   wfa_cta_kernel  : persistent, one CTA per pair taken from a work queue; the per-score
 *                     barrier is __syncthreads().  Used for batches.
 *   wfa_grid_kernel : cooperative launch, every SM works on the same pair; the per-score
 *                     barrier is a grid barrier.  Used for a few large pairs.
 *
 * Data layout in HBM (per slot = per resident CTA in cta mode, one slot in grid mode):
 *   ring  : int32 [nring][5][pitch], nring = max_pen+1 score slices x {H,E1,F1,E2,F2};
 *           cell of diagonal d lives at index d + doff; every slice is written together
 *           with nring cells of NEG_INF on both sides (same invariant as miniwfa.c:96-99),
 *           so the +-1 neighbour reads never need a range test.
 *   arena : traceback bytes, one row per score (bump-allocated, 1 byte per cell exactly as
 *           miniwfa.c:42) -- or, during low-memory pass 1, the snapshots.
 *   rowtab: int64 per score: arena offset of that score's row minus its first index.
 * Sequences: each padded to a 16-byte boundary with >=64 zero bytes of slack; match runs
 * are clamped to the matrix, so no sentinel characters are required.
 *
 * No CPU fallback exists: every entry point aborts if no CUDA device is usable.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <algorithm>
#include <vector>
#include <atomic>
#include <mutex>
#include <thread>
#include <unordered_map>
#include "mwf_b200.h"
#include "kalloc.h"

#define NEG_INF (-0x40000000) /* WF_NEG_INF, miniwfa.c:67 */

enum { MODE_SCORE = 0, MODE_TB = 1, MODE_SEG1 = 2 };
enum { ST_OK = 0, ST_STOPPED = 1, ST_ARENA = 2, ST_SHRINK = 3, ST_CORRUPT = 4 };
enum { FL_LO = 1, FL_HI = 2, FL_DONE = 4, FL_LAST_SHIFT = 4 };

#define CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	fprintf(stderr, "[mwf_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
	abort(); } } while (0)

struct Pen { int x, oe1, e1, oe2, e2, nring; };

struct PairDesc {
	long long t_off, q_off; /* byte offsets into the sequence buffer (16-byte aligned) */
	int tl, ql;
	long long cigar_off;    /* word offset of this pair's CIGAR buffer */
	int cigar_cap, pad_;
};

struct PairOut {
	int s, n_cigar;
	long long n_iter;
	long long cigar_pos;    /* word offset of the first CIGAR word */
	int status, pad_;
	int end_s, end_i, end_k, pad2_; /* where wf_traceback's walk stopped: what MWF_F_DEBUG prints (miniwfa.c:367) */
};

struct KParams {
	Pen pen;
	int is_tb, step, max_s;
	long long max_iter;
	int n_pairs, single_pair;
	const int *order;
	unsigned int *queue;
	const PairDesc *pairs;
	PairOut *outs;
	const uint8_t *seq;
	uint32_t *cigar;
	int pitch;
	int32_t *ring, *ring2;
	long long ring_stride;      /* int32 per slot */
	uint8_t *arena;
	long long arena_stride;     /* bytes per slot */
	long long *rowtab;
	long long rowtab_stride;    /* entries per slot */
	int *snaphdr;
	long long *snapoff;
	int *seg;
	int snap_cap;
	unsigned int *bar;          /* grid barrier: {count, generation} */
	int *gflags;                /* grid mode flag words [3] */
};

/* ------------------------------------------------------------------------------------------ */
/* small device helpers                                                                        */
/* ------------------------------------------------------------------------------------------ */

__device__ __forceinline__ int4 ld_ring4(const int32_t *p) { return __ldcg(reinterpret_cast<const int4*>(p)); }
__device__ __forceinline__ int  ld_ring1(const int32_t *p) { return __ldcg(p); }
__device__ __forceinline__ void st_ring4(int32_t *p, int4 v) { __stcg(reinterpret_cast<int4*>(p), v); }
__device__ __forceinline__ void st_ring1(int32_t *p, int v) { __stcg(p, v); }

/* 4 bytes starting at byte position pos (little endian), from a 4-byte aligned word array */
__device__ __forceinline__ uint32_t seq_word(const uint32_t *__restrict__ w, int pos)
{
	const int i = pos >> 2;
	return __funnelshift_r(__ldg(w + i), __ldg(w + i + 1), (pos & 3) << 3);
}

/* longest common prefix of ts[k+1..] and qs[d+k+1..], clamped so that k <= kmax = min(tl-1, ql-1-d).
 * Same value as wf_extend1_padded (miniwfa.c:212-226), whose sentinels stop at the same place. */
__device__ __forceinline__ int extend_run(const uint32_t *__restrict__ T, const uint32_t *__restrict__ Q, int k, int d, int kmax)
{
	int tp = k + 1, qp = d + k + 1;
	while (k < kmax) {
		const uint32_t x = seq_word(T, tp) ^ seq_word(Q, qp);
		if (x) { k += (__ffs(x) - 1) >> 3; break; }
		k += 4, tp += 4, qp += 4;
	}
	return min(k, kmax);
}

__device__ __forceinline__ bool on_matrix(int d, int k, int tl, int ql) /* good_diag, miniwfa.c:139-142 */
{
	return k >= -1 && k < tl && d + k >= -1 && d + k < ql;
}

/* ------------------------------------------------------------------------------------------ */
/* thread groups: the set of threads that cooperates on one pair                               */
/* ------------------------------------------------------------------------------------------ */

struct CtaGroup {
	int *fl; /* shared memory flag words [3] */
	__device__ __forceinline__ int rank() const { return threadIdx.x; }
	__device__ __forceinline__ int size() const { return blockDim.x; }
	__device__ __forceinline__ bool leader_cta() const { return true; }
	__device__ __forceinline__ void sync() { __syncthreads(); }
	__device__ __forceinline__ void flag_or(int k, int v) { atomicOr(&fl[k], v); }
	__device__ __forceinline__ int flag_get(int k) const { return ((volatile int*)fl)[k]; }
	__device__ __forceinline__ void flag_zero(int k) { if (threadIdx.x == 0) fl[k] = 0; }
};

struct GridGroup {
	int *fl;            /* global flag words [3] */
	unsigned int *bar;  /* {count, generation} */
	unsigned int gen;   /* generation this CTA waits to leave */
	__device__ __forceinline__ int rank() const { return blockIdx.x * blockDim.x + threadIdx.x; }
	__device__ __forceinline__ int size() const { return gridDim.x * blockDim.x; }
	__device__ __forceinline__ bool leader_cta() const { return blockIdx.x == 0; }
	__device__ __forceinline__ void sync()
	{
		__syncthreads();
		if (threadIdx.x == 0) {
			__threadfence();
			const unsigned int arrived = atomicAdd(&bar[0], 1u);
			if (arrived == gridDim.x - 1) {
				bar[0] = 0;
				__threadfence();
				atomicAdd(&bar[1], 1u);
			} else {
				while (*((volatile unsigned int*)&bar[1]) == gen) { }
			}
			__threadfence();
		}
		++gen;
		__syncthreads();
	}
	__device__ __forceinline__ void flag_or(int k, int v) { atomicOr(&fl[k], v); }
	__device__ __forceinline__ int flag_get(int k) const { return __ldcg(&fl[k]); }
	__device__ __forceinline__ void flag_zero(int k) { if (blockIdx.x == 0 && threadIdx.x == 0) __stcg(&fl[k], 0); }
};

/* everything one pair needs, resolved for the slot it runs in */
struct Job {
	int tl, ql, doff;
	const uint32_t *T, *Q;
	const uint8_t *T8, *Q8;
	int32_t *ring, *ring2;
	uint8_t *arena;
	long long arena_cap;
	long long *rowtab;
	long long rowtab_cap;
	int *snaphdr;
	long long *snapoff;
	int *seg;
	int *slo, *shi; /* shared: [lo,hi] of every ring slice */
	int *scr;       /* shared scratch, >= 4 ints */
};

/* ------------------------------------------------------------------------------------------ */
/* wf_stripe_shrink (miniwfa.c:144-171)                                                        */
/* ------------------------------------------------------------------------------------------ */

__device__ bool diag_alive(const Job &J, const Pen &pen, int pitch, int d)
{
	for (int j = 0; j < pen.nring; ++j) {
		if (d < J.slo[j] || d > J.shi[j]) continue;
		const int32_t *p = J.ring + (size_t)j * 5 * pitch + d + J.doff;
		for (int a = 0; a < 5; ++a)
			if (on_matrix(d, ld_ring1(p + (size_t)a * pitch), J.tl, J.ql)) return true;
	}
	return false;
}

/* every CTA of the group runs this redundantly and gets the same answer */
__device__ bool shrink_band(const Job &J, const Pen &pen, int pitch, int &wflo, int &wfhi)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (warp == 0) {
		int nl = wfhi + 1;
		for (int base = wflo; base <= wfhi; base += 32) {
			const int d = base + lane;
			const unsigned m = __ballot_sync(0xffffffffu, d <= wfhi && diag_alive(J, pen, pitch, d));
			if (m) { nl = base + __ffs(m) - 1; break; }
		}
		if (lane == 0) J.scr[0] = nl;
	}
	__syncthreads();
	const int nl = J.scr[0];
	if (warp == 0) {
		int nh = nl - 1;
		for (int base = wfhi; base >= nl; base -= 32) {
			const int d = base - lane;
			const unsigned m = __ballot_sync(0xffffffffu, d >= nl && diag_alive(J, pen, pitch, d));
			if (m) { nh = base - (__ffs(m) - 1); break; }
		}
		if (lane == 0) J.scr[1] = nh;
	}
	__syncthreads();
	const int nh = J.scr[1];
	__syncthreads(); /* scr is reused */
	if (nl > wfhi || nh < nl) return false; /* the reference asserts here (:157, :169) */
	wflo = nl, wfhi = nh;
	return true;
}

/* ------------------------------------------------------------------------------------------ */
/* one score step over the band [lo,hi]                                                        */
/* ------------------------------------------------------------------------------------------ */

struct StepRows {
	const int32_t *Hx, *Ho1, *Ho2, *pE1, *pF1, *pE2, *pF2; /* wf_next_prep, miniwfa.c:252-257 */
	int32_t *nH, *nE1, *nF1, *nE2, *nF2;
};

__device__ __forceinline__ StepRows step_rows(int32_t *ring, const Pen &pen, int pitch, int nslot)
{
	StepRows r;
	const int n = pen.nring;
	int sx = nslot - pen.x, so1 = nslot - pen.oe1, so2 = nslot - pen.oe2, se1 = nslot - pen.e1, se2 = nslot - pen.e2;
	if (sx < 0) sx += n;
	if (so1 < 0) so1 += n;
	if (so2 < 0) so2 += n;
	if (se1 < 0) se1 += n;
	if (se2 < 0) se2 += n;
	const size_t p = (size_t)pitch;
	r.Hx  = ring + (size_t)sx  * 5 * p;
	r.Ho1 = ring + (size_t)so1 * 5 * p;
	r.Ho2 = ring + (size_t)so2 * 5 * p;
	r.pE1 = ring + ((size_t)se1 * 5 + 1) * p;
	r.pF1 = ring + ((size_t)se1 * 5 + 2) * p;
	r.pE2 = ring + ((size_t)se2 * 5 + 3) * p;
	r.pF2 = ring + ((size_t)se2 * 5 + 4) * p;
	r.nH  = ring + (size_t)nslot * 5 * p;
	r.nE1 = r.nH + p, r.nF1 = r.nH + 2 * p, r.nE2 = r.nH + 3 * p, r.nF2 = r.nH + 4 * p;
	return r;
}

#define I4(v, j) ((j) == 0 ? (v).x : (j) == 1 ? (v).y : (j) == 2 ? (v).z : (v).w)

/*
 * Each thread owns 4 consecutive diagonals (one aligned int4 of every row); the d-1 / d+1
 * neighbours inside the int4 come from registers, across threads from warp shuffles, and
 * across warps from two scalar loads.  Returns this thread's flag bits.
 */
template<int MODE, class G>
__device__ __forceinline__ int step_cells(G &g, const Job &J, const StepRows &R, const StepRows &S,
                                          int lo, int hi, uint8_t *tbrow)
{
	const int lane = threadIdx.x & 31;
	const int doff = J.doff, tl = J.tl, ql = J.ql;
	const int ilo = lo + doff, ihi = hi + doff, i0 = ilo & ~3;
	const int dfin = ql - tl;
	int myfl = 0;
	for (int wbase = i0 + 4 * (g.rank() - lane); wbase <= ihi; wbase += 4 * g.size()) {
		const int idx = wbase + 4 * lane;
		const bool act = idx <= ihi + 1;
		int4 ho1, pe1, pf1, ho2, pe2, pf2, hx;
		if (act) {
			ho1 = ld_ring4(R.Ho1 + idx); pe1 = ld_ring4(R.pE1 + idx); pf1 = ld_ring4(R.pF1 + idx);
			ho2 = ld_ring4(R.Ho2 + idx); pe2 = ld_ring4(R.pE2 + idx); pf2 = ld_ring4(R.pF2 + idx);
			hx  = ld_ring4(R.Hx + idx);
		} else {
			ho1 = pe1 = pf1 = ho2 = pe2 = pf2 = hx = make_int4(NEG_INF, NEG_INF, NEG_INF, NEG_INF);
		}
		int4 sho1, spe1, spf1, sho2, spe2, spf2, shx; /* provenance labels (low-memory pass 1) */
		if (MODE == MODE_SEG1) {
			if (act) {
				sho1 = ld_ring4(S.Ho1 + idx); spe1 = ld_ring4(S.pE1 + idx); spf1 = ld_ring4(S.pF1 + idx);
				sho2 = ld_ring4(S.Ho2 + idx); spe2 = ld_ring4(S.pE2 + idx); spf2 = ld_ring4(S.pF2 + idx);
				shx  = ld_ring4(S.Hx + idx);
			} else {
				sho1 = spe1 = spf1 = sho2 = spe2 = spf2 = shx = make_int4(NEG_INF, NEG_INF, NEG_INF, NEG_INF);
			}
		}
		/* what this cell offers to its right neighbour (insertion, d+1) and left neighbour (deletion, d-1) */
		int A1[6], A2[6], C1[6], C2[6];       /* index j+1 for element j; [0] = from the left, [5] = from the right */
		int bA1[6], bA2[6], bC1[6], bC2[6];   /* 1 iff the extend source strictly beats the open source */
		int LA1[6], LA2[6], LC1[6], LC2[6];   /* offered labels */
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int o1 = I4(ho1, j), o2 = I4(ho2, j);
			const int e1 = I4(pe1, j), e2 = I4(pe2, j), f1 = I4(pf1, j), f2 = I4(pf2, j);
			A1[j + 1] = max(o1, e1), A2[j + 1] = max(o2, e2);
			C1[j + 1] = max(o1, f1) + 1, C2[j + 1] = max(o2, f2) + 1;
			if (MODE != MODE_SCORE) {
				bA1[j + 1] = o1 < e1, bA2[j + 1] = o2 < e2, bC1[j + 1] = o1 < f1, bC2[j + 1] = o2 < f2;
			}
			if (MODE == MODE_SEG1) {
				LA1[j + 1] = bA1[j + 1] ? I4(spe1, j) : I4(sho1, j);
				LA2[j + 1] = bA2[j + 1] ? I4(spe2, j) : I4(sho2, j);
				LC1[j + 1] = bC1[j + 1] ? I4(spf1, j) : I4(sho1, j);
				LC2[j + 1] = bC2[j + 1] ? I4(spf2, j) : I4(sho2, j);
			}
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
		if (MODE == MODE_SEG1) {
			LA1[0] = __shfl_up_sync(0xffffffffu, LA1[4], 1);
			LA2[0] = __shfl_up_sync(0xffffffffu, LA2[4], 1);
			LC1[5] = __shfl_down_sync(0xffffffffu, LC1[1], 1);
			LC2[5] = __shfl_down_sync(0xffffffffu, LC2[1], 1);
		}
		if (lane == 0 && act) { /* left neighbour belongs to another warp */
			const int o1 = ld_ring1(R.Ho1 + idx - 1), e1 = ld_ring1(R.pE1 + idx - 1);
			const int o2 = ld_ring1(R.Ho2 + idx - 1), e2 = ld_ring1(R.pE2 + idx - 1);
			A1[0] = max(o1, e1), A2[0] = max(o2, e2);
			if (MODE != MODE_SCORE) bA1[0] = o1 < e1, bA2[0] = o2 < e2;
			if (MODE == MODE_SEG1) {
				LA1[0] = bA1[0] ? ld_ring1(S.pE1 + idx - 1) : ld_ring1(S.Ho1 + idx - 1);
				LA2[0] = bA2[0] ? ld_ring1(S.pE2 + idx - 1) : ld_ring1(S.Ho2 + idx - 1);
			}
		}
		if (lane == 31 && act) { /* right neighbour belongs to another warp */
			const int o1 = ld_ring1(R.Ho1 + idx + 4), f1 = ld_ring1(R.pF1 + idx + 4);
			const int o2 = ld_ring1(R.Ho2 + idx + 4), f2 = ld_ring1(R.pF2 + idx + 4);
			C1[5] = max(o1, f1) + 1, C2[5] = max(o2, f2) + 1;
			if (MODE != MODE_SCORE) bC1[5] = o1 < f1, bC2[5] = o2 < f2;
			if (MODE == MODE_SEG1) {
				LC1[5] = bC1[5] ? ld_ring1(S.pF1 + idx + 4) : ld_ring1(S.Ho1 + idx + 4);
				LC2[5] = bC2[5] ? ld_ring1(S.pF2 + idx + 4) : ld_ring1(S.Ho2 + idx + 4);
			}
		}
		if (!act) continue; /* after the shuffles: inactive lanes have nothing to store */

		int vH[4], vE1[4], vF1[4], vE2[4], vF2[4], st[4], kmax[4];
		int lH[4], lE1[4], lF1[4], lE2[4], lF2[4];
		uint32_t tbw = 0;
		bool ext[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int E1 = A1[j], E2 = A2[j], F1 = C1[j + 2], F2 = C2[j + 2];
			const int e = max(E1, E2), f = max(F1, F2), gmx = max(e, f), hxp = I4(hx, j) + 1;
			const int H = max(hxp, gmx);
			vE1[j] = E1, vE2[j] = E2, vF1[j] = F1, vF2[j] = F2, vH[j] = H;
			if (MODE != MODE_SCORE) { /* the 7-bit pack, miniwfa.c:290-306 */
				const int z = hxp >= gmx ? 0 : (e >= f ? (E1 >= E2 ? 1 : 3) : (F1 >= F2 ? 2 : 4));
				st[j] = z;
				tbw |= (uint32_t)(z | bA1[j] << 3 | bC1[j + 2] << 4 | bA2[j] << 5 | bC2[j + 2] << 6) << (8 * j);
				if (MODE == MODE_SEG1) { /* replay, miniwfa.c:506-522 */
					lE1[j] = LA1[j], lE2[j] = LA2[j], lF1[j] = LC1[j + 2], lF2[j] = LC2[j + 2];
					lH[j] = z == 0 ? I4(shx, j) : z == 1 ? lE1[j] : z == 2 ? lF1[j] : z == 3 ? lE2[j] : lF2[j];
				}
			}
			const int ii = idx + j, d = ii - doff;
			const bool in = ii >= ilo && ii <= ihi;
			if (in && (H >= -1 || E1 >= -1 || F1 >= -1 || E2 >= -1 || F2 >= -1)) { /* edge rule, :325-326 */
				if (d == lo) myfl |= FL_LO;
				if (d == hi) myfl |= FL_HI;
			}
			kmax[j] = min(tl - 1, ql - 1 - d);
			ext[j] = in && H >= -1 && d + H >= -1 && H <= kmax[j]; /* :402 */
		}
		/* first probe of the match run for all four cells (loads issued together) */
		uint32_t px[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int d = idx + j - doff;
			const int tp = ext[j] ? vH[j] + 1 : 0, qp = ext[j] ? d + vH[j] + 1 : 0;
			px[j] = seq_word(J.T, tp) ^ seq_word(J.Q, qp);
		}
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			if (!ext[j]) continue;
			const int d = idx + j - doff, h0 = vH[j];
			int k;
			if (px[j]) k = min(h0 + ((__ffs(px[j]) - 1) >> 3), kmax[j]);
			else k = extend_run(J.T, J.Q, min(h0 + 4, kmax[j]), d, kmax[j]);
			if (d == dfin && k == tl - 1) { /* reached the end of both sequences, :405-409 */
				myfl |= FL_DONE;
				if (MODE == MODE_TB && k == h0) myfl |= st[j] << FL_LAST_SHIFT;
			}
			vH[j] = k;
		}
		/* stores */
		if (idx >= ilo && idx + 3 <= ihi) {
			st_ring4(R.nH + idx, make_int4(vH[0], vH[1], vH[2], vH[3]));
			st_ring4(R.nE1 + idx, make_int4(vE1[0], vE1[1], vE1[2], vE1[3]));
			st_ring4(R.nF1 + idx, make_int4(vF1[0], vF1[1], vF1[2], vF1[3]));
			st_ring4(R.nE2 + idx, make_int4(vE2[0], vE2[1], vE2[2], vE2[3]));
			st_ring4(R.nF2 + idx, make_int4(vF2[0], vF2[1], vF2[2], vF2[3]));
			if (MODE == MODE_SEG1) {
				st_ring4(S.nH + idx, make_int4(lH[0], lH[1], lH[2], lH[3]));
				st_ring4(S.nE1 + idx, make_int4(lE1[0], lE1[1], lE1[2], lE1[3]));
				st_ring4(S.nF1 + idx, make_int4(lF1[0], lF1[1], lF1[2], lF1[3]));
				st_ring4(S.nE2 + idx, make_int4(lE2[0], lE2[1], lE2[2], lE2[3]));
				st_ring4(S.nF2 + idx, make_int4(lF2[0], lF2[1], lF2[2], lF2[3]));
			}
		} else {
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const int ii = idx + j;
				if (ii < ilo || ii > ihi) continue;
				st_ring1(R.nH + ii, vH[j]); st_ring1(R.nE1 + ii, vE1[j]); st_ring1(R.nF1 + ii, vF1[j]);
				st_ring1(R.nE2 + ii, vE2[j]); st_ring1(R.nF2 + ii, vF2[j]);
				if (MODE == MODE_SEG1) {
					st_ring1(S.nH + ii, lH[j]); st_ring1(S.nE1 + ii, lE1[j]); st_ring1(S.nF1 + ii, lF1[j]);
					st_ring1(S.nE2 + ii, lE2[j]); st_ring1(S.nF2 + ii, lF2[j]);
				}
			}
		}
		if (MODE == MODE_TB && idx <= ihi) __stcs(reinterpret_cast<uint32_t*>(tbrow + idx), tbw);
	}
	return myfl;
}

/* ------------------------------------------------------------------------------------------ */
/* wf_snapshot1 (miniwfa.c:451-474): copy the provenance ring out, relabel every cell          */
/* ------------------------------------------------------------------------------------------ */

template<class G>
__device__ bool take_snapshot(G &g, const Job &J, const Pen &pen, int pitch, int cur, int s,
                              int snap_idx, int snap_cap, long long &used)
{
	const int n = pen.nring, hw = 2 + 2 * n;
	long long total = 0;
	for (int j = 0; j < n; ++j) {
		int slot = cur + 1 + j; if (slot >= n) slot -= n;
		total += 5LL * (J.shi[slot] - J.slo[slot] + 1);
	}
	if (snap_idx >= snap_cap || used + total * 4 > J.arena_cap || total > 0x7fffffffLL) return false;
	int32_t *x = reinterpret_cast<int32_t*>(J.arena + used);
	int *hdr = J.snaphdr + (size_t)snap_idx * hw;
	if (g.rank() == 0) { hdr[0] = s; hdr[1] = (int)total; J.snapoff[snap_idx] = used; }
	int pref = 0;
	for (int j = 0; j < n; ++j) { /* oldest slice first */
		int slot = cur + 1 + j; if (slot >= n) slot -= n;
		const int slo = J.slo[slot], cnt = 5 * (J.shi[slot] - slo + 1);
		if (g.rank() == 0) { hdr[2 + 2 * j] = slo; hdr[3 + 2 * j] = cnt; }
		int32_t *base = J.ring2 + (size_t)slot * 5 * pitch + J.doff;
		for (int u = g.rank(); u < cnt; u += g.size()) {
			const int d = slo + u / 5, a = u % 5;
			int32_t *p = base + (size_t)a * pitch + d;
			x[pref + u] = ld_ring1(p);
			st_ring1(p, pref + u);
		}
		pref += cnt;
	}
	used += (total * 4 + 15) & ~15LL;
	return true;
}

/* ------------------------------------------------------------------------------------------ */
/* the score loop of mwf_wfa_core (:397-426) / mwf_wfa_seg (:569-589)                           */
/* ------------------------------------------------------------------------------------------ */

struct PassOut {
	int s, last, n_snap;
	long long n_iter;
};

template<int MODE, class G>
__device__ int run_pass(G &g, const KParams &P, const Job &J, int n_seg, PassOut &out)
{
	const Pen pen = P.pen;
	const int n = pen.nring, m1 = n, pitch = P.pitch;
	const int tl = J.tl, ql = J.ql, doff = J.doff;
	/* wf_stripe_init (:103-121): every slice is [0,0] holding NEG_INF (pads included) */
	for (int t = g.rank(); t < n * 5 * (2 * m1 + 1); t += g.size()) {
		const int row = t / (2 * m1 + 1), c = t % (2 * m1 + 1);
		st_ring1(J.ring + (size_t)row * pitch + doff - m1 + c, NEG_INF);
		if (MODE == MODE_SEG1) st_ring1(J.ring2 + (size_t)row * pitch + doff - m1 + c, NEG_INF);
	}
	for (int t = threadIdx.x; t < n; t += blockDim.x) J.slo[t] = J.shi[t] = 0;
	g.flag_zero(0); g.flag_zero(1); g.flag_zero(2);
	g.sync();
	if (g.rank() == 0) { /* score 0: H[0] = -1, then its match run */
		const int kmax = min(tl - 1, ql - 1);
		const int k = extend_run(J.T, J.Q, -1, 0, kmax);
		st_ring1(J.ring + doff, k);
		if (MODE == MODE_SEG1) st_ring1(J.ring2 + doff, -1);
		if (k == tl - 1 && k == ql - 1) g.flag_or(0, FL_DONE);
	}
	g.sync();
	int s = 0, cur = 0, wflo = 0, wfhi = 0, sid = 0, n_snap = 0;
	long long n_iter = 0, used = 0;
	out.s = 0, out.last = MODE == MODE_SEG1 ? -1 : 0, out.n_snap = 0, out.n_iter = 0;
	if (g.flag_get(0) & FL_DONE) return ST_OK;

	for (;;) {
		if (MODE == MODE_TB && sid < n_seg && J.seg[2 * sid] == s) { /* band collapse, :413-416 */
			wflo = wfhi = J.seg[2 * sid + 1];
			++sid;
		}
		const int lo = wflo > -tl ? wflo - 1 : -tl, hi = wfhi < ql ? wfhi + 1 : ql; /* :417-418 */
		const int snew = s + 1, nslot = cur + 1 == n ? 0 : cur + 1;
		if (MODE == MODE_SEG1 && snew % P.step == 0) { /* :585-586 */
			if (!take_snapshot(g, J, pen, pitch, cur, s, n_snap, P.snap_cap, used)) return ST_ARENA;
			++n_snap;
			g.sync();
		}
		uint8_t *tbrow = 0;
		if (MODE == MODE_TB) { /* wf_tb_add (:33-44): one byte per cell of this score */
			const int i0 = (lo + doff) & ~3;
			const long long rowsize = ((hi + doff - i0) | 3) + 1;
			if (used + rowsize > J.arena_cap || snew >= J.rowtab_cap) return ST_ARENA;
			tbrow = J.arena + used - i0;
			if (g.rank() == 0) J.rowtab[snew] = used - i0;
			used += rowsize;
		}
		const StepRows R = step_rows(J.ring, pen, pitch, nslot);
		StepRows S = R;
		if (MODE == MODE_SEG1) S = step_rows(J.ring2, pen, pitch, nslot);
		const int myfl = step_cells<MODE>(g, J, R, S, lo, hi, tbrow);
		for (int t = g.rank(); t < 2 * m1 * 5; t += g.size()) { /* NEG_INF pads, :96-99 */
			const int a = t / (2 * m1), c = t % (2 * m1);
			const int idx = c < m1 ? lo + doff - 1 - c : hi + doff + 1 + (c - m1);
			st_ring1(R.nH + (size_t)a * pitch + idx, NEG_INF);
			if (MODE == MODE_SEG1) st_ring1(S.nH + (size_t)a * pitch + idx, NEG_INF);
		}
		if (threadIdx.x == 0) J.slo[nslot] = lo, J.shi[nslot] = hi;
		g.flag_zero((snew + 1) % 3);
		if (myfl) g.flag_or(snew % 3, myfl);
		g.sync();
		const int fl = g.flag_get(snew % 3);
		s = snew, cur = nslot;
		n_iter += hi - lo + 1; /* :421 */
		if (fl & FL_LO) wflo = lo;
		if (fl & FL_HI) wfhi = hi;
		if ((s & 0xff) == 0) { /* :420 */
			if (!shrink_band(J, pen, pitch, wflo, wfhi)) return ST_SHRINK;
			g.sync(); /* the oldest slice is overwritten next step; everyone must be done reading it */
		}
		out.s = s, out.n_iter = n_iter, out.n_snap = n_snap;
		if (MODE != MODE_SEG1 && ((P.max_iter > 0 && n_iter > P.max_iter) || (P.max_s > 0 && s > P.max_s)))
			return ST_STOPPED; /* :422-425; tested before the next extend, so it wins over FL_DONE */
		if (fl & FL_DONE) {
			if (MODE == MODE_TB) out.last = fl >> FL_LAST_SHIFT;
			if (MODE == MODE_SEG1) out.last = ld_ring1(J.ring2 + (size_t)cur * 5 * pitch + (ql - tl) + doff); /* :577 */
			return ST_OK;
		}
	}
}

/* wf_traceback_seg (miniwfa.c:528-549); one thread */
__device__ bool trace_checkpoints(const Job &J, const Pen &pen, int n_snap, int last)
{
	const int n = pen.nring, hw = 2 + 2 * n;
	for (int j = n_snap - 1; j >= 0; --j) {
		const int *hdr = J.snaphdr + (size_t)j * hw;
		const int32_t *x = reinterpret_cast<const int32_t*>(J.arena + J.snapoff[j]);
		int k, m = 0;
		for (k = 0; k < n; ++k) {
			if (last >= m && last < m + hdr[3 + 2 * k]) break;
			m += hdr[3 + 2 * k];
		}
		if (k == n) return false;
		J.seg[2 * j] = hdr[0] - (n - k - 1);
		J.seg[2 * j + 1] = hdr[2 + 2 * k] + (last - m) / 5;
		last = __ldcg(x + last);
	}
	return last == -1;
}

/* wf_traceback (miniwfa.c:329-377); one warp, all lanes walk in lock step, match runs 32 bytes at a time.
 * CIGAR words are written backwards from cig_end so that no final reversal is needed.
 *
 * The walk is a chain of dependent DRAM reads, one traceback byte per difference.  Where the next byte lies depends only on the
 * (state, extension bit) the current byte resolves to: nine possible (row, diagonal) successors -- a mismatch, and open / extend
 * of the two insertion and the two deletion states.  Lanes 0..8 fetch all nine as soon as the current cell is known, so the
 * row-table lookup and the byte are in flight during the match run of the cell instead of after it, and the match runs' sequence
 * lines are prefetched 256 bases ahead (150 kb pair, 10 924 CIGAR operations: 7.5 -> 6.8 ms). */
#define TB_ROWWIN 2048 /* entries of the row table a traceback warp keeps in shared memory */

/* The bytes a walk can reach next lie in a cone below its position: at most max-penalty rows down and one diagonal sideways per
 * step.  TbCone keeps TBC_ROWS rows x TBC_COLS diagonals around the position in shared memory, filled by the whole warp with
 * aligned word loads that are all in flight together: one DRAM latency per ~25 steps of the walk instead of one per step. */
#define TBC_ROWS 128
#define TBC_COLS 24 /* (+ 2 words of slack = 8 words per row: a warp instruction covers four rows, index arithmetic by shifts) */
#define TBC_WORDS (TBC_COLS / 4 + 2) /* per row: the columns plus the slack of rounding the row's address down to a word */
struct TbCone {
	uint32_t w[TBC_ROWS * TBC_WORDS];
	long long rt[TBC_ROWS]; /* row-table entries of the cached rows, rt[r] of row top - r */
};
struct TbConePos { int top, lo; long long c0; }; /* rows [lo, top], columns [c0, c0 + TBC_COLS) (column = diagonal + doff) */

/* the match runs walk both sequences backwards, 32 bytes at a time: a window of each in shared memory, refilled every ~1000 bases */
#define TBS_WIN 1024
struct TbSeqWin { uint8_t q[TBS_WIN], t[TBS_WIN]; };
__device__ __forceinline__ int tbseq_fill(uint8_t *dst, const uint8_t *src, int pos) /* bytes [base, pos] of src (16-byte aligned); returns base */
{
	const int lane = threadIdx.x & 31, b = max(0, pos - (TBS_WIN - 16)) & ~15;
	uint4 v[TBS_WIN / 512];
	__syncwarp();
#pragma unroll
	for (int u = 0; u < TBS_WIN / 512; ++u) {
		const int off = 16 * (lane + 32 * u);
		v[u] = b + off <= pos ? __ldg(reinterpret_cast<const uint4*>(src + b + off)) : make_uint4(0, 0, 0, 0);
	}
#pragma unroll
	for (int u = 0; u < TBS_WIN / 512; ++u) *reinterpret_cast<uint4*>(dst + 16 * (lane + 32 * u)) = v[u];
	__syncwarp();
	return b;
}

template<class RowTab>
__device__ __forceinline__ void tbcone_fill(TbCone *cn, TbConePos &cp, const uint8_t *arena, long long arena_cap, const RowTab &rowtab, int row, long long col)
{
	const int lane = threadIdx.x & 31;
	cp.top = row, cp.lo = max(1, row - TBC_ROWS + 1), cp.c0 = col - TBC_COLS / 2;
	const int nrows = cp.top - cp.lo + 1;
	__syncwarp();
	{ /* (loads first, stores after: every load of the warp is in flight before the first is waited for) */
		long long v[TBC_ROWS / 32];
#pragma unroll
		for (int u = 0; u < TBC_ROWS / 32; ++u) v[u] = lane + 32 * u < nrows ? rowtab(row - (lane + 32 * u)) : 0;
#pragma unroll
		for (int u = 0; u < TBC_ROWS / 32; ++u) cn->rt[lane + 32 * u] = v[u];
	}
	__syncwarp();
	const long long last_word = (arena_cap - 4) & ~3LL;
	constexpr int PER = TBC_ROWS * TBC_WORDS / 32;
	uint32_t v[PER];
#pragma unroll
	for (int u = 0; u < PER; ++u) {
		const int idx = lane + 32 * u, r = idx / TBC_WORDS, q = idx - r * TBC_WORDS;
		long long a = ((cn->rt[r] + cp.c0) & ~3LL) + 4 * q;
		a = a < 0 ? 0 : a > last_word ? last_word : a; /* (a word off the arena is never one the walk uses) */
		v[u] = r < nrows ? __ldcg(reinterpret_cast<const uint32_t*>(arena + a)) : 0u;
	}
#pragma unroll
	for (int u = 0; u < PER; ++u) cn->w[lane + 32 * u] = v[u];
	__syncwarp();
}
/* the same fill with cp.async: nothing is waited for and no register holds the data; the walk goes on in the cone it is in and
 * switches to this one, long landed, when it leaves the other (cp.async.wait_group 0 + __syncwarp before the first read) */
template<class RowTab>
__device__ __forceinline__ void tbcone_prefetch(TbCone *cn, TbConePos &cp, const uint8_t *arena, long long arena_cap, const RowTab &rowtab, int row, long long col)
{
	const int lane = threadIdx.x & 31;
	cp.top = row, cp.lo = max(1, row - TBC_ROWS + 1), cp.c0 = col - TBC_COLS / 2;
	const int nrows = cp.top - cp.lo + 1;
	__syncwarp();
#pragma unroll
	for (int u = 0; u < TBC_ROWS / 32; ++u) cn->rt[lane + 32 * u] = lane + 32 * u < nrows ? rowtab(row - (lane + 32 * u)) : 0;
	__syncwarp();
	const long long last_word = (arena_cap - 4) & ~3LL;
	constexpr int PER = TBC_ROWS * TBC_WORDS / 32;
#pragma unroll 8
	for (int u = 0; u < PER; ++u) {
		const int idx = lane + 32 * u, r = idx / TBC_WORDS, q = idx - r * TBC_WORDS;
		long long a = ((cn->rt[r] + cp.c0) & ~3LL) + 4 * q;
		a = a < 0 ? 0 : a > last_word ? last_word : a;
		if (r < nrows) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(&cn->w[idx])), "l"(arena + a) : "memory");
	}
	asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ bool tbcone_has(const TbConePos &cp, int row, long long col)
{
	return row <= cp.top && row >= cp.lo && col >= cp.c0 && col < cp.c0 + TBC_COLS;
}
__device__ __forceinline__ int tbcone_get(const TbCone *cn, const TbConePos &cp, int row, long long col)
{
	const int r = cp.top - row;
	const long long rt = cn->rt[r];
	const int b = (int)((rt + col) - ((rt + cp.c0) & ~3LL));
	return reinterpret_cast<const uint8_t*>(cn->w)[r * (TBC_WORDS * 4) + b];
}

/* rtw: TB_ROWWIN entries of shared memory for a sliding window of the row table (the walk needs rowtab[row - penalty] before every
 * traceback byte: from shared memory that is one dependent global load less per step), or null */
__device__ int traceback_warp(const Job &J, const Pen &pen, int s_final, int last, uint32_t *cig_end, int *end_state, long long *rtw = 0, TbCone *cone = 0, TbSeqWin *sw = 0, TbCone *cone_b = 0)
{
	const int lane = threadIdx.x & 31, doff = J.doff;
	int i = J.ql - 1, k = J.tl - 1, row = s_final, n_out = 0, cur_op = -1;
	uint32_t cur_len = 0;
	uint32_t *wp = cig_end;
	/* successor c of a cell: c = 0 mismatch; c = 1 + 2 g + (open ? 1 : 0), g = 0 state 1 (I, e1), 1 state 3 (I, e2), 2 state 2 (D, e1), 3 state 4 (D, e2) */
	const int cg = (lane - 1) >> 1, copen = (lane - 1) & 1;
	const int cpen = lane == 0 ? pen.x : (cg & 1) ? (copen ? pen.oe2 : pen.e2) : (copen ? pen.oe1 : pen.e1);
	const int cdd = lane == 0 ? 0 : cg < 2 ? -1 : 1; /* an insertion steps to diagonal d - 1, a deletion to d + 1 */
	int win_lo = 0x7fffffff; /* rows [win_lo, ...] of the table are in rtw (rows only go down) */
	const int maxpen = max(pen.x, max(pen.oe1, pen.oe2));
#define RT(row_) ((row_) >= win_lo ? rtw[(row_) - win_lo] : J.rowtab[row_])
#define RT_REFILL() do { \
		if (rtw && win_lo > 1 && row - maxpen < win_lo) { /* slide the window down to end at the current row */ \
			const int nlo = max(1, row - TB_ROWWIN + 1); \
			__syncwarp(); \
			for (int r_ = nlo + lane; r_ <= row; r_ += 32) rtw[r_ - nlo] = J.rowtab[r_]; \
			__syncwarp(); \
			win_lo = nlo; \
		} \
	} while (0)
	RT_REFILL();
	TbConePos cp;
	cp.top = -1, cp.lo = 0, cp.c0 = 0;
	int qwb = 0x7fffffff, twb = 0x7fffffff; /* first byte of each sequence window */
	auto rowtab_at = [&](int r_) -> long long { return RT(r_); };
	TbConePos pp; /* the cone being fetched ahead into cone_b (or into cone, they swap), when `pending` */
	bool pending = false;
	pp.top = -1, pp.lo = 0, pp.c0 = 0;
#define CONE_NEEDS(cp_, col_) (row > (cp_).top || (row - maxpen < (cp_).lo && (cp_).lo > 1) || (col_) - 1 < (cp_).c0 || (col_) + 1 >= (cp_).c0 + TBC_COLS)
#define CONE_REFILL() do { \
		const long long col_ = (long long)(i - k + doff); \
		if (cone && row >= 1 && CONE_NEEDS(cp, col_)) { \
			bool have_ = false; \
			if (pending) { /* the cone fetched ahead has landed long ago */ \
				asm volatile("cp.async.wait_group 0;" ::: "memory"); \
				__syncwarp(); \
				pending = false; \
				if (!CONE_NEEDS(pp, col_)) { TbCone *t_ = cone; cone = cone_b; cone_b = t_; cp = pp; have_ = true; } \
			} \
			if (!have_) tbcone_fill(cone, cp, J.arena, J.arena_cap, rowtab_at, row, col_); \
		} else if (cone_b && !pending && row >= 1 && cp.lo > 1 && row - cp.lo < 56) /* 24 rows further down is where the walk will be when it leaves this cone */ \
			tbcone_prefetch(cone_b, pp, J.arena, J.arena_cap, rowtab_at, max(1, row - 24), col_), pending = true; \
	} while (0)
	CONE_REFILL();
	int x = row < 1 ? 0 : cone ? tbcone_get(cone, cp, row, i - k + doff) : (int)__ldcg(J.arena + RT(row) + (i - k + doff)); /* the byte of (row, i - k): a match run keeps the diagonal */
#define CIG_PUSH(op_, len_) do { \
		if ((op_) == cur_op) cur_len += (len_); \
		else { if (cur_op >= 0) { --wp; if (lane == 0) *wp = cur_len << 4 | (uint32_t)cur_op; ++n_out; } cur_op = (op_), cur_len = (len_); } \
	} while (0)
	/* the bytes of the nine successors of cell (row, i - k); a successor off the band reads a neighbouring row's byte, never used */
#define SUCCESSORS(dst_) do { \
		(dst_) = 0; \
		if (lane < 9 && row - cpen >= 1) { \
			if (cone && tbcone_has(cp, row - cpen, (long long)(i - k + cdd + doff))) (dst_) = tbcone_get(cone, cp, row - cpen, (long long)(i - k + cdd + doff)); \
			else { \
				long long off_ = RT(row - cpen) + (i - k + cdd + doff); \
				off_ = off_ < 0 ? 0 : off_ >= J.arena_cap ? J.arena_cap - 1 : off_; \
				(dst_) = __ldcg(J.arena + off_); \
			} \
		} \
	} while (0)
	int xs;
	SUCCESSORS(xs);
#ifdef MWF_PHASE_PROF
	long long tw_run = 0, tw_rest = 0, tw_fill = 0, tw_n = 0, tw_chunks = 0, tw_t = clock64();
#define TW(acc) do { const long long n_ = clock64(); acc += n_ - tw_t; tw_t = n_; } while (0)
#else
#define TW(acc) do {} while (0)
#endif
	while (i >= 0 && k >= 0) {
		if (!sw && lane == 9 && i >= 256) asm volatile("prefetch.global.L1 [%0];" :: "l"(J.Q8 + (i - 256))); /* the match runs walk both */
		if (!sw && lane == 10 && k >= 256) asm volatile("prefetch.global.L1 [%0];" :: "l"(J.T8 + (k - 256))); /* sequences backwards */
		if (last == 0) { /* greedy backward matches, :335-341 */
			int run = 0;
			for (;;) {
				const int ii = i - lane, kk = k - lane;
				bool same;
				if (sw) {
					if (qwb > max(0, i - 31)) qwb = tbseq_fill(sw->q, J.Q8, i);
					if (twb > max(0, k - 31)) twb = tbseq_fill(sw->t, J.T8, k);
					same = ii >= 0 && kk >= 0 && sw->q[ii - qwb] == sw->t[kk - twb];
				} else same = ii >= 0 && kk >= 0 && __ldg(J.Q8 + ii) == __ldg(J.T8 + kk);
				const unsigned m = __ballot_sync(0xffffffffu, !same);
				if (m) { const int c = __ffs(m) - 1; run += c, i -= c, k -= c; break; }
				run += 32, i -= 32, k -= 32;
			}
			if (run > 0) CIG_PUSH(7, (uint32_t)run);
			if (i < 0 || k < 0) break;
		}
		TW(tw_run);
		const int state = last == 0 ? (x & 7) : last;
		const int ext = state > 0 ? (x >> (state + 2)) & 1 : 0;
		int csel = 0, xs_next;
		if (state == 0) { CIG_PUSH(8, 1u); --i, --k; row -= pen.x; }
		else if (state == 1) { CIG_PUSH(1, 1u); --i; row -= ext ? pen.e1 : pen.oe1; csel = 1; }
		else if (state == 3) { CIG_PUSH(1, 1u); --i; row -= ext ? pen.e2 : pen.oe2; csel = 3; }
		else if (state == 2) { CIG_PUSH(2, 1u); --k; row -= ext ? pen.e1 : pen.oe1; csel = 5; }
		else { CIG_PUSH(2, 1u); --k; row -= ext ? pen.e2 : pen.oe2; csel = 7; }
		if (state > 0 && !ext) ++csel;
		last = (state > 0 && ext) ? state : 0;
		x = __shfl_sync(0xffffffffu, xs, csel);
		RT_REFILL();
		TW(tw_rest);
		CONE_REFILL();
		TW(tw_fill);
		SUCCESSORS(xs_next); /* (issuing these before the shuffle was slower on the box: 8.1 ms against 6.8 ms on the 150 kb pair) */
		xs = xs_next;
		TW(tw_rest);
#ifdef MWF_PHASE_PROF
		++tw_n;
#endif
	}
#ifdef MWF_PHASE_PROF
	if (lane == 0) printf("[walk] %lld steps: match runs %lld, cone fills %lld, rest %lld cycles per step\n", tw_n, tw_run / max(tw_n, 1LL), tw_fill / max(tw_n, 1LL), tw_rest / max(tw_n, 1LL));
#endif
#undef TW
	if (pending) asm volatile("cp.async.wait_group 0;" ::: "memory");
#undef SUCCESSORS
#undef CONE_REFILL
#undef CONE_NEEDS
#undef RT_REFILL
#undef RT
	end_state[0] = row, end_state[1] = i, end_state[2] = k; /* :367 */
	if (i >= 0) CIG_PUSH(1, (uint32_t)(i + 1));       /* :368 */
	else if (k >= 0) CIG_PUSH(2, (uint32_t)(k + 1));  /* :369 */
	if (cur_op >= 0) { --wp; if (lane == 0) *wp = cur_len << 4 | (uint32_t)cur_op; ++n_out; }
#undef CIG_PUSH
	return n_out;
}

/* mwf_wfa_exact (:603-615) for one pair in one slot */
template<class G>
__device__ void run_pair(G &g, const KParams &P, int pi, int slot, int *smem)
{
	const PairDesc pd = P.pairs[pi];
	Job J;
	const int n = P.pen.nring;
	J.tl = pd.tl, J.ql = pd.ql;
	J.doff = pd.tl + n + 8;
	J.T8 = P.seq + pd.t_off, J.Q8 = P.seq + pd.q_off;
	J.T = reinterpret_cast<const uint32_t*>(J.T8), J.Q = reinterpret_cast<const uint32_t*>(J.Q8);
	J.ring = P.ring + (size_t)slot * P.ring_stride;
	J.ring2 = P.ring2 ? P.ring2 + (size_t)slot * P.ring_stride : 0;
	J.arena = P.arena ? P.arena + (size_t)slot * P.arena_stride : 0;
	J.arena_cap = P.arena_stride;
	J.rowtab = P.rowtab ? P.rowtab + (size_t)slot * P.rowtab_stride : 0;
	J.rowtab_cap = P.rowtab_stride;
	J.snaphdr = P.snaphdr ? P.snaphdr + (size_t)slot * P.snap_cap * (2 + 2 * n) : 0;
	J.snapoff = P.snapoff ? P.snapoff + (size_t)slot * P.snap_cap : 0;
	J.seg = P.seg ? P.seg + (size_t)slot * P.snap_cap * 2 : 0;
	J.scr = smem + 4;
	J.slo = smem + 8, J.shi = smem + 8 + n;

	PassOut po;
	int st, n_cigar = 0, n_seg = 0, end_state[3] = { 0, 0, 0 };
	if (!P.is_tb) {
		st = run_pass<MODE_SCORE>(g, P, J, 0, po);
	} else {
		st = ST_OK;
		if (P.step > 0) { /* low-memory mode: pass 1 finds the checkpoints (:610-611) */
			st = run_pass<MODE_SEG1>(g, P, J, 0, po);
			if (st == ST_OK) {
				if (g.rank() == 0 && !trace_checkpoints(J, P.pen, po.n_snap, po.last)) g.flag_or(0, 1 << 30);
				g.sync();
				if (g.flag_get(0) & (1 << 30)) st = ST_CORRUPT;
				n_seg = po.n_snap;
				g.sync();
			}
		}
		if (st == ST_OK) st = run_pass<MODE_TB>(g, P, J, n_seg, po);
		if (st == ST_OK && g.leader_cta() && threadIdx.x < 32)
			n_cigar = traceback_warp(J, P.pen, po.s, po.last, P.cigar + pd.cigar_off + pd.cigar_cap, end_state);
	}
	if (g.rank() == 0) {
		PairOut o;
		o.s = st == ST_OK ? po.s : -1;
		o.n_cigar = n_cigar;
		o.n_iter = po.n_iter;
		o.cigar_pos = pd.cigar_off + pd.cigar_cap - n_cigar;
		o.status = st, o.pad_ = 0;
		o.end_s = end_state[0], o.end_i = end_state[1], o.end_k = end_state[2], o.pad2_ = 0;
		P.outs[pi] = o;
	}
}

__global__ void __launch_bounds__(512, 1) wfa_cta_kernel(const KParams P)
{
	extern __shared__ int smem[]; /* [0..2] flags, [3] queue slot, [4..7] scratch, then slo[n], shi[n] */
	CtaGroup g;
	g.fl = smem;
	for (;;) {
		__syncthreads();
		if (threadIdx.x == 0) smem[3] = (int)atomicAdd(P.queue, 1u);
		__syncthreads();
		const int qi = smem[3];
		if (qi >= P.n_pairs) break;
		run_pair(g, P, P.order[qi], blockIdx.x, smem);
	}
}

__global__ void __launch_bounds__(512, 1) wfa_grid_kernel(const KParams P)
{
	extern __shared__ int smem[];
	GridGroup g;
	g.fl = P.gflags, g.bar = P.bar;
	g.gen = *((volatile unsigned int*)&P.bar[1]);
	run_pair(g, P, P.single_pair, 0, smem);
}

#include "wfa_tile.cuh"

/* ------------------------------------------------------------------------------------------ */
/* host side                                                                                   */
/* ------------------------------------------------------------------------------------------ */

static std::atomic<int> g_device(-1), g_kernel(-1), g_threads(0);
static std::atomic<int> g_device_explicit(0), g_n_devices(0);
static thread_local int tl_device = -1; /* set by the per-device workers of mwf_wfa_exact_batch: the device of batches this thread creates */

static int env_int(const char *name, int dflt)
{
	const char *v = getenv(name);
	return v && *v ? atoi(v) : dflt;
}

extern "C" int mwf_b200_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

extern "C" void mwf_b200_set_device(int dev) { g_device = dev; g_device_explicit = 1; }
extern "C" void mwf_b200_set_devices(int n) { g_n_devices = n < 0 ? 0 : n; }

extern "C" int mwf_b200_get_device(void)
{
	int d = g_device;
	if (d < 0) {
		d = env_int("MWF_B200_DEVICE", -1);
		if (d < 0) d = env_int("LOCAL_RANK", 0);
		const int n = mwf_b200_device_count();
		if (n > 0) d %= n;
		g_device = d;
	}
	return d;
}

extern "C" void mwf_b200_set_kernel(int kernel) { g_kernel = kernel; }
extern "C" void mwf_b200_set_block_threads(int threads) { g_threads = threads; }

static int pick_kernel_pref(void)
{
	int k = g_kernel;
	if (k < 0) {
		const char *v = getenv("MWF_B200_KERNEL");
		k = MWF_B200_KERNEL_AUTO;
		if (v && !strcmp(v, "cta")) k = MWF_B200_KERNEL_CTA;
		else if (v && !strcmp(v, "grid")) k = MWF_B200_KERNEL_GRID;
		else if (v && !strcmp(v, "tile")) k = MWF_B200_KERNEL_TILE;
	}
	return k;
}

/* ------------------------------------------------------------------------------------------ */
/* workspace cache: device and pinned-host buffers survive mwf_b200_batch_destroy and are handed */
/* to the next batch of a similar size, so that a one-shot mwf_wfa_exact_batch() call does not   */
/* pay cudaMalloc / cudaMallocHost / cudaFree every time ("own them in the shim, */
/* lazily created").  mwf_b200_release_cache() gives everything back.                            */
/* ------------------------------------------------------------------------------------------ */

struct WsEntry { void *p; size_t bytes; int dev; bool host; };
static std::mutex g_ws_mu;
static std::vector<WsEntry> g_ws_free;                 /* oldest first */
static std::unordered_map<void*, WsEntry> g_ws_live;

static void ws_really_free(const WsEntry &e)
{
	if (e.host) cudaFreeHost(e.p);
	else { int cur = 0; cudaGetDevice(&cur); cudaSetDevice(e.dev); cudaFree(e.p); cudaSetDevice(cur); }
}

extern "C" void mwf_b200_release_cache(void)
{
	std::lock_guard<std::mutex> lk(g_ws_mu);
	for (size_t i = 0; i < g_ws_free.size(); ++i) ws_really_free(g_ws_free[i]);
	g_ws_free.clear();
}

static size_t ws_cached_bytes(int dev)
{
	std::lock_guard<std::mutex> lk(g_ws_mu);
	size_t t = 0;
	for (size_t i = 0; i < g_ws_free.size(); ++i)
		if (!g_ws_free[i].host && g_ws_free[i].dev == dev) t += g_ws_free[i].bytes;
	return t;
}

static thread_local bool g_ws_soft = false; /* ws_alloc reports an exhausted device instead of aborting (ws_dev_shrinking) */

/* returns true when the memory is fresh (never used by an earlier batch) */
static bool ws_alloc(void **out, size_t bytes, bool host, int dev)
{
	bytes = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
	{
		std::lock_guard<std::mutex> lk(g_ws_mu);
		int best = -1; /* best fit: a small request must not take the buffer the next, larger one would have matched */
		for (size_t i = 0; i < g_ws_free.size(); ++i) {
			const WsEntry &e = g_ws_free[i];
			if (e.host == host && (host || e.dev == dev) && e.bytes >= bytes && e.bytes <= bytes + bytes / 4 + 65536 &&
			    (best < 0 || e.bytes < g_ws_free[best].bytes)) best = (int)i;
		}
		if (best >= 0) {
			const WsEntry e = g_ws_free[best];
			g_ws_free.erase(g_ws_free.begin() + best);
			g_ws_live[e.p] = e;
			*out = e.p;
			return false;
		}
	}
	void *p = 0;
	cudaError_t err = host ? cudaMallocHost(&p, bytes) : cudaMalloc(&p, bytes);
	if (err == cudaErrorMemoryAllocation) { /* give the cached buffers back and try once more */
		cudaGetLastError();
		mwf_b200_release_cache();
		err = host ? cudaMallocHost(&p, bytes) : cudaMalloc(&p, bytes);
	}
	if (err == cudaErrorMemoryAllocation && g_ws_soft) { cudaGetLastError(); *out = 0; return false; }
	CUDA_OK(err);
	WsEntry e = { p, bytes, dev, host };
	std::lock_guard<std::mutex> lk(g_ws_mu);
	g_ws_live[p] = e;
	*out = p;
	return true;
}

/* Per-device byte cap of the cache: $MWF_B200_CACHE_GB, default half of the device's memory.  After a batch that took nearly
 * all of HBM for its traceback arena the library hands that buffer back to the driver instead of sitting on it, so that other
 * allocators of the process (torch, NCCL, the caller's own) are not starved; mwf_b200_release_cache() frees the rest. */
static size_t ws_cache_cap(void)
{
	const int gb = env_int("MWF_B200_CACHE_GB", -1);
	if (gb >= 0) return (size_t)gb << 30;
	size_t free_b = 0, total_b = 0;
	if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return (size_t)64 << 30; }
	return total_b / 2;
}

static void ws_free(void *p)
{
	if (p == 0) return;
	std::vector<WsEntry> drop;
	{
		std::lock_guard<std::mutex> lk(g_ws_mu);
		std::unordered_map<void*, WsEntry>::iterator it = g_ws_live.find(p);
		if (it == g_ws_live.end()) return;
		const WsEntry e = it->second;
		g_ws_live.erase(it);
		/* the cache is bounded in entries and, per device, in bytes (ws_cache_cap) */
		static const size_t cap = ws_cache_cap();
		if (!e.host && e.bytes > cap) drop.push_back(e);
		else {
			g_ws_free.push_back(e);
			while (g_ws_free.size() > 192) { drop.push_back(g_ws_free.front()); g_ws_free.erase(g_ws_free.begin()); }
			if (!e.host) {
				size_t held = 0;
				for (size_t i = 0; i < g_ws_free.size(); ++i)
					if (!g_ws_free[i].host && g_ws_free[i].dev == e.dev) held += g_ws_free[i].bytes;
				for (size_t i = 0; held > cap && i < g_ws_free.size(); ) { /* oldest first */
					if (g_ws_free[i].host || g_ws_free[i].dev != e.dev) { ++i; continue; }
					held -= g_ws_free[i].bytes;
					drop.push_back(g_ws_free[i]);
					g_ws_free.erase(g_ws_free.begin() + i);
				}
			}
		}
	}
	for (size_t i = 0; i < drop.size(); ++i) ws_really_free(drop[i]);
}

template<class T> static bool ws_dev(T **out, size_t bytes, int dev) { return ws_alloc((void**)out, bytes, false, dev); }

/* The traceback arena is sized from the memory that is free at that moment; another thread or another allocator of the process
 * may take some of it before the cudaMalloc.  Rather than aborting, ask for less: every mode works with a smaller arena (the
 * batch then goes through waves, a rerun with the grown arena, or the segmented traceback).  Returns the bytes obtained. */
static long long ws_dev_shrinking(uint8_t **out, long long want, long long floor_bytes, int dev)
{
	for (long long bytes = want; bytes >= floor_bytes; bytes = (bytes / 2) & ~255LL) {
		g_ws_soft = true;
		ws_alloc((void**)out, (size_t)bytes, false, dev);
		g_ws_soft = false;
		if (*out) return bytes;
		fprintf(stderr, "[mwf_b200] %lld bytes of device memory for the traceback arena are not available (another allocator took them): trying half\n", bytes);
	}
	fprintf(stderr, "[mwf_b200] not enough free device memory for the traceback arena\n");
	abort();
	return 0;
}
template<class T> static bool ws_host(T **out, size_t bytes) { return ws_alloc((void**)out, bytes, true, 0); }

/* an "as much as is free" buffer (the traceback arena when the worst case exceeds the budget): any cached device buffer of at
 * least min_bytes will do -- the free memory differs a little from call to call, and missing the cache by a few MB would mean
 * freeing and reallocating >100 GB.  Returns the size actually obtained. */
static long long ws_dev_flex(uint8_t **out, long long want, long long min_bytes, int dev)
{
	{
		std::lock_guard<std::mutex> lk(g_ws_mu);
		int best = -1;
		for (size_t i = 0; i < g_ws_free.size(); ++i) {
			const WsEntry &e = g_ws_free[i];
			if (!e.host && e.dev == dev && (long long)e.bytes >= min_bytes && (best < 0 || e.bytes > g_ws_free[best].bytes)) best = (int)i;
		}
		if (best >= 0) {
			const WsEntry e = g_ws_free[best];
			g_ws_free.erase(g_ws_free.begin() + best);
			g_ws_live[e.p] = e;
			*out = (uint8_t*)e.p;
			return (long long)e.bytes;
		}
	}
	ws_dev(out, (size_t)want, dev);
	return want;
}

struct mwf_b200_batch {
	int dev, kernel, n, n_sm, threads, n_slots;
	mwf_opt_t opt;
	Pen pen;
	bool is_tb;
	cudaStream_t stream;
	bool own_stream;
	std::vector<int> tl, ql, order;
	std::vector<PairDesc> pairs;
	size_t seq_bytes, cigar_words;
	uint8_t *d_seq, *h_seq;
	PairDesc *d_pairs;
	PairOut *d_outs, *h_outs;
	int *d_order;
	unsigned int *d_ctl; /* [0] queue, [1..2] grid barrier, [4..6] grid flags */
	uint32_t *d_cigar;
	int pitch;
	long long ring_stride, arena_total, rowtab_stride;
	int snap_cap;
	int32_t *d_ring, *d_ring2;
	uint8_t *d_arena;
	long long *d_rowtab, *d_snapoff;
	int *d_snaphdr, *d_seg;
	cudaEvent_t ev0, ev1;
	double kernel_ms;
	int64_t launches, h2d, d2h;
	bool ran, timed; /* timed: kernel_ms of the last run has been taken (mwf_b200_batch_wait may be called more than once) */
	/* tile engine (wfa_tile.cuh) */
	struct TileGeom { int CPT, NT, T, HL, W, grid, grid_p; size_t smem; tile_kernel_fn fn, fn_score; tile_persist_fn pfn, pfn_score; } geom[2]; /* [0] few tiles (latency), [1] many (throughput) */
	int n_geom, tR, wave_pairs, s_limit;
	long long max_len, max_sbound, arena_full; /* arena_full: the arena when everything that is free is taken */
	int *d_nseg;
	uint32_t *d_seqp; /* two- or four-bit packed sequences */
	uint2 *d_seqp2;   /* the same as overlapping pairs of words */
	int *d_packed;
	/* segmented traceback */
	int seg_P;
	int32_t *d_snap;
	long long snap_words;
	SnapDir *d_snapdir;
	int snapdir_stride;
	int *d_nsnap, *d_sstop, *h_nsnap;
	TraceState *d_trace;
	size_t items_cap;
	TileCtl *d_tctl;
	int32_t *d_state, *d_alive;
	int2 *d_items;
	unsigned long long *d_qitems; /* persistent scheduling: the work queue, one 64-bit word per entry (q_word) */
	unsigned int q_mask;
	bool persist;
	int seg_K;             /* traceback segments recomputed at once (virtual slots: seg_K x wave_pairs) */
	double s_est;          /* a high estimate of the largest score of the batch (shared 13-mers), 0 when unknown */
	bool arena_deferred;   /* a few very long pairs: the traceback arena is sized in mwf_b200_batch_run, from the pairs' shared 13-mers */
	unsigned char *d_tmisc; /* TileCounters[2] @0, n_running @32, arena_used @64 */
	int *h_running;
	cudaEvent_t evc[2];
};

static void die(const char *msg)
{
	fprintf(stderr, "[mwf_b200] %s\n", msg);
	abort();
}

static long long gap_cost(const mwf_opt_t *o, long long len)
{
	if (len <= 0) return 0;
	const long long a = o->o1 + len * (long long)o->e1, b = o->o2 + len * (long long)o->e2;
	return a < b ? a : b;
}

/* workspaces of the streaming kernels (ring in HBM): low-memory mode of any size, tiny pairs */
static void alloc_streaming(mwf_b200_batch_t *b)
{
	const mwf_opt_t *opt = &b->opt;
	const int n = b->pen.nring;
	const bool seg = b->is_tb && opt->step > 0;
	const long long max_len = b->max_len, max_sbound = b->max_sbound;
	b->n_slots = 1;
	if (b->kernel == MWF_B200_KERNEL_CTA) {
		/* One CTA per pair in flight.  A band is at most tl + ql + 1 diagonals wide: short pairs (the gap fills of mwf_wfa_chain,
		 * read-sized pairs) get small CTAs, and as many of them per SM as registers allow -- 148 pairs in flight with 512-thread
		 * CTAs leave the GPU waiting on barriers and L2 round trips. */
		if (g_threads <= 0 && getenv("MWF_B200_THREADS") == 0) {
			int t = max_len <= 512 ? 64 : max_len <= 2048 ? 128 : max_len <= 4096 ? 256 : 512;
			while (t < 512 && (long long)b->n_sm * (512 / t) > 2LL * b->n) t *= 2; /* too few pairs for that many CTAs: larger ones */
			b->threads = t;
		}
		int per_sm = 1;
		CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wfa_cta_kernel, b->threads, sizeof(int) * (8 + 2 * (size_t)n)));
		per_sm = std::max(1, std::min(per_sm, env_int("MWF_B200_CTA_PER_SM", 16)));
		b->n_slots = std::max(1, (int)std::min<long long>(b->n, (long long)b->n_sm * per_sm));
	}
	b->pitch = (int)((max_len + 2LL * n + 1 + 24 + 31) & ~31LL);
	b->ring_stride = (long long)n * 5 * b->pitch;
	ws_dev(&b->d_ring, sizeof(int32_t) * b->ring_stride * b->n_slots, b->dev);
	if (b->is_tb) {
		b->rowtab_stride = max_sbound + 2;
		ws_dev(&b->d_rowtab, sizeof(long long) * b->rowtab_stride * b->n_slots, b->dev);
		if (seg) {
			ws_dev(&b->d_ring2, sizeof(int32_t) * b->ring_stride * b->n_slots, b->dev);
			b->snap_cap = (int)(max_sbound / opt->step + 2);
			ws_dev(&b->d_snaphdr, sizeof(int) * (size_t)b->snap_cap * (2 + 2 * n) * b->n_slots, b->dev);
			ws_dev(&b->d_snapoff, sizeof(long long) * (size_t)b->snap_cap * b->n_slots, b->dev);
			ws_dev(&b->d_seg, sizeof(int) * (size_t)b->snap_cap * 2 * b->n_slots, b->dev);
		}
		/* traceback / snapshot arena: the worst case when it is small, else most of what is free */
		size_t free_b = 0, total_b = 0;
		CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
		free_b += ws_cached_bytes(b->dev);
		const long long budget = (long long)((double)free_b * env_int("MWF_B200_ARENA_PCT", 85) / 100.0);
		long long worst = (max_sbound + 2) * (max_len + 16);
		if (seg) worst = std::max(worst, (long long)b->snap_cap * (5LL * n * (max_len + 1) * 4 + 16));
		worst = std::max((worst + 255) & ~255LL, 65536LL);
		long long per_slot = std::min(worst, (budget / b->n_slots) & ~255LL);
		if (per_slot < 4096) die("not enough free device memory for the traceback arena");
		b->arena_total = per_slot * b->n_slots;
		ws_dev(&b->d_arena, (size_t)b->arena_total, b->dev);
	}
}

/* give the tile engine's workspaces back (the low-memory pass did not fit the arena: the streaming kernels take over) */
static void free_tile(mwf_b200_batch_t *b)
{
	ws_free(b->d_tctl); ws_free(b->d_state); ws_free(b->d_alive); ws_free(b->d_items); ws_free(b->d_qitems); ws_free(b->d_tmisc); ws_free(b->d_nseg);
	ws_free(b->d_rowtab); ws_free(b->d_arena); ws_free(b->d_seg);
	b->d_tctl = 0, b->d_state = 0, b->d_alive = 0, b->d_items = 0, b->d_qitems = 0, b->d_tmisc = 0, b->d_nseg = 0, b->d_rowtab = 0, b->d_arena = 0, b->d_seg = 0;
	b->arena_total = 0, b->rowtab_stride = 0;
}

/* The dynamic shared-memory limit is an attribute of the kernel function, process-wide per device: it is raised to the device's
 * opt-in maximum once and never lowered, so that batches with different penalties (different tile sizes) created concurrently
 * on several host threads cannot shrink it under one another between cudaFuncSetAttribute and the launch. */
static void tile_smem_optin(tile_kernel_fn fn, int dev, int smem_optin)
{
	static std::mutex mu;
	static std::vector<std::pair<void*, int> > done;
	std::lock_guard<std::mutex> lk(mu);
	for (size_t i = 0; i < done.size(); ++i)
		if (done[i].first == (void*)fn && done[i].second == dev) return;
	cudaFuncAttributes fa;
	CUDA_OK(cudaFuncGetAttributes(&fa, fn));
	CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - (int)fa.sharedSizeBytes)); /* (static shared memory counts against the limit) */
	done.push_back(std::make_pair((void*)fn, dev));
}

extern "C" mwf_b200_batch_t *mwf_b200_batch_create(const mwf_opt_t *opt, int32_t n_pairs, const int32_t *tl, const int32_t *ql)
{
	if (mwf_b200_device_count() <= 0) die("no CUDA device: this library has no CPU fallback");
	if (opt->x <= 0 || opt->e1 <= 0 || opt->e2 <= 0 || opt->o1 < 0 || opt->o2 < 0)
		die("penalties must satisfy x>0, e1>0, e2>0, o1>=0, o2>=0");
	mwf_b200_batch_t *b = new mwf_b200_batch_t();
	b->dev = tl_device >= 0 ? tl_device : mwf_b200_get_device();
	CUDA_OK(cudaSetDevice(b->dev));
	int smem_optin = 0; /* (cudaGetDeviceProperties costs tens of milliseconds; two attributes are all that is needed) */
	CUDA_OK(cudaDeviceGetAttribute(&b->n_sm, cudaDevAttrMultiProcessorCount, b->dev));
	CUDA_OK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, b->dev));
	b->opt = *opt;
	b->n = n_pairs;
	b->is_tb = !!(opt->flag & MWF_F_CIGAR);
	if (b->is_tb && opt->step > 0) {
		/* The first snapshot of low-memory mode is taken at score opt->step (miniwfa.c:585-586).  When no pair of the batch can
		 * score that much (deleting one sequence and inserting the other bounds the score), pass 1 finds no checkpoint and
		 * pass 2 is the high-memory run: skip pass 1.  (The gap fills of mwf_wfa_chain: ~1e5 pairs of a few dozen bases.) */
		bool reachable = false;
		for (int i = 0; i < n_pairs && !reachable; ++i)
			reachable = gap_cost(opt, tl[i]) + gap_cost(opt, ql[i]) >= opt->step;
		if (!reachable) b->opt.step = 0;
	}
	opt = &b->opt;
	const bool seg = b->is_tb && opt->step > 0;
	int max_pen = opt->x;
	max_pen = std::max(max_pen, opt->o1 + opt->e1);
	max_pen = std::max(max_pen, opt->o2 + opt->e2);
	b->pen.x = opt->x, b->pen.oe1 = opt->o1 + opt->e1, b->pen.e1 = opt->e1;
	b->pen.oe2 = opt->o2 + opt->e2, b->pen.e2 = opt->e2, b->pen.nring = max_pen + 1;
	b->tl.assign(tl, tl + n_pairs);
	b->ql.assign(ql, ql + n_pairs);
	b->threads = g_threads > 0 ? (int)g_threads : env_int("MWF_B200_THREADS", 512);
	if (b->threads < 64 || b->threads > 512 || b->threads % 32) die("block threads must be a multiple of 32 in [64,512]");

	/* layout of sequences, CIGAR buffers, work order */
	long long max_len = 0, max_sbound = 0;
	size_t off = 256, cw = 0; /* (front slack: the probe of a cell off the matrix reads position d < 0 of its query, never used) */
	b->pairs.resize(n_pairs);
	b->order.resize(n_pairs);
	for (int i = 0; i < n_pairs; ++i) {
		if (tl[i] < 0 || ql[i] < 0) die("negative sequence length");
		PairDesc &p = b->pairs[i];
		p.tl = tl[i], p.ql = ql[i];
		p.t_off = (long long)off; off += ((size_t)tl[i] + 64 + 15) & ~(size_t)15;
		p.q_off = (long long)off; off += ((size_t)ql[i] + 64 + 15) & ~(size_t)15;
		p.cigar_off = (long long)cw, p.cigar_cap = b->is_tb ? tl[i] + ql[i] + 2 : 0, p.pad_ = 0;
		cw += (size_t)p.cigar_cap;
		max_len = std::max(max_len, (long long)tl[i] + ql[i]);
		long long sb = gap_cost(opt, tl[i]) + gap_cost(opt, ql[i]);
		if (opt->max_s > 0 && !seg) sb = std::min(sb, (long long)opt->max_s + 1); /* (pass 1 of low-memory mode has no stop tests) */
		max_sbound = std::max(max_sbound, sb);
		b->order[i] = i;
	}
	if (max_len >= 1024) /* longest first, for balance; pairs this short take microseconds each and sorting 1e5 of them costs more than it saves */
		std::stable_sort(b->order.begin(), b->order.end(), [&](int a, int c) {
			return (long long)tl[a] + ql[a] > (long long)tl[c] + ql[c]; });
	b->seq_bytes = off + 64, b->cigar_words = cw;
	b->s_limit = (int)std::min<long long>(max_sbound + 1, 0x7ffffff0);
	b->max_len = max_len, b->max_sbound = max_sbound, b->arena_full = 0, b->d_nseg = 0, b->d_seqp = 0, b->d_seqp2 = 0, b->d_packed = 0;
	b->seg_P = 0, b->d_snap = 0, b->snap_words = 0, b->d_snapdir = 0, b->snapdir_stride = 0, b->d_nsnap = 0, b->d_sstop = 0, b->h_nsnap = 0, b->d_trace = 0;

	/* kernel family */
	int pref = pick_kernel_pref();
	const int n = b->pen.nring;
	/* tile geometries.  [1] throughput: 4 cells per thread, 128 threads, blocks of 32 scores (lowest instruction count per
	 * cell, most CTAs per SM) -- used while a launch has enough tiles to fill the GPU.  [0] latency: 1 cell per thread, 512 threads
	 * on the same 512-wide tile, blocks of 64 scores -- used while there are few tiles (single pairs, narrow bands): the GPU is
	 * not full anyway and the dependent chain of one score step is shorter.  The state layout does not depend on the geometry,
	 * so run_tile switches between them from launch to launch.  The environment forces a single geometry (tests, sweeps). */
	b->tR = n + 2 * (opt->e1 + 1) + 2 * (opt->e2 + 1);
	const bool forced = getenv("MWF_B200_TILE_CPT") || getenv("MWF_B200_TILE_T") || getenv("MWF_B200_TILE_THREADS");
	b->n_geom = forced ? 1 : 2;
	bool tile_ok = n <= TILE_NRING_MAX && opt->e1 < TILE_EDEPTH_MAX && opt->e2 < TILE_EDEPTH_MAX;
	for (int g = 0; g < b->n_geom; ++g) {
		mwf_b200_batch::TileGeom &G = b->geom[g];
		const bool lat = forced ? n_pairs < 16 : g == 0;
		/* (with the default gap extensions the interior step keeps the gap rows in registers -- tile_cells_fast2 -- and 2 cells per
		 * thread on 256 threads make the shortest step of a lone tile: 150 kb pair 25.6 ms against 29.0 with 1 x 512) */
		const bool lean = opt->e1 == 2 && opt->e2 == 1 && env_int("MWF_B200_TILE_FAST", 2) >= 2;
		G.CPT = env_int("MWF_B200_TILE_CPT", lat ? (lean ? 2 : 1) : 4);
		if (G.CPT != 1 && G.CPT != 2 && G.CPT != 4) die("MWF_B200_TILE_CPT must be 1, 2 or 4");
		G.T = std::max(4, std::min(env_int("MWF_B200_TILE_T", lat ? 64 : 32) & ~3, TILE_TMAX));
		G.HL = G.T;
		G.NT = env_int("MWF_B200_TILE_THREADS", lat ? (lean ? 256 : 512) : 128);
		G.W = G.CPT * G.NT;
		G.smem = (size_t)b->tR * G.W * 4 + 64 + sizeof(StepTab) * (size_t)std::max(G.T, 2) + 2 * XCH_BUF;
		G.fn = 0, G.fn_score = 0, G.grid = 0;
		tile_ok = tile_ok && G.NT % 32 == 0 && G.NT >= 64 && G.NT <= TILE_MAX_THREADS(G.CPT) && G.W % 4 == 0 && G.smem + 1024 <= (size_t)smem_optin &&
			(G.W - 2 * G.HL) / 2 - 4 >= 2 * G.HL + n + 8;
	}
	int umax = b->geom[0].W - 2 * b->geom[0].HL, wmax = b->geom[0].W;
	if (b->n_geom > 1) umax = std::min(umax, b->geom[1].W - 2 * b->geom[1].HL), wmax = std::max(wmax, b->geom[1].W);
	if (pref == MWF_B200_KERNEL_TILE && !tile_ok) pref = MWF_B200_KERNEL_AUTO;
	if (pref == MWF_B200_KERNEL_AUTO) {
		if (tile_ok && max_len >= env_int("MWF_B200_TILE_MINLEN", 8192)) pref = MWF_B200_KERNEL_TILE;
		else pref = (n_pairs >= b->n_sm / 4 || max_len < 32768) ? MWF_B200_KERNEL_CTA : MWF_B200_KERNEL_GRID;
	}
	b->kernel = pref;
	b->n_slots = pref == MWF_B200_KERNEL_GRID ? 1 : std::max(1, std::min(n_pairs, b->n_sm));

	CUDA_OK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
	b->own_stream = true;
	CUDA_OK(cudaEventCreate(&b->ev0));
	CUDA_OK(cudaEventCreate(&b->ev1));

	/* device workspaces */
	ws_dev(&b->d_seq, b->seq_bytes, b->dev);
	if (ws_host(&b->h_seq, b->seq_bytes)) memset(b->h_seq, 0, b->seq_bytes); /* padding bytes are never interpreted: runs are clamped */
	ws_dev(&b->d_pairs, sizeof(PairDesc) * std::max(1, n_pairs), b->dev);
	ws_dev(&b->d_outs, sizeof(PairOut) * std::max(1, n_pairs), b->dev);
	ws_host(&b->h_outs, sizeof(PairOut) * std::max(1, n_pairs));
	ws_dev(&b->d_order, sizeof(int) * std::max(1, n_pairs), b->dev);
	ws_dev(&b->d_ctl, 64, b->dev);
	b->d_ring = 0, b->d_ring2 = 0, b->d_arena = 0, b->d_rowtab = 0, b->d_snapoff = 0, b->d_snaphdr = 0, b->d_seg = 0, b->d_cigar = 0;
	b->d_tctl = 0, b->d_state = 0, b->d_alive = 0, b->d_items = 0, b->d_qitems = 0, b->q_mask = 0, b->persist = false, b->arena_deferred = false, b->s_est = 0, b->seg_K = 1, b->d_tmisc = 0, b->h_running = 0;
	b->arena_total = 0, b->rowtab_stride = 0, b->snap_cap = 0, b->wave_pairs = 0;
	if (b->is_tb) ws_dev(&b->d_cigar, sizeof(uint32_t) * std::max<size_t>(1, cw), b->dev);
	if (pref == MWF_B200_KERNEL_TILE) {
		/* state: two buffers of R rows per pair in flight; pairs beyond the memory budget run in later waves */
		b->pitch = (int)((max_len + 2LL * n + 2LL * TILE_TMAX + 32 + wmax + 31) & ~31LL);
		b->rowtab_stride = b->is_tb ? max_sbound + 2 : 0;
		const size_t per_pair = (size_t)b->pitch * 4 * (2 * b->tR + 1) + sizeof(TileCtl) + (size_t)b->rowtab_stride * 8 +
			sizeof(int2) * ((size_t)b->pitch / umax + 2);
		size_t free_b = 0, total_b = 0;
		CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
		free_b += ws_cached_bytes(b->dev); /* cached workspaces are reused or given back on demand */
		const double frac = b->is_tb ? 0.35 : 0.85;
		b->wave_pairs = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, n_pairs), (size_t)(free_b * frac) / per_pair));
		if (b->is_tb && n_pairs > 1) { /* ... and few enough that their expected traceback bytes (s ~ 0.3 n) fit the arena in one go */
			const double len = 0.5 * (double)max_len, expect = 0.09 * len * len + 4096.0 * len;
			const double room = (double)free_b * 0.5;
			b->wave_pairs = (int)std::max(1.0, std::min((double)b->wave_pairs, room / expect));
		}
		if (env_int("MWF_B200_TILE_WAVE", 0) > 0) b->wave_pairs = std::min(b->wave_pairs, env_int("MWF_B200_TILE_WAVE", 0)); /* tests */
		const int wp = b->wave_pairs;
		if (ws_dev(&b->d_state, (size_t)wp * 2 * b->tR * b->pitch * 4, b->dev)) /* fresh memory: a large negative int32 everywhere */
			CUDA_OK(cudaMemsetAsync(b->d_state, 0xC0, (size_t)wp * 2 * b->tR * b->pitch * 4, b->stream));
		ws_dev(&b->d_alive, (size_t)wp * b->pitch * 4, b->dev);
		ws_dev(&b->d_tctl, sizeof(TileCtl) * wp, b->dev);
		b->items_cap = (size_t)wp * ((size_t)b->pitch / umax + 2);
		ws_dev(&b->d_items, sizeof(int2) * b->items_cap, b->dev);
		ws_dev(&b->d_tmisc, 256, b->dev);
		if (env_int("MWF_B200_TILE_PACK", 1)) {
			ws_dev(&b->d_seqp, b->seq_bytes / 2 + (size_t)max_len / 4 + 512, b->dev); /* (slack at the end: the clamped probes of cells off the matrix read up to tl codes past a query) */
			ws_dev(&b->d_packed, sizeof(int) * std::max(1, n_pairs), b->dev);
			if (env_int("MWF_B200_TILE_FAST", 2) >= 2 && (long long)b->seq_bytes < (1LL << 28))
				ws_dev(&b->d_seqp2, 2 * (b->seq_bytes / 2 + (size_t)max_len / 4 + 512), b->dev);
		}
		ws_host(&b->h_running, 2 * 256);
		CUDA_OK(cudaEventCreateWithFlags(&b->evc[0], cudaEventDisableTiming));
		CUDA_OK(cudaEventCreateWithFlags(&b->evc[1], cudaEventDisableTiming));
		if (seg) { /* low-memory mode: checkpoints found by walking a high-memory pass (wfa_tile_checkpoint_kernel) */
			b->snap_cap = (int)(max_sbound / opt->step + 2);
			ws_dev(&b->d_seg, sizeof(int) * (size_t)b->snap_cap * 2 * wp, b->dev);
			ws_dev(&b->d_nseg, sizeof(int) * wp, b->dev);
		}
		if (b->is_tb) {
			ws_dev(&b->d_rowtab, sizeof(long long) * b->rowtab_stride * wp, b->dev);
			CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
			free_b += ws_cached_bytes(b->dev);
			const long long budget = (long long)((double)free_b * env_int("MWF_B200_ARENA_PCT", 85) / 100.0);
			long long worst = (max_sbound + 2) * (max_len + 2LL * n + 2LL * TILE_TMAX + 16);
			worst = std::max((worst + 255) & ~255LL, 65536LL);
			b->arena_full = (long long)std::min((double)budget, (double)worst * wp) & ~255LL;
			/* first try: what pairs of up to ~6 % divergence need (s ~ 0.3 n, s^2 bytes); a batch that needs more is rerun with all
			 * that is free, and beyond that with the segmented traceback.  Asking for 150 GB up front costs seconds in cudaMalloc. */
			double expect = 0;
			for (int i = 0; i < std::min(wp, n_pairs); ++i) {
				const double len = std::max(tl[b->order[i]], ql[b->order[i]]);
				double e = 0.09 * len * len + 4096.0 * len;
				/* a max_iter budget (mwf_wfa_auto's first attempt: 1e8 cells) bounds the bytes too: the run stops within one block
				 * of scores of the budget (pass 1 of low-memory mode has no stop tests) */
				if (opt->max_iter > 0 && !seg)
					e = std::min(e, (double)opt->max_iter + (TILE_TMAX + 2.0) * ((double)tl[b->order[i]] + ql[b->order[i]] + 2.0 * n + 2.0 * TILE_TMAX + 16));
				expect += e;
			}
			b->arena_total = (long long)std::min((double)b->arena_full, std::max(expect, 64.0 * 1048576)) & ~255LL;
			if (env_int("MWF_B200_TILE_ARENA_MAX", 0) > 0) /* tests */
				b->arena_total = b->arena_full = std::min<long long>(b->arena_total, env_int("MWF_B200_TILE_ARENA_MAX", 0));
			if (b->arena_total < 4096) die("not enough free device memory for the traceback arena");
			/* a few very long pairs whose expected arena is everything that is free: wait for the sequences (mwf_b200_batch_run
			 * estimates the score from the shared 13-mers, 3 ms) instead of taking ~150 GB that the run may never touch */
			b->arena_deferred = b->arena_total >= b->arena_full && env_int("MWF_B200_TILE_PREDICT", 1) && n_pairs <= 8 &&
			                    max_len >= env_int("MWF_B200_TILE_PREDICT_MINLEN", 2000000) && !env_int("MWF_B200_TILE_ARENA_MAX", 0);
			if (!b->arena_deferred) {
				b->arena_total = ws_dev_shrinking(&b->d_arena, b->arena_total, 4096, b->dev);
				b->arena_full = std::max(b->arena_full, b->arena_total);
			}
		}
		for (int g = 0; g < b->n_geom; ++g) {
			mwf_b200_batch::TileGeom &G = b->geom[g];
			int per_sm = 0;
			G.fn = tile_kernel_for(b->is_tb, G.CPT);
			G.fn_score = tile_kernel_for(false, G.CPT);
			tile_smem_optin(G.fn, b->dev, smem_optin);
			tile_smem_optin(G.fn_score, b->dev, smem_optin);
			CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, G.fn, G.NT, G.smem));
			if (per_sm < 1) die("tile kernel does not fit on an SM");
			G.grid = b->n_sm * std::min(per_sm, env_int("MWF_B200_TILE_CTAS_PER_SM", 8));
			G.pfn = tile_persist_for(b->is_tb, G.CPT);
			G.pfn_score = tile_persist_for(false, G.CPT);
			tile_smem_optin((tile_kernel_fn)(void*)G.pfn, b->dev, smem_optin);
			tile_smem_optin((tile_kernel_fn)(void*)G.pfn_score, b->dev, smem_optin);
			CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, G.pfn, G.NT, G.smem));
			if (per_sm < 1) die("tile kernel does not fit on an SM");
			G.grid_p = b->n_sm * std::min(per_sm, env_int("MWF_B200_TILE_CTAS_PER_SM", 8));
		}
		b->persist = env_int("MWF_B200_TILE_PERSIST", 1) != 0;
		if (b->persist) { /* the work queue: every tile in flight, one plan item per pair, one retire item per CTA, with room to spare */
			size_t need = 2 * (b->items_cap + (size_t)wp + (size_t)std::max(b->geom[0].grid_p, b->geom[b->n_geom - 1].grid_p)) + 4096, cap = 1;
			while (cap < need) cap <<= 1;
			b->q_mask = (unsigned int)(cap - 1);
			ws_dev(&b->d_qitems, sizeof(unsigned long long) * cap, b->dev);
		}
	} else alloc_streaming(b);
	b->kernel_ms = 0, b->launches = 0, b->h2d = 0, b->d2h = 0, b->ran = false, b->timed = false;
	return b;
}

extern "C" void mwf_b200_batch_set_stream(mwf_b200_batch_t *b, void *cuda_stream)
{
	if (b->own_stream) { CUDA_OK(cudaStreamSynchronize(b->stream)); CUDA_OK(cudaStreamDestroy(b->stream)); b->own_stream = false; }
	b->stream = (cudaStream_t)cuda_stream;
}

extern "C" void mwf_b200_batch_upload(mwf_b200_batch_t *b, const char *const *ts, const char *const *qs)
{
	CUDA_OK(cudaSetDevice(b->dev));
	/* staging into pinned memory: a few host threads when there is enough to copy (one core moves ~10 GB/s: 25 MB of a config-3
	 * batch are 2.5 ms of the call) */
	const int n_thr = b->seq_bytes >= (4u << 20) && b->n >= 8 ? std::min(4, env_int("MWF_B200_STAGE_THREADS", 4)) : 1;
	auto stage = [&](int first, int step) {
		for (int i = first; i < b->n; i += step) {
			const PairDesc &p = b->pairs[i];
			if (p.tl) memcpy(b->h_seq + p.t_off, ts[i], (size_t)p.tl);
			if (p.ql) memcpy(b->h_seq + p.q_off, qs[i], (size_t)p.ql);
		}
	};
	if (n_thr > 1) {
		std::vector<std::thread> th;
		for (int t = 1; t < n_thr; ++t) th.push_back(std::thread(stage, t, n_thr));
		stage(0, n_thr);
		for (size_t t = 0; t < th.size(); ++t) th[t].join();
	} else stage(0, 1);
	CUDA_OK(cudaMemcpyAsync(b->d_seq, b->h_seq, b->seq_bytes, cudaMemcpyHostToDevice, b->stream));
	CUDA_OK(cudaMemcpyAsync(b->d_pairs, b->pairs.data(), sizeof(PairDesc) * b->n, cudaMemcpyHostToDevice, b->stream));
	CUDA_OK(cudaMemcpyAsync(b->d_order, b->order.data(), sizeof(int) * b->n, cudaMemcpyHostToDevice, b->stream));
	b->h2d = (int64_t)b->seq_bytes + (int64_t)(sizeof(PairDesc) + sizeof(int)) * b->n;
	if (b->d_seqp && b->n > 0) { /* two- or four-bit copies for the tile engine's match-run probes */
		wfa_pack_kernel<<<b->n, 256, 0, b->stream>>>(b->d_seq, b->d_pairs, b->d_seqp, b->d_seqp2, b->d_packed);
		CUDA_OK(cudaGetLastError());
	}
}

static KParams make_params(const mwf_b200_batch_t *b, int n_slots)
{
	KParams P;
	memset(&P, 0, sizeof(P));
	P.pen = b->pen;
	P.is_tb = b->is_tb, P.step = b->is_tb ? b->opt.step : 0, P.max_s = b->opt.max_s, P.max_iter = b->opt.max_iter;
	P.n_pairs = b->n, P.single_pair = 0;
	P.order = b->d_order, P.queue = b->d_ctl;
	P.pairs = b->d_pairs, P.outs = b->d_outs, P.seq = b->d_seq, P.cigar = b->d_cigar;
	P.pitch = b->pitch;
	P.ring = b->d_ring, P.ring2 = b->d_ring2, P.ring_stride = b->ring_stride;
	P.arena = b->d_arena, P.arena_stride = b->d_arena ? (b->arena_total / n_slots) & ~255LL : 0;
	P.rowtab = b->d_rowtab, P.rowtab_stride = b->rowtab_stride;
	P.snaphdr = b->d_snaphdr, P.snapoff = b->d_snapoff, P.seg = b->d_seg, P.snap_cap = b->snap_cap;
	P.bar = b->d_ctl + 1, P.gflags = (int*)(b->d_ctl + 4);
	return P;
}

static void launch_cta(mwf_b200_batch_t *b, const KParams &P, int grid)
{
	const size_t smem = sizeof(int) * (8 + 2 * (size_t)b->pen.nring);
	CUDA_OK(cudaMemsetAsync(b->d_ctl, 0, 64, b->stream));
	wfa_cta_kernel<<<grid, b->threads, smem, b->stream>>>(P);
	CUDA_OK(cudaGetLastError());
	++b->launches;
}

static void launch_grid(mwf_b200_batch_t *b, KParams P, int pair)
{
	const size_t smem = sizeof(int) * (8 + 2 * (size_t)b->pen.nring);
	int per_sm = 0;
	CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wfa_grid_kernel, b->threads, smem));
	if (per_sm < 1) die("grid kernel does not fit on an SM");
	const int grid = b->n_sm * std::min(per_sm, env_int("MWF_B200_GRID_CTAS_PER_SM", 1));
	P.single_pair = pair;
	CUDA_OK(cudaMemsetAsync(b->d_ctl, 0, 64, b->stream));
	void *args[] = { (void*)&P };
	CUDA_OK(cudaLaunchCooperativeKernel((void*)wfa_grid_kernel, dim3(grid), dim3(b->threads), args, smem, b->stream));
	++b->launches;
}

/* one pass of the tile engine over a wave of pairs: score 0, then plan + tile kernels per time block until every pair has ended.
 * The number of running pairs and of tiles per launch is read back one chunk of launches behind, so the device never waits for
 * the host; the tile count picks the geometry of the next chunk. */
/* returns the error bits the planner raised during the pass: 1 << TS_ARENA, 1 << TS_SHRINK */
static int tile_pass_persist(mwf_b200_batch_t *b, const TParams *PP, int np, int seg_j, bool score_kernel, int group);

/* group > 1 (persistent kernel only): segments seg_j, seg_j - 1, ..., seg_j - group + 1 of every pair are recomputed in one pass,
 * each on its own virtual slot (state, TileCtl, alive words), so that a few very long pairs keep the GPU full */
static int tile_pass(mwf_b200_batch_t *b, const TParams *PP, int np, int seg_j = -1, bool score_kernel = false, int group = 1)
{
	if (b->persist && (long long)np * group < 0x3ffffe && b->pitch / 256 < 0x3ffffe) /* (slot and tile numbers are 22-bit fields of a queue word) */
		return tile_pass_persist(b, PP, np, seg_j, score_kernel, group);
	int err = 0;
	const int chunk_len = std::max(1, env_int("MWF_B200_TILE_CHUNK", 8));
	const unsigned int many = (unsigned int)env_int("MWF_B200_TILE_SWITCH", 2 * b->n_sm);
	CUDA_OK(cudaMemsetAsync(b->d_tmisc, 0, 256, b->stream));
	CUDA_OK(cudaMemsetAsync(b->d_alive, 0, (size_t)np * b->pitch * 4, b->stream));
	if (seg_j <= 0) { wfa_tile_init_kernel<<<np, 128, 0, b->stream>>>(PP[0], 0); ++b->launches; }
	if (seg_j >= 0) { wfa_tile_segstart_kernel<<<dim3(np, seg_j > 0 ? 64 : 1), 256, 0, b->stream>>>(PP[0], seg_j); ++b->launches; }
	CUDA_OK(cudaGetLastError());
	int it = 0, g = b->n_geom > 1 && np >= 16 ? 1 : 0;
	for (int chunk = 0;; ++chunk) {
		const TParams &P = PP[g];
		const mwf_b200_batch::TileGeom &G = b->geom[g];
		for (int k = 0; k < chunk_len; ++k, ++it) {
			wfa_plan_kernel<<<np, 128, 0, b->stream>>>(P, it);
			(score_kernel ? G.fn_score : G.fn)<<<G.grid, G.NT, G.smem, b->stream>>>(P, it);
			b->launches += 2;
		}
		CUDA_OK(cudaGetLastError());
		if (env_int("MWF_B200_DEBUG", 0)) {
			TileCtl h;
			CUDA_OK(cudaStreamSynchronize(b->stream));
			CUDA_OK(cudaMemcpy(&h, b->d_tctl, sizeof(h), cudaMemcpyDeviceToHost));
			fprintf(stderr, "[tile dbg] it=%d geom=%d status=%d s=%d band=[%d,%d] cur=%d n_iter=%lld Tb=%d A4=%d total4=%d n_tiles=%d done_t=%d fin=[%d,%d] lo0=%d hi0=%d sid=%d\n",
			        it, g, h.status, h.s, h.wflo, h.wfhi, h.cur, h.n_iter, h.Tb, h.A4, h.total4, h.n_tiles, h.done_t, h.fin_lo, h.fin_hi, h.lo_log[0], h.hi_log[0], h.sid);
		}
		int *hr = b->h_running + 16 * (chunk & 1); /* TileCounters[2] @0, n_running @32 */
		CUDA_OK(cudaMemcpyAsync(hr, b->d_tmisc, 64, cudaMemcpyDeviceToHost, b->stream));
		CUDA_OK(cudaEventRecord(b->evc[chunk & 1], b->stream));
		if (chunk >= 1) {
			const int *pr = b->h_running + 16 * ((chunk - 1) & 1);
			CUDA_OK(cudaEventSynchronize(b->evc[(chunk - 1) & 1]));
			if (pr[8] == 0) { err = pr[10]; break; }
			if (b->n_geom > 1) g = std::max((unsigned int)pr[0], (unsigned int)pr[2]) >= many ? 1 : 0;
		}
	}
	return err;
}

/* the same pass with the persistent kernel: one launch per tile geometry in use instead of two per block of scores.  Batches
 * run the throughput geometry throughout; a few large pairs start in the latency geometry, and the planner asks for the other
 * one when the number of tiles in flight crosses `many` (back below many / 2): the pairs park after their block in flight, the
 * kernel retires, and the host launches the other geometry. */
static int tile_pass_persist(mwf_b200_batch_t *b, const TParams *PP, int np_pairs, int seg_j, bool score_kernel, int group)
{
	const int np = np_pairs * group, vmod = group > 1 ? np_pairs : 0; /* np: (virtual) slots of this pass */
	TParams P0 = PP[0];
	P0.n_pairs = np, P0.vmod = vmod;
	CUDA_OK(cudaMemsetAsync(b->d_tmisc, 0, 256, b->stream));
	CUDA_OK(cudaMemsetAsync(b->d_alive, 0, (size_t)np * b->pitch * 4, b->stream));
	if (seg_j < 0) { wfa_tile_init_kernel<<<np_pairs, 128, 0, b->stream>>>(P0, 0); ++b->launches; }
	else if (seg_j < group) { wfa_tile_init_kernel<<<np_pairs, 128, 0, b->stream>>>(P0, seg_j * np_pairs); ++b->launches; } /* the slots of segment 0 */
	if (seg_j >= 0) { wfa_tile_segstart_kernel<<<dim3(np, seg_j > 0 ? 64 : 1), 256, 0, b->stream>>>(P0, seg_j); ++b->launches; }
	CUDA_OK(cudaGetLastError());
	const bool may_switch = b->n_geom > 1 && np < 16;
	int g = b->n_geom > 1 && np >= 16 ? 1 : 0;
	for (int round = 0;; ++round) {
		TParams P = PP[g];
		const mwf_b200_batch::TileGeom &G = b->geom[g];
		P.n_pairs = np, P.vmod = vmod;
		P.geom_id = g, P.n_geom = may_switch ? 2 : 1, P.many = env_int("MWF_B200_TILE_SWITCH", 2 * b->n_sm);
		CUDA_OK(cudaMemsetAsync(b->d_qitems, 0, sizeof(unsigned long long) * ((size_t)b->q_mask + 1), b->stream));
		wfa_tile_persist_begin_kernel<<<1, 256, 0, b->stream>>>(P, G.grid_p);
		(score_kernel ? G.pfn_score : G.pfn)<<<G.grid_p, G.NT, G.smem, b->stream>>>(P);
		b->launches += 2;
		CUDA_OK(cudaGetLastError());
		int *hr = b->h_running;
		CUDA_OK(cudaMemcpyAsync(hr, b->d_tmisc, 192, cudaMemcpyDeviceToHost, b->stream));
		CUDA_OK(cudaStreamSynchronize(b->stream));
		const PersistCtl *pq = (const PersistCtl*)(hr + 24);
		if (env_int("MWF_B200_DEBUG", 0))
		{
			TileCtl h;
			CUDA_OK(cudaMemcpy(&h, b->d_tctl, sizeof(h), cudaMemcpyDeviceToHost));
			fprintf(stderr, "[persist dbg] round %d geometry %d: running %d, err %d, head %u tail %u, stop %d -> geometry %d, tiles in flight %d; pair 0: s %d band [%d, %d] last shrink at %d\n",
			        round, g, hr[8], hr[10], pq->head, pq->tail, pq->stop_req, pq->switch_to, pq->total_tiles, h.s, h.wflo, h.wfhi, h.shrink_s);
		}
#ifdef MWF_PHASE_PROF
		{
			unsigned long long ph[16], z[16] = {0};
			CUDA_OK(cudaMemcpyFromSymbol(ph, g_phase, sizeof(ph)));
			CUDA_OK(cudaMemcpyToSymbol(g_phase, z, sizeof(z)));
			double tot = 0;
			for (int k = 0; k < 14; ++k) tot += (double)ph[k];
			static const char *nm[16] = {"take", "setup-after-issue", "load-wait", "steps", "post-commit..return", "atom-done", "plan", "steps(special tiles)", "store:fence", "store:issue", "store:commit", "ctl-loads", "load-issue", "wait_group0+fence", "#interior", "#special"};
			fprintf(stderr, "[phase] geometry %d grid %d (cycles per CTA %.3g):", g, G.grid_p, tot / G.grid_p);
			for (int k = 0; k < 14; ++k) if (ph[k]) fprintf(stderr, " %s %.1f%%", nm[k], 100 * ph[k] / tot);
			fprintf(stderr, " | tiles: interior %llu special %llu", ph[14], ph[15]);
			fprintf(stderr, "\n");
		}
#endif
		if (hr[8] == 0) return hr[10]; /* n_running, err */
		if (!may_switch || !pq->stop_req) die("internal error: the persistent tile kernel retired with pairs still running");
		g = pq->switch_to;
	}
}

/* kernel parameters of the tile engine, one set per geometry */
static void tile_params(mwf_b200_batch_t *b, TParams *PP)
{
	TParams P;
	memset(&P, 0, sizeof(P));
	P.pen = b->pen, P.is_tb = b->is_tb, P.max_s = b->opt.max_s, P.max_iter = b->opt.max_iter;
	P.order = b->d_order, P.pairs = b->d_pairs, P.outs = b->d_outs, P.seq = b->d_seq, P.seqp = b->d_seqp, P.seqp2 = b->d_seqp2, P.packed = b->d_packed, P.cigar = b->d_cigar;
	P.ctl = b->d_tctl, P.state = b->d_state, P.alive = b->d_alive;
	P.pitch = b->pitch, P.R = b->tR;
	P.pq = (PersistCtl*)(b->d_tmisc + 96), P.q_items = b->d_qitems, P.q_mask = b->q_mask, P.q_bits = (unsigned int)__builtin_popcount(b->q_mask);
	P.items = b->d_items, P.cnt = (TileCounters*)b->d_tmisc, P.n_running = (int*)(b->d_tmisc + 32), P.err = (int*)(b->d_tmisc + 40);
	P.arena = b->d_arena, P.arena_cap = b->arena_total, P.arena_used = (unsigned long long*)(b->d_tmisc + 64);
	P.rowtab = b->d_rowtab, P.rowtab_stride = b->rowtab_stride;
	P.s_limit = b->s_limit;
	/* (the fast path addresses the sequences by 32-bit bit positions inside the sequence buffer) */
	P.fast = env_int("MWF_B200_TILE_FAST", 2) * (int)(b->pen.e1 <= 2 && b->pen.e2 <= 2 && (long long)b->seq_bytes < (1LL << 28));
	if (P.fast >= 2 && !b->d_seqp2) P.fast = 1;
	P.fast_edge = env_int("MWF_B200_TILE_FASTEDGE", 1);
	P.seg = b->d_seg, P.n_seg = b->d_nseg, P.seg_stride = 2 * b->snap_cap, P.step = b->opt.step;
	for (int g = 0; g < 2; ++g) { /* per geometry: tile width, block length, row tables */
		const mwf_b200_batch::TileGeom &G = b->geom[g < b->n_geom ? g : 0];
		const int n = b->pen.nring, d1 = b->pen.e1 + 1, d2 = b->pen.e2 + 1, rb = G.W * 4;
		const int bE1 = n, bF1 = bE1 + d1, bE2 = bF1 + d1, bF2 = bE2 + d2;
		TParams &Q = PP[g];
		Q = P;
		Q.W = G.W, Q.HL = G.HL, Q.T = G.T;
		for (int h = 0; h < n; ++h)
			Q.tabH[h] = make_int4(((h - b->pen.x + n) % n) * rb, ((h - b->pen.oe1 + n) % n) * rb, ((h - b->pen.oe2 + n) % n) * rb, h * rb);
		for (int e = 0; e < d1; ++e) {
			const int pe = (e - b->pen.e1 + d1) % d1;
			Q.tabE1[e] = make_int4((bE1 + pe) * rb, (bF1 + pe) * rb, (bE1 + e) * rb, (bF1 + e) * rb);
		}
		for (int e = 0; e < d2; ++e) {
			const int pe = (e - b->pen.e2 + d2) % d2;
			Q.tabE2[e] = make_int4((bE2 + pe) * rb, (bF2 + pe) * rb, (bE2 + e) * rb, (bF2 + e) * rb);
		}
	}
}

/* the tile engine over all waves of a batch; false when a low-memory request did not fit the arena (nothing is lost: the
 * caller reruns the batch on the streaming kernels) */
static bool run_tile(mwf_b200_batch_t *b)
{
	const bool lowmem = b->is_tb && b->opt.step > 0;
	TParams PP[2];
	tile_params(b, PP);
	for (int p0 = 0; p0 < b->n; p0 += b->wave_pairs) {
		const int np = std::min(b->wave_pairs, b->n - p0);
		for (int g = 0; g < 2; ++g) PP[g].pair0 = p0, PP[g].n_pairs = np;
		if (lowmem) {
			/* low-memory mode (miniwfa.c:603-615).  Pass 1 of the reference only serves to find the checkpoints; here they come
			 * from an unbanded high-memory pass (no stop tests, like mwf_wfa_seg) whose traceback bytes are walked backwards. */
			for (int g = 0; g < 2; ++g) PP[g].seg_use = 0, PP[g].max_s = 0, PP[g].max_iter = 0;
			const int err = tile_pass(b, PP, np);
			if (err & (1 << TS_SHRINK)) die("internal error: empty band after shrink");
			if (err & (1 << TS_ARENA)) return false; /* s^2 bytes do not fit: the caller falls back */
			wfa_tile_checkpoint_kernel<<<np, 32, 0, b->stream>>>(PP[0]);
			CUDA_OK(cudaGetLastError());
			++b->launches;
			for (int g = 0; g < 2; ++g) PP[g].seg_use = 1, PP[g].max_s = b->opt.max_s, PP[g].max_iter = b->opt.max_iter; /* pass 2: mwf_wfa_core with the checkpoints */
		}
		const int err2 = tile_pass(b, PP, np);
		if (err2 & (1 << TS_SHRINK)) die("internal error: empty band after shrink");
		if ((err2 & (1 << TS_ARENA)) && b->is_tb) return false; /* the caller grows the arena, or goes segmented */
		if (b->is_tb) {
			wfa_tile_traceback_kernel<<<np, 32, 0, b->stream>>>(PP[0]);
			CUDA_OK(cudaGetLastError());
			++b->launches;
		}
	}
	return true;
}

/* High-memory CIGAR whose s^2 traceback bytes do not fit the arena (or MWF_B200_TILE_SEGP set): segmented traceback.
 * A score-only forward pass saves the ring state every seg_P scores; then, from the last segment to the first, the state is
 * restored, seg_P scores are recomputed with traceback bytes and the traceback warp walks that segment.  The bytes are the
 * reference's high-memory bytes (miniwfa.c:281-308), so the CIGAR is the reference's; only seg_P rows exist at a time. */
static void run_tile_segmented(mwf_b200_batch_t *b)
{
	const int wp = b->wave_pairs;
	const bool lowmem = b->opt.step > 0; /* the segments are walked for checkpoints (then the banded pass 2 follows), not for the CIGAR */
	if (b->d_snap == 0) { /* workspaces of this mode, on first use: the big arena shrinks to make room for the snapshots */
		CUDA_OK(cudaStreamSynchronize(b->stream));
		ws_free(b->d_arena);
		b->d_arena = 0;
		size_t free_b = 0, total_b = 0;
		CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
		free_b += ws_cached_bytes(b->dev); /* (cached buffers count as free: ws_alloc hands them out, or releases them when it must) */
		b->seg_P = std::max(256, env_int("MWF_B200_TILE_SEGP", 4096) & ~255);
		b->snapdir_stride = (int)(b->max_sbound / b->seg_P + 2);
		const long long half = (long long)((double)free_b * 0.45) & ~255LL;
		const long long width = b->max_len + 2LL * b->pen.nring + 2LL * TILE_TMAX + 16;
		long long need = (long long)(b->seg_P + 2 * TILE_TMAX) * width * wp; /* traceback rows of one segment (at the full width: a loose bound) */
		if (lowmem) need = half; /* ... and of pass 2, which is banded only while the reference's checkpoint matching keeps up: with
		                            step below the penalties two snapshots can share a checkpoint and the band stops collapsing (:413-416) */
		b->arena_total = std::min(half, (need + 255) & ~255LL);
		/* snapshots: R rows as wide as the band, one every seg_P scores.  Worst case: the all-gap score bound at full width; with
		 * a score estimate (a few very long pairs): the band of snapshot k is ~ 2 k seg_P wide.  Too small an estimate is found
		 * out by the forward pass, which is then rerun with the worst case. */
		long long snap_bytes = std::min(half, (long long)b->snapdir_stride * b->tR * width * 4 * wp);
		if (b->s_est > 0) {
			const double n_snap = b->s_est / b->seg_P + 2, w0 = 2.0 * b->pen.nring + 2.0 * TILE_TMAX + 64;
			const double est = 1.25 * wp * b->tR * 4.0 * (b->seg_P * n_snap * n_snap + n_snap * w0) + (64 << 20);
			snap_bytes = std::min(snap_bytes, (long long)est & ~255LL);
		}
		b->snap_words = snap_bytes / 4;
		/* a few very long pairs: recompute several segments at once, each on a virtual slot of its own (state, TileCtl, alive
		 * words, queue room): a single pair's block of scores leaves the last wave of tiles nearly empty, K segments interleave */
		b->seg_K = 1;
		if (!lowmem && b->persist && wp <= 8) {
			const size_t per_slot = (size_t)2 * b->tR * b->pitch * 4 + (size_t)b->pitch * 4;
			int K = std::max(1, std::min(env_int("MWF_B200_TILE_SEGPAR", 3), 8));
			while (K > 1 && (double)K * wp * per_slot > 0.1 * (double)free_b) --K;
			if (K > 1) {
				b->seg_K = K;
				ws_free(b->d_state); ws_free(b->d_alive); ws_free(b->d_tctl); ws_free(b->d_qitems);
				if (ws_dev(&b->d_state, (size_t)K * wp * 2 * b->tR * b->pitch * 4, b->dev))
					CUDA_OK(cudaMemsetAsync(b->d_state, 0xC0, (size_t)K * wp * 2 * b->tR * b->pitch * 4, b->stream));
				ws_dev(&b->d_alive, (size_t)K * wp * b->pitch * 4, b->dev);
				ws_dev(&b->d_tctl, sizeof(TileCtl) * K * wp, b->dev);
				size_t need_q = 2 * ((size_t)K * b->items_cap + (size_t)K * wp + (size_t)std::max(b->geom[0].grid_p, b->geom[b->n_geom - 1].grid_p)) + 4096, cap = 1;
				while (cap < need_q) cap <<= 1;
				b->q_mask = (unsigned int)(cap - 1);
				ws_dev(&b->d_qitems, sizeof(unsigned long long) * cap, b->dev);
				b->arena_total = std::min(half, ((long long)K * need + 255) & ~255LL);
			}
		}
		ws_dev(&b->d_arena, (size_t)b->arena_total, b->dev);
		ws_dev(&b->d_snap, (size_t)b->snap_words * 4, b->dev);
		ws_dev(&b->d_snapdir, sizeof(SnapDir) * (size_t)b->snapdir_stride * wp, b->dev);
		ws_dev(&b->d_nsnap, sizeof(int) * wp, b->dev);
		ws_dev(&b->d_sstop, sizeof(int) * wp * b->seg_K, b->dev);
		ws_dev(&b->d_trace, sizeof(TraceState) * wp, b->dev);
		ws_host(&b->h_nsnap, sizeof(int) * wp);
	}
	TParams PP[2];
	tile_params(b, PP);
	for (int g = 0; g < 2; ++g) {
		PP[g].snap_P = b->seg_P, PP[g].snapdir_stride = b->snapdir_stride, PP[g].snap_arena = b->d_snap, PP[g].snap_cap = b->snap_words;
		PP[g].snap_used = (unsigned long long*)(b->d_tmisc + 72), PP[g].snapdir = b->d_snapdir, PP[g].n_snap = b->d_nsnap;
		PP[g].trace = b->d_trace;
	}
	for (int p0 = 0; p0 < b->n; p0 += wp) {
		const int np = std::min(wp, b->n - p0);
		/* forward: scores only, stop tests as in mwf_wfa_core, snapshots every seg_P scores */
		for (int g = 0; g < 2; ++g) {
			PP[g].pair0 = p0, PP[g].n_pairs = np, PP[g].is_tb = 0, PP[g].snap_take = 1, PP[g].s_stop = 0;
			PP[g].max_s = lowmem ? 0 : b->opt.max_s, PP[g].max_iter = lowmem ? 0 : b->opt.max_iter; /* pass 1 has no stop tests (:569-589) */
			PP[g].seg_use = 0;
		}
		CUDA_OK(cudaMemsetAsync(b->d_nsnap, 0, sizeof(int) * np, b->stream));
		const bool timing = getenv("MWF_B200_BATCH_TIMING") != 0;
		struct timespec tsx;
		double t_fwd = 0, t_pass = 0, t_walk = 0, t0x = 0, t1x = 0;
#define SEG_NOW(v_) do { if (timing) { CUDA_OK(cudaStreamSynchronize(b->stream)); clock_gettime(CLOCK_MONOTONIC, &tsx); v_ = 1e3 * tsx.tv_sec + 1e-6 * tsx.tv_nsec; } } while (0)
		SEG_NOW(t0x);
		if (tile_pass(b, PP, np, -1, true)) {
			if (b->s_est <= 0) die("device workspace exhausted (snapshots of the segmented traceback)");
			/* the score estimate was too low: the worst-case snapshot arena, and once more from score 0 */
			b->s_est = 0;
			CUDA_OK(cudaStreamSynchronize(b->stream));
			ws_free(b->d_snap); ws_free(b->d_snapdir); ws_free(b->d_nsnap); ws_free(b->d_sstop); ws_free(b->d_trace); ws_free(b->h_nsnap);
			b->d_snap = 0, b->d_snapdir = 0, b->d_nsnap = 0, b->d_sstop = 0, b->d_trace = 0, b->h_nsnap = 0;
			run_tile_segmented(b);
			return;
		}
		SEG_NOW(t1x); t_fwd = t1x - t0x;
		wfa_tile_trace_begin_kernel<<<(np + 127) / 128, 128, 0, b->stream>>>(PP[0]);
		CUDA_OK(cudaGetLastError());
		++b->launches;
		CUDA_OK(cudaMemcpyAsync(b->h_nsnap, b->d_nsnap, sizeof(int) * np, cudaMemcpyDeviceToHost, b->stream));
		CUDA_OK(cudaMemcpyAsync(b->h_outs, b->d_outs, sizeof(PairOut) * b->n, cudaMemcpyDeviceToHost, b->stream));
		CUDA_OK(cudaStreamSynchronize(b->stream));
		int max_seg = 0;
		for (int i = 0; i < np; ++i) {
			if (b->h_outs[b->order[p0 + i]].status == ST_ARENA) die("device workspace exhausted (snapshots of the segmented traceback)");
			max_seg = std::max(max_seg, b->h_nsnap[i]);
		}
		/* backward: one segment at a time, traceback bytes of that segment only */
		for (int g = 0; g < 2; ++g)
			PP[g].is_tb = 1, PP[g].snap_take = 0, PP[g].s_stop = b->d_sstop, PP[g].max_s = 0, PP[g].max_iter = 0;
		const int K = b->seg_K;
		for (int j = max_seg; j >= 0; j -= K) {
			const int kk = std::min(K, j + 1); /* segments j, j - 1, ..., j - kk + 1 in one pass */
			SEG_NOW(t0x);
			if (tile_pass(b, PP, np, j, false, kk)) die("device workspace exhausted (traceback bytes of one segment); lower MWF_B200_TILE_SEGP");
			SEG_NOW(t1x); t_pass += t1x - t0x;
			for (int k = 0; k < kk; ++k) { /* the walks, in order: each resumes where the one above it stopped */
				if (lowmem) wfa_tile_ckpt_seg_kernel<<<np, 32, 0, b->stream>>>(PP[0], j - k);
				else wfa_tile_trace_seg_kernel<<<np, 32, 0, b->stream>>>(PP[0], j - k, k * np);
				++b->launches;
			}
			CUDA_OK(cudaGetLastError());
			SEG_NOW(t0x); t_walk += t0x - t1x;
		}
		if (timing) fprintf(stderr, "[mwf_b200] segmented traceback: forward pass with snapshots %.1f ms, %d segments recomputed %.1f ms, walked %.1f ms\n", t_fwd, max_seg + 1, t_pass, t_walk);
#undef SEG_NOW
		if (lowmem && env_int("MWF_B200_DEBUG", 0)) {
			int ns = 0, sg[16];
			CUDA_OK(cudaStreamSynchronize(b->stream));
			CUDA_OK(cudaMemcpy(&ns, b->d_nseg, sizeof(int), cudaMemcpyDeviceToHost));
			CUDA_OK(cudaMemcpy(sg, b->d_seg, sizeof(int) * 16, cudaMemcpyDeviceToHost));
			fprintf(stderr, "[seg dbg] n_seg=%d seg0=(%d,%d) seg1=(%d,%d)\n", ns, sg[0], sg[1], sg[2], sg[3]);
		}
		if (lowmem) { /* pass 2: mwf_wfa_core with the checkpoints, its own (banded) traceback bytes and traceback */
			for (int g = 0; g < 2; ++g)
				PP[g].seg_use = 1, PP[g].s_stop = 0, PP[g].max_s = b->opt.max_s, PP[g].max_iter = b->opt.max_iter;
			tile_pass(b, PP, np);
			wfa_tile_traceback_kernel<<<np, 32, 0, b->stream>>>(PP[0]);
			CUDA_OK(cudaGetLastError());
			++b->launches;
		}
	}
}

/* the first, expected-size arena was too small: take all that is free (false when that is what we already have) */
static bool grow_arena(mwf_b200_batch_t *b)
{
	if (b->arena_total >= b->arena_full) return false;
	CUDA_OK(cudaStreamSynchronize(b->stream));
	ws_free(b->d_arena);
	b->d_arena = 0;
	b->arena_total = ws_dev_flex(&b->d_arena, b->arena_full, b->arena_full / 2, b->dev) & ~255LL;
	b->arena_full = b->arena_total;
	return true;
}

/* High-memory CIGAR of a few very long pairs: will the s^2 traceback bytes overflow everything that is free?  The all-at-once
 * attempt finds out the hard way (1.5 s on a 5 Mb pair at 3 %, all of it thrown away); the fraction f of shared 13-mers tells
 * beforehand (3 ms, kmer_front.cuh): f ~ (1 - p)^13 for a per-base difference rate p, s >~ 0.8 x p n (every difference costs
 * at least about a mismatch; 0.8 leaves room for the estimate), and the bytes are ~ s^2.  A wrong guess only costs time -- both
 * routes give the same CIGAR -- so the test is one-sided: skip the attempt only when the low estimate already overflows. */
static bool predict_arena_overflow(mwf_b200_batch_t *b, double *bytes_likely = 0)
{
	if (bytes_likely) *bytes_likely = 0;
	if (!env_int("MWF_B200_TILE_PREDICT", 1) || b->n > 8 || b->max_len < env_int("MWF_B200_TILE_PREDICT_MINLEN", 2000000)) return false;
	double bytes = 0, likely = 0;
	for (int i = 0; i < b->n; ++i) {
		const PairDesc &p = b->pairs[i];
		int64_t n1 = 0, n2 = 0, shared = 0;
		if (p.tl < 13 || p.ql < 13) continue;
		mwf_b200_kmer_shared(p.tl, (const char*)b->h_seq + p.t_off, p.ql, (const char*)b->h_seq + p.q_off, 13, &n1, &n2, &shared);
		const double f = std::min(n1, n2) > 0 ? (double)shared / (double)std::min(n1, n2) : 0.0;
		const double diff = 1.0 - pow(std::min(1.0, std::max(f, 1e-9)), 1.0 / 13.0);
		const double gap = gap_cost(&b->opt, p.tl > p.ql ? p.tl - p.ql : p.ql - p.tl), xn = b->opt.x * diff * std::max(p.tl, p.ql);
		const double s_low = 0.8 * xn + gap, s_hi = 1.35 * xn + gap + 4096; /* (synthetic 5 Mb pairs: s = 1.19 x n diff) */
		bytes += s_low * s_low;
		likely += s_hi * s_hi + (TILE_TMAX + 2.0) * ((double)p.tl + p.ql);
		b->s_est = std::max(b->s_est, s_hi);
	}
	if (bytes_likely) *bytes_likely = likely;
	return bytes > (double)b->arena_full;
}

/* the deferred arena of a few very long pairs: what the pairs are likely to need (a low guess only costs a rerun with the grown
 * arena), or nothing at all when even the low estimate overflows the device -- the segmented traceback brings its own */
static bool size_deferred_arena(mwf_b200_batch_t *b)
{
	double likely = 0;
	const bool overflow = predict_arena_overflow(b, &likely);
	b->arena_deferred = false;
	if (overflow) return true;
	size_t free_b = 0, total_b = 0;
	CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
	free_b += ws_cached_bytes(b->dev);
	b->arena_full = std::min(b->arena_full, (long long)((double)free_b * env_int("MWF_B200_ARENA_PCT", 85) / 100.0) & ~255LL);
	const long long want = (long long)std::min((double)b->arena_full, std::max(likely, 64.0 * 1048576)) & ~255LL;
	b->arena_total = ws_dev_shrinking(&b->d_arena, want, 4096, b->dev);
	return false;
}

extern "C" void mwf_b200_batch_run(mwf_b200_batch_t *b)
{
	CUDA_OK(cudaSetDevice(b->dev));
	b->launches = 0;
	CUDA_OK(cudaEventRecord(b->ev0, b->stream));
	if (b->n > 0) {
		if (b->kernel == MWF_B200_KERNEL_TILE && b->is_tb && b->opt.step <= 0) { /* high-memory CIGAR */
			bool segmented = getenv("MWF_B200_TILE_SEGP") != 0 || b->d_snap != 0;
			if (b->arena_deferred) { if (size_deferred_arena(b)) segmented = true; }
			else if (!segmented && b->arena_total >= b->arena_full && predict_arena_overflow(b)) segmented = true;
			while (!segmented) { /* optimistic: all s^2 traceback bytes at once */
				if (run_tile(b)) break;
				if (!grow_arena(b)) segmented = true;
			}
			if (segmented) run_tile_segmented(b);
		} else if (b->kernel == MWF_B200_KERNEL_TILE && b->is_tb) { /* low-memory mode */
			bool segmented = getenv("MWF_B200_TILE_SEGP") != 0 || b->d_snap != 0;
			if (b->arena_deferred && size_deferred_arena(b)) segmented = true;
			while (!segmented && !run_tile(b)) /* the unbanded pass does not fit the arena as s^2 bytes */
				if (!grow_arena(b)) segmented = true;
			if (segmented) {
				if (env_int("MWF_B200_LOWMEM_STREAMING", 0)) { /* the reference's two-stripe pass 1 on the streaming kernels */
					CUDA_OK(cudaStreamSynchronize(b->stream));
					free_tile(b);
					b->kernel = (b->n >= b->n_sm / 4 || b->max_len < 32768) ? MWF_B200_KERNEL_CTA : MWF_B200_KERNEL_GRID;
					alloc_streaming(b);
				} else run_tile_segmented(b);
			}
		} else if (b->kernel == MWF_B200_KERNEL_TILE) run_tile(b);
		if (b->kernel == MWF_B200_KERNEL_TILE) {
		} else if (b->kernel == MWF_B200_KERNEL_CTA) {
			launch_cta(b, make_params(b, b->n_slots), b->n_slots);
		} else {
			const KParams P = make_params(b, 1);
			for (int i = 0; i < b->n; ++i) launch_grid(b, P, b->order[i]);
		}
	}
	CUDA_OK(cudaEventRecord(b->ev1, b->stream));
	CUDA_OK(cudaMemcpyAsync(b->h_outs, b->d_outs, sizeof(PairOut) * b->n, cudaMemcpyDeviceToHost, b->stream));
	b->ran = true, b->timed = false;
}

extern "C" void mwf_b200_batch_wait(mwf_b200_batch_t *b)
{
	CUDA_OK(cudaSetDevice(b->dev));
	CUDA_OK(cudaStreamSynchronize(b->stream));
	if (!b->ran || b->timed) return;
	float ms = 0;
	CUDA_OK(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
	b->kernel_ms = ms;
	b->timed = true;
	/* pairs whose slot ran out of arena are retried one at a time with the whole arena */
	std::vector<int> retry;
	for (int i = 0; i < b->n; ++i) {
		const int st = b->h_outs[i].status;
		if (st == ST_ARENA && b->kernel == MWF_B200_KERNEL_CTA && b->n_slots > 1) retry.push_back(i);
		else if (st == ST_ARENA) die("device workspace exhausted (traceback/snapshot arena); use opt.step>0 or a smaller batch");
		else if (st == ST_SHRINK) die("internal error: empty band after shrink");
		else if (st == ST_CORRUPT) die("internal error: checkpoint chain is corrupt");
	}
	if (!retry.empty()) {
		CUDA_OK(cudaEventRecord(b->ev0, b->stream));
		KParams P = make_params(b, 1);
		std::vector<int> one(1);
		for (size_t r = 0; r < retry.size(); ++r) {
			one[0] = retry[r];
			CUDA_OK(cudaMemcpyAsync(b->d_order, one.data(), sizeof(int), cudaMemcpyHostToDevice, b->stream));
			P.n_pairs = 1;
			launch_cta(b, P, 1);
			CUDA_OK(cudaStreamSynchronize(b->stream));
		}
		CUDA_OK(cudaEventRecord(b->ev1, b->stream));
		CUDA_OK(cudaMemcpyAsync(b->h_outs, b->d_outs, sizeof(PairOut) * b->n, cudaMemcpyDeviceToHost, b->stream));
		CUDA_OK(cudaMemcpyAsync(b->d_order, b->order.data(), sizeof(int) * b->n, cudaMemcpyHostToDevice, b->stream));
		CUDA_OK(cudaStreamSynchronize(b->stream));
		CUDA_OK(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
		b->kernel_ms += ms;
		for (size_t r = 0; r < retry.size(); ++r)
			if (b->h_outs[retry[r]].status == ST_ARENA)
				die("device workspace exhausted (traceback/snapshot arena); use opt.step>0");
	}
}

/* CIGARs of many pairs: packed back to back on the device (one warp per pair), so that the batch needs one copy to the host
 * instead of one per pair (~10 us each: 1 s for the 100 000 gap fills of a 5 Mb mwf_wfa_chain) */
__global__ void __launch_bounds__(256) cigar_gather_kernel(const PairOut *__restrict__ outs, const long long *__restrict__ off,
                                                           const uint32_t *__restrict__ cigar, uint32_t *__restrict__ dst, int n)
{
	const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (w >= n) return;
	const long long o0 = off[w], cnt = off[w + 1] - o0;
	if (cnt <= 0) return;
	const uint32_t *src = cigar + outs[w].cigar_pos;
	for (long long j = lane; j < cnt; j += 32) dst[o0 + j] = src[j];
}

extern "C" void mwf_b200_batch_fetch(mwf_b200_batch_t *b, void *km, mwf_rst_t *r)
{
	mwf_b200_batch_wait(b);
	b->d2h = (int64_t)sizeof(PairOut) * b->n;
	long long total = 0;
	int with_cigar = 0;
	for (int i = 0; i < b->n; ++i) {
		const PairOut &o = b->h_outs[i];
		r[i].s = o.s, r[i].n_iter = o.n_iter, r[i].n_cigar = 0, r[i].cigar = 0;
		if ((b->opt.flag & MWF_F_DEBUG) && b->is_tb && o.s >= 0) /* wf_traceback's line, miniwfa.c:367: it counts traceback rows, row = score - 1 */
			fprintf(stderr, "s0=%d, s=%d, i=%d, k=%d\n", o.s - 1, o.end_s - 1, o.end_i, o.end_k);
		if (o.s >= 0 && o.n_cigar > 0) {
			r[i].n_cigar = o.n_cigar;
			r[i].cigar = (uint32_t*)kmalloc(km, sizeof(uint32_t) * (size_t)o.n_cigar);
			total += o.n_cigar, ++with_cigar;
		}
	}
	b->d2h += (int64_t)sizeof(uint32_t) * total;
	if (with_cigar > 8) { /* gather on the device, one copy, scatter on the host */
		long long *h_off = 0, *d_off = 0;
		uint32_t *d_all = 0, *h_all = 0;
		ws_host(&h_off, sizeof(long long) * ((size_t)b->n + 1));
		ws_dev(&d_off, sizeof(long long) * ((size_t)b->n + 1), b->dev);
		ws_dev(&d_all, sizeof(uint32_t) * (size_t)total, b->dev);
		ws_host(&h_all, sizeof(uint32_t) * (size_t)total);
		h_off[0] = 0;
		for (int i = 0; i < b->n; ++i) h_off[i + 1] = h_off[i] + r[i].n_cigar;
		CUDA_OK(cudaMemcpyAsync(d_off, h_off, sizeof(long long) * ((size_t)b->n + 1), cudaMemcpyHostToDevice, b->stream));
		cigar_gather_kernel<<<(unsigned)(((long long)b->n * 32 + 255) / 256), 256, 0, b->stream>>>(b->d_outs, d_off, b->d_cigar, d_all, b->n);
		CUDA_OK(cudaGetLastError());
		++b->launches;
		CUDA_OK(cudaMemcpyAsync(h_all, d_all, sizeof(uint32_t) * (size_t)total, cudaMemcpyDeviceToHost, b->stream));
		CUDA_OK(cudaStreamSynchronize(b->stream));
		for (int i = 0; i < b->n; ++i)
			if (r[i].n_cigar > 0) memcpy(r[i].cigar, h_all + h_off[i], sizeof(uint32_t) * (size_t)r[i].n_cigar);
		ws_free(h_off); ws_free(d_off); ws_free(d_all); ws_free(h_all);
	} else if (with_cigar > 0) { /* few (possibly very long) CIGARs: straight into the caller's buffers */
		for (int i = 0; i < b->n; ++i)
			if (r[i].n_cigar > 0)
				CUDA_OK(cudaMemcpyAsync(r[i].cigar, b->d_cigar + b->h_outs[i].cigar_pos, sizeof(uint32_t) * (size_t)r[i].n_cigar,
				                        cudaMemcpyDeviceToHost, b->stream));
		CUDA_OK(cudaStreamSynchronize(b->stream));
	}
}

extern "C" void mwf_b200_batch_destroy(mwf_b200_batch_t *b)
{
	if (!b) return;
	CUDA_OK(cudaSetDevice(b->dev));
	CUDA_OK(cudaStreamSynchronize(b->stream)); /* the workspaces go back to the cache: nothing may still use them */
	ws_free(b->d_seq); ws_free(b->h_seq); ws_free(b->d_pairs); ws_free(b->d_outs); ws_free(b->h_outs);
	ws_free(b->d_order); ws_free(b->d_ctl); ws_free(b->d_ring); ws_free(b->d_ring2); ws_free(b->d_arena);
	ws_free(b->d_rowtab); ws_free(b->d_snapoff); ws_free(b->d_snaphdr); ws_free(b->d_seg); ws_free(b->d_cigar);
	ws_free(b->d_tctl); ws_free(b->d_state); ws_free(b->d_alive); ws_free(b->d_items); ws_free(b->d_qitems); ws_free(b->d_tmisc); ws_free(b->d_nseg);
	ws_free(b->d_seqp); ws_free(b->d_seqp2); ws_free(b->d_packed);
	ws_free(b->d_snap); ws_free(b->d_snapdir); ws_free(b->d_nsnap); ws_free(b->d_sstop); ws_free(b->h_nsnap); ws_free(b->d_trace);
	if (b->h_running) { ws_free(b->h_running); cudaEventDestroy(b->evc[0]); cudaEventDestroy(b->evc[1]); }
	cudaEventDestroy(b->ev0); cudaEventDestroy(b->ev1);
	if (b->own_stream) cudaStreamDestroy(b->stream);
	delete b;
}

extern "C" double mwf_b200_batch_kernel_ms(const mwf_b200_batch_t *b) { return b->kernel_ms; }
extern "C" int64_t mwf_b200_batch_launches(const mwf_b200_batch_t *b) { return b->launches; }
extern "C" int mwf_b200_batch_kernel_used(const mwf_b200_batch_t *b) { return b->kernel; }
extern "C" int64_t mwf_b200_batch_h2d_bytes(const mwf_b200_batch_t *b) { return b->h2d; }
extern "C" int64_t mwf_b200_batch_d2h_bytes(const mwf_b200_batch_t *b) { return b->d2h; }

/* How many devices a batch is spread over.  mwf_b200_set_devices(n): n = 1 keeps every batch on one device, n > 1 uses n
 * devices starting at the current one, 0 (default) decides per batch: all visible devices ($MWF_B200_DEVICES caps them) when
 * no device was pinned -- by mwf_b200_set_device(), $MWF_B200_DEVICE or $LOCAL_RANK (one process per GPU: torchrun, MPI) --
 * and the batch is worth it (at least two pairs and a sum of squared lengths of 1e11 -- ten 100 kb pairs: a device costs
 * about a millisecond to set up, and the many tiny gap fills of mwf_wfa_chain must stay on one). */
static int batch_devices(int n_pairs, const int32_t *tl, const int32_t *ql)
{
	int want = g_n_devices;
	if (want == 1 || n_pairs < 2) return 1;
	const int vis = mwf_b200_device_count();
	if (want == 0) {
		if (g_device_explicit || getenv("MWF_B200_DEVICE") || getenv("LOCAL_RANK")) return 1;
		want = env_int("MWF_B200_DEVICES", vis);
		double work = 0; /* wavefront cells grow with the square of the length */
		for (int i = 0; i < n_pairs; ++i) work += (double)std::max(tl[i], ql[i]) * std::max(tl[i], ql[i]);
		if (work < 1e11) return 1;
	}
	return std::max(1, std::min(std::min(want, vis), n_pairs));
}

static void exact_batch_one_device(void *km, const mwf_opt_t *opt, int32_t n_pairs, const int32_t *tl, const char *const *ts,
                                   const int32_t *ql, const char *const *qs, mwf_rst_t *r)
{
	const bool timing = getenv("MWF_B200_BATCH_TIMING") != 0; /* phase times to stderr */
	double t[6] = { 0, 0, 0, 0, 0, 0 };
	struct timespec ts0;
#define BATCH_T(i_) do { if (timing) { clock_gettime(CLOCK_MONOTONIC, &ts0); t[i_] = 1e3 * ts0.tv_sec + 1e-6 * ts0.tv_nsec; } } while (0)
	BATCH_T(0);
	mwf_b200_batch_t *b = mwf_b200_batch_create(opt, n_pairs, tl, ql);
	BATCH_T(1);
	mwf_b200_batch_upload(b, ts, qs);
	BATCH_T(2);
	mwf_b200_batch_run(b);
	mwf_b200_batch_wait(b);
	BATCH_T(3);
	mwf_b200_batch_fetch(b, km, r);
	BATCH_T(4);
	mwf_b200_batch_destroy(b);
	BATCH_T(5);
#undef BATCH_T
	if (timing)
		fprintf(stderr, "[mwf_b200] batch of %d: create %.2f, upload %.2f, run %.2f, fetch %.2f, destroy %.2f ms\n", n_pairs,
		        t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4]);
}

/* n independent pairs (the reference's CLI loops over them, main.c:67).  With several devices the pairs are dealt
 * out by cost (wavefront cells grow with the square of the length; longest first, each to the least loaded device), one host
 * thread per device creates, uploads and runs its shard on its own stream, and the calling thread then fetches shard after
 * shard: the results -- CIGARs from the caller's km, which is not thread-safe -- land in r in input order.  No data crosses
 * between devices. */
extern "C" void mwf_wfa_exact_batch(void *km, const mwf_opt_t *opt, int32_t n_pairs,
                                    const int32_t *tl, const char *const *ts,
                                    const int32_t *ql, const char *const *qs, mwf_rst_t *r)
{
	const int n_dev = batch_devices(n_pairs, tl, ql);
	{ /* The register-resident steps of the tile engine address the sequences of a batch by 32-bit bit positions, which holds for
	   * up to 2^28 bytes of sequence per device batch (about 1300 pairs of 100 kb); a larger submission is cut into consecutive
	   * parts of at most that size per device -- the same results, and the same rate, as one batch would give */
		const double cap = 200e6 * n_dev;
		double bytes = 0;
		for (int i = 0; i < n_pairs; ++i) bytes += (double)tl[i] + ql[i] + 160;
		if (bytes > cap && n_pairs > 1) {
			int i0 = 0;
			while (i0 < n_pairs) {
				double part = 0;
				int i1 = i0;
				while (i1 < n_pairs && (i1 == i0 || part + tl[i1] + ql[i1] + 160 <= cap)) part += (double)tl[i1] + ql[i1] + 160, ++i1;
				mwf_wfa_exact_batch(km, opt, i1 - i0, tl + i0, ts + i0, ql + i0, qs + i0, r + i0);
				i0 = i1;
			}
			return;
		}
	}
	if (n_dev <= 1) { exact_batch_one_device(km, opt, n_pairs, tl, ts, ql, qs, r); return; }
	const int dev0 = mwf_b200_get_device(), vis = mwf_b200_device_count();
	struct Shard { std::vector<int> idx; std::vector<int32_t> tl, ql; std::vector<const char*> ts, qs; std::vector<mwf_rst_t> r; mwf_b200_batch_t *b; double load; };
	std::vector<Shard> sh(n_dev);
	for (int d = 0; d < n_dev; ++d) sh[d].b = 0, sh[d].load = 0;
	std::vector<int> order(n_pairs);
	for (int i = 0; i < n_pairs; ++i) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return std::max(tl[a], ql[a]) > std::max(tl[c], ql[c]); });
	for (int k = 0; k < n_pairs; ++k) {
		const int i = order[k];
		int best = 0;
		for (int d = 1; d < n_dev; ++d) if (sh[d].load < sh[best].load) best = d;
		const double n = std::max(tl[i], ql[i]);
		sh[best].load += n * n + 1.0;
		sh[best].idx.push_back(i);
	}
	std::vector<std::thread> workers;
	for (int d = 0; d < n_dev; ++d) {
		Shard &S = sh[d];
		std::sort(S.idx.begin(), S.idx.end());
		for (size_t k = 0; k < S.idx.size(); ++k) {
			const int i = S.idx[k];
			S.tl.push_back(tl[i]), S.ql.push_back(ql[i]), S.ts.push_back(ts[i]), S.qs.push_back(qs[i]);
		}
		S.r.resize(S.idx.size());
		workers.push_back(std::thread([&S, d, dev0, vis, opt]() {
			tl_device = (dev0 + d) % vis;
			S.b = mwf_b200_batch_create(opt, (int32_t)S.idx.size(), S.tl.data(), S.ql.data());
			mwf_b200_batch_upload(S.b, S.ts.data(), S.qs.data());
			mwf_b200_batch_run(S.b);
			mwf_b200_batch_wait(S.b);
			tl_device = -1;
		}));
	}
	for (size_t d = 0; d < workers.size(); ++d) workers[d].join();
	for (int d = 0; d < n_dev; ++d) {
		Shard &S = sh[d];
		mwf_b200_batch_fetch(S.b, km, S.r.data());
		for (size_t k = 0; k < S.idx.size(); ++k) r[S.idx[k]] = S.r[k];
		mwf_b200_batch_destroy(S.b);
	}
	CUDA_OK(cudaSetDevice(dev0));
}

#include "kmer_front.cuh"
