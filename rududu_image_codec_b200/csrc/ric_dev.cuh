// ric_dev.cuh -- device-side building blocks shared by the forward and inverse level kernels.
//
// Exact integer semantics of the reference:
//   lifting        src/lib/wavelet2d.cpp:307-405 (9/7), :593-634 (5/3)
//   block quantiser src/lib/bandcodec.cpp:115-237
// "C-typed store" truncation is TR<SH>() (SH: the level works on `short`).  All >> are arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ric_quant_pk.cuh"

namespace ric {

enum { T97 = 0, T53 = 1, THAAR = 2 };  // Haar: S1/S2 (U2/U1) only, pairs without neighbours; even sizes only

// Strip geometry: one warp owns STRIP_W output columns and loads 8 more on each side
// (lane 0 and lane 31 are halo lanes); every lane holds 8 consecutive columns.
constexpr int STRIP_W = 240;
constexpr int LANE_W = 8;
constexpr unsigned FULL = 0xffffffffu;

// 8 bytes through the read-only path if `ok`, else zeros: one predicated LDG, never a branch
__device__ __forceinline__ uint2 ldg_u2_if(const void *p, bool ok)
{
	uint2 v;
	asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\t@q ld.global.nc.v2.u32 {%0, %1}, [%2];\n\t}"
	    : "=&r"(v.x), "=&r"(v.y) : "l"(p), "r"((int)ok));
	return v;
}

__device__ __forceinline__ uint4 ldg_u4_if(const void *p, bool ok)  // 16 bytes, likewise
{
	uint4 v;
	asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
	    "@q ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
	    : "=&r"(v.x), "=&r"(v.y), "=&r"(v.z), "=&r"(v.w) : "l"(p), "r"((int)ok));
	return v;
}

template <bool SH>
__device__ __forceinline__ int TR(int v) { return SH ? (int)(short)v : v; }

// The same store-to-C truncation on the FMA pipe: IDP.2A.LO.S16.U8 with the byte pair (1, 0) sign-extends the low
// half.  The inverse kernels use it (their ALU pipe is the busier one: -4 % per launch, measured); the forward
// kernels keep the ALU form (their FMA pipe already carries the quantiser's multiplies: +2 % with this form).
template <bool SH>
__device__ __forceinline__ int TRI(int v)
{
	if (SH) {
		int d;
		asm("dp2a.lo.s32.u32 %0, %1, 1, 0;" : "=r"(d) : "r"(v));
		return d;
	}
	return v;
}

__device__ __forceinline__ int m08i(int a)  // mult08<int>, wavelet2d.cpp:307-318
{
	a -= a >> 2;
	a += a >> 4;
	return a + (a >> 8);
}

template <bool SH>
__device__ __forceinline__ int m08c(int a)  // mult08<C>: every assignment truncates to C
{
	a = TR<SH>(a - (a >> 2));
	a = TR<SH>(a + (a >> 4));
	return TR<SH>(a + (a >> 8));
}
template <bool SH>
__device__ __forceinline__ int m08ci(int a)  // the same for the inverse kernels (TRI)
{
	a = TRI<SH>(a - (a >> 2));
	a = TRI<SH>(a + (a >> 4));
	return TRI<SH>(a + (a >> 8));
}

// ---- forward lifting steps (Appendix A.1 / A.3).  x: centre, l/r: neighbours (proper C values).
// KEEP: result feeds a shift next, so it must be truncated now; otherwise truncation is deferred
// to the next truncating use (the low 16 bits are always right: + - * are ring homomorphisms).
template <bool SH, int TRANS>
__device__ __forceinline__ int fS1(int x, int l, int r)
{
	if (TRANS == T97) { int t = TR<SH>(l + r); return TR<SH>(x - (t + (t >> 1))); }
	if (TRANS == THAAR) return TR<SH>(x - r);  // i[0] -= i[1], wavelet2d.cpp:772
	return TR<SH>(x + ((1 - l - r) >> 1));  // x - ((l + r) >> 1), same identity
}
template <bool SH, int TRANS>
__device__ __forceinline__ int fS1_first(int x, int r) { return TRANS == T97 ? TR<SH>(x - 3 * r) : TR<SH>(x - r); }
template <bool SH, int TRANS>
__device__ __forceinline__ int fS1_last(int x, int l) { return TRANS == T97 ? TR<SH>(x - 3 * l) : TR<SH>(x - l); }

template <bool SH, int TRANS>
__device__ __forceinline__ int fS2(int x, int l, int r)
{
	// x - ((l + r) >> 4), written as x + ((15 - l - r) >> 4): -floor(s / 16) = floor((15 - s) / 16) for every integer s,
	// and this form is two instructions (a three-input add and a fused shift-add) instead of three
	if (TRANS == T97) return TR<SH>(x + ((15 - l - r) >> 4));
	if (TRANS == THAAR) return TR<SH>(x + (l >> 1));  // i[1] += i[0] >> 1, wavelet2d.cpp:773
	return TR<SH>(x + ((l + r) >> 2));
}
template <bool SH, int TRANS>
__device__ __forceinline__ int fS2_last(int x, int l) { return TRANS == T97 ? TR<SH>(x - (l >> 3)) : TR<SH>(x + (l >> 1)); }

// S3 / S4 (and U2 / U1 below) results are only ever consumed through a truncating use -- the
// (C)(l + r) temporary of the next step, a 16-bit store, or an edge formula that truncates its
// operand itself -- so their own C-typed store is deferred (the low 16 bits are always right:
// + - * are ring homomorphisms mod 2^16).  Everything that feeds a shift stays truncated.
template <bool SH, int TRANS>
__device__ __forceinline__ int fS3(int x, int l, int r) { return TRANS == T97 ? x + m08i(l + r) : x; }
template <bool SH, int TRANS>
__device__ __forceinline__ int fS3_edge(int x, int n) { return TRANS == T97 ? TR<SH>(x + 2 * m08c<SH>(n)) : x; }

template <bool SH, int TRANS>
__device__ __forceinline__ int fS4(int x, int l, int r)
{
	if (TRANS == T97) { int t = TR<SH>(l + r); return x + ((t >> 1) - (t >> 5)); }
	return x;
}
template <bool SH, int TRANS>
__device__ __forceinline__ int fS4_last(int x, int l)
{
	if (TRANS != T97) return x;
	l = TR<SH>(l);  // S3 results arrive un-truncated
	return TR<SH>(x + (l - (l >> 4)));
}

// ---- inverse lifting steps (Appendix A.2 / A.3): U4 undoes S4, ... U1 undoes S1.
template <bool SH, int TRANS>
__device__ __forceinline__ int iU4(int x, int l, int r)
{
	if (TRANS == T97) { int t = TRI<SH>(l + r); return TRI<SH>(x - ((t >> 1) - (t >> 5))); }
	return x;
}
template <bool SH, int TRANS>
__device__ __forceinline__ int iU4_last(int x, int l)
{
	if (TRANS != T97) return x;
	l = TRI<SH>(l);  // row pass: vertical U1 results arrive un-truncated
	return TRI<SH>(x - (l - (l >> 4)));
}
template <bool SH, int TRANS>
__device__ __forceinline__ int iU3(int x, int l, int r) { return TRANS == T97 ? TRI<SH>(x - m08i(l + r)) : x; }
template <bool SH, int TRANS>
__device__ __forceinline__ int iU3_edge(int x, int n) { return TRANS == T97 ? TRI<SH>(x - 2 * m08ci<SH>(n)) : x; }
template <bool SH, int TRANS>
__device__ __forceinline__ int iU2(int x, int l, int r)
{
	if (TRANS == T97) return x + ((l + r) >> 4);
	if (TRANS == THAAR) return TRI<SH>(x - (l >> 1));  // i[1] -= i[0] >> 1, wavelet2d.cpp:784
	return TRI<SH>(x + ((3 - l - r) >> 2));  // 5/3: x - ((l + r) >> 2) (same identity); feeds the un-truncated (l + r) >> 1 of U1
}
template <bool SH, int TRANS>
__device__ __forceinline__ int iU2_last(int x, int l) { return TRANS == T97 ? TRI<SH>(x + (l >> 3)) : TRI<SH>(x - (l >> 1)); }
template <bool SH, int TRANS>
__device__ __forceinline__ int iU1(int x, int l, int r)
{
	if (TRANS == T97) { int t = TRI<SH>(l + r); return x + (t + (t >> 1)); }
	if (TRANS == THAAR) return TRI<SH>(x + r);  // i[0] += i[1], wavelet2d.cpp:785
	return TRI<SH>(x + ((l + r) >> 1));
}
template <bool SH, int TRANS>
__device__ __forceinline__ int iU1_first(int x, int r) { return TRANS == T97 ? TRI<SH>(x + 3 * r) : TRI<SH>(x + r); }
template <bool SH, int TRANS>
__device__ __forceinline__ int iU1_last(int x, int l) { return TRANS == T97 ? TRI<SH>(x + 3 * l) : TRI<SH>(x + l); }

// ---- horizontal passes on the 8 columns a lane holds (cb = absolute column of v[0], even).
// Neighbours across lanes come from warp shuffles.  Every element is first computed with the
// interior formula; strips that contain column 0 or column w-1 (edge_x, warp-uniform) then
// overwrite those two elements with the absolute-coordinate edge formulas of the reference
// (the values they replace were never used by another step).
// NT (forward only): "no truncation" -- for rows lifted straight from 8-bit pixels (|x| <= 2048)
// no intermediate of the four steps can leave the int16 range (bounds: S1 8192, S2 3072,
// S3 13108, S4 17000), so the C-typed stores are identities and are skipped.
struct EdgeX {
	bool on;        // this strip holds column 0 or column w-1
	bool first;     // this lane holds column 0 (as its element 0)
	bool last;      // this lane holds column w-1 ...
	int kl;         // ... as its element kl
};

__device__ __forceinline__ EdgeX make_edge_x(int cb, int w, bool on)
{
	EdgeX e;
	e.on = on;
	e.first = cb == 0;
	e.last = cb == ((w - 1) & ~7);
	e.kl = (w - 1) & 7;
	return e;
}

template <bool SH, int TRANS, bool NT, bool EDGE>
__device__ __forceinline__ void row_fwd(int (&v)[8], const EdgeX &ee)
{
	constexpr bool S = SH && !NT;
	EdgeX e = ee;
	if (!EDGE) e.on = false;  // interior strips: every fix-up below folds away at compile time
	int nl, nr, o[8];
	// S1: even columns
	nl = __shfl_up_sync(FULL, v[7], 1);
#pragma unroll
	for (int k = 0; k < 8; k += 2) { o[k] = v[k]; v[k] = fS1<S, TRANS>(v[k], k ? v[k - 1] : nl, v[k + 1]); }
	if (e.on) {
		if (e.first) v[0] = fS1_first<S, TRANS>(o[0], v[1]);
		if (e.last) {
#pragma unroll
			for (int k = 0; k < 8; k += 2) if (e.kl == k) v[k] = fS1_last<S, TRANS>(o[k], k ? v[k - 1] : nl);
		}
	}
	// S2: odd columns
	nr = __shfl_down_sync(FULL, v[0], 1);
#pragma unroll
	for (int k = 1; k < 8; k += 2) { o[k] = v[k]; v[k] = fS2<S, TRANS>(v[k], v[k - 1], k < 7 ? v[k + 1] : nr); }
	if (e.on && e.last) {
#pragma unroll
		for (int k = 1; k < 8; k += 2) if (e.kl == k) v[k] = fS2_last<S, TRANS>(o[k], v[k - 1]);
	}
	if (TRANS != T97) return;
	// S3
	nl = __shfl_up_sync(FULL, v[7], 1);
#pragma unroll
	for (int k = 0; k < 8; k += 2) { o[k] = v[k]; v[k] = fS3<S, TRANS>(v[k], k ? v[k - 1] : nl, v[k + 1]); }
	if (e.on) {
		if (e.first) v[0] = fS3_edge<S, TRANS>(o[0], v[1]);
		if (e.last) {
#pragma unroll
			for (int k = 0; k < 8; k += 2) if (e.kl == k) v[k] = fS3_edge<S, TRANS>(o[k], k ? v[k - 1] : nl);
		}
	}
	// S4
	nr = __shfl_down_sync(FULL, v[0], 1);
#pragma unroll
	for (int k = 1; k < 8; k += 2) { o[k] = v[k]; v[k] = fS4<S, TRANS>(v[k], v[k - 1], k < 7 ? v[k + 1] : nr); }
	if (e.on && e.last) {
#pragma unroll
		for (int k = 1; k < 8; k += 2) if (e.kl == k) v[k] = fS4_last<S, TRANS>(o[k], v[k - 1]);
	}
}

template <bool SH, int TRANS, bool EDGE>
__device__ __forceinline__ void row_inv(int (&v)[8], const EdgeX &ee)
{
	EdgeX e = ee;
	if (!EDGE) e.on = false;  // interior strips: every fix-up below folds away at compile time
	int nl, nr, o[8];
	if (TRANS == T97) {
		// U4: odd columns
		nr = __shfl_down_sync(FULL, v[0], 1);
#pragma unroll
		for (int k = 1; k < 8; k += 2) { o[k] = v[k]; v[k] = iU4<SH, TRANS>(v[k], v[k - 1], k < 7 ? v[k + 1] : nr); }
		if (e.on && e.last) {
#pragma unroll
			for (int k = 1; k < 8; k += 2) if (e.kl == k) v[k] = iU4_last<SH, TRANS>(o[k], v[k - 1]);
		}
		// U3: even columns
		nl = __shfl_up_sync(FULL, v[7], 1);
#pragma unroll
		for (int k = 0; k < 8; k += 2) { o[k] = v[k]; v[k] = iU3<SH, TRANS>(v[k], k ? v[k - 1] : nl, v[k + 1]); }
		if (e.on) {
			if (e.first) v[0] = iU3_edge<SH, TRANS>(o[0], v[1]);
			if (e.last) {
#pragma unroll
				for (int k = 0; k < 8; k += 2) if (e.kl == k) v[k] = iU3_edge<SH, TRANS>(o[k], k ? v[k - 1] : nl);
			}
		}
	}
	// U2: odd columns
	nr = __shfl_down_sync(FULL, v[0], 1);
#pragma unroll
	for (int k = 1; k < 8; k += 2) { o[k] = v[k]; v[k] = iU2<SH, TRANS>(v[k], v[k - 1], k < 7 ? v[k + 1] : nr); }
	if (e.on && e.last) {
#pragma unroll
		for (int k = 1; k < 8; k += 2) if (e.kl == k) v[k] = iU2_last<SH, TRANS>(o[k], v[k - 1]);
	}
	// U1: even columns
	nl = __shfl_up_sync(FULL, v[7], 1);
#pragma unroll
	for (int k = 0; k < 8; k += 2) { o[k] = v[k]; v[k] = iU1<SH, TRANS>(v[k], k ? v[k - 1] : nl, v[k + 1]); }
	if (e.on) {
		if (e.first) v[0] = iU1_first<SH, TRANS>(o[0], v[1]);
		if (e.last) {
#pragma unroll
			for (int k = 0; k < 8; k += 2) if (e.kl == k) v[k] = iU1_last<SH, TRANS>(o[k], k ? v[k - 1] : nl);
		}
	}
}

// ---- vertical steps on 8-column row vectors held by one lane (edge choice is warp-uniform) ----
#define RIC_VSTEP(NAME, GEN, FIRST, LAST)                                                           \
	template <bool SH, int TRANS, bool EDGE>                                                        \
	__device__ __forceinline__ void NAME(int (&x)[8], const int (&l)[8], const int (&r)[8], bool first, \
	                                     bool last)                                                 \
	{                                                                                               \
		if (EDGE && first) {                                                                        \
			_Pragma("unroll") for (int k = 0; k < 8; k++) x[k] = FIRST;                             \
		} else if (EDGE && last) {                                                                  \
			_Pragma("unroll") for (int k = 0; k < 8; k++) x[k] = LAST;                              \
		} else {                                                                                    \
			_Pragma("unroll") for (int k = 0; k < 8; k++) x[k] = GEN;                               \
		}                                                                                           \
	}
RIC_VSTEP(vS1, (fS1<SH, TRANS>(x[k], l[k], r[k])), (fS1_first<SH, TRANS>(x[k], r[k])), (fS1_last<SH, TRANS>(x[k], l[k])))
RIC_VSTEP(vS2, (fS2<SH, TRANS>(x[k], l[k], r[k])), (x[k]), (fS2_last<SH, TRANS>(x[k], l[k])))
RIC_VSTEP(vS3, (fS3<SH, TRANS>(x[k], l[k], r[k])), (fS3_edge<SH, TRANS>(x[k], r[k])), (fS3_edge<SH, TRANS>(x[k], l[k])))
RIC_VSTEP(vS4, (fS4<SH, TRANS>(x[k], l[k], r[k])), (x[k]), (fS4_last<SH, TRANS>(x[k], l[k])))
RIC_VSTEP(vU4, (iU4<SH, TRANS>(x[k], l[k], r[k])), (x[k]), (iU4_last<SH, TRANS>(x[k], l[k])))
RIC_VSTEP(vU3, (iU3<SH, TRANS>(x[k], l[k], r[k])), (iU3_edge<SH, TRANS>(x[k], r[k])), (iU3_edge<SH, TRANS>(x[k], l[k])))
RIC_VSTEP(vU2, (iU2<SH, TRANS>(x[k], l[k], r[k])), (x[k]), (iU2_last<SH, TRANS>(x[k], l[k])))
RIC_VSTEP(vU1, (iU1<SH, TRANS>(x[k], l[k], r[k])), (iU1_first<SH, TRANS>(x[k], r[k])), (iU1_last<SH, TRANS>(x[k], l[k])))
#undef RIC_VSTEP

// Quantise one 4x4 block held in registers (c[4*row+col], proper C values), in place, returning
// the number of non-zero outputs.  bw/bh: valid columns/rows (4,4 = full block -> rank-threshold
// path; otherwise the plain dead-zone path of the partial-block overload).  qb points to shared
// memory.  The candidate ranking reproduces the reference's stable insertion sort (descending
// unsigned value, ties in raster order) through distinct keys (value<<4 | 15-pos):
//   survivors = the m* largest candidates, m* = 1 + max{i : f_(i) >= thr[cnt+i]}  (bandcodec.cpp:188-199).
template <bool SH>
__device__ __forceinline__ int quant_block(int (&c)[16], const QuantBand *qb, int bw, int bh)
{
	const bool full = bw == 4 && bh == 4;
	const int T = full ? qb->T : qb->Te;
	const unsigned T2 = (unsigned)(2 * T);
	if (!full) {
#pragma unroll
		for (int k = 0; k < 16; k++)
			if ((k & 3) >= bw || (k >> 2) >= bh) c[k] = 0;  // outside the band: treated as dead
	}
	// pass 1: dead zone only.  Warps whose blocks are all dead (chroma, coarse quantisers) stop here.
	unsigned alive = 0;
#pragma unroll
	for (int k = 0; k < 16; k++) alive |= ((unsigned)(c[k] + T) > T2) ? (1u << k) : 0u;
	if (!__any_sync(FULL, alive != 0)) {  // callers keep the warp converged (flush_blocks / quant_level_kernel)
#pragma unroll
		for (int k = 0; k < 16; k++) c[k] = 0;
		return 0;
	}
	// pass 2 (branch-free): fold, split into sure non-zeros (quantised now) and rank candidates
	const int iQ = qb->iQ;
	const unsigned uthr0 = full ? uview<SH>(qb->thr[0]) : 0u;  // partial blocks have no candidates
	int key[16];
	int cnt = 0, nc = 0;
#pragma unroll
	for (int k = 0; k < 16; k++) {
		const int v = c[k];
		const bool live = (alive >> k) & 1;
		const int sgn = (int)((unsigned)v >> 31);
		const unsigned uf = uview<SH>(2 * abs(v) + sgn);        // s2u_ (utils.h:95-99), C-typed
		const bool cand = live && uf < uthr0;
		const int qq = ((int)(uf >> 1) * iQ + (1 << 15)) >> 16;  // int arithmetic as in the reference (:172)
		c[k] = !live ? 0 : cand ? (2 | sgn) : ((qq << 1) | sgn);
		key[k] = cand ? (int)(uf << 4 | (unsigned)(15 - k)) : 0;
		cnt += (live && !cand) ? 1 : 0;
		nc += cand ? 1 : 0;
	}
	const int ncm = __reduce_max_sync(FULL, nc);  // largest candidate count among this warp's blocks
	if (ncm > 0) {
		int kstar = 0x7fffffff, m = 0;
		if (ncm == 1) {
			// at most one candidate per block (the common case on chroma planes): it is rank 0, so it
			// survives iff f >= thr[cnt] -- no sort needed
			int s0 = 0;
#pragma unroll
			for (int k = 0; k < 16; k++) s0 = max(s0, key[k]);
			bool pass;
			if (qb->fast) pass = s0 >= qb->kthr[cnt];
			else pass = s0 != 0 && !(TR<SH>(s0 >> 4) < qb->thr[cnt & 15]);
			if (pass && s0 != 0) { kstar = s0; m = 1; }
		} else {
			int s[16];
#pragma unroll
			for (int k = 0; k < 16; k++) s[k] = key[k];
			sort16_desc(s);
			if (qb->fast) {
				const int *kt = qb->kthr + cnt;
#pragma unroll
				for (int i = 0; i < 16; i++) {
					if ((i & 3) == 0 && i >= ncm) break;  // warp-uniform: ranks beyond the largest candidate count are empty
					const bool pass = s[i] >= kt[i];  // non-candidates (key 0) never pass: kthr > 0 whenever candidates exist
					kstar = pass ? s[i] : kstar;
					m = pass ? i + 1 : m;
				}
			} else {
#pragma unroll
				for (int i = 0; i < 16; i++) {
					if (s[i] != 0) {
						const int fs = TR<SH>(s[i] >> 4);  // signed C compare, :191
						if (!(fs < qb->thr[(cnt + i) & 15])) { kstar = s[i]; m = i + 1; }
					}
				}
			}
		}
#pragma unroll
		for (int k = 0; k < 16; k++)
			c[k] = ((unsigned)(key[k] - 1) < (unsigned)(kstar - 1)) ? 0 : c[k];  // candidate ranked below the last survivor
		cnt += m;
	}
	return cnt;
}

// CBand::TSUQ<C> (band.h:65-92): uniform dead-zone quantiser used for the LL band on encode.
template <bool SH>
__device__ __forceinline__ int tsuq1(int c, int T, int iQ)
{
	if ((unsigned)(c + T) <= (unsigned)(2 * T)) return 0;
	return TR<SH>((c * iQ + (1 << 15)) >> 16);
}

}  // namespace ric
