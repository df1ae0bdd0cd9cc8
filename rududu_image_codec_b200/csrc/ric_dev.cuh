// ric_dev.cuh -- device-side building blocks shared by the forward and inverse level kernels.
//
// Exact integer semantics of the reference:
//   lifting        src/lib/wavelet2d.cpp:307-405 (9/7), :593-634 (5/3)
//   block quantiser src/lib/bandcodec.cpp:115-237
// "C-typed store" truncation is TR<SH>() (SH: the level works on `short`).  All >> are arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ric {

enum { T97 = 0, T53 = 1 };

// Strip geometry: one warp owns STRIP_W output columns and loads 8 more on each side
// (lane 0 and lane 31 are halo lanes); every lane holds 8 consecutive columns.
constexpr int STRIP_W = 240;
constexpr int LANE_W = 8;
constexpr unsigned FULL = 0xffffffffu;

template <bool SH>
__device__ __forceinline__ int TR(int v) { return SH ? (int)(short)v : v; }

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

// ---- forward lifting steps (Appendix A.1 / A.3).  x: centre, l/r: neighbours (proper C values).
// KEEP: result feeds a shift next, so it must be truncated now; otherwise truncation is deferred
// to the next truncating use (the low 16 bits are always right: + - * are ring homomorphisms).
template <bool SH, int TRANS>
__device__ __forceinline__ int fS1(int x, int l, int r)
{
	if (TRANS == T97) { int t = TR<SH>(l + r); return TR<SH>(x - (t + (t >> 1))); }
	return TR<SH>(x - ((l + r) >> 1));
}
template <bool SH, int TRANS>
__device__ __forceinline__ int fS1_first(int x, int r) { return TRANS == T97 ? TR<SH>(x - 3 * r) : TR<SH>(x - r); }
template <bool SH, int TRANS>
__device__ __forceinline__ int fS1_last(int x, int l) { return TRANS == T97 ? TR<SH>(x - 3 * l) : TR<SH>(x - l); }

template <bool SH, int TRANS>
__device__ __forceinline__ int fS2(int x, int l, int r)
{
	if (TRANS == T97) return TR<SH>(x - ((l + r) >> 4));
	return TR<SH>(x + ((l + r) >> 2));
}
template <bool SH, int TRANS>
__device__ __forceinline__ int fS2_last(int x, int l) { return TRANS == T97 ? TR<SH>(x - (l >> 3)) : TR<SH>(x + (l >> 1)); }

template <bool SH, int TRANS>
__device__ __forceinline__ int fS3(int x, int l, int r) { return TRANS == T97 ? TR<SH>(x + m08i(l + r)) : x; }
template <bool SH, int TRANS>
__device__ __forceinline__ int fS3_edge(int x, int n) { return TRANS == T97 ? TR<SH>(x + 2 * m08c<SH>(n)) : x; }

template <bool SH, int TRANS>
__device__ __forceinline__ int fS4(int x, int l, int r)
{
	if (TRANS == T97) { int t = TR<SH>(l + r); return TR<SH>(x + ((t >> 1) - (t >> 5))); }
	return x;
}
template <bool SH, int TRANS>
__device__ __forceinline__ int fS4_last(int x, int l) { return TRANS == T97 ? TR<SH>(x + (l - (l >> 4))) : x; }

// ---- inverse lifting steps (Appendix A.2 / A.3): U4 undoes S4, ... U1 undoes S1.
template <bool SH, int TRANS>
__device__ __forceinline__ int iU4(int x, int l, int r)
{
	if (TRANS == T97) { int t = TR<SH>(l + r); return TR<SH>(x - ((t >> 1) - (t >> 5))); }
	return x;
}
template <bool SH, int TRANS>
__device__ __forceinline__ int iU4_last(int x, int l) { return TRANS == T97 ? TR<SH>(x - (l - (l >> 4))) : x; }
template <bool SH, int TRANS>
__device__ __forceinline__ int iU3(int x, int l, int r) { return TRANS == T97 ? TR<SH>(x - m08i(l + r)) : x; }
template <bool SH, int TRANS>
__device__ __forceinline__ int iU3_edge(int x, int n) { return TRANS == T97 ? TR<SH>(x - 2 * m08c<SH>(n)) : x; }
template <bool SH, int TRANS>
__device__ __forceinline__ int iU2(int x, int l, int r)
{
	if (TRANS == T97) return TR<SH>(x + ((l + r) >> 4));
	return TR<SH>(x - ((l + r) >> 2));
}
template <bool SH, int TRANS>
__device__ __forceinline__ int iU2_last(int x, int l) { return TRANS == T97 ? TR<SH>(x + (l >> 3)) : TR<SH>(x - (l >> 1)); }
template <bool SH, int TRANS>
__device__ __forceinline__ int iU1(int x, int l, int r)
{
	if (TRANS == T97) { int t = TR<SH>(l + r); return TR<SH>(x + (t + (t >> 1))); }
	return TR<SH>(x + ((l + r) >> 1));
}
template <bool SH, int TRANS>
__device__ __forceinline__ int iU1_first(int x, int r) { return TRANS == T97 ? TR<SH>(x + 3 * r) : TR<SH>(x + r); }
template <bool SH, int TRANS>
__device__ __forceinline__ int iU1_last(int x, int l) { return TRANS == T97 ? TR<SH>(x + 3 * l) : TR<SH>(x + l); }

// ---- horizontal passes on the 8 columns a lane holds (cb = absolute column of v[0], even).
// Neighbours across lanes come from warp shuffles.  EDGE selects the absolute-coordinate edge
// formulas (column 0 and column w-1); interior strips instantiate EDGE=false.
template <bool SH, int TRANS, bool EDGE>
__device__ __forceinline__ void row_fwd(int (&v)[8], int cb, int w)
{
	int nl, nr;
	// S1: even columns
	nl = __shfl_up_sync(FULL, v[7], 1);
#pragma unroll
	for (int k = 0; k < 8; k += 2) {
		int l = k ? v[k - 1] : nl, r = v[k + 1], c = cb + k;
		int g = fS1<SH, TRANS>(v[k], l, r);
		if (EDGE) {
			if (c == 0) g = fS1_first<SH, TRANS>(v[k], r);
			else if (c == w - 1) g = fS1_last<SH, TRANS>(v[k], l);
		}
		v[k] = g;
	}
	// S2: odd columns
	nr = __shfl_down_sync(FULL, v[0], 1);
#pragma unroll
	for (int k = 1; k < 8; k += 2) {
		int l = v[k - 1], r = k < 7 ? v[k + 1] : nr, c = cb + k;
		int g = fS2<SH, TRANS>(v[k], l, r);
		if (EDGE && c == w - 1) g = fS2_last<SH, TRANS>(v[k], l);
		v[k] = g;
	}
	if (TRANS != T97) return;
	// S3
	nl = __shfl_up_sync(FULL, v[7], 1);
#pragma unroll
	for (int k = 0; k < 8; k += 2) {
		int l = k ? v[k - 1] : nl, r = v[k + 1], c = cb + k;
		int g = fS3<SH, TRANS>(v[k], l, r);
		if (EDGE) {
			if (c == 0) g = fS3_edge<SH, TRANS>(v[k], r);
			else if (c == w - 1) g = fS3_edge<SH, TRANS>(v[k], l);
		}
		v[k] = g;
	}
	// S4
	nr = __shfl_down_sync(FULL, v[0], 1);
#pragma unroll
	for (int k = 1; k < 8; k += 2) {
		int l = v[k - 1], r = k < 7 ? v[k + 1] : nr, c = cb + k;
		int g = fS4<SH, TRANS>(v[k], l, r);
		if (EDGE && c == w - 1) g = fS4_last<SH, TRANS>(v[k], l);
		v[k] = g;
	}
}

template <bool SH, int TRANS, bool EDGE>
__device__ __forceinline__ void row_inv(int (&v)[8], int cb, int w)
{
	int nl, nr;
	if (TRANS == T97) {
		// U4: odd columns
		nr = __shfl_down_sync(FULL, v[0], 1);
#pragma unroll
		for (int k = 1; k < 8; k += 2) {
			int l = v[k - 1], r = k < 7 ? v[k + 1] : nr, c = cb + k;
			int g = iU4<SH, TRANS>(v[k], l, r);
			if (EDGE && c == w - 1) g = iU4_last<SH, TRANS>(v[k], l);
			v[k] = g;
		}
		// U3: even columns
		nl = __shfl_up_sync(FULL, v[7], 1);
#pragma unroll
		for (int k = 0; k < 8; k += 2) {
			int l = k ? v[k - 1] : nl, r = v[k + 1], c = cb + k;
			int g = iU3<SH, TRANS>(v[k], l, r);
			if (EDGE) {
				if (c == 0) g = iU3_edge<SH, TRANS>(v[k], r);
				else if (c == w - 1) g = iU3_edge<SH, TRANS>(v[k], l);
			}
			v[k] = g;
		}
	}
	// U2: odd columns
	nr = __shfl_down_sync(FULL, v[0], 1);
#pragma unroll
	for (int k = 1; k < 8; k += 2) {
		int l = v[k - 1], r = k < 7 ? v[k + 1] : nr, c = cb + k;
		int g = iU2<SH, TRANS>(v[k], l, r);
		if (EDGE && c == w - 1) g = iU2_last<SH, TRANS>(v[k], l);
		v[k] = g;
	}
	// U1: even columns
	nl = __shfl_up_sync(FULL, v[7], 1);
#pragma unroll
	for (int k = 0; k < 8; k += 2) {
		int l = k ? v[k - 1] : nl, r = v[k + 1], c = cb + k;
		int g = iU1<SH, TRANS>(v[k], l, r);
		if (EDGE) {
			if (c == 0) g = iU1_first<SH, TRANS>(v[k], r);
			else if (c == w - 1) g = iU1_last<SH, TRANS>(v[k], l);
		}
		v[k] = g;
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

// ---- encode quantiser (CBandCodec::tsuqBlock, bandcodec.cpp:159-237) ----------------------------
// Host-computed scalars of one band (buildTree :243-247, makeThres :149-157).
struct QuantBand {
	int Q, iQ, T, Te;  // T = Q>>1 (full blocks), Te = (Q+((Q-(Q>>2))>>1))>>1 (partial blocks)
	int thr[16];
};

template <bool SH>
__device__ __forceinline__ unsigned uview(int v) { return SH ? (unsigned)(v & 0xFFFF) : (unsigned)v; }

__device__ __forceinline__ int fold_s2u(int c)  // s2u_, utils.h:95-99
{
	int m = c >> 31;
	return (2 * c + m) ^ (m * 2);
}

// Batcher odd-even merge sort of 16 keys, descending: 63 compare-exchanges written out so that the
// keys provably stay in registers (a rolled network would index them dynamically -> local memory).
__device__ __forceinline__ void sort16_desc(int (&s)[16])
{
#define RIC_CE(i, j) { const int a_ = s[i], b_ = s[j]; s[i] = max(a_, b_); s[j] = min(a_, b_); }
	RIC_CE(0, 1) RIC_CE(2, 3) RIC_CE(4, 5) RIC_CE(6, 7) RIC_CE(8, 9) RIC_CE(10, 11) RIC_CE(12, 13)
	RIC_CE(14, 15) RIC_CE(0, 2) RIC_CE(1, 3) RIC_CE(4, 6) RIC_CE(5, 7) RIC_CE(8, 10) RIC_CE(9, 11)
	RIC_CE(12, 14) RIC_CE(13, 15) RIC_CE(1, 2) RIC_CE(5, 6) RIC_CE(9, 10) RIC_CE(13, 14) RIC_CE(0, 4)
	RIC_CE(1, 5) RIC_CE(2, 6) RIC_CE(3, 7) RIC_CE(8, 12) RIC_CE(9, 13) RIC_CE(10, 14) RIC_CE(11, 15)
	RIC_CE(2, 4) RIC_CE(3, 5) RIC_CE(10, 12) RIC_CE(11, 13) RIC_CE(1, 2) RIC_CE(3, 4) RIC_CE(5, 6)
	RIC_CE(9, 10) RIC_CE(11, 12) RIC_CE(13, 14) RIC_CE(0, 8) RIC_CE(1, 9) RIC_CE(2, 10) RIC_CE(3, 11)
	RIC_CE(4, 12) RIC_CE(5, 13) RIC_CE(6, 14) RIC_CE(7, 15) RIC_CE(4, 8) RIC_CE(5, 9) RIC_CE(6, 10)
	RIC_CE(7, 11) RIC_CE(2, 4) RIC_CE(3, 5) RIC_CE(6, 8) RIC_CE(7, 9) RIC_CE(10, 12) RIC_CE(11, 13)
	RIC_CE(1, 2) RIC_CE(3, 4) RIC_CE(5, 6) RIC_CE(7, 8) RIC_CE(9, 10) RIC_CE(11, 12) RIC_CE(13, 14)
#undef RIC_CE
}

// Quantise one 4x4 block held in registers (c[4*row+col], proper C values), in place, returning
// the number of non-zero outputs.  bw/bh: valid columns/rows (4,4 = full block -> rank-threshold
// path; otherwise the plain dead-zone path of the partial-block overload).  qb points to shared
// memory.  The candidate ranking reproduces the reference's stable insertion sort (descending
// unsigned value, ties in raster order) through distinct keys (value<<4 | 15-pos).
template <bool SH>
__device__ __forceinline__ int quant_block(int (&c)[16], const QuantBand *qb, int bw, int bh)
{
	const int Q_iQ = qb->iQ;
	int cnt = 0;
	if (bw == 4 && bh == 4) {
		const int T = qb->T, thr0 = qb->thr[0];
		const unsigned uthr0 = uview<SH>(thr0);
		int key[16];
		int nc = 0;
#pragma unroll
		for (int k = 0; k < 16; k++) {
			int v = c[k];
			key[k] = 0;
			if ((unsigned)(v + T) <= (unsigned)(2 * T)) { c[k] = 0; continue; }
			int f = TR<SH>(fold_s2u(v));
			unsigned uf = uview<SH>(f);
			if (uf < uthr0) {
				key[k] = (int)(uf << 4) | (15 - k);
				c[k] = f;
				nc++;
			} else {
				cnt++;
				int a = (int)(uf >> 1);
				int q = (a * Q_iQ + (1 << 15)) >> 16;
				c[k] = TR<SH>((q << 1) | (f & 1));
			}
		}
		if (__any_sync(__activemask(), nc > 0)) {
			int s[16];
#pragma unroll
			for (int k = 0; k < 16; k++) s[k] = key[k];
			sort16_desc(s);
			int kstar = 0x7fffffff;
#pragma unroll
			for (int i = 0; i < 16; i++) {
				if (s[i] != 0) {
					int fs = TR<SH>(s[i] >> 4);
					if (!(fs < qb->thr[(cnt + i) & 15])) kstar = s[i];
				}
			}
#pragma unroll
			for (int k = 0; k < 16; k++) {
				if (key[k] != 0) {
					if (key[k] >= kstar) { c[k] = 2 | (c[k] & 1); cnt++; }
					else c[k] = 0;
				}
			}
		}
	} else {
		const int T = qb->Te;
#pragma unroll
		for (int k = 0; k < 16; k++) {
			if ((k & 3) >= bw || (k >> 2) >= bh) continue;
			int v = c[k];
			if ((unsigned)(v + T) <= (unsigned)(2 * T)) { c[k] = 0; continue; }
			int f = TR<SH>(fold_s2u(v));
			unsigned uf = uview<SH>(f);
			cnt++;
			int a = (int)(uf >> 1);
			int q = (a * Q_iQ + (1 << 15)) >> 16;
			c[k] = TR<SH>((q << 1) | (f & 1));
		}
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
