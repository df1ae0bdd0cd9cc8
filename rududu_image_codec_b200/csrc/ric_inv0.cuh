// ric_inv0.cuh -- the finest inverse level (to 8-bit pixels) with two samples per register.
//
// Same job as inv_level_kernel<true, T97, DST_U8_*> (ric_inv.cuh): CBand::TSUQi (src/lib/band.h:94-107), one level
// of CWavelet2D::Transform97I (src/lib/wavelet2d.cpp:361-405, 494-591), YCoCgtoRGB / gray un-shift and clip
// (src/ric/ric.cpp:93-112, 237-240) -- same strips, segments, 3-warp RGB groups and streaming structure; what
// changes is the arithmetic, which is the packed-linear form of ric_swar.cuh:
//
//  * The band rows arrive as pairs of int16 (4 samples per 8-byte load) and stay pairs: dequantisation is one
//    multiply per pair, the column pass works on (column c, column c+2) pairs exactly as loaded, one PRMT per
//    register regroups the two finished rows as (row 2t-4, row 2t-3) pairs of one column for the row pass (both
//    rows lifted by the same instructions, half the shuffles), and colour conversion / clipping / byte packing
//    run on those pairs too.
//  * Packed arithmetic is only the reference's arithmetic while nothing wraps in int16.  Every register an
//    iteration produces carries the same constant, so ONE OR over all of them plus one mask tests "everything in
//    [-8192, 8191]" (sw::GUARD_I), which bounds every intermediate of the following step; the coefficients are
//    range-checked before the multiply.  An iteration (or row pair) that fails, and the first / last rows of the
//    image, run the scalar, exactly-wrapping steps of ric_dev.cuh instead (inv0_v_slow / inv0_h_slow).  Decoded natural images never leave the range; arbitrary coefficient arenas do,
//    and still decode bit-exactly.
//
// Host side (ric_b200.cu launch_inverse) uses this kernel when: 9/7, short level 0 fed by a short level 1,
// u8 output, q != 0.
#pragma once
#include "ric_inv.cuh"
#include "ric_swar.cuh"

namespace ric {

struct Inv0Scratch {  // lane-private exchange with the scalar path: 32 words per lane
	unsigned w[32][32];
};

// Scalar column-pass iteration.  in: w[0..3] new even row, [4..7] new odd row, [8..11] se0, [12..15] so4,
// [16..19] se3, [20..23] so2 (pairs (c, c+2), constant G, any int16 value).  out: w[8..23] the new state (same
// form), w[24..27] / w[28..31] the finished even / odd row as two's-complement pairs.  Returns whether the new
// state is inside the packed path's range.
__device__ __noinline__ bool inv0_v_slow(unsigned (*w)[32], int lane, int t, int h)
{
	using namespace sw;
	int xe[8], xo[8], se0[8], so4[8], se3[8], so2[8];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		xe[2 * i] = dec_lo(w[i][lane], G); xe[2 * i + 1] = dec_hi(w[i][lane], G);
		xo[2 * i] = dec_lo(w[4 + i][lane], G); xo[2 * i + 1] = dec_hi(w[4 + i][lane], G);
		se0[2 * i] = dec_lo(w[8 + i][lane], G); se0[2 * i + 1] = dec_hi(w[8 + i][lane], G);
		so4[2 * i] = dec_lo(w[12 + i][lane], G); so4[2 * i + 1] = dec_hi(w[12 + i][lane], G);
		se3[2 * i] = dec_lo(w[16 + i][lane], G); se3[2 * i + 1] = dec_hi(w[16 + i][lane], G);
		so2[2 * i] = dec_lo(w[20 + i][lane], G); so2[2 * i + 1] = dec_hi(w[20 + i][lane], G);
	}
	const int r4 = 2 * t - 1, r3 = 2 * t - 2, r2 = 2 * t - 3, r1 = 2 * t - 4;
	if (r4 >= 0 && r4 < h) vU4<true, T97, true>(xo, se0, xe, false, r4 == h - 1);
	if (r3 >= 0 && r3 < h) vU3<true, T97, true>(se0, so4, xo, r3 == 0, r3 == h - 1);
	if (r2 >= 0 && r2 < h) vU2<true, T97, true>(so4, se3, se0, false, r2 == h - 1);
	if (r1 >= 0 && r1 < h) vU1<true, T97, true>(se3, so2, so4, r1 == 0, r1 == h - 1);
	bool ok = true;
#pragma unroll
	for (int i = 0; i < 4; i++) {
		// U2 / U1 results arrive un-truncated (ric_dev.cuh): the (short) casts are their C-typed stores
		const int e0 = (short)se3[2 * i], e1 = (short)se3[2 * i + 1], o0 = (short)so4[2 * i], o1 = (short)so4[2 * i + 1];
		w[24 + i][lane] = (unsigned)(e0 & 0xFFFF) | (unsigned)e1 << 16;  // finished even row 2t-4
		w[28 + i][lane] = (unsigned)(o0 & 0xFFFF) | (unsigned)o1 << 16;  // finished odd row 2t-3
		// rotated state: se0 <- xe, so4 <- xo (U4'd), se3 <- se0 (U3'd), so2 <- so4 (U2'd)
		const int a0 = (short)xe[2 * i], a1 = (short)xe[2 * i + 1], b0 = (short)xo[2 * i], b1 = (short)xo[2 * i + 1];
		const int c0 = (short)se0[2 * i], c1 = (short)se0[2 * i + 1];
		w[8 + i][lane] = enc(a0, a1, G); w[12 + i][lane] = enc(b0, b1, G); w[16 + i][lane] = enc(c0, c1, G); w[20 + i][lane] = enc(o0, o1, G);
		ok = ok && abs(a0) <= 8191 && abs(a1) <= 8191 && abs(b0) <= 8191 && abs(b1) <= 8191 && abs(c0) <= 8191 && abs(c1) <= 8191 &&
		     abs(o0) <= 8191 && abs(o1) <= 8191;
	}
	return ok;
}

// Scalar row pass of a finished row pair.  in: w[24..27] / w[28..31] even / odd row as two's-complement (c, c+2)
// pairs.  out: w[0..7] = column c of (even row | odd row << 16), two's complement.  Warp-collective (shuffles).
__device__ __noinline__ void inv0_h_slow(unsigned (*w)[32], int lane, int cb, int width, bool edge)
{
	const EdgeX ex = make_edge_x(cb, width, edge);
	int o[2][8];
#pragma unroll
	for (int r = 0; r < 2; r++) {
		const unsigned a = w[24 + 4 * r][lane], b = w[25 + 4 * r][lane], c = w[26 + 4 * r][lane], d = w[27 + 4 * r][lane];
		o[r][0] = (short)(a & 0xFFFF); o[r][2] = (int)a >> 16; o[r][4] = (short)(b & 0xFFFF); o[r][6] = (int)b >> 16;
		o[r][1] = (short)(c & 0xFFFF); o[r][3] = (int)c >> 16; o[r][5] = (short)(d & 0xFFFF); o[r][7] = (int)d >> 16;
		if (edge) row_inv<true, T97, true>(o[r], ex); else row_inv<true, T97, false>(o[r], ex);
	}
#pragma unroll
	for (int k = 0; k < 8; k++) w[k][lane] = (unsigned)(o[0][k] & 0xFFFF) | (unsigned)o[1][k] << 16;
}

// packed row pass, interior strips: X[c] = column c of (even row | odd row << 16), constant G; returns the OR of
// everything it produced (for the guard)
__device__ __forceinline__ unsigned inv0_row_pass(unsigned (&X)[8])
{
	using namespace sw;
	unsigned acc = 0, nl, nr;
	nr = __shfl_down_sync(FULL, X[0], 1);
#pragma unroll
	for (int k = 1; k < 8; k += 2) { X[k] = u4(X[k], X[k - 1], k < 7 ? X[k + 1] : nr); acc |= X[k]; }
	nl = __shfl_up_sync(FULL, X[7], 1);
#pragma unroll
	for (int k = 0; k < 8; k += 2) { X[k] = u3(X[k], k ? X[k - 1] : nl, X[k + 1]); acc |= X[k]; }
	nr = __shfl_down_sync(FULL, X[0], 1);
#pragma unroll
	for (int k = 1; k < 8; k += 2) { X[k] = u2(X[k], X[k - 1], k < 7 ? X[k + 1] : nr); acc |= X[k]; }
	nl = __shfl_up_sync(FULL, X[7], 1);
#pragma unroll
	for (int k = 0; k < 8; k += 2) { X[k] = u1(X[k], k ? X[k - 1] : nl, X[k + 1]); acc |= X[k]; }
	return acc;
}

// The same with the reference's edge formulas (wavelet2d.cpp:365-368,388-403) for strips that hold column 0 or
// column w-1: every element is first computed with the interior formula, then the edge elements are overwritten
// (cf. row_inv in ric_dev.cuh).  Out of line: 2 strips of 16 at 4K, and the hot loop has to stay small.
__device__ __noinline__ unsigned inv0_row_pass_edge(unsigned *Xp, int cb, int w)
{
	using namespace sw;
	const EdgeX e = make_edge_x(cb, w, true);
	unsigned X[8], o[8], acc = 0, nl, nr;
#pragma unroll
	for (int k = 0; k < 8; k++) X[k] = Xp[k];
	nr = __shfl_down_sync(FULL, X[0], 1);
#pragma unroll
	for (int k = 1; k < 8; k += 2) { o[k] = X[k]; X[k] = u4(X[k], X[k - 1], k < 7 ? X[k + 1] : nr); }
	if (e.last) {
#pragma unroll
		for (int k = 1; k < 8; k += 2) if (e.kl == k) X[k] = u4_last(o[k], X[k - 1]);
	}
	nl = __shfl_up_sync(FULL, X[7], 1);
#pragma unroll
	for (int k = 0; k < 8; k += 2) { o[k] = X[k]; X[k] = u3(X[k], k ? X[k - 1] : nl, X[k + 1]); }
	if (e.first) X[0] = u3_edge(o[0], X[1]);
	if (e.last) {
#pragma unroll
		for (int k = 0; k < 8; k += 2) if (e.kl == k) X[k] = u3_edge(o[k], k ? X[k - 1] : nl);
	}
	acc |= X[1] | X[3] | X[5] | X[7] | X[0] | X[2] | X[4] | X[6];
	nr = __shfl_down_sync(FULL, X[0], 1);
#pragma unroll
	for (int k = 1; k < 8; k += 2) { o[k] = X[k]; X[k] = u2(X[k], X[k - 1], k < 7 ? X[k + 1] : nr); }
	if (e.last) {
#pragma unroll
		for (int k = 1; k < 8; k += 2) if (e.kl == k) X[k] = u2_last(o[k], X[k - 1]);
	}
	nl = __shfl_up_sync(FULL, X[7], 1);
#pragma unroll
	for (int k = 0; k < 8; k += 2) { o[k] = X[k]; X[k] = u1(X[k], k ? X[k - 1] : nl, X[k + 1]); }
	if (e.first) X[0] = u1_edge(o[0], X[1]);
	if (e.last) {
#pragma unroll
		for (int k = 0; k < 8; k += 2) if (e.kl == k) X[k] = u1_edge(o[k], k ? X[k - 1] : nl);
	}
#pragma unroll
	for (int k = 0; k < 8; k++) { acc |= X[k]; Xp[k] = X[k]; }
	return acc;
}

// four pixel pairs (value + PIXK per half, low byte = pixel) -> the bytes of 4 columns of the even / odd row
__device__ __forceinline__ void pix_words(unsigned a, unsigned b, unsigned c, unsigned d, unsigned &even, unsigned &odd)
{
	const unsigned p01 = __byte_perm(a, b, 0x6240), p23 = __byte_perm(c, d, 0x6240);  // a.b0 b.b0 a.b2 b.b2
	even = __byte_perm(p01, p23, 0x5410);
	odd = __byte_perm(p01, p23, 0x7632);
}

// staged rows of one iteration: [set][slot][plane][0..1][lane] = X[0..7]; fmt: 0 constant-G pairs in range, 1 two's complement
typedef int Inv0Fmt[2][3][3];

__device__ __forceinline__ void inv0_scalar_pair(unsigned X, int fmt, int &e, int &o)
{
	if (fmt) { e = (short)(X & 0xFFFF); o = (int)X >> 16; }
	else { e = sw::dec_lo(X, sw::G); o = sw::dec_hi(X, sw::G); }
}

template <int DST>
__device__ __forceinline__ void inv0_job(const InvParams &P, long long job, RgbStage *stage, Inv0Fmt *sfmt, Inv0Scratch &scr, int grp,
                                         int wig, int lane)
{
	using namespace sw;
	constexpr bool RGB = DST == DST_U8_RGB;
	const int plane = RGB ? wig : 0;
	const int sx = (int)(job % P.nstrips); job /= P.nstrips;
	const int sy = (int)(job % P.nsegs);
	const int img = (int)(job / P.nsegs);

	const int w = P.w, h = P.h;
	const int x0 = sx * STRIP_W;
	const int cb = x0 - LANE_W + lane * LANE_W;
	const bool col_ok = cb >= 0 && cb < w;
	const bool lane_out = lane >= 1 && lane <= 30 && cb < w;
	const bool edge_x = (x0 == 0) || (w <= x0 + STRIP_W + LANE_W);
	const int y0 = sy * P.seg_rows;
	const int y1 = min(h, y0 + P.seg_rows);
	const int bc = cb >> 1;
	const char *arena = P.arena + img * P.arena_img_stride + plane * P.arena_plane_stride;
	const char *llp = (const char *)P.ll + (img * P.ll_img_stride + plane * P.ll_plane_stride) * 2;

	const unsigned qd = (unsigned)P.dq[plane][0], qh = (unsigned)P.dq[plane][1], qv = (unsigned)P.dq[plane][2];
	const unsigned kd = G - OB * qd, kh = G - OB * qh, kv = G - OB * qv, kl = G - OB;
	const int clim = 8191 / (int)max(qd, max(qh, qv));  // |coefficient| <= clim keeps every product inside the range

	unsigned se0[4], so4[4], se3[4], so2[4];
#pragma unroll
	for (int i = 0; i < 4; i++) se0[i] = so4[i] = se3[i] = so2[i] = G;
	bool state_ok = true;

	const int t_begin = max((y0 >> 1) - 2, 0), t_last = (y1 + 3) >> 1;
	RawIn<true> in;
	InPtrs ip = in_ptrs<true>(P, arena, llp, t_begin, bc);
	load_in<true, T97>(in, P, arena, ip, t_begin, bc, col_ok);

#pragma unroll 1
	for (int t = t_begin; t <= t_last; t++) {
		// TSUQi on the pairs as loaded (LL comes dequantised from the level below)
		unsigned xe[4], xo[4];
		xe[0] = dequant(in.d.x, qd, kd); xe[1] = dequant(in.d.y, qd, kd); xe[2] = dequant(in.h.x, qh, kh); xe[3] = dequant(in.h.y, qh, kh);
		xo[0] = dequant(in.v.x, qv, kv); xo[1] = dequant(in.v.y, qv, kv);
		xo[2] = dequant((unsigned)in.l.x, 1u, kl); xo[3] = dequant((unsigned)in.l.y, 1u, kl);
		// the products are only meaningful if the coefficients were small enough
		const unsigned mx = __vmaxs2(__vimax3_s16x2(in.d.x, in.d.y, in.h.x), __vimax3_s16x2(in.h.y, in.v.x, in.v.y));
		const unsigned mn = __vmins2(__vimin3_s16x2(in.d.x, in.d.y, in.h.x), __vimin3_s16x2(in.h.y, in.v.x, in.v.y));
		const bool in_ok = max((int)(short)(mx & 0xFFFF), (int)mx >> 16) <= clim && min((int)(short)(mn & 0xFFFF), (int)mn >> 16) >= -clim;
		const uint2 cd = in.d, chh = in.h, cv = in.v;
		const int2 cl = make_int2(in.l.x, in.l.y);
		load_in<true, T97>(in, P, arena, ip, t + 1, bc, col_ok);  // prefetch

		const int r1 = 2 * t - 4;
		const bool interior = !((r1 - 1 <= 0) || (2 * t >= h - 1));
		unsigned re[4], ro[4];  // finished even row 2t-4 / odd row 2t-3: constant-G pairs (rows_g) or two's complement
		bool rows_g = false;
		if (interior && state_ok) {
			unsigned nxo[4], nse0[4], nso4[4], nse3[4], acc = 0;
#pragma unroll
			for (int i = 0; i < 4; i++) {
				nxo[i] = u4(xo[i], se0[i], xe[i]);
				nse0[i] = u3(se0[i], so4[i], nxo[i]);
				nso4[i] = u2(so4[i], se3[i], nse0[i]);
				nse3[i] = u1(se3[i], so2[i], nso4[i]);
				acc |= xe[i] | xo[i] | nxo[i] | nse0[i] | nso4[i] | nse3[i];
			}
			if (__all_sync(FULL, in_ok && (acc & GUARD_I) == 0)) {
				rows_g = true;
#pragma unroll
				for (int i = 0; i < 4; i++) {
					re[i] = nse3[i]; ro[i] = nso4[i];
					so2[i] = nso4[i]; se3[i] = nse0[i]; se0[i] = xe[i]; so4[i] = nxo[i];
				}
			}
		}
		if (!rows_g) {
			// scalar iteration from the two's-complement coefficients (exact TSUQi wrap included)
			const int q4[4] = {(int)qd, (int)qh, (int)qv, 1};
			const unsigned c2[8] = {cd.x, cd.y, chh.x, chh.y, cv.x, cv.y, (unsigned)cl.x, (unsigned)cl.y};
#pragma unroll
			for (int i = 0; i < 8; i++) {
				const int q = q4[i >> 1];
				const int lo = (short)((short)(c2[i] & 0xFFFF) * q), hi = (short)(((int)c2[i] >> 16) * q);
				scr.w[i][lane] = enc(lo, hi, G);
			}
#pragma unroll
			for (int i = 0; i < 4; i++) { scr.w[8 + i][lane] = se0[i]; scr.w[12 + i][lane] = so4[i]; scr.w[16 + i][lane] = se3[i]; scr.w[20 + i][lane] = so2[i]; }
			const bool ok = inv0_v_slow(scr.w, lane, t, h);
			state_ok = __all_sync(FULL, ok);
#pragma unroll
			for (int i = 0; i < 4; i++) {
				se0[i] = scr.w[8 + i][lane]; so4[i] = scr.w[12 + i][lane]; se3[i] = scr.w[16 + i][lane]; so2[i] = scr.w[20 + i][lane];
				re[i] = scr.w[24 + i][lane]; ro[i] = scr.w[28 + i][lane];
			}
		}

		if (P.stats && lane == 0) { atomicAdd(P.stats + 2, 1ull); if (!rows_g) atomicAdd(P.stats + 3, 1ull); }
		// row pass on the pair of finished rows: X[c] = column c of (even row | odd row << 16)
		unsigned X[8];
		int fmt = 1;
		if (rows_g) {
			X[0] = prmt(re[0], ro[0], 0x5410u); X[2] = prmt(re[0], ro[0], 0x7632u);
			X[4] = prmt(re[1], ro[1], 0x5410u); X[6] = prmt(re[1], ro[1], 0x7632u);
			X[1] = prmt(re[2], ro[2], 0x5410u); X[3] = prmt(re[2], ro[2], 0x7632u);
			X[5] = prmt(re[3], ro[3], 0x5410u); X[7] = prmt(re[3], ro[3], 0x7632u);
			unsigned Y[8];
#pragma unroll
			for (int k = 0; k < 8; k++) Y[k] = X[k];
			const unsigned acc = edge_x ? inv0_row_pass_edge(Y, cb, w) : inv0_row_pass(Y);  // (warp-uniform)
			if (__all_sync(FULL, (acc & GUARD_I) == 0)) {
				fmt = 0;
#pragma unroll
				for (int k = 0; k < 8; k++) X[k] = Y[k];
			}
		}
		if (P.stats && lane == 0 && fmt) atomicAdd(P.stats + 4, 1ull);
		if (fmt) {
			if (rows_g) {  // (packed rows, scalar row pass: the row pass left the range)
#pragma unroll
				for (int i = 0; i < 4; i++) { scr.w[24 + i][lane] = to_c2(re[i], G); scr.w[28 + i][lane] = to_c2(ro[i], G); }
			} else {
#pragma unroll
				for (int i = 0; i < 4; i++) { scr.w[24 + i][lane] = re[i]; scr.w[28 + i][lane] = ro[i]; }
			}
			inv0_h_slow(scr.w, lane, cb, w, edge_x);
#pragma unroll
			for (int k = 0; k < 8; k++) X[k] = scr.w[k][lane];
		}

		const int slot = RGB ? (t - t_begin) % 3 : 0;
		const int set = RGB ? ((t - t_begin) / 3) & 1 : 0;
		if (RGB) {
			(*stage)[set][slot][plane][0][lane] = make_uint4(X[0], X[1], X[2], X[3]);
			(*stage)[set][slot][plane][1][lane] = make_uint4(X[4], X[5], X[6], X[7]);
			if (lane == 0) (*sfmt)[set][slot][plane] = fmt;
			if (slot == 2 || t == t_last) {
				asm volatile("bar.sync %0, 96;" ::"r"(grp + 1) : "memory");
				const int my = plane;  // this warp converts the rows staged in slot `plane`
				if (my <= slot) {
					const int tt = t - (slot - my);
					const int rowe = 2 * tt - 4, rowo = 2 * tt - 3;
					const bool oe = rowe >= y0 && rowe < y1 && lane_out, oo = rowo >= y0 && rowo < y1 && lane_out;
					unsigned C[3][8];
#pragma unroll
					for (int p = 0; p < 3; p++) {
						const uint4 a = (*stage)[set][my][p][0][lane], b = (*stage)[set][my][p][1][lane];
						C[p][0] = a.x; C[p][1] = a.y; C[p][2] = a.z; C[p][3] = a.w; C[p][4] = b.x; C[p][5] = b.y; C[p][6] = b.z; C[p][7] = b.w;
					}
					const int f0 = (*sfmt)[set][my][0], f1 = (*sfmt)[set][my][1], f2 = (*sfmt)[set][my][2];
					unsigned R[8], Gc[8], B[8];  // pixel pairs: low byte of each half
					if ((f0 | f1 | f2) == 0) {
#pragma unroll
						for (int k = 0; k < 8; k++) ycocg_out(C[0][k], C[1][k], C[2][k], R[k], Gc[k], B[k]);
					} else {
#pragma unroll
						for (int k = 0; k < 8; k++) {  // YCoCgtoRGB<4>, ric.cpp:98-110, on both rows of column k
							int co[2], cg[2], y[2];
							inv0_scalar_pair(C[0][k], f0, co[0], co[1]);
							inv0_scalar_pair(C[1][k], f1, cg[0], cg[1]);
							inv0_scalar_pair(C[2][k], f2, y[0], y[1]);
							unsigned r = 0, g = 0, b = 0;
#pragma unroll
							for (int hf = 0; hf < 2; hf++) {
								int c = co[hf], m = cg[hf], yy = y[hf];
								c = (c + 4) >> 3; m = (m + 4) >> 3; yy = (yy + 8) >> 4;
								yy -= (m >> 1) - 128;
								m += yy;
								yy -= c >> 1;
								c += yy;
								r |= (unsigned)clip255(c) << (16 * hf); g |= (unsigned)clip255(m) << (16 * hf); b |= (unsigned)clip255(yy) << (16 * hf);
							}
							R[k] = r; Gc[k] = g; B[k] = b;
						}
					}
					unsigned char *dpe = (unsigned char *)P.dst + img * P.dst_img_stride + (long long)rowe * P.dst_pitch + cb;
					unsigned char *dpo = dpe + P.dst_pitch;
					unsigned e0, e1, o0, o1;
					pix_words(R[0], R[1], R[2], R[3], e0, o0); pix_words(R[4], R[5], R[6], R[7], e1, o1);
					if (oe) *(uint2 *)dpe = make_uint2(e0, e1);
					if (oo) *(uint2 *)dpo = make_uint2(o0, o1);
					pix_words(Gc[0], Gc[1], Gc[2], Gc[3], e0, o0); pix_words(Gc[4], Gc[5], Gc[6], Gc[7], e1, o1);
					if (oe) *(uint2 *)(dpe + P.dst_plane_stride) = make_uint2(e0, e1);
					if (oo) *(uint2 *)(dpo + P.dst_plane_stride) = make_uint2(o0, o1);
					pix_words(B[0], B[1], B[2], B[3], e0, o0); pix_words(B[4], B[5], B[6], B[7], e1, o1);
					if (oe) *(uint2 *)(dpe + 2 * P.dst_plane_stride) = make_uint2(e0, e1);
					if (oo) *(uint2 *)(dpo + 2 * P.dst_plane_stride) = make_uint2(o0, o1);
				}
			}
		} else {  // gray: clip(128 + ((v + 8) >> 4)), ric.cpp:237-240
			const int rowe = 2 * t - 4, rowo = 2 * t - 3;
			const bool oe = rowe >= y0 && rowe < y1 && lane_out, oo = rowo >= y0 && rowo < y1 && lane_out;
			unsigned V[8];
			if (fmt == 0) {
#pragma unroll
				for (int k = 0; k < 8; k++) V[k] = gray_out(X[k]);
			} else {
#pragma unroll
				for (int k = 0; k < 8; k++) {
					const int e = (short)(X[k] & 0xFFFF), o = (int)X[k] >> 16;
					V[k] = (unsigned)clip255((int)(short)(128 + ((e + 8) >> 4))) | (unsigned)clip255((int)(short)(128 + ((o + 8) >> 4))) << 16;
				}
			}
			unsigned char *dpe = (unsigned char *)P.dst + img * P.dst_img_stride + (long long)rowe * P.dst_pitch + cb;
			unsigned e0, e1, o0, o1;
			pix_words(V[0], V[1], V[2], V[3], e0, o0); pix_words(V[4], V[5], V[6], V[7], e1, o1);
			if (oe) *(uint2 *)dpe = make_uint2(e0, e1);
			if (oo) *(uint2 *)(dpe + P.dst_pitch) = make_uint2(o0, o1);
		}
	}
}

template <int DST>
__global__ void __launch_bounds__(DST == DST_U8_RGB ? 96 : INV_WARPS * 32, DST == DST_U8_RGB ? 6 : 4) inv0_kernel(const __grid_constant__ InvParams P)
{
	constexpr bool RGB = DST == DST_U8_RGB;
	constexpr int NW = RGB ? 3 : INV_WARPS;
	__shared__ uint4 s_stage_raw[RGB ? sizeof(RgbStage) / sizeof(uint4) : 1];
	__shared__ Inv0Fmt s_fmt;
	__shared__ Inv0Scratch s_scr[NW];
	__shared__ unsigned long long s_job;
	RgbStage *s_stage = (RgbStage *)s_stage_raw;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const long long njobs = (long long)P.nstrips * P.nsegs * P.nimages;
	for (;;) {
		unsigned long long job = 0;
		if (RGB) {
			asm volatile("bar.sync 1, 96;" ::: "memory");  // everybody has read the previous job id
			if (wib == 0 && lane == 0) s_job = atomicAdd(P.counter, 1ull);
			asm volatile("bar.sync 1, 96;" ::: "memory");
			job = s_job;
		} else {
			if (lane == 0) job = atomicAdd(P.counter, 1ull);
			job = __shfl_sync(FULL, job, 0);
		}
		if ((long long)job >= njobs) break;
		inv0_job<DST>(P, (long long)job, s_stage, &s_fmt, s_scr[wib], 0, wib, lane);
	}
}

}  // namespace ric
