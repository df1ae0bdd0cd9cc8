// ric_inv.cuh -- one inverse wavelet level, fused with dequantisation on the way in and (finest
// level) the inverse colour transform / level shift / clip to 8 bit on the way out.
//
// Replaces, for one level of every plane of a batch of images:
//   CBand::TSUQi                   src/lib/band.h:94-107 (via CWavelet2D::TSUQi wavelet2d.cpp:248-268)
//   CWavelet2D::Transform97I/53I   src/lib/wavelet2d.cpp:494-591,694-764 (+TransLine*I :361-405,:613-634)
//   int->short narrowing           src/lib/wavelet2d.cpp:971-980
//   YCoCgtoRGB / gray un-shift     src/ric/ric.cpp:93-112,227-240      (DST_U8_*)
//
// Same warp-per-strip streaming structure as the forward kernel (ric_fwd.cuh): vertical inverse
// lifting as a 4-row register pipeline, then the horizontal inverse on each finished row with
// shuffles for the neighbours.  For RGB output the Co, Cg and Y warps of one strip segment form a
// group that swaps finished rows through shared memory every third iteration, so that each warp
// converts one iteration's pixels to RGB (no s16 plane ever goes to HBM).
#pragma once
#include "ric_dev.cuh"
#include "ric_fwd.cuh"  // BandRef

namespace ric {

enum { LLSRC_S16 = 0, LLSRC_S32 = 1, LLSRC_BAND = 2 };
enum { DST_PLANE = 0, DST_U8_GRAY = 1, DST_U8_RGB = 2 };  // DST_PLANE: s16 (short level) / s32 scratch or plane

constexpr int INV_WARPS = 4;       // warps per CTA, independent jobs (DST_PLANE / DST_U8_GRAY)
constexpr int INV_RGB_GROUPS = 1;  // DST_U8_RGB: groups of 3 warps (Co, Cg, Y of one strip segment) per CTA; 6 CTAs per SM

struct InvParams {
	const char *arena;
	long long arena_img_stride, arena_plane_stride;
	const void *ll;            // LL scratch written by the coarser level (LLSRC_S16/S32)
	long long ll_img_stride, ll_plane_stride;
	int ll_pitch;
	int llsrc;                 // LLSRC_*
	void *dst;                 // scratch plane / s16 plane (DST_PLANE) or u8 image
	long long dst_img_stride, dst_plane_stride;
	int dst_pitch;
	BandRef band[3];           // D, H, V
	BandRef lband;             // coarsest LL (LLSRC_BAND)
	int h_row1_off;            // sample offset of H band row 1: H.stride, or D.stride to reproduce the
	                           // reference's Transform53I stride slip (wavelet2d.cpp:715)
	int w, h;
	int nstrips, nsegs, seg_rows, nplanes, nimages;
	int shift;                 // q != 0: undo the fixed-point up-shift and clip (ric.cpp:98-110,237-240)
	int dq[3][4];              // TSUQi multiplier per plane for D,H,V,L (1 = no dequantisation)
	unsigned long long *counter;  // dynamic job fetch (zeroed before the launch)
	unsigned long long *stats;    // optional (profiling): [2] packed-kernel plane iterations, [3] scalar column passes, [4] scalar row passes
};

__device__ __forceinline__ int clip255(int v) { return __vimin_s32_relu(v, 255); }  // max(min(v, 255), 0), one VIMNMX.RELU

// raw (not yet unpacked) inputs of one iteration: D/H row t, V/LL row t-1, 4 samples each
template <bool SH>
struct RawIn {
	typedef typename std::conditional<SH, uint2, int4>::type vec;
	vec d, h, v;
	int4 l;
};

template <bool SH>
__device__ __forceinline__ typename RawIn<SH>::vec ldvec(const char *rowp, int col, bool ok)
{
	typename RawIn<SH>::vec z;
	if (SH) { const uint2 t = ldg_u2_if(rowp + 2 * (long long)col, ok); *(uint2 *)&z = t; }  // predicated, no branch
	else { const uint4 t = ldg_u4_if(rowp + 4 * (long long)col, ok); *(uint4 *)&z = t; }
	return z;
}

// halves of a packed pair, sign-extended on the FMA pipe (IDP.2A.LO.S16.U8 with the byte pairs (1, 0) / (0, 1))
__device__ __forceinline__ int s16_lo_fma(unsigned w) { int d; asm("dp2a.lo.s32.u32 %0, %1, 0x0001, 0;" : "=r"(d) : "r"(w)); return d; }
__device__ __forceinline__ int s16_hi_fma(unsigned w) { int d; asm("dp2a.lo.s32.u32 %0, %1, 0x0100, 0;" : "=r"(d) : "r"(w)); return d; }

// c + a.s16[0] * b.u8[0] + a.s16[1] * b.u8[1], one FMA-pipe instruction (IDP.2A.LO.S16.U8)
__device__ __forceinline__ int dp2a_lo_u(unsigned a, int b, int c)
{
	int d;
	asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
// sign-extended half `hi` of a packed pair plus a constant, one FMA-pipe instruction
__device__ __forceinline__ int s16_plus(unsigned w, int hi, int c)
{
	int d;
	if (hi) asm("dp2a.lo.s32.u32 %0, %1, 0x0100, %2;" : "=r"(d) : "r"(w), "r"(c));
	else asm("dp2a.lo.s32.u32 %0, %1, 0x0001, %2;" : "=r"(d) : "r"(w), "r"(c));
	return d;
}
// four ints clipped to 0..255 and packed, b0 in the low byte: two I2IP.U8.S32.SAT (d = c << 16 | sat(a) << 8 | sat(b))
__device__ __forceinline__ unsigned pack4_sat(int b0, int b1, int b2, int b3)
{
	unsigned t, d;
	asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(b3), "r"(b2), "r"(0));
	asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(b1), "r"(b0), "r"(t));
	return d;
}

template <bool SH>
__device__ __forceinline__ void unpack4(const typename RawIn<SH>::vec &r, int (&o)[4])
{
	if (SH) {
		const uint2 a = *(const uint2 *)&r;
		o[0] = s16_lo_fma(a.x); o[1] = s16_hi_fma(a.x);
		o[2] = s16_lo_fma(a.y); o[3] = s16_hi_fma(a.y);
	} else {
		const int4 a = *(const int4 *)&r;
		o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
	}
}

// Running row pointers of one job's band loads: element (row, bc) of D / H (row t), V / LL (row t - 1), advanced by
// one row per load_in call (the calls of a job come with t, t + 1, t + 2, ...), so the row loop carries four
// 64-bit adds instead of four 64-bit multiplies and their address chains.
struct InPtrs { const char *d, *h, *v, *l; long long sd, sh, sv, sl; };

template <bool SH>
__device__ __forceinline__ InPtrs in_ptrs(const InvParams &P, const char *arena, const char *llp, int t, int bc)
{
	constexpr int ES = SH ? 2 : 4;
	const BandRef &D = P.band[0], &H = P.band[1], &V = P.band[2];
	InPtrs p;
	p.sd = (long long)D.stride * ES; p.sh = (long long)H.stride * ES; p.sv = (long long)V.stride * ES;
	p.d = arena + D.off + t * p.sd + (long long)bc * ES;
	p.h = arena + H.off + t * p.sh + (long long)bc * ES;
	p.v = arena + V.off + (t - 1) * p.sv + (long long)bc * ES;
	if (P.llsrc == LLSRC_BAND) { p.sl = (long long)P.lband.stride * ES; p.l = arena + P.lband.off + (t - 1) * p.sl + (long long)bc * ES; }
	else {
		const int es = P.llsrc == LLSRC_S16 ? 2 : 4;
		p.sl = (long long)P.ll_pitch * es; p.l = llp + (t - 1) * p.sl + (long long)bc * es;
	}
	return p;
}

template <bool SH>
__device__ __forceinline__ typename RawIn<SH>::vec ldvec_at(const char *p, bool ok)
{
	typename RawIn<SH>::vec z;
	if (SH) { const uint2 t = ldg_u2_if(p, ok); *(uint2 *)&z = t; }  // predicated, no branch
	else { const uint4 t = ldg_u4_if(p, ok); *(uint4 *)&z = t; }
	return z;
}

template <bool SH, int TRANS>
__device__ __forceinline__ void load_in(RawIn<SH> &in, const InvParams &P, const char *arena, InPtrs &p, int t, int bc, bool col_ok)
{
	constexpr int ES = SH ? 2 : 4;
	const int h = P.h;
	const bool e_ok = col_ok && 2 * t < h;
	const bool o_ok = col_ok && t >= 1 && 2 * t - 1 < h;
	const BandRef &D = P.band[0], &H = P.band[1], &V = P.band[2];
	in.d = ldvec_at<SH>(p.d, e_ok && bc < D.dimx);
	const char *hp = p.h;
	if (TRANS == T53 && t == 1) hp = arena + H.off + ((long long)P.h_row1_off + bc) * ES;  // the reference's stride slip, SURVEY Q1
	in.h = ldvec_at<SH>(hp, e_ok && bc < H.dimx);
	in.v = ldvec_at<SH>(p.v, o_ok && bc < V.dimx);
	in.l = make_int4(0, 0, 0, 0);
	const bool l_ok = o_ok && bc < (P.w >> 1);
	if (P.llsrc == LLSRC_BAND) {
		typename RawIn<SH>::vec r = ldvec_at<SH>(p.l, l_ok);
		if (SH) { const uint2 a = *(const uint2 *)&r; in.l.x = (int)a.x; in.l.y = (int)a.y; }
		else in.l = *(const int4 *)&r;
	} else if (P.llsrc == LLSRC_S16) {
		const uint2 a = ldg_u2_if(p.l, l_ok);
		in.l.x = (int)a.x; in.l.y = (int)a.y;
	} else {
		const uint4 a = ldg_u4_if(p.l, l_ok);
		in.l = make_int4((int)a.x, (int)a.y, (int)a.z, (int)a.w);
	}
	p.d += p.sd; p.h += p.sh; p.v += p.sv; p.l += p.sl;
}

// unpack + dequantise (TSUQi: pBand[n] *= Quant, truncating store) into interleaved even/odd rows.
// 9/7: the product's own truncation is deferred -- every first use of these rows is a truncating one
// (the (C)(l + r) temporary of U4, the C-typed result of U4 / U3, or iU4_last which truncates its
// operand) -- 5/3 and Haar feed them straight into a shift and truncate here.
template <bool SH, int TRANS>
__device__ __forceinline__ void unpack_in(const RawIn<SH> &in, const InvParams &P, int qd, int qh, int qv, int ql, bool smallq,
                                          int (&xe)[8], int (&xo)[8])
{
	constexpr bool S = SH && TRANS != T97;
	if (SH && TRANS == T97 && smallq) {
		// every multiplier fits a byte (fine quantisers -- the common case -- and q == 0): IDP.2A sign-extends one half
		// of a packed pair AND multiplies it, one FMA-pipe instruction per sample instead of unpack + IMAD
		const uint2 D = *(const uint2 *)&in.d, H = *(const uint2 *)&in.h, V = *(const uint2 *)&in.v;
		const unsigned L0 = (unsigned)in.l.x, L1 = (unsigned)in.l.y;
		xe[0] = dp2a_lo_u(D.x, qd, 0); xe[2] = dp2a_lo_u(D.x, qd << 8, 0); xe[4] = dp2a_lo_u(D.y, qd, 0); xe[6] = dp2a_lo_u(D.y, qd << 8, 0);
		xe[1] = dp2a_lo_u(H.x, qh, 0); xe[3] = dp2a_lo_u(H.x, qh << 8, 0); xe[5] = dp2a_lo_u(H.y, qh, 0); xe[7] = dp2a_lo_u(H.y, qh << 8, 0);
		xo[0] = dp2a_lo_u(V.x, qv, 0); xo[2] = dp2a_lo_u(V.x, qv << 8, 0); xo[4] = dp2a_lo_u(V.y, qv, 0); xo[6] = dp2a_lo_u(V.y, qv << 8, 0);
		xo[1] = dp2a_lo_u(L0, ql, 0); xo[3] = dp2a_lo_u(L0, ql << 8, 0); xo[5] = dp2a_lo_u(L1, ql, 0); xo[7] = dp2a_lo_u(L1, ql << 8, 0);
		return;
	}
	int d[4], hh[4], v[4], l[4];
	unpack4<SH>(in.d, d);
	unpack4<SH>(in.h, hh);
	unpack4<SH>(in.v, v);
	if (P.llsrc == LLSRC_S32 || (!SH && P.llsrc == LLSRC_BAND)) {
		l[0] = in.l.x; l[1] = in.l.y; l[2] = in.l.z; l[3] = in.l.w;
	} else {
		l[0] = s16_lo_fma((unsigned)in.l.x); l[1] = s16_hi_fma((unsigned)in.l.x);
		l[2] = s16_lo_fma((unsigned)in.l.y); l[3] = s16_hi_fma((unsigned)in.l.y);
	}
#pragma unroll
	for (int k = 0; k < 4; k++) {
		xe[2 * k] = TRI<S>(d[k] * qd);
		xe[2 * k + 1] = TRI<S>(hh[k] * qh);
		xo[2 * k] = TRI<S>(v[k] * qv);
		xo[2 * k + 1] = TRI<S>(l[k] * ql);  // also the (C) narrowing of an int LL, wavelet2d.cpp:971-980
	}
}

__device__ __forceinline__ unsigned pack2(int a, int b) { return (unsigned)(a & 0xFFFF) | ((unsigned)b << 16); }
__device__ __forceinline__ unsigned pack4b(unsigned a, unsigned b, unsigned c, unsigned d) { return a | b << 8 | c << 16 | d << 24; }

// RGB staging: [set][slot][plane][row parity][lane] -> 8 s16 samples.  Double-buffered (set = batch
// parity): one barrier per batch of three iterations is enough, because a set is rewritten only after
// every warp of the group has passed the barrier that follows its last reads.
typedef uint4 RgbStage[2][3][3][2][32];

// One job = one (image, row segment, strip) of one plane -- or, for RGB output, of all three planes
// (the three warps of a group run the same job with plane = their index in the group).
// SH: this level works on short; DST: what is written.
template <bool SH, int TRANS, int DST>
__device__ __forceinline__ void inv_job(const InvParams &P, long long job, RgbStage *stage, int grp, int wig, int lane)
{
	constexpr bool RGB = DST == DST_U8_RGB;
	const int jplanes = RGB ? 1 : P.nplanes;
	const int plane = RGB ? wig : (int)(job % jplanes);
	job /= jplanes;
	const int sx = (int)(job % P.nstrips); job /= P.nstrips;
	const int sy = (int)(job % P.nsegs);
	const int img = (int)(job / P.nsegs);

	const int w = P.w, h = P.h;
	const int x0 = sx * STRIP_W;
	const int cb = x0 - LANE_W + lane * LANE_W;
	const bool col_ok = cb >= 0 && cb < w;
	const bool lane_out = lane >= 1 && lane <= 30 && cb < w;
	const EdgeX ex = make_edge_x(cb, w, (x0 == 0) || (w <= x0 + STRIP_W + LANE_W));
	const int y0 = sy * P.seg_rows;
	const int y1 = min(h, y0 + P.seg_rows);
	const int bc = cb >> 1;  // band column of this lane's first even/odd sample
	const char *arena = P.arena + img * P.arena_img_stride + plane * P.arena_plane_stride;
	const char *llp = (const char *)P.ll + (img * P.ll_img_stride + plane * P.ll_plane_stride) * (P.llsrc == LLSRC_S32 ? 4 : 2);

	// vertical state: se0 raw even row 2t-2, so4 U4'd odd row 2t-3, se3 U3'd even row 2t-4, so2 U2'd odd row 2t-5
	int se0[8], so4[8], se3[8], so2[8];
#pragma unroll
	for (int k = 0; k < 8; k++) se0[k] = so4[k] = se3[k] = so2[k] = 0;

	const int t_begin = max((y0 >> 1) - 2, 0), t_last = (y1 + 3) >> 1;
	const int qd = P.dq[plane][0], qh = P.dq[plane][1], qv = P.dq[plane][2];
	const int ql = P.llsrc == LLSRC_BAND ? P.dq[plane][3] : 1;
	// all four multipliers in 1..255 and the LL input a packed short plane: the one-instruction dequantiser of unpack_in
	const bool smallq = SH && P.llsrc != LLSRC_S32 && (unsigned)(qd - 1) < 255u && (unsigned)(qh - 1) < 255u && (unsigned)(qv - 1) < 255u &&
	                    (unsigned)(ql - 1) < 255u;
	RawIn<SH> in;
	InPtrs ip = in_ptrs<SH>(P, arena, llp, t_begin, bc);
	load_in<SH, TRANS>(in, P, arena, ip, t_begin, bc, col_ok);

	int slot_c = 0, set_c = 0;
#pragma unroll 1
	for (int t = t_begin; t <= t_last; t++) {
		int xe[8], xo[8];
		unpack_in<SH, TRANS>(in, P, qd, qh, qv, ql, smallq, xe, xo);
		load_in<SH, TRANS>(in, P, arena, ip, t + 1, bc, col_ok);  // prefetch

		const int r4 = 2 * t - 1, r3 = 2 * t - 2, r2 = 2 * t - 3, r1 = 2 * t - 4;
		const bool edge_y = (r1 - 1 <= 0) || (2 * t >= h - 1);
		if (edge_y) {
			if (r4 >= 0 && r4 < h) vU4<SH, TRANS, true>(xo, se0, xe, false, r4 == h - 1);
			if (r3 >= 0 && r3 < h) vU3<SH, TRANS, true>(se0, so4, xo, r3 == 0, r3 == h - 1);
			if (r2 >= 0 && r2 < h) vU2<SH, TRANS, true>(so4, se3, se0, false, r2 == h - 1);
			if (r1 >= 0 && r1 < h) vU1<SH, TRANS, true>(se3, so2, so4, r1 == 0, r1 == h - 1);
		} else {
			vU4<SH, TRANS, false>(xo, se0, xe, false, false);
			vU3<SH, TRANS, false>(se0, so4, xo, false, false);
			vU2<SH, TRANS, false>(so4, se3, se0, false, false);
			vU1<SH, TRANS, false>(se3, so2, so4, false, false);
		}
		// finished: even row r1 (se3), odd row r2 (so4); both rows' horizontal passes are unrolled so
		// that their dependency chains interleave (this kernel fits the instruction cache either way)
		const int slot = RGB ? slot_c : 0, set = RGB ? set_c : 0;  // (t - t_begin) % 3 and ((t - t_begin) / 3) & 1, kept as counters
#pragma unroll
		for (int half = 0; half < 2; half++) {
			int o[8];
#pragma unroll
			for (int k = 0; k < 8; k++) o[k] = half ? so4[k] : se3[k];
			if (ex.on) row_inv<SH, TRANS, true>(o, ex); else row_inv<SH, TRANS, false>(o, ex);
			const int row = half ? r2 : r1;
			if (DST == DST_U8_RGB) {
				(*stage)[set][slot][plane][half][lane] = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
				continue;
			}
			if (!(row >= y0 && row < y1 && lane_out)) continue;
			if (DST == DST_PLANE) {
				if (SH) {
					short *dp = (short *)P.dst + img * P.dst_img_stride + plane * P.dst_plane_stride + (long long)row * P.dst_pitch + cb;
					*(uint4 *)dp = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
				} else {
					int *dp = (int *)P.dst + img * P.dst_img_stride + plane * P.dst_plane_stride + (long long)row * P.dst_pitch + cb;
					*(int4 *)dp = make_int4(o[0], o[1], o[2], o[3]);
					*(int4 *)(dp + 4) = make_int4(o[4], o[5], o[6], o[7]);
				}
			} else {  // DST_U8_GRAY, ric.cpp:229 / :237-240
				unsigned char *dp = (unsigned char *)P.dst + img * P.dst_img_stride + plane * P.dst_plane_stride +
				                    (long long)row * P.dst_pitch + cb;
				if (P.shift) {
					// clip(128 + ((v + 8) >> 4)) with v the truncated lifting result: the truncation, the rounding
					// constant and the + 128 (= 2048 >> 4) fold into one IDP.2A, the clip and the packing into I2IP
					int b[8];
#pragma unroll
					for (int k = 0; k < 8; k++) b[k] = s16_plus((unsigned)o[k], 0, 8 + 2048) >> 4;
					*(uint2 *)dp = make_uint2(pack4_sat(b[0], b[1], b[2], b[3]), pack4_sat(b[4], b[5], b[6], b[7]));
				} else {
					unsigned b[8];
#pragma unroll
					for (int k = 0; k < 8; k++) b[k] = (unsigned)(TRI<SH>(o[k]) + 128) & 0xFF;
					*(uint2 *)dp = make_uint2(pack4b(b[0], b[1], b[2], b[3]), pack4b(b[4], b[5], b[6], b[7]));
				}
			}
		}
		// rotate
#pragma unroll
		for (int k = 0; k < 8; k++) { so2[k] = so4[k]; se3[k] = se0[k]; se0[k] = xe[k]; so4[k] = xo[k]; }

		if (DST == DST_U8_RGB) {
			// stage this plane's two rows; every third iteration (and at the end) the three warps of the
			// group swap planes through shared memory and each converts one iteration's pixel rows
			if (slot == 2 || t == t_last) {
				asm volatile("bar.sync %0, 96;" ::"r"(grp + 1) : "memory");
				const int my = plane;  // this warp converts the rows staged in slot `plane`
				if (my <= slot) {
					const int tt = t - (slot - my);
#pragma unroll
					for (int half = 0; half < 2; half++) {
						const int row = half ? 2 * tt - 3 : 2 * tt - 4;
						if (!(row >= y0 && row < y1 && lane_out)) continue;
						const uint4 c0 = (*stage)[set][my][0][half][lane], c1 = (*stage)[set][my][1][half][lane],
						            c2 = (*stage)[set][my][2][half][lane];
						// YCoCgtoRGB<shift>, ric.cpp:93-112 (planes 0 Co, 1 Cg, 2 Y)
						unsigned char *dp = (unsigned char *)P.dst + img * P.dst_img_stride + (long long)row * P.dst_pitch + cb;
						if (P.shift) {
							// The rounding constants of the down-shifts -- and Y's + 128 = 2048 >> 4 -- ride on the
							// IDP.2A that sign-extends each staged half; after the down-shifts every intermediate is
							// < 2^14 in magnitude, so the reference's short stores cannot wrap; the clips to 0..255 and
							// the byte packing are I2IP.U8.S32.SAT, two per four pixels.
							int R[8], G[8], B[8];
#pragma unroll
							for (int k = 0; k < 8; k++) {
								const unsigned w0 = k < 2 ? c0.x : k < 4 ? c0.y : k < 6 ? c0.z : c0.w;
								const unsigned w1 = k < 2 ? c1.x : k < 4 ? c1.y : k < 6 ? c1.z : c1.w;
								const unsigned w2 = k < 2 ? c2.x : k < 4 ? c2.y : k < 6 ? c2.z : c2.w;
								int co = s16_plus(w0, k & 1, 4) >> 3, cg = s16_plus(w1, k & 1, 4) >> 3;
								int y = s16_plus(w2, k & 1, 8 + 2048) >> 4;  // ((y + 8) >> 4) + 128
								y -= cg >> 1;
								cg += y;
								y -= co >> 1;
								co += y;
								R[k] = co; G[k] = cg; B[k] = y;
							}
							*(uint2 *)dp = make_uint2(pack4_sat(R[0], R[1], R[2], R[3]), pack4_sat(R[4], R[5], R[6], R[7]));
							*(uint2 *)(dp + P.dst_plane_stride) = make_uint2(pack4_sat(G[0], G[1], G[2], G[3]), pack4_sat(G[4], G[5], G[6], G[7]));
							*(uint2 *)(dp + 2 * P.dst_plane_stride) = make_uint2(pack4_sat(B[0], B[1], B[2], B[3]), pack4_sat(B[4], B[5], B[6], B[7]));
						} else {  // q == 0: every store is a wrapping short store, no clip
							unsigned R[8], G[8], B[8];
#pragma unroll
							for (int k = 0; k < 8; k++) {
								const unsigned w0 = k < 2 ? c0.x : k < 4 ? c0.y : k < 6 ? c0.z : c0.w;
								const unsigned w1 = k < 2 ? c1.x : k < 4 ? c1.y : k < 6 ? c1.z : c1.w;
								const unsigned w2 = k < 2 ? c2.x : k < 4 ? c2.y : k < 6 ? c2.z : c2.w;
								int co = (k & 1) ? (int)w0 >> 16 : (int)(short)(w0 & 0xFFFF);
								int cg = (k & 1) ? (int)w1 >> 16 : (int)(short)(w1 & 0xFFFF);
								int y = (k & 1) ? (int)w2 >> 16 : (int)(short)(w2 & 0xFFFF);
								y = (int)(short)(y - ((cg >> 1) - 128));
								cg = (int)(short)(cg + y);
								y = (int)(short)(y - (co >> 1));
								co = (int)(short)(co + y);
								R[k] = (unsigned)co & 0xFF; G[k] = (unsigned)cg & 0xFF; B[k] = (unsigned)y & 0xFF;
							}
							*(uint2 *)dp = make_uint2(pack4b(R[0], R[1], R[2], R[3]), pack4b(R[4], R[5], R[6], R[7]));
							*(uint2 *)(dp + P.dst_plane_stride) = make_uint2(pack4b(G[0], G[1], G[2], G[3]), pack4b(G[4], G[5], G[6], G[7]));
							*(uint2 *)(dp + 2 * P.dst_plane_stride) = make_uint2(pack4b(B[0], B[1], B[2], B[3]), pack4b(B[4], B[5], B[6], B[7]));
						}
					}
				}
			}
			if (++slot_c == 3) { slot_c = 0; set_c ^= 1; }
		}
	}
}

// Persistent warps with dynamic job fetch (see fwd_level_kernel); for RGB output the three warps of a
// group claim one job together (broadcast through shared memory under the group's named barrier).
template <bool SH, int TRANS, int DST>
__global__ void __launch_bounds__(DST == DST_U8_RGB ? INV_RGB_GROUPS * 96 : INV_WARPS * 32, DST == DST_U8_RGB ? 6 / INV_RGB_GROUPS : 4)
    inv_level_kernel(const __grid_constant__ InvParams P)
{
	constexpr bool RGB = DST == DST_U8_RGB;
	__shared__ uint4 s_stage_raw[RGB ? INV_RGB_GROUPS * (sizeof(RgbStage) / sizeof(uint4)) : 1];
	__shared__ unsigned long long s_job[RGB ? INV_RGB_GROUPS : 1];
	RgbStage *s_stage = (RgbStage *)s_stage_raw;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int grp = RGB ? wib / 3 : 0, wig = RGB ? wib % 3 : 0;
	asm volatile("griddepcontrol.launch_dependents;");  // programmatic dependent launch, see fwd_level_kernel
	asm volatile("griddepcontrol.wait;" ::: "memory");
	const long long njobs = (long long)P.nstrips * (RGB ? 1 : P.nplanes) * P.nsegs * P.nimages;
	for (;;) {
		unsigned long long job = 0;
		if (RGB) {
			asm volatile("bar.sync %0, 96;" ::"r"(grp + 1) : "memory");  // everybody has read the previous job id
			if (wig == 0 && lane == 0) s_job[grp] = atomicAdd(P.counter, 1ull);
			asm volatile("bar.sync %0, 96;" ::"r"(grp + 1) : "memory");
			job = s_job[grp];
		} else {
			if (lane == 0) job = atomicAdd(P.counter, 1ull);
			job = __shfl_sync(FULL, job, 0);
		}
		if ((long long)job >= njobs) break;
		inv_job<SH, TRANS, DST>(P, (long long)job, s_stage + grp, grp, wig, lane);
	}
}

}  // namespace ric
