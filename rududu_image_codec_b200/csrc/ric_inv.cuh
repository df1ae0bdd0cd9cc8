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
// shuffles for the neighbours.  For RGB output one warp carries all three planes of its strip so
// that the colour transform happens in registers (NPL = 3).
#pragma once
#include "ric_dev.cuh"
#include "ric_fwd.cuh"  // BandRef

namespace ric {

enum { LLSRC_S16 = 0, LLSRC_S32 = 1, LLSRC_BAND = 2 };
enum { DST_S16 = 0, DST_S32 = 1, DST_U8_GRAY = 2, DST_U8_RGB = 3 };

struct InvParams {
	const char *arena;
	long long arena_img_stride, arena_plane_stride;
	const void *ll;            // LL scratch written by the coarser level (LLSRC_S16/S32)
	long long ll_img_stride, ll_plane_stride;
	int ll_pitch;
	void *dst;                 // scratch plane (DST_S16/S32) or u8 image
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
};

template <bool SH>
__device__ __forceinline__ void load4(const char *rowp, int col, bool ok, int (&o)[4])
{
	if (!ok) { o[0] = o[1] = o[2] = o[3] = 0; return; }
	if (SH) {
		uint2 a = __ldg((const uint2 *)(rowp + 2 * (long long)col));
		o[0] = (int)(short)(a.x & 0xFFFF); o[1] = (int)a.x >> 16;
		o[2] = (int)(short)(a.y & 0xFFFF); o[3] = (int)a.y >> 16;
	} else {
		int4 a = __ldg((const int4 *)(rowp + 4 * (long long)col));
		o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
	}
}

__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }

// SH: this level works on short; LLSRC: where the LL samples come from; DST: what is written.
template <bool SH, int TRANS, int LLSRC, int DST>
__global__ void __launch_bounds__(128) inv_level_kernel(const __grid_constant__ InvParams P)
{
	constexpr int NPL = DST == DST_U8_RGB ? 3 : 1;
	const int lane = threadIdx.x & 31;
	long long job = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const int jplanes = NPL == 3 ? 1 : P.nplanes;
	const long long njobs = (long long)P.nstrips * jplanes * P.nsegs * P.nimages;
	if (job >= njobs) return;
	const int plane0 = (int)(job % jplanes); job /= jplanes;
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
	const int bc = cb >> 1;  // band column of this lane's first even/odd sample
	constexpr int ES = SH ? 2 : 4;

	// vertical state per plane: se0 raw even row 2t-2, so4 U4'd odd row 2t-3, se3 U3'd even row 2t-4, so2 U2'd odd row 2t-5
	int se0[NPL][8], so4[NPL][8], se3[NPL][8], so2[NPL][8];
#pragma unroll
	for (int p = 0; p < NPL; p++)
#pragma unroll
		for (int k = 0; k < 8; k++) se0[p][k] = so4[p][k] = se3[p][k] = so2[p][k] = 0;

	const int t_begin = max((y0 >> 1) - 2, 0), t_last = (y1 + 3) >> 1;
#pragma unroll 2
	for (int t = t_begin; t <= t_last; t++) {
		int outE[NPL][8], outO[NPL][8];
		const int re = 2 * t, ro = 2 * t - 1;  // rows arriving now
		const int r4 = 2 * t - 1, r3 = 2 * t - 2, r2 = 2 * t - 3, r1 = 2 * t - 4;
		const bool edge_y = (r1 - 1 <= 0) || (re >= h - 1);
#pragma unroll
		for (int p = 0; p < NPL; p++) {
			const int plane = NPL == 3 ? p : plane0;
			const char *arena = P.arena + img * P.arena_img_stride + plane * P.arena_plane_stride;
			int xe[8], xo[8];
			{
				int d[4], hh[4], v[4], l[4];
				const bool e_ok = re < h;
				const bool o_ok = ro >= 0 && ro < h;
				const BandRef &D = P.band[0], &H = P.band[1], &V = P.band[2];
				load4<SH>(arena + D.off + (long long)t * D.stride * ES, bc, col_ok && e_ok && bc < D.dimx, d);
				const long long hoff = t == 1 ? (long long)P.h_row1_off : (long long)t * H.stride;
				load4<SH>(arena + H.off + hoff * ES, bc, col_ok && e_ok && bc < H.dimx, hh);
				load4<SH>(arena + V.off + (long long)(t - 1) * V.stride * ES, bc, col_ok && o_ok && bc < V.dimx, v);
				if (LLSRC == LLSRC_BAND) {
					const BandRef &L = P.lband;
					load4<SH>(arena + L.off + (long long)(t - 1) * L.stride * ES, bc, col_ok && o_ok && bc < L.dimx, l);
#pragma unroll
					for (int k = 0; k < 4; k++) l[k] = TR<SH>(l[k] * P.dq[plane][3]);
				} else if (LLSRC == LLSRC_S16) {
					const short *lp = (const short *)P.ll + img * P.ll_img_stride + plane * P.ll_plane_stride +
					                  (long long)(t - 1) * P.ll_pitch;
					load4<true>((const char *)lp, bc, col_ok && o_ok && bc < (w >> 1), l);
				} else {
					const int *lp = (const int *)P.ll + img * P.ll_img_stride + plane * P.ll_plane_stride +
					                (long long)(t - 1) * P.ll_pitch;
					load4<false>((const char *)lp, bc, col_ok && o_ok && bc < (w >> 1), l);
#pragma unroll
					for (int k = 0; k < 4; k++) l[k] = TR<SH>(l[k]);  // (C) narrowing, wavelet2d.cpp:971-980
				}
				const int qd = P.dq[plane][0], qh = P.dq[plane][1], qv = P.dq[plane][2];
#pragma unroll
				for (int k = 0; k < 4; k++) {  // TSUQi: pBand[n] *= Quant (truncating store)
					xe[2 * k] = TR<SH>(d[k] * qd);
					xe[2 * k + 1] = TR<SH>(hh[k] * qh);
					xo[2 * k] = TR<SH>(v[k] * qv);
					xo[2 * k + 1] = l[k];
				}
			}
			if (edge_y) {
				if (r4 >= 0 && r4 < h) vU4<SH, TRANS, true>(xo, se0[p], xe, false, r4 == h - 1);
				if (r3 >= 0 && r3 < h) vU3<SH, TRANS, true>(se0[p], so4[p], xo, r3 == 0, r3 == h - 1);
				if (r2 >= 0 && r2 < h) vU2<SH, TRANS, true>(so4[p], se3[p], se0[p], false, r2 == h - 1);
				if (r1 >= 0 && r1 < h) vU1<SH, TRANS, true>(se3[p], so2[p], so4[p], r1 == 0, r1 == h - 1);
			} else {
				vU4<SH, TRANS, false>(xo, se0[p], xe, false, false);
				vU3<SH, TRANS, false>(se0[p], so4[p], xo, false, false);
				vU2<SH, TRANS, false>(so4[p], se3[p], se0[p], false, false);
				vU1<SH, TRANS, false>(se3[p], so2[p], so4[p], false, false);
			}
			// finished: even row r1 (se3), odd row r2 (so4)
#pragma unroll
			for (int k = 0; k < 8; k++) { outE[p][k] = se3[p][k]; outO[p][k] = so4[p][k]; }
			if (edge_x) { row_inv<SH, TRANS, true>(outE[p], cb, w); row_inv<SH, TRANS, true>(outO[p], cb, w); }
			else { row_inv<SH, TRANS, false>(outE[p], cb, w); row_inv<SH, TRANS, false>(outO[p], cb, w); }
			// rotate
#pragma unroll
			for (int k = 0; k < 8; k++) { so2[p][k] = so4[p][k]; se3[p][k] = se0[p][k]; se0[p][k] = xe[k]; so4[p][k] = xo[k]; }
		}
		// ---- write rows r1 (even) and r2 (odd)
#pragma unroll
		for (int half = 0; half < 2; half++) {
			const int row = half ? r2 : r1;
			if (!(row >= y0 && row < y1 && lane_out)) continue;
			if (DST == DST_S16) {
				const int *o = half ? outO[0] : outE[0];
				short *dp = (short *)P.dst + img * P.dst_img_stride + plane0 * P.dst_plane_stride + (long long)row * P.dst_pitch + cb;
				uint4 pk;
				pk.x = (unsigned)(o[0] & 0xFFFF) | ((unsigned)o[1] << 16);
				pk.y = (unsigned)(o[2] & 0xFFFF) | ((unsigned)o[3] << 16);
				pk.z = (unsigned)(o[4] & 0xFFFF) | ((unsigned)o[5] << 16);
				pk.w = (unsigned)(o[6] & 0xFFFF) | ((unsigned)o[7] << 16);
				*(uint4 *)dp = pk;
			} else if (DST == DST_S32) {
				const int *o = half ? outO[0] : outE[0];
				int *dp = (int *)P.dst + img * P.dst_img_stride + plane0 * P.dst_plane_stride + (long long)row * P.dst_pitch + cb;
				*(int4 *)dp = make_int4(o[0], o[1], o[2], o[3]);
				*(int4 *)(dp + 4) = make_int4(o[4], o[5], o[6], o[7]);
			} else if (DST == DST_U8_GRAY) {
				const int *o = half ? outO[0] : outE[0];
				unsigned b[8];
#pragma unroll
				for (int k = 0; k < 8; k++) {  // ric.cpp:229 / :237-240
					int v = o[k];
					v = P.shift ? clip255((int)(short)(128 + ((v + 8) >> 4))) : (v + 128);
					b[k] = (unsigned)v & 0xFF;
				}
				unsigned char *dp = (unsigned char *)P.dst + img * P.dst_img_stride + plane0 * P.dst_plane_stride +
				                    (long long)row * P.dst_pitch + cb;
				*(uint2 *)dp = make_uint2(b[0] | b[1] << 8 | b[2] << 16 | b[3] << 24, b[4] | b[5] << 8 | b[6] << 16 | b[7] << 24);
			} else {
				unsigned R[8], G[8], B[8];
#pragma unroll
				for (int k = 0; k < 8; k++) {  // YCoCgtoRGB<shift>, ric.cpp:93-112 (planes 0 Co, 1 Cg, 2 Y)
					int co = half ? outO[0][k] : outE[0][k];
					int cg = half ? outO[NPL > 1 ? 1 : 0][k] : outE[NPL > 1 ? 1 : 0][k];
					int y = half ? outO[NPL > 2 ? 2 : 0][k] : outE[NPL > 2 ? 2 : 0][k];
					if (P.shift) {
						co = (int)(short)((co + 4) >> 3);
						cg = (int)(short)((cg + 4) >> 3);
						y = (int)(short)((y + 8) >> 4);
					}
					y = (int)(short)(y - ((cg >> 1) - 128));
					cg = (int)(short)(cg + y);
					y = (int)(short)(y - (co >> 1));
					co = (int)(short)(co + y);
					if (P.shift) { co = clip255(co); cg = clip255(cg); y = clip255(y); }
					R[k] = (unsigned)co & 0xFF; G[k] = (unsigned)cg & 0xFF; B[k] = (unsigned)y & 0xFF;
				}
				unsigned char *dp = (unsigned char *)P.dst + img * P.dst_img_stride + (long long)row * P.dst_pitch + cb;
				*(uint2 *)dp = make_uint2(R[0] | R[1] << 8 | R[2] << 16 | R[3] << 24, R[4] | R[5] << 8 | R[6] << 16 | R[7] << 24);
				*(uint2 *)(dp + P.dst_plane_stride) =
				    make_uint2(G[0] | G[1] << 8 | G[2] << 16 | G[3] << 24, G[4] | G[5] << 8 | G[6] << 16 | G[7] << 24);
				*(uint2 *)(dp + 2 * P.dst_plane_stride) =
				    make_uint2(B[0] | B[1] << 8 | B[2] << 16 | B[3] << 24, B[4] | B[5] << 8 | B[6] << 16 | B[7] << 24);
			}
		}
	}
}

}  // namespace ric
