// ric_fwd.cuh -- one forward wavelet level, fused with colour/level-shift on the way in and the
// encode quantiser on the way out.
//
// Replaces, for one level of every plane of a batch of images:
//   RGBtoYCoCg / gray shift        src/ric/ric.cpp:76-91,143-148      (SRC_U8_*)
//   CWavelet2D::Transform97/53     src/lib/wavelet2d.cpp:407-492,636-692 (+TransLine :320-359,:593-611)
//   short->int widening            src/lib/wavelet2d.cpp:938-950
//   CBandCodec::buildTree          src/lib/bandcodec.cpp:239-322 (this level's D/H/V bands)
//   CBand::TSUQ (LL, last level)   src/lib/band.h:65-92
//
// Work decomposition: one WARP per (image, plane, row segment, 240-column strip).  A lane holds 8
// consecutive columns; the horizontal lifting takes its neighbours from warp shuffles, the
// vertical lifting is a streaming filter whose 4-row state lives in registers while the warp
// walks down its segment two rows at a time.  Finished coefficients are collected per lane into
// 4x4 blocks (exactly the reference's block grid: strip and segment origins are multiples of 8
// level samples), quantised in registers and written once, in the band layout the entropy coder
// reads.  No shared-memory staging of samples, no block-level synchronisation in the main loop.
#pragma once
#include "ric_dev.cuh"

namespace ric {

enum { SRC_U8_GRAY = 0, SRC_U8_RGB = 1, SRC_S16 = 2, SRC_S32 = 3 };
enum { LL_S16 = 0, LL_S32 = 1, LL_BAND = 2 };

struct BandRef {
	long long off;  // byte offset inside one plane arena
	int dimx, dimy, stride;
	int fl_off;     // offset of this band's block flags inside one plane's flag area (bytes)
	int fl_bw;      // blocks per row = ceil(dimx/4)
};

struct FwdParams {
	const void *src;           // u8 image / s16 plane / LL scratch
	long long src_img_stride;  // elements between images
	long long src_plane_stride;  // elements between planes (u8: between channels)
	int src_pitch;             // elements between rows
	void *ll;                  // LL scratch of this level's output (next level's input)
	long long ll_img_stride, ll_plane_stride;
	int ll_pitch;
	char *arena;               // band arenas
	long long arena_img_stride, arena_plane_stride;  // bytes
	unsigned char *flags;      // block non-zero flags (pRD != 0), all levels
	long long flags_img_stride, flags_plane_stride;
	BandRef band[3];           // D, H, V of this level
	BandRef child[3];          // same orientation one level finer (has_child)
	BandRef lband;             // LL band (LL_BAND only)
	int has_child;
	int w, h;                  // level input size
	int nstrips, nsegs, seg_rows, nplanes, nimages;
	int shift;                 // colour path: q != 0 (fixed-point up-shift, ric.cpp:85-89,147)
	int quant;                 // 0: store raw coefficients (CWavelet2D::Transform only)
	int plane_class[3];        // quantiser class of each plane (0 luma / 1 chroma)
	QuantBand qb[2][3];        // [class][orientation]
	int llQ[2], lliQ[2], llT[2];  // LL TSUQ scalars per class
};

template <int SRC>
struct RawRow {
	static constexpr int N = SRC == SRC_U8_GRAY ? 2 : SRC == SRC_U8_RGB ? 6 : SRC == SRC_S16 ? 4 : 8;
	unsigned r[N];
};

// issue the global loads of one row (8 columns starting at cb) -- conversion happens later
template <int SRC>
__device__ __forceinline__ void load_raw(RawRow<SRC> &raw, const void *base, long long row_off, int cb, bool ok,
                                         long long plane_stride)
{
#pragma unroll
	for (int i = 0; i < RawRow<SRC>::N; i++) raw.r[i] = 0;
	if (!ok) return;
	if (SRC == SRC_U8_GRAY) {
		uint2 a = __ldg((const uint2 *)((const unsigned char *)base + row_off + cb));
		raw.r[0] = a.x; raw.r[1] = a.y;
	} else if (SRC == SRC_U8_RGB) {
		const unsigned char *p = (const unsigned char *)base + row_off + cb;
#pragma unroll
		for (int ch = 0; ch < 3; ch++) {
			uint2 a = __ldg((const uint2 *)(p + ch * plane_stride));
			raw.r[2 * ch] = a.x; raw.r[2 * ch + 1] = a.y;
		}
	} else if (SRC == SRC_S16) {
		uint4 a = __ldg((const uint4 *)((const short *)base + row_off + cb));
		raw.r[0] = a.x; raw.r[1] = a.y; raw.r[2] = a.z; raw.r[3] = a.w;
	} else {
		const int *p = (const int *)base + row_off + cb;
		uint4 a = __ldg((const uint4 *)p), b = __ldg((const uint4 *)(p + 4));
		raw.r[0] = a.x; raw.r[1] = a.y; raw.r[2] = a.z; raw.r[3] = a.w;
		raw.r[4] = b.x; raw.r[5] = b.y; raw.r[6] = b.z; raw.r[7] = b.w;
	}
}

__device__ __forceinline__ int byte_of(unsigned lo, unsigned hi, int k)
{
	return (int)(((k < 4 ? lo : hi) >> (8 * (k & 3))) & 0xFF);
}

// raw registers -> 8 level-input samples (colour transform / level shift fused here)
template <int SRC>
__device__ __forceinline__ void convert_raw(const RawRow<SRC> &raw, int (&v)[8], int plane, int shift)
{
	if (SRC == SRC_U8_GRAY) {
#pragma unroll
		for (int k = 0; k < 8; k++) {
			int p = byte_of(raw.r[0], raw.r[1], k) - 128;   // ric.cpp:144 / :147
			v[k] = shift ? p << 4 : p;
		}
	} else if (SRC == SRC_U8_RGB) {
#pragma unroll
		for (int k = 0; k < 8; k++) {  // RGBtoYCoCg<shift>, ric.cpp:76-91 (planes 0 Co, 1 Cg, 2 Y)
			int R = byte_of(raw.r[0], raw.r[1], k), G = byte_of(raw.r[2], raw.r[3], k), B = byte_of(raw.r[4], raw.r[5], k);
			int co = R - B;
			int t = B + (co >> 1);
			int cg = G - t;
			int y = t + ((cg >> 1) - 128);
			int o = plane == 0 ? co : plane == 1 ? cg : y;
			int sh = plane == 2 ? 4 : 3;
			v[k] = shift ? o << sh : o;
		}
	} else if (SRC == SRC_S16) {
#pragma unroll
		for (int k = 0; k < 8; k++) {
			unsigned wd = raw.r[k >> 1];
			v[k] = (k & 1) ? (int)wd >> 16 : (int)(short)(wd & 0xFFFF);
		}
	} else {
#pragma unroll
		for (int k = 0; k < 8; k++) v[k] = (int)raw.r[k];
	}
}

template <bool SH>
__device__ __forceinline__ void store4(char *rowp, int col, int c0, int c1, int c2, int c3)
{
	if (SH) {
		uint2 o;
		o.x = (unsigned)(c0 & 0xFFFF) | ((unsigned)c1 << 16);
		o.y = (unsigned)(c2 & 0xFFFF) | ((unsigned)c3 << 16);
		*(uint2 *)(rowp + 2 * (long long)col) = o;
	} else {
		*(int4 *)(rowp + 4 * (long long)col) = make_int4(c0, c1, c2, c3);
	}
}

// Quantise (optionally) and write one lane's 4x4 block of band `b`; bx/by: block coordinates.
template <bool SH>
__device__ __forceinline__ void flush_block(const FwdParams &P, const BandRef &b, const BandRef &ch, char *arena,
                                            unsigned char *flags, const QuantBand *qb, int (&c)[16], int bx, int by,
                                            bool lane_ok)
{
	const int x0 = bx * 4, y0 = by * 4;
	if (!lane_ok || x0 >= b.dimx || y0 >= b.dimy) return;  // (warp-divergent exit is fine: no syncs below but __any)
	const int bw = min(4, b.dimx - x0), bh = min(4, b.dimy - y0);
	if (P.quant) {
		int cnt = quant_block<SH>(c, qb, bw, bh);
		int nz = cnt;
		if (P.has_child && bw == 4 && bh == 4) {  // buildTree :267-270: add the four child blocks
			const unsigned char *cf = flags + ch.fl_off + (2 * by) * ch.fl_bw + 2 * bx;
			nz += cf[0] + cf[1] + cf[ch.fl_bw] + cf[ch.fl_bw + 1];
		}
		flags[b.fl_off + by * b.fl_bw + bx] = nz != 0;
		if (nz == 0) c[0] = -0x8000;  // INSIGNIF_BLOCK, bandcodec.cpp:113,272
	}
	char *base = arena + b.off;
	const int es = SH ? 2 : 4;
	if (bw == 4) {
#pragma unroll
		for (int r = 0; r < 4; r++)
			if (r < bh) store4<SH>(base + (long long)(y0 + r) * b.stride * es, x0, c[4 * r], c[4 * r + 1], c[4 * r + 2], c[4 * r + 3]);
	} else {
#pragma unroll
		for (int r = 0; r < 4; r++)
#pragma unroll
			for (int k = 0; k < 4; k++)
				if (r < bh && k < bw) {
					char *p = base + ((long long)(y0 + r) * b.stride + x0 + k) * es;
					if (SH) *(short *)p = (short)c[4 * r + k]; else *(int *)p = c[4 * r + k];
				}
	}
}

template <bool SH, int TRANS, int SRC, int LLDST>
__global__ void __launch_bounds__(128) fwd_level_kernel(const __grid_constant__ FwdParams P)
{
	__shared__ QuantBand s_qb[2][3];
	for (int i = threadIdx.x; i < (int)(sizeof(s_qb) / 4); i += blockDim.x) ((int *)s_qb)[i] = ((const int *)P.qb)[i];
	__syncthreads();

	const int lane = threadIdx.x & 31;
	long long job = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const long long njobs = (long long)P.nstrips * P.nplanes * P.nsegs * P.nimages;
	if (job >= njobs) return;
	// plane fastest: the planes of one RGB strip share their u8 loads through L1
	const int plane = (int)(job % P.nplanes); job /= P.nplanes;
	const int sx = (int)(job % P.nstrips); job /= P.nstrips;
	const int sy = (int)(job % P.nsegs);
	const int img = (int)(job / P.nsegs);

	const int w = P.w, h = P.h;
	const int x0 = sx * STRIP_W;
	const int cb = x0 - LANE_W + lane * LANE_W;       // first column of this lane
	const bool col_ok = cb >= 0 && cb < w;            // lane has at least one real column
	const bool lane_out = lane >= 1 && lane <= 30;    // lane owns output columns
	const bool edge_x = (x0 == 0) || (w <= x0 + STRIP_W + LANE_W);
	const int y0 = sy * P.seg_rows;
	const int y1 = min(h, y0 + P.seg_rows);
	const int y1r = (y1 + 7) & ~7;

	const void *src = SRC == SRC_U8_RGB
	                      ? (const void *)((const unsigned char *)P.src + img * P.src_img_stride)
	                      : SRC == SRC_U8_GRAY
	                            ? (const void *)((const unsigned char *)P.src + img * P.src_img_stride + plane * P.src_plane_stride)
	                            : SRC == SRC_S16
	                                  ? (const void *)((const short *)P.src + img * P.src_img_stride + plane * P.src_plane_stride)
	                                  : (const void *)((const int *)P.src + img * P.src_img_stride + plane * P.src_plane_stride);
	char *arena = P.arena + img * P.arena_img_stride + plane * P.arena_plane_stride;
	unsigned char *flags = P.flags + img * P.flags_img_stride + plane * P.flags_plane_stride;
	const int cls = P.plane_class[plane];
	const QuantBand *qbD = &s_qb[cls][0], *qbH = &s_qb[cls][1], *qbV = &s_qb[cls][2];
	const int bx = cb >> 3;  // block column of this lane in every band of this level

	// vertical state: so1 raw odd row 2t-1, se1 S1'd even row 2t-2, so2 S2'd odd row 2t-3, se3 S3'd even row 2t-4
	int so1[8], se1[8], so2[8], se3[8];
#pragma unroll
	for (int k = 0; k < 8; k++) so1[k] = se1[k] = so2[k] = se3[k] = 0;
	int bD[16], bH[16], bV[16];
#pragma unroll
	for (int k = 0; k < 16; k++) bD[k] = bH[k] = bV[k] = 0;

	const int t_begin = (y0 >> 1) - 2, t_end = (y1r >> 1) + 2;  // exclusive; (t_end - t_begin) % 4 == 0
	RawRow<SRC> rawE, rawO;
	{
		int re = 2 * t_begin, ro = re + 1;
		load_raw<SRC>(rawE, src, (long long)re * P.src_pitch, cb, col_ok && re >= 0 && re < h, P.src_plane_stride);
		load_raw<SRC>(rawO, src, (long long)ro * P.src_pitch, cb, col_ok && ro >= 0 && ro < h, P.src_plane_stride);
	}

	for (int t0 = t_begin; t0 < t_end; t0 += 4) {
		// rows touched by this group: 2*t0-4 .. 2*t0+7; edge formulas needed if that range meets row 0 or h-1
		const bool edge_y = (2 * t0 - 4 <= 0) || (2 * t0 + 7 >= h - 1);
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const int t = t0 + u;
			int ne[8], no[8];
			convert_raw<SRC>(rawE, ne, plane, P.shift);
			convert_raw<SRC>(rawO, no, plane, P.shift);
			{  // prefetch the next row pair
				int re = 2 * t + 2, ro = re + 1;
				load_raw<SRC>(rawE, src, (long long)re * P.src_pitch, cb, col_ok && re >= 0 && re < h, P.src_plane_stride);
				load_raw<SRC>(rawO, src, (long long)ro * P.src_pitch, cb, col_ok && ro >= 0 && ro < h, P.src_plane_stride);
			}
			if (edge_x) { row_fwd<SH, TRANS, true>(ne, cb, w); row_fwd<SH, TRANS, true>(no, cb, w); }
			else { row_fwd<SH, TRANS, false>(ne, cb, w); row_fwd<SH, TRANS, false>(no, cb, w); }

			const int r1 = 2 * t, r2 = 2 * t - 1, r3 = 2 * t - 2, r4 = 2 * t - 3;
			if (edge_y) {
				if (r1 >= 0 && r1 < h) vS1<SH, TRANS, true>(ne, so1, no, r1 == 0, r1 == h - 1);
				if (r2 >= 0 && r2 < h) vS2<SH, TRANS, true>(so1, se1, ne, false, r2 == h - 1);
				if (r3 >= 0 && r3 < h) vS3<SH, TRANS, true>(se1, so2, so1, r3 == 0, r3 == h - 1);
				if (r4 >= 0 && r4 < h) vS4<SH, TRANS, true>(so2, se3, se1, false, r4 == h - 1);
			} else {
				vS1<SH, TRANS, false>(ne, so1, no, false, false);
				vS2<SH, TRANS, false>(so1, se1, ne, false, false);
				vS3<SH, TRANS, false>(se1, so2, so1, false, false);
				vS4<SH, TRANS, false>(so2, se3, se1, false, false);
			}
			// finished: even row r3 (band row jd = t-1: D even cols, H odd cols), odd row r4 (band row jv = t-2: V, LL)
			const int jd = t - 1, jv = t - 2;
			constexpr int UD = 0, UV = 0;
			(void)UD; (void)UV;
			const int sd = (u + 1) & 3, sv = u & 3;  // slot of jd / jv inside its 4-row block (y0/2 is a multiple of 4)
#pragma unroll
			for (int k = 0; k < 4; k++) {
				bD[4 * sd + k] = se1[2 * k];
				bH[4 * sd + k] = se1[2 * k + 1];
				bV[4 * sv + k] = so2[2 * k];
			}
			// LL row jv
			if (jv >= (y0 >> 1) && jv < (y1 >> 1) && lane_out) {
				int llv[4];
#pragma unroll
				for (int k = 0; k < 4; k++) llv[k] = so2[2 * k + 1];
				const int lx = cb >> 1;
				if (LLDST == LL_BAND) {
					const BandRef &L = P.lband;
					if (P.quant) {
#pragma unroll
						for (int k = 0; k < 4; k++) llv[k] = tsuq1<SH>(llv[k], P.llT[cls], P.lliQ[cls]);
					}
					char *rowp = arena + L.off + (long long)jv * L.stride * (SH ? 2 : 4);
					if (lx + 4 <= L.dimx) store4<SH>(rowp, lx, llv[0], llv[1], llv[2], llv[3]);
					else {
#pragma unroll
						for (int k = 0; k < 4; k++)
							if (lx + k < L.dimx) {
								if (SH) ((short *)rowp)[lx + k] = (short)llv[k]; else ((int *)rowp)[lx + k] = llv[k];
							}
					}
				} else if (lx < (w >> 1)) {
					// scratch rows are padded to a multiple of 8 samples: a full 4-sample store is always in bounds
					if (LLDST == LL_S16) {
						short *rowp = (short *)P.ll + img * P.ll_img_stride + plane * P.ll_plane_stride + (long long)jv * P.ll_pitch;
						store4<true>((char *)rowp, lx, llv[0], llv[1], llv[2], llv[3]);
					} else {
						int *rowp = (int *)P.ll + img * P.ll_img_stride + plane * P.ll_plane_stride + (long long)jv * P.ll_pitch;
						store4<false>((char *)rowp, lx, llv[0], llv[1], llv[2], llv[3]);
					}
				}
			}
			if (u == 2) {  // D/H block row (jd>>2) complete
				const int by = jd >> 2;
				if (by >= (y0 >> 3) && by < (y1r >> 3)) {
					flush_block<SH>(P, P.band[0], P.child[0], arena, flags, qbD, bD, bx, by, lane_out);
					flush_block<SH>(P, P.band[1], P.child[1], arena, flags, qbH, bH, bx, by, lane_out);
				}
			}
			if (u == 3) {  // V block row (jv>>2) complete
				const int by = jv >> 2;
				if (by >= (y0 >> 3) && by < (y1r >> 3))
					flush_block<SH>(P, P.band[2], P.child[2], arena, flags, qbV, bV, bx, by, lane_out);
			}
			// rotate the vertical state
#pragma unroll
			for (int k = 0; k < 8; k++) { se3[k] = se1[k]; so2[k] = so1[k]; se1[k] = ne[k]; so1[k] = no[k]; }
		}
	}
}

}  // namespace ric
