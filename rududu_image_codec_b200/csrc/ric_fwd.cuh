// ric_fwd.cuh -- one forward wavelet level, fused with colour/level-shift on the way in and the
// encode quantiser on the way out.
//
// Replaces, for one level of every plane of a batch of images:
//   RGBtoYCoCg / gray shift        src/ric/ric.cpp:76-91,143-148      (SRC_U8_*)
//   CWavelet2D::Transform97/53/Haar src/lib/wavelet2d.cpp:407-492,636-692,788-819 (+TransLine :320-359,:593-611,:766-775)
//   short->int widening            src/lib/wavelet2d.cpp:938-950
//   CBandCodec::buildTree          src/lib/bandcodec.cpp:239-322 (this level's D/H/V bands)
//   CBand::TSUQ (LL, last level)   src/lib/band.h:65-92
//
// Work decomposition: one warp JOB per (image, plane, row segment, 240-column strip); persistent
// warps claim jobs from a global counter.  A lane holds 8 consecutive columns; the horizontal
// lifting takes its neighbours from warp shuffles, the vertical lifting is a streaming filter whose
// 4-row state lives in registers while the warp walks down its segment two rows at a time.
// Finished band rows go to a lane-private shared-memory ring; every fourth iteration each lane owns
// complete 4x4 blocks (exactly the reference's block grid: strip and segment origins are multiples
// of 8 level samples), quantises them and writes them once, in the band layout the entropy coder
// reads.  No block-level synchronisation in the main loop.
#pragma once
#include <type_traits>

#include "ric_dev.cuh"

namespace ric {

enum { SRC_U8_GRAY = 0, SRC_U8_RGB = 1, SRC_S16 = 2, SRC_S32 = 3 };

__host__ __device__ constexpr int fwd_warps(bool sh) { return sh ? 4 : 2; }  // warps per CTA (independent jobs)
constexpr int RING_ROWS = 8;   // band rows staged per warp before a 4x4 block row is flushed

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
	void *ll;                  // LL scratch of this level's output (next level's input); s16 (short level) / s32
	long long ll_img_stride, ll_plane_stride;
	int ll_pitch;
	int ll_to_band;            // last level: LL goes to the L band (with TSUQ when quant)
	char *arena;               // band arenas
	long long arena_img_stride, arena_plane_stride;  // bytes
	unsigned char *flags;      // block non-zero flags (pRD != 0), all levels
	long long flags_img_stride, flags_plane_stride;
	BandRef band[3];           // D, H, V of this level
	BandRef child[3];          // same orientation one level finer (has_child)
	BandRef lband;             // LL band (ll_to_band)
	int has_child;
	int w, h;                  // level input size
	int nstrips, nsegs, seg_rows, nplanes, nimages;
	int shift;                 // colour path: q != 0 (fixed-point up-shift, ric.cpp:85-89,147)
	int quant;                 // 0: store raw coefficients (CWavelet2D::Transform only)
	int plane_class[3];        // quantiser class of each plane (0 luma / 1 chroma)
	QuantBand qb[2][3];        // [class][orientation]
	int llQ[2], lliQ[2], llT[2];  // LL TSUQ scalars per class
	unsigned long long *counter;  // dynamic job fetch (zeroed before the launch)
	unsigned long long *stats;    // optional (profiling): [0] packed-kernel plane iterations, [1] of those on the scalar path
};

template <int SRC>
struct RawRow {
	static constexpr int N = SRC == SRC_U8_GRAY ? 2 : SRC == SRC_U8_RGB ? 6 : SRC == SRC_S16 ? 4 : 8;
	unsigned r[N];
};

// issue the global loads of one row (8 columns starting at cb) -- conversion happens later.
// RGB: plane 0 (Co) does not need G.
template <int SRC>
__device__ __forceinline__ void load_raw(RawRow<SRC> &raw, const char *p, bool ok, long long plane_stride, int plane)
{
	// p: address of this lane's first sample of the row (the row loop advances it by two rows per iteration: no
	// per-load 64-bit multiply).  Predicated loads, no branch.
	if constexpr (SRC == SRC_U8_GRAY) {
		const uint2 a = ldg_u2_if(p, ok);
		raw.r[0] = a.x; raw.r[1] = a.y;
	} else if constexpr (SRC == SRC_U8_RGB) {
		const uint2 a = ldg_u2_if(p, ok), g = ldg_u2_if(p + plane_stride, ok && plane != 0), b = ldg_u2_if(p + 2 * plane_stride, ok);
		raw.r[0] = a.x; raw.r[1] = a.y; raw.r[2] = g.x; raw.r[3] = g.y; raw.r[4] = b.x; raw.r[5] = b.y;
	} else if constexpr (SRC == SRC_S16) {
		const uint4 a = ldg_u4_if(p, ok);
		raw.r[0] = a.x; raw.r[1] = a.y; raw.r[2] = a.z; raw.r[3] = a.w;
	} else {
		const uint4 a = ldg_u4_if(p, ok), b = ldg_u4_if(p + 16, ok);
		raw.r[0] = a.x; raw.r[1] = a.y; raw.r[2] = a.z; raw.r[3] = a.w;
		raw.r[4] = b.x; raw.r[5] = b.y; raw.r[6] = b.z; raw.r[7] = b.w;
	}
}

// c + mult * (byte k & 3 of word): one IDP.4A (u8 x s8 dot product, FMA pipe; scripts/ubench/pipes3.cu).
// mult in [-128, 127], compile-time or warp-uniform; the byte position is a compile-time shift.
__device__ __forceinline__ int dp4a_pick(unsigned word, int mult, int k, int c)
{
	int d;
	asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(word), "r"((int)((unsigned)(mult & 0xFF) << (8 * (k & 3)))), "r"(c));
	return d;
}


// Per-plane constants of the RGB closed forms below, one row per plane in shared memory (filled once per CTA):
// the byte-position-shifted multipliers of R, B and G, then {c, m, down-shift for q == 0, unused}.  Read back with
// four 128-bit loads per row pair: kept in registers across the loop they would cost 15 registers, recomputed
// from `plane` they cost 2 instructions per sample (both measured).
struct ColourTab { int4 aR, aB, aG, cm; };

__device__ __forceinline__ void colour_tab_fill(ColourTab *tab, int shift)  // tab[3], any thread layout
{
	for (int i = threadIdx.x; i < 3 * 16; i += blockDim.x) {
		const int plane = i >> 4, j = i & 15;
		const int aR = plane == 0 ? 8 : plane == 1 ? -4 : 4, aB = plane == 0 ? -8 : plane == 1 ? -4 : 4, aG = plane == 0 ? 0 : 8;
		int v;
		if (j < 12) v = (int)((unsigned)((j < 4 ? aR : j < 8 ? aB : aG) & 0xFF) << (8 * (j & 3)));
		else v = j == 12 ? (plane == 0 ? 0 : plane == 1 ? 4 : -2048) : j == 13 ? (plane == 0 ? ~0 : plane == 1 ? ~7 : ~15)
		                 : j == 14 ? (shift ? 0 : plane == 2 ? 4 : 3) : 0;
		((int *)tab)[i] = v;
	}
}

__device__ __forceinline__ int4 lds_v4(const int4 *p)  // volatile: reloaded where it is used, never hoisted out of the row loop
{
	int4 r;
	asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"((unsigned)__cvta_generic_to_shared(p)));
	return r;
}

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c)  // IDP.4A.U8.S8: c + sum of a.u8[i] * b.s8[i]
{
	int d;
	asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}

// RGBtoYCoCg<shift> (ric.cpp:76-91; planes 0 Co, 1 Cg, 2 Y) of two rows of 8 pixels.  The reference's lifting chain
// Co = R - B, t = B + (Co >> 1), Cg = G - t, Y = t + (Cg >> 1) - 128, then << 3 (chroma) / << 4 (luma), has closed
// forms, because floors of integers nest (t = (R + B) >> 1, Y + 128 = (R + 2G + B) >> 2):
//   Co << 3 =  8R - 8B
//   Cg << 3 = (8G - 4R - 4B + 4) & ~7        (8G - 4(R + B) is 8 Cg or 8 Cg - 4)
//   Y  << 4 = (4R + 8G + 4B - 2048) & ~15
// (checked over all 2^24 pixels in tests/test_host_logic.py).  Each is a dot product of pixel bytes with small
// constants plus one mask: IDP.4A picks the byte, scales it and accumulates in ONE FMA-pipe instruction -- no byte
// extraction and no shifts on the ALU pipe.  One sequence, (aR*R + aB*B + aG*G + c) & m, serves the three planes
// with the warp's constants from `tab`: no per-plane branches, small code (the kernel is instruction-fetch
// sensitive: profiles/README.md, round 2).  q == 0 (no up-shift): the values are exact multiples and are shifted
// back down.
__device__ __forceinline__ void convert_rgb2(const RawRow<SRC_U8_RGB> &rawE, const RawRow<SRC_U8_RGB> &rawO, int (&ve)[8], int (&vo)[8],
                                             const ColourTab *tab)
{
	const int4 aR = lds_v4(&tab->aR), aB = lds_v4(&tab->aB), aG = lds_v4(&tab->aG), cm = lds_v4(&tab->cm);
#pragma unroll
	for (int k = 0; k < 8; k++) {
		const int j = k & 3;
		const int sR = j == 0 ? aR.x : j == 1 ? aR.y : j == 2 ? aR.z : aR.w, sB = j == 0 ? aB.x : j == 1 ? aB.y : j == 2 ? aB.z : aB.w;
		const int sG = j == 0 ? aG.x : j == 1 ? aG.y : j == 2 ? aG.z : aG.w;
		ve[k] = dp4a_us(rawE.r[2 + (k >> 2)], sG, dp4a_us(rawE.r[k >> 2], sR, dp4a_us(rawE.r[4 + (k >> 2)], sB, cm.x))) & cm.y;
		vo[k] = dp4a_us(rawO.r[2 + (k >> 2)], sG, dp4a_us(rawO.r[k >> 2], sR, dp4a_us(rawO.r[4 + (k >> 2)], sB, cm.x))) & cm.y;
	}
	if (cm.z) {  // warp-uniform, rare
#pragma unroll
		for (int k = 0; k < 8; k++) { ve[k] >>= cm.z; vo[k] >>= cm.z; }
	}
}

// raw registers -> 8 level-input samples (colour transform / level shift fused here).
// (RGB sources go through convert_rgb2 above.)
template <int SRC>
__device__ __forceinline__ void convert_raw(const RawRow<SRC> &raw, int (&v)[8], int plane, int shift)
{
	if (SRC == SRC_U8_GRAY) {
#pragma unroll
		for (int k = 0; k < 8; k++) v[k] = dp4a_pick(raw.r[k >> 2], 16, k, -2048);  // (pixel - 128) << 4, ric.cpp:147
		if (!shift) {  // ric.cpp:144 (q == 0): no up-shift
#pragma unroll
			for (int k = 0; k < 8; k++) v[k] >>= 4;
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

// per-warp staging ring of finished band rows: [band][row & 7][lane] -> 4 samples (lane-private)
template <bool SH>
struct Ring {
	typedef typename std::conditional<SH, uint2, int4>::type vec;
	vec v[3][RING_ROWS][32];
};

template <bool SH>
__device__ __forceinline__ void ring_put(Ring<SH> &rg, int o, int row, int lane, int c0, int c1, int c2, int c3)
{
	if (SH) {
		uint2 t;
		t.x = (unsigned)(c0 & 0xFFFF) | ((unsigned)c1 << 16);
		t.y = (unsigned)(c2 & 0xFFFF) | ((unsigned)c3 << 16);
		*(uint2 *)&rg.v[o][row & (RING_ROWS - 1)][lane] = t;
	} else {
		*(int4 *)&rg.v[o][row & (RING_ROWS - 1)][lane] = make_int4(c0, c1, c2, c3);
	}
}

template <bool SH>
__device__ __forceinline__ void ring_get(const Ring<SH> &rg, int o, int row, int lane, int &c0, int &c1, int &c2, int &c3)
{
	if (SH) {
		uint2 t = *(const uint2 *)&rg.v[o][row & (RING_ROWS - 1)][lane];
		c0 = (int)(short)(t.x & 0xFFFF); c1 = (int)t.x >> 16;
		c2 = (int)(short)(t.y & 0xFFFF); c3 = (int)t.y >> 16;
	} else {
		int4 t = *(const int4 *)&rg.v[o][row & (RING_ROWS - 1)][lane];
		c0 = t.x; c1 = t.y; c2 = t.z; c3 = t.w;
	}
}

// Quantise (optionally) and write block row `by` of the three bands from the staging ring.
// One lane = one 4x4 block per band (bx = this lane's block column).  Called by all 32 lanes of the
// warp (the quantiser uses full-warp votes); lanes without a block in the band (halo lanes, columns
// or rows beyond the band) run it on an empty block and store nothing.
template <bool SH>
__device__ __forceinline__ void flush_blocks(const FwdParams &P, const Ring<SH> &rg, char *arena, unsigned char *flags,
                                             const QuantBand *qb3, int bx, int by, int lane, bool lane_out)
{
#pragma unroll 1
	for (int o = 0; o < 3; o++) {
		const BandRef &b = P.band[o];
		const int x0 = bx * 4, y0 = by * 4;
		if (y0 >= b.dimy) continue;  // warp-uniform
		const bool have = lane_out && x0 < b.dimx;
		const int bw = have ? min(4, b.dimx - x0) : 0, bh = min(4, b.dimy - y0);
		int c[16];
#pragma unroll
		for (int r = 0; r < 4; r++) ring_get<SH>(rg, o, y0 + r, lane, c[4 * r], c[4 * r + 1], c[4 * r + 2], c[4 * r + 3]);
		if (P.quant) {
			int nz = quant_block<SH>(c, qb3 + o, bw, bh);
			if (P.has_child && bw == 4 && bh == 4) {  // buildTree :267-270: add the four child blocks
				const BandRef &ch = P.child[o];
				const unsigned char *cf = flags + ch.fl_off + (2 * by) * ch.fl_bw + 2 * bx;
				nz += cf[0] + cf[1] + cf[ch.fl_bw] + cf[ch.fl_bw + 1];  // (odd fl_bw: the second row is not 2-byte aligned)
			}
			if (have) flags[b.fl_off + by * b.fl_bw + bx] = nz != 0;
			if (nz == 0) c[0] = -0x8000;  // INSIGNIF_BLOCK, bandcodec.cpp:113,272
		}
		if (!have) continue;
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
}

// ---- short levels: the same flush on PACKED rows ---------------------------------------------------
// The block never exists as 16 registers: the element-wise passes run as rolled loops over the four
// ring rows (4 packed int16 each), write their results back into the ring in place and park the
// candidate keys in a second lane-private shared-memory array; only the 16 keys are pulled into
// registers for the sort.  Same arithmetic as quant_block (ric_dev.cuh), a quarter of the code:
// the loop body of the kernel has to stay inside the instruction cache.
typedef uint4 KeyRows[4][32];

__device__ __forceinline__ int s16lo(unsigned w) { return (int)(short)(w & 0xFFFF); }
__device__ __forceinline__ int s16hi(unsigned w) { return (int)w >> 16; }

// Element-wise pass of the block quantiser on packed rows, scalar form: any quantiser (Q > 16383 thresholds
// wrap in int16), partial blocks, the -32768 corner of s2u_ (utils.h:95-99).  Rewrites the ring rows in place
// with the folded quantised values, parks the rank candidates' keys (value << 4 | 15 - raster) and returns
// the number of sure non-zeros; nc = number of candidates (tsuqBlock, bandcodec.cpp:166-186).
template <int RR, class RingT>
__device__ __forceinline__ int quant_rows_scalar(RingT &rg, KeyRows &keys, const QuantBand *qb, int o, int y0, int bw, int bh,
                                                 int lane, int &nc)
{
	const bool full = bw == 4 && bh == 4;
	const int T = full ? qb->T : qb->Te;
	const unsigned T2 = (unsigned)(2 * T);
	const int iQ = qb->iQ;
	const unsigned uthr0 = full ? (unsigned)(qb->thr[0] & 0xFFFF) : 0u;  // partial blocks have no candidates
	int cnt = 0;
	nc = 0;
#pragma unroll 1
	for (int r = 0; r < 4; r++) {
		uint2 &row = rg.v[o][(y0 + r) & (RR - 1)][lane];
		const uint2 wv = row;
		const int v[4] = {s16lo(wv.x), s16hi(wv.x), s16lo(wv.y), s16hi(wv.y)};
		int out[4], key[4];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const bool live = j < bw && r < bh && (unsigned)(v[j] + T) > T2;
			const int sgn = (int)((unsigned)v[j] >> 31);
			const unsigned uf = (unsigned)(2 * abs(v[j]) + sgn) & 0xFFFFu;  // s2u_ (utils.h:95-99), C-typed
			const bool cand = live && uf < uthr0;
			const int qq = ((int)(uf >> 1) * iQ + (1 << 15)) >> 16;          // int arithmetic as in the reference (:172)
			out[j] = !live ? 0 : cand ? (2 | sgn) : ((qq << 1) | sgn);
			key[j] = cand ? (int)(uf << 4) | (15 - 4 * r - j) : 0;
			cnt += (live && !cand) ? 1 : 0;
			nc += cand ? 1 : 0;
		}
		row = make_uint2((unsigned)(out[0] & 0xFFFF) | ((unsigned)out[1] << 16), (unsigned)(out[2] & 0xFFFF) | ((unsigned)out[3] << 16));
		keys[r][lane] = make_uint4((unsigned)key[0], (unsigned)key[1], (unsigned)key[2], (unsigned)key[3]);
	}
	return cnt;
}

// Rank stage on the parked 32-bit keys (tsuqBlock :188-199): which candidates survive.  Returns how many.
template <int RR, class RingT>
__device__ __forceinline__ int rank_rows_scalar(RingT &rg, KeyRows &keys, const QuantBand *qb, int o, int y0, int lane, int cnt, int ncm)
{
	int s[16];
#pragma unroll
	for (int r = 0; r < 4; r++) {
		const uint4 kk = keys[r][lane];
		s[4 * r] = (int)kk.x; s[4 * r + 1] = (int)kk.y; s[4 * r + 2] = (int)kk.z; s[4 * r + 3] = (int)kk.w;
	}
	int kstar = 0x7fffffff, m = 0;
	if (ncm == 1) {  // at most one candidate per block: rank 0, survives iff f >= thr[cnt]
		int s0 = 0;
#pragma unroll
		for (int k = 0; k < 16; k++) s0 = max(s0, s[k]);
		bool pass;
		if (qb->fast) pass = s0 >= qb->kthr[cnt];
		else pass = s0 != 0 && !((int)(short)(s0 >> 4) < qb->thr[cnt & 15]);
		if (pass && s0 != 0) { kstar = s0; m = 1; }
	} else {
		sort16_desc(s);
		if (__builtin_expect(qb->fast, 1)) {
			const int *kt = qb->kthr + cnt;
#pragma unroll
			for (int i = 0; i < 16; i++) {
				if ((i & 3) == 0 && i >= ncm) break;  // warp-uniform
				const bool pass = s[i] >= kt[i];
				kstar = pass ? s[i] : kstar;
				m = pass ? i + 1 : m;
			}
		} else {
			// irregular thresholds (Q > 16383: int16 wrap in thr[]): rank every candidate against the
			// parked keys with the reference's signed compare (:191); rolled, rare, small
#pragma unroll 1
			for (int i = 0; i < 16; i++) {
				const int ki = ((const int *)&keys[i >> 2][lane])[i & 3];
				if (ki == 0) continue;
				int rank = 0;
#pragma unroll 1
				for (int j = 0; j < 16; j++) rank += ((const int *)&keys[j >> 2][lane])[j & 3] > ki;
				if (!((int)(short)(ki >> 4) < qb->thr[(cnt + rank) & 15]) && rank + 1 > m) { m = rank + 1; kstar = ki; }
			}
		}
	}
	// drop the candidates ranked below the last survivor (:191-192)
	const unsigned lim = (unsigned)(kstar - 1);
#pragma unroll 1
	for (int r = 0; r < 4; r++) {
		const uint4 kk = keys[r][lane];
		uint2 &row = rg.v[o][(y0 + r) & (RR - 1)][lane];
		uint2 wv = row;
		const unsigned m0 = (kk.x - 1u < lim) ? 0xFFFF0000u : 0xFFFFFFFFu, m1 = (kk.y - 1u < lim) ? 0x0000FFFFu : 0xFFFFFFFFu;
		const unsigned m2 = (kk.z - 1u < lim) ? 0xFFFF0000u : 0xFFFFFFFFu, m3 = (kk.w - 1u < lim) ? 0x0000FFFFu : 0xFFFFFFFFu;
		wv.x &= m0 & m1;
		wv.y &= m2 & m3;
		row = wv;
	}
	return m;
}

// lane-private row / key staging of one block for the packed quantiser (ric_quant_pk.cuh): in shared memory ...
template <int RR, class RingT>
struct BlockIO {
	static constexpr int UNROLL = 1;
	RingT &rg;
	KeyRows &keys;
	int o, y0, lane;
	__device__ __forceinline__ uint2 get(int r) const { return rg.v[o][(y0 + r) & (RR - 1)][lane]; }
	__device__ __forceinline__ void put(int r, uint2 v) { rg.v[o][(y0 + r) & (RR - 1)][lane] = v; }
	__device__ __forceinline__ uint2 get_key(int r) const { return *(const uint2 *)&keys[r][lane]; }
	__device__ __forceinline__ void put_key(int r, uint2 v) { *(uint2 *)&keys[r][lane] = v; }
};
// ... or in registers (every index is a compile-time constant once the loops are flat)
struct RegIO {
	static constexpr int UNROLL = 2;
	uint2 rows[4], keyr[4];
	__device__ __forceinline__ uint2 get(int r) const { return rows[r]; }
	__device__ __forceinline__ void put(int r, uint2 v) { rows[r] = v; }
	__device__ __forceinline__ uint2 get_key(int r) const { return keyr[r]; }
	__device__ __forceinline__ void put_key(int r, uint2 v) { keyr[r] = v; }
};

// RingT: anything with `uint2 v[3][RR][32]` (RR rows per band, a power of two); bands [o_begin, o_end).
#ifndef RIC_EXP_FLATQ
#define RIC_EXP_FLATQ true  // level kernels: quantise the block in registers (false: rolled loops over the shared-memory ring)
#endif
template <int RR, bool FLAT = true, class RingT>
__device__ __forceinline__ void flush_blocks_packed(const FwdParams &P, RingT &rg, KeyRows &keys, char *arena,
                                                    unsigned char *flags, const QuantBand *qb3, int bx, int by, int lane,
                                                    bool lane_out, int o_begin = 0, int o_end = 3)
{
#pragma unroll 1
	for (int o = o_begin; o < o_end; o++) {
		const BandRef &b = P.band[o];
		const int x0 = bx * 4, y0 = by * 4;
		if (y0 >= b.dimy) continue;  // warp-uniform
		const bool have = lane_out && x0 < b.dimx;
		const int bw = have ? min(4, b.dimx - x0) : 0, bh = min(4, b.dimy - y0);
		bool mark = false;
		RegIO io;  // FLAT: the block stays in registers from here to the band stores
#pragma unroll
		for (int r = 0; r < 4; r++) io.rows[r] = rg.v[o][(y0 + r) & (RR - 1)][lane];
#ifdef RIC_EXP_NOQUANT  // diagnosis only (wrong results): how fast is the kernel without the quantiser's code?
		if (false) {
#else
		if (P.quant) {
#endif
			const QuantBand *qb = qb3 + o;
			const bool full = bw == 4 && bh == 4;
			const int T = full ? qb->T : qb->Te;
			int cnt = 0;
			// pass 1: anything outside the dead zone in this warp's blocks?  (packed min / max trees; rows and
			// columns beyond the band hold finite garbage, which can only make the answer conservative)
			const uint2 r0 = io.rows[0], r1 = io.rows[1], r2 = io.rows[2], r3 = io.rows[3];
			const unsigned mx = __vmaxs2(__vimax3_s16x2(__vimax3_s16x2(r0.x, r0.y, r1.x), __vimax3_s16x2(r1.y, r2.x, r2.y), r3.x), r3.y);
			const unsigned mn = __vmins2(__vimin3_s16x2(__vimin3_s16x2(r0.x, r0.y, r1.x), __vimin3_s16x2(r1.y, r2.x, r2.y), r3.x), r3.y);
			const int vmax = max(s16lo(mx), s16hi(mx)), vmin = min(s16lo(mn), s16hi(mn));
			const bool any_alive = have && (!qb->fast || vmax > T || vmin < -T);
			bool in_regs = FLAT;  // where the quantised block is: io.rows or the ring
			if (__any_sync(FULL, any_alive)) {
				int nc;
				// the packed form does not cover the -32768 corner of s2u_ (utils.h:95-99): warp-uniform choice
				if (qb->pk && !__any_sync(FULL, vmin == -32768)) {
					if constexpr (FLAT) {
						cnt = quant_rows_pk(qb, bw, bh, io, nc);
						const int ncm = __reduce_max_sync(FULL, nc);  // largest candidate count among this warp's blocks
						if (ncm > 0) cnt += rank_rows_pk(qb, cnt, ncm, io);
					} else {
						BlockIO<RR, RingT> sio{rg, keys, o, y0, lane};
						cnt = quant_rows_pk(qb, bw, bh, sio, nc);
						const int ncm = __reduce_max_sync(FULL, nc);
						if (ncm > 0) cnt += rank_rows_pk(qb, cnt, ncm, sio);
					}
				} else {
					cnt = quant_rows_scalar<RR>(rg, keys, qb, o, y0, bw, bh, lane, nc);
					const int ncm = __reduce_max_sync(FULL, nc);
					if (ncm > 0) cnt += rank_rows_scalar<RR>(rg, keys, qb, o, y0, lane, cnt, ncm);
					in_regs = false;
				}
			} else {
#pragma unroll
				for (int r = 0; r < 4; r++) io.rows[r] = make_uint2(0u, 0u);
				in_regs = true;
			}
			if (!in_regs) {
#pragma unroll
				for (int r = 0; r < 4; r++) io.rows[r] = rg.v[o][(y0 + r) & (RR - 1)][lane];
			}
			int nz = cnt;
			if (P.has_child && bw == 4 && bh == 4) {  // buildTree :267-270: add the four child blocks
				const BandRef &ch = P.child[o];
				const unsigned char *cf = flags + ch.fl_off + (2 * by) * ch.fl_bw + 2 * bx;
				nz += cf[0] + cf[1] + cf[ch.fl_bw] + cf[ch.fl_bw + 1];  // (odd fl_bw: the second row is not 2-byte aligned)
			}
			if (have) flags[b.fl_off + by * b.fl_bw + bx] = nz != 0;
			mark = nz == 0;  // INSIGNIF_BLOCK in the block's first sample, bandcodec.cpp:113,272
		}
		if (!have) continue;
		if (mark) io.rows[0].x = (io.rows[0].x & 0xFFFF0000u) | 0x8000u;
		char *rowp = arena + b.off + ((long long)y0 * b.stride + x0) * 2;
		const long long pitch = (long long)b.stride * 2;
		// (predicated stores instead of these branches were measured 2.5 % slower: profiles/README.md, round 2)
		if (bw == 4 && bh == 4) {  // the usual case: four 8-byte stores off one running pointer
			*(uint2 *)rowp = io.rows[0]; rowp += pitch;
			*(uint2 *)rowp = io.rows[1]; rowp += pitch;
			*(uint2 *)rowp = io.rows[2]; rowp += pitch;
			*(uint2 *)rowp = io.rows[3];
		} else {
#pragma unroll 1
			for (int r = 0; r < bh; r++, rowp += pitch) {
				const uint2 wv = r == 0 ? io.rows[0] : r == 1 ? io.rows[1] : r == 2 ? io.rows[2] : io.rows[3];
				if (bw == 4) *(uint2 *)rowp = wv;
				else {
					short *q = (short *)rowp;
					if (bw > 0) q[0] = (short)(wv.x & 0xFFFF);
					if (bw > 1) q[1] = (short)(wv.x >> 16);
					if (bw > 2) q[2] = (short)(wv.y & 0xFFFF);
				}
			}
		}
	}
}

// One job = one (image, plane, row segment, strip); see the header comment.
template <bool SH, int TRANS, int SRC>
__device__ __forceinline__ void fwd_job(const FwdParams &P, unsigned job, Ring<SH> &rg, KeyRows &keys,
                                        const QuantBand (&s_qb)[2][3], const ColourTab *ctab, int lane)
{
	// plane fastest: the planes of one RGB strip share their u8 loads through L1
	const int plane = (int)(job % (unsigned)P.nplanes); job /= (unsigned)P.nplanes;  // (job ids fit 32 bits: host checks)
	const int sx = (int)(job % (unsigned)P.nstrips); job /= (unsigned)P.nstrips;
	const int sy = (int)(job % (unsigned)P.nsegs);
	const int img = (int)(job / (unsigned)P.nsegs);

	const int w = P.w, h = P.h;
	const int x0 = sx * STRIP_W;
	const int cb = x0 - LANE_W + lane * LANE_W;       // first column of this lane
	const bool col_ok = cb >= 0 && cb < w;            // lane has at least one real column
	const bool lane_out = lane >= 1 && lane <= 30;    // lane owns output columns
	const EdgeX ex = make_edge_x(cb, w, (x0 == 0) || (w <= x0 + STRIP_W + LANE_W));
	constexpr bool NT = SRC == SRC_U8_GRAY || SRC == SRC_U8_RGB;  // rows straight from 8-bit pixels cannot wrap
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
	const int bx = cb >> 3;  // block column of this lane in every band of this level
	const int lx = cb >> 1;  // first LL column of this lane
	char *ll_base;           // LL destination (row 0)
	int ll_rowbytes;
	if (P.ll_to_band) {
		ll_base = arena + P.lband.off;
		ll_rowbytes = P.lband.stride * (SH ? 2 : 4);
	} else {
		ll_base = (char *)P.ll + (img * P.ll_img_stride + plane * P.ll_plane_stride) * (SH ? 2 : 4);
		ll_rowbytes = P.ll_pitch * (SH ? 2 : 4);
	}
	const int ll_dimx = w >> 1;

	// vertical state: so1 raw odd row 2t-1, se1 S1'd even row 2t-2, so2 S2'd odd row 2t-3, se3 S3'd even row 2t-4
	int so1[8], se1[8], so2[8], se3[8];
#pragma unroll
	for (int k = 0; k < 8; k++) so1[k] = se1[k] = so2[k] = se3[k] = 0;

	const int t_begin = (y0 >> 1) - 2, t_last = (y1r >> 1) + 1;
	RawRow<SRC> rawE, rawO;
	constexpr int SES = (SRC == SRC_U8_GRAY || SRC == SRC_U8_RGB) ? 1 : SRC == SRC_S16 ? 2 : 4;  // bytes per source sample
	const long long rowstep = (long long)P.src_pitch * SES;
	const char *prow = (const char *)src + (long long)(2 * t_begin) * rowstep + (long long)cb * SES;  // row 2t, this lane's columns
	{
		const int re = 2 * t_begin, ro = re + 1;
		load_raw<SRC>(rawE, prow, col_ok && re >= 0 && re < h, P.src_plane_stride, plane);
		load_raw<SRC>(rawO, prow + rowstep, col_ok && ro >= 0 && ro < h, P.src_plane_stride, plane);
	}

	// Two source rows per iteration: their conversions and horizontal passes are independent
	// instruction streams that the scheduler interleaves.
#pragma unroll 1
	for (int t = t_begin; t <= t_last; t++) {
		int ne[8], no[8];
		if constexpr (SRC == SRC_U8_RGB) convert_rgb2(rawE, rawO, ne, no, ctab + plane);
		else { convert_raw<SRC>(rawE, ne, plane, P.shift); convert_raw<SRC>(rawO, no, plane, P.shift); }
		{  // prefetch the next row pair
			const int re = 2 * t + 2, ro = re + 1;
			prow += 2 * rowstep;
			load_raw<SRC>(rawE, prow, col_ok && re >= 0 && re < h, P.src_plane_stride, plane);
			load_raw<SRC>(rawO, prow + rowstep, col_ok && ro >= 0 && ro < h, P.src_plane_stride, plane);
		}
		if (ex.on) { row_fwd<SH, TRANS, NT, true>(ne, ex); row_fwd<SH, TRANS, NT, true>(no, ex); }
		else { row_fwd<SH, TRANS, NT, false>(ne, ex); row_fwd<SH, TRANS, NT, false>(no, ex); }

		const int r1 = 2 * t, r2 = 2 * t - 1, r3 = 2 * t - 2, r4 = 2 * t - 3;
		// rows r4-1 .. r1+1 are touched; edge formulas if that range meets row 0 or row h-1
		const bool edge_y = (r4 - 1 <= 0) || (r1 + 1 >= h - 1);
		if (__builtin_expect(edge_y, 0)) {
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
		ring_put<SH>(rg, 0, jd, lane, se1[0], se1[2], se1[4], se1[6]);
		ring_put<SH>(rg, 1, jd, lane, se1[1], se1[3], se1[5], se1[7]);
		ring_put<SH>(rg, 2, jv, lane, so2[0], so2[2], so2[4], so2[6]);
		if (jv >= (y0 >> 1) && jv < (y1 >> 1) && lane_out && lx < ll_dimx) {  // LL row jv
			int l0 = so2[1], l1 = so2[3], l2 = so2[5], l3 = so2[7];
			char *rowp = ll_base + (long long)jv * ll_rowbytes;
			if (P.ll_to_band) {
				if (P.quant) {  // CBand::TSUQ(Quant, 0.5), wavelet2d.cpp:121,124
					const int T = P.llT[cls], iQ = P.lliQ[cls];
					l0 = TR<SH>(l0); l1 = TR<SH>(l1); l2 = TR<SH>(l2); l3 = TR<SH>(l3);  // S4 results arrive un-truncated
					l0 = tsuq1<SH>(l0, T, iQ); l1 = tsuq1<SH>(l1, T, iQ); l2 = tsuq1<SH>(l2, T, iQ); l3 = tsuq1<SH>(l3, T, iQ);
				}
				if (lx + 4 <= ll_dimx) store4<SH>(rowp, lx, l0, l1, l2, l3);
				else {
#pragma unroll
					for (int k = 0; k < 4; k++)
						if (lx + k < ll_dimx) {
							const int lv = k == 0 ? l0 : k == 1 ? l1 : k == 2 ? l2 : l3;
							if (SH) ((short *)rowp)[lx + k] = (short)lv; else ((int *)rowp)[lx + k] = lv;
						}
				}
			} else {
				// scratch rows are padded to a multiple of 8 samples: a full 4-sample store is always in bounds
				store4<SH>(rowp, lx, l0, l1, l2, l3);
			}
		}
		if ((jv & 3) == 3) {  // block row jv>>2 of D, H and V is complete (D/H rows sit one slot ahead in the ring)
			const int by = jv >> 2;
			if (by >= (y0 >> 3)) {
				if constexpr (SH) flush_blocks_packed<RING_ROWS, RIC_EXP_FLATQ>(P, rg, keys, arena, flags, &s_qb[cls][0], bx, by, lane, lane_out);
				else flush_blocks<SH>(P, rg, arena, flags, &s_qb[cls][0], bx, by, lane, lane_out);
			}
		}
		// rotate the vertical state
#pragma unroll
		for (int k = 0; k < 8; k++) { se3[k] = se1[k]; so2[k] = so1[k]; se1[k] = ne[k]; so1[k] = no[k]; }
	}
}

// Persistent warps: every warp repeatedly claims the next job from a global counter, so a launch has
// no wave-quantisation tail beyond one job and the heavier luma jobs (more quantiser work) do not
// hold back the chroma warps of their CTA.  Consecutive job ids are the planes of one strip: they
// are claimed at about the same time, so their shared pixel loads still meet in L1/L2.
template <bool SH, int TRANS, int SRC>
__global__ void __launch_bounds__(fwd_warps(SH) * 32, 4) fwd_level_kernel(const __grid_constant__ FwdParams P)
{
	constexpr int FWD_WARPS = fwd_warps(SH);
	__shared__ QuantBand s_qb[2][3];
	__shared__ Ring<SH> s_ring[FWD_WARPS];
	__shared__ KeyRows s_keys[SH ? FWD_WARPS : 1];
	__shared__ ColourTab s_ctab[SRC == SRC_U8_RGB ? 3 : 1];
	for (int i = threadIdx.x; i < (int)(sizeof(s_qb) / 4); i += blockDim.x) ((int *)s_qb)[i] = ((const int *)P.qb)[i];
	if (SRC == SRC_U8_RGB) colour_tab_fill(s_ctab, P.shift);
	// programmatic dependent launch (ric_b200.cu launch_level): let the next level's CTAs become resident as this
	// grid drains, and do not touch what the previous level wrote before it is complete and visible
	asm volatile("griddepcontrol.launch_dependents;");
	asm volatile("griddepcontrol.wait;" ::: "memory");
	__syncthreads();
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const long long njobs = (long long)P.nstrips * P.nplanes * P.nsegs * P.nimages;
	for (;;) {
		unsigned long long job = 0;
		if (lane == 0) job = atomicAdd(P.counter, 1ull);
		job = __shfl_sync(FULL, job, 0);
		if ((long long)job >= njobs) break;
		fwd_job<SH, TRANS, SRC>(P, (unsigned)job, s_ring[wib], s_keys[SH ? wib : 0], s_qb, s_ctab, lane);
	}
}

}  // namespace ric
