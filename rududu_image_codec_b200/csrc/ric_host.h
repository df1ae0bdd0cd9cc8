// ric_host.h -- host-side geometry and quantiser scalars (plain C++, no CUDA).
//
// Restates, in our own form, the parts of the reference that are host scalars by nature:
//   CWavelet2D::Init     src/lib/wavelet2d.cpp:69-81   band sizes / recursion stop
//   CBand::Init          src/lib/band.cpp:51-65        DimXAlign
//   SetWeight            src/lib/wavelet2d.cpp:1009-1032
//   buildTree prologue   src/lib/bandcodec.cpp:243-247 + makeThres :149-157 + clen :135-147
//   TSUQ / TSUQi prologue src/lib/band.h:68-72,97-100
//   Quants               src/ric/ric.cpp:42-49
// All float arithmetic is single precision in the reference's expression order.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/ric_b200.h"

namespace ric {

struct HostGeom {
	int width, height, channels, levels, level_chg, align, trans;
	int nlev, nbands;
	int lev_w[RIC_MAX_LEVELS], lev_h[RIC_MAX_LEVELS], lev_int[RIC_MAX_LEVELS];
	ric_band_info band[RIC_MAX_BANDS];
	size_t arena_bytes;
	// block-flag area of one plane (one byte per 4x4 block of every D/H/V band)
	int flag_off[RIC_MAX_BANDS], flag_bw[RIC_MAX_BANDS];
	size_t flag_bytes;
};

inline void set_band(ric_band_info &b, int x, int y, int is_int, int align)
{
	const int sz = is_int ? 4 : 2;
	b.dimx = x;
	b.dimy = y;
	b.is_int = is_int;
	b.stride = ((x * sz + align - 1) & -align) / sz;
	b.weight = 1.f;
	b.offset = 0;
}

inline int geom_init(HostGeom &g, int w, int h, int ch, int levels, int level_chg, int align, int trans)
{
	memset(&g, 0, sizeof g);
	if (w < 16 || h < 16 || w > 65535 || h > 65535) return RIC_E_ARG;  // u16 header fields, ric.cpp:150-153
	if (ch != 1 && ch != 3) return RIC_E_ARG;
	if (levels < 1 || levels > RIC_MAX_LEVELS || level_chg < 0 || level_chg >= levels) return RIC_E_ARG;
	if (align < 32 || (align & (align - 1))) return RIC_E_ARG;  // kernels rely on >= 32-byte rows
	if (trans != RIC_CDF97 && trans != RIC_CDF53 && trans != RIC_HAAR) return RIC_E_UNSUPPORTED;
	g.width = w; g.height = h; g.channels = ch; g.levels = levels; g.level_chg = level_chg;
	g.align = align; g.trans = trans;
	int x = w, y = h, lv = levels, n = 0;
	for (;;) {
		const int is_int = lv <= level_chg;
		g.lev_w[n] = x; g.lev_h[n] = y; g.lev_int[n] = is_int;
		set_band(g.band[3 * n + 0], (x + 1) >> 1, (y + 1) >> 1, is_int, align);  // D: even rows, even cols
		set_band(g.band[3 * n + 1], x >> 1, (y + 1) >> 1, is_int, align);        // H: even rows, odd cols
		set_band(g.band[3 * n + 2], (x + 1) >> 1, y >> 1, is_int, align);        // V: odd rows, even cols
		n++;
		if (lv > 1 && x > 15 && y > 15) { x >>= 1; y >>= 1; lv--; continue; }
		set_band(g.band[3 * n], x >> 1, y >> 1, is_int, align);                  // LL of the coarsest level
		break;
	}
	if (x < 8 || y < 8) return RIC_E_ARG;  // lifting needs a few samples per line
	if (trans == RIC_HAAR)  // the reference's Haar leaves odd trailing rows/columns unprocessed and the matching
		for (int i = 0; i < n; i++)  // band samples uninitialised (wavelet2d.cpp:802,838): even sizes only
			if ((g.lev_w[i] | g.lev_h[i]) & 1) return RIC_E_UNSUPPORTED;
	g.nlev = n;
	g.nbands = 3 * n + 1;
	const float scale = trans == RIC_CDF97 ? 1.149604398f * 1.149604398f : 2.f;
	float d = 1.f / scale, v = 1.f, l = 1.f * scale;
	for (int i = 0; i < n; i++) {
		if (i > 0) { d = v; v = l; l = v * scale; }
		g.band[3 * i + 0].weight = d;
		g.band[3 * i + 1].weight = v;
		g.band[3 * i + 2].weight = v;
	}
	g.band[3 * n].weight = l;
	size_t off = 0, foff = 0;
	for (int i = 0; i < g.nbands; i++) {
		ric_band_info &b = g.band[i];
		b.offset = off;
		off += ((size_t)b.stride * b.dimy * (b.is_int ? 4 : 2) + 31) & ~(size_t)31;
		g.flag_off[i] = (int)foff;
		g.flag_bw[i] = (b.dimx + 3) / 4;
		if (i < 3 * n) foff += (((size_t)g.flag_bw[i] * ((b.dimy + 3) / 4)) + 15) & ~(size_t)15;
	}
	g.arena_bytes = off;
	g.flag_bytes = foff ? foff : 16;
	return RIC_OK;
}

inline int quants(int idx)  // ric.cpp:42-49
{
	static const unsigned q5[5] = {0x8000, 0x9000, 0xA800, 0xC000, 0xE000};
	if (idx <= 0) return 0;
	const int k = idx - 1, r = 14 - k / 5;
	return (int16_t)((q5[k % 5] + (1u << (r - 1))) >> r);
}

inline void plane_quant(int q, int ch, int p, int *Quant, int *lambda)  // ric.cpp:163-171
{
	const int boost = (ch == 3 && p != 2) ? 8 : 0;  // C_Q_BOOST on Co (plane 0) and Cg (plane 1)
	*Quant = q ? quants(q + 20 + boost) : 0;
	*lambda = q ? quants(q + 13 + boost) : 0;
}

struct HostQuantBand { int Q, iQ, T, Te, thr[16]; };

inline int trc(int v, int sh) { return sh ? (int)(int16_t)v : v; }

inline void make_quant_band(HostQuantBand &o, int Quant, int lambda, float weight, int is_int)
{
	static const int blen[17] = {20, 40, 55, 66, 75, 81, 85, 88, 89, 88, 85, 81, 75, 66, 55, 40, 20};
	static const int kk[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2};
	static const int mps[16] = {1, 1, 2, 2, 2, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5};
	const int sh = !is_int;
	const int lbda = (int)((float)lambda / weight);
	const int Qarg = trc(Quant, sh);
	int Q = trc((int16_t)((float)Qarg / weight), sh);
	if (Q == 0) Q = 1;
	o.Q = Q;
	o.iQ = (1 << 16) / Q;
	for (int i = 0; i < 16; i++) {
		const int cl = (kk[i] + 1) * 5 + mps[i];  // clen(1, i+1)
		int t = trc((Q + ((lbda * (blen[i + 1] - blen[i] + cl) + 8) >> 4)) & 0xFFFE, sh);
		if (t > Q * 2) t = trc(Q * 2, sh);
		if (t < (Q & 0xFFFE)) t = trc(Q & 0xFFFE, sh);
		o.thr[i] = t;
	}
	o.T = trc(Q >> 1, sh);
	o.Te = trc((Q + ((Q - (Q >> 2)) >> 1)) >> 1, sh);
}

inline void make_tsuq(int Quant, float thres, float weight, int is_int, int *Q, int *iQ, int *T)
{
	int q = (int)((float)Quant / weight);
	if (q == 0) q = 1;
	*Q = q;
	*iQ = (1 << 16) / q;
	*T = trc((int)(thres * (float)q), !is_int);
}

inline int make_tsuqi(int Quant, float weight, int is_int)
{
	const int sh = !is_int;
	int q = trc((int)((float)trc(Quant, sh) / weight), sh);
	return q == 0 ? 1 : q;
}

}  // namespace ric
