/* oracle/ric_oracle.c -- CPU restatement of the RIC transform+quant hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see ric_oracle.h).  Written from the arithmetic description in
 * SURVEY.md Appendix A as "whole-array" lifting steps (each step reads only results of earlier
 * steps), NOT as the reference's streaming row-pair loop; equality with the reference is
 * established by differential tests against oracle/_ref (tests/test_oracle_vs_ref.py).
 *
 * All working data is held as int32; "C-typed store" (C = short for sshort bands/levels) is
 * modelled by TR(), a wrap to int16.  Every >> is an arithmetic shift (gcc, -fwrapv).
 */
#include "ric_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef int32_t i32;

/* ------------------------------------------------------------------------------------------ */
/* scalars */

int rico_quants(int idx) /* ric.cpp:42-49 */
{
	static const unsigned q5[5] = {0x8000, 0x9000, 0xA800, 0xC000, 0xE000};
	if (idx <= 0) return 0;
	int k = idx - 1, r = 14 - k / 5;
	return (int16_t)((q5[k % 5] + (1u << (r - 1))) >> r);
}

void rico_plane_quant(int q, int ch, int p, int *Quant, int *lambda) /* ric.cpp:163-171 */
{
	int boost = (ch == 3 && p != 2) ? 8 : 0; /* C_Q_BOOST on Co (0) and Cg (1) */
	*Quant = q ? rico_quants(q + 20 + boost) : 0;
	*lambda = q ? rico_quants(q + 13 + boost) : 0;
}

/* ------------------------------------------------------------------------------------------ */
/* geometry: CWavelet2D::Init wavelet2d.cpp:69-81, CBand::Init band.cpp:51-65, SetWeight :1009-1032 */

static void band_set(rico_band *b, int x, int y, int is_int, int align)
{
	int sz = is_int ? 4 : 2;
	b->dimx = x;
	b->dimy = y;
	b->is_int = is_int;
	b->stride = ((x * sz + align - 1) & -align) / sz;
	b->weight = 1.f;
}

int rico_geom_init(rico_geom *g, int w, int h, int levels, int level_chg, int align, int trans)
{
	memset(g, 0, sizeof *g);
	if (w < 8 || h < 8 || levels < 1 || levels > RICO_MAX_LEVELS || align < 4 || (align & (align - 1)))
		return -1;
	g->width = w; g->height = h; g->levels = levels; g->level_chg = level_chg;
	g->align = align; g->trans = trans;
	int x = w, y = h, lv = levels, n = 0;
	for (;;) {
		int is_int = lv <= level_chg;
		g->lev_w[n] = x; g->lev_h[n] = y; g->lev_is_int[n] = is_int;
		band_set(&g->band[3 * n + 0], (x + 1) >> 1, (y + 1) >> 1, is_int, align); /* D */
		band_set(&g->band[3 * n + 1], x >> 1, (y + 1) >> 1, is_int, align);       /* H */
		band_set(&g->band[3 * n + 2], (x + 1) >> 1, y >> 1, is_int, align);       /* V */
		n++;
		if (lv > 1 && x > 15 && y > 15) { x >>= 1; y >>= 1; lv--; continue; }
		band_set(&g->band[3 * n], x >> 1, y >> 1, is_int, align);                 /* LL */
		break;
	}
	g->nlev = n;
	g->nbands = 3 * n + 1;
	/* weights, finest -> coarsest; all float, same expression order as the reference */
	float scale = (trans == RICO_CDF97) ? 1.149604398f * 1.149604398f : 2.f;
	float d = 1.f / scale, v = 1.f, l = 1.f * scale;
	for (int i = 0; i < n; i++) {
		if (i > 0) { float nd = v, nv = l; d = nd; v = nv; l = v * scale; }
		g->band[3 * i + 0].weight = d;
		g->band[3 * i + 1].weight = v;
		g->band[3 * i + 2].weight = v;
	}
	g->band[3 * n].weight = l;
	size_t off = 0;
	for (int i = 0; i < g->nbands; i++) {
		rico_band *b = &g->band[i];
		g->band[i].offset = off;
		off += ((size_t)b->stride * b->dimy * (b->is_int ? 4 : 2) + 31) & ~(size_t)31;
	}
	g->arena_bytes = off;
	return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* 1-D lifting as whole-array steps (wavelet2d.cpp:307-405,593-634,766-786) */

#define TR(v) (sh ? (i32)(int16_t)(v) : (i32)(v))

static i32 m08_int(i32 a) /* mult08<int> */
{
	a -= a >> 2;
	a += a >> 4;
	return a + (a >> 8);
}

static i32 m08_c(i32 a, int sh) /* mult08<C>: every assignment is a C-typed store */
{
	a = TR(a - (a >> 2));
	a = TR(a + (a >> 4));
	return TR(a + (a >> 8));
}

/* x[i*s], i in [0,n); sh: element type is short */
static void fwd97(i32 *x, int n, int s, int sh)
{
#define X(i) x[(size_t)(i) * s]
	int i;
	for (i = 0; i < n; i += 2) { /* S1, even */
		if (i == 0) X(0) = TR(X(0) - 3 * X(1));
		else if (i == n - 1) X(i) = TR(X(i) - 3 * X(i - 1));
		else { i32 t = TR(X(i - 1) + X(i + 1)); X(i) = TR(X(i) - (t + (t >> 1))); }
	}
	for (i = 1; i < n; i += 2) { /* S2, odd */
		if (i == n - 1) X(i) = TR(X(i) - (X(i - 1) >> 3));
		else X(i) = TR(X(i) - ((X(i - 1) + X(i + 1)) >> 4));
	}
	for (i = 0; i < n; i += 2) { /* S3, even */
		if (i == 0) X(0) = TR(X(0) + 2 * m08_c(X(1), sh));
		else if (i == n - 1) X(i) = TR(X(i) + 2 * m08_c(X(i - 1), sh));
		else X(i) = TR(X(i) + m08_int(X(i - 1) + X(i + 1)));
	}
	for (i = 1; i < n; i += 2) { /* S4, odd */
		if (i == n - 1) X(i) = TR(X(i) + (X(i - 1) - (X(i - 1) >> 4)));
		else { i32 t = TR(X(i - 1) + X(i + 1)); X(i) = TR(X(i) + ((t >> 1) - (t >> 5))); }
	}
}

static void inv97(i32 *x, int n, int s, int sh)
{
	int i;
	for (i = 1; i < n; i += 2) { /* U4 */
		if (i == n - 1) X(i) = TR(X(i) - (X(i - 1) - (X(i - 1) >> 4)));
		else { i32 t = TR(X(i - 1) + X(i + 1)); X(i) = TR(X(i) - ((t >> 1) - (t >> 5))); }
	}
	for (i = 0; i < n; i += 2) { /* U3 */
		if (i == 0) X(0) = TR(X(0) - 2 * m08_c(X(1), sh));
		else if (i == n - 1) X(i) = TR(X(i) - 2 * m08_c(X(i - 1), sh));
		else X(i) = TR(X(i) - m08_int(X(i - 1) + X(i + 1)));
	}
	for (i = 1; i < n; i += 2) { /* U2 */
		if (i == n - 1) X(i) = TR(X(i) + (X(i - 1) >> 3));
		else X(i) = TR(X(i) + ((X(i - 1) + X(i + 1)) >> 4));
	}
	for (i = 0; i < n; i += 2) { /* U1 */
		if (i == 0) X(0) = TR(X(0) + 3 * X(1));
		else if (i == n - 1) X(i) = TR(X(i) + 3 * X(i - 1));
		else { i32 t = TR(X(i - 1) + X(i + 1)); X(i) = TR(X(i) + (t + (t >> 1))); }
	}
}

static void fwd53(i32 *x, int n, int s, int sh)
{
	int i;
	for (i = 0; i < n; i += 2) {
		if (i == 0) X(0) = TR(X(0) - X(1));
		else if (i == n - 1) X(i) = TR(X(i) - X(i - 1));
		else X(i) = TR(X(i) - ((X(i - 1) + X(i + 1)) >> 1));
	}
	for (i = 1; i < n; i += 2) {
		if (i == n - 1) X(i) = TR(X(i) + (X(i - 1) >> 1));
		else X(i) = TR(X(i) + ((X(i - 1) + X(i + 1)) >> 2));
	}
}

static void inv53(i32 *x, int n, int s, int sh)
{
	int i;
	for (i = 1; i < n; i += 2) {
		if (i == n - 1) X(i) = TR(X(i) - (X(i - 1) >> 1));
		else X(i) = TR(X(i) - ((X(i - 1) + X(i + 1)) >> 2));
	}
	for (i = 0; i < n; i += 2) {
		if (i == 0) X(0) = TR(X(0) + X(1));
		else if (i == n - 1) X(i) = TR(X(i) + X(i - 1));
		else X(i) = TR(X(i) + ((X(i - 1) + X(i + 1)) >> 1));
	}
}

static void fwd_haar(i32 *x, int n, int s, int sh) /* trailing odd sample untouched */
{
	for (int i = 0; i + 1 < n; i += 2) {
		X(i) = TR(X(i) - X(i + 1));
		X(i + 1) = TR(X(i + 1) + (X(i) >> 1));
	}
}

static void inv_haar(i32 *x, int n, int s, int sh)
{
	for (int i = 0; i + 1 < n; i += 2) {
		X(i + 1) = TR(X(i + 1) - (X(i) >> 1));
		X(i) = TR(X(i) + X(i + 1));
	}
}
#undef X

static void lift1d(i32 *x, int n, int s, int sh, int trans, int inverse)
{
	switch (trans) {
	case RICO_CDF97: (inverse ? inv97 : fwd97)(x, n, s, sh); break;
	case RICO_CDF53: (inverse ? inv53 : fwd53)(x, n, s, sh); break;
	default: (inverse ? inv_haar : fwd_haar)(x, n, s, sh); break;
	}
}

/* ------------------------------------------------------------------------------------------ */
/* band element access */

static i32 bget(const rico_band *b, const void *arena, int y, int x)
{
	const char *p = (const char *)arena + b->offset;
	size_t k = (size_t)y * b->stride + x;
	return b->is_int ? ((const i32 *)p)[k] : ((const int16_t *)p)[k];
}

static i32 bget_flat(const rico_band *b, const void *arena, size_t k)
{
	const char *p = (const char *)arena + b->offset;
	return b->is_int ? ((const i32 *)p)[k] : ((const int16_t *)p)[k];
}

static void bput(const rico_band *b, void *arena, int y, int x, i32 v)
{
	char *p = (char *)arena + b->offset;
	size_t k = (size_t)y * b->stride + x;
	if (b->is_int) ((i32 *)p)[k] = v; else ((int16_t *)p)[k] = (int16_t)v;
}

/* ------------------------------------------------------------------------------------------ */
/* forward: Transform<short> wavelet2d.cpp:926-958 + Transform97/53/Haar */

void rico_forward(const rico_geom *g, const int16_t *plane, int stride, void *arena)
{
	int W = g->width, H = g->height;
	i32 *buf = (i32 *)malloc((size_t)W * H * sizeof(i32));
	for (int y = 0; y < H; y++)
		for (int x = 0; x < W; x++) buf[(size_t)y * W + x] = plane[(size_t)y * stride + x];
	memset(arena, 0, g->arena_bytes);

	for (int lv = 0; lv < g->nlev; lv++) {
		int w = g->lev_w[lv], h = g->lev_h[lv], sh = !g->lev_is_int[lv];
		const rico_band *D = &g->band[3 * lv], *Hb = &g->band[3 * lv + 1], *V = &g->band[3 * lv + 2];
		int last = (lv == g->nlev - 1);
		int hh = h; /* rows that take part (Haar drops an odd last row, wavelet2d.cpp:802) */
		if (g->trans == RICO_HAAR) hh = h & ~1;
		for (int y = 0; y < hh; y++) lift1d(buf + (size_t)y * W, w, 1, sh, g->trans, 0);
		for (int x = 0; x < w; x++) lift1d(buf + x, hh, W, sh, g->trans, 0);
		/* de-interleave: (y even,x even)->D (even,odd)->H (odd,even)->V (odd,odd)->LL */
		for (int y = 0; y < hh; y++)
			for (int x = 0; x < w; x++) {
				i32 v = buf[(size_t)y * W + x];
				if (!(y & 1)) bput((x & 1) ? Hb : D, arena, y >> 1, x >> 1, v);
				else if (!(x & 1)) bput(V, arena, y >> 1, x >> 1, v);
				else if (last) bput(&g->band[3 * g->nlev], arena, y >> 1, x >> 1, v);
			}
		if (!last) { /* LL compacts into the top-left (w/2 x h/2) for the next level */
			int w2 = w >> 1, h2 = h >> 1;
			for (int y = 0; y < h2; y++)
				for (int x = 0; x < w2; x++) buf[(size_t)y * W + x] = buf[(size_t)(2 * y + 1) * W + 2 * x + 1];
		}
	}
	free(buf);
}

/* inverse: TransformI<short> wavelet2d.cpp:960-992 + Transform97I/53I/HaarI */
void rico_inverse(const rico_geom *g, const void *arena, int16_t *plane, int stride, int q1_quirk)
{
	int W = g->width, H = g->height;
	i32 *buf = (i32 *)calloc((size_t)W * H, sizeof(i32));
	i32 *ll = (i32 *)calloc((size_t)W * H, sizeof(i32));
	const rico_band *L = &g->band[3 * g->nlev];
	for (int y = 0; y < L->dimy; y++)
		for (int x = 0; x < L->dimx; x++) ll[(size_t)y * W + x] = bget(L, arena, y, x);

	for (int lv = g->nlev - 1; lv >= 0; lv--) {
		int w = g->lev_w[lv], h = g->lev_h[lv], sh = !g->lev_is_int[lv];
		const rico_band *D = &g->band[3 * lv], *Hb = &g->band[3 * lv + 1], *V = &g->band[3 * lv + 2];
		int hh = (g->trans == RICO_HAAR) ? (h & ~1) : h;
		for (int y = 0; y < hh; y++)
			for (int x = 0; x < w; x++) {
				i32 v;
				if (!(y & 1)) {
					if (x & 1) {
						/* Q1: Transform53I reads image row 2 of H with D's stride (wavelet2d.cpp:715) */
						if (q1_quirk && g->trans == RICO_CDF53 && y == 2)
							v = bget_flat(Hb, arena, (size_t)D->stride + (x >> 1));
						else
							v = bget(Hb, arena, y >> 1, x >> 1);
					} else
						v = bget(D, arena, y >> 1, x >> 1);
				} else if (!(x & 1))
					v = bget(V, arena, y >> 1, x >> 1);
				else
					v = TR(ll[(size_t)(y >> 1) * W + (x >> 1)]); /* int -> short narrowing :971-980 */
				buf[(size_t)y * W + x] = v;
			}
		for (int x = 0; x < w; x++) lift1d(buf + x, hh, W, sh, g->trans, 1);
		for (int y = 0; y < hh; y++) lift1d(buf + (size_t)y * W, w, 1, sh, g->trans, 1);
		for (int y = 0; y < h; y++)
			for (int x = 0; x < w; x++) ll[(size_t)y * W + x] = buf[(size_t)y * W + x];
	}
	for (int y = 0; y < H; y++)
		for (int x = 0; x < W; x++) plane[(size_t)y * stride + x] = (int16_t)ll[(size_t)y * W + x];
	free(buf);
	free(ll);
}

/* ------------------------------------------------------------------------------------------ */
/* encode quantiser: CBandCodec::buildTree/tsuqBlock/makeThres bandcodec.cpp:129-322 */

static const int kBlen[17] = {20, 40, 55, 66, 75, 81, 85, 88, 89, 88, 85, 81, 75, 66, 55, 40, 20};

static int clen1(int cnt) /* clen(1, cnt), bandcodec.cpp:135-147 */
{
	static const int k[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2};
	static const int mps[16] = {1, 1, 2, 2, 2, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5};
	return (k[cnt - 1] + 1) * 5 + mps[cnt - 1];
}

/* unsigned view of a C-typed value: U(short) is 16-bit, U(int) 32-bit (utils.h:107-115) */
static uint32_t uview(i32 v, int sh) { return sh ? (uint32_t)(uint16_t)v : (uint32_t)v; }

static i32 fold(i32 c) /* s2u_, utils.h:95-99: 2|c| + (c<0) */
{
	i32 m = c >> 31;
	return (2 * c + m) ^ (m * 2);
}

static int quant_block_full(const rico_band *b, void *arena, int bx, int by, i32 Q, i32 iQ, const i32 *thr)
{
	int sh = !b->is_int, cnt = 0, nc = 0;
	i32 T = TR(Q >> 1);
	int cand_pos[16];
	i32 val[16];
	for (int k = 0; k < 16; k++) {
		i32 c = bget(b, arena, by * 4 + (k >> 2), bx * 4 + (k & 3));
		if ((uint32_t)(c + T) <= (uint32_t)(2 * T)) { val[k] = 0; continue; }
		i32 f = TR(fold(c));
		if (uview(f, sh) < uview(thr[0], sh)) { val[k] = f; cand_pos[nc++] = k; continue; }
		cnt++;
		i32 a = (i32)(uview(f, sh) >> 1);
		i32 q = (i32)((a * iQ + (1 << 15)) >> 16);
		val[k] = TR((q << 1) | (f & 1));
	}
	if (nc > 0) {
		/* stable descending order on the unsigned view (insertion sort semantics, :115-127) */
		int ord[16];
		for (int i = 0; i < nc; i++) {
			int j = i;
			while (j > 0 && uview(val[ord[j - 1]], sh) < uview(val[cand_pos[i]], sh)) { ord[j] = ord[j - 1]; j--; }
			ord[j] = cand_pos[i];
		}
		int i = nc - 1;
		while (i >= 0 && val[ord[i]] < thr[i + cnt]) val[ord[i--]] = 0; /* signed C compare, :191 */
		cnt += i + 1;
		for (; i >= 0; i--) val[ord[i]] = 2 | (val[ord[i]] & 1);
	}
	for (int k = 0; k < 16; k++) bput(b, arena, by * 4 + (k >> 2), bx * 4 + (k & 3), val[k]);
	return cnt;
}

static int quant_block_edge(const rico_band *b, void *arena, int x0, int y0, int bw, int bh, i32 Q, i32 iQ)
{
	int sh = !b->is_int, cnt = 0;
	i32 T = TR((Q + ((Q - (Q >> 2)) >> 1)) >> 1);
	for (int j = 0; j < bh; j++)
		for (int i = 0; i < bw; i++) {
			i32 c = bget(b, arena, y0 + j, x0 + i);
			if ((uint32_t)(c + T) <= (uint32_t)(2 * T)) { bput(b, arena, y0 + j, x0 + i, 0); continue; }
			i32 f = TR(fold(c));
			cnt++;
			i32 a = (i32)(uview(f, sh) >> 1);
			i32 q = (i32)((a * iQ + (1 << 15)) >> 16);
			bput(b, arena, y0 + j, x0 + i, TR((q << 1) | (f & 1)));
		}
	return cnt;
}

static void quant_chain(const rico_geom *g, void *arena, int orient, int Quant, int lambda)
{
	uint32_t *rd_child = NULL;
	int child_bw = 0;
	for (int lv = 0; lv < g->nlev; lv++) {
		const rico_band *b = &g->band[3 * lv + orient];
		int sh = !b->is_int;
		/* host scalars, float exactly as written (:243-247) */
		int lbda = (int)(lambda / b->weight);
		i32 Qarg = TR(Quant); /* parameter `const C Quant` */
		i32 Q = TR((int16_t)(Qarg / b->weight));
		if (Q == 0) Q = 1;
		i32 iQ = (1 << 16) / Q;
		i32 thr[16];
		for (int i = 0; i < 16; i++) { /* makeThres :149-157 */
			thr[i] = TR((Q + ((lbda * (kBlen[i + 1] - kBlen[i] + clen1(i + 1)) + 8) >> 4)) & 0xFFFE);
			if (thr[i] > Q * 2) thr[i] = TR(Q * 2);
			if (thr[i] < (Q & 0xFFFE)) thr[i] = TR(Q & 0xFFFE);
		}
		int bw = (b->dimx + 3) / 4, bh = (b->dimy + 3) / 4;
		uint32_t *rd = (uint32_t *)calloc((size_t)bw * bh, sizeof(uint32_t));
		for (int by = 0; by < bh; by++)
			for (int bx = 0; bx < bw; bx++) {
				int fw = b->dimx - bx * 4, fh = b->dimy - by * 4;
				if (fw > 4) fw = 4;
				if (fh > 4) fh = 4;
				long long dist;
				if (fw == 4 && fh == 4) {
					dist = quant_block_full(b, arena, bx, by, Q, iQ, thr);
					if (lv > 0) /* !high_band: add the 4 child blocks' counts (:267-270) */
						dist += (long long)rd_child[(size_t)(2 * by) * child_bw + 2 * bx] +
						        rd_child[(size_t)(2 * by) * child_bw + 2 * bx + 1] +
						        rd_child[(size_t)(2 * by + 1) * child_bw + 2 * bx] +
						        rd_child[(size_t)(2 * by + 1) * child_bw + 2 * bx + 1];
				} else
					dist = quant_block_edge(b, arena, bx * 4, by * 4, fw, fh, Q, iQ);
				if (dist <= 0) {
					bput(b, arena, by * 4, bx * 4, -0x8000); /* INSIGNIF_BLOCK :113 */
					rd[(size_t)by * bw + bx] = 0;
				} else
					rd[(size_t)by * bw + bx] = dist > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)dist;
			}
		free(rd_child);
		rd_child = rd;
		child_bw = bw;
	}
	free(rd_child);
}

/* CBand::TSUQ<C> band.h:65-92 */
static unsigned tsuq_band(const rico_band *b, void *arena, int Quant, float thres)
{
	int sh = !b->is_int;
	int Q = (int)(Quant / b->weight);
	if (Q == 0) Q = 1;
	int iQ = (1 << 16) / Q;
	i32 T = TR((i32)(thres * Q));
	unsigned count = 0;
	for (int y = 0; y < b->dimy; y++)
		for (int x = 0; x < b->dimx; x++) {
			i32 c = bget(b, arena, y, x);
			if ((uint32_t)(c + T) <= (uint32_t)(2 * T)) bput(b, arena, y, x, 0);
			else { count++; bput(b, arena, y, x, TR((c * iQ + (1 << 15)) >> 16)); }
		}
	return count;
}

void rico_quant(const rico_geom *g, void *arena, int Quant, int lambda)
{
	for (int o = 0; o < 3; o++) quant_chain(g, arena, o, Quant, lambda);
	tsuq_band(&g->band[3 * g->nlev], arena, Quant, 0.5f);
}

unsigned rico_tsuq_all(const rico_geom *g, void *arena, int Quant, float thres)
{
	unsigned n = 0;
	for (int i = 0; i < 3 * g->nlev; i++) n += tsuq_band(&g->band[i], arena, Quant, thres);
	return n + tsuq_band(&g->band[3 * g->nlev], arena, Quant, 0.5f); /* LL always 0.5 (:240-243) */
}

/* CBand::TSUQi<C> band.h:94-107 via CWavelet2D::TSUQi wavelet2d.cpp:248-268 */
void rico_tsuqi(const rico_geom *g, void *arena, int Quant)
{
	for (int i = 0; i < g->nbands; i++) {
		const rico_band *b = &g->band[i];
		int sh = !b->is_int;
		i32 Qc = TR(Quant);
		i32 Q = TR((i32)(Qc / b->weight));
		if (Q == 0) Q = 1;
		for (int y = 0; y < b->dimy; y++)
			for (int x = 0; x < b->dimx; x++) bput(b, arena, y, x, TR(bget(b, arena, y, x) * Q));
	}
}

void rico_unfold(const rico_geom *g, void *arena) /* u2s_ utils.h:101-105 on folded HF bands */
{
	for (int i = 0; i < 3 * g->nlev; i++) {
		const rico_band *b = &g->band[i];
		int sh = !b->is_int;
		for (int y = 0; y < b->dimy; y++)
			for (int x = 0; x < b->dimx; x++) {
				i32 f = bget(b, arena, y, x);
				if (f == -0x8000) { bput(b, arena, y, x, 0); continue; }
				i32 u = (i32)uview(f, sh);
				i32 m = (i32)((uint32_t)u << 31) >> 31;
				bput(b, arena, y, x, TR(((u >> 1) + m) ^ m));
			}
	}
}

/* ------------------------------------------------------------------------------------------ */
/* colour: RGBtoYCoCg / YCoCgtoRGB ric.cpp:76-112, gray paths :143-148,:227-240 (SHIFT = 4) */

static i32 s16(i32 v) { return (int16_t)v; }
static i32 clip255(i32 v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

void rico_colour_fwd(const uint8_t *src, int w, int h, int ch, int q, int16_t *planes)
{
	size_t n = (size_t)w * h;
	for (size_t i = 0; i < n; i++) {
		if (ch == 3) {
			i32 r = src[i], gg = src[n + i], bl = src[2 * n + i];
			i32 co = s16(r - bl);
			i32 t = s16(bl + (co >> 1));
			i32 cg = s16(gg - t);
			i32 y = s16(t + ((cg >> 1) - 128));
			if (q) { co = s16(co << 3); cg = s16(cg << 3); y = s16(y << 4); }
			planes[i] = (int16_t)co; planes[n + i] = (int16_t)cg; planes[2 * n + i] = (int16_t)y;
		} else
			planes[i] = (int16_t)(q ? (src[i] - 128) << 4 : src[i] - 128);
	}
}

void rico_colour_inv(const int16_t *planes, int w, int h, int ch, int q, uint8_t *dst)
{
	size_t n = (size_t)w * h;
	for (size_t i = 0; i < n; i++) {
		if (ch == 3) {
			i32 co = planes[i], cg = planes[n + i], y = planes[2 * n + i];
			if (q) { co = s16((co + 4) >> 3); cg = s16((cg + 4) >> 3); y = s16((y + 8) >> 4); }
			y = s16(y - ((cg >> 1) - 128));
			cg = s16(cg + y);
			y = s16(y - (co >> 1));
			co = s16(co + y);
			if (q) { co = clip255(co); cg = clip255(cg); y = clip255(y); }
			dst[i] = (uint8_t)co; dst[n + i] = (uint8_t)cg; dst[2 * n + i] = (uint8_t)y;
		} else {
			i32 v = planes[i];
			if (q) v = clip255(s16(128 + ((v + 8) >> 4)));
			else v = s16(v + 128);
			dst[i] = (uint8_t)v;
		}
	}
}

/* ------------------------------------------------------------------------------------------ */
/* whole stage as ric drives it (CompressImage ric.cpp:159-171, DecompressImage :209-246) */

void rico_encode_image(const rico_geom *g, const uint8_t *src, int ch, int q, void *arenas)
{
	size_t n = (size_t)g->width * g->height;
	int16_t *planes = (int16_t *)malloc(n * ch * sizeof(int16_t));
	rico_colour_fwd(src, g->width, g->height, ch, q, planes);
	for (int p = 0; p < ch; p++) {
		int Q, lam;
		char *arena = (char *)arenas + (size_t)p * g->arena_bytes;
		rico_plane_quant(q, ch, p, &Q, &lam);
		rico_forward(g, planes + (size_t)p * n, g->width, arena);
		rico_quant(g, arena, Q, lam);
	}
	free(planes);
}

void rico_decode_image(const rico_geom *g, const void *arenas, int ch, int q, uint8_t *dst)
{
	size_t n = (size_t)g->width * g->height;
	int16_t *planes = (int16_t *)malloc(n * ch * sizeof(int16_t));
	char *tmp = (char *)malloc(g->arena_bytes);
	for (int p = 0; p < ch; p++) {
		int Q, lam;
		rico_plane_quant(q, ch, p, &Q, &lam);
		memcpy(tmp, (const char *)arenas + (size_t)p * g->arena_bytes, g->arena_bytes);
		if (q) rico_tsuqi(g, tmp, Q);
		rico_inverse(g, tmp, planes + (size_t)p * n, g->width, 1);
	}
	rico_colour_inv(planes, g->width, g->height, ch, q, dst);
	free(tmp);
	free(planes);
}
