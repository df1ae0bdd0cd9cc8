// CPU check of the packed encode quantiser (csrc/ric_quant_pk.cuh: quant_rows_pk + rank_rows_pk) against the
// oracle's restatement of CBandCodec::buildTree / tsuqBlock (oracle/ric_oracle.c; src/lib/bandcodec.cpp:159-322).
// Host code only; links oracle/libric_oracle.so (test infrastructure).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../oracle/ric_oracle.h"
#include "../../rududu_image_codec_b200/csrc/ric_host.h"
#include "../../rududu_image_codec_b200/csrc/ric_quant_pk.cuh"

using namespace ric;

static unsigned rng_state = 777;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }
static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } fails++; } } while (0)

// the host half of the quantiser exactly as ric_b200.cu fill_qb does it (kept in step by hand: small)
static void fill_qb(QuantBand &q, int Quant, int lambda, float weight)
{
	HostQuantBand hq;
	make_quant_band(hq, Quant, lambda, weight, 0);
	q.Q = hq.Q; q.iQ = hq.iQ; q.T = hq.T; q.Te = hq.Te;
	q.fast = hq.Q >= 1 && hq.Q <= 16383;
	for (int i = 0; i < 32; i++) q.kthr[i] = q.kt16[i] = 0x7fffffff;
	for (int i = 0; i < 16; i++) { q.thr[i] = hq.thr[i]; if (q.fast) q.kthr[i] = hq.thr[i] << 4; }
	q.pk = q.fast && hq.Q >= 4 && !(hq.thr[0] & 1);
	for (int i = 0; i < 16 && q.pk; i++) {
		const int d = hq.thr[i] - 2 * hq.T;
		if (d < 0 || d > 2046) q.pk = 0;
		else q.kt16[i] = 0x8000 | (d << 4);
	}
	q.h0 = std::max(hq.thr[0] >> 1, hq.T + 1);
}

template <int U>
struct TestIOT {
	static constexpr int UNROLL = U;
	uint2 rows[4], keys[4];
	uint2 get(int r) const { return rows[r]; }
	void put(int r, uint2 v) { rows[r] = v; }
	uint2 get_key(int r) const { return keys[r]; }
	void put_key(int r, uint2 v) { keys[r] = v; }
};

template <int U>
static void run()
{
	int tested = 0, skipped = 0;
	// one-level geometry: every band is a finest band (no children), 9/7 weights D 1/s, H = V = 1
	for (int iter = 0; iter < 400; iter++) {
		const int w = 32 + (int)(rnd() % 40), h = 32 + (int)(rnd() % 40);
		rico_geom g;
		if (rico_geom_init(&g, w, h, 1, 0, 32, RICO_CDF97)) { printf("geom_init failed\n"); exit(1); }
		static const int qs[] = {32, 42, 96, 126, 288, 380, 672, 1024, 2048, 2705, 9, 4, 5, 1500};
		const int Quant = qs[iter % 14];
		const int lambda = (iter % 5 == 4) ? 0 : (iter % 7 == 6) ? Quant : (int)(Quant / 2.4);
		std::vector<unsigned char> arena(g.arena_bytes), want;
		// coefficients: mostly a few Q wide (dead / candidate / sure mix), some large, rarely the extremes
		for (int i = 0; i < 3; i++) {
			const rico_band &b = g.band[i];
			short *p = (short *)(arena.data() + b.offset);
			for (int k = 0; k < b.stride * b.dimy; k++) {
				const unsigned r = rnd();
				int v;
				if (r % 97 == 0) v = (int)(rnd() % 65536) - 32768;
				else if (r % 13 == 0) v = (int)(rnd() % (8 * Quant + 1)) - 4 * Quant;
				else v = (int)(rnd() % (2 * Quant + Quant / 2 + 1)) - (Quant + Quant / 4);
				if (iter % 3 == 0 && v == -32768) v = -32767;  // keep two thirds of the runs free of the corner that the caller diverts
				p[k] = (short)std::max(-32768, std::min(32767, v));
			}
		}
		want = arena;
		rico_quant(&g, want.data(), Quant, lambda);
		for (int o = 0; o < 3; o++) {
			const rico_band &b = g.band[o];
			QuantBand qb;
			fill_qb(qb, Quant, lambda, b.weight);
			if (!qb.pk) { skipped++; continue; }
			short *p = (short *)(arena.data() + b.offset);
			const short *e = (const short *)(want.data() + b.offset);
			for (int by = 0; by * 4 < b.dimy; by++)
				for (int bx = 0; bx * 4 < b.dimx; bx++) {
					const int bw = std::min(4, b.dimx - 4 * bx), bh = std::min(4, b.dimy - 4 * by);
					TestIOT<U> io;
					uint2 (&rows)[4] = io.rows;
					for (int r = 0; r < 4; r++) io.keys[r] = make_uint2(0xDEADBEEFu, 0x12345678u);  // stale staging
					bool corner = false;
					for (int r = 0; r < 4; r++) {
						unsigned short v[4];
						for (int k = 0; k < 4; k++) {
							// rows / columns beyond the band: finite garbage, as in the kernels' rings
							const bool in = r < bh && k < bw;
							v[k] = in ? (unsigned short)p[(4 * by + r) * b.stride + 4 * bx + k] : (unsigned short)(rnd() % 65535 + 1 - 32768 + 32768);
							if ((short)v[k] == -32768) { if (in) corner = true; else v[k] = 0x8001; }
						}
						rows[r] = make_uint2(v[0] | (unsigned)v[1] << 16, v[2] | (unsigned)v[3] << 16);
					}
					if (corner) { skipped++; continue; }  // flush_blocks_packed sends such warps to the scalar path
					int nc;
					int cnt = quant_rows_pk(&qb, bw, bh, io, nc);
					// the kernel passes the warp's largest candidate count: any value >= nc must give the same result
					const int ncm = nc == 0 ? 0 : std::min(16, nc + (int)(rnd() % 3) * (int)(rnd() % 8));
					if (ncm > 0) cnt += rank_rows_pk(&qb, cnt, ncm, io);
					if (cnt == 0) rows[0].x = (rows[0].x & 0xFFFF0000u) | 0x8000u;  // INSIGNIF_BLOCK (single level: no children)
					for (int r = 0; r < bh; r++)
						for (int k = 0; k < bw; k++) {
							const unsigned wd = k < 2 ? rows[r].x : rows[r].y;
							const short got = (short)((k & 1) ? wd >> 16 : wd & 0xFFFF);
							const short exp = e[(4 * by + r) * b.stride + 4 * bx + k];
							CHECK(got == exp, "Quant %d lambda %d band %d (Q %d T %d h0 %d) block (%d,%d) %dx%d sample (%d,%d): in %d got %d want %d  nc %d ncm %d",
							      Quant, lambda, o, qb.Q, qb.T, qb.h0, bx, by, bw, bh, k, r, p[(4 * by + r) * b.stride + 4 * bx + k], got, exp, nc, ncm);
						}
					tested++;
				}
		}
	}
	printf("unroll %d: blocks tested %d, skipped (not pk / -32768 corner) %d\n", U, tested, skipped);
	CHECK(tested > 50000, "too few blocks tested");
}

int main()
{
	run<1>();
	run<2>();
	if (fails) { printf("quant_pk_test: %d FAILURES\n", fails); return 1; }
	printf("quant_pk_test: ok\n");
	return 0;
}
