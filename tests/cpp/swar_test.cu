// CPU check of the packed-linear lifting arithmetic (csrc/ric_swar.cuh) against a direct transcription of
// the reference's 1-D 9/7 forward lifting on `short` (src/lib/wavelet2d.cpp:320-359).
// Host code only (built by nvcc because the header is __host__ __device__); no GPU needed.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../rududu_image_codec_b200/csrc/ric_swar.cuh"

using namespace ric::sw;

static int m08i(int a) { a -= a >> 2; a += a >> 4; return a + (a >> 8); }
static short m08s(short a) { a -= a >> 2; a += a >> 4; return a + (a >> 8); }

// reference: in place, C = short (every store truncates)
static void ref_line(std::vector<short> &x)
{
	const int n = (int)x.size();
	for (int i = 0; i < n; i += 2) {  // S1
		if (i == 0) x[i] -= 3 * x[1];
		else if (i == n - 1) x[i] -= 3 * x[i - 1];
		else { short t = x[i - 1] + x[i + 1]; x[i] -= t + (t >> 1); }
	}
	for (int i = 1; i < n; i += 2) {  // S2
		if (i == n - 1) x[i] -= x[i - 1] >> 3;
		else x[i] -= (x[i - 1] + x[i + 1]) >> 4;
	}
	for (int i = 0; i < n; i += 2) {  // S3
		if (i == 0) x[i] += 2 * m08s(x[1]);
		else if (i == n - 1) x[i] += 2 * m08s(x[i - 1]);
		else x[i] += m08i(x[i - 1] + x[i + 1]);
	}
	for (int i = 1; i < n; i += 2) {  // S4
		if (i == n - 1) x[i] += x[i - 1] - (x[i - 1] >> 4);
		else { short t = x[i - 1] + x[i + 1]; x[i] += (t >> 1) - (t >> 5); }
	}
}

// packed: two lines at once (lo = line a, hi = line b); edges = also apply the edge formulas
template <class K>
static void packed_line(std::vector<unsigned> &X, bool edges)
{
	const int n = (int)X.size();
	std::vector<unsigned> o(X);
	for (int i = 0; i < n; i += 2) {
		if (i == 0) { if (edges) X[i] = s1_edge<K>(o[i], X[1]); }
		else if (i == n - 1) { if (edges) X[i] = s1_edge<K>(o[i], X[i - 1]); }
		else X[i] = s1<K>(o[i], X[i - 1], X[i + 1]);
	}
	o = X;
	for (int i = 1; i < n; i += 2) {
		if (i == n - 1) { if (edges) X[i] = s2_last<K>(o[i], X[i - 1]); }
		else X[i] = s2<K>(o[i], X[i - 1], X[i + 1]);
	}
	o = X;
	for (int i = 0; i < n; i += 2) {
		if (i == 0) { if (edges) X[i] = s3_edge<K>(o[i], X[1]); }
		else if (i == n - 1) { if (edges) X[i] = s3_edge<K>(o[i], X[i - 1]); }
		else X[i] = s3<K>(o[i], X[i - 1], X[i + 1]);
	}
	o = X;
	for (int i = 1; i < n; i += 2) {
		if (i == n - 1) { if (edges) X[i] = s4_last<K>(o[i], X[i - 1]); }
		else X[i] = s4<K>(o[i], X[i - 1], X[i + 1]);
	}
}

static unsigned rng_state = 12345;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } fails++; } } while (0)

// row pass: any input in [-2048, 2047] (what 8-bit pixels give after the colour transform and up-shift)
static void test_rows()
{
	for (int iter = 0; iter < 4000; iter++) {
		const int n = 6 + (int)(rnd() % 60);
		const int mode = iter % 5;
		std::vector<short> a(n), b(n);
		for (int i = 0; i < n; i++) {
			if (mode == 0) { a[i] = (short)((int)(rnd() % 4096) - 2048); b[i] = (short)((int)(rnd() % 4096) - 2048); }
			else if (mode == 1) { a[i] = (i & 1) ? 2047 : -2048; b[i] = (i & 1) ? -2048 : 2047; }      // worst-case alternation
			else if (mode == 2) { a[i] = ((i >> 1) & 1) ? 2047 : -2048; b[i] = (rnd() & 1) ? 2047 : -2048; }
			else if (mode == 3) { a[i] = -2048; b[i] = 2047; }
			else { a[i] = (short)(((rnd() & 1) ? 2047 : -2048)); b[i] = (short)((rnd() % 3 == 0) ? -2048 : (int)(rnd() % 4096) - 2048); }
		}
		std::vector<unsigned> X(n);
		for (int i = 0; i < n; i++) X[i] = enc(a[i], b[i], (i & 1) ? KH::O0 : KH::E0);
		ref_line(a); ref_line(b);
		packed_line<KH>(X, true);
		for (int i = 0; i < n; i++) {
			const unsigned K = (i & 1) ? KH::O4 : KH::E3;
			CHECK(dec_lo(X[i], K) == a[i] && dec_hi(X[i], K) == b[i], "row n=%d i=%d mode=%d: got (%d,%d) want (%d,%d)", n, i, mode,
			      dec_lo(X[i], K), dec_hi(X[i], K), a[i], b[i]);
			// the halves must be directly readable (no borrow) for the transposition that follows
			const unsigned c2 = to_c2(X[i], K);
			CHECK((short)(c2 & 0xFFFF) == a[i] && (short)(c2 >> 16) == b[i], "row c2 n=%d i=%d", n, i);
			CHECK(from_c2(c2, K) == X[i], "from_c2");
		}
	}
}

// column pass: inputs that pass the guard (|odd| <= 4095, |S1'd even| <= 8191, |even| <= 17002) must be exact
static void test_cols()
{
	int used = 0;
	for (int iter = 0; iter < 20000; iter++) {
		const int n = 8 + (int)(rnd() % 40);
		const int mode = iter % 6;
		std::vector<short> a(n), b(n);
		for (int i = 0; i < n; i++) {
			int oa, ob, ea, eb;
			if (mode <= 1) { oa = (int)(rnd() % 8191) - 4095; ob = (int)(rnd() % 8191) - 4095; }
			else if (mode == 2) { oa = (rnd() & 1) ? 4095 : -4095; ob = ((i >> 1) & 1) ? 4095 : -4095; }
			else if (mode == 3) { oa = 4095; ob = -4095; }
			else { oa = (rnd() % 3) ? ((rnd() & 1) ? 4095 : -4095) : (int)(rnd() % 8191) - 4095; ob = (int)(rnd() % 2001) - 1000; }
			ea = (int)(rnd() % 34005) - 17002; eb = (int)(rnd() % 34005) - 17002;
			a[i] = (short)((i & 1) ? oa : ea); b[i] = (short)((i & 1) ? ob : eb);
		}
		// choose even samples so that the S1 result lands on a target inside +-8191 (often at the bound)
		for (int i = 2; i + 1 < n; i += 2) {
			for (int f = 0; f < 2; f++) {
				std::vector<short> &v = f ? b : a;
				short t = v[i - 1] + v[i + 1];
				int target = (mode & 1) ? ((rnd() & 1) ? 8191 : -8191) : (int)(rnd() % 16383) - 8191;
				if (mode == 5) target = (rnd() & 1) ? 8191 : -8191;
				int e = target + t + (t >> 1);
				if (e > 17002) e = 17002;
				if (e < -17002) e = -17002;
				v[i] = (short)e;
			}
		}
		std::vector<unsigned> X(n);
		for (int i = 0; i < n; i++) X[i] = enc(a[i], b[i], (i & 1) ? KV::O0 : KV::E0);
		// guard on the raw odd samples and on the interior S1 results, as the kernel applies it
		bool ok = true;
		for (int i = 1; i < n; i += 2) ok = ok && (X[i] & GUARD_O) == 0;
		std::vector<unsigned> Y(X);
		for (int i = 2; i + 1 < n; i += 2) ok = ok && (s1<KV>(Y[i], Y[i - 1], Y[i + 1]) & GUARD_E) == 0;
		if (!ok) continue;
		used++;
		std::vector<short> ra(a), rb(b);
		ref_line(ra); ref_line(rb);
		packed_line<KV>(X, false);  // interior formulas only: the kernel sends edge rows to the scalar path
		// samples whose dependency cone touches an edge sample are not comparable: skip 4 at each end
		for (int i = 5; i < n - 5; i++) {
			const unsigned K = (i & 1) ? KV::O4 : KV::E3;
			CHECK(dec_lo(X[i], K) == ra[i] && dec_hi(X[i], K) == rb[i], "col n=%d i=%d mode=%d: got (%d,%d) want (%d,%d)", n, i, mode,
			      dec_lo(X[i], K), dec_hi(X[i], K), ra[i], rb[i]);
		}
	}
	CHECK(used > 4000, "guard rejected too many column tests (%d)", used);
	printf("column tests that passed the guard: %d\n", used);
}

// the guard must reject exactly the out-of-range operands
static void test_guard()
{
	for (int v = -32768; v <= 32767; v++) {
		const unsigned r1 = enc(v, 0, KV::O0), r2 = enc(0, v, KV::O0);
		const bool in = v >= -4096 && v <= 4095;
		CHECK(((r1 & GUARD_O) == 0) == in && ((r2 & GUARD_O) == 0) == in, "GUARD_O v=%d", v);
		const unsigned e1 = enc(v, 0, KV::E1), e2 = enc(0, v, KV::E1);
		const bool ine = v >= -8192 && v <= 8191;
		CHECK(((e1 & GUARD_E) == 0) == ine && ((e2 & GUARD_E) == 0) == ine, "GUARD_E v=%d", v);
	}
	// two out-of-range halves never cancel into a pass
	for (int i = 0; i < 200000; i++) {
		const int a = (int)(rnd() % 65536) - 32768, b = (int)(rnd() % 65536) - 32768;
		const bool in = a >= -4096 && a <= 4095 && b >= -4096 && b <= 4095;
		CHECK(((enc(a, b, KV::O0) & GUARD_O) == 0) == in, "GUARD_O pair %d %d", a, b);
		const bool ine = a >= -8192 && a <= 8191 && b >= -8192 && b <= 8191;
		CHECK(((enc(a, b, KV::E1) & GUARD_E) == 0) == ine, "GUARD_E pair %d %d", a, b);
	}
}

// ---- inverse ----------------------------------------------------------------------------------------------
// reference: the step-parallel form of TransLine97I, C = short
static void ref_line_inv(std::vector<short> &x)
{
	const int n = (int)x.size();
	for (int i = 1; i < n; i += 2) {  // U4
		if (i == n - 1) x[i] -= x[i - 1] - (x[i - 1] >> 4);
		else { short t = x[i - 1] + x[i + 1]; x[i] -= (t >> 1) - (t >> 5); }
	}
	for (int i = 0; i < n; i += 2) {  // U3
		if (i == 0) x[i] -= 2 * m08s(x[1]);
		else if (i == n - 1) x[i] -= 2 * m08s(x[i - 1]);
		else x[i] -= m08i(x[i - 1] + x[i + 1]);
	}
	for (int i = 1; i < n; i += 2) {  // U2
		if (i == n - 1) x[i] += x[i - 1] >> 3;
		else x[i] += (x[i - 1] + x[i + 1]) >> 4;
	}
	for (int i = 0; i < n; i += 2) {  // U1
		if (i == 0) x[i] += 3 * x[1];
		else if (i == n - 1) x[i] += 3 * x[i - 1];
		else { short t = x[i - 1] + x[i + 1]; x[i] += t + (t >> 1); }
	}
}

// packed inverse of two lines at once; returns false when any produced value fails the guard
static bool packed_line_inv(std::vector<unsigned> &X)
{
	const int n = (int)X.size();
	unsigned acc = 0;
	for (int i = 0; i < n; i++) acc |= X[i];
	std::vector<unsigned> o(X);
	for (int i = 1; i < n; i += 2) { X[i] = i == n - 1 ? u4_last(o[i], o[i - 1]) : u4(o[i], o[i - 1], o[i + 1]); acc |= X[i]; }
	o = X;
	for (int i = 0; i < n; i += 2) { X[i] = i == 0 ? u3_edge(o[i], o[1]) : i == n - 1 ? u3_edge(o[i], o[i - 1]) : u3(o[i], o[i - 1], o[i + 1]); acc |= X[i]; }
	o = X;
	for (int i = 1; i < n; i += 2) { X[i] = i == n - 1 ? u2_last(o[i], o[i - 1]) : u2(o[i], o[i - 1], o[i + 1]); acc |= X[i]; }
	o = X;
	for (int i = 0; i < n; i += 2) { X[i] = i == 0 ? u1_edge(o[i], o[1]) : i == n - 1 ? u1_edge(o[i], o[i - 1]) : u1(o[i], o[i - 1], o[i + 1]); acc |= X[i]; }
	return (acc & GUARD_I) == 0;
}

static void test_inverse()
{
	int used = 0, rejected = 0;
	for (int iter = 0; iter < 30000; iter++) {
		const int n = 6 + (int)(rnd() % 50);
		const int mode = iter % 6;
		std::vector<short> a(n), b(n);
		if (mode < 3) {  // coefficients of a smooth-ish signal: forward-transform bounded data, perturb a little
			const int amp = mode == 0 ? 2048 : mode == 1 ? 1000 : 300;
			for (int i = 0; i < n; i++) { a[i] = (short)((int)(rnd() % (2 * amp)) - amp); b[i] = (short)((i * 37 % (2 * amp)) - amp); }
			ref_line(a); ref_line(b);
			for (int i = 0; i < n; i++) { a[i] += (short)((int)(rnd() % 9) - 4); b[i] += (short)((int)(rnd() % 65) - 32); }
		} else {
			const int amp = mode == 3 ? 8191 : mode == 4 ? 3000 : 600;  // arbitrary coefficients, up to the bound itself
			for (int i = 0; i < n; i++) { a[i] = (short)((int)(rnd() % (2 * amp + 1)) - amp); b[i] = (short)((rnd() & 1) ? amp : -amp); }
		}
		std::vector<unsigned> X(n);
		for (int i = 0; i < n; i++) X[i] = enc(a[i], b[i], G);
		const bool ok = packed_line_inv(X);
		if (!ok) { rejected++; continue; }
		used++;
		ref_line_inv(a); ref_line_inv(b);
		for (int i = 0; i < n; i++)
			CHECK(dec_lo(X[i], G) == a[i] && dec_hi(X[i], G) == b[i], "inverse n=%d i=%d mode=%d: got (%d,%d) want (%d,%d)", n, i, mode,
			      dec_lo(X[i], G), dec_hi(X[i], G), a[i], b[i]);
	}
	printf("inverse lines that passed the guard: %d (rejected %d)\n", used, rejected);
	CHECK(used > 8000 && rejected > 1000, "inverse test mix is off");
	// the guard is exactly the range test
	for (int v = -32768; v <= 32767; v++) {
		const bool in = v >= -8192 && v <= 8191;
		CHECK(((enc(v, 0, G) & GUARD_I) == 0) == in && ((enc(0, v, G) & GUARD_I) == 0) == in, "GUARD_I v=%d", v);
	}
	// dequantisation
	for (int i = 0; i < 200000; i++) {
		const int q = 1 + (int)(rnd() % 700), lim = 8191 / q;
		const int c0 = (int)(rnd() % (2 * lim + 1)) - lim, c1 = (int)(rnd() % (2 * lim + 1)) - lim;
		const unsigned c2 = (unsigned)(c0 & 0xFFFF) | (unsigned)c1 << 16;
		const unsigned r = dequant(c2, (unsigned)q, G - OB * (unsigned)q);
		CHECK(dec_lo(r, G) == c0 * q && dec_hi(r, G) == c1 * q && (r & GUARD_I) == 0, "dequant %d %d q %d", c0, c1, q);
	}
}

static int clip255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
static void test_pixels()
{
	for (int i = 0; i < 400000; i++) {
		int v[2][3];
		for (int h = 0; h < 2; h++)
			for (int c = 0; c < 3; c++) v[h][c] = (i % 3 == 0) ? ((rnd() & 1) ? 8191 : -8192) : (int)(rnd() % 16384) - 8192;
		unsigned r, g, b;
		ycocg_out(enc(v[0][0], v[1][0], G), enc(v[0][1], v[1][1], G), enc(v[0][2], v[1][2], G), r, g, b);
		for (int h = 0; h < 2; h++) {
			short co = (short)v[h][0], cg = (short)v[h][1], y = (short)v[h][2];  // ric.cpp:98-110
			co = (co + 4) >> 3; cg = (cg + 4) >> 3; y = (y + 8) >> 4;
			y -= (cg >> 1) - 128;
			cg += y;
			y -= co >> 1;
			co += y;
			const int R = clip255(co), Gv = clip255(cg), B = clip255(y);
			const int gr = (r >> (16 * h)) & 0xFF, gg = (g >> (16 * h)) & 0xFF, gb = (b >> (16 * h)) & 0xFF;
			CHECK(gr == R && gg == Gv && gb == B, "ycocg half %d in (%d,%d,%d): got (%d,%d,%d) want (%d,%d,%d)", h, v[h][0], v[h][1], v[h][2], gr, gg, gb, R, Gv, B);
			CHECK(((r >> (16 * h)) & 0xFF00) == 0x4100 && ((g >> (16 * h)) & 0xFF00) == 0x4100 && ((b >> (16 * h)) & 0xFF00) == 0x4100, "pixel constant");
		}
		const unsigned go = gray_out(enc(v[0][0], v[1][0], G));
		for (int h = 0; h < 2; h++)
			CHECK((int)((go >> (16 * h)) & 0xFF) == clip255(128 + ((v[h][0] + 8) >> 4)), "gray %d", v[h][0]);
	}
}

int main()
{
	test_inverse();
	test_pixels();
	test_rows();
	test_cols();
	test_guard();
	if (fails) { printf("swar_test: %d FAILURES\n", fails); return 1; }
	printf("swar_test: ok\n");
	return 0;
}
