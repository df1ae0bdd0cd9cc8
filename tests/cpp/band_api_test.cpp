// The per-band methods of the class shim (include/rududu_b200/wavelet2d.h), used the way the reference's own code
// uses them:  CBandCodec::buildTree<high_band, C> + LBand.TSUQ<C>(Quant, 0.5) = the quantiser half of CodeBand
// (src/lib/wavelet2d.cpp:110-126);  CBand::TSUQ<C> / TSUQi<C> band by band = CWavelet2D::TSUQ / TSUQi
// (wavelet2d.cpp:224-268);  SetWeight(t, baseWeight) (wavelet2d.cpp:1009-1032);  CBand::Init (band.cpp:51-65).
// usage: band_api_test W H LEVELS Quant lambda plane.s16 signed_arena.bin OUT_PREFIX
// writes OUT_PREFIX{bt,tsuq,tsuqi,bw}.bin (arenas) and prints counts; tests/test_cpp_shim.py compares with the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rududu_b200/wavelet2d.h"

using namespace rududu_b200;

static std::vector<char> slurp(const char *path)
{
	FILE *f = fopen(path, "rb");
	if (!f) { perror(path); exit(2); }
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	std::vector<char> b(n);
	if (fread(b.data(), 1, n, f) != (size_t)n) exit(2);
	fclose(f);
	return b;
}

static void dump(const std::string &path, const char *p, size_t n)
{
	FILE *f = fopen(path.c_str(), "wb");
	if (!f) { perror(path.c_str()); exit(2); }
	fwrite(p, 1, n, f);
	fclose(f);
}

template <class F>
static void each_band(CWavelet2D &W, F f)  // every band of the chain, finest level first, LL last
{
	CWavelet2D *c = &W;
	for (; c; c = c->pLow) {
		f(c->DBand, false); f(c->HBand, false); f(c->VBand, false);
		if (!c->pLow) f(c->LBand, true);
	}
}

int main(int argc, char **argv)
{
	if (argc != 9) return 2;
	const int w = atoi(argv[1]), h = atoi(argv[2]), levels = atoi(argv[3]), Quant = atoi(argv[4]), lambda = atoi(argv[5]);
	const std::string out = argv[8];
	try {
		std::vector<char> plane = slurp(argv[6]);
		CWavelet2D Wavelet(w, h, levels, levels > 4 ? levels - 4 : 0);
		Wavelet.SetWeight(cdf97);
		Wavelet.Transform((short *)plane.data(), w, cdf97);
		const std::vector<char> raw(Wavelet.arena(), Wavelet.arena() + Wavelet.arena_bytes());

		// 1. the quantiser half of CodeBand, written out as in wavelet2d.cpp:110-126
		if (Wavelet.DBand.type == sshort) {
			Wavelet.DBand.buildTree<true, short>(Quant, lambda);
			Wavelet.HBand.buildTree<true, short>(Quant, lambda);
			Wavelet.VBand.buildTree<true, short>(Quant, lambda);
		} else {
			Wavelet.DBand.buildTree<true, int>(Quant, lambda);
			Wavelet.HBand.buildTree<true, int>(Quant, lambda);
			Wavelet.VBand.buildTree<true, int>(Quant, lambda);
		}
		CWavelet2D *pCurWav = &Wavelet;
		while (pCurWav->pLow) pCurWav = pCurWav->pLow;
		const unsigned ll_count = pCurWav->LBand.type == sshort ? pCurWav->LBand.TSUQ<short>(Quant, 0.5f) : pCurWav->LBand.TSUQ<int>(Quant, 0.5f);
		dump(out + "bt.bin", Wavelet.arena(), Wavelet.arena_bytes());
		printf("ll_count %u ll_min %d ll_max %d\n", ll_count, pCurWav->LBand.Min, pCurWav->LBand.Max);
		// a parent band on its own: buildTree<false> re-quantising level 1 from its raw samples with the child's pRD
		if (Wavelet.pLow) {
			CBandCodec &P = Wavelet.pLow->VBand;
			const size_t bytes = (size_t)P.BandSize * (P.type == sint ? 4 : 2);
			std::vector<char> after(P.pBand, P.pBand + bytes);
			std::vector<std::vector<char>> coarser;  // the chain above it is re-quantised too: restore all of it
			for (CBand *b = &P; b; b = b->pParent) {
				const size_t nb = (size_t)b->BandSize * (b->type == sint ? 4 : 2);
				coarser.push_back(std::vector<char>(b->pBand, b->pBand + nb));
				memcpy(b->pBand, raw.data() + (b->pBand - Wavelet.arena()), nb);
			}
			if (P.type == sshort) P.buildTree<false, short>(Quant, lambda); else P.buildTree<false, int>(Quant, lambda);
			size_t k = 0;
			for (CBand *b = &P; b; b = b->pParent, k++)
				if (memcmp(b->pBand, coarser[k].data(), coarser[k].size())) { fprintf(stderr, "buildTree<false> differs at chain band %zu\n", k); return 5; }
		}

		// 2. CWavelet2D::TSUQ band by band (LL always with 0.5, wavelet2d.cpp:240-243)
		memcpy(Wavelet.arena(), raw.data(), raw.size());
		unsigned total = 0;
		each_band(Wavelet, [&](CBandCodec &b, bool ll) {
			const float th = ll ? 0.5f : 0.7f;
			total += b.type == sshort ? b.TSUQ<short>(Quant, th) : b.TSUQ<int>(Quant, th);
		});
		dump(out + "tsuq.bin", Wavelet.arena(), Wavelet.arena_bytes());
		printf("tsuq_count %u d_min %d d_max %d\n", total, Wavelet.DBand.Min, Wavelet.DBand.Max);

		// 3. CWavelet2D::TSUQi band by band on signed coefficients
		std::vector<char> sg = slurp(argv[7]);
		if (sg.size() != Wavelet.arena_bytes()) return 4;
		memcpy(Wavelet.arena(), sg.data(), sg.size());
		each_band(Wavelet, [&](CBandCodec &b, bool) {
			if (b.type == sshort) b.TSUQi<short>((short)Quant); else b.TSUQi<int>(Quant);
		});
		dump(out + "tsuqi.bin", Wavelet.arena(), Wavelet.arena_bytes());

		// 4. SetWeight with a base weight: every weight halves, so (Quant, lambda) act like (2 Quant, 2 lambda)
		memcpy(Wavelet.arena(), raw.data(), raw.size());
		const float w1 = Wavelet.VBand.Weight;
		Wavelet.SetWeight(cdf97, 0.5f);
		if (Wavelet.VBand.Weight != 0.5f * w1 || Wavelet.DBand.Weight != 0.5f / (1.149604398f * 1.149604398f)) return 6;
		Wavelet.QuantBands(Quant, lambda);
		dump(out + "bw.bin", Wavelet.arena(), Wavelet.arena_bytes());

		// 5. a band of its own: CBand::Init geometry (band.cpp:51-65) and TSUQ on caller-owned data
		CBandCodec b;
		b.Init(sshort, 37, 21, 32);
		if (b.DimXAlign != 48 || b.BandSize != 48 * 21 || ((uintptr_t)b.pBand & 31)) return 7;
		short *p = (short *)b.pBand;
		for (unsigned i = 0; i < b.BandSize; i++) p[i] = (short)((int)(i * 2654435761u >> 20) % 601 - 300);
		std::vector<short> want(p, p + b.BandSize);
		b.Weight = 2.f;
		const int Q = (int)(Quant / b.Weight), iQ = (1 << 16) / Q;
		const short T = (short)(0.5f * Q);
		unsigned cnt = 0;
		for (unsigned y = 0; y < 21; y++)
			for (unsigned x = 0; x < 37; x++) {
				short &c = want[y * 48 + x];
				if ((unsigned)(c + T) <= (unsigned)(2 * T)) c = 0; else { cnt++; c = (short)((c * iQ + (1 << 15)) >> 16); }
			}
		if (b.TSUQ<short>(Quant, 0.5f) != cnt || memcmp(p, want.data(), want.size() * 2)) { fprintf(stderr, "standalone TSUQ differs\n"); return 8; }
		b.Init(sint, 5, 3, 32);
		if (b.DimXAlign != 8 || b.type != sint) return 9;
		printf("ok\n");
	} catch (const std::exception &e) {
		fprintf(stderr, "%s\n", e.what());
		return 1;
	}
	return 0;
}
