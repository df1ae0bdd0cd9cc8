// Drives the C++ class shim (include/rududu_b200/wavelet2d.h) the way ric.cpp drives the reference:
//   encode: Transform -> QuantBands (the quantiser half of CodeBand)      ric.cpp:159-171
//   decode: TSUQi -> TransformI with the one-past-end pointer              ric.cpp:209-225
// usage: shim_test W H LEVELS Quant lambda plane.s16 arena_out.bin signed_arena_in.bin plane_out.s16
#include <cstdio>
#include <cstdlib>
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

int main(int argc, char **argv)
{
	if (argc != 10) return 2;
	const int w = atoi(argv[1]), h = atoi(argv[2]), levels = atoi(argv[3]), Quant = atoi(argv[4]), lambda = atoi(argv[5]);
	try {
		std::vector<char> plane = slurp(argv[6]);
		CWavelet2D Wavelet(w, h, levels, levels - 4);
		Wavelet.SetWeight(cdf97);
		Wavelet.Transform((short *)plane.data(), w, cdf97);
		Wavelet.QuantBands(Quant, lambda);
		// band geometry sanity, as the entropy stage would walk it
		int nlev = 0;
		for (CWavelet2D *c = &Wavelet; c; c = c->pLow) nlev++;
		if (Wavelet.DBand.pBand != Wavelet.arena() || Wavelet.DBand.type != sshort) return 3;
		FILE *f = fopen(argv[7], "wb");
		fwrite(Wavelet.arena(), 1, Wavelet.arena_bytes(), f);
		fclose(f);
		// decode side
		std::vector<char> sg = slurp(argv[8]);
		if (sg.size() != Wavelet.arena_bytes()) return 4;
		for (size_t i = 0; i < sg.size(); i++) Wavelet.arena()[i] = sg[i];
		Wavelet.TSUQi(Quant);
		std::vector<short> out((size_t)w * h);
		Wavelet.TransformI(out.data() + (size_t)w * h, w, cdf97);
		f = fopen(argv[9], "wb");
		fwrite(out.data(), 2, out.size(), f);
		fclose(f);
		printf("ok nlev=%d arena=%zu\n", nlev, Wavelet.arena_bytes());
	} catch (const std::exception &e) {
		fprintf(stderr, "%s\n", e.what());
		return 1;
	}
	return 0;
}
