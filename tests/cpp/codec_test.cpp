// The body of CompressImage / DecompressImage (src/ric/ric.cpp:157-176, 203-225) written against the shim
// classes exactly as the reference writes it against its own: one CMuxCodec shared by the planes,
// Transform + CodeBand per plane (Y, Cg, Co), endCoding; then DecodeBand + TSUQi + TransformI per plane.
// usage: codec_test W H CH q planes.s16 payload_out.bin planes_out.s16
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "rududu_b200/wavelet2d.h"

using namespace rududu_b200;

#define WAV_LEVELS 5
#define SHIFT 4
#define C_Q_BOOST 8

int main(int argc, char **argv)
{
	if (argc != 8) return 2;
	const int w = atoi(argv[1]), h = atoi(argv[2]), ch = atoi(argv[3]), Quant = atoi(argv[4]);
	const size_t n = (size_t)w * h;
	std::vector<short> img(n * ch), out(n * ch);
	FILE *f = fopen(argv[5], "rb");
	if (!f || fread(img.data(), 2, n * ch, f) != n * ch) return 2;
	fclose(f);
	try {
		const trans Trans = cdf97;
		std::vector<unsigned char> stream(n * ch + 4096);
		unsigned char *pEnd;
		{
			CMuxCodec Codec(stream.data(), 0);
			CWavelet2D Wavelet(w, h, WAV_LEVELS, WAV_LEVELS - 4);
			Wavelet.SetWeight(Trans);
			const int ql = ric_quants(Quant + SHIFT * 5), ll = ric_quants(Quant + SHIFT * 5 - 7);
			const int qc = ric_quants(Quant + SHIFT * 5 + C_Q_BOOST), lc = ric_quants(Quant + SHIFT * 5 - 7 + C_Q_BOOST);
			if (ch == 3) {
				Wavelet.Transform(img.data() + 2 * n, w, Trans);
				Wavelet.CodeBand(&Codec, ql, ll);
				Wavelet.Transform(img.data() + n, w, Trans);
				Wavelet.CodeBand(&Codec, qc, lc);
				Wavelet.Transform(img.data(), w, Trans);
				Wavelet.CodeBand(&Codec, qc, lc);
			} else {
				Wavelet.Transform(img.data(), w, Trans);
				Wavelet.CodeBand(&Codec, ql, ll);
			}
			pEnd = Codec.endCoding();
		}
		f = fopen(argv[6], "wb");
		fwrite(stream.data() + 2, 1, pEnd - stream.data() - 2, f);  // ric.cpp:176
		fclose(f);
		{
			CMuxCodec Codec(stream.data());
			CWavelet2D Wavelet(w, h, WAV_LEVELS, WAV_LEVELS - 4);
			Wavelet.SetWeight(Trans);
			Wavelet.DecodeBand(&Codec);
			Wavelet.TSUQi(ric_quants(Quant + SHIFT * 5));
			if (ch == 3) {
				Wavelet.TransformI(out.data() + n * 3, w, Trans);
				Wavelet.DecodeBand(&Codec);
				Wavelet.TSUQi(ric_quants(Quant + SHIFT * 5 + C_Q_BOOST));
				Wavelet.TransformI(out.data() + n * 2, w, Trans);
				Wavelet.DecodeBand(&Codec);
				Wavelet.TSUQi(ric_quants(Quant + SHIFT * 5 + C_Q_BOOST));
			}
			Wavelet.TransformI(out.data() + n, w, Trans);
			// the reference class's host helpers exist on the shim's bands too
			std::vector<unsigned char> vis((size_t)Wavelet.HBand.DimX * Wavelet.HBand.DimY);
			Wavelet.HBand.GetBand<short, unsigned char>(vis.data());
			float mean, var;
			Wavelet.HBand.Mean<short>(mean, var);
			Wavelet.HBand.Add<short>(1);
			Wavelet.HBand.Clear();
			if (Wavelet.HBand.pBand[0] != 0) return 3;
		}
		f = fopen(argv[7], "wb");
		fwrite(out.data(), 2, out.size(), f);
		fclose(f);
		printf("ok payload=%ld\n", (long)(pEnd - stream.data() - 2));
	} catch (const std::exception &e) {
		fprintf(stderr, "%s\n", e.what());
		return 1;
	}
	return 0;
}
