// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// extern "C" wrappers around the UNMODIFIED reference library (compiled from
// /root/reference/src/lib where it lies, see oracle/Makefile) so that python tests and
// bench.py's cpu_baseline / --impl reference leg can drive the reference's own
// implementation of the hot path on raw buffers.  Nothing in here is reference source:
// it only *calls* rududu::CWavelet2D / CBandCodec / CMuxCodec through their public API
// and restates the ~40-line call sequence of the (CImg-dependent, hence unbuildable)
// CLI: CompressImage  src/ric/ric.cpp:123-180, DecompressImage src/ric/ric.cpp:182-251,
// Quants src/ric/ric.cpp:42-49, RGBtoYCoCg/YCoCgtoRGB src/ric/ric.cpp:76-112.
//
// Band index convention used by every harness call ("canonical order"):
//   lev = 0 is the FINEST level (the CWavelet2D the caller owns), lev = nlev-1 the coarsest;
//   band id = 3*lev + {0:D, 1:H, 2:V};  the last id (3*nlev) is the coarsest level's LL band.
#include <wavelet2d.h>
#include <muxcodec.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

using namespace rududu;

namespace {

const int kShift = 4;   // SHIFT   src/ric/ric.cpp:39
const int kCBoost = 8;  // C_Q_BOOST src/ric/ric.cpp:38

short quants(int idx)  // restates Quants(), src/ric/ric.cpp:42-49
{
	static const unsigned short Q[5] = {0x8000, 0x9000, 0xA800, 0xC000, 0xE000};
	if (idx <= 0) return 0;
	idx--;
	int r = 14 - idx / 5;
	return (short)((Q[idx % 5] + (1 << (r - 1))) >> r);
}

struct RefWav {
	CWavelet2D *w;
	int nlev;
	std::vector<CWavelet2D *> lev;  // lev[0] finest
	trans t;
};

CBandCodec *band_of(RefWav *h, int id)
{
	if (id == 3 * h->nlev) return &h->lev[h->nlev - 1]->LBand;
	CWavelet2D *w = h->lev[id / 3];
	switch (id % 3) {
	case 0: return &w->DBand;
	case 1: return &w->HBand;
	default: return &w->VBand;
	}
}

// wavelet2d.cpp:110-126 without the pred() call: the GPU-able half of CodeBand
void quant_half(CWavelet2D *w, int Quant, int lambda)
{
	if (w->DBand.type == sshort) {
		w->DBand.buildTree<true, short>(Quant, lambda);
		w->HBand.buildTree<true, short>(Quant, lambda);
		w->VBand.buildTree<true, short>(Quant, lambda);
	} else {
		w->DBand.buildTree<true, int>(Quant, lambda);
		w->HBand.buildTree<true, int>(Quant, lambda);
		w->VBand.buildTree<true, int>(Quant, lambda);
	}
	CWavelet2D *c = w;
	while (c->pLow) c = c->pLow;
	if (c->LBand.type == sshort)
		c->LBand.TSUQ<short>(Quant, 0.5f);
	else
		c->LBand.TSUQ<int>(Quant, 0.5f);
}

// wavelet2d.cpp:119-159 without the quantisers: the host (entropy) half of CodeBand
void entropy_half(CWavelet2D *w, CMuxCodec *codec)
{
	CWavelet2D *c = w;
	while (c->pLow) c = c->pLow;
	if (c->LBand.type == sshort)
		c->LBand.pred<encode, short>(codec);
	else
		c->LBand.pred<encode, int>(codec);
	for (;;) {
		bool fine = (c->pHigh == 0);
		bool par_int = c->pLow && c->pLow->DBand.type == sint;
#define TREE3(HB, C, P)                                   \
	do {                                                  \
		c->VBand.tree<encode, HB, C, P>(codec);           \
		c->HBand.tree<encode, HB, C, P>(codec);           \
		c->DBand.tree<encode, HB, C, P>(codec);           \
	} while (0)
		if (c->DBand.type == sshort) {
			if (!par_int) { if (fine) TREE3(true, short, short); else TREE3(false, short, short); }
			else          { if (fine) TREE3(true, short, int);   else TREE3(false, short, int); }
		} else {
			if (fine) TREE3(true, int, int); else TREE3(false, int, int);
		}
#undef TREE3
		if (fine) break;
		c = c->pHigh;
	}
}

void rgb_to_ycocg(short *co, short *cg, short *y, size_t n, int shift)  // ric.cpp:76-91
{
	for (size_t i = 0; i < n; i++) {
		co[i] -= y[i];
		y[i] += co[i] >> 1;
		cg[i] -= y[i];
		y[i] += (cg[i] >> 1) - 128;
		if (shift > 0) {
			co[i] <<= shift - 1;
			cg[i] <<= shift - 1;
			y[i] <<= shift;
		}
	}
}

void ycocg_to_rgb(short *co, short *cg, short *y, size_t n, int shift)  // ric.cpp:93-112
{
	for (size_t i = 0; i < n; i++) {
		if (shift > 0) {
			co[i] = (co[i] + (1 << (shift - 2))) >> (shift - 1);
			cg[i] = (cg[i] + (1 << (shift - 2))) >> (shift - 1);
			y[i] = (y[i] + (1 << (shift - 1))) >> shift;
		}
		y[i] -= (cg[i] >> 1) - 128;
		cg[i] += y[i];
		y[i] -= co[i] >> 1;
		co[i] += y[i];
		if (shift > 0) {
			co[i] = CLIP(co[i], 0, 255);
			cg[i] = CLIP(cg[i], 0, 255);
			y[i] = CLIP(y[i], 0, 255);
		}
	}
}

void forward_colour(const uint8_t *src, short *img, int w, int h, int ch, int q)
{
	size_t n = (size_t)w * h;
	for (size_t i = 0; i < n * ch; i++) img[i] = src[i];
	if (ch == 3) {
		rgb_to_ycocg(img, img + n, img + 2 * n, n, q == 0 ? 0 : kShift);
	} else {
		if (q == 0) for (size_t i = 0; i < n; i++) img[i] -= 128;           // ric.cpp:144
		else for (size_t i = 0; i < n; i++) img[i] = (img[i] - 128) << kShift;  // ric.cpp:147
	}
}

void inverse_colour(short *img, uint8_t *dst, int w, int h, int ch, int q)
{
	size_t n = (size_t)w * h;
	if (ch == 3) {
		ycocg_to_rgb(img, img + n, img + 2 * n, n, q == 0 ? 0 : kShift);
	} else {
		if (q == 0) for (size_t i = 0; i < n; i++) img[i] += 128;           // ric.cpp:229
		else for (size_t i = 0; i < n; i++) {                                 // ric.cpp:237-240
			img[i] = 128 + ((img[i] + (1 << (kShift - 1))) >> kShift);
			img[i] = CLIP(img[i], 0, 255);
		}
	}
	for (size_t i = 0; i < n * ch; i++) dst[i] = (uint8_t)img[i];
}

}  // namespace

extern "C" {

int ref_quants(int idx) { return quants(idx); }

void *ref_wavelet_new(int w, int h, int levels, int level_chg, int t)
{
	RefWav *r = new RefWav;
	r->w = new CWavelet2D(w, h, levels, level_chg);
	r->t = (trans)t;
	r->w->SetWeight(r->t);
	for (CWavelet2D *c = r->w; c; c = c->pLow) r->lev.push_back(c);
	r->nlev = (int)r->lev.size();
	return r;
}

void ref_wavelet_free(void *hv)
{
	RefWav *h = (RefWav *)hv;
	delete h->w;
	delete h;
}

int ref_levels(void *hv) { return ((RefWav *)hv)->nlev; }

// info[0..4] = DimX, DimY, DimXAlign, type (0 short, 1 int), sample bytes
void ref_band_info(void *hv, int id, int *info, float *weight, void **ptr)
{
	CBandCodec *b = band_of((RefWav *)hv, id);
	info[0] = b->DimX; info[1] = b->DimY; info[2] = b->DimXAlign;
	info[3] = b->type == sint; info[4] = b->type == sint ? 4 : 2;
	if (weight) *weight = b->Weight;
	if (ptr) *ptr = b->pBand;
}

// copy a band out/in as the reference lays it out (DimXAlign stride, padding columns zeroed on read)
void ref_band_get(void *hv, int id, void *dst)
{
	CBandCodec *b = band_of((RefWav *)hv, id);
	int sz = b->type == sint ? 4 : 2;
	memset(dst, 0, (size_t)b->BandSize * sz);
	for (unsigned j = 0; j < b->DimY; j++)
		memcpy((char *)dst + (size_t)j * b->DimXAlign * sz, (char *)b->pBand + (size_t)j * b->DimXAlign * sz,
		       (size_t)b->DimX * sz);
}

void ref_band_set(void *hv, int id, const void *src)
{
	CBandCodec *b = band_of((RefWav *)hv, id);
	int sz = b->type == sint ? 4 : 2;
	memcpy(b->pBand, src, (size_t)b->BandSize * sz);
}

// CWavelet2D::Transform<short>  wavelet2d.cpp:926 (plane is destroyed)
void ref_transform(void *hv, short *plane, int stride)
{
	RefWav *h = (RefWav *)hv;
	h->w->Transform(plane, stride, h->t);
}

// CWavelet2D::TransformI<short> wavelet2d.cpp:960; takes the plane START here, passes the END
void ref_transform_inv(void *hv, short *plane, int stride, int height)
{
	RefWav *h = (RefWav *)hv;
	h->w->TransformI(plane + (size_t)stride * height, stride, h->t);
}

void ref_quant(void *hv, int Quant, int lambda) { quant_half(((RefWav *)hv)->w, Quant, lambda); }
void ref_tsuqi(void *hv, int Quant) { ((RefWav *)hv)->w->TSUQi(Quant); }
unsigned ref_tsuq(void *hv, int Quant, float thres) { return ((RefWav *)hv)->w->TSUQ(Quant, thres); }

// ---- codec objects -------------------------------------------------------------------
void *ref_codec_new_enc(unsigned char *buf) { return new CMuxCodec(buf, 0); }
void *ref_codec_new_enc_word(unsigned char *buf, int first_word) { return new CMuxCodec(buf, (unsigned short)first_word); }
void *ref_codec_new_dec(unsigned char *buf) { return new CMuxCodec(buf); }
long ref_codec_end(void *c, unsigned char *buf)
{
	CMuxCodec *m = (CMuxCodec *)c;
	unsigned char *e = m->endCoding();
	return (long)(e - buf);
}
void ref_codec_free(void *c) { delete (CMuxCodec *)c; }

void ref_codeband(void *hv, void *codec, int Quant, int lambda)
{
	((RefWav *)hv)->w->CodeBand((CMuxCodec *)codec, Quant, lambda);
}
void ref_entropy_encode(void *hv, void *codec) { entropy_half(((RefWav *)hv)->w, (CMuxCodec *)codec); }
void ref_decodeband(void *hv, void *codec) { ((RefWav *)hv)->w->DecodeBand((CMuxCodec *)codec); }

// ---- whole-image restatement of the CLI ------------------------------------------------
// src: planar u8 (R,G,B planes or one gray plane).  out: payload exactly as ric writes it after
// the 9-byte header (pStream+2 .. pEnd).  Returns payload size.
long ref_compress(const uint8_t *src, int w, int h, int ch, int q, int t, int levels, int level_chg,
                  uint8_t *out, long out_cap)
{
	size_t n = (size_t)w * h;
	std::vector<short> img(n * ch + 64);
	forward_colour(src, img.data(), w, h, ch, q);
	std::vector<unsigned char> stream(n * ch * 2 + 4096);
	CMuxCodec codec(stream.data(), 0);
	CWavelet2D wav(w, h, levels, level_chg);
	wav.SetWeight((trans)t);
	int ql = q ? quants(q + kShift * 5) : 0, ll = q ? quants(q + kShift * 5 - 7) : 0;
	int qc = q ? quants(q + kShift * 5 + kCBoost) : 0, lc = q ? quants(q + kShift * 5 - 7 + kCBoost) : 0;
	if (ch == 3) {
		wav.Transform(img.data() + 2 * n, w, (trans)t); wav.CodeBand(&codec, ql, ll);
		wav.Transform(img.data() + n, w, (trans)t);     wav.CodeBand(&codec, qc, lc);
		wav.Transform(img.data(), w, (trans)t);         wav.CodeBand(&codec, qc, lc);
	} else {
		wav.Transform(img.data(), w, (trans)t);         wav.CodeBand(&codec, ql, ll);
	}
	unsigned char *end = codec.endCoding();
	long sz = (long)(end - stream.data()) - 2;
	if (sz > out_cap) return -sz;
	memcpy(out, stream.data() + 2, sz);
	return sz;
}

void ref_decompress(const uint8_t *payload, long size, int w, int h, int ch, int q, int t, int levels,
                    int level_chg, uint8_t *dst)
{
	size_t n = (size_t)w * h;
	std::vector<short> img(n * ch + 64);
	std::vector<unsigned char> stream(n * ch * 2 + 4096 + size);
	memcpy(stream.data() + 2, payload, size);
	CMuxCodec codec(stream.data());
	CWavelet2D wav(w, h, levels, level_chg);
	wav.SetWeight((trans)t);
	wav.DecodeBand(&codec);
	if (q) wav.TSUQi(quants(q + kShift * 5));
	if (ch == 3) {
		wav.TransformI(img.data() + n * 3, w, (trans)t);
		wav.DecodeBand(&codec);
		if (q) wav.TSUQi(quants(q + kShift * 5 + kCBoost));
		wav.TransformI(img.data() + n * 2, w, (trans)t);
		wav.DecodeBand(&codec);
		if (q) wav.TSUQi(quants(q + kShift * 5 + kCBoost));
	}
	wav.TransformI(img.data() + n, w, (trans)t);
	inverse_colour(img.data(), dst, w, h, ch, q);
}

// ---- CPU baseline timing of the hot-path stage only (bench.py --impl reference / cpu_baseline) -----
// Each worker thread owns one CWavelet2D and processes images round-robin:
//   mode 0 (encode stage): colour + Transform + buildTree x3 + LL TSUQ            per plane
//   mode 1 (decode stage): TSUQi + TransformI + inverse colour                     per plane
// Returns seconds for `reps` passes over `n_images` copies of the given image (best-of handled
// by the caller).  The decode stage runs on the bands produced by the encode stage's quantiser
// un-folded to signed form (what DecodeBand would leave there).
static void unfold_bands(RefWav &h)
{
	for (int id = 0; id < 3 * h.nlev; id++) {
		CBandCodec *b = band_of(&h, id);
		for (unsigned j = 0; j < b->DimY; j++)
			for (unsigned i = 0; i < b->DimX; i++) {
				if (b->type == sint) {
					int *p = (int *)b->pBand + (size_t)j * b->DimXAlign + i;
					*p = (*p == -0x8000) ? 0 : u2s_(*p);
				} else {
					short *p = (short *)b->pBand + (size_t)j * b->DimXAlign + i;
					*p = (*p == -0x8000) ? 0 : (short)u2s_((unsigned short)*p);
				}
			}
	}
}

// per-thread stage seconds: mode 0: [0] colour, [1] Transform, [2] buildTree x3 + LL TSUQ;
//                           mode 1: [0] TSUQi,  [1] TransformI, [2] inverse colour
static double g_stage_secs[3];

double ref_bench_stage(const uint8_t *src, int w, int h, int ch, int q, int t, int levels, int level_chg,
                       int n_images, int n_threads, int mode)
{
	size_t n = (size_t)w * h;
	int ql = q ? quants(q + kShift * 5) : 0, ll = q ? quants(q + kShift * 5 - 7) : 0;
	int qc = q ? quants(q + kShift * 5 + kCBoost) : 0, lc = q ? quants(q + kShift * 5 - 7 + kCBoost) : 0;
	std::vector<std::thread> th;
	std::vector<double> secs(n_threads, 0.0);
	std::vector<double> stage(3 * n_threads, 0.0);
	typedef std::chrono::steady_clock clk;
	for (int k = 0; k < n_threads; k++) {
		th.emplace_back([&, k]() {
			std::vector<short> img(n * ch + 64);
			std::vector<uint8_t> out(n * ch);
			std::vector<RefWav> wav(ch);
			std::vector<std::vector<char>> saved(ch);
			double *st = &stage[3 * k];
			for (int c = 0; c < ch; c++) {
				wav[c].w = new CWavelet2D(w, h, levels, level_chg);
				wav[c].t = (trans)t;
				wav[c].w->SetWeight((trans)t);
				for (CWavelet2D *p = wav[c].w; p; p = p->pLow) wav[c].lev.push_back(p);
				wav[c].nlev = (int)wav[c].lev.size();
			}
			auto enc = [&](double *acc) {
				auto a = clk::now();
				forward_colour(src, img.data(), w, h, ch, q);
				auto b = clk::now();
				double tt = 0, tq = 0;
				for (int c = ch - 1; c >= 0; c--) {
					auto c0 = clk::now();
					wav[c].w->Transform(img.data() + c * n, w, (trans)t);
					auto c1 = clk::now();
					bool luma = (ch == 1) || (c == 2);
					quant_half(wav[c].w, luma ? ql : qc, luma ? ll : lc);
					auto c2 = clk::now();
					tt += std::chrono::duration<double>(c1 - c0).count();
					tq += std::chrono::duration<double>(c2 - c1).count();
				}
				if (acc) { acc[0] += std::chrono::duration<double>(b - a).count(); acc[1] += tt; acc[2] += tq; }
			};
			double acc = 0;
			if (mode == 1) {  // prepare quantised signed bands once and snapshot them
				enc(nullptr);
				for (int c = 0; c < ch; c++) {
					unfold_bands(wav[c]);
					for (int id = 0; id <= 3 * wav[c].nlev; id++) {
						CBandCodec *b = band_of(&wav[c], id);
						size_t bytes = (size_t)b->BandSize * (b->type == sint ? 4 : 2);
						size_t o = saved[c].size();
						saved[c].resize(o + bytes);
						memcpy(saved[c].data() + o, b->pBand, bytes);
					}
				}
			}
			for (int i = k; i < n_images; i += n_threads) {
				if (mode == 1) {  // restore (untimed): TSUQi works in place
					for (int c = 0; c < ch; c++) {
						size_t o = 0;
						for (int id = 0; id <= 3 * wav[c].nlev; id++) {
							CBandCodec *b = band_of(&wav[c], id);
							size_t bytes = (size_t)b->BandSize * (b->type == sint ? 4 : 2);
							memcpy(b->pBand, saved[c].data() + o, bytes);
							o += bytes;
						}
					}
				}
				auto t0 = clk::now();
				if (mode == 0) {
					enc(st);
				} else {
					double td = 0, ti = 0;
					for (int c = ch - 1; c >= 0; c--) {
						bool luma = (ch == 1) || (c == 2);
						auto c0 = clk::now();
						if (q) wav[c].w->TSUQi(luma ? ql : qc);
						auto c1 = clk::now();
						wav[c].w->TransformI(img.data() + (c + 1) * n, w, (trans)t);
						auto c2 = clk::now();
						td += std::chrono::duration<double>(c1 - c0).count();
						ti += std::chrono::duration<double>(c2 - c1).count();
					}
					auto c3 = clk::now();
					inverse_colour(img.data(), out.data(), w, h, ch, q);
					auto c4 = clk::now();
					st[0] += td; st[1] += ti; st[2] += std::chrono::duration<double>(c4 - c3).count();
				}
				auto t1 = clk::now();
				acc += std::chrono::duration<double>(t1 - t0).count();
			}
			secs[k] = acc;
			for (int c = 0; c < ch; c++) delete wav[c].w;
		});
	}
	for (auto &x : th) x.join();
	double mx = 0;
	int kmax = 0;
	for (int k = 0; k < n_threads; k++) if (secs[k] > mx) { mx = secs[k]; kmax = k; }
	for (int i = 0; i < 3; i++) g_stage_secs[i] = stage[3 * kmax + i];
	return mx;  // busiest thread's compute time == stage makespan
}

// stage split of the busiest thread of the last ref_bench_stage call (see g_stage_secs)
void ref_bench_stage_split(double *out3) { for (int i = 0; i < 3; i++) out3[i] = g_stage_secs[i]; }

// CPU baseline of the WHOLE codec (bench.py "codec" figures): the CompressImage / DecompressImage
// restatements above, one image per thread round-robin.  mode 0: compress n_images copies of src;
// mode 1: decompress n_images copies of the payload of src.  Returns the busiest thread's seconds.
double ref_bench_codec(const uint8_t *src, int w, int h, int ch, int q, int t, int levels, int level_chg,
                       int n_images, int n_threads, int mode)
{
	const size_t n = (size_t)w * h * ch;
	std::vector<uint8_t> payload(n * 2 + 4096);
	const long psz = ref_compress(src, w, h, ch, q, t, levels, level_chg, payload.data(), (long)payload.size());
	if (psz < 0) return -1.0;
	std::vector<std::thread> th;
	std::vector<double> secs(n_threads, 0.0);
	typedef std::chrono::steady_clock clk;
	for (int k = 0; k < n_threads; k++)
		th.emplace_back([&, k]() {
			std::vector<uint8_t> out(n * 2 + 4096);
			auto t0 = clk::now();
			for (int i = k; i < n_images; i += n_threads) {
				if (mode == 0) ref_compress(src, w, h, ch, q, t, levels, level_chg, out.data(), (long)out.size());
				else ref_decompress(payload.data(), psz, w, h, ch, q, t, levels, level_chg, out.data());
			}
			secs[k] = std::chrono::duration<double>(clk::now() - t0).count();
		});
	for (auto &x : th) x.join();
	double mx = 0;
	for (double v : secs) mx = v > mx ? v : mx;
	return mx;
}

}  // extern "C"
