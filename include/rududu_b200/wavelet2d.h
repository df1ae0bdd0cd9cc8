// rududu_b200/wavelet2d.h -- the reference's C++ class API for the transform+quant path,
// re-created on top of the C ABI (ric_b200.h).  Header-only; link with -lrududu_b200.
//
// What it mirrors (names, argument meaning, ownership):
//   enum trans / cmode                     src/lib/utils.h:27-28
//   enum band_t, ALIGN                     src/lib/band.h:33-35
//   class CBand (public fields)            src/lib/band.h:37-161
//   class CBandCodec::buildTree            src/lib/bandcodec.h:42   (per band, and CWavelet2D::QuantBands for a whole plane)
//   CBand::Init / TSUQ<C> / TSUQi<C>       src/lib/band.h:61-107
//   class CWavelet2D                       src/lib/wavelet2d.h:27-88
//        CWavelet2D(x, y, level, level_chg, Align), SetWeight, Transform<short>, TransformI<short>,
//        CodeBand, DecodeBand, TSUQ, TSUQi, public DBand/HBand/VBand/LBand, pLow/pHigh chain
//   class CMuxCodec (the constructors and endCoding that ric.cpp uses)   src/lib/muxcodec.h:60-139
// Differences a maintainer must know (INTEGRATION.md has the full list):
//   * CodeBand = {quantiser half} + {entropy half}.  The quantiser half (buildTree x3 + LL TSUQ,
//     wavelet2d.cpp:110-126) is QuantBands(Quant, lambda) here and runs on the GPU; the entropy half
//     (CBandCodec::pred/tree, wavelet2d.cpp:119-159) is this library's host entropy stage
//     (ric_mux_code_plane) reading the same pBand buffers.  CodeBand() does both, so the call sequence of
//     CompressImage / DecompressImage (ric.cpp:157-176, 203-225) compiles against these classes as it is.
//   * Transform() leaves the caller's plane untouched (the reference destroys it).
//   * Band buffers of all levels live in ONE pinned host arena owned by the top-level object; every
//     CBand::pBand points into it with the reference's DimXAlign stride and 32-byte alignment.
//   * Errors: the reference has none; here a failing C-ABI call throws std::runtime_error with
//     ric_last_error() (never across the C ABI itself).
// The namespace is rududu_b200 so that the shim can be linked next to the reference library in
// tests; a drop-in build adds `namespace rududu = rududu_b200;`.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#include "../ric_b200.h"

namespace rududu_b200 {

typedef enum cmode { encode, decode } cmode;
typedef enum trans { cdf97 = 0, cdf53 = 1, haar = 2 } trans;
typedef enum band_t { sshort, sint } band_t;
#ifndef ALIGN
#define ALIGN 32
#endif

class CBand {
public:
	unsigned int DimX = 0, DimY = 0, DimXAlign = 0, BandSize = 0;
	int Max = 0, Min = 0;
	unsigned int Dist = 0, Count = 0;
	float Weight = 1.f;
	CBand *pParent = 0, *pChild = 0, *pNeighbor[3] = {0, 0, 0};
	char *pBand = 0;  // points into the owning CWavelet2D's pinned arena
	band_t type = sshort;
	int device = 0;  // (not in the reference) the GPU the per-band methods run on

	CBand() {}
	~CBand() { release(); }
	CBand(const CBand &) = delete;
	CBand &operator=(const CBand &) = delete;

	// band.cpp:51-65: a row-padded, 32-byte aligned buffer of its own (bands of a CWavelet2D live in its arena instead)
	void Init(band_t t = sshort, unsigned int x = 0, unsigned int y = 0, int Align = ALIGN)
	{
		release();
		type = t;
		DimX = x; DimY = y;
		const unsigned sz = t == sint ? 4 : 2;
		DimXAlign = ((x * sz + Align - 1) & -Align) / sz;
		BandSize = DimXAlign * y;
		if (BandSize) {
			void *p = 0;
			if (ric_host_alloc(&p, (size_t)BandSize * sz) < 0) fail();
			pBand = (char *)p;
			own_ = true;
		}
	}

	// band.h:65-92: dead-zone quantiser on pBand (wherever the caller left it), sets Count / Min / Max
	template <class C>
	unsigned int TSUQ(int Quant, float Thres)
	{
		check_type<C>();
		const ric_band_buf b = buf();
		unsigned n = 0;
		if (ric_buf_tsuq(device, &b, Quant, Thres, &n, &Min, &Max) < 0) fail();
		Count = n;
		return n;
	}
	// band.h:94-107
	template <class C>
	void TSUQi(C Quant)
	{
		check_type<C>();
		const ric_band_buf b = buf();
		if (ric_buf_tsuqi(device, &b, (int)Quant) < 0) fail();
	}

	ric_band_buf buf(unsigned char *flags = 0) const
	{
		ric_band_buf b;
		b.data = pBand; b.dimx = (int)DimX; b.dimy = (int)DimY; b.stride = (int)DimXAlign; b.is_int = type == sint;
		b.weight = Weight; b.flags = flags;
		return b;
	}
	void attach(char *p) { release(); pBand = p; }  // (CWavelet2D: the band lives in the plane's arena)

protected:
	template <class C>
	void check_type() const
	{
		if ((sizeof(C) == 4) != (type == sint)) throw std::runtime_error("rududu_b200: sample type does not match the band's type");
	}
	static void fail() { throw std::runtime_error(std::string("rududu_b200: ") + ric_last_error()); }
	void release()
	{
		if (own_ && pBand) ric_host_free(pBand);
		pBand = 0;
		own_ = false;
	}
	bool own_ = false;

public:
	// Host-side helpers of the reference class (band.h:114-159, band.cpp:162-167): statistics and debugging
	// aids over the band buffer, not part of the hot path.
	template <class C>
	void Mean(float &Mean, float &Var) const
	{
		long long sum = 0, sq = 0;
		const C *p = (const C *)pBand;
		for (unsigned j = 0; j < DimY; j++, p += DimXAlign)
			for (unsigned i = 0; i < DimX; i++) { sum += p[i]; sq += (long long)(p[i] * p[i]); }
		const float n = (float)(DimX * DimY);
		Mean = (float)sum * Weight / n;
		Var = (float)(sq - sum * sum) * Weight * Weight / (n * n);
	}
	template <class C>
	void Add(C val)
	{
		C *p = (C *)pBand;
		for (unsigned i = 0; i < BandSize; i++) p[i] = (C)(p[i] + val);
	}
	void Clear(bool recurse = false)
	{
		const size_t bytes = (size_t)BandSize * (type == sint ? 4 : 2);
		for (size_t i = 0; i < bytes; i++) pBand[i] = 0;
		if (recurse && pParent) pParent->Clear(true);
	}
	template <class C, class T>
	void GetBand(T *pOut) const  // samples re-centred on the middle of T's range, rows packed
	{
		const C *p = (const C *)pBand;
		const int mid = 1 << (sizeof(T) * 8 - 1);
		for (unsigned j = 0; j < DimY; j++, p += DimXAlign, pOut += DimX)
			for (unsigned i = 0; i < DimX; i++) pOut[i] = (T)(p[i] + mid);
	}
};

// CBandCodec: the quantiser half (buildTree, bandcodec.h:42, bandcodec.cpp:239-322).  The entropy half (pred / tree)
// is this library's host entropy stage, reached through CWavelet2D::CodeBand / DecodeBand.
class CBandCodec : public CBand {
public:
	CBandCodec() {}
	~CBandCodec() { delete[] pRD; }

	// Quantises this band and then, as the reference recurses through pParent, every coarser band of the chain.
	// high_band: this band is the finest of its orientation (no children); otherwise its child's pRD -- left by an
	// earlier buildTree on the child -- is added, as in the reference.  Works on pBand in host memory.
	template <bool high_band, class C>
	void buildTree(const C Quant, const int lambda)
	{
		check_type<C>();
		ric_band_buf chain[RIC_MAX_LEVELS];
		int n = 0;
		for (CBandCodec *b = this; b && n < RIC_MAX_LEVELS; b = (CBandCodec *)b->pParent) {
			if (!b->pRD) b->pRD = new unsigned char[((b->DimX + 3) / 4) * ((b->DimY + 3) / 4)];  // bandcodec.cpp:252-253
			chain[n++] = b->buf(b->pRD);
		}
		const CBandCodec *ch = (const CBandCodec *)pChild;
		if (!high_band && !(ch && ch->pRD)) throw std::runtime_error("rududu_b200: buildTree<false> needs the child's pRD (run buildTree on it first)");
		if (ric_buf_build_tree(device, chain, n, high_band, high_band ? 0 : ch->pRD, high_band ? 0 : (int)ch->DimX, (int)Quant, lambda) < 0) fail();
	}

	unsigned char *pRD = 0;  // one byte per 4x4 block: the reference's pRD != 0 (all it ever asks of pRD)
};

// CMuxCodec as ric.cpp uses it: CMuxCodec(pStream, firstWord) to write, CMuxCodec(pStream) to read,
// endCoding() (muxcodec.cpp:25-64,92-113).  The reference knows no buffer bounds; pass `capacity` when
// you have it and the writer reports overflow instead of running past the end.
class CMuxCodec {
public:
	CMuxCodec(unsigned char *pStream, unsigned short firstWord, size_t capacity = 0) : base_(pStream), mux_(0)
	{
		if (ric_mux_encoder(&mux_, pStream, bound(pStream, capacity), firstWord) < 0) fail();
	}
	explicit CMuxCodec(unsigned char *pStream) : base_(pStream), mux_(0)
	{
		if (ric_mux_decoder(&mux_, pStream, bound(pStream, 0)) < 0) fail();
	}
	~CMuxCodec() { ric_mux_destroy(mux_); }
	CMuxCodec(const CMuxCodec &) = delete;
	CMuxCodec &operator=(const CMuxCodec &) = delete;

	unsigned char *endCoding()
	{
		size_t end = 0;
		if (ric_mux_finish(mux_, &end) < 0) fail();
		return base_ + end;
	}
	ric_mux *handle() const { return mux_; }

private:
	static size_t bound(const unsigned char *p, size_t n)
	{
		const size_t room = (size_t)(UINTPTR_MAX - (uintptr_t)p) - 1;
		return n && n < room ? n : room;
	}
	static void fail() { throw std::runtime_error(std::string("rududu_b200: ") + ric_last_error()); }
	unsigned char *base_;
	ric_mux *mux_;
};

class CWavelet2D {
public:
	CWavelet2D(int x, int y, int level, int level_chg = 0, int Align = ALIGN, int device = 0)
	    : DimX(x), DimY(y), levels_(level), level_chg_(level_chg), align_(Align), device_(device)
	{
		ctx_[0] = ctx_[1] = ctx_[2] = 0;
		arena_ = 0;
		try {
			ric_ctx *c = ctx(cdf97);
			ric_info inf;
			check(ric_get_info(c, &inf));
			nlev_ = inf.nlev;
			arena_bytes_ = inf.arena_bytes;
			void *p = 0;
			check(ric_host_alloc(&p, arena_bytes_));
			arena_ = (char *)p;
			build_chain(c);
		} catch (...) {  // a constructor that throws runs no destructor: release what exists
			for (CWavelet2D *w = pLow; w;) { CWavelet2D *n = w->pLow; w->pLow = 0; delete w; w = n; }
			pLow = 0;
			if (arena_) ric_host_free(arena_);
			for (int i = 0; i < 3; i++)
				if (ctx_[i]) ric_destroy(ctx_[i]);
			throw;
		}
	}
	~CWavelet2D()
	{
		for (CWavelet2D *w = pLow; w;) { CWavelet2D *n = w->pLow; w->pLow = 0; delete w; w = n; }
		if (!pHigh) {
			if (arena_) ric_host_free(arena_);
			for (int i = 0; i < 3; i++)
				if (ctx_[i]) ric_destroy(ctx_[i]);
		}
	}

	// CWavelet2D::SetWeight, wavelet2d.cpp:1009-1032 (weights come from the C ABI's band table)
	void SetWeight(trans t, float baseWeight = 1.f)
	{
		trans_ = t;
		ric_ctx *c = ctx(t);
		check(ric_set_base_weight(c, baseWeight));
		int id = 0;
		for (CWavelet2D *w = this; w; w = w->pLow, id += 3) {
			ric_band_info b;
			check(ric_get_band(c, id, &b)); w->DBand.Weight = b.weight;
			check(ric_get_band(c, id + 1, &b)); w->HBand.Weight = b.weight;
			check(ric_get_band(c, id + 2, &b)); w->VBand.Weight = b.weight;
			if (!w->pLow) { check(ric_get_band(c, id + 3, &b)); w->LBand.Weight = b.weight; }
		}
	}

	// wavelet2d.cpp:926-958.  pImage: top-left of a DimX x DimY short plane (left untouched).
	template <class C>
	void Transform(C *pImage, int Stride, trans t)
	{
		static_assert(sizeof(C) == 2, "only Transform<short> exists in the reference (wavelet2d.cpp:958)");
		trans_ = t;
		check(ric_transform(ctx(t), (const int16_t *)pImage, Stride, arena_));
	}

	// wavelet2d.cpp:960-992.  pImage is the ONE-PAST-END pointer of the output plane, as in the
	// reference (ric.cpp:216,220,225 pass data() + W*H*(c+1)).
	template <class C>
	void TransformI(C *pImage, int Stride, trans t)
	{
		static_assert(sizeof(C) == 2, "only TransformI<short> exists in the reference (wavelet2d.cpp:992)");
		trans_ = t;
		check(ric_transform_inv(ctx(t), arena_, (int16_t *)pImage - (size_t)Stride * DimY, Stride));
	}

	// The quantiser half of CodeBand (wavelet2d.cpp:110-126): buildTree on the D/H/V chains and
	// TSUQ(Quant, 0.5) on the LL band, on the coefficients in the band buffers (host edits since Transform()
	// included: the arena is uploaded first).  Afterwards the bands hold exactly what CBandCodec::pred /
	// tree<encode> expect.
	void QuantBands(int Quant, int lambda) { check(ric_quant_host(ctx(trans_), Quant, lambda, arena_)); }

	// wavelet2d.cpp:83-159: quantiser half on the GPU, then the entropy half on the host into pCodec's stream.
	void CodeBand(CMuxCodec *pCodec, int Quant, int lambda)
	{
		QuantBands(Quant, lambda);
		check(ric_mux_code_plane(pCodec->handle(), DimX, DimY, levels_, level_chg_, align_, arena_));
	}

	// wavelet2d.cpp:183-221: the bands receive signed quantised coefficients, ready for TSUQi + TransformI.
	void DecodeBand(CMuxCodec *pCodec)
	{
		check(ric_mux_decode_plane(pCodec->handle(), DimX, DimY, levels_, level_chg_, align_, arena_));
	}

	// wavelet2d.cpp:224-246 / :248-268
	unsigned int TSUQ(int Quant, float Thres)
	{
		unsigned n = 0;
		check(ric_tsuq_host(ctx(trans_), Quant, Thres, arena_, &n));
		return n;
	}
	void TSUQi(int Quant) { check(ric_tsuqi(ctx(trans_), Quant, arena_)); }

	CBandCodec DBand, HBand, VBand, LBand;
	CWavelet2D *pLow = 0, *pHigh = 0;
	int DimX, DimY;

	char *arena() const { return arena_; }
	size_t arena_bytes() const { return arena_bytes_; }

private:
	CWavelet2D(CWavelet2D *high, int x, int y) : pHigh(high), DimX(x), DimY(y) { ctx_[0] = ctx_[1] = ctx_[2] = 0; arena_ = 0; }

	static void check(int rc)
	{
		if (rc < 0) throw std::runtime_error(std::string("rududu_b200: ") + ric_last_error());
	}

	CWavelet2D *top() { CWavelet2D *w = this; while (w->pHigh) w = w->pHigh; return w; }

	ric_ctx *ctx(trans t)
	{
		CWavelet2D *T = top();
		if (!T->ctx_[t]) check(ric_create(&T->ctx_[t], T->device_, T->DimX, T->DimY, 1, T->levels_, T->level_chg_, T->align_, t, 1));
		return T->ctx_[t];
	}

	void fill(CBandCodec &b, ric_ctx *c, int id)
	{
		ric_band_info i;
		check(ric_get_band(c, id, &i));
		b.DimX = i.dimx; b.DimY = i.dimy; b.DimXAlign = i.stride; b.BandSize = i.stride * i.dimy;
		b.Weight = i.weight; b.type = i.is_int ? sint : sshort;
		b.attach(top()->arena_ + i.offset);
		b.device = top()->device_;
	}

	void build_chain(ric_ctx *c)
	{
		CWavelet2D *w = this;
		for (int lv = 0; lv < nlev_; lv++) {
			fill(w->DBand, c, 3 * lv); fill(w->HBand, c, 3 * lv + 1); fill(w->VBand, c, 3 * lv + 2);
			if (w->pHigh) {  // wavelet2d.cpp:54-59
				w->DBand.pChild = &w->pHigh->DBand; w->pHigh->DBand.pParent = &w->DBand;
				w->HBand.pChild = &w->pHigh->HBand; w->pHigh->HBand.pParent = &w->HBand;
				w->VBand.pChild = &w->pHigh->VBand; w->pHigh->VBand.pParent = &w->VBand;
			}
			if (lv == nlev_ - 1) { fill(w->LBand, c, 3 * nlev_); break; }
			w->pLow = new CWavelet2D(w, w->DimX >> 1, w->DimY >> 1);
			w = w->pLow;
		}
	}

	ric_ctx *ctx_[3];
	char *arena_;
	size_t arena_bytes_ = 0;
	int levels_ = 0, level_chg_ = 0, align_ = ALIGN, device_ = 0, nlev_ = 0;
	trans trans_ = cdf97;
};

}  // namespace rududu_b200
