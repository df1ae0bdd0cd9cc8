// ric_entropy_core.h -- the entropy stage's stream objects, adaptive models and band walkers, written so that
// the same source compiles for the host (ric_entropy.cpp, one image per thread) and for the device
// (ric_entropy_gpu.cu, one image per warp).  See ric_entropy.h for the reference map.
//
// Structure (ours): one MuxWriter / MuxReader per image carries the interleaved range-coder bytes and raw bit
// fields; small adaptive models (GeomModel, BitModel) are created per band; the band walkers are written once
// over a Port (writer or reader) so that both directions share the traversal.
#pragma once
#include <stdint.h>
#include <string.h>
#if defined(__SSE2__) && !defined(__CUDA_ARCH__)
#include <emmintrin.h>
#endif

#include "ric_host.h"
#include "ric_huff_data.h"

#if defined(__CUDACC__)
#define RIC_HD __host__ __device__
#else
#define RIC_HD
#endif

namespace ric {
namespace ent {

RIC_HD inline int count_lz(uint32_t v)
{
#ifdef __CUDA_ARCH__
	return __clz((int)v);
#else
	return __builtin_clz(v);
#endif
}
RIC_HD inline int count_tz(uint32_t v)
{
#ifdef __CUDA_ARCH__
	return __ffs((int)v) - 1;
#else
	return __builtin_ctz(v);
#endif
}
RIC_HD inline int count_ones(uint32_t v)
{
#ifdef __CUDA_ARCH__
	return __popc(v);
#else
	return __builtin_popcount(v);
#endif
}


enum : uint32_t { kProbBits = 12, kProbOne = 1u << kProbBits, kProbHalf = kProbOne >> 1, kMinRange = 1u << 12 };
enum : int { kMarker = -0x8000 };  // insignificant-block marker, bandcodec.cpp:112-113

RIC_HD inline int bit_length(uint32_t v) { return v ? 32 - count_lz(v) : 0; }  // utils.h:132-140
RIC_HD inline int fold_signed(int s) { int u = -(2 * s + 1); return u ^ (u >> 31); }  // utils.h:80-85 (0,-1,1,-2.. -> 0,1,2,3..)
RIC_HD inline int unfold_signed(int u) { return (u >> 1) ^ -(u & 1); }                // utils.h:87-90
RIC_HD inline int unfold_sign_lsb(int u) { const int m = -(u & 1); return ((u >> 1) + m) ^ m; }  // utils.h:98-102

// ---------------------------------------------------------------------------------------------
// Static code tables, all derived at start-up.
// ---------------------------------------------------------------------------------------------
struct CountCode {          // one canonical prefix code over k
	uint16_t code[17];      // encoder: code word per symbol
	uint8_t len[17];
	uint16_t floor16[17];   // decoder: per length L (1..16), left-aligned lowest code of that length ...
	uint8_t first_rank[17]; // ... rank of the HIGHEST code of that length, and
	uint8_t count[17];      // ... number of codes of that length
	uint8_t by_rank[17];
	uint8_t quick_len[256], quick_sym[256];  // decoder: codes of up to 8 bits, indexed by the next 8 stream bits
};

struct Tables {
	uint32_t taboo_n[32], taboo_sum[32];  // muxcodec.cpp:120-137, taboo length 2
	uint16_t choose[8][17];               // choose[r][n] = C(n, r+1)
	uint8_t enum_len[17][9];              // [n][k]: bits of the truncated binary code over C(n,k) values
	uint16_t enum_short[17][9];           // [n][k]: how many values get the (len-1)-bit form
	uint8_t edge_ctx[17][16];             // geometric-model context of an edge block: [samples][k-1]
	CountCode low[17], fine[16];
	uint16_t geo_bound[11];               // geometric model: probability bounds of the adaptation rates (geomcodec.cpp:45-47)
	uint16_t bit_lim[11];                 // binary model: same for its rates (bitcodec.cpp:40-42)
	uint8_t ll_init[16], band_init[16];   // initial geometric-model states (bandcodec.cpp:66-67, 486)
	uint16_t kmean_init[16];              // initial running mean of k per parent class, 6.10 fixed point (bandcodec.cpp:487-489)

	Tables()
	{
		static const uint16_t gb[11] = {1512, 2584, 3351, 3725, 3911, 4004, 4050, 4073, 4084, 4090, 4093};
		static const uint16_t bl[11] = {2584, 1512, 745, 371, 185, 92, 46, 23, 12, 6, 3};
		static const uint8_t li[16] = {9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 15};
		static const uint8_t bi[16] = {5, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 10, 10, 10, 11};
		static const uint8_t km[16] = {2, 3, 4, 5, 8, 11, 13, 14, 15, 15, 15, 15, 15, 15, 15, 15};
		for (int i = 0; i < 11; i++) { geo_bound[i] = gb[i]; bit_lim[i] = bl[i]; }
		for (int i = 0; i < 16; i++) { ll_init[i] = li[i]; band_init[i] = bi[i]; kmean_init[i] = (uint16_t)(km[i] << 10); }

		taboo_n[0] = taboo_n[1] = 1;
		for (int i = 2; i < 32; i++) taboo_n[i] = taboo_n[i - 2] + taboo_n[i - 1];
		taboo_sum[0] = 1;
		for (int i = 1; i < 32; i++) taboo_sum[i] = taboo_sum[i - 1] + taboo_n[i];

		uint32_t binom[17][17];
		memset(binom, 0, sizeof binom);
		for (int n = 0; n <= 16; n++) {
			binom[n][0] = 1;
			for (int k = 1; k <= n; k++) binom[n][k] = binom[n - 1][k - 1] + (k <= n - 1 ? binom[n - 1][k] : 0);
		}
		memset(choose, 0, sizeof choose);
		for (int r = 0; r < 8; r++)
			for (int n = 0; n < 16; n++) choose[r][n] = (uint16_t)binom[n][r + 1];
		memset(enum_len, 0, sizeof enum_len);
		memset(enum_short, 0, sizeof enum_short);
		for (int n = 1; n <= 16; n++)
			for (int k = 1; k <= 8 && k <= n; k++) {
				const uint32_t c = binom[n][k];
				const int L = bit_length(c - 1);  // ceil(log2 c)
				enum_len[n][k] = (uint8_t)L;
				enum_short[n][k] = (uint16_t)((1u << L) - c);
			}

		// Edge blocks: the context for a block of n samples with k set is round(16k/n)-1, tabulated for
		// n in {1,2,3,4,6,8,9,12,16}.  The reference selects the row through a 16-entry map that was laid
		// out for n-1 but is indexed with n (bandcodec.cpp:423-424,452,475): reproduce that selection.
		static const int sizes[9] = {1, 2, 3, 4, 6, 8, 9, 12, 16};
		int row_of[17];
		for (int i = 0; i < 17; i++) row_of[i] = 0;
		for (int r = 0; r < 9; r++) row_of[sizes[r] - 1] = r;
		memset(edge_ctx, 0, sizeof edge_ctx);
		for (int n = 1; n < 16; n++) {
			const int m = sizes[row_of[n]];
			for (int k = 1; k <= m && k <= 16; k++) edge_ctx[n][k - 1] = (uint8_t)((k * 32 + m) / (2 * m) - 1);
		}

		for (int t = 0; t < 17; t++) build(low[t], kCountCodeLow[t], 17);
		for (int t = 0; t < 16; t++) build(fine[t], kCountCodeFine[t], 16);
	}

	static void build(CountCode &c, const uint8_t *desc, int nsym)
	{
		memset(&c, 0, sizeof c);
		const uint8_t *order = desc + 16;
		uint32_t top = 1u << 16;
		int rank = 0;
		for (int L = 1; L <= 16; L++) {
			c.first_rank[L] = (uint8_t)rank;
			c.count[L] = desc[L - 1];
			for (int i = 0; i < desc[L - 1]; i++, rank++) {
				top -= 1u << (16 - L);
				const int s = order[rank];
				c.code[s] = (uint16_t)(top >> (16 - L));
				c.len[s] = (uint8_t)L;
				c.by_rank[rank] = (uint8_t)s;
			}
			c.floor16[L] = (uint16_t)top;
		}
		for (int s = 0; s < nsym; s++) {
			if (c.len[s] > 8) continue;
			const unsigned first = (unsigned)c.code[s] << (8 - c.len[s]);
			for (unsigned i = 0; i < (1u << (8 - c.len[s])); i++) { c.quick_len[first + i] = c.len[s]; c.quick_sym[first + i] = (uint8_t)s; }
		}
	}
};


// ---------------------------------------------------------------------------------------------
// Stream multiplexer.  Range-coder bytes and raw bit fields share one byte stream; the range coder's
// bytes trail by four positions (carry-less coder with deferred slots) and a partially filled bit byte
// is parked in a slot claimed the moment the range coder next emits, which is exactly where the
// reader will look for it.
// ---------------------------------------------------------------------------------------------
class MuxWriter {
public:
	RIC_HD MuxWriter(uint8_t *out, size_t cap, unsigned first_word = 0) : wr_(out + 2), lim_(out + cap), low_((uint32_t)first_word << 16)
	{
		// The first two range-coder bytes carry the coder's 16-bit start word; a .ric file does not store them.
		slot_[0] = &lead_[0]; slot_[1] = &lead_[1]; slot_[2] = out; slot_[3] = out + 1;
		if (cap < 2) { slot_[2] = slot_[3] = &sink_; wr_ = lim_ = out; overflow_ = true; }
	}

	RIC_HD inline void bin(uint32_t p, int bit)  // muxcodec.h:165-172
	{
		if (range_ <= kMinRange) renorm();
		const uint32_t cut = (range_ * p) >> kProbBits;
		if (bit) { low_ += cut; range_ -= cut; } else range_ = cut;
	}

	RIC_HD inline void bits(uint32_t v, unsigned n)  // muxcodec.h:224-230
	{
		if (nacc_ + n > 32) drain();
		acc_ = (acc_ << n) | v;
		nacc_ += n;
	}

	RIC_HD uint8_t *finish()  // muxcodec.cpp:92-113
	{
		park(true);
		if (range_ <= kMinRange) renorm();
		const uint32_t tail = 0x257 & (kMinRange - 1);
		if ((low_ & (kMinRange - 1)) > tail) low_ += kMinRange;
		low_ = (low_ & ~(uint32_t)(kMinRange - 1)) | tail;
		for (int i = 0; i < 4; i++) *slot_[(head_ + i) & 3] = (uint8_t)(low_ >> (24 - 8 * i));
		return wr_;
	}

	RIC_HD bool overflow() const { return overflow_; }
	RIC_HD void poison() { overflow_ = true; }  // input that no encode stage produces: the call reports failure
	uint8_t lead(int i) const { return lead_[i]; }

private:
	RIC_HD inline uint8_t *claim()
	{
		if (wr_ < lim_) return wr_++;
		overflow_ = true;
		return &sink_;
	}

	RIC_HD void renorm()  // muxcodec.cpp:67-78
	{
		park(false);
		do {
			*slot_[head_] = (uint8_t)(low_ >> 24);
			if (((low_ + range_ - 1) ^ low_) >= 0x01000000u) range_ = (0u - low_) & (kMinRange - 1);
			slot_[head_] = claim();
			head_ = (head_ + 1) & 3;
			range_ <<= 8;
			low_ <<= 8;
		} while (range_ <= kMinRange);
	}

	RIC_HD inline void put_byte(uint8_t b)
	{
		if (parked_) { *parked_ = b; parked_ = 0; } else *claim() = b;
	}

	RIC_HD void drain()  // muxcodec.cpp:517-530
	{
		do {
			nacc_ -= 8;
			put_byte((uint8_t)(acc_ >> nacc_));
		} while (nacc_ >= 8);
	}

	RIC_HD void park(bool final)  // muxcodec.cpp:532-552
	{
		if (nacc_ >= 8) drain();
		if (nacc_ == 0) return;
		if (final) { put_byte((uint8_t)(acc_ << (8 - nacc_))); nacc_ = 0; }
		else if (!parked_) parked_ = claim();
	}

	uint8_t *wr_, *lim_;
	uint8_t *slot_[4];
	unsigned head_ = 0;
	uint32_t low_, range_ = kMinRange << 4;
	uint32_t acc_ = 0;
	unsigned nacc_ = 0;
	uint8_t *parked_ = 0;
	uint8_t lead_[2] = {0, 0}, sink_ = 0;
	bool overflow_ = false;
};

class MuxReader {
public:
	RIC_HD MuxReader(const uint8_t *p, size_t size) : rd_(p), end_(p + size)
	{
		low_ = code_ = (uint32_t)(next() << 8);
		const uint32_t b = next();
		low_ |= b; code_ |= b;
	}

	RIC_HD inline int bin(uint32_t p)  // muxcodec.h:203-211
	{
		if (range_ <= kMinRange) renorm();
		const uint32_t cut = (range_ * p) >> kProbBits;
		if (low_ < cut) { range_ = cut; return 0; }
		low_ -= cut; range_ -= cut;
		return 1;
	}

	RIC_HD inline uint32_t bits(unsigned n)  // muxcodec.h:232-238
	{
		if (nacc_ < n) fill(n);
		nacc_ -= n;
		return (acc_ >> nacc_) & ((1u << n) - 1);
	}

	// Prefix code over k: 16-bit look-ahead across the bit accumulator and the next two stream bytes,
	// then give back whole bytes that were not needed (muxcodec.h:241-252).
	RIC_HD inline int count_symbol(const CountCode &c)
	{
		const uint32_t b0 = rd_ < end_ ? rd_[0] : 0, b1 = rd_ + 1 < end_ ? rd_[1] : 0;
		const uint32_t win = (((acc_ << 16) | (b0 << 8) | b1) >> nacc_) & 0xFFFF;
		unsigned L = c.quick_len[win >> 8];
		int sym;
		if (L) sym = c.quick_sym[win >> 8];
		else {
			L = 9;
			while (win < c.floor16[L] || !c.count[L]) L++;
			sym = c.by_rank[c.first_rank[L] + (int)(((uint32_t)c.floor16[L] + ((uint32_t)c.count[L] << (16 - L)) - 1 - win) >> (16 - L))];
		}
		if (nacc_ >= L) nacc_ -= L;
		else {
			const unsigned need = L - nacc_, nbytes = (need + 7) >> 3;
			for (unsigned i = 0; i < nbytes; i++) acc_ = next();
			nacc_ = nbytes * 8 - need;
		}
		return sym;
	}

	RIC_HD uint32_t taboo(const Tables &T)  // muxcodec.cpp:235-276, taboo length 2
	{
		if (nacc_ < 2) fill(2);
		unsigned l = 2;
		uint32_t t = 3u << (nacc_ - 2);
		while ((~acc_ & t) != t) {
			l++;
			// a valid code is at most ~28 bits long (values are folded 16/32-bit samples); a stream
			// without two consecutive zero bits (0xFF.., 0x55..) would otherwise run the length past the
			// accumulator and index the 32-entry tables far out of bounds
			if (l > 32) { past_ += 1000; nacc_ = 0; return 0; }
			if (l > nacc_) { fill(l); t <<= 8; }
			t >>= 1;
		}
		nacc_ -= l;
		const uint32_t cd = acc_ >> (nacc_ + 3);
		int i = (int)l - 2;
		uint32_t v = 0;
		if (i > 0) { i--; v += T.taboo_sum[i]; }
		while (i > 2) {
			int j = 1;
			while (((cd >> (i - j)) & 1) == 0) j++;
			v += T.taboo_sum[i - j] - T.taboo_sum[i - 2];
			i -= j;
		}
		if (i == 2) v -= 1;
		return v + (cd & ((1u << i) - 1));
	}

	RIC_HD bool overrun() const { return past_ > 8; }

private:
	RIC_HD inline uint32_t next()
	{
		if (rd_ < end_) return *rd_++;
		past_++;
		return 0;
	}

	RIC_HD void renorm()  // muxcodec.cpp:80-90
	{
		do {
			const uint32_t d = code_ - low_;
			if (((d + range_ - 1) ^ d) >= 0x01000000u) range_ = (low_ - code_) & (kMinRange - 1);
			const uint32_t b = next();
			low_ = (low_ << 8) | b;
			code_ = (code_ << 8) | b;
			range_ <<= 8;
			if (range_ == 0) { range_ = kMinRange << 4; past_ += 1000; }  // impossible in a valid stream: flag it, keep moving
		} while (range_ <= kMinRange);
	}

	RIC_HD void fill(unsigned n)  // muxcodec.cpp:570-577
	{
		do {
			nacc_ += 8;
			acc_ = (acc_ << 8) | next();
		} while (nacc_ < n);
	}

	const uint8_t *rd_, *end_;
	uint32_t low_ = 0, code_ = 0, range_ = kMinRange << 4;
	uint32_t acc_ = 0;
	unsigned nacc_ = 0;
	unsigned past_ = 0;
};

// Ports: the band walkers below call these; `value`/`flag` arguments are inputs when writing and
// ignored when reading, the return value is what the stream holds.
struct WritePort {
	static constexpr bool writing = true;
	MuxWriter &m;
	const Tables *T;
	RIC_HD inline int bin(uint32_t p, int bit) { m.bin(p, bit); return bit; }
	RIC_HD inline uint32_t bits(uint32_t v, unsigned n) { m.bits(v, n); return v; }

	RIC_HD void taboo(uint32_t v)  // muxcodec.cpp:198-233, taboo length 2
	{
		const Tables &T = *this->T;
		int i = 0;
		while (T.taboo_sum[i] <= v) i++;
		if (i == 0) { m.bits(0, 2); return; }
		const int l = i;
		i--;
		v -= T.taboo_sum[i];
		uint32_t r = 0;
		while (i > 2) {
			const int k = i - 1;
			uint32_t cnt = T.taboo_n[k];
			int j = 0;
			while (v >= cnt) cnt += T.taboo_n[k + ++j];
			v -= cnt - T.taboo_n[k + j];
			j = 2 - j;
			r = (r << j) | 1;
			i -= j;
		}
		if (i == 2) v++;
		r = ((((r << i) | (v & ((1u << i) - 1))) << 1) | 1) << 2;
		m.bits(r, l + 2);
	}
};

struct ReadPort {
	static constexpr bool writing = false;
	MuxReader &m;
	const Tables *T;
	RIC_HD inline int bin(uint32_t p, int) { return m.bin(p); }
	RIC_HD inline uint32_t bits(uint32_t, unsigned n) { return m.bits(n); }
};

// value in [0, max] with a truncated binary code (muxcodec.cpp:496-515).  The reference reader takes one
// bit for max == 0 where its writer emits none; we follow the writer.
template <class Port>
RIC_HD inline uint32_t truncated(Port &io, uint32_t value, uint32_t max)
{
	const int len = bit_length(max);
	if (len == 0) return 0;
	const uint32_t spare = (1u << len) - max - 1;
	if (Port::writing) {
		if (value < spare) io.bits(value, len - 1); else io.bits(value + spare, len);
		return value;
	}
	uint32_t v = len > 1 ? io.bits(0, len - 1) : 0;
	if (v >= spare) v = ((v << 1) | io.bits(0, 1)) - spare;
	return v;
}

// Which k of n positions are set, as the index of the combination (muxcodec.cpp:341-401).  Bit i of `mask`
// is scan position i (row-major); the reference numbers positions from the other end, hence n-1-i.
template <class Port>
RIC_HD inline uint32_t combination(Port &io, uint32_t mask, unsigned k, unsigned n)
{
	const Tables &T = *io.T;
	const uint32_t all = (1u << n) - 1;
	const bool flip = k > ((n + 1) >> 1);
	if (flip) k = n - k;
	const unsigned len = T.enum_len[n][k];
	const uint32_t nshort = T.enum_short[n][k];
	if (Port::writing) {
		uint32_t m = flip ? mask ^ all : mask, idx = 0;
		for (unsigned r = 0; m; r++) {
			const unsigned i = 31 - (unsigned)count_lz(m);  // last scanned position first
			idx += T.choose[r][n - 1 - i];
			m ^= 1u << i;
		}
		if (idx < nshort) io.bits(idx, len - 1); else io.bits(idx + nshort, len);
		return mask;
	}
	uint32_t idx = io.bits(0, len - 1);
	if (idx >= nshort) idx = ((idx << 1) | io.bits(0, 1)) - nshort;
	uint32_t m = 0;
	int r = (int)k - 1;
	for (int pos = (int)n - 1; r >= 0 && pos >= 0; pos--)
		if (idx >= T.choose[r][pos]) { m |= 1u << (n - 1 - pos); idx -= T.choose[r][pos]; r--; }
	return flip ? m ^ all : m;
}

// Adaptive geometric (Golomb-like) model: unary part through the range coder with one adaptive
// probability per context, k low bits raw; the state index moves the Golomb parameter and the
// adaptation rate (geomcodec.h:40-97).
class GeomModel {
public:
	RIC_HD GeomModel(const uint8_t init[16], const uint16_t *bounds) : bound_(bounds)
	{
		for (int c = 0; c < 16; c++) {
			state_[c] = init[c];
			prob_[c] = init[c] >= 9 ? (uint16_t)kProbHalf : (uint16_t)((bound(init[c] - 1) + bound(init[c])) >> 1);
		}
	}

	template <class Port>
	RIC_HD inline uint32_t code(Port &io, uint32_t sym, unsigned ctx)
	{
		const unsigned st = state_[ctx];
		// (k <= 24: a corrupt stream can walk the state up without bound; valid ones never come near)
		const unsigned k = st > 9 ? (st - 9 < 24 ? st - 9 : 24) : 0, s = st < 9 ? 10 - st : 1, rate = 3 + s;
		const uint32_t p = prob_[ctx];
		uint32_t pr = p, hi = Port::writing ? sym >> k : 0, n = 0;
		if (Port::writing) {
			for (; n < hi; n++) { io.bin(p, 1); pr -= pr >> rate; }
			io.bin(p, 0);
		} else {
			while (io.bin(p, 0) && n < (1u << 16)) { pr -= pr >> rate; n++; }  // the cap only matters for corrupt input
		}
		uint32_t v = n;
		if (k) v = (n << k) | io.bits(sym & ((1u << k) - 1), k);
		pr += (kProbOne - pr) >> rate;
		prob_[ctx] = (uint16_t)pr;
		if ((uint16_t)(pr - bound(s - 1)) > bound(s) - bound(s - 1)) {
			if (pr < bound(s - 1)) { if (state_[ctx] < 40) state_[ctx]++; }
			else if (state_[ctx] > 0) state_[ctx]--;
			if (state_[ctx] >= 9) prob_[ctx] = (uint16_t)kProbHalf;
		}
		return v;
	}

private:
	RIC_HD inline uint32_t bound(int i) const { return bound_[i]; }
	const uint16_t *bound_;
	uint16_t prob_[16];
	uint8_t state_[16];
};

// Adaptive binary model with a most-probable-symbol flag and a state-dependent rate (bitcodec.h:28-93).
class BitModel {
public:
	RIC_HD explicit BitModel(const uint16_t *lim) : lim_(lim)
	{
		for (int c = 0; c < 16; c++) { prob_[c] = (uint16_t)kProbHalf; mps_[c] = 0; slow_[c] = 0; }
	}

	template <class Port>
	RIC_HD inline int code(Port &io, int sym, unsigned ctx)
	{
		const uint16_t *lim = lim_;
		const unsigned sh = slow_[ctx];
		const int miss = io.bin(prob_[ctx], (sym ^ mps_[ctx]) ^ 1) ^ 1;
		const uint16_t p = (uint16_t)(prob_[ctx] + (miss << (9 - sh)) - (prob_[ctx] >> (3 + sh)));
		prob_[ctx] = p;
		const int out = miss ^ mps_[ctx];
		if ((uint16_t)(p - lim[sh + 1]) > lim[sh] - lim[sh + 1]) {
			if (p > lim[sh]) {
				if (sh == 0) { mps_[ctx] ^= 1; prob_[ctx] = (uint16_t)(kProbOne - p); slow_[ctx] = 1; }
				else slow_[ctx]--;
			} else if (sh < 9) slow_[ctx]++;
		}
		return out;
	}

private:
	const uint16_t *lim_;
	uint16_t prob_[16];
	uint8_t mps_[16], slow_[16];
};

// ---------------------------------------------------------------------------------------------
// Band walkers
// ---------------------------------------------------------------------------------------------
template <class C>
struct BandRef {
	C *p;
	int dimx, dimy, stride;
};

template <class C>
RIC_HD inline BandRef<C> band_ref(const HostGeom &g, char *plane, int id)
{
	const ric_band_info &b = g.band[id];
	return BandRef<C>{(C *)(plane + b.offset), b.dimx, b.dimy, b.stride};
}

// LL band: 2-D DPCM, residuals through the geometric model with the local gradient as context
// (bandcodec.cpp:62-110).
template <class Port, class C>
RIC_HD void walk_ll(Port &io, BandRef<C> b)
{
	GeomModel geo(io.T->ll_init, io.T->geo_bound);
	C *row = b.p;
	if (Port::writing) io.m_taboo(fold_signed(row[0]));
	else row[0] = (C)unfold_signed((int)io.m_taboo(0));
	for (int x = 1; x < b.dimx; x++) {
		const uint32_t u = geo.code(io, Port::writing ? fold_signed(row[x] - row[x - 1]) : 0, 15);
		if (!Port::writing) row[x] = (C)(row[x - 1] + unfold_signed((int)u));
	}
	for (int y = 1; y < b.dimy; y++) {
		if constexpr (!Port::writing) if (io.m.overrun()) return;
		const C *up = row;
		row += b.stride;
		const uint32_t u0 = geo.code(io, Port::writing ? fold_signed(row[0] - up[0]) : 0, 15);
		if (!Port::writing) row[0] = (C)(up[0] + unfold_signed((int)u0));
		for (int x = 1; x < b.dimx; x++) {
			const int w = row[x - 1], n = up[x], nw = up[x - 1];
			const int dw = w - nw, dn = n - nw;
			int ctx = bit_length((uint32_t)((dw < 0 ? -dw : dw) + (dn < 0 ? -dn : dn)));
			if (ctx > 15) ctx = 15;  // the reference would index past its 16 contexts; unreachable for real data
			const uint32_t u = geo.code(io, Port::writing ? fold_signed(row[x] - w - n + nw) : 0, ctx);
			if (!Port::writing) row[x] = (C)(w + n - nw + unfold_signed((int)u));
		}
	}
}

// Bit i set <=> sample i (row-major) of a full 4x4 block is non-zero.
RIC_HD inline uint32_t nonzero_mask(const int16_t *blk, int stride)
{
#if defined(__SSE2__) && !defined(__CUDA_ARCH__)
	const __m128i z = _mm_setzero_si128();
	const __m128i r01 = _mm_unpacklo_epi64(_mm_loadl_epi64((const __m128i *)blk), _mm_loadl_epi64((const __m128i *)(blk + stride)));
	const __m128i r23 = _mm_unpacklo_epi64(_mm_loadl_epi64((const __m128i *)(blk + 2 * stride)),
	                                       _mm_loadl_epi64((const __m128i *)(blk + 3 * stride)));
	const __m128i zero = _mm_packs_epi16(_mm_cmpeq_epi16(r01, z), _mm_cmpeq_epi16(r23, z));  // 0xFF per zero sample
	return (uint32_t)~_mm_movemask_epi8(zero) & 0xFFFFu;
#else
	uint32_t m = 0;
	for (int i = 0; i < 16; i++) m |= (uint32_t)(blk[(i >> 2) * stride + (i & 3)] != 0) << i;
	return m;
#endif
}
RIC_HD inline uint32_t nonzero_mask(const int32_t *blk, int stride)
{
	uint32_t m = 0;
	for (int i = 0; i < 16; i++) m |= (uint32_t)(blk[(i >> 2) * stride + (i & 3)] != 0) << i;
	return m;
}

// Coefficients of one block (w x h samples, 16 when full): count k, which positions, then magnitude-1
// through the geometric model and a raw sign bit each (bandcodec.cpp:347-482).
template <bool FINE, class Port, class C>
RIC_HD inline unsigned code_block(Port &io, GeomModel &geo, C *blk, int stride, int w, int h, int table)
{
	const Tables &T = *io.T;
	const unsigned n = (unsigned)(w * h);
	const bool full = n == 16;
	uint32_t mask = 0;  // bit i = scan position i
	unsigned k = 0;
	if (Port::writing) {
		if (full) mask = nonzero_mask(blk, stride);
		else {
			const C *r = blk;
			for (int y = 0, i = 0; y < h; y++, r += stride)
				for (int x = 0; x < w; x++, i++) mask |= (uint32_t)(r[x] != 0) << i;
		}
		k = (unsigned)count_ones(mask);
		if constexpr (Port::writing)
			if (FINE && k == 0) { io.m.poison(); return 0; }  // an all-zero block without the INSIGNIF_BLOCK mark: malformed arena
	}
	if (full) {
		if constexpr (Port::writing) {
			const CountCode &c = FINE ? T.fine[table] : T.low[table];
			const unsigned s = FINE ? k - 1 : k;
			io.bits(c.code[s], c.len[s]);
		} else
			k = FINE ? io.m.count_symbol(T.fine[table]) + 1 : io.m.count_symbol(T.low[table]);
	} else
		k = FINE ? truncated(io, k - 1, n - 1) + 1 : truncated(io, k, n);
	if (FINE || k != 0) {
		if (k != n) mask = combination(io, mask, k, n);
		else mask = (1u << n) - 1;
		const unsigned ctx = full ? k - 1 : T.edge_ctx[n][k - 1];
		for (uint32_t m = mask; m; m &= m - 1) {  // set positions in scan order
			const unsigned i = (unsigned)count_tz(m);
			C *p = full ? blk + (i >> 2) * stride + (i & 3) : blk + (i / (unsigned)w) * stride + i % (unsigned)w;
			if (Port::writing) {
				const uint32_t u = sizeof(C) == 2 ? (uint16_t)*p : (uint32_t)*p;
				geo.code(io, (u >> 1) - 1, ctx);
				io.bits(u & 1, 1);
			} else {
				const uint32_t mag = geo.code(io, 0, ctx) + 1;
				*p = (C)unfold_sign_lsb((int)((mag << 1) | io.bits(0, 1)));
			}
		}
	}
	return k - (FINE ? 1 : 0);
}

// Largest magnitude of the 2x2 parent samples that cover one block (bandcodec.cpp:324-345): folded
// values while writing, signed values while reading.
template <bool WRITING, class P>
RIC_HD inline int parent_ctx(const P *par, int stride)
{
	if (WRITING) {
		P m = 0;
		for (int y = 0; y < 2; y++, par += stride)
			for (int x = 0; x < 2; x++) m = par[x] > m ? par[x] : m;
		const uint32_t u = sizeof(P) == 2 ? (uint16_t)m : (uint32_t)m;
		return bit_length(u >> 1);
	}
	P hi = 0, lo = 0;
	for (int y = 0; y < 2; y++, par += stride)
		for (int x = 0; x < 2; x++) { hi = par[x] > hi ? par[x] : hi; lo = par[x] < lo ? par[x] : lo; }
	P a = lo < 0 ? (P)-lo : lo;
	return bit_length((uint32_t)(a > hi ? a : hi));
}

// One D/H/V band: 4x4 blocks in serpentine order; a full block is either insignificant (one adaptive
// bit, context = parent magnitude class; implied when the parent's covering 2x2 cell is marked) or coded
// with code_block; partial blocks at the right/bottom edge use their own one-context bit
// (bandcodec.cpp:484-590).  FINE = finest level (no children).
template <bool FINE, class Port, class C, class P>
RIC_HD void walk_band(Port &io, BandRef<C> b, const BandRef<P> *parent, bool has_child)
{
	uint16_t kmean[16];
	for (int i = 0; i < 16; i++) kmean[i] = io.T->kmean_init[i];
	GeomModel geo(io.T->band_init, io.T->geo_bound);
	BitModel tree_bit(io.T->bit_lim), edge_bit(io.T->bit_lim);
	const C mark = (C)(has_child ? kMarker : 0);
	const int nfull = b.dimx >> 2, rem = b.dimx & 3;
	for (int y = 0; y < b.dimy; y += 4) {
		if constexpr (!Port::writing) if (io.m.overrun()) return;  // ran off the end of a truncated stream
		const int h = b.dimy - y < 4 ? b.dimy - y : 4;
		C *row = b.p + (size_t)y * b.stride;
		P *prow = parent ? parent->p + (size_t)(y >> 1) * parent->stride : 0;
		const bool backwards = (y & 4) != 0;
		const int nblk = nfull + (rem ? 1 : 0);
		for (int t = 0; t < nblk; t++) {
			const int bx = backwards ? nblk - 1 - t : t;
			const int x = bx * 4, w = bx < nfull ? 4 : rem;
			C *blk = row + x;
			if (w == 4 && h == 4) {
				int ctx = 15;
				if (parent) {
					P *cell = prow + (x >> 1);
					if (*cell == kMarker) {
						*cell = 0;
						blk[0] = blk[2] = blk[2 * b.stride] = blk[2 * b.stride + 2] = mark;
						continue;
					}
					ctx = parent_ctx<Port::writing, P>(cell, parent->stride);
					if (ctx > 15) ctx = 15;  // as in walk_ll: out of the reference's model range, unreachable
				}
				if (tree_bit.code(io, Port::writing ? blk[0] == kMarker : 0, ctx)) {
					blk[0] = blk[2] = blk[2 * b.stride] = blk[2 * b.stride + 2] = mark;
				} else {
					const int table = (kmean[ctx] + (1 << 9)) >> 10;
					const unsigned k = code_block<FINE>(io, geo, blk, b.stride, 4, 4, table);
					kmean[ctx] = (uint16_t)(kmean[ctx] + (k << 7) - (kmean[ctx] >> 3));
				}
			} else {
				if (parent && (x >> 1) < parent->dimx && (y >> 1) < parent->dimy && prow[x >> 1] == kMarker) prow[x >> 1] = 0;
				if (edge_bit.code(io, Port::writing ? blk[0] == kMarker : 0, 0)) {
					if (Port::writing) blk[0] = 0;
				} else
					code_block<FINE>(io, geo, blk, b.stride, w, h, 0);
			}
		}
	}
}

struct WPort : WritePort {
	RIC_HD WPort(MuxWriter &w, const Tables *t) : WritePort{w, t} {}
	RIC_HD inline uint32_t m_taboo(uint32_t v) { taboo(v); return v; }
};
struct RPort : ReadPort {
	RIC_HD RPort(MuxReader &r, const Tables *t) : ReadPort{r, t} {}
	RIC_HD inline uint32_t m_taboo(uint32_t) { return m.taboo(*T); }
};

template <bool FINE, class Port, class C, class P>
RIC_HD void walk_level(Port &io, const HostGeom &g, char *plane, int lev, bool has_parent)
{
	for (int o = 0; o < 3; o++) {
		const int id = 3 * lev + 2 - o;  // V, H, D (wavelet2d.cpp:130-132)
		BandRef<P> par = has_parent ? band_ref<P>(g, plane, id + 3) : BandRef<P>{0, 0, 0, 0};
		walk_band<FINE, Port, C, P>(io, band_ref<C>(g, plane, id), has_parent ? &par : 0, lev > 0);
	}
}

template <class Port>
RIC_HD void walk_plane(Port &io, const HostGeom &g, char *plane)
{
	const int n = g.nlev, ll = 3 * n;
#ifndef __CUDA_ARCH__  // on the device the caller clears the arenas with one wide memset
	if (!Port::writing) memset(plane, 0, g.arena_bytes);
#endif
	;  // CBand::Clear of every band (bandcodec.cpp:503) + deterministic padding
	if (g.band[ll].is_int) walk_ll(io, band_ref<int32_t>(g, plane, ll));
	else walk_ll(io, band_ref<int16_t>(g, plane, ll));
	for (int lev = n - 1; lev >= 0; lev--) {
		const bool has_parent = lev < n - 1;
		const bool ci = g.lev_int[lev] != 0, pi = has_parent ? g.lev_int[lev + 1] != 0 : ci;
		if (lev == 0) {
			if (ci) walk_level<true, Port, int32_t, int32_t>(io, g, plane, lev, has_parent);
			else if (pi) walk_level<true, Port, int16_t, int32_t>(io, g, plane, lev, has_parent);
			else walk_level<true, Port, int16_t, int16_t>(io, g, plane, lev, has_parent);
		} else {
			if (ci) walk_level<false, Port, int32_t, int32_t>(io, g, plane, lev, has_parent);
			else if (pi) walk_level<false, Port, int16_t, int32_t>(io, g, plane, lev, has_parent);
			else walk_level<false, Port, int16_t, int16_t>(io, g, plane, lev, has_parent);
		}
	}
}


// ---------------------------------------------------------------------------------------------
// Encoding with a parallel pre-pass ("block hints").
//
// Everything the band walker derives from the band DATA alone -- whether a block is skipped because its
// parent block is insignificant, whether it is insignificant itself, its parent-magnitude context, which of
// its samples are non-zero and the index of that combination -- does not depend on the coder's adaptive state,
// so it can be computed for all blocks at once (one thread per block on the device) and leave the serial
// coder with the models, the range coder and the bit fields.  The hinted walker reads the bands but never
// writes them (the in-place marker bookkeeping of walk_band is what the hints replace).
//
// The skip rule relies on the quantiser's own invariant (buildTree accumulates children into parents,
// bandcodec.cpp:291-296): a block under an insignificant full parent block is insignificant itself, so
// "parent cell marked at walk time" == "the parent block is full and carries the marker".  Arenas from the
// encode stage always satisfy it; for arbitrary arenas use the plain walker.
// ---------------------------------------------------------------------------------------------
struct BlockHint {
	uint32_t a;  // bits 0-15 non-zero mask (scan order), 16-19 parent context, 20-21 state
	uint32_t idx;  // index of the combination (flipped when k > 8), ready for the truncated code
};
enum : uint32_t { kHintDead = 0, kHintInsig = 1, kHintSig = 2 };

// Hint slot of block (bx, by) of band `id`: the block-flag indexing of HostGeom (one slot per 4x4 block).
RIC_HD inline size_t hint_slot(const HostGeom &g, int id, int bx, int by) { return (size_t)g.flag_off[id] + (size_t)by * g.flag_bw[id] + bx; }

template <class C, class P>
RIC_HD inline BlockHint make_hint_t(const HostGeom &g, const Tables &T, const char *plane, int id, int bx, int by)
{
	const ric_band_info &b = g.band[id];
	const C *blk = (const C *)(plane + b.offset) + (size_t)(4 * by) * b.stride + 4 * bx;
	const int w = b.dimx - 4 * bx < 4 ? b.dimx - 4 * bx : 4, h = b.dimy - 4 * by < 4 ? b.dimy - 4 * by : 4;
	const bool full = w == 4 && h == 4;
	BlockHint r{0, 0};
	uint32_t state = blk[0] == kMarker ? kHintInsig : kHintSig, ctx = 15;
	const int lev = id / 3;
	if (full && lev < g.nlev - 1) {
		const ric_band_info &pb = g.band[id + 3];
		const P *par = (const P *)(plane + pb.offset);
		const int pbx = bx >> 1, pby = by >> 1;
		const bool pfull = pb.dimx - 4 * pbx >= 4 && pb.dimy - 4 * pby >= 4;
		if (pfull && par[(size_t)(4 * pby) * pb.stride + 4 * pbx] == kMarker) state = kHintDead;
		else {
			int c = parent_ctx<true, P>(par + (size_t)(2 * by) * pb.stride + 2 * bx, pb.stride);
			ctx = (uint32_t)(c > 15 ? 15 : c);
		}
	}
	uint32_t mask = 0;
	if (full && state == kHintSig) {
		mask = nonzero_mask(blk, b.stride);
		const unsigned k = (unsigned)count_ones(mask);
		if (k != 0 && k != 16) {
			uint32_t m = k > 8 ? mask ^ 0xFFFFu : mask, idx = 0;
			for (unsigned rr = 0; m; rr++) {
				const unsigned i = 31 - (unsigned)count_lz(m);
				idx += T.choose[rr][15 - i];
				m ^= 1u << i;
			}
			r.idx = idx;
		}
	}
	r.a = mask | ctx << 16 | state << 20;
	return r;
}

RIC_HD inline BlockHint make_hint(const HostGeom &g, const Tables &T, const char *plane, int id, int bx, int by)
{
	const int lev = id / 3;
	const bool ci = g.lev_int[lev] != 0, pi = lev < g.nlev - 1 ? g.lev_int[lev + 1] != 0 : ci;
	if (ci) return make_hint_t<int32_t, int32_t>(g, T, plane, id, bx, by);
	if (pi) return make_hint_t<int16_t, int32_t>(g, T, plane, id, bx, by);
	return make_hint_t<int16_t, int16_t>(g, T, plane, id, bx, by);
}

template <bool FINE, class C>
RIC_HD void walk_band_hinted(WPort &io, const C *base, const ric_band_info &b, const BlockHint *hints, int bw)
{
	const Tables &T = *io.T;
	uint16_t kmean[16];
	for (int i = 0; i < 16; i++) kmean[i] = T.kmean_init[i];
	GeomModel geo(T.band_init, T.geo_bound);
	BitModel tree_bit(T.bit_lim), edge_bit(T.bit_lim);
	const int nfull = b.dimx >> 2, rem = b.dimx & 3, nblk = nfull + (rem ? 1 : 0);
	for (int y = 0, by = 0; y < b.dimy; y += 4, by++) {
		const int h = b.dimy - y < 4 ? b.dimy - y : 4;
		const C *row = base + (size_t)y * b.stride;
		const BlockHint *hrow = hints + (size_t)by * bw;
		const bool backwards = (y & 4) != 0;
		for (int t = 0; t < nblk; t++) {
			const int bx = backwards ? nblk - 1 - t : t;
			const BlockHint hint = hrow[bx];
			const uint32_t state = (hint.a >> 20) & 3;
			const int w = bx < nfull ? 4 : rem;
			C *blk = const_cast<C *>(row) + 4 * bx;  // code_block's writing side only reads
			if (w == 4 && h == 4) {
				if (state == kHintDead) continue;
				const unsigned ctx = (hint.a >> 16) & 15;
				if (tree_bit.code(io, state == kHintInsig, ctx)) continue;
				const int table = (kmean[ctx] + (1 << 9)) >> 10;
				const uint32_t mask = hint.a & 0xFFFFu;
				const unsigned k = (unsigned)count_ones(mask);
				if (FINE && k == 0) { io.m.poison(); continue; }  // malformed arena (see code_block)
				const CountCode &c = FINE ? T.fine[table] : T.low[table];
				const unsigned s = FINE ? k - 1 : k;
				io.bits(c.code[s], c.len[s]);
				if (k != 0) {
					if (k != 16) {
						const unsigned kk = k > 8 ? 16 - k : k;
						const unsigned len = T.enum_len[16][kk];
						const uint32_t nshort = T.enum_short[16][kk];
						if (hint.idx < nshort) io.bits(hint.idx, len - 1); else io.bits(hint.idx + nshort, len);
					}
					for (uint32_t m = mask; m; m &= m - 1) {
						const unsigned i = (unsigned)count_tz(m);
						const C v = blk[(i >> 2) * b.stride + (i & 3)];
						const uint32_t u = sizeof(C) == 2 ? (uint16_t)v : (uint32_t)v;
						geo.code(io, (u >> 1) - 1, k - 1);
						io.bits(u & 1, 1);
					}
				}
				const unsigned kk2 = k - (FINE ? 1 : 0);
				kmean[ctx] = (uint16_t)(kmean[ctx] + (kk2 << 7) - (kmean[ctx] >> 3));
			} else {
				if (edge_bit.code(io, state == kHintInsig, 0)) continue;
				code_block<FINE>(io, geo, blk, b.stride, w, h, 0);
			}
		}
	}
}

// One plane, luma-first order handled by the caller.  `hints`: the plane's hint slots (g.flag_bytes of them).
RIC_HD inline void walk_plane_hinted(WPort &io, const HostGeom &g, char *plane, const BlockHint *hints)
{
	const int n = g.nlev, ll = 3 * n;
	if (g.band[ll].is_int) walk_ll(io, band_ref<int32_t>(g, plane, ll));
	else walk_ll(io, band_ref<int16_t>(g, plane, ll));
	for (int lev = n - 1; lev >= 0; lev--)
		for (int o = 0; o < 3; o++) {
			const int id = 3 * lev + 2 - o;  // V, H, D
			const ric_band_info &b = g.band[id];
			const BlockHint *h = hints + g.flag_off[id];
			if (g.lev_int[lev]) {
				const int32_t *p = (const int32_t *)(plane + b.offset);
				if (lev == 0) walk_band_hinted<true>(io, p, b, h, g.flag_bw[id]); else walk_band_hinted<false>(io, p, b, h, g.flag_bw[id]);
			} else {
				const int16_t *p = (const int16_t *)(plane + b.offset);
				if (lev == 0) walk_band_hinted<true>(io, p, b, h, g.flag_bw[id]); else walk_band_hinted<false>(io, p, b, h, g.flag_bw[id]);
			}
		}
}


}  // namespace ent
}  // namespace ric
