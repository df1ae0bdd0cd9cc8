// ric_swar.cuh -- two int16 samples per 32-bit register ("packed-linear" arithmetic) for the short levels.
//
// Why: the level kernels are bound by the integer issue slots (ALU and FMA pipe: one warp instruction
// per two clocks each, profiles/README.md), so the only way to go faster is to do two samples per
// instruction.  sm_100a has no packed arithmetic shift, but the lifting steps of the reference
// (src/lib/wavelet2d.cpp:307-359) are adds and right shifts only, and those can be done on two samples at
// once with plain 32-bit instructions as long as no sample leaves the int16 range:
//
//   A register R represents the pair (hi, lo) as  R = 65536*hi + lo + K  (mod 2^32), K a compile-time
//   constant that is tracked per pipeline role.  Adds and subtracts of such registers are exact (the
//   representation is linear; a negative lo simply borrows from the upper half).  A right shift by k is
//   done on a register B whose constant puts BOTH halves into [0, 65536) and is a multiple of 2^k in both
//   halves: (B >> k) & mask gives 65536*((hi + bh) >> k) + ((lo + bl) >> k) -- a logical shift, plus one
//   AND that removes the k bits of the upper half that fell into the lower one -- and
//   (v + b) >> k == (v >> k) + b / 2^k exactly (floor semantics) when 2^k divides b.
//
// Exactness: the reference works on `short` with wrap-around on every store; packed-linear arithmetic is
// only equal to it while nothing wraps.  Rows lifted straight from 8-bit pixels cannot wrap (bounds in
// ric_dev.cuh, row_fwd); everywhere else the kernels check a bound on the operands first (a few LOP3 per
// row) and fall back to the scalar, exactly-wrapping code of ric_dev.cuh for the rows that fail it.
// Every function here is __host__ __device__ so that tests/cpp/swar_test.cu can check it on the CPU
// against a direct transcription of the reference formulas.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ric {
namespace sw {

#define RIC_HD __host__ __device__ __forceinline__

// (b >> k) with the bits that crossed from the upper half into the lower half cleared
template <int K>
struct Msk { static constexpr unsigned v = ~(((1u << K) - 1u) << (16 - K)); };
template <int K>
RIC_HD unsigned lsr(unsigned b) { return (b >> K) & Msk<K>::v; }

RIC_HD unsigned prmt(unsigned a, unsigned b, unsigned sel)
{
#ifdef __CUDA_ARCH__
	return __byte_perm(a, b, sel);
#else
	unsigned long long v = ((unsigned long long)b << 32) | a;
	unsigned r = 0;
	for (int i = 0; i < 4; i++) {
		const unsigned n = (sel >> (4 * i)) & 7u;  // (like __byte_perm: bit 3 of a selector nibble is ignored)
		r |= (unsigned)((v >> (8 * n)) & 0xFF) << (8 * i);
	}
	return r;
#endif
}

// 0xFFFF in every half whose sign bit is set, 0 elsewhere.  (PTX prmt replicates the sign of the selected byte
// when bit 3 of a selector nibble is set; the __byte_perm intrinsic ignores that bit, hence the asm.)
RIC_HD unsigned smear(unsigned x)
{
#ifdef __CUDA_ARCH__
	unsigned r;
	asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(r) : "r"(x));
	return r;
#else
	return ((x & 0x8000u) ? 0xFFFFu : 0u) | ((x & 0x80000000u) ? 0xFFFF0000u : 0u);
#endif
}

constexpr unsigned OB = 0x80008000u;  // "offset binary": both halves hold value + 32768, nothing borrowed

// encode / decode for tests and the scalar fall-back: K must keep both halves inside [0, 65536)
RIC_HD unsigned enc(int lo, int hi, unsigned K) { return (unsigned)hi * 65536u + (unsigned)lo + K; }
RIC_HD int dec_lo(unsigned R, unsigned K) { return (int)((R - K + OB) & 0xFFFFu) - 32768; }
RIC_HD int dec_hi(unsigned R, unsigned K) { return (int)((R - K + OB) >> 16) - 32768; }
// two's-complement halves (what the band arenas and the LL scratch hold) from any representation
RIC_HD unsigned to_c2(unsigned R, unsigned K) { return (R + (OB - K)) ^ OB; }
RIC_HD unsigned from_c2(unsigned c2, unsigned K) { return (c2 ^ OB) + (K - OB); }

// ---- constants of one 9/7 lifting pass --------------------------------------------------------------
// O0 / E0: constants of the raw odd / even samples.  Every derived constant follows from the step
// formulas below; C3 / C4 are the corrections added to the neighbour sums of S3 / S4 so that their shift
// constants become OB (zero when the chain already ends there).
template <unsigned O0_, unsigned E0_>
struct LiftK {
	static constexpr unsigned O0 = O0_, E0 = E0_;
	static constexpr unsigned B1 = 2u * O0;               // S1: b = l + r
	static constexpr unsigned E1 = E0 - B1 - (B1 >> 1);   //     x - b - (b >> 1)
	static constexpr unsigned B2 = 2u * E1;               // S2: b = l + r
	static constexpr unsigned O2 = O0 - (B2 >> 4);        //     x - (b >> 4)
	static constexpr unsigned C3 = OB - 2u * O2;          // S3: a = l + r + C3 (constant OB)
	static constexpr unsigned A3a = OB - (OB >> 2);       //     a -= a >> 2      0x60006000
	static constexpr unsigned A3b = A3a + (A3a >> 4);     //     a += a >> 4      0x66006600
	static constexpr unsigned E3 = E1 + A3b + (A3b >> 8); //     x + a + (a >> 8)
	static constexpr unsigned C4 = OB - 2u * E3;          // S4: b = l + r + C4 (constant OB)
	static constexpr unsigned O4 = O2 + (OB >> 1) - (OB >> 5);  // x + (b >> 1) - (b >> 5)
	static_assert((B1 & 0x00010001u) == 0, "S1 shift constant must be even in both halves");
	static_assert((B2 & 0x000F000Fu) == 0, "S2 shift constant must be a multiple of 16 in both halves");
};

// forward 9/7 steps, interior formulas; x: centre, l / r: neighbours
template <class K>
RIC_HD unsigned s1(unsigned x, unsigned l, unsigned r) { const unsigned b = l + r; return x - b - lsr<1>(b); }
template <class K>
RIC_HD unsigned s2(unsigned x, unsigned l, unsigned r) { return x - lsr<4>(l + r); }
template <class K>
RIC_HD unsigned s3(unsigned x, unsigned l, unsigned r)
{
	unsigned a = K::C3 ? l + r + K::C3 : l + r;
	a -= lsr<2>(a);
	a += lsr<4>(a);
	return x + a + lsr<8>(a);
}
template <class K>
RIC_HD unsigned s4(unsigned x, unsigned l, unsigned r)
{
	const unsigned b = K::C4 ? l + r + K::C4 : l + r;
	return x + lsr<1>(b) - lsr<5>(b);
}

// edge formulas of the row pass (wavelet2d.cpp:324-359): one neighbour n.  Results carry the same
// constants as the interior formulas, so they can overwrite an interior result in place.
template <class K>
RIC_HD unsigned s1_edge(unsigned x, unsigned n) { return x - 3u * n + (K::E1 - (K::E0 - 3u * K::O0)); }  // x -= 3 n
template <class K>
RIC_HD unsigned s2_last(unsigned x, unsigned l)  // x -= l >> 3
{
	constexpr unsigned CB = OB - K::E1;
	return x - lsr<3>(l + CB) + (K::O2 - (K::O0 - (OB >> 3)));
}
template <class K>
RIC_HD unsigned s3_edge(unsigned x, unsigned n)  // x += 2 * mult08(n)
{
	unsigned a = n + (OB - K::O2);
	a -= lsr<2>(a);
	a += lsr<4>(a);
	const unsigned m = a + lsr<8>(a);  // constant A3b + (A3b >> 8)
	return x + 2u * m + (K::E3 - (K::E1 + 2u * (K::A3b + (K::A3b >> 8))));
}
template <class K>
RIC_HD unsigned s4_last(unsigned x, unsigned l)  // x += l - (l >> 4)
{
	const unsigned b = l + (OB - K::E3);
	return x + b - lsr<4>(b) + (K::O4 - (K::O2 + OB - (OB >> 4)));
}

// ---- the two passes of the level-0 kernels ---------------------------------------------------------------
// Row pass on rows lifted straight from 8-bit pixels (|x| <= 2048: S1 <= 8192, S2 <= 3073, S3 <= 13109,
// S4 <= 17002, neighbour sums <= 26218): constants chosen so that S1..S3 need no correction at all.
typedef LiftK<0x48004800u, 0x40004000u + 0x90009000u + 0x48004800u> KH;
static_assert(KH::E1 == 0x40004000u && KH::O2 == 0x40004000u && KH::C3 == 0u, "row-pass constants");
// Column pass: raw odd rows carry 0x1000 per half and S1'd even rows 0x2000, so that the operand bounds
// the fast path needs -- |odd| <= 4095, |S1'd even| <= 8191 -- are plain bit tests on the registers.
typedef LiftK<0x10001000u, 0x50005000u> KV;
static_assert(KV::E1 == 0x20002000u, "column-pass constants");
constexpr unsigned GUARD_O = 0xE000E000u;  // (R & GUARD_O) == 0  <=>  both raw odd samples in [-4096, 4095]
constexpr unsigned GUARD_E = 0xC000C000u;  // (R & GUARD_E) == 0  <=>  both S1'd even samples in [-8192, 8191]
// bounds that then hold for the derived rows (ric_fwd0.cuh, v_slow re-checks them on the scalar path):
constexpr int BOUND_O0 = 4095, BOUND_E1 = 8191, BOUND_O2 = 5119, BOUND_E3 = 16381;


// ---- inverse 9/7 steps (src/lib/wavelet2d.cpp:361-405) ---------------------------------
// Here every register, at every stage, carries the SAME constant G = 0x2000 per half: a value is inside the
// range the packed path may use, [-8192, 8191], exactly when (R & GUARD_I) == 0, so one OR over everything an
// iteration produced tests all of it at once.  With every stored value in that range no intermediate of the next
// step can leave int16 (sums <= 16382, t + (t >> 1) <= 24573, x + that <= 32764), which is what makes the linear
// arithmetic equal to the reference's wrapping `short` arithmetic.  The constants that restore G ride on the
// three-input adds.
constexpr unsigned G = 0x20002000u;
constexpr unsigned GUARD_I = 0xC000C000u;
constexpr unsigned CSUM = OB - 2u * G;  // neighbour sum l + r + CSUM has constant OB
RIC_HD unsigned u4(unsigned x, unsigned l, unsigned r)  // x -= (t >> 1) - (t >> 5), t = l + r
{
	const unsigned b = l + r + CSUM;
	return x + ((OB >> 1) - (OB >> 5)) + lsr<5>(b) - lsr<1>(b);
}
RIC_HD unsigned u3(unsigned x, unsigned l, unsigned r)  // x -= mult08(l + r)
{
	unsigned a = l + r + CSUM;
	a -= lsr<2>(a);
	a += lsr<4>(a);  // constant 0x6600 per half, then + (0x6600 >> 8)
	return x + 0x66666666u - a - lsr<8>(a);
}
RIC_HD unsigned u2(unsigned x, unsigned l, unsigned r) { return x - (OB >> 4) + lsr<4>(l + r + CSUM); }  // x += (l + r) >> 4
RIC_HD unsigned u1(unsigned x, unsigned l, unsigned r)  // x += t + (t >> 1)
{
	const unsigned b = l + r + CSUM;
	return x - (OB + (OB >> 1)) + b + lsr<1>(b);
}
// edge formulas of the row pass: one neighbour n (wavelet2d.cpp:365-368,388-403)
RIC_HD unsigned u4_last(unsigned x, unsigned l)  // x -= l - (l >> 4)
{
	const unsigned b = l + (OB - G);
	return x + (OB - (OB >> 4)) - b + lsr<4>(b);
}
RIC_HD unsigned u3_edge(unsigned x, unsigned n)  // x -= 2 * mult08(n)
{
	unsigned a = n + (OB - G);
	a -= lsr<2>(a);
	a += lsr<4>(a);
	const unsigned m = a + lsr<8>(a);  // constant 0x6666 per half
	return x + 2u * 0x66666666u - 2u * m;
}
RIC_HD unsigned u2_last(unsigned x, unsigned l) { return x - (OB >> 3) + lsr<3>(l + (OB - G)); }  // x += l >> 3
RIC_HD unsigned u1_edge(unsigned x, unsigned n) { return x - 3u * G + 3u * n; }                    // x += 3 n

// TSUQi (src/lib/band.h:94-107) on a pair of two's-complement coefficients: c * q per half, constant G.
// Only valid while |c * q| <= 8191 in both halves (the caller bounds |c| first): kfix = G - OB * q.
RIC_HD unsigned dequant(unsigned c2, unsigned q, unsigned kfix) { return (c2 ^ OB) * q + kfix; }

// Pixel outputs are clamped as value + PIXK per half: both halves non-negative (so the register is readable half by
// half) and the low byte of each half is the pixel.
RIC_HD unsigned vmaxs2(unsigned a, unsigned b)
{
#ifdef __CUDA_ARCH__
	return __vmaxs2(a, b);
#else
	const int al = (short)(a & 0xFFFF), ah = (short)(a >> 16), bl = (short)(b & 0xFFFF), bh = (short)(b >> 16);
	return (unsigned)((al > bl ? al : bl) & 0xFFFF) | (unsigned)((ah > bh ? ah : bh) & 0xFFFF) << 16;
#endif
}
RIC_HD unsigned vmins2(unsigned a, unsigned b)
{
#ifdef __CUDA_ARCH__
	return __vmins2(a, b);
#else
	const int al = (short)(a & 0xFFFF), ah = (short)(a >> 16), bl = (short)(b & 0xFFFF), bh = (short)(b >> 16);
	return (unsigned)((al < bl ? al : bl) & 0xFFFF) | (unsigned)((ah < bh ? ah : bh) & 0xFFFF) << 16;
#endif
}
constexpr unsigned PIXK = 0x41004100u;
RIC_HD unsigned clip_pix(unsigned v) { return vmins2(vmaxs2(v, PIXK), PIXK + 0x00FF00FFu); }  // clamp(value, 0, 255) + PIXK
// inverse level shift of a gray pair: clip(128 + ((v + 8) >> 4)), ric.cpp:237-240
RIC_HD unsigned gray_out(unsigned v)  // v: constant G
{
	const unsigned s = lsr<4>(v + (OB - G) + 0x00080008u);     // ((v + 8) >> 4) + 0x0800
	return clip_pix(s + (PIXK + 0x00800080u - (OB >> 4)));     // + 128
}
// YCoCgtoRGB<4> (ric.cpp:93-112) on pairs, constant G in, clamped value + PIXK out (r, g, b by reference)
RIC_HD void ycocg_out(unsigned co, unsigned cg, unsigned y, unsigned &r, unsigned &g, unsigned &b)
{
	const unsigned co3 = lsr<3>(co + (OB - G) + 0x00040004u);  // ((co + 4) >> 3) + 0x1000
	const unsigned cg3 = lsr<3>(cg + (OB - G) + 0x00040004u);  // ((cg + 4) >> 3) + 0x1000
	const unsigned y4 = lsr<4>(y + (OB - G) + 0x00080008u);    // ((y + 8) >> 4) + 0x0800
	const unsigned yb = y4 + 0x00800080u - lsr<1>(cg3);        // y -= (cg >> 1) - 128      constant 0: may be "negative"
	const unsigned gg = cg3 + yb;                              // cg += y                  constant 0x1000
	const unsigned bb = yb - lsr<1>(co3);                      // y -= co >> 1             constant -0x0800
	const unsigned rr = co3 + bb;                              // co += y                  constant 0x0800
	r = clip_pix(rr + (PIXK - 0x08000800u));
	g = clip_pix(gg + (PIXK - 0x10001000u));
	b = clip_pix(bb + (PIXK + 0x08000800u));
}

}  // namespace sw
}  // namespace ric
