// ric_quant_pk.cuh -- the encode quantiser's scalars (QuantBand) and its two-samples-per-instruction form.
// __host__ __device__ throughout: tests/cpp/quant_pk_test.cu checks it on the CPU against the oracle's
// restatement of CBandCodec::tsuqBlock (src/lib/bandcodec.cpp:159-237).
#pragma once
#include "ric_swar.cuh"

namespace ric {

namespace sw {
RIC_HD unsigned vadd2(unsigned a, unsigned b)  // per-half add, wrapping (VIADD.16x2)
{
#ifdef __CUDA_ARCH__
	return __vadd2(a, b);
#else
	return ((a & 0xFFFFu) + (b & 0xFFFFu) & 0xFFFFu) | ((a >> 16) + (b >> 16)) << 16;
#endif
}
RIC_HD unsigned viaddmax2(unsigned a, unsigned b, unsigned c)  // per-half max(a + b, c), signed (VIADDMNMX.S16x2)
{
#ifdef __CUDA_ARCH__
	return __viaddmax_s16x2(a, b, c);
#else
	const unsigned s = vadd2(a, b);
	const int sl = (short)(s & 0xFFFF), sh = (short)(s >> 16), cl = (short)(c & 0xFFFF), ch = (short)(c >> 16);
	return (unsigned)((sl > cl ? sl : cl) & 0xFFFF) | (unsigned)((sh > ch ? sh : ch) & 0xFFFF) << 16;
#endif
}
}  // namespace sw

// ---- encode quantiser (CBandCodec::tsuqBlock, bandcodec.cpp:159-237) ----------------------------
// Host-computed scalars of one band (buildTree :243-247, makeThres :149-157).
struct QuantBand {
	int Q, iQ, T, Te;  // T = Q>>1 (full blocks), Te = (Q+((Q-(Q>>2))>>1))>>1 (partial blocks)
	int thr[16];
	int fast;          // 1 <= Q <= 16383: every candidate / threshold is a non-negative int16, so signed and
	                   // unsigned views agree and the rank thresholds can be compared in the key domain
	int kthr[32];      // fast: thr[n] << 4 (n < 16), INT_MAX beyond
	int pk;            // the two-samples-per-instruction form applies (ric_fwd.cuh quant_rows_pk; host-checked in fill_qb)
	int h0;            // pk: max(thr[0] >> 1, T + 1): |c| >= h0 <=> quantised for sure
	int kt16[32];      // pk: 0x8000 | (thr[n] - 2T) << 4 (n < 16), INT_MAX beyond: thr[] in the domain of the 16-bit keys
};

template <bool SH>
RIC_HD unsigned uview(int v) { return SH ? (unsigned)(v & 0xFFFF) : (unsigned)v; }

// Batcher odd-even merge sort of 16 keys, descending: 63 compare-exchanges written out so that the
// keys provably stay in registers (a rolled network would index them dynamically -> local memory).
RIC_HD void sort16_desc(int (&s)[16])
{
#define RIC_CE(i, j) { const int a_ = s[i], b_ = s[j]; s[i] = max(a_, b_); s[j] = min(a_, b_); }
	RIC_CE(0, 1) RIC_CE(2, 3) RIC_CE(4, 5) RIC_CE(6, 7) RIC_CE(8, 9) RIC_CE(10, 11) RIC_CE(12, 13)
	RIC_CE(14, 15) RIC_CE(0, 2) RIC_CE(1, 3) RIC_CE(4, 6) RIC_CE(5, 7) RIC_CE(8, 10) RIC_CE(9, 11)
	RIC_CE(12, 14) RIC_CE(13, 15) RIC_CE(1, 2) RIC_CE(5, 6) RIC_CE(9, 10) RIC_CE(13, 14) RIC_CE(0, 4)
	RIC_CE(1, 5) RIC_CE(2, 6) RIC_CE(3, 7) RIC_CE(8, 12) RIC_CE(9, 13) RIC_CE(10, 14) RIC_CE(11, 15)
	RIC_CE(2, 4) RIC_CE(3, 5) RIC_CE(10, 12) RIC_CE(11, 13) RIC_CE(1, 2) RIC_CE(3, 4) RIC_CE(5, 6)
	RIC_CE(9, 10) RIC_CE(11, 12) RIC_CE(13, 14) RIC_CE(0, 8) RIC_CE(1, 9) RIC_CE(2, 10) RIC_CE(3, 11)
	RIC_CE(4, 12) RIC_CE(5, 13) RIC_CE(6, 14) RIC_CE(7, 15) RIC_CE(4, 8) RIC_CE(5, 9) RIC_CE(6, 10)
	RIC_CE(7, 11) RIC_CE(2, 4) RIC_CE(3, 5) RIC_CE(6, 8) RIC_CE(7, 9) RIC_CE(10, 12) RIC_CE(11, 13)
	RIC_CE(1, 2) RIC_CE(3, 4) RIC_CE(5, 6) RIC_CE(7, 8) RIC_CE(9, 10) RIC_CE(11, 12) RIC_CE(13, 14)
#undef RIC_CE
}

// ---- the same two stages, two samples per instruction (QuantBand::pk) -----------------------------------------
// Per packed register w (two int16): a = |w| (one VIADDMNMX), "outside the dead zone" and "at least thr[0]" as
// the sign bits of a + const, spread to half-word masks by one PRMT each; both halves are multiplied by iQ
// (the only scalar work) and the result is merged, sign-folded and masked with three LOP3.  A candidate
// (live, below thr[0]) quantises to 0 or 1 by itself, so OR-ing 2 into its half gives the 2|sign the reference
// stores.  Keys are 16-bit: 0x8000 | (f - 2T) << 4 | 15 - raster for candidates (f = 2|c| + sign), else 0 --
// as signed half-words candidates are negative and everything else is zero, so "ranked below the last survivor"
// is one packed subtract and a sign spread.  Needs (host-checked, fill_qb): 4 <= Q, thr[0] - 2T < 2047,
// sure samples quantise to >= 1; and no -32768 in the block (the caller checks: s2u_ wraps there).
// IO: get(r) / put(r, row) the block's row r (4 packed samples), get_key(r) / put_key(r, keys): lane-private staging.
// IO::UNROLL == 1: rows two at a time in a rolled loop over shared-memory staging (kernels whose hot loop is at the
// limit of the instruction cache: ric_fwd0.cuh); IO::UNROLL == 2: flat code on registers (ric_fwd.cuh).
template <class IO>
RIC_HD int quant_rows_pk(const QuantBand *qb, int bw, int bh, IO &io, int &nc)  // IO::UNROLL: 1 (rolled) or 2 (flat)
{
	const bool full = bw == 4 && bh == 4;
	const int T = full ? qb->T : qb->Te;
	const int hl = full ? qb->h0 : T + 1;  // partial blocks: everything outside the dead zone is quantised (:215-237)
	// columns beyond the band (bw < 4): constant 0 keeps the sign bit of a + const clear (a <= 32767)
	const unsigned vm0 = bw >= 2 ? 0xFFFFFFFFu : bw == 1 ? 0x0000FFFFu : 0u, vm1 = bw == 4 ? 0xFFFFFFFFu : bw == 3 ? 0x0000FFFFu : 0u;
	const unsigned CL = (unsigned)(0x7FFF - T) * 0x10001u, CS = (unsigned)(0x8000 - hl) * 0x10001u;
	const unsigned cl[2] = {CL & vm0, CL & vm1}, cs[2] = {CS & vm0, CS & vm1};
	const unsigned NT = (unsigned)((-T) & 0xFFFF) * 0x10001u;
	const unsigned iQ = (unsigned)qb->iQ, iQ2 = 2u * iQ;
	unsigned accL = 0, accS = 0;
#pragma unroll IO::UNROLL
	for (int r0 = 0; r0 < 4; r0 += 2) {
#pragma unroll
		for (int rr = 0; rr < 2; rr++) {
			const int r = r0 + rr;
			if (r >= bh) { io.put(r, make_uint2(0u, 0u)); io.put_key(r, make_uint2(0u, 0u)); continue; }  // warp-uniform
			const uint2 row = io.get(r);
			unsigned out[2], key[2];
#pragma unroll
			for (int g = 0; g < 2; g++) {
				const unsigned w = g ? row.y : row.x;
				const unsigned a = sw::viaddmax2(~w, 0x00010001u, w);                  // |w|
				const unsigned LM = sw::smear(sw::vadd2(a, cl[g]));                       // outside the dead zone
				const unsigned SM = sw::smear(sw::vadd2(a, cs[g]));                       // f >= thr[0]: quantised for sure
				const unsigned sb = (w >> 15) & LM & 0x00010001u;                          // sign bit of the live samples
				const unsigned al = a & LM;
				const unsigned p0 = (al & 0xFFFFu) * iQ + 32768u;                          // (|c| * iQ + 32768) >> 16, :172
				const unsigned p1 = (al >> 16) * iQ2 + 65536u;                             // the same, already doubled
				const unsigned t = (p1 & 0xFFFE0000u) | ((p0 >> 15) & 0x0000FFFFu);
				const unsigned CM = LM & ~SM;                                              // rank candidates
				out[g] = (t & 0xFFFEFFFEu) | sb | (CM & 0x00020002u);
				accL = sw::vadd2(accL, LM);
				accS = sw::vadd2(accS, SM);
				const unsigned m = sw::vadd2(a, NT);                                      // |c| - T
				const unsigned f2 = sw::vadd2(sw::vadd2(m, m), sb) & CM;                  // f - 2T (candidates only)
				const unsigned pos = 4u * (unsigned)r + 2u * (unsigned)g;                  // tag | 15 - raster position, per half
				key[g] = ((f2 * 16u) | (0x800E800Fu - pos * 0x00010001u)) & CM;
			}
			io.put(r, make_uint2(out[0], out[1]));
			io.put_key(r, make_uint2(key[0], key[1]));
		}
	}
	const int nl = -((int)(short)(accL & 0xFFFF) + ((int)accL >> 16)), ns = -((int)(short)(accS & 0xFFFF) + ((int)accS >> 16));
	nc = nl - ns;
	return ns;
}

// Rank stage on the 16-bit keys: clears the halves of the rows that belong to dropped candidates, returns the survivors.
template <class IO>
RIC_HD int rank_rows_pk(const QuantBand *qb, int cnt, int ncm, IO &io)
{
	int s[16];
#pragma unroll
	for (int r = 0; r < 4; r++) {
		const uint2 k = io.get_key(r);
		s[4 * r] = (int)(k.x & 0xFFFFu); s[4 * r + 1] = (int)(k.x >> 16);
		s[4 * r + 2] = (int)(k.y & 0xFFFFu); s[4 * r + 3] = (int)(k.y >> 16);
	}
	int kstar = 0, m = 0;  // kstar: key of the last survivor (0: nobody survives)
	if (ncm == 1) {
		int s0 = 0;
#pragma unroll
		for (int k = 0; k < 16; k++) s0 = max(s0, s[k]);
		if (s0 >= qb->kt16[cnt]) { kstar = s0; m = 1; }  // kt16 > 0: an empty block (s0 == 0) never passes
	} else {
		sort16_desc(s);
		const int *kt = qb->kt16 + cnt;
#pragma unroll
		for (int i = 0; i < 16; i++) {
			if ((i & 3) == 0 && i >= ncm) break;  // warp-uniform
			const bool pass = s[i] >= kt[i];
			kstar = pass ? s[i] : kstar;
			m = pass ? i + 1 : m;
		}
	}
	// as signed half-words: candidates' keys are negative, everything else 0 -> (key - kstar) < 0 <=> dropped candidate
	const unsigned nk = (unsigned)((-kstar) & 0xFFFF) * 0x10001u;
#pragma unroll (IO::UNROLL == 2 ? 4 : 1)
	for (int r = 0; r < 4; r++) {
		const uint2 k = io.get_key(r);
		uint2 row = io.get(r);
		row.x &= ~sw::smear(sw::vadd2(k.x, nk));
		row.y &= ~sw::smear(sw::vadd2(k.y, nk));
		io.put(r, row);
	}
	return m;
}

}  // namespace ric
