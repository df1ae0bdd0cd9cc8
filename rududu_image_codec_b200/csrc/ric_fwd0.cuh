// ric_fwd0.cuh -- level 0 of the forward 9/7 transform from 8-bit pixels, two samples per register.
//
// Same job as fwd_level_kernel<true, T97, SRC_U8_*> (ric_fwd.cuh) -- colour transform / level shift
// (src/ric/ric.cpp:76-91,143-148), one level of CWavelet2D::Transform97 (src/lib/wavelet2d.cpp:320-359,
// 407-492), CBandCodec::buildTree of the level's D/H/V bands (src/lib/bandcodec.cpp:159-322) -- with the
// lifting done in the packed-linear arithmetic of ric_swar.cuh (two int16 per 32-bit register):
//
//  * One warp job = a strip of 240 columns x seg_rows rows of ALL planes of one image: the u8 pixels are
//    loaded once and the YCoCg chain runs once per pixel pair (the scalar kernel ran one plane per warp).
//  * Row pass: a register holds one column of the row pair (2t, 2t+1) the iteration brings in, so both
//    rows are lifted by the same instructions and the neighbours still come from warp shuffles.  Rows
//    lifted from 8-bit pixels cannot leave the int16 range: no checks needed.
//  * Column pass: registers are transposed (one PRMT each) to hold two columns of one row.  The 4-row
//    pipeline state of every plane lives in lane-private shared memory (no rotation moves, and the plane
//    loop stays rolled: the hot loop has to fit the instruction cache).  The packed arithmetic is only
//    valid while nothing wraps, so every iteration tests the operand bounds (ric_swar.cuh GUARD_*, a few
//    LOP3 and one vote); rows that fail -- and the first / last rows of the image, which use the
//    reference's edge formulas -- go through v_slow(): the scalar, exactly-wrapping steps of ric_dev.cuh.
//  * Finished band rows go to a 4-row lane-private ring per band; D/H and V block rows complete one
//    iteration apart and are quantised by flush_blocks_packed (ric_fwd.cuh) as soon as they do.
//
// Host side (ric_b200.cu launch_forward) uses this kernel when: 9/7, short level 0, more than one level,
// q != 0 (up-shifted pixels; q == 0 keeps the scalar kernel).
#pragma once
#include <cuda.h>  // CUtensorMap (type only: the encoder is fetched through the runtime, no libcuda link)

#include "ric_fwd.cuh"
#include "ric_swar.cuh"

namespace ric {

constexpr int F0_WARPS = 4;  // warps per CTA (independent jobs)
constexpr int F0_RR = 4;     // ring rows per band

struct F0Ring { uint2 v[3][F0_RR][32]; };  // finished band rows of one plane: [band][row & 3][lane]
template <int NP>
struct F0Smem {  // lane-private staging of one warp
	F0Ring ring[NP];
	union {
		uint4 stage[NP < 2 ? 2 : NP][2][32];  // converted input columns of every plane (8 registers per lane) ...
		KeyRows keys;                         // ... dead by the time the quantiser parks its candidate keys here
	};
	uint4 state[NP][4][32];  // column-pass pipeline: raw odd row, S1'd even row, S2'd odd row, S3'd even row
};

// ---- TMA experiment (gray level 0): one bulk tensor copy per warp and row pair instead of 64 per-lane loads ----
// Two 2-D tensor maps over the u8 source (width W, all image rows back to back, row pitch `pitch`), boxes of
// 256 x 2 and 16 x 2 bytes: the unit wants the box to start on a 16-byte boundary (a start at x0 - 8 traps with
// "illegal instruction", scripts/ubench/tma_probe.cu), so a strip's columns x0-8 .. x0+247 of rows 2t, 2t+1 come
// as the box at x0-16 plus a 16-byte box at x0+240.  Columns left of 0 / right of W-1 are zero-filled by the unit
// (what the `col_ok ? load : 0` predicates did).  Lane 0 arms an mbarrier with the byte count and issues the
// copies; every lane waits on the barrier and reads its 8 bytes per row from shared memory.
struct F0Tma {
	alignas(128) unsigned char tile[2][640];  // double-buffered: [stage]: rows of box A [2][256], rows of box B [2][16], pad
	unsigned long long bar[2];
};
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f0_mbar_init(unsigned long long *bar)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void f0_tma_issue(const CUtensorMap *mapA, const CUtensorMap *mapB, unsigned char *dst, unsigned long long *bar,
                                             int x0, int y)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 544;" ::"r"(smem_u32(bar)) : "memory");
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
	             ::"r"(smem_u32(dst)), "l"(mapA), "r"(x0 - 16), "r"(y), "r"(smem_u32(bar))
	             : "memory");
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
	             ::"r"(smem_u32(dst + 512)), "l"(mapB), "r"(x0 + 240), "r"(y), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void f0_mbar_wait(unsigned long long *bar, unsigned parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "RIC_F0_WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra RIC_F0_DONE_%=;\n"
	    "bra RIC_F0_WAIT_%=;\n"
	    "RIC_F0_DONE_%=:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}

__host__ __device__ constexpr size_t f0_qb_bytes() { return (sizeof(QuantBand) * 6 + 15) & ~(size_t)15; }
template <int NP>
__host__ __device__ constexpr size_t f0_smem_bytes() { return f0_qb_bytes() + F0_WARPS * sizeof(F0Smem<NP>); }
__host__ __device__ constexpr size_t f0_tma_off() { return (f0_smem_bytes<1>() + 127) & ~(size_t)127; }
__host__ __device__ constexpr size_t f0_smem_bytes_tma() { return f0_tma_off() + F0_WARPS * sizeof(F0Tma); }

// ---- pixels -> row-pass input ---------------------------------------------------------------------------
// X[c] = (row 2t | row 2t+1 << 16) of column c, constants KH::E0 (even c) / KH::O0 (odd c).
__device__ __forceinline__ void f0_convert_gray(const uint2 e, const uint2 o, unsigned (&X)[8])
{
#pragma unroll
	for (int c = 0; c < 8; c += 2) {
		const unsigned w = sw::prmt(c < 4 ? e.x : e.y, c < 4 ? o.x : o.y, (c & 2) ? 0x7632u : 0x5410u);  // e[c], e[c+1], o[c], o[c+1]
		const unsigned p0 = sw::prmt(w, 0u, 0x4240u), p1 = sw::prmt(w, 0u, 0x4341u);
		X[c] = p0 * 16u + (sw::KH::E0 - 128u * 16u * 0x10001u);      // (p - 128) << 4, ric.cpp:147
		X[c + 1] = p1 * 16u + (sw::KH::O0 - 128u * 16u * 0x10001u);
	}
}

// RGBtoYCoCg<4> (ric.cpp:76-91) on pixel pairs; every intermediate is kept non-negative by a bias that the
// byte permute puts into the upper byte of each half for free.  Xp[0] Co, Xp[1] Cg, Xp[2] Y.
__device__ __forceinline__ void f0_convert_rgb(const uint2 (&e)[3], const uint2 (&o)[3], unsigned (&Xp)[3][8])
{
	const unsigned BIAS = 0x03020100u;  // byte 4 + k of the permute = k
#pragma unroll
	for (int c = 0; c < 8; c += 2) {
		const unsigned sel = (c & 2) ? 0x7632u : 0x5410u;
		const unsigned wr = sw::prmt(c < 4 ? e[0].x : e[0].y, c < 4 ? o[0].x : o[0].y, sel);
		const unsigned wg = sw::prmt(c < 4 ? e[1].x : e[1].y, c < 4 ? o[1].x : o[1].y, sel);
		const unsigned wb = sw::prmt(c < 4 ? e[2].x : e[2].y, c < 4 ? o[2].x : o[2].y, sel);
#pragma unroll
		for (int k = 0; k < 2; k++) {
			const unsigned r = sw::prmt(wr, BIAS, k ? 0x5351u : 0x5250u);  // R + 256
			const unsigned g = sw::prmt(wg, BIAS, k ? 0x7371u : 0x7270u);  // G + 768
			const unsigned b = sw::prmt(wb, BIAS, k ? 0x4341u : 0x4240u);  // B
			const unsigned co = r - b;                 // Co + 256          Co = R - B
			const unsigned t = b + sw::lsr<1>(co);     // t + 128           t = B + (Co >> 1)
			const unsigned cg = g - t;                 // Cg + 640          Cg = G - t
			const unsigned y = t + sw::lsr<1>(cg);     // Y + 576           Y = t + (Cg >> 1) - 128
			const unsigned K = k ? sw::KH::O0 : sw::KH::E0;
			Xp[0][c + k] = co * 8u + (K - 256u * 8u * 0x10001u);   // Co <<= 3
			Xp[1][c + k] = cg * 8u + (K - 640u * 8u * 0x10001u);   // Cg <<= 3
			Xp[2][c + k] = y * 16u + (K - 576u * 16u * 0x10001u);  // Y <<= 4
		}
	}
}

// ---- row pass on the 8 columns of a lane (cf. row_fwd in ric_dev.cuh) -----------------------------------
template <bool EDGE>
__device__ __forceinline__ void f0_row_pass(unsigned (&X)[8], const EdgeX &ee)
{
	using namespace sw;
	EdgeX e = ee;
	if (!EDGE) e.on = false;
	unsigned nl, nr, o[8];
	nl = __shfl_up_sync(FULL, X[7], 1);
#pragma unroll
	for (int k = 0; k < 8; k += 2) { o[k] = X[k]; X[k] = s1<KH>(X[k], k ? X[k - 1] : nl, X[k + 1]); }
	if (e.on) {
		if (e.first) X[0] = s1_edge<KH>(o[0], X[1]);
		if (e.last) {
#pragma unroll
			for (int k = 0; k < 8; k += 2) if (e.kl == k) X[k] = s1_edge<KH>(o[k], k ? X[k - 1] : nl);
		}
	}
	nr = __shfl_down_sync(FULL, X[0], 1);
#pragma unroll
	for (int k = 1; k < 8; k += 2) { o[k] = X[k]; X[k] = s2<KH>(X[k], X[k - 1], k < 7 ? X[k + 1] : nr); }
	if (e.on && e.last) {
#pragma unroll
		for (int k = 1; k < 8; k += 2) if (e.kl == k) X[k] = s2_last<KH>(o[k], X[k - 1]);
	}
	nl = __shfl_up_sync(FULL, X[7], 1);
#pragma unroll
	for (int k = 0; k < 8; k += 2) { o[k] = X[k]; X[k] = s3<KH>(X[k], k ? X[k - 1] : nl, X[k + 1]); }
	if (e.on) {
		if (e.first) X[0] = s3_edge<KH>(o[0], X[1]);
		if (e.last) {
#pragma unroll
			for (int k = 0; k < 8; k += 2) if (e.kl == k) X[k] = s3_edge<KH>(o[k], k ? X[k - 1] : nl);
		}
	}
	nr = __shfl_down_sync(FULL, X[0], 1);
#pragma unroll
	for (int k = 1; k < 8; k += 2) { o[k] = X[k]; X[k] = s4<KH>(X[k], X[k - 1], k < 7 ? X[k + 1] : nr); }
	if (e.on && e.last) {
#pragma unroll
		for (int k = 1; k < 8; k += 2) if (e.kl == k) X[k] = s4_last<KH>(o[k], X[k - 1]);
	}
}

// ---- scalar column-pass iteration (edge rows, operands outside the packed bounds) ---------------------
// scr: the lane's 8 staging words (new even row in 0-3, new odd row in 4-7, column-pass constants);
// st: the lane's pipeline state.  Leaves the finished even / odd row in scr as two's-complement pairs,
// the new state in st, and returns whether the new state is inside the bounds the packed path assumes.
__device__ __noinline__ bool f0_v_slow(uint4 *scr, uint4 *st, int t, int h)
{
	using namespace sw;
	int ne[8], no[8], so1[8], se1[8], so2[8], se3[8];
	{
		const uint4 a = scr[0], b = scr[32], s0 = st[0], s1_ = st[32], s2_ = st[64], s3_ = st[96];
		const unsigned A[4] = {a.x, a.y, a.z, a.w}, B[4] = {b.x, b.y, b.z, b.w};
		const unsigned S0[4] = {s0.x, s0.y, s0.z, s0.w}, S1[4] = {s1_.x, s1_.y, s1_.z, s1_.w};
		const unsigned S2[4] = {s2_.x, s2_.y, s2_.z, s2_.w}, S3[4] = {s3_.x, s3_.y, s3_.z, s3_.w};
#pragma unroll
		for (int i = 0; i < 4; i++) {
			ne[2 * i] = dec_lo(A[i], KV::E0); ne[2 * i + 1] = dec_hi(A[i], KV::E0);
			no[2 * i] = dec_lo(B[i], KV::O0); no[2 * i + 1] = dec_hi(B[i], KV::O0);
			so1[2 * i] = dec_lo(S0[i], KV::O0); so1[2 * i + 1] = dec_hi(S0[i], KV::O0);
			se1[2 * i] = dec_lo(S1[i], KV::E1); se1[2 * i + 1] = dec_hi(S1[i], KV::E1);
			so2[2 * i] = dec_lo(S2[i], KV::O2); so2[2 * i + 1] = dec_hi(S2[i], KV::O2);
			se3[2 * i] = dec_lo(S3[i], KV::E3); se3[2 * i + 1] = dec_hi(S3[i], KV::E3);
		}
	}
	const int r1 = 2 * t, r2 = 2 * t - 1, r3 = 2 * t - 2, r4 = 2 * t - 3;
	if (r1 >= 0 && r1 < h) vS1<true, T97, true>(ne, so1, no, r1 == 0, r1 == h - 1);
	if (r2 >= 0 && r2 < h) vS2<true, T97, true>(so1, se1, ne, false, r2 == h - 1);
	if (r3 >= 0 && r3 < h) vS3<true, T97, true>(se1, so2, so1, r3 == 0, r3 == h - 1);
	if (r4 >= 0 && r4 < h) vS4<true, T97, true>(so2, se3, se1, false, r4 == h - 1);
	bool ok = true;
	unsigned E[4], O[4], N0[4], N1[4], N2[4], N3[4];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		// S3 / S4 results arrive un-truncated (ric_dev.cuh): the (short) casts are their C-typed stores
		const int e0 = (short)se1[2 * i], e1 = (short)se1[2 * i + 1], o0 = (short)so2[2 * i], o1 = (short)so2[2 * i + 1];
		E[i] = (unsigned)(e0 & 0xFFFF) | ((unsigned)e1 << 16);
		O[i] = (unsigned)(o0 & 0xFFFF) | ((unsigned)o1 << 16);
		const int a0 = (short)no[2 * i], a1 = (short)no[2 * i + 1], b0 = (short)ne[2 * i], b1 = (short)ne[2 * i + 1];
		const int c0 = (short)so1[2 * i], c1 = (short)so1[2 * i + 1];
		N0[i] = enc(a0, a1, KV::O0); N1[i] = enc(b0, b1, KV::E1); N2[i] = enc(c0, c1, KV::O2); N3[i] = enc(e0, e1, KV::E3);
		ok = ok && abs(a0) <= BOUND_O0 && abs(a1) <= BOUND_O0 && abs(b0) <= BOUND_E1 && abs(b1) <= BOUND_E1 &&
		     abs(c0) <= BOUND_O2 && abs(c1) <= BOUND_O2 && abs(e0) <= BOUND_E3 && abs(e1) <= BOUND_E3;
	}
	scr[0] = make_uint4(E[0], E[1], E[2], E[3]);
	scr[32] = make_uint4(O[0], O[1], O[2], O[3]);
	st[0] = make_uint4(N0[0], N0[1], N0[2], N0[3]);
	st[32] = make_uint4(N1[0], N1[1], N1[2], N1[3]);
	st[64] = make_uint4(N2[0], N2[1], N2[2], N2[3]);
	st[96] = make_uint4(N3[0], N3[1], N3[2], N3[3]);
	return ok;
}

// Strips that hold column 0 or column w-1 (2 of 16 at 4K): the row pass with the reference's edge formulas, and
// the columns outside the image zeroed (they hold garbage, and the packed column pass must stay in range).
// Out of line: the hot loop must stay small (instruction cache), and this is the rarer case.
__device__ __noinline__ void f0_row_pass_edge(unsigned (&X)[8], int cb, int w, int nvalid)
{
	using namespace sw;
	const EdgeX ex = make_edge_x(cb, w, true);
	f0_row_pass<true>(X, ex);
#pragma unroll
	for (int k = 0; k < 8; k++) X[k] = k < nvalid ? X[k] : ((k & 1) ? KH::O4 : KH::E3);
}

// One job = one (image, row segment, strip), all NP planes.
template <int NP, bool TMA>
__device__ __forceinline__ void fwd0_job(const FwdParams &P, int img, int sy, int sx, bool EDGE, F0Smem<NP> &sm,
                                         const QuantBand (*s_qb)[3], int lane, const CUtensorMap *tmap, const CUtensorMap *tmapB, F0Tma *tm,
                                         unsigned &tma_par)
{
	using namespace sw;
	const int w = P.w, h = P.h;
	const int x0 = sx * STRIP_W;
	const int cb = x0 - LANE_W + lane * LANE_W;
	const bool col_ok = cb >= 0 && cb < w;
	const bool lane_out = lane >= 1 && lane <= 30;
	const int nvalid = cb < 0 ? 0 : min(8, max(0, w - cb));  // real columns of this lane (EDGE strips only)
	const int y0 = sy * P.seg_rows;
	const int y1 = min(h, y0 + P.seg_rows);
	const int y1r = (y1 + 7) & ~7;
	const unsigned char *src = (const unsigned char *)P.src + img * P.src_img_stride + cb;
	const int bx = cb >> 3, lx = cb >> 1;
	const int ll_dimx = w >> 1;
	short *ll_base = (short *)P.ll + img * P.ll_img_stride;

	// pipeline state: the representation of zero in every role
#pragma unroll
	for (int p = 0; p < NP; p++) {
		sm.state[p][0][lane] = make_uint4(KV::O0, KV::O0, KV::O0, KV::O0);
		sm.state[p][1][lane] = make_uint4(KV::E1, KV::E1, KV::E1, KV::E1);
		sm.state[p][2][lane] = make_uint4(KV::O2, KV::O2, KV::O2, KV::O2);
		sm.state[p][3][lane] = make_uint4(KV::E3, KV::E3, KV::E3, KV::E3);
	}
	unsigned state_ok = (1u << NP) - 1u;  // bit p: plane p's pipeline state is inside the packed path's bounds

	const int t_begin = (y0 >> 1) - 2, t_last = (y1r >> 1) + 1;
	uint2 rawE[NP], rawO[NP];
	auto load_rows = [&](int t) {
		if constexpr (TMA) {  // (NP == 1) issue the bulk copy of rows 2t, 2t+1 into stage (t & 1); read back by take_rows
			__syncwarp();     // every lane is done with the stage this copy overwrites
			if (lane == 0) f0_tma_issue(tmap, tmapB, tm->tile[t & 1], &tm->bar[t & 1], x0, img * h + 2 * t);
			return;
		}
		const int re = 2 * t, ro = re + 1;
		const bool oke = col_ok && re >= 0 && re < h, oko = col_ok && ro >= 0 && ro < h;
#pragma unroll
		for (int p = 0; p < NP; p++) {
			rawE[p] = make_uint2(0u, 0u); rawO[p] = make_uint2(0u, 0u);
			if (oke) rawE[p] = __ldg((const uint2 *)(src + p * P.src_plane_stride + (long long)re * P.src_pitch));
			if (oko) rawO[p] = __ldg((const uint2 *)(src + p * P.src_plane_stride + (long long)ro * P.src_pitch));
		}
	};
	auto take_rows = [&](int t) {
		if constexpr (TMA) {
			const int st = t & 1;
			f0_mbar_wait(&tm->bar[st], (tma_par >> st) & 1u);
			tma_par ^= 1u << st;
			// box A holds columns x0-16 .. x0+239, this lane's start at x0-8+8*lane; lane 31's eight are box B's first
			const unsigned char *rowE = lane < 31 ? tm->tile[st] + 8 + lane * 8 : tm->tile[st] + 512;
			rawE[0] = *(const uint2 *)rowE;
			rawO[0] = *(const uint2 *)(rowE + (lane < 31 ? 256 : 16));
		}
	};
	load_rows(t_begin);

#pragma unroll 1
	for (int t = t_begin; t <= t_last; t++) {
		take_rows(t);
		{  // pixels -> row-pass input of every plane, staged in shared memory for the rolled plane loop
			if constexpr (NP == 3) {
				unsigned Xp[3][8];
				f0_convert_rgb(rawE, rawO, Xp);
#pragma unroll
				for (int p = 0; p < 3; p++) {
					sm.stage[p][0][lane] = make_uint4(Xp[p][0], Xp[p][1], Xp[p][2], Xp[p][3]);
					sm.stage[p][1][lane] = make_uint4(Xp[p][4], Xp[p][5], Xp[p][6], Xp[p][7]);
				}
			} else {
				unsigned X[8];
				f0_convert_gray(rawE[0], rawO[0], X);
				sm.stage[0][0][lane] = make_uint4(X[0], X[1], X[2], X[3]);
				sm.stage[0][1][lane] = make_uint4(X[4], X[5], X[6], X[7]);
			}
		}
		load_rows(t + 1);  // prefetch

		const int r1 = 2 * t, r4 = 2 * t - 3;
		const bool interior = !((r4 - 1 <= 0) || (r1 + 1 >= h - 1));
		const int jd = t - 1, jv = t - 2;  // band rows that finish in this iteration: D/H row jd, V/LL row jv

#pragma unroll 1
		for (int p = 0; p < NP; p++) {
			unsigned X[8];
			{
				const uint4 a = sm.stage[p][0][lane], b = sm.stage[p][1][lane];
				X[0] = a.x; X[1] = a.y; X[2] = a.z; X[3] = a.w; X[4] = b.x; X[5] = b.y; X[6] = b.z; X[7] = b.w;
			}
			// the plane's pipeline state is fetched now, so that the row pass hides the shared-memory latency
			uint4 *st = &sm.state[p][0][lane];
			const uint4 q0 = st[0], q1 = st[32], q2 = st[64], q3 = st[96];
			if (EDGE) {  // warp-uniform; a copy goes through the call so that X itself stays in registers
				unsigned Y[8];
#pragma unroll
				for (int k = 0; k < 8; k++) Y[k] = X[k];
				f0_row_pass_edge(Y, cb, w, nvalid);
#pragma unroll
				for (int k = 0; k < 8; k++) X[k] = Y[k];
			} else {
				EdgeX none;
				none.on = false;
				f0_row_pass<false>(X, none);
			}
			// transpose: two columns of one row per register (even columns first: D / V, then odd: H / LL)
			unsigned NE[4], NO[4];
			NE[0] = prmt(X[0], X[2], 0x5410u) + (KV::E0 - KH::E3); NO[0] = prmt(X[0], X[2], 0x7632u) + (KV::O0 - KH::E3);
			NE[1] = prmt(X[4], X[6], 0x5410u) + (KV::E0 - KH::E3); NO[1] = prmt(X[4], X[6], 0x7632u) + (KV::O0 - KH::E3);
			NE[2] = prmt(X[1], X[3], 0x5410u) + (KV::E0 - KH::O4); NO[2] = prmt(X[1], X[3], 0x7632u) + (KV::O0 - KH::O4);
			NE[3] = prmt(X[5], X[7], 0x5410u) + (KV::E0 - KH::O4); NO[3] = prmt(X[5], X[7], 0x7632u) + (KV::O0 - KH::O4);

			unsigned E3[4], O4[4];  // finished even row 2t-2 / odd row 2t-3 as two's-complement pairs
			bool fast = false;
			unsigned E1[4];
			const unsigned so1[4] = {q0.x, q0.y, q0.z, q0.w}, se1[4] = {q1.x, q1.y, q1.z, q1.w};
			if (interior && ((state_ok >> p) & 1u)) {
#pragma unroll
				for (int i = 0; i < 4; i++) E1[i] = s1<KV>(NE[i], so1[i], NO[i]);
				const unsigned go = (NO[0] | NO[1] | NO[2] | NO[3]) & GUARD_O, ge = (E1[0] | E1[1] | E1[2] | E1[3]) & GUARD_E;
				fast = __all_sync(FULL, (go | ge) == 0);
			}
			if (fast) {
				const unsigned so2[4] = {q2.x, q2.y, q2.z, q2.w}, se3[4] = {q3.x, q3.y, q3.z, q3.w};
				unsigned O2[4], E3k[4];
#pragma unroll
				for (int i = 0; i < 4; i++) {
					O2[i] = s2<KV>(so1[i], se1[i], E1[i]);
					E3k[i] = s3<KV>(se1[i], so2[i], O2[i]);
					const unsigned o4 = s4<KV>(so2[i], se3[i], E3k[i]);
					E3[i] = to_c2(E3k[i], KV::E3);
					O4[i] = to_c2(o4, KV::O4);
				}
				st[0] = make_uint4(NO[0], NO[1], NO[2], NO[3]);
				st[32] = make_uint4(E1[0], E1[1], E1[2], E1[3]);
				st[64] = make_uint4(O2[0], O2[1], O2[2], O2[3]);
				st[96] = make_uint4(E3k[0], E3k[1], E3k[2], E3k[3]);
			} else {
				uint4 *scr = &sm.stage[p][0][lane];
				scr[0] = make_uint4(NE[0], NE[1], NE[2], NE[3]);
				scr[32] = make_uint4(NO[0], NO[1], NO[2], NO[3]);
				const bool ok = f0_v_slow(scr, st, t, h);
				state_ok = (state_ok & ~(1u << p)) | ((unsigned)__all_sync(FULL, ok) << p);
				const uint4 a = scr[0], b = scr[32];
				E3[0] = a.x; E3[1] = a.y; E3[2] = a.z; E3[3] = a.w;
				O4[0] = b.x; O4[1] = b.y; O4[2] = b.z; O4[3] = b.w;
			}
			if (P.stats && lane == 0) { atomicAdd(P.stats, 1ull); if (!fast) atomicAdd(P.stats + 1, 1ull); }
			sm.ring[p].v[0][jd & (F0_RR - 1)][lane] = make_uint2(E3[0], E3[1]);
			sm.ring[p].v[1][jd & (F0_RR - 1)][lane] = make_uint2(E3[2], E3[3]);
			sm.ring[p].v[2][jv & (F0_RR - 1)][lane] = make_uint2(O4[0], O4[1]);
			if (jv >= (y0 >> 1) && jv < (y1 >> 1) && lane_out && lx < ll_dimx)  // LL row jv (scratch rows are padded)
				*(uint2 *)(ll_base + p * P.ll_plane_stride + (long long)jv * P.ll_pitch + lx) = make_uint2(O4[2], O4[3]);
		}
		// a block row of D / H is complete one iteration before the same block row of V
		if ((jd & 3) == 3 || (jv & 3) == 3) {
			const bool dh = (jd & 3) == 3;
			const int by = dh ? jd >> 2 : jv >> 2;
			if (by >= (y0 >> 3)) {
#pragma unroll 1
				for (int p = 0; p < NP; p++) {
					char *arena = P.arena + img * P.arena_img_stride + p * P.arena_plane_stride;
					unsigned char *flags = P.flags + img * P.flags_img_stride + p * P.flags_plane_stride;
					flush_blocks_packed<F0_RR, false>(P, sm.ring[p], sm.keys, arena, flags, &s_qb[P.plane_class[p]][0], bx, by, lane, lane_out,
					                           dh ? 0 : 2, dh ? 2 : 3);
				}
			}
		}
	}
	take_rows(t_last + 1);  // (TMA) the last prefetch is still in flight: its stage must be free before the next job
}

template <int NP, bool TMA>
__device__ __forceinline__ void fwd0_body(const FwdParams &P, const CUtensorMap *tmap, const CUtensorMap *tmapB)
{
	extern __shared__ __align__(128) unsigned char f0_smem[];
	QuantBand(*s_qb)[3] = (QuantBand(*)[3])f0_smem;
	for (int i = threadIdx.x; i < (int)(sizeof(QuantBand) * 6 / 4); i += blockDim.x) ((int *)f0_smem)[i] = ((const int *)P.qb)[i];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	F0Smem<NP> &sm = *((F0Smem<NP> *)(f0_smem + f0_qb_bytes()) + wib);
	F0Tma *tm = TMA ? (F0Tma *)(f0_smem + f0_tma_off()) + wib : nullptr;
	unsigned tma_par = 0;
	if (TMA && lane == 0) {
		f0_mbar_init(&tm->bar[0]);
		f0_mbar_init(&tm->bar[1]);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	const unsigned njobs = (unsigned)P.nstrips * P.nsegs * P.nimages;
	for (;;) {
		unsigned long long j64 = 0;
		if (lane == 0) j64 = atomicAdd(P.counter, 1ull);
		j64 = __shfl_sync(FULL, j64, 0);
		if (j64 >= njobs) break;
		unsigned job = (unsigned)j64;
		const int sx = (int)(job % (unsigned)P.nstrips); job /= (unsigned)P.nstrips;
		const int sy = (int)(job % (unsigned)P.nsegs);
		const int img = (int)(job / (unsigned)P.nsegs);
		const int x0 = sx * STRIP_W;
		fwd0_job<NP, TMA>(P, img, sy, sx, (x0 == 0) || (P.w <= x0 + STRIP_W + LANE_W), sm, s_qb, lane, tmap, tmapB, tm, tma_par);
	}
}

template <int NP>
__global__ void __launch_bounds__(F0_WARPS * 32, NP == 3 ? 3 : 4) fwd0_kernel(const __grid_constant__ FwdParams P)
{
	fwd0_body<NP, false>(P, nullptr, nullptr);
}

// gray only: the source rows come through the TMA unit (RIC_TMA=1, with RIC_FWD0=1)
__global__ void __launch_bounds__(F0_WARPS * 32, 4) fwd0_tma_kernel(const __grid_constant__ FwdParams P, const __grid_constant__ CUtensorMap tmap,
                                                                    const __grid_constant__ CUtensorMap tmapB)
{
	fwd0_body<1, true>(P, &tmap, &tmapB);
}

}  // namespace ric
