// ric_b200.cu -- C ABI (include/ric_b200.h) over the sm_100a level kernels.
//
// There is no CPU implementation of the hot path in this library: every entry point that computes
// launches CUDA kernels, and ric_create fails with RIC_E_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "ric_entropy.h"
#include "ric_entropy_gpu.h"
#include "ric_fwd.cuh"
#include "ric_fwd0.cuh"
#include "ric_host.h"
#include "ric_landed.h"
#include "ric_inv.cuh"
#include "ric_inv0.cuh"

using namespace ric;

static thread_local char g_err[512] = "";

static int getenv(const char *name, int dflt)
{
	const char *v = ::getenv(name);
	return v ? atoi(v) : dflt;
}


static int set_err(int code, const char *fmt, const char *a = "", const char *b = "")
{
	snprintf(g_err, sizeof g_err, fmt, a, b);
	return code;
}

#define CK(call)                                                                               \
	do {                                                                                       \
		cudaError_t e_ = (call);                                                               \
		if (e_ != cudaSuccess) return set_err(RIC_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
	} while (0)

struct ric_ctx {
	HostGeom g;
	int device, max_batch;
	int sm_count;
	cudaStream_t stream;
	cudaStream_t pipe[3];    // host-buffer calls: chunks of the batch alternate over these (copy/compute overlap)
	struct ChunkNote { ric_chunk_fn fn; void *user; int first, count; };
	std::vector<ChunkNote> notes;  // one per chunk of the last *_stream call (must outlive the callbacks)
	int img0;                // first image slot used by the launch functions (chunked pipelining)
	// device buffers
	unsigned char *d_src;   // [max_batch][channels][height][src_pitch] u8
	size_t src_pitch;
	char *d_arena;          // [max_batch][channels][arena_bytes]  encode output (padding columns stay zero)
	char *d_arena_in;       // same size, decode-side input (allocated on first use)
	HostGeom *d_geom;       // device copies for the GPU entropy stage (allocated on first use)
	void *d_tables;         // ent::Tables, a POD block
	int *d_bad;
	void *d_hints;          // block hints of the device entropy encoder: [max_batch][channels][flag_bytes] x 8 bytes
	uint8_t *d_payload;     // [max_batch][payload_stride] device-side .ric payload slots
	size_t payload_stride;
	long long *d_psizes, *h_psizes;  // payload lengths, device and pinned host
	cudaStream_t ent[8];    // one per chunk: the per-image-serial entropy kernels of different chunks run side by side
	cudaEvent_t ent_ev[8];
	char *h_stage;          // pinned [max_batch][channels][arena_bytes]: ric_compress_u8 / ric_decompress_u8 staging
	unsigned char *d_flags; // [max_batch][channels][flag_bytes]
	short *d_plane;         // [channels][height][plane_pitch] s16 (plane-level API, slot 0)
	int plane_pitch;
	void *d_ll[RIC_MAX_LEVELS];  // LL scratch between level i and i+1: [max_batch][channels][lev_h[i+1]][ll_pitch[i]]
	int ll_pitch[RIC_MAX_LEVELS];
	int ll_es[RIC_MAX_LEVELS];
	unsigned *d_count;
	unsigned long long *d_stats;         // path statistics of the packed kernels (filled while profiling is on)
	unsigned long long *d_jobctr;        // job counters of the persistent kernels: [4 stream sets][2 directions][levels]
	int cset;                            // counter set in use (0-2: internal pipeline streams, 3: caller's stream)
	int launches;
	int profiling;                       // record CUDA events around every level launch
	cudaEvent_t ev[2][RIC_MAX_LEVELS + 1];  // [direction][launch boundary]
	int ev_n[2];
	int target_warps;
	int use_tma;
	int use_pdl;            // programmatic dependent launch between the level kernels of one call (RIC_PDL, default 1)
	int use_inv0;
	int use_fwd0;                        // packed level-0 forward kernel, experimental (RIC_FWD0=1)
};

// ---------------------------------------------------------------------------------------------
// kernel dispatch tables

// Level kernels of one call are launched back to back on one stream; with programmatic dependent launch the next
// level's CTAs become resident while the previous level drains and wait in griddepcontrol.wait (top of the kernels)
// until its memory is visible: the launch latency and the kernel preamble leave the critical path of a single image.
template <class P>
static cudaError_t launch_level(void (*fn)(const P), unsigned grid, unsigned block, cudaStream_t st, const P &params, bool pdl)
{
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, fn, params);
}

typedef void (*fwd_fn)(const FwdParams);
typedef void (*inv_fn)(const InvParams);

static fwd_fn pick_fwd(bool sh, int trans, int src)
{
#define RIC_F(SHV, TR_, SRC_) if (sh == SHV && trans == (TR_ == T97 ? RIC_CDF97 : TR_ == T53 ? RIC_CDF53 : RIC_HAAR) && src == SRC_) return fwd_level_kernel<SHV, TR_, SRC_>;
	RIC_F(true, T97, SRC_U8_GRAY) RIC_F(true, T97, SRC_U8_RGB) RIC_F(true, T97, SRC_S16)
	RIC_F(false, T97, SRC_S16) RIC_F(false, T97, SRC_S32)
	RIC_F(true, T53, SRC_U8_GRAY) RIC_F(true, T53, SRC_U8_RGB) RIC_F(true, T53, SRC_S16)
	RIC_F(false, T53, SRC_S16) RIC_F(false, T53, SRC_S32)
	RIC_F(true, THAAR, SRC_U8_GRAY) RIC_F(true, THAAR, SRC_U8_RGB) RIC_F(true, THAAR, SRC_S16)
	RIC_F(false, THAAR, SRC_S16) RIC_F(false, THAAR, SRC_S32)
#undef RIC_F
	return nullptr;
}

static inv_fn pick_inv(bool sh, int trans, int dst)
{
#define RIC_I(SHV, TR_, DST_) if (sh == SHV && trans == (TR_ == T97 ? RIC_CDF97 : TR_ == T53 ? RIC_CDF53 : RIC_HAAR) && dst == DST_) return inv_level_kernel<SHV, TR_, DST_>;
	RIC_I(true, T97, DST_PLANE) RIC_I(true, T97, DST_U8_GRAY) RIC_I(true, T97, DST_U8_RGB) RIC_I(false, T97, DST_PLANE)
	RIC_I(true, T53, DST_PLANE) RIC_I(true, T53, DST_U8_GRAY) RIC_I(true, T53, DST_U8_RGB) RIC_I(false, T53, DST_PLANE)
	RIC_I(true, THAAR, DST_PLANE) RIC_I(true, THAAR, DST_U8_GRAY) RIC_I(true, THAAR, DST_U8_RGB) RIC_I(false, THAAR, DST_PLANE)
#undef RIC_I
	return nullptr;
}

// ---------------------------------------------------------------------------------------------
// standalone kernels behind the class-level API (CBandCodec::buildTree on resident bands,
// CWavelet2D::TSUQ / TSUQi on all bands)

struct QuantLevelParams {
	char *arena;
	unsigned char *flags;
	BandRef band[3], child[3];
	int has_child, is_int;
	int o_first;  // blockIdx.y + o_first = orientation (single-band launches of ric_buf_build_tree)
	QuantBand qb[3];
};

template <bool SH>
__global__ void quant_level_kernel(const __grid_constant__ QuantLevelParams P)
{
	__shared__ QuantBand s_qb[3];
	for (int i = threadIdx.x; i < (int)(sizeof(s_qb) / 4); i += blockDim.x) ((int *)s_qb)[i] = ((const int *)P.qb)[i];
	__syncthreads();
	const int o = blockIdx.y + P.o_first;
	const BandRef &b = P.band[o];
	const int nbx = b.fl_bw, nby = (b.dimy + 3) / 4;
	const int id = blockIdx.x * blockDim.x + threadIdx.x;
	const bool have = id < nbx * nby;  // threads past the last block run the quantiser on an empty block (full-warp votes inside)
	const int bx = have ? id % nbx : 0, by = have ? id / nbx : 0;
	const int bw = have ? min(4, b.dimx - 4 * bx) : 0, bh = have ? min(4, b.dimy - 4 * by) : 0;
	constexpr int ES = SH ? 2 : 4;
	char *base = P.arena + b.off;
	int c[16];
#pragma unroll
	for (int k = 0; k < 16; k++) {
		const int r = k >> 2, x = k & 3;
		c[k] = 0;
		if (r < bh && x < bw) {
			const char *p = base + ((long long)(4 * by + r) * b.stride + 4 * bx + x) * ES;
			c[k] = SH ? (int)*(const short *)p : *(const int *)p;
		}
	}
	int nz = quant_block<SH>(c, &s_qb[o], bw, bh);
	if (!have) return;
	if (P.has_child && bw == 4 && bh == 4) {
		const BandRef &ch = P.child[o];
		const unsigned char *cf = P.flags + ch.fl_off + (2 * by) * ch.fl_bw + 2 * bx;
		nz += cf[0] + cf[1] + cf[ch.fl_bw] + cf[ch.fl_bw + 1];
	}
	P.flags[b.fl_off + by * b.fl_bw + bx] = nz != 0;
	if (nz == 0) c[0] = -0x8000;
#pragma unroll
	for (int k = 0; k < 16; k++) {
		const int r = k >> 2, x = k & 3;
		if (r < bh && x < bw) {
			char *p = base + ((long long)(4 * by + r) * b.stride + 4 * bx + x) * ES;
			if (SH) *(short *)p = (short)c[k]; else *(int *)p = c[k];
		}
	}
}

// mode 0: TSUQ (dead-zone quantise, count non-zeros); mode 1: TSUQi (multiply).  count[0] += Count; count[1] /
// count[2] (when count is given) = min / max of the quantised values over 0, as CBand::TSUQ keeps Min / Max.
template <bool SH>
__global__ void band_pointwise_kernel(char *arena, long long off, int dimx, int dimy, int stride, int mode, int Q,
                                      int iQ, int T, unsigned *count)
{
	const long long n = (long long)dimx * dimy;
	unsigned local = 0;
	int lmin = 0, lmax = 0;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const int y = (int)(i / dimx), x = (int)(i % dimx);
		char *p = arena + off + ((long long)y * stride + x) * (SH ? 2 : 4);
		int c = SH ? (int)*(short *)p : *(int *)p;
		if (mode == 0) {
			// Count counts every coefficient outside the dead zone (band.h:77-80), even if it rounds to 0
			local += !((unsigned)(c + T) <= (unsigned)(2 * T));
			c = tsuq1<SH>(c, T, iQ);
			lmin = min(lmin, c);
			lmax = max(lmax, c);
		} else {
			c = TR<SH>(c * Q);
		}
		if (SH) *(short *)p = (short)c; else *(int *)p = c;
	}
	if (mode == 0 && count) {
		local = __reduce_add_sync(0xffffffffu, local);
		lmin = __reduce_min_sync(0xffffffffu, lmin);
		lmax = __reduce_max_sync(0xffffffffu, lmax);
		if ((threadIdx.x & 31) == 0) {
			if (local) atomicAdd(count, local);
			if (lmin < 0) atomicMin((int *)count + 1, lmin);
			if (lmax > 0) atomicMax((int *)count + 2, lmax);
		}
	}
}

// ---------------------------------------------------------------------------------------------

static void fill_qb(QuantBand &q, int Quant, int lambda, float weight, int is_int)
{
	HostQuantBand hq;
	make_quant_band(hq, Quant, lambda, weight, is_int);
	q.Q = hq.Q; q.iQ = hq.iQ; q.T = hq.T; q.Te = hq.Te;
	q.fast = hq.Q >= 1 && hq.Q <= (is_int ? 32767 : 16383);
	for (int i = 0; i < 32; i++) q.kthr[i] = q.kt16[i] = 0x7fffffff;
	for (int i = 0; i < 16; i++) {
		q.thr[i] = hq.thr[i];
		if (q.fast) q.kthr[i] = hq.thr[i] << 4;
	}
	// packed form (short levels only): 16-bit keys hold f - 2T in 11 bits, the quantised value 2q|sign in 15
	q.pk = !is_int && q.fast && hq.Q >= 4 && !(hq.thr[0] & 1) && getenv("RIC_QUANT_PK", 1);
	for (int i = 0; i < 16 && q.pk; i++) {
		const int d = hq.thr[i] - 2 * hq.T;
		if (d < 0 || d > 2046) q.pk = 0;
		else q.kt16[i] = 0x8000 | (d << 4);
	}
	q.h0 = std::max(hq.thr[0] >> 1, hq.T + 1);
}

// 2-D tensor map over a u8 source for the TMA experiment of ric_fwd0.cuh (width x rows, row pitch in bytes).
// The encoder lives in the driver; it is fetched through the runtime so that the library needs no libcuda link.
static bool make_src_tensor_map(CUtensorMap *map, const void *base, int width, unsigned long long rows, int pitch, unsigned box_w)
{
	typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
	                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
	                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	static encode_fn fn = nullptr;
	if (!fn) {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
		fn = (encode_fn)p;
	}
	if ((pitch & 15) || ((uintptr_t)base & 15)) return false;  // the unit wants 16-byte aligned rows
	const cuuint64_t dim[2] = {(cuuint64_t)width, rows}, stride[1] = {(cuuint64_t)pitch};
	const cuuint32_t box[2] = {box_w, 2}, es[2] = {1, 1};
	return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dim, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
	          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static BandRef band_ref(const HostGeom &g, int id)
{
	BandRef r;
	const ric_band_info &b = g.band[id];
	r.off = (long long)b.offset;
	r.dimx = b.dimx; r.dimy = b.dimy; r.stride = b.stride;
	r.fl_off = g.flag_off[id];
	r.fl_bw = g.flag_bw[id];
	return r;
}

// Rows per job.  Jobs are equal-sized and claimed dynamically by persistent CTAs, so a launch takes
// about (total loop iterations / concurrent jobs) + one job (the tail); a job of `seg` rows runs
// seg/2 + 4 iterations (the +4 is the vertical warm-up).  The segment height minimising that wins.
static int choose_seg_rows(const ric_ctx *c, int w, int h, long long planes_images, int jobs_per_sm)
{
	const long long nstrips = (w + STRIP_W - 1) / STRIP_W;
	const long long slots = c->target_warps > 0 ? c->target_warps : (long long)c->sm_count * jobs_per_sm;
	static const int cand[] = {256, 192, 160, 128, 112, 96, 80, 64, 56, 48, 40, 32, 24, 16, 8};
	int best = 8;
	long long best_cost = -1;
	for (int seg : cand) {
		if (seg > 8 && seg >= 2 * ((h + 7) & ~7)) continue;  // taller than the level: same as a smaller candidate
		const long long nseg = (h + seg - 1) / seg, rem = h - (nseg - 1) * seg;  // last segment may be short
		const long long it = (seg < h ? seg : ((h + 7) & ~7)) / 2 + 4, it_last = ((rem + 7) & ~7) / 2 + 4;
		const long long cols = nstrips * planes_images;          // independent columns of segments
		const long long work = cols * ((nseg - 1) * it + it_last);
		// jobs differ in weight (luma blocks cost more quantiser work than chroma), so even a single wave
		// benefits from several jobs per slot: charge a full job as the tail
		const long long cost = (work + slots - 1) / slots + it;
		if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = seg; }
	}
	return best;
}

extern "C" {

const char *ric_last_error(void) { return g_err; }
int ric_quants(int idx) { return quants(idx); }

int ric_plane_quant(int q, int channels, int plane, int *Quant, int *lambda)
{
	if (q < 0 || q > 31 || (channels != 1 && channels != 3) || plane < 0 || plane >= channels || !Quant || !lambda)
		return set_err(RIC_E_ARG, "ric_plane_quant: bad argument");
	plane_quant(q, channels, plane, Quant, lambda);
	return RIC_OK;
}

int ric_destroy(ric_ctx *c)
{
	if (!c) return RIC_OK;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	for (int i = 0; i < 3; i++)
		if (c->pipe[i]) cudaStreamSynchronize(c->pipe[i]);  // outstanding *_stream chunks and their callbacks
	for (int i = 0; i < 8; i++)
		if (c->ent[i]) cudaStreamSynchronize(c->ent[i]);
	cudaFree(c->d_src);
	cudaFree(c->d_arena);
	cudaFree(c->d_arena_in);
	if (c->h_stage) cudaFreeHost(c->h_stage);
	cudaFree(c->d_geom);
	cudaFree(c->d_tables);
	cudaFree(c->d_bad);
	cudaFree(c->d_hints);
	cudaFree(c->d_payload);
	cudaFree(c->d_psizes);
	if (c->h_psizes) cudaFreeHost(c->h_psizes);
	for (int i = 0; i < 8; i++) {
		if (c->ent[i]) cudaStreamDestroy(c->ent[i]);
		if (c->ent_ev[i]) cudaEventDestroy(c->ent_ev[i]);
	}
	cudaFree(c->d_flags);
	cudaFree(c->d_plane);
	cudaFree(c->d_count);
	cudaFree(c->d_stats);
	cudaFree(c->d_jobctr);
	for (int d = 0; d < 2; d++)
		for (int i = 0; i <= RIC_MAX_LEVELS; i++)
			if (c->ev[d][i]) cudaEventDestroy(c->ev[d][i]);
	for (int i = 0; i < RIC_MAX_LEVELS; i++) cudaFree(c->d_ll[i]);
	if (c->stream) cudaStreamDestroy(c->stream);
	for (int i = 0; i < 3; i++)
		if (c->pipe[i]) cudaStreamDestroy(c->pipe[i]);
	delete c;
	return RIC_OK;
}

int ric_create(ric_ctx **out, int device, int width, int height, int channels, int levels, int level_chg, int align,
               int trans, int max_batch)
{
	if (!out || max_batch < 1) return set_err(RIC_E_ARG, "ric_create: bad argument");
	*out = nullptr;
	HostGeom g;
	int rc = geom_init(g, width, height, channels, levels, level_chg, align, trans);
	if (rc != RIC_OK) return set_err(rc, "ric_create: unsupported geometry/transform");
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return set_err(RIC_E_CUDA, "ric_create: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
	if (device < 0 || device >= ndev) return set_err(RIC_E_ARG, "ric_create: bad device index");
	CK(cudaSetDevice(device));
	ric_ctx *c = new (std::nothrow) ric_ctx();
	if (!c) return set_err(RIC_E_NOMEM, "ric_create: out of host memory");
	c->g = g;
	c->device = device;
	c->max_batch = max_batch;
	{
		int sms = 0;
		cudaError_t pe = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
		if (pe != cudaSuccess) {
			delete c;
			return set_err(RIC_E_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(pe));
		}
		c->sm_count = sms;
	}
	const char *tw = getenv("RIC_TARGET_WARPS");
	c->target_warps = tw ? atoi(tw) : 0;  // override of the concurrent-job count used by choose_seg_rows
	// ric_fwd0.cuh: measured 6 % slower than the scalar kernel with the packed quantiser (12 warps per SM against
	// 16: profiles/README.md), so it is opt-in
	c->use_fwd0 = getenv("RIC_FWD0", 0);
	c->use_inv0 = getenv("RIC_INV0", 0);  // packed finest inverse level (ric_inv0.cuh)
	c->use_pdl = getenv("RIC_PDL", 1);  // programmatic dependent launch between the level kernels of a call
	c->use_tma = getenv("RIC_TMA", 0);  // gray level 0 of the packed kernel through the TMA unit (experiment, profiles/README.md)
#define CKD(call)                                                                    \
	do {                                                                             \
		cudaError_t e_ = (call);                                                     \
		if (e_ != cudaSuccess) {                                                     \
			set_err(e_ == cudaErrorMemoryAllocation ? RIC_E_NOMEM : RIC_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
			ric_destroy(c);                                                          \
			return e_ == cudaErrorMemoryAllocation ? RIC_E_NOMEM : RIC_E_CUDA;       \
		}                                                                            \
	} while (0)
	CKD(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	CKD(cudaFuncSetAttribute(fwd0_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f0_smem_bytes<3>()));
	CKD(cudaFuncSetAttribute(fwd0_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f0_smem_bytes<1>()));
	CKD(cudaFuncSetAttribute(fwd0_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f0_smem_bytes_tma()));
	for (int i = 0; i < 3; i++) CKD(cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking));
	const size_t nb = (size_t)max_batch, ch = (size_t)channels;
	c->src_pitch = ((size_t)width + 7) & ~(size_t)7;  // roundup8(w): dense rows (one contiguous copy) when w % 8 == 0
	CKD(cudaMalloc(&c->d_src, nb * ch * height * c->src_pitch + 64));
	CKD(cudaMalloc(&c->d_arena, nb * ch * g.arena_bytes + 64));
	CKD(cudaMemset(c->d_arena, 0, nb * ch * g.arena_bytes + 64));  // padding columns stay zero
	CKD(cudaMalloc(&c->d_flags, nb * ch * g.flag_bytes));
	CKD(cudaMemset(c->d_flags, 0, nb * ch * g.flag_bytes));
	c->plane_pitch = (width + 7 + 8) & ~7;
	CKD(cudaMalloc(&c->d_plane, ch * height * (size_t)c->plane_pitch * sizeof(short) + 64));
	for (int i = 0; i + 1 < g.nlev; i++) {
		c->ll_pitch[i] = (g.lev_w[i + 1] + 7 + 8) & ~7;
		c->ll_es[i] = (g.lev_int[i] || g.lev_int[i + 1]) ? 4 : 2;
		const size_t bytes = nb * ch * g.lev_h[i + 1] * (size_t)c->ll_pitch[i] * c->ll_es[i] + 64;
		CKD(cudaMalloc(&c->d_ll[i], bytes));
		CKD(cudaMemset(c->d_ll[i], 0, bytes));
	}
	CKD(cudaMalloc(&c->d_count, 4 * sizeof(unsigned)));
	CKD(cudaMalloc(&c->d_stats, 8 * sizeof(unsigned long long)));
	CKD(cudaMemset(c->d_stats, 0, 8 * sizeof(unsigned long long)));
	CKD(cudaMalloc(&c->d_jobctr, 4 * 2 * RIC_MAX_LEVELS * sizeof(unsigned long long)));
	c->cset = 3;
	for (int d = 0; d < 2; d++)
		for (int i = 0; i <= g.nlev; i++) CKD(cudaEventCreate(&c->ev[d][i]));
	CKD(cudaDeviceSynchronize());
#undef CKD
	*out = c;
	return RIC_OK;
}

int ric_get_info(const ric_ctx *c, ric_info *info)
{
	if (!c || !info) return set_err(RIC_E_ARG, "ric_get_info: null");
	const HostGeom &g = c->g;
	info->width = g.width; info->height = g.height; info->channels = g.channels;
	info->levels = g.levels; info->level_chg = g.level_chg; info->align = g.align; info->trans = g.trans;
	info->nlev = g.nlev; info->nbands = g.nbands; info->max_batch = c->max_batch;
	info->arena_bytes = g.arena_bytes;
	info->image_arena_bytes = g.arena_bytes * g.channels;
	return RIC_OK;
}

int ric_get_band(const ric_ctx *c, int id, ric_band_info *info)
{
	if (!c || !info || id < 0 || id >= c->g.nbands) return set_err(RIC_E_ARG, "ric_get_band: bad argument");
	*info = c->g.band[id];
	return RIC_OK;
}

int ric_last_launch_count(const ric_ctx *c) { return c ? c->launches : 0; }

int ric_set_profiling(ric_ctx *c, int on)
{
	if (!c) return set_err(RIC_E_ARG, "ric_set_profiling: null");
	c->profiling = on != 0;
	return RIC_OK;
}

int ric_get_level_times(ric_ctx *c, int direction, float *ms, int cap)
{
	if (!c || !ms || direction < 0 || direction > 1) return set_err(RIC_E_ARG, "ric_get_level_times: bad argument");
	const int n = c->ev_n[direction];
	if (n == 0 || cap < n) return set_err(RIC_E_ARG, "ric_get_level_times: nothing recorded (enable ric_set_profiling) or cap too small");
	CK(cudaSetDevice(c->device));
	CK(cudaEventSynchronize(c->ev[direction][n]));
	for (int i = 0; i < n; i++) CK(cudaEventElapsedTime(&ms[i], c->ev[direction][i], c->ev[direction][i + 1]));
	return n;
}

int ric_get_path_stats(ric_ctx *c, unsigned long long *out, int n)
{
	if (!c || !out || n < 1 || n > 8) return set_err(RIC_E_ARG, "ric_get_path_stats: bad argument");
	CK(cudaSetDevice(c->device));
	CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(out, c->d_stats, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	CK(cudaMemset(c->d_stats, 0, 8 * sizeof(unsigned long long)));
	return RIC_OK;
}

int ric_header_write(uint8_t *out, int width, int height, int q, int color, int trans)
{
	if (!out || width < 1 || width > 65535 || height < 1 || height > 65535 || q < 0 || q > 31 || (color & ~1) || trans < 0 || trans > 2)
		return set_err(RIC_E_ARG, "ric_header_write: bad argument");
	out[0] = 'R'; out[1] = 'U'; out[2] = 'D'; out[3] = '2';
	out[4] = (uint8_t)(width & 0xFF); out[5] = (uint8_t)(width >> 8);
	out[6] = (uint8_t)(height & 0xFF); out[7] = (uint8_t)(height >> 8);
	out[8] = (uint8_t)(q | color << 5 | trans << 6);
	return RIC_OK;
}

int ric_header_parse(const uint8_t *in, int *width, int *height, int *q, int *color, int *trans)
{
	if (!in || !width || !height || !q || !color || !trans) return set_err(RIC_E_ARG, "ric_header_parse: null");
	if (in[0] != 'R' || in[1] != 'U' || in[2] != 'D' || in[3] != '2') return set_err(RIC_E_ARG, "ric_header_parse: bad magic (ric.cpp:189-190)");
	*width = in[4] | in[5] << 8;
	*height = in[6] | in[7] << 8;
	*q = in[8] & 31;
	*color = (in[8] >> 5) & 1;
	*trans = in[8] >> 6;
	return RIC_OK;
}

int ric_host_alloc(void **p, size_t bytes)
{
	if (!p) return set_err(RIC_E_ARG, "ric_host_alloc: null");
	cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault);
	if (e != cudaSuccess) return set_err(RIC_E_NOMEM, "cudaHostAlloc: %s", cudaGetErrorString(e));
	return RIC_OK;
}

int ric_host_free(void *p)
{
	if (p) cudaFreeHost(p);
	return RIC_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// launch sequences

// Forward: all levels of n images.  src_kind SRC_U8_* (d_src, pitch bytes) or SRC_S16 (plane API).
// Quant[p]/lambda[p] per plane; do_quant = 0 stores raw coefficients.
static int launch_forward(ric_ctx *c, const void *d_src, int src_kind, long long src_img_stride,
                          long long src_plane_stride, int src_pitch, int n, int nplanes, int shift, int do_quant,
                          const int *Quant, const int *lambda, char *d_arena, cudaStream_t st)
{
	const HostGeom &g = c->g;
	c->launches = 0;
	c->ev_n[0] = 0;
	unsigned long long *ctr = c->d_jobctr + (size_t)(c->cset * 2 + 0) * RIC_MAX_LEVELS;
	CK(cudaMemsetAsync(ctr, 0, RIC_MAX_LEVELS * sizeof(unsigned long long), st));
	if (c->profiling) { CK(cudaEventRecord(c->ev[0][0], st)); }
	for (int lv = 0; lv < g.nlev; lv++) {
		FwdParams P;
		memset(&P, 0, sizeof P);
		const bool last = lv == g.nlev - 1;
		const bool sh = !g.lev_int[lv];
		int src;
		if (lv == 0) {
			P.src = d_src; src = src_kind;
			P.src_img_stride = src_img_stride; P.src_plane_stride = src_plane_stride; P.src_pitch = src_pitch;
		} else {
			src = c->ll_es[lv - 1] == 4 && g.lev_int[lv - 1] ? SRC_S32 : SRC_S16;
			P.src_pitch = c->ll_pitch[lv - 1];
			P.src_plane_stride = (long long)g.lev_h[lv] * P.src_pitch;
			P.src_img_stride = P.src_plane_stride * g.channels;
			P.src = (const char *)c->d_ll[lv - 1] + (size_t)c->img0 * P.src_img_stride * (src == SRC_S32 ? 4 : 2);
		}
		P.ll_to_band = last;
		if (!last) {
			P.ll_pitch = c->ll_pitch[lv];
			P.ll_plane_stride = (long long)g.lev_h[lv + 1] * P.ll_pitch;
			P.ll_img_stride = P.ll_plane_stride * g.channels;
			P.ll = (char *)c->d_ll[lv] + (size_t)c->img0 * P.ll_img_stride * (g.lev_int[lv] ? 4 : 2);
		}
		P.arena = d_arena;
		P.arena_plane_stride = (long long)g.arena_bytes;
		P.arena_img_stride = (long long)g.arena_bytes * g.channels;
		P.flags_plane_stride = (long long)g.flag_bytes;
		P.flags_img_stride = (long long)g.flag_bytes * g.channels;
		P.flags = c->d_flags + (size_t)c->img0 * P.flags_img_stride;
		for (int o = 0; o < 3; o++) {
			P.band[o] = band_ref(g, 3 * lv + o);
			if (lv > 0) P.child[o] = band_ref(g, 3 * (lv - 1) + o);
		}
		P.lband = band_ref(g, 3 * g.nlev);
		P.has_child = lv > 0;
		P.w = g.lev_w[lv]; P.h = g.lev_h[lv];
		P.nstrips = (P.w + STRIP_W - 1) / STRIP_W;
		P.seg_rows = choose_seg_rows(c, P.w, P.h, (long long)nplanes * n, 16);
		P.nsegs = (P.h + P.seg_rows - 1) / P.seg_rows;
		P.nplanes = nplanes; P.nimages = n;
		P.shift = shift; P.quant = do_quant;
		// quantiser classes: plane 2 of an RGB image is luma (class 0), planes 0/1 chroma (class 1)
		int cls_plane[2] = {nplanes == 3 ? 2 : 0, 0};
		for (int p = 0; p < 3; p++) P.plane_class[p] = (nplanes == 3 && p != 2) ? 1 : 0;
		if (do_quant) {
			for (int cls = 0; cls < 2; cls++) {
				const int p = cls_plane[cls];
				for (int o = 0; o < 3; o++) {
					fill_qb(P.qb[cls][o], Quant[p], lambda[p], g.band[3 * lv + o].weight, g.lev_int[lv]);
				}
				make_tsuq(Quant[p], 0.5f, g.band[3 * g.nlev].weight, g.lev_int[g.nlev - 1], &P.llQ[cls], &P.lliQ[cls], &P.llT[cls]);
			}
		}
		P.counter = ctr + lv;
		P.stats = c->profiling ? c->d_stats : nullptr;
		// level 0 from 8-bit pixels: the packed kernel (ric_fwd0.cuh), one warp job = all planes of a strip segment
		const bool packed0 = c->use_fwd0 && lv == 0 && sh && !last && shift && g.trans == RIC_CDF97 &&
		                     ((src == SRC_U8_RGB && nplanes == 3) || (src == SRC_U8_GRAY && nplanes == 1));
		if (packed0) {
			const int occ = nplanes == 3 ? 3 : 4;
			P.seg_rows = choose_seg_rows(c, P.w, P.h, (long long)n, F0_WARPS * occ);
			P.nsegs = (P.h + P.seg_rows - 1) / P.seg_rows;
			const long long njobs = (long long)P.nstrips * P.nsegs * n;
			if (njobs >= (1ll << 31)) return set_err(RIC_E_ARG, "forward: batch too large for one launch");
			const unsigned grid = (unsigned)std::min<long long>((njobs + F0_WARPS - 1) / F0_WARPS, (long long)c->sm_count * occ);
			CUtensorMap tmap, tmapB;
			if (nplanes == 3) fwd0_kernel<3><<<grid, F0_WARPS * 32, f0_smem_bytes<3>(), st>>>(P);
			else if (c->use_tma && make_src_tensor_map(&tmap, d_src, P.w, (unsigned long long)n * P.h, src_pitch, 256) &&
			         make_src_tensor_map(&tmapB, d_src, P.w, (unsigned long long)n * P.h, src_pitch, 16))
				fwd0_tma_kernel<<<grid, F0_WARPS * 32, f0_smem_bytes_tma(), st>>>(P, tmap, tmapB);
			else fwd0_kernel<1><<<grid, F0_WARPS * 32, f0_smem_bytes<1>(), st>>>(P);
		} else {
			fwd_fn fn = pick_fwd(sh, g.trans, src);
			if (!fn) return set_err(RIC_E_UNSUPPORTED, "forward: unsupported level type combination");
			const long long njobs = (long long)P.nstrips * P.nsegs * nplanes * n;
			if (njobs >= (1ll << 31)) return set_err(RIC_E_ARG, "forward: batch too large for one launch");
			const int wpb = fwd_warps(sh);
			const unsigned grid = (unsigned)std::min<long long>((njobs + wpb - 1) / wpb, (long long)c->sm_count * 4);
			CK(launch_level(fn, grid, wpb * 32, st, P, c->use_pdl && !c->profiling && lv > 0));
		}
		CK(cudaGetLastError());
		c->launches++;
		if (c->profiling) { CK(cudaEventRecord(c->ev[0][c->launches], st)); c->ev_n[0] = c->launches; }
	}
	return RIC_OK;
}

// Inverse: all levels of n images.  dq_Quant[p] = TSUQi argument per plane (0: no dequantisation).
static int launch_inverse(ric_ctx *c, const char *d_arena, int n, int nplanes, int shift, const int *dq_Quant,
                          int dst_kind, void *d_dst, long long dst_img_stride, long long dst_plane_stride,
                          int dst_pitch, int q1_quirk, cudaStream_t st)
{
	const HostGeom &g = c->g;
	c->launches = 0;
	c->ev_n[1] = 0;
	unsigned long long *ctr = c->d_jobctr + (size_t)(c->cset * 2 + 1) * RIC_MAX_LEVELS;
	CK(cudaMemsetAsync(ctr, 0, RIC_MAX_LEVELS * sizeof(unsigned long long), st));
	if (c->profiling) { CK(cudaEventRecord(c->ev[1][0], st)); }
	for (int lv = g.nlev - 1; lv >= 0; lv--) {
		InvParams P;
		memset(&P, 0, sizeof P);
		const bool coarsest = lv == g.nlev - 1;
		const bool sh = !g.lev_int[lv];
		P.arena = d_arena;
		P.arena_plane_stride = (long long)g.arena_bytes;
		P.arena_img_stride = (long long)g.arena_bytes * g.channels;
		int dst;
		if (coarsest) P.llsrc = LLSRC_BAND;
		else {
			P.llsrc = g.lev_int[lv + 1] ? LLSRC_S32 : LLSRC_S16;
			P.ll_pitch = c->ll_pitch[lv];
			P.ll_plane_stride = (long long)g.lev_h[lv + 1] * P.ll_pitch;
			P.ll_img_stride = P.ll_plane_stride * g.channels;
			P.ll = (const char *)c->d_ll[lv] + (size_t)c->img0 * P.ll_img_stride * (g.lev_int[lv + 1] ? 4 : 2);
		}
		if (lv == 0) {
			dst = dst_kind;
			P.dst = d_dst; P.dst_img_stride = dst_img_stride; P.dst_plane_stride = dst_plane_stride; P.dst_pitch = dst_pitch;
		} else {
			dst = DST_PLANE;
			P.dst_pitch = c->ll_pitch[lv - 1];
			P.dst_plane_stride = (long long)g.lev_h[lv] * P.dst_pitch;
			P.dst_img_stride = P.dst_plane_stride * g.channels;
			P.dst = (char *)c->d_ll[lv - 1] + (size_t)c->img0 * P.dst_img_stride * (g.lev_int[lv] ? 4 : 2);
		}
		for (int o = 0; o < 3; o++) P.band[o] = band_ref(g, 3 * lv + o);
		P.lband = band_ref(g, 3 * g.nlev);
		P.h_row1_off = (g.trans == RIC_CDF53 && q1_quirk) ? P.band[0].stride : P.band[1].stride;
		P.w = g.lev_w[lv]; P.h = g.lev_h[lv];
		P.nstrips = (P.w + STRIP_W - 1) / STRIP_W;
		const int jplanes = dst == DST_U8_RGB ? 1 : nplanes;
		P.seg_rows = choose_seg_rows(c, P.w, P.h, (long long)jplanes * n, dst == DST_U8_RGB ? 6 : 16);
		P.nsegs = (P.h + P.seg_rows - 1) / P.seg_rows;
		P.nplanes = nplanes; P.nimages = n;
		P.shift = shift;
		for (int p = 0; p < 3; p++)
			for (int o = 0; o < 4; o++) {
				const int id = o < 3 ? 3 * lv + o : 3 * g.nlev;
				P.dq[p][o] = (p < nplanes && dq_Quant[p]) ? make_tsuqi(dq_Quant[p], g.band[id].weight, g.band[id].is_int) : 1;
			}
		inv_fn fn = pick_inv(sh, g.trans, dst);
		if (!fn) return set_err(RIC_E_UNSUPPORTED, "inverse: unsupported level type combination");
		const long long njobs = (long long)P.nstrips * P.nsegs * jplanes * n;
		if (njobs >= (1ll << 31)) return set_err(RIC_E_ARG, "inverse: batch too large for one launch");
		P.counter = ctr + lv;
		P.stats = c->profiling ? c->d_stats : nullptr;
		// finest level to 8-bit pixels: the packed kernel (ric_inv0.cuh)
		const bool packed0 = c->use_inv0 && lv == 0 && sh && shift && g.trans == RIC_CDF97 && P.llsrc == LLSRC_S16 &&
		                     ((dst == DST_U8_RGB && nplanes == 3) || (dst == DST_U8_GRAY && nplanes == 1));
		if (packed0) {
			if (dst == DST_U8_RGB) {
				const unsigned grid = (unsigned)std::min<long long>(njobs, (long long)c->sm_count * 6);
				inv0_kernel<DST_U8_RGB><<<grid, 96, 0, st>>>(P);
			} else {
				const unsigned grid = (unsigned)std::min<long long>((njobs + INV_WARPS - 1) / INV_WARPS, (long long)c->sm_count * 4);
				inv0_kernel<DST_U8_GRAY><<<grid, INV_WARPS * 32, 0, st>>>(P);
			}
		} else if (dst == DST_U8_RGB) {
			const unsigned grid = (unsigned)std::min<long long>((njobs + INV_RGB_GROUPS - 1) / INV_RGB_GROUPS, (long long)c->sm_count * getenv("RIC_INV_RGB_OCC", 6 / INV_RGB_GROUPS));
			CK(launch_level(fn, grid, INV_RGB_GROUPS * 96, st, P, c->use_pdl && !c->profiling && lv < g.nlev - 1));
		} else {
			const unsigned grid = (unsigned)std::min<long long>((njobs + INV_WARPS - 1) / INV_WARPS, (long long)c->sm_count * 4);
			CK(launch_level(fn, grid, INV_WARPS * 32, st, P, c->use_pdl && !c->profiling && lv < g.nlev - 1));
		}
		CK(cudaGetLastError());
		c->launches++;
		if (c->profiling) { CK(cudaEventRecord(c->ev[1][c->launches], st)); c->ev_n[1] = c->launches; }
	}
	return RIC_OK;
}

static int need_arena_in(ric_ctx *c)
{
	if (c->d_arena_in) return RIC_OK;
	const size_t bytes = (size_t)c->max_batch * c->g.channels * c->g.arena_bytes + 64;
	cudaError_t e = cudaMalloc(&c->d_arena_in, bytes);
	if (e != cudaSuccess) return set_err(RIC_E_NOMEM, "cudaMalloc(decode arena): %s", cudaGetErrorString(e));
	return RIC_OK;
}

static int check_batch(const ric_ctx *c, int n, int q, const char *who)
{
	if (!c) return set_err(RIC_E_ARG, "%s: null context", who);
	if (n < 1 || n > c->max_batch) return set_err(RIC_E_ARG, "%s: batch size outside [1, max_batch]", who);
	if (q < 0 || q > 31) return set_err(RIC_E_ARG, "%s: q outside [0, 31]", who);
	return RIC_OK;
}

extern "C" {

int ric_encode_u8_device(ric_ctx *c, const uint8_t *d_src, size_t pitch, int n, int q, void *d_arenas, void *stream)
{
	int rc = check_batch(c, n, q, "ric_encode_u8_device");
	if (rc) return rc;
	const HostGeom &g = c->g;
	if (!d_src || !d_arenas || (pitch & 7) || pitch < (size_t)((g.width + 7) & ~7) || ((uintptr_t)d_src & 7) ||
	    ((uintptr_t)d_arenas & 31))
		return set_err(RIC_E_ARG, "ric_encode_u8_device: bad pointer/pitch (pitch must be a multiple of 8 and >= roundup8(width))");
	CK(cudaSetDevice(c->device));
	int Q[3] = {0, 0, 0}, L[3] = {0, 0, 0};
	for (int p = 0; p < g.channels; p++) plane_quant(q, g.channels, p, &Q[p], &L[p]);
	return launch_forward(c, d_src, g.channels == 3 ? SRC_U8_RGB : SRC_U8_GRAY, (long long)pitch * g.height * g.channels,
	                      (long long)pitch * g.height, (int)pitch, n, g.channels, q != 0, 1, Q, L, (char *)d_arenas,
	                      (cudaStream_t)stream);
}

int ric_decode_u8_device(ric_ctx *c, const void *d_arenas, int n, int q, uint8_t *d_dst, size_t pitch, void *stream)
{
	int rc = check_batch(c, n, q, "ric_decode_u8_device");
	if (rc) return rc;
	const HostGeom &g = c->g;
	if (!d_dst || !d_arenas || (pitch & 7) || pitch < (size_t)((g.width + 7) & ~7) || ((uintptr_t)d_dst & 7) ||
	    ((uintptr_t)d_arenas & 31))
		return set_err(RIC_E_ARG, "ric_decode_u8_device: bad pointer/pitch (pitch must be a multiple of 8 and >= roundup8(width))");
	CK(cudaSetDevice(c->device));
	int Q[3] = {0, 0, 0}, L[3];
	for (int p = 0; p < g.channels; p++) plane_quant(q, g.channels, p, &Q[p], &L[p]);
	return launch_inverse(c, (const char *)d_arenas, n, g.channels, q != 0, Q, g.channels == 3 ? DST_U8_RGB : DST_U8_GRAY,
	                      d_dst, (long long)pitch * g.height * g.channels, (long long)pitch * g.height, (int)pitch, 1,
	                      (cudaStream_t)stream);
}

// Host-buffer entry points: the batch is cut into chunks that alternate over three streams, so the
// H2D copy of chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlap (each chunk
// owns its image slots of every device buffer, LL scratch and flags included).
static cudaError_t copy_pixels(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows,
                               cudaMemcpyKind kind, cudaStream_t st)
{
	if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * rows, kind, st);
	return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, st);
}

static int chunk_images(int n)
{
	static const int parts = getenv("RIC_CHUNKS", 8);  // chunks per call of the pipelined host-buffer paths
	return n >= 12 ? (n + parts - 1) / parts : n >= 4 ? 2 : 1;
}

static int sync_pipe(ric_ctx *c)
{
	for (int i = 0; i < 3; i++) CK(cudaStreamSynchronize(c->pipe[i]));
	return RIC_OK;
}

static void CUDART_CB chunk_trampoline(void *p)
{
	ric_ctx::ChunkNote *n = (ric_ctx::ChunkNote *)p;
	n->fn(n->user, n->first, n->count);
}

int ric_sync(ric_ctx *c)
{
	if (!c) return set_err(RIC_E_ARG, "ric_sync: null context");
	CK(cudaSetDevice(c->device));
	return sync_pipe(c);
}

int ric_encode_u8_stream(ric_ctx *c, const uint8_t *src, int n, int q, void *arenas, ric_chunk_fn done, void *user)
{
	int rc = check_batch(c, n, q, "ric_encode_u8_stream");
	if (rc) return rc;
	if (!src || !arenas) return set_err(RIC_E_ARG, "ric_encode_u8_stream: null buffer");
	const HostGeom &g = c->g;
	CK(cudaSetDevice(c->device));
	if ((rc = sync_pipe(c))) return rc;  // the previous call's chunk notes are about to be reused
	const size_t img_px = (size_t)g.channels * g.height * g.width, img_dev = (size_t)g.channels * g.height * c->src_pitch;
	const size_t img_ar = (size_t)g.channels * g.arena_bytes;
	const int step = chunk_images(n);
	c->notes.assign((size_t)(n + step - 1) / step, ric_ctx::ChunkNote{done, user, 0, 0});
	int total = 0, k = 0;
	for (int i0 = 0; i0 < n; i0 += step, k++) {
		const int m = std::min(step, n - i0);
		cudaStream_t st = c->pipe[k % 3];
		CK(copy_pixels(c->d_src + i0 * img_dev, c->src_pitch, src + i0 * img_px, g.width, g.width,
		               (size_t)m * g.channels * g.height, cudaMemcpyHostToDevice, st));
		c->img0 = i0;
		c->cset = k % 3;
		rc = ric_encode_u8_device(c, c->d_src + i0 * img_dev, c->src_pitch, m, q, c->d_arena + i0 * img_ar, st);
		c->img0 = 0;
		c->cset = 3;
		if (rc) { sync_pipe(c); return rc; }
		total += c->launches;
		CK(cudaMemcpyAsync((char *)arenas + i0 * img_ar, c->d_arena + i0 * img_ar, (size_t)m * img_ar, cudaMemcpyDeviceToHost, st));
		if (done) {
			c->notes[k].first = i0;
			c->notes[k].count = m;
			CK(cudaLaunchHostFunc(st, chunk_trampoline, &c->notes[k]));
		}
	}
	c->launches = total;
	return RIC_OK;
}

int ric_encode_u8(ric_ctx *c, const uint8_t *src, int n, int q, void *arenas)
{
	int rc = ric_encode_u8_stream(c, src, n, q, arenas, nullptr, nullptr);
	return rc ? rc : sync_pipe(c);
}

// H2D of the arenas of images [i0, i0+m), decode stage, D2H of their pixels, all on pipeline stream k % 3.
static int decode_chunk(ric_ctx *c, const void *arenas, int i0, int m, int q, uint8_t *dst, int k)
{
	const HostGeom &g = c->g;
	const size_t img_px = (size_t)g.channels * g.height * g.width, img_dev = (size_t)g.channels * g.height * c->src_pitch;
	const size_t img_ar = (size_t)g.channels * g.arena_bytes;
	cudaStream_t st = c->pipe[k % 3];
	CK(cudaMemcpyAsync(c->d_arena_in + i0 * img_ar, (const char *)arenas + i0 * img_ar, (size_t)m * img_ar, cudaMemcpyHostToDevice, st));
	c->img0 = i0;
	c->cset = k % 3;
	const int rc = ric_decode_u8_device(c, c->d_arena_in + i0 * img_ar, m, q, c->d_src + i0 * img_dev, c->src_pitch, st);
	c->img0 = 0;
	c->cset = 3;
	if (rc) return rc;
	CK(copy_pixels(dst + i0 * img_px, g.width, c->d_src + i0 * img_dev, c->src_pitch, g.width, (size_t)m * g.channels * g.height,
	               cudaMemcpyDeviceToHost, st));
	return RIC_OK;
}

int ric_decode_u8_stream(ric_ctx *c, const void *arenas, int n, int q, uint8_t *dst, ric_chunk_fn done, void *user)
{
	int rc = check_batch(c, n, q, "ric_decode_u8_stream");
	if (rc) return rc;
	if (!dst || !arenas) return set_err(RIC_E_ARG, "ric_decode_u8_stream: null buffer");
	CK(cudaSetDevice(c->device));
	if ((rc = need_arena_in(c))) return rc;
	if ((rc = sync_pipe(c))) return rc;
	const int step = chunk_images(n);
	c->notes.assign((size_t)(n + step - 1) / step, ric_ctx::ChunkNote{done, user, 0, 0});
	int total = 0, k = 0;
	for (int i0 = 0; i0 < n; i0 += step, k++) {
		const int m = std::min(step, n - i0);
		if ((rc = decode_chunk(c, arenas, i0, m, q, dst, k))) { sync_pipe(c); return rc; }
		total += c->launches;
		if (done) {
			c->notes[k].first = i0;
			c->notes[k].count = m;
			CK(cudaLaunchHostFunc(c->pipe[k % 3], chunk_trampoline, &c->notes[k]));
		}
	}
	c->launches = total;
	return RIC_OK;
}

int ric_decode_u8(ric_ctx *c, const void *arenas, int n, int q, uint8_t *dst)
{
	int rc = ric_decode_u8_stream(c, arenas, n, q, dst, nullptr, nullptr);
	return rc ? rc : sync_pipe(c);
}

// ---- plane-level API (one plane, batch slot 0, plane slot 0) ------------------------------------

int ric_transform(ric_ctx *c, const int16_t *plane, int stride, void *arena)
{
	if (!c || !plane) return set_err(RIC_E_ARG, "ric_transform: null");
	const HostGeom &g = c->g;
	if (stride < g.width) return set_err(RIC_E_ARG, "ric_transform: stride < width");
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpy2DAsync(c->d_plane, (size_t)c->plane_pitch * 2, plane, (size_t)stride * 2, (size_t)g.width * 2, g.height,
	                     cudaMemcpyHostToDevice, c->stream));
	int Q[3] = {0, 0, 0};
	int rc = launch_forward(c, c->d_plane, SRC_S16, 0, 0, c->plane_pitch, 1, 1, 0, 0, Q, Q, c->d_arena, c->stream);
	if (rc) return rc;
	if (arena) CK(cudaMemcpyAsync(arena, c->d_arena, g.arena_bytes, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return RIC_OK;
}

int ric_quant(ric_ctx *c, int Quant, int lambda, void *arena)
{
	if (!c || Quant < 0 || lambda < 0) return set_err(RIC_E_ARG, "ric_quant: bad argument");
	const HostGeom &g = c->g;
	CK(cudaSetDevice(c->device));
	for (int lv = 0; lv < g.nlev; lv++) {
		QuantLevelParams P;
		memset(&P, 0, sizeof P);
		P.arena = c->d_arena;
		P.flags = c->d_flags;
		P.has_child = lv > 0;
		P.is_int = g.lev_int[lv];
		int maxblk = 0;
		for (int o = 0; o < 3; o++) {
			P.band[o] = band_ref(g, 3 * lv + o);
			if (lv > 0) P.child[o] = band_ref(g, 3 * (lv - 1) + o);
			fill_qb(P.qb[o], Quant, lambda, g.band[3 * lv + o].weight, g.lev_int[lv]);
			maxblk = std::max(maxblk, P.band[o].fl_bw * ((P.band[o].dimy + 3) / 4));
		}
		dim3 grid((maxblk + 127) / 128, 3);
		if (g.lev_int[lv]) quant_level_kernel<false><<<grid, 128, 0, c->stream>>>(P);
		else quant_level_kernel<true><<<grid, 128, 0, c->stream>>>(P);
		CK(cudaGetLastError());
	}
	{
		const ric_band_info &b = g.band[3 * g.nlev];
		int Q, iQ, T;
		make_tsuq(Quant, 0.5f, b.weight, b.is_int, &Q, &iQ, &T);
		if (b.is_int) band_pointwise_kernel<false><<<32, 128, 0, c->stream>>>(c->d_arena, (long long)b.offset, b.dimx, b.dimy, b.stride, 0, Q, iQ, T, nullptr);
		else band_pointwise_kernel<true><<<32, 128, 0, c->stream>>>(c->d_arena, (long long)b.offset, b.dimx, b.dimy, b.stride, 0, Q, iQ, T, nullptr);
		CK(cudaGetLastError());
	}
	if (arena) CK(cudaMemcpyAsync(arena, c->d_arena, g.arena_bytes, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return RIC_OK;
}

static int pointwise_all(ric_ctx *c, char *d_arena, int mode, int Quant, float thres, unsigned *count)
{
	const HostGeom &g = c->g;
	CK(cudaMemsetAsync(c->d_count, 0, 4 * sizeof(unsigned), c->stream));
	for (int id = 0; id < g.nbands; id++) {
		const ric_band_info &b = g.band[id];
		int Q = 1, iQ = 0, T = 0;
		if (mode == 0) make_tsuq(Quant, id == 3 * g.nlev ? 0.5f : thres, b.weight, b.is_int, &Q, &iQ, &T);  // LL: wavelet2d.cpp:240-243
		else Q = make_tsuqi(Quant, b.weight, b.is_int);
		const int blocks = std::max(1, std::min(1024, (b.dimx * b.dimy + 1023) / 1024));
		if (b.is_int) band_pointwise_kernel<false><<<blocks, 256, 0, c->stream>>>(d_arena, (long long)b.offset, b.dimx, b.dimy, b.stride, mode, Q, iQ, T, c->d_count);
		else band_pointwise_kernel<true><<<blocks, 256, 0, c->stream>>>(d_arena, (long long)b.offset, b.dimx, b.dimy, b.stride, mode, Q, iQ, T, c->d_count);
		CK(cudaGetLastError());
	}
	if (count) CK(cudaMemcpyAsync(count, c->d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
	return RIC_OK;
}

int ric_tsuq(ric_ctx *c, int Quant, float thres, void *arena, unsigned *count)
{
	if (!c || Quant < 0) return set_err(RIC_E_ARG, "ric_tsuq: bad argument");
	CK(cudaSetDevice(c->device));
	int rc = pointwise_all(c, c->d_arena, 0, Quant, thres, count);
	if (rc) return rc;
	if (arena) CK(cudaMemcpyAsync(arena, c->d_arena, c->g.arena_bytes, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return RIC_OK;
}

int ric_tsuqi(ric_ctx *c, int Quant, void *arena)
{
	if (!c || !arena || Quant < 0) return set_err(RIC_E_ARG, "ric_tsuqi: bad argument");
	CK(cudaSetDevice(c->device));
	int rc = need_arena_in(c);
	if (rc) return rc;
	CK(cudaMemcpyAsync(c->d_arena_in, arena, c->g.arena_bytes, cudaMemcpyHostToDevice, c->stream));
	rc = pointwise_all(c, c->d_arena_in, 1, Quant, 0.f, nullptr);
	if (rc) return rc;
	CK(cudaMemcpyAsync(arena, c->d_arena_in, c->g.arena_bytes, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return RIC_OK;
}

int ric_set_base_weight(ric_ctx *c, float baseWeight)
{
	if (!c || !(baseWeight > 0.f)) return set_err(RIC_E_ARG, "ric_set_base_weight: bad argument");
	HostGeom &g = c->g;
	const float scale = g.trans == RIC_CDF97 ? 1.149604398f * 1.149604398f : 2.f;  // wavelet2d.cpp:1011-1016
	float d = baseWeight / scale, v = baseWeight, l = baseWeight * scale;
	for (int i = 0; i < g.nlev; i++) {
		if (i > 0) { d = v; v = l; l = v * scale; }
		g.band[3 * i + 0].weight = d;
		g.band[3 * i + 1].weight = v;
		g.band[3 * i + 2].weight = v;
	}
	g.band[3 * g.nlev].weight = l;
	return RIC_OK;
}

int ric_quant_host(ric_ctx *c, int Quant, int lambda, void *arena)
{
	if (!c || !arena) return set_err(RIC_E_ARG, "ric_quant_host: null");
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync(c->d_arena, arena, c->g.arena_bytes, cudaMemcpyHostToDevice, c->stream));
	return ric_quant(c, Quant, lambda, arena);
}

int ric_tsuq_host(ric_ctx *c, int Quant, float thres, void *arena, unsigned *count)
{
	if (!c || !arena) return set_err(RIC_E_ARG, "ric_tsuq_host: null");
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync(c->d_arena, arena, c->g.arena_bytes, cudaMemcpyHostToDevice, c->stream));
	return ric_tsuq(c, Quant, thres, arena, count);
}

// ---- single bands without a context ------------------------------------------------------------------------

static int buf_check(int device, const ric_band_buf *b, const char *who)
{
	if (!b || !b->data || b->dimx < 1 || b->dimy < 1 || b->stride < b->dimx || !(b->weight > 0.f))
		return set_err(RIC_E_ARG, "%s: bad band description", who);
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return set_err(RIC_E_CUDA, "%s: no CUDA device (%s); this library has no CPU path", who, cudaGetErrorString(e));
	if (device < 0 || device >= ndev) return set_err(RIC_E_ARG, "%s: bad device index", who);
	CK(cudaSetDevice(device));
	return RIC_OK;
}

static int buf_pointwise(int device, const ric_band_buf *b, int mode, int Quant, float thres, unsigned *count, int *mn, int *mx,
                         const char *who)
{
	int rc = buf_check(device, b, who);
	if (rc) return rc;
	if (Quant < 0) return set_err(RIC_E_ARG, "%s: Quant < 0", who);
	const size_t bytes = (size_t)b->stride * b->dimy * (b->is_int ? 4 : 2);
	char *d = nullptr;
	unsigned *dc = nullptr;
	cudaError_t e = cudaMalloc(&d, bytes);
	if (e == cudaSuccess) e = cudaMalloc(&dc, 4 * sizeof(unsigned));
	if (e == cudaSuccess) e = cudaMemset(dc, 0, 4 * sizeof(unsigned));
	if (e == cudaSuccess) e = cudaMemcpy(d, b->data, bytes, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) {
		int Q = 1, iQ = 0, T = 0;
		if (mode == 0) make_tsuq(Quant, thres, b->weight, b->is_int, &Q, &iQ, &T);
		else Q = make_tsuqi(Quant, b->weight, b->is_int);
		const int blocks = std::max(1, std::min(1024, (b->dimx * b->dimy + 1023) / 1024));
		if (b->is_int) band_pointwise_kernel<false><<<blocks, 256>>>(d, 0, b->dimx, b->dimy, b->stride, mode, Q, iQ, T, dc);
		else band_pointwise_kernel<true><<<blocks, 256>>>(d, 0, b->dimx, b->dimy, b->stride, mode, Q, iQ, T, dc);
		e = cudaGetLastError();
	}
	unsigned hc[4] = {0, 0, 0, 0};
	if (e == cudaSuccess) e = cudaMemcpy(b->data, d, bytes, cudaMemcpyDeviceToHost);
	if (e == cudaSuccess) e = cudaMemcpy(hc, dc, sizeof hc, cudaMemcpyDeviceToHost);
	cudaFree(d);
	cudaFree(dc);
	if (e != cudaSuccess) return set_err(RIC_E_CUDA, "%s: %s", who, cudaGetErrorString(e));
	if (count) *count = hc[0];
	if (mn) *mn = (int)hc[1];
	if (mx) *mx = (int)hc[2];
	return RIC_OK;
}

int ric_buf_tsuq(int device, const ric_band_buf *band, int Quant, float thres, unsigned *count, int *mn, int *mx)
{
	return buf_pointwise(device, band, 0, Quant, thres, count, mn, mx, "ric_buf_tsuq");
}

int ric_buf_tsuqi(int device, const ric_band_buf *band, int Quant)
{
	return buf_pointwise(device, band, 1, Quant, 0.f, nullptr, nullptr, nullptr, "ric_buf_tsuqi");
}

int ric_buf_build_tree(int device, const ric_band_buf *chain, int n, int high_band, const unsigned char *child_flags,
                       int child_dimx, int Quant, int lambda)
{
	if (n < 1 || n > RIC_MAX_LEVELS || Quant < 0 || lambda < 0 || (!high_band && (!child_flags || child_dimx < 1)))
		return set_err(RIC_E_ARG, "ric_buf_build_tree: bad argument");
	for (int i = 0; i < n; i++) {
		int rc = buf_check(device, chain + i, "ric_buf_build_tree");
		if (rc) return rc;
		// a parent band is half its child in both directions (wavelet2d.cpp:73-79): the kernel indexes child blocks 2bx, 2by
		if (i > 0 && (chain[i].dimx > (chain[i - 1].dimx + 1) / 2 + 1 || chain[i].dimy > (chain[i - 1].dimy + 1) / 2 + 1))
			return set_err(RIC_E_ARG, "ric_buf_build_tree: chain[i] is not the next coarser band of chain[i-1]");
	}
	// one device block: [child flags][band 0][flags 0][band 1][flags 1]...
	size_t off = 0, boff[RIC_MAX_LEVELS], foff[RIC_MAX_LEVELS], cf_off = 0;
	const int child_bw = high_band ? 0 : (child_dimx + 3) / 4;
	int fbw[RIC_MAX_LEVELS], fbh[RIC_MAX_LEVELS];
	if (!high_band) {
		fbw[0] = (chain[0].dimx + 3) / 4; fbh[0] = (chain[0].dimy + 3) / 4;
		off = (((size_t)child_bw * (2 * fbh[0] + 2) + 2 * fbw[0] + 2) + 31) & ~(size_t)31;
	}
	for (int i = 0; i < n; i++) {
		fbw[i] = (chain[i].dimx + 3) / 4; fbh[i] = (chain[i].dimy + 3) / 4;
		boff[i] = off; off += (((size_t)chain[i].stride * chain[i].dimy * (chain[i].is_int ? 4 : 2)) + 31) & ~(size_t)31;
		foff[i] = off; off += ((size_t)fbw[i] * fbh[i] + 64 + 31) & ~(size_t)31;
	}
	char *d = nullptr;
	cudaError_t e = cudaMalloc(&d, off + 64);
	if (e == cudaSuccess) e = cudaMemset(d, 0, off + 64);
	if (e == cudaSuccess && !high_band) {
		// only the rows / columns the band's full blocks look at need to exist; the caller's array covers them
		const size_t need = (size_t)child_bw * std::min(2 * fbh[0], (chain[0].dimy / 4) * 2);
		e = cudaMemcpy(d + cf_off, child_flags, need, cudaMemcpyHostToDevice);
	}
	for (int i = 0; i < n && e == cudaSuccess; i++)
		e = cudaMemcpy(d + boff[i], chain[i].data, (size_t)chain[i].stride * chain[i].dimy * (chain[i].is_int ? 4 : 2), cudaMemcpyHostToDevice);
	for (int i = 0; i < n && e == cudaSuccess; i++) {
		QuantLevelParams P;
		memset(&P, 0, sizeof P);
		P.arena = d;
		P.flags = (unsigned char *)d;
		P.has_child = i > 0 || !high_band;
		P.is_int = chain[i].is_int;
		P.o_first = 0;
		BandRef &b = P.band[0];
		b.off = (long long)boff[i]; b.dimx = chain[i].dimx; b.dimy = chain[i].dimy; b.stride = chain[i].stride;
		b.fl_off = (int)foff[i]; b.fl_bw = fbw[i];
		if (P.has_child) {
			BandRef &ch = P.child[0];
			ch.fl_off = i > 0 ? (int)foff[i - 1] : (int)cf_off;
			ch.fl_bw = i > 0 ? fbw[i - 1] : child_bw;
		}
		fill_qb(P.qb[0], Quant, lambda, chain[i].weight, chain[i].is_int);
		dim3 grid((fbw[i] * fbh[i] + 127) / 128, 1);
		if (chain[i].is_int) quant_level_kernel<false><<<grid, 128>>>(P);
		else quant_level_kernel<true><<<grid, 128>>>(P);
		e = cudaGetLastError();
	}
	for (int i = 0; i < n && e == cudaSuccess; i++) {
		e = cudaMemcpy(chain[i].data, d + boff[i], (size_t)chain[i].stride * chain[i].dimy * (chain[i].is_int ? 4 : 2), cudaMemcpyDeviceToHost);
		if (e == cudaSuccess && chain[i].flags) e = cudaMemcpy(chain[i].flags, d + foff[i], (size_t)fbw[i] * fbh[i], cudaMemcpyDeviceToHost);
	}
	cudaFree(d);
	if (e != cudaSuccess) return set_err(RIC_E_CUDA, "ric_buf_build_tree: %s", cudaGetErrorString(e));
	return RIC_OK;
}

int ric_transform_inv(ric_ctx *c, const void *arena, int16_t *plane, int stride)
{
	if (!c || !arena || !plane) return set_err(RIC_E_ARG, "ric_transform_inv: null");
	const HostGeom &g = c->g;
	if (stride < g.width) return set_err(RIC_E_ARG, "ric_transform_inv: stride < width");
	CK(cudaSetDevice(c->device));
	int rc = need_arena_in(c);
	if (rc) return rc;
	CK(cudaMemcpyAsync(c->d_arena_in, arena, g.arena_bytes, cudaMemcpyHostToDevice, c->stream));
	int Q[3] = {0, 0, 0};
	rc = launch_inverse(c, c->d_arena_in, 1, 1, 0, Q, DST_PLANE, c->d_plane, 0, 0, c->plane_pitch, 1, c->stream);
	if (rc) return rc;
	CK(cudaMemcpy2DAsync(plane, (size_t)stride * 2, c->d_plane, (size_t)c->plane_pitch * 2, (size_t)g.width * 2, g.height,
	                     cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return RIC_OK;
}

// ---- host entropy stage (no GPU) ------------------------------------------------------------------

static int entropy_geom(HostGeom &g, int w, int h, int ch, int levels, int level_chg, int align, const char *who)
{
	const int rc = geom_init(g, w, h, ch, levels, level_chg, align, RIC_CDF97);  // the transform does not shape the bands
	return rc ? set_err(rc, "%s: unsupported geometry", who) : RIC_OK;
}

int ric_entropy_encode(int width, int height, int channels, int levels, int level_chg, int align, void *image_arena,
                       uint8_t *out, size_t cap, size_t *size)
{
	HostGeom g;
	int rc = entropy_geom(g, width, height, channels, levels, level_chg, align, "ric_entropy_encode");
	if (rc) return rc;
	if (!image_arena || !out || !size) return set_err(RIC_E_ARG, "ric_entropy_encode: null");
	const long n = entropy_encode_image(g, (char *)image_arena, out, cap);
	if (n < 0) return set_err(RIC_E_NOMEM, "ric_entropy_encode: output buffer too small");
	*size = (size_t)n;
	return RIC_OK;
}

int ric_entropy_encode_hinted(int width, int height, int channels, int levels, int level_chg, int align, const void *image_arena,
                              uint8_t *out, size_t cap, size_t *size)
{
	HostGeom g;
	int rc = entropy_geom(g, width, height, channels, levels, level_chg, align, "ric_entropy_encode_hinted");
	if (rc) return rc;
	if (!image_arena || !out || !size) return set_err(RIC_E_ARG, "ric_entropy_encode_hinted: null");
	const long n = entropy_encode_image_hinted(g, (const char *)image_arena, out, cap);
	if (n < 0) return set_err(RIC_E_NOMEM, "ric_entropy_encode_hinted: output buffer too small");
	*size = (size_t)n;
	return RIC_OK;
}

int ric_entropy_decode(int width, int height, int channels, int levels, int level_chg, int align, const uint8_t *payload,
                       size_t size, void *image_arena)
{
	HostGeom g;
	int rc = entropy_geom(g, width, height, channels, levels, level_chg, align, "ric_entropy_decode");
	if (rc) return rc;
	if (!image_arena || !payload) return set_err(RIC_E_ARG, "ric_entropy_decode: null");
	if (entropy_decode_image(g, payload, size, (char *)image_arena)) return set_err(RIC_E_ARG, "ric_entropy_decode: truncated payload");
	return RIC_OK;
}

struct ric_mux {
	MuxEncoder *enc;
	MuxDecoder *dec;
};

int ric_mux_encoder(ric_mux **mux, uint8_t *stream, size_t cap, unsigned first_word)
{
	if (!mux || !stream || first_word > 0xFFFF) return set_err(RIC_E_ARG, "ric_mux_encoder: bad argument");
	MuxEncoder *e = mux_encoder_new(stream, cap, first_word);
	if (!e) return set_err(RIC_E_NOMEM, "ric_mux_encoder: buffer too small");
	*mux = new ric_mux{e, nullptr};
	return RIC_OK;
}

int ric_mux_decoder(ric_mux **mux, const uint8_t *stream, size_t size)
{
	if (!mux || !stream) return set_err(RIC_E_ARG, "ric_mux_decoder: null");
	MuxDecoder *d = mux_decoder_new(stream, size);
	if (!d) return set_err(RIC_E_ARG, "ric_mux_decoder: stream too short");
	*mux = new ric_mux{nullptr, d};
	return RIC_OK;
}

int ric_mux_code_plane(ric_mux *mux, int width, int height, int levels, int level_chg, int align, void *plane_arena)
{
	HostGeom g;
	if (!mux || !mux->enc || !plane_arena) return set_err(RIC_E_ARG, "ric_mux_code_plane: not an encoder / null arena");
	int rc = entropy_geom(g, width, height, 1, levels, level_chg, align, "ric_mux_code_plane");
	if (rc) return rc;
	mux_encoder_plane(mux->enc, g, (char *)plane_arena);
	return RIC_OK;
}

int ric_mux_decode_plane(ric_mux *mux, int width, int height, int levels, int level_chg, int align, void *plane_arena)
{
	HostGeom g;
	if (!mux || !mux->dec || !plane_arena) return set_err(RIC_E_ARG, "ric_mux_decode_plane: not a decoder / null arena");
	int rc = entropy_geom(g, width, height, 1, levels, level_chg, align, "ric_mux_decode_plane");
	if (rc) return rc;
	if (mux_decoder_plane(mux->dec, g, (char *)plane_arena)) return set_err(RIC_E_ARG, "ric_mux_decode_plane: truncated stream");
	return RIC_OK;
}

int ric_mux_finish(ric_mux *mux, size_t *end)
{
	if (!mux || !mux->enc || !end) return set_err(RIC_E_ARG, "ric_mux_finish: not an encoder / null");
	const long n = mux_encoder_finish(mux->enc);
	if (n < 0) return set_err(RIC_E_NOMEM, "ric_mux_finish: stream buffer too small");
	*end = (size_t)n;
	return RIC_OK;
}

int ric_mux_destroy(ric_mux *mux)
{
	if (!mux) return RIC_OK;
	if (mux->enc) mux_encoder_free(mux->enc);
	if (mux->dec) mux_decoder_free(mux->dec);
	delete mux;
	return RIC_OK;
}

// ---- entropy stage on the device (large batches; one image per warp) ---------------------------------

static int need_entropy_tables(ric_ctx *c)
{
	if (c->d_tables) return RIC_OK;
	size_t bytes = 0;
	const void *host = entropy_tables(&bytes);
	CK(cudaMalloc(&c->d_tables, bytes));
	CK(cudaMemcpy(c->d_tables, host, bytes, cudaMemcpyHostToDevice));
	CK(cudaMalloc(&c->d_geom, sizeof(HostGeom)));
	CK(cudaMemcpy(c->d_geom, &c->g, sizeof(HostGeom), cudaMemcpyHostToDevice));
	CK(cudaMalloc(&c->d_bad, sizeof(int)));
	CK(cudaMemset(c->d_bad, 0, sizeof(int)));
	cudaError_t e = cudaMalloc(&c->d_hints, hint_bytes(c->g, c->max_batch));
	if (e != cudaSuccess) return set_err(RIC_E_NOMEM, "cudaMalloc(block hints): %s", cudaGetErrorString(e));
	return RIC_OK;
}

int ric_entropy_encode_device(ric_ctx *c, void *d_arenas, int n, uint8_t *d_out, size_t stride, long long *d_sizes, void *stream)
{
	if (!c || !d_arenas || !d_out || !d_sizes || n < 1 || stride < 8) return set_err(RIC_E_ARG, "ric_entropy_encode_device: bad argument");
	CK(cudaSetDevice(c->device));
	int rc = need_entropy_tables(c);
	if (rc) return rc;
	if (n > c->max_batch) return set_err(RIC_E_ARG, "ric_entropy_encode_device: n > max_batch");
	if (getenv("RIC_ENTROPY_PLAIN"))  // the plain walker (in-place marker bookkeeping, no pre-pass), for comparison
		CK(launch_entropy_encode(c->d_geom, c->d_tables, (char *)d_arenas, (size_t)c->g.channels * c->g.arena_bytes, d_out, stride, d_sizes, n,
		                         (cudaStream_t)stream));
	else
		CK(launch_entropy_encode_hinted(c->d_geom, c->g, c->d_tables, (char *)d_arenas, (size_t)c->g.channels * c->g.arena_bytes, c->d_hints,
		                                d_out, stride, d_sizes, n, (cudaStream_t)stream));
	return RIC_OK;
}

int ric_entropy_decode_device(ric_ctx *c, const uint8_t *d_payloads, size_t stride, const long long *d_sizes, int n, void *d_arenas,
                              int *d_status, void *stream)
{
	if (!c || !d_arenas || !d_payloads || !d_sizes || n < 1) return set_err(RIC_E_ARG, "ric_entropy_decode_device: bad argument");
	CK(cudaSetDevice(c->device));
	int rc = need_entropy_tables(c);
	if (rc) return rc;
	const size_t img_ar = (size_t)c->g.channels * c->g.arena_bytes;
	CK(cudaMemsetAsync(d_arenas, 0, (size_t)n * img_ar, (cudaStream_t)stream));
	CK(cudaMemsetAsync(c->d_bad, 0, sizeof(int), (cudaStream_t)stream));  // (the context-wide flag of ric_decompress_u8_gpu)
	CK(launch_entropy_decode(c->d_geom, c->d_tables, d_payloads, stride, d_sizes, (char *)d_arenas, img_ar, c->d_bad, d_status, n,
	                         (cudaStream_t)stream));
	return RIC_OK;
}

// ---- whole .ric files with the entropy stage on the device (large batches) -----------------------------

static int need_payload(ric_ctx *c, size_t pstride)
{
	int rc = need_entropy_tables(c);
	if (rc) return rc;
	if (!c->ent[0])
		for (int i = 0; i < 8; i++) {
			CK(cudaStreamCreateWithFlags(&c->ent[i], cudaStreamNonBlocking));
			CK(cudaEventCreateWithFlags(&c->ent_ev[i], cudaEventDisableTiming));
		}
	if (!c->d_psizes) {
		CK(cudaMalloc(&c->d_psizes, sizeof(long long) * c->max_batch));
		CK(cudaHostAlloc((void **)&c->h_psizes, sizeof(long long) * c->max_batch, cudaHostAllocDefault));
	}
	if (pstride > c->payload_stride) {
		cudaFree(c->d_payload);
		c->d_payload = nullptr;
		c->payload_stride = 0;
		cudaError_t e = cudaMalloc(&c->d_payload, (size_t)c->max_batch * pstride);
		if (e != cudaSuccess) return set_err(RIC_E_NOMEM, "cudaMalloc(payload slots): %s", cudaGetErrorString(e));
		c->payload_stride = pstride;
	}
	return RIC_OK;
}

// Chunks of the device-entropy calls: at most four.  Streams are multiplexed onto a few hardware queues (eight
// by default, shared with every other stream of the process); with eight entropy streams, stream i+4 was
// observed to queue behind stream i's 300 ms kernel.
static int ent_chunk_images(int n) { return n >= 8 ? (n + 3) / 4 : n >= 2 ? (n + 1) / 2 : 1; }

static int sync_ent(ric_ctx *c)
{
	for (int i = 0; i < 8; i++) CK(cudaStreamSynchronize(c->ent[i]));
	return RIC_OK;
}

int ric_compress_u8_gpu(ric_ctx *c, const uint8_t *src, int n, int q, uint8_t *files, size_t stride, size_t *sizes)
{
	int rc = check_batch(c, n, q, "ric_compress_u8_gpu");
	if (rc) return rc;
	if (!src || !files || !sizes) return set_err(RIC_E_ARG, "ric_compress_u8_gpu: null buffer");
	if (stride < RIC_HEADER_BYTES + 16) return set_err(RIC_E_NOMEM, "ric_compress_u8_gpu: stride too small");
	const HostGeom &g = c->g;
	CK(cudaSetDevice(c->device));
	const size_t img_px = (size_t)g.channels * g.height * g.width, img_dev = (size_t)g.channels * g.height * c->src_pitch;
	const size_t img_ar = (size_t)g.channels * g.arena_bytes;
	// device payload slots follow the caller's stride (like the host path), up to 2 x W*H*C + 4096 -- twice the
	// reference's own buffer (W*H*C, ric.cpp:131-132; incompressible input at q = 0 can exceed it)
	const size_t pstride = std::min(stride - RIC_HEADER_BYTES, 2 * img_px + 4096) & ~(size_t)15;
	if ((rc = need_payload(c, pstride))) return rc;
	if ((rc = sync_pipe(c))) return rc;
	// Pixel copies + encode stage chunk by chunk on the pipeline streams, then ONE entropy launch over the whole
	// batch: the stage is latency-bound per image, so it wants every image in flight at once, and entropy kernels
	// started per chunk were observed to hold back the encode stages of the later chunks (RIC_TRACE timeline).
	const int step = chunk_images(n);
	int total = 0, k = 0;
	const bool trace = getenv("RIC_TRACE") != nullptr;
	const auto t_start = std::chrono::steady_clock::now();
	auto stamp = [&](const char *what, int idx) {
		if (trace) fprintf(stderr, "[ric trace] %s %d at %.1f ms\n", what, idx,
		                   1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
	};
	for (int i0 = 0; i0 < n; i0 += step, k++) {
		const int m = std::min(step, n - i0);
		cudaStream_t st = c->pipe[k % 3];
		CK(copy_pixels(c->d_src + i0 * img_dev, c->src_pitch, src + i0 * img_px, g.width, g.width,
		               (size_t)m * g.channels * g.height, cudaMemcpyHostToDevice, st));
		c->img0 = i0;
		c->cset = k % 3;
		rc = ric_encode_u8_device(c, c->d_src + i0 * img_dev, c->src_pitch, m, q, c->d_arena + i0 * img_ar, st);
		c->img0 = 0;
		c->cset = 3;
		if (rc) { sync_pipe(c); sync_ent(c); return rc; }
		total += c->launches;
	}
	cudaStream_t es = c->ent[0];
	for (int i = 0; i < 3; i++) {
		CK(cudaEventRecord(c->ent_ev[i], c->pipe[i]));
		CK(cudaStreamWaitEvent(es, c->ent_ev[i], 0));
	}
	CK(launch_entropy_encode_hinted(c->d_geom, c->g, c->d_tables, c->d_arena, img_ar, c->d_hints, c->d_payload, c->payload_stride, c->d_psizes,
	                                n, es));
	CK(cudaMemcpyAsync(c->h_psizes, c->d_psizes, sizeof(long long) * n, cudaMemcpyDeviceToHost, es));
	c->launches = total + 1;
	stamp("enqueued", 0);
	if (trace) { sync_pipe(c); stamp("encode stage done", n); }
	CK(cudaStreamSynchronize(es));
	stamp("entropy coded", n);
	// fetch exactly the bytes of every file
	bool small = false;
	for (int i = 0; i < n; i++) {
		uint8_t *f = files + (size_t)i * stride;
		ric_header_write(f, g.width, g.height, q, g.channels == 3, g.trans);
		if (c->h_psizes[i] < 0) { small = true; sizes[i] = 0; continue; }
		sizes[i] = (size_t)c->h_psizes[i] + RIC_HEADER_BYTES;
		cudaError_t ce = cudaMemcpyAsync(f + RIC_HEADER_BYTES, c->d_payload + (size_t)i * c->payload_stride, (size_t)c->h_psizes[i],
		                                 cudaMemcpyDeviceToHost, c->pipe[i % 3]);
		if (ce != cudaSuccess) {  // copies into the caller's buffer are in flight: drain them before returning
			sync_pipe(c);
			return set_err(RIC_E_CUDA, "ric_compress_u8_gpu: %s", cudaGetErrorString(ce));
		}
	}
	if ((rc = sync_pipe(c))) return rc;
	stamp("files fetched", n);
	if (small) return set_err(RIC_E_NOMEM, "ric_compress_u8_gpu: a file did not fit in its slot (min(stride, 2*W*H*C + 4096) bytes)");
	return RIC_OK;
}

int ric_decompress_u8_gpu(ric_ctx *c, const uint8_t *files, size_t stride, const size_t *sizes, int n, uint8_t *dst)
{
	int rc = check_batch(c, n, 0, "ric_decompress_u8_gpu");
	if (rc) return rc;
	if (!files || !sizes || !dst) return set_err(RIC_E_ARG, "ric_decompress_u8_gpu: null buffer");
	const HostGeom &g = c->g;
	int q = -1;
	size_t longest = 16;
	for (int i = 0; i < n; i++) {
		int w, h, qi, color, trans;
		if (sizes[i] < RIC_HEADER_BYTES || sizes[i] > stride) return set_err(RIC_E_ARG, "ric_decompress_u8_gpu: bad file size");
		if ((rc = ric_header_parse(files + (size_t)i * stride, &w, &h, &qi, &color, &trans))) return rc;
		if (w != g.width || h != g.height || color != (g.channels == 3) || trans != g.trans)
			return set_err(RIC_E_ARG, "ric_decompress_u8_gpu: file header does not match the context (size, colour or transform)");
		if (q >= 0 && qi != q) return set_err(RIC_E_ARG, "ric_decompress_u8_gpu: files of one batch must share the quantiser index");
		q = qi;
		longest = std::max(longest, sizes[i] - RIC_HEADER_BYTES);
	}
	CK(cudaSetDevice(c->device));
	if ((rc = need_payload(c, (longest + 15) & ~(size_t)15))) return rc;
	if ((rc = need_arena_in(c))) return rc;
	if ((rc = sync_pipe(c))) return rc;
	const size_t img_px = (size_t)g.channels * g.height * g.width, img_dev = (size_t)g.channels * g.height * c->src_pitch;
	const size_t img_ar = (size_t)g.channels * g.arena_bytes;
	CK(cudaMemsetAsync(c->d_bad, 0, sizeof(int), c->ent[0]));
	CK(cudaStreamSynchronize(c->ent[0]));
	const int step = ent_chunk_images(n);
	int total = 0, k = 0;
	for (int i0 = 0; i0 < n; i0 += step, k++) {
		const int m = std::min(step, n - i0);
		cudaStream_t es = c->ent[k];
		for (int i = i0; i < i0 + m; i++) {
			c->h_psizes[i] = (long long)(sizes[i] - RIC_HEADER_BYTES);
			CK(cudaMemcpyAsync(c->d_payload + (size_t)i * c->payload_stride, files + (size_t)i * stride + RIC_HEADER_BYTES,
			                   sizes[i] - RIC_HEADER_BYTES, cudaMemcpyHostToDevice, es));
		}
		CK(cudaMemcpyAsync(c->d_psizes + i0, c->h_psizes + i0, sizeof(long long) * m, cudaMemcpyHostToDevice, es));
		CK(cudaMemsetAsync(c->d_arena_in + i0 * img_ar, 0, (size_t)m * img_ar, es));
		CK(launch_entropy_decode(c->d_geom, c->d_tables, c->d_payload + (size_t)i0 * c->payload_stride, c->payload_stride, c->d_psizes + i0,
		                         c->d_arena_in + i0 * img_ar, img_ar, c->d_bad, nullptr, m, es));
		CK(cudaEventRecord(c->ent_ev[k], es));
	}
	k = 0;
	for (int i0 = 0; i0 < n; i0 += step, k++) {
		const int m = std::min(step, n - i0);
		cudaStream_t st = c->pipe[k % 3];
		CK(cudaStreamWaitEvent(st, c->ent_ev[k], 0));
		c->img0 = i0;
		c->cset = k % 3;
		rc = ric_decode_u8_device(c, c->d_arena_in + i0 * img_ar, m, q, c->d_src + i0 * img_dev, c->src_pitch, st);
		c->img0 = 0;
		c->cset = 3;
		if (rc) { sync_ent(c); sync_pipe(c); return rc; }
		total += c->launches + 1;
		CK(copy_pixels(dst + i0 * img_px, g.width, c->d_src + i0 * img_dev, c->src_pitch, g.width,
		               (size_t)m * g.channels * g.height, cudaMemcpyDeviceToHost, st));
	}
	c->launches = total;
	if ((rc = sync_ent(c))) return rc;
	if ((rc = sync_pipe(c))) return rc;
	int bad = 0;
	CK(cudaMemcpy(&bad, c->d_bad, sizeof(int), cudaMemcpyDeviceToHost));
	if (bad) return set_err(RIC_E_ARG, "ric_decompress_u8_gpu: truncated payload");
	return RIC_OK;
}

// ---- whole .ric files: GPU stage + host entropy threads ----------------------------------------------

static int need_stage(ric_ctx *c)
{
	if (c->h_stage) return RIC_OK;
	const size_t bytes = (size_t)c->max_batch * c->g.channels * c->g.arena_bytes;
	cudaError_t e = cudaHostAlloc((void **)&c->h_stage, bytes, cudaHostAllocDefault);
	if (e != cudaSuccess) { c->h_stage = nullptr; return set_err(RIC_E_NOMEM, "cudaHostAlloc(staging arenas): %s", cudaGetErrorString(e)); }
	return RIC_OK;
}

static int worker_count(int threads, int n)
{
	if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
	return std::max(1, std::min(threads, n));
}

int ric_compress_u8(ric_ctx *c, const uint8_t *src, int n, int q, uint8_t *files, size_t stride, size_t *sizes, int threads)
{
	int rc = check_batch(c, n, q, "ric_compress_u8");
	if (rc) return rc;
	if (!src || !files || !sizes) return set_err(RIC_E_ARG, "ric_compress_u8: null buffer");
	if (stride < RIC_HEADER_BYTES + 4) return set_err(RIC_E_NOMEM, "ric_compress_u8: stride too small");
	const HostGeom &g = c->g;
	CK(cudaSetDevice(c->device));
	if ((rc = need_stage(c))) return rc;
	const size_t img_ar = (size_t)g.channels * g.arena_bytes;
	LandedQueue queue;
	queue.total = n;
	std::atomic<int> failed{0}, cancelled{0};
	std::vector<std::thread> pool;
	const int nw = worker_count(threads, n);
	for (int t = 0; t < nw; t++)
		pool.emplace_back([&] {
			for (int i; (i = queue.take()) >= 0;) {
				if (cancelled) break;  // the GPU stage failed: nothing valid to code
				uint8_t *f = files + (size_t)i * stride;
				ric_header_write(f, g.width, g.height, q, g.channels == 3, g.trans);
				const long sz = entropy_encode_image(g, c->h_stage + (size_t)i * img_ar, f + RIC_HEADER_BYTES, stride - RIC_HEADER_BYTES);
				if (sz < 0) { failed = 1; sizes[i] = 0; } else sizes[i] = (size_t)sz + RIC_HEADER_BYTES;
			}
		});
	rc = ric_encode_u8_stream(c, src, n, q, c->h_stage, LandedQueue::landed, &queue);
	int rc2 = sync_pipe(c);
	if (rc || rc2) { cancelled = 1; queue.release_all(); }  // let the workers drain
	for (auto &t : pool) t.join();
	if (rc) return rc;
	if (rc2) return rc2;
	if (failed) return set_err(RIC_E_NOMEM, "ric_compress_u8: a file did not fit in `stride` bytes");
	return RIC_OK;
}

int ric_decompress_u8(ric_ctx *c, const uint8_t *files, size_t stride, const size_t *sizes, int n, uint8_t *dst, int threads)
{
	int rc = check_batch(c, n, 0, "ric_decompress_u8");
	if (rc) return rc;
	if (!files || !sizes || !dst) return set_err(RIC_E_ARG, "ric_decompress_u8: null buffer");
	const HostGeom &g = c->g;
	int q = -1;
	for (int i = 0; i < n; i++) {
		int w, h, qi, color, trans;
		if (sizes[i] < RIC_HEADER_BYTES || sizes[i] > stride) return set_err(RIC_E_ARG, "ric_decompress_u8: bad file size");
		if ((rc = ric_header_parse(files + (size_t)i * stride, &w, &h, &qi, &color, &trans))) return rc;
		if (w != g.width || h != g.height || color != (g.channels == 3) || trans != g.trans)
			return set_err(RIC_E_ARG, "ric_decompress_u8: file header does not match the context (size, colour or transform)");
		if (q >= 0 && qi != q) return set_err(RIC_E_ARG, "ric_decompress_u8: files of one batch must share the quantiser index");
		q = qi;
	}
	CK(cudaSetDevice(c->device));
	if ((rc = need_stage(c))) return rc;
	if ((rc = need_arena_in(c))) return rc;
	if ((rc = sync_pipe(c))) return rc;
	const size_t img_ar = (size_t)g.channels * g.arena_bytes;
	// Workers take the files in order; as soon as every image of a chunk is decoded the main thread sends that
	// chunk through the decode stage, while the workers go on with the later chunks.
	const int step = chunk_images(n), nchunks = (n + step - 1) / step;
	std::vector<int> left(nchunks);
	for (int k = 0; k < nchunks; k++) left[k] = std::min(step, n - k * step);
	std::mutex mu;
	std::condition_variable cv;
	std::atomic<int> next{0}, failed{0};
	std::vector<std::thread> pool;
	const int nw = worker_count(threads, n);
	for (int t = 0; t < nw; t++)
		pool.emplace_back([&] {
			for (int i; (i = next.fetch_add(1)) < n;) {
				const uint8_t *f = files + (size_t)i * stride;
				if (entropy_decode_image(g, f + RIC_HEADER_BYTES, sizes[i] - RIC_HEADER_BYTES, c->h_stage + (size_t)i * img_ar)) failed = 1;
				bool chunk_done;
				{ std::lock_guard<std::mutex> l(mu); chunk_done = --left[i / step] == 0; }
				if (chunk_done) cv.notify_all();
			}
		});
	int total = 0;
	for (int k = 0; k < nchunks && !rc; k++) {
		{ std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return left[k] == 0; }); }
		if (failed) break;  // a truncated file: its arenas are not worth decoding
		rc = decode_chunk(c, c->h_stage, k * step, std::min(step, n - k * step), q, dst, k);
		total += c->launches;
	}
	for (auto &t : pool) t.join();
	const int rc2 = sync_pipe(c);
	c->launches = total;
	if (rc) return rc;
	if (failed) return set_err(RIC_E_ARG, "ric_decompress_u8: truncated payload");
	return rc2;
}

}  // extern "C"
