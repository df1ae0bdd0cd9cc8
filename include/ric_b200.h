/* ric_b200.h -- C ABI of the B200-native RIC transform+quantisation front end.
 *
 * This is the drop-in boundary for the one data-parallel hot path of rududu/RIC:
 *   colour / level shift      src/ric/ric.cpp:76-112,143-148,227-246
 *   CWavelet2D::Transform     src/lib/wavelet2d.cpp:926-958  (+Transform97/53 :407-492,:636-692)
 *   CBandCodec::buildTree     src/lib/bandcodec.cpp:239-322  (encode quantiser, folded output)
 *   CBand::TSUQ / TSUQi       src/lib/band.h:65-107
 *   CWavelet2D::TransformI    src/lib/wavelet2d.cpp:960-992  (+Transform97I/53I :494-591,:694-764)
 * The reference has no FFI layer: its boundary is the C++ class API (CWavelet2D/CBand).  The C++
 * shim in include/rududu_b200/ re-creates those classes on top of the functions declared here;
 * INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions: every function returns 0 on success or a negative RIC_E_* code (never throws).
 * All kernels are hand-written sm_100a CUDA; there is NO CPU fallback: if no CUDA device is
 * usable, ric_create fails with RIC_E_CUDA.
 *
 * Band ("canonical") order used everywhere: id = 3*lev + {0:D, 1:H, 2:V}, lev 0 = finest level,
 * id = 3*nlev = the coarsest level's LL band.  A plane's bands live in one "arena": each band
 * row-major with the reference's DimXAlign stride (band.cpp:57) at a 32-byte aligned offset --
 * exactly the buffers the reference's entropy stage (CBandCodec::pred/tree) consumes.
 * An image with C planes has C consecutive arenas, plane order = CImg channel order after the
 * colour transform: 0 = Co, 1 = Cg, 2 = Y (ric.cpp:76-91), or the single gray plane.
 */
#ifndef RIC_B200_H
#define RIC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIC_MAX_LEVELS 16
#define RIC_MAX_BANDS (3 * RIC_MAX_LEVELS + 1)

enum { RIC_CDF97 = 0, RIC_CDF53 = 1, RIC_HAAR = 2 }; /* enum trans, src/lib/utils.h:28 */

enum {
	RIC_OK = 0,
	RIC_E_ARG = -1,      /* invalid argument / unsupported geometry */
	RIC_E_CUDA = -2,     /* CUDA runtime error (ric_last_error() has the text) */
	RIC_E_NOMEM = -3,
	RIC_E_UNSUPPORTED = -4
};

typedef struct ric_ctx ric_ctx;

typedef struct {
	int dimx, dimy;   /* CBand::DimX, DimY       band.h:43-44 */
	int stride;       /* CBand::DimXAlign        band.cpp:57 (samples) */
	int is_int;       /* band_t sint             band.h:35 */
	float weight;     /* CBand::Weight           wavelet2d.cpp:1009-1032 */
	size_t offset;    /* byte offset inside one plane arena */
} ric_band_info;

typedef struct {
	int width, height, channels;
	int levels, level_chg, align, trans;
	int nlev;          /* levels actually built (wavelet2d.cpp:76) */
	int nbands;        /* 3*nlev + 1 */
	int max_batch;
	size_t arena_bytes;       /* one plane */
	size_t image_arena_bytes; /* channels * arena_bytes */
} ric_info;

/* ---- context -------------------------------------------------------------------------------
 * Replaces CWavelet2D::CWavelet2D / Init (wavelet2d.cpp:38-81), CBand::Init (band.cpp:51-65) and
 * SetWeight (wavelet2d.cpp:1009-1032) for `channels` planes of up to `max_batch` images on CUDA
 * device `device`.  trans selects the lifting: cdf97, cdf53, or haar (haar only when every level has
 * even width and height -- the reference's Haar leaves odd trailing rows/columns undefined -- else
 * RIC_E_UNSUPPORTED). */
int ric_create(ric_ctx **out, int device, int width, int height, int channels, int levels,
               int level_chg, int align, int trans, int max_batch);
int ric_destroy(ric_ctx *ctx);
int ric_get_info(const ric_ctx *ctx, ric_info *info);
int ric_get_band(const ric_ctx *ctx, int band_id, ric_band_info *info);
const char *ric_last_error(void);

/* Quants(), ric.cpp:42-49, and the per-plane (Quant, lambda) pair ric derives from -q
 * (ric.cpp:163-171): plane `plane` of `channels`. */
int ric_quants(int idx);
int ric_plane_quant(int q, int channels, int plane, int *Quant, int *lambda);

/* ---- whole-stage entry points, HOST buffers (H2D + kernels + D2H inside) ---------------------
 * ric_encode_u8: what CompressImage does between loading the image and the first entropy call:
 *   colour/level shift + Transform + buildTree x3 + LL TSUQ for every plane of n images.
 *   src: n planar u8 images (channels*height*width each, densely packed).
 *   arenas: n * image_arena_bytes, filled with quantised bands ready for CBandCodec::pred/tree.
 *   q: ric's -q value 0..31 (0 = lossless: only folding).  Quant/lambda as ric.cpp:163-171.
 * ric_decode_u8: what DecompressImage does after the last DecodeBand: TSUQi + TransformI +
 *   inverse colour + clip.  arenas hold signed quantised coefficients (DecodeBand output). */
int ric_encode_u8(ric_ctx *ctx, const uint8_t *src, int n, int q, void *arenas);
int ric_decode_u8(ric_ctx *ctx, const void *arenas, int n, int q, uint8_t *dst);

/* ---- streaming variants: overlap with host entropy threads ------------------------------------
 * ric_encode_u8_stream returns as soon as every chunk of the batch is enqueued (H2D, kernels, D2H
 * alternate over three internal streams).  `done(user, first_image, n_images)` is called once per
 * chunk, from a CUDA callback thread, when the arenas of those images have landed in `arenas`
 * (chunks alternate over three streams and may complete OUT OF ORDER: act on exactly [first, first + n))
 * (use pinned memory, ric_host_alloc): the callback must not call CUDA or ric_* functions -- it
 * hands the chunk to a host worker, e.g. one running the reference's entropy half of CodeBand
 * (wavelet2d.cpp:119-159) while the GPU works on the next chunk.  ric_sync waits for everything.
 * ric_decode_u8_stream is the mirror image (done fires when the pixels of a chunk are in `dst`).
 * ric_encode_u8 / ric_decode_u8 are these calls followed by ric_sync. */
typedef void (*ric_chunk_fn)(void *user, int first_image, int n_images);
int ric_encode_u8_stream(ric_ctx *ctx, const uint8_t *src, int n, int q, void *arenas, ric_chunk_fn done, void *user);
int ric_decode_u8_stream(ric_ctx *ctx, const void *arenas, int n, int q, uint8_t *dst, ric_chunk_fn done, void *user);
int ric_sync(ric_ctx *ctx);

/* ---- device-resident variants (no copies; asynchronous on `stream`, a cudaStream_t) ------------
 * d_src pitch: bytes between rows (multiple of 8); planes are pitch*height apart, images
 * channels*pitch*height apart.  d_arenas as above but in device memory.  d_dst likewise.
 * CONCURRENCY: a context is ONE CWavelet2D-like object -- its LL scratch planes, block flags and job counters are
 * shared by every call.  Keep at most one call per context in flight: issue the *_device calls of a context on
 * one stream (or order them yourself with events), and let a *_stream / ric_compress / plane-level call return
 * (ric_sync for the *_stream pair) before the next call on the same context.  Independent work goes on separate
 * contexts, which share nothing (bench.py runs the encode and the decode stage concurrently on two). */
int ric_encode_u8_device(ric_ctx *ctx, const uint8_t *d_src, size_t pitch, int n, int q,
                         void *d_arenas, void *stream);
int ric_decode_u8_device(ric_ctx *ctx, const void *d_arenas, int n, int q, uint8_t *d_dst,
                         size_t pitch, void *stream);
/* number of kernel launches the last *_device call enqueued (for bench accounting) */
int ric_last_launch_count(const ric_ctx *ctx);

/* Optional per-launch timing: when on, a CUDA event is recorded on the launching stream around every
 * level kernel of the *_device calls.  ric_get_level_times waits for the last recorded call of the
 * given direction (0 encode: finest level first; 1 decode: coarsest level first) and writes one
 * duration per launch in milliseconds; returns the number of launches (<= RIC_MAX_LEVELS) or < 0. */
int ric_set_profiling(ric_ctx *ctx, int on);
int ric_get_level_times(ric_ctx *ctx, int direction, float *ms, int cap);
/* Path statistics of the packed (two samples per register) kernels, counted while profiling is on and reset
 * by this call: out[0] = plane iterations (one row pair of one plane of one strip) run by the packed forward
 * kernel, out[1] = how many of them took the scalar, exactly-wrapping path (image edge rows, or operands outside
 * the bounds the packed arithmetic needs); out[2] = plane iterations of the packed inverse kernel, out[3] / out[4] = how
 * many of their column / row passes took the scalar path.  n <= 8. */
int ric_get_path_stats(ric_ctx *ctx, unsigned long long *out, int n);

/* ---- plane-level entry points mirroring the reference class API (HOST buffers) -------------------
 * One plane at a time on batch slot 0, like a CWavelet2D object:
 * ric_transform:   CWavelet2D::Transform<short>(pImage, Stride, t)  wavelet2d.cpp:926.
 *                  Unlike the reference the caller's plane is left untouched. arena receives the
 *                  unquantised bands (may be NULL to keep them on the device only).
 * ric_quant:       the quantiser half of CodeBand (wavelet2d.cpp:110-126): buildTree x3 + LL
 *                  TSUQ(0.5) on the device-resident bands of the last ric_transform; result -> arena.
 * ric_tsuq:        CWavelet2D::TSUQ(Quant, Thres)   wavelet2d.cpp:224-246; *count gets Count.
 * ric_tsuqi:       CWavelet2D::TSUQi(Quant)          wavelet2d.cpp:248-268 on arena (in/out, host).
 * ric_transform_inv: CWavelet2D::TransformI wavelet2d.cpp:960; takes the plane START (not the
 *                  one-past-end pointer of the reference) and reads the bands from arena. */
int ric_transform(ric_ctx *ctx, const int16_t *plane, int stride, void *arena);
int ric_quant(ric_ctx *ctx, int Quant, int lambda, void *arena);
int ric_tsuq(ric_ctx *ctx, int Quant, float thres, void *arena, unsigned *count);
int ric_tsuqi(ric_ctx *ctx, int Quant, void *arena);
int ric_transform_inv(ric_ctx *ctx, const void *arena, int16_t *plane, int stride);

/* ric_set_base_weight: the baseWeight argument of CWavelet2D::SetWeight(t, baseWeight) (wavelet2d.h:36,
 * wavelet2d.cpp:1009-1032): the band weights of the context are recomputed in the reference's expression order
 * (top level D = base / scale, V = H = base, L = base * scale, then down the chain).  1.0 at ric_create. */
int ric_set_base_weight(ric_ctx *ctx, float baseWeight);
/* ric_quant_host / ric_tsuq_host: like ric_quant / ric_tsuq, but on the bands in `arena` (HOST, in/out), which are
 * uploaded first -- the reference's CodeBand / TSUQ work on pBand wherever the caller left it, so a caller that
 * edits the bands between Transform() and CodeBand() needs this form.  The C++ shim always uses it. */
int ric_quant_host(ric_ctx *ctx, int Quant, int lambda, void *arena);
int ric_tsuq_host(ric_ctx *ctx, int Quant, float thres, void *arena, unsigned *count);

/* ---- single bands, no context: the public per-band methods of the reference classes --------------------------
 * A band is described the way CBand holds it (src/lib/band.h:43-59): `data` = pBand (HOST, in/out, rows of
 * `stride` samples, short or int), dimx/dimy/stride = DimX/DimY/DimXAlign, weight = Weight.  Each call uploads
 * the band(s), runs the kernel on `device` and downloads the result; these exist for API fidelity (the class
 * shim's CBand::TSUQ / TSUQi / CBandCodec::buildTree call them), not for throughput.
 * ric_buf_tsuq:   CBand::TSUQ<C>(Quant, Thres)   band.h:65-92;  count, min and max receive Count, Min and Max (each may be NULL).
 * ric_buf_tsuqi:  CBand::TSUQi<C>(Quant)         band.h:94-107.
 * ric_buf_build_tree: CBandCodec::buildTree<high_band, C>(Quant, lambda)  bandcodec.h:42, bandcodec.cpp:239-322
 *   on chain[0] and then, as the reference recurses through pParent, on chain[1..n-1] (each the next coarser
 *   band of the same orientation).  `flags` of every band (one byte per 4x4 block, ceil(dimx/4) per row) receives
 *   "pRD != 0" -- the only thing the reference ever asks of pRD.  high_band == 0: chain[0] adds its child's blocks
 *   like a parent band does; child_flags / child_dimx then describe that child (a previous call's output). */
typedef struct ric_band_buf {
	void *data;
	int dimx, dimy, stride, is_int;
	float weight;
	unsigned char *flags; /* ric_buf_build_tree only: out, ceil(dimx/4) * ceil(dimy/4) bytes; may be NULL */
} ric_band_buf;
int ric_buf_tsuq(int device, const ric_band_buf *band, int Quant, float thres, unsigned *count, int *min, int *max);
int ric_buf_tsuqi(int device, const ric_band_buf *band, int Quant);
int ric_buf_build_tree(int device, const ric_band_buf *chain, int n, int high_band, const unsigned char *child_flags,
                       int child_dimx, int Quant, int lambda);

/* ---- .ric container header (src/ric/ric.cpp:114-121,135,150-154,187-200) --------------------------
 * 9 bytes: "RUD2", u16 LE width, u16 LE height, one byte Quant:5 | Color:1 << 5 | Trans:2 << 6.
 * The payload that follows is the entropy coder's buffer from offset 2 (ric.cpp:176). */
#define RIC_HEADER_BYTES 9
int ric_header_write(uint8_t *out, int width, int height, int q, int color, int trans);
int ric_header_parse(const uint8_t *in, int *width, int *height, int *q, int *color, int *trans);

/* ---- host entropy stage of the .ric format ----------------------------------------
 * The second half of CWavelet2D::CodeBand (src/lib/wavelet2d.cpp:119-159: CBandCodec::pred + tree<encode>
 * over one shared CMuxCodec, src/lib/muxcodec.cpp) and CWavelet2D::DecodeBand (wavelet2d.cpp:183-221).
 * HOST code by nature -- one adaptive, serial bit stream per image -- and not a fallback for anything: the
 * transform/quantiser stage above has no CPU path.  No GPU, no context: the geometry arguments are those of
 * ric_create, so band offsets equal ric_get_band's.  Re-entrant; run one image per host thread.
 *
 * ric_entropy_encode: image_arena = the `channels` plane arenas of ONE image as the encode stage wrote them
 *   (folded values + markers).  Planes are coded luma first (Y, Cg, Co; ric.cpp:163-168).  The arenas are
 *   consumed (markers are cleared in place, as the reference does).  out receives the payload exactly as a
 *   .ric file holds it after the 9-byte header; *size its length.  RIC_E_NOMEM if cap is too small.
 * ric_entropy_decode: the inverse; image_arena receives signed quantised coefficients (decode-stage input).
 *   Differs from the reference decoder in one place: a 1-sample edge block of the finest level is read the
 *   way the encoder wrote it (the reference reader takes one bit too many there). */
int ric_entropy_encode(int width, int height, int channels, int levels, int level_chg, int align,
                       void *image_arena, uint8_t *out, size_t cap, size_t *size);
int ric_entropy_decode(int width, int height, int channels, int levels, int level_chg, int align,
                       const uint8_t *payload, size_t size, void *image_arena);
/* ric_entropy_encode_hinted: the same payload through the form the device stage uses -- a data-parallel pre-pass
 * computes per 4x4 block what depends on the band data only (skipped / insignificant / significant, parent
 * context, non-zero mask, combination index), the serial coder consumes those hints and never writes the bands.
 * Valid for arenas produced by the encode stage (it relies on the quantiser's parent/child invariant); the
 * arenas are left untouched.  On the host it exists to check that walker without a GPU. */
int ric_entropy_encode_hinted(int width, int height, int channels, int levels, int level_chg, int align,
                              const void *image_arena, uint8_t *out, size_t cap, size_t *size);

/* Plane-at-a-time form of the same stage, the granularity of the reference API: ONE coder object shared by
 * the planes of an image (CMuxCodec, src/lib/muxcodec.h:60-139) and one call per plane (the entropy half of
 * CodeBand / DecodeBand).  `stream` is laid out like the reference's buffer: bytes 0-1 belong to the coder's
 * start word (a .ric file drops them, ric.cpp:176,203-205) and the payload starts at stream + 2.
 *   ric_mux_encoder  = CMuxCodec(pStream, firstWord)        ric_mux_decoder = CMuxCodec(pStream)
 *   ric_mux_code_plane / ric_mux_decode_plane take ONE plane arena and the plane geometry of ric_create
 *   ric_mux_finish   = CMuxCodec::endCoding(): *end = offset of the returned pointer from `stream`
 *   ric_mux_destroy frees either kind. */
typedef struct ric_mux ric_mux;
int ric_mux_encoder(ric_mux **mux, uint8_t *stream, size_t cap, unsigned first_word);
int ric_mux_decoder(ric_mux **mux, const uint8_t *stream, size_t size);
int ric_mux_code_plane(ric_mux *mux, int width, int height, int levels, int level_chg, int align, void *plane_arena);
int ric_mux_decode_plane(ric_mux *mux, int width, int height, int levels, int level_chg, int align, void *plane_arena);
int ric_mux_finish(ric_mux *mux, size_t *end);
int ric_mux_destroy(ric_mux *mux);

/* The same stage ON THE DEVICE, for large batches: one image per warp, the identical coder source.  The
 * stream of an image is serial, so this only pays when thousands of images are resident (BASELINE configs[3]);
 * what it buys is that the band arenas never cross PCIe -- only the finished payloads do.
 * d_arenas: n image arenas in device memory, as ric_encode_u8_device wrote them / as ric_decode_u8_device
 * reads them (written).  The encoder runs a data-parallel pre-pass over all 4x4 blocks (skipped / insignificant /
 * significant, parent context, non-zero mask, combination index) and a serial coder per image that consumes
 * those hints; it relies on the encode stage's parent/child invariant, so feed it encode-stage arenas only
 * (environment variable RIC_ENTROPY_PLAIN=1 selects the plain walker, which accepts any arenas and consumes them).  d_out / d_payloads: n slots of `stride` bytes; d_sizes: n
 * payload lengths (encode writes -1 where `stride` was too small).  Asynchronous on `stream`. */
int ric_entropy_encode_device(ric_ctx *ctx, void *d_arenas, int n, uint8_t *d_out, size_t stride, long long *d_sizes,
                              void *stream);
/* d_status (device, n ints, may be NULL): 0 where image i decoded cleanly, 1 where its payload was truncated or
 * corrupt (the reader ran past the end, or met a code no encoder writes) or d_sizes[i] is negative / larger than
 * `stride` -- such an image's arenas hold garbage (or stay cleared) and must not be used.  Without d_status a
 * caller has no way to learn of a bad payload from this call; ric_decompress_u8_gpu checks internally. */
int ric_entropy_decode_device(ric_ctx *ctx, const uint8_t *d_payloads, size_t stride, const long long *d_sizes, int n,
                              void *d_arenas, int *d_status, void *stream);

/* ---- whole .ric files, batch (CompressImage / DecompressImage without the image-file I/O, ric.cpp:123-251) --
 * ric_compress_u8: n planar u8 images -> n complete .ric files (header + payload), file i at files + i*stride,
 *   its length in sizes[i].  The GPU stage runs chunk by chunk on the context's streams while `threads` host
 *   threads (<= 0: all hardware threads) entropy-code the chunks that have already landed in pinned memory.
 *   RIC_E_NOMEM if a file would not fit in `stride` bytes (width*height*channels + RIC_HEADER_BYTES always does
 *   for the reference, which allocates exactly that, ric.cpp:131-132).
 * ric_decompress_u8: n .ric files of the context's geometry and one common quantiser index -> planar u8.
 *   RIC_E_ARG if a header disagrees with the context (size, colour, transform) or the q values differ. */
int ric_compress_u8(ric_ctx *ctx, const uint8_t *src, int n, int q, uint8_t *files, size_t stride, size_t *sizes,
                    int threads);
int ric_decompress_u8(ric_ctx *ctx, const uint8_t *files, size_t stride, const size_t *sizes, int n, uint8_t *dst,
                      int threads);

/* The same two calls with the entropy stage ON THE DEVICE (ric_entropy_*_device): pixels go up, only finished
 * files come back (and the reverse), the band arenas stay in HBM.  Meant for batches of hundreds to thousands
 * of images per call -- the per-image stream is serial, so the device needs many images in flight to beat the
 * host threads (DESIGN.md section 7 has the measured crossover).  Pinned host buffers recommended. */
int ric_compress_u8_gpu(ric_ctx *ctx, const uint8_t *src, int n, int q, uint8_t *files, size_t stride, size_t *sizes);
int ric_decompress_u8_gpu(ric_ctx *ctx, const uint8_t *files, size_t stride, const size_t *sizes, int n, uint8_t *dst);

/* pinned host memory helpers (arenas handed to host entropy threads should be pinned) */
int ric_host_alloc(void **p, size_t bytes);
int ric_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
