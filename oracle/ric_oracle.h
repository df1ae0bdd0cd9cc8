/* oracle/ric_oracle.h -- CPU restatement of the RIC transform+quant hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker the CUDA path is diffed against; the product
 * (rududu_image_codec_b200/csrc) never includes, links or calls anything in oracle/.
 * Parity of this restatement is PINNED: tests/test_oracle_vs_ref.py diffs every function below
 * against the compiled, unmodified reference (oracle/_ref/libric_ref.so) band by band, and
 * tests/golden/ holds CRCs generated from that reference (SURVEY.md Appendix C KATs included).
 */
#ifndef RIC_ORACLE_H
#define RIC_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RICO_MAX_LEVELS 16
#define RICO_MAX_BANDS (3 * RICO_MAX_LEVELS + 1)

enum { RICO_CDF97 = 0, RICO_CDF53 = 1, RICO_HAAR = 2 }; /* enum trans, src/lib/utils.h:28 */

typedef struct {
	int dimx, dimy;  /* CBand::DimX/DimY            src/lib/band.h:43-44 */
	int stride;      /* CBand::DimXAlign (samples)  src/lib/band.cpp:57 */
	int is_int;      /* band_t sint                 src/lib/band.h:35 */
	float weight;    /* CBand::Weight set by SetWeight, src/lib/wavelet2d.cpp:1009-1032 */
	size_t offset;   /* byte offset of the band inside the arena (32-byte aligned) */
} rico_band;

/* Canonical band order: id = 3*lev + {0:D,1:H,2:V}, lev 0 = finest; id 3*nlev = coarsest LL. */
typedef struct {
	int width, height, levels, level_chg, align, trans;
	int nlev;                          /* levels actually built (wavelet2d.cpp:76) */
	int lev_w[RICO_MAX_LEVELS];        /* input size of each level */
	int lev_h[RICO_MAX_LEVELS];
	int lev_is_int[RICO_MAX_LEVELS];
	int nbands;
	rico_band band[RICO_MAX_BANDS];
	size_t arena_bytes;
} rico_geom;

int rico_quants(int idx);                                                      /* ric.cpp:42-49 */
int rico_geom_init(rico_geom *g, int w, int h, int levels, int level_chg, int align, int trans);

/* plane: int16 W x H (not modified).  Writes every band of the arena (padding columns zeroed). */
void rico_forward(const rico_geom *g, const int16_t *plane, int stride, void *arena);
/* arena -> int16 plane (arena not modified).  q1_quirk!=0 reproduces the reference's 5/3 stride bug. */
void rico_inverse(const rico_geom *g, const void *arena, int16_t *plane, int stride, int q1_quirk);
/* encode quantiser: buildTree on D/H/V chains + TSUQ(0.5) on LL, in place (wavelet2d.cpp:110-126) */
void rico_quant(const rico_geom *g, void *arena, int Quant, int lambda);
/* CWavelet2D::TSUQ on all bands (wavelet2d.cpp:224-246); returns Count */
unsigned rico_tsuq_all(const rico_geom *g, void *arena, int Quant, float thres);
/* CWavelet2D::TSUQi (wavelet2d.cpp:248-268) */
void rico_tsuqi(const rico_geom *g, void *arena, int Quant);
/* folded (q<<1|sign, -0x8000 markers) -> signed coefficients, i.e. what DecodeBand leaves in HF bands */
void rico_unfold(const rico_geom *g, void *arena);

/* colour / level shift, ric.cpp:76-112,143-148,227-246.  src/dst planar u8; planes planar int16 */
void rico_colour_fwd(const uint8_t *src, int w, int h, int ch, int q, int16_t *planes);
void rico_colour_inv(const int16_t *planes, int w, int h, int ch, int q, uint8_t *dst);

/* whole stage, as ric drives it: src planar u8 -> ch arenas (plane order 0..ch-1 = Co,Cg,Y or gray) */
void rico_encode_image(const rico_geom *g, const uint8_t *src, int ch, int q, void *arenas);
void rico_decode_image(const rico_geom *g, const void *arenas, int ch, int q, uint8_t *dst);
/* luma/chroma quantiser pair for ric's -q value: plane index p of ch */
void rico_plane_quant(int q, int ch, int p, int *Quant, int *lambda);

#ifdef __cplusplus
}
#endif
#endif
