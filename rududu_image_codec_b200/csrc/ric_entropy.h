// ric_entropy.h -- host entropy stage of the .ric format: the byte stream the
// reference's CodeBand/DecodeBand entropy half produces from / turns back into quantised band arenas.
//
// Host code by nature: one adaptive, strictly serial bit stream per image (all planes share one coder,
// src/ric/ric.cpp:157-176), so the parallelism is one image per host thread, fed by the GPU stage.
// Format restated from (reference file:line):
//   stream multiplexer + binary range coder   src/lib/muxcodec.h:140-212, muxcodec.cpp:33-118,517-579
//   taboo / enumerative / truncated codes     muxcodec.cpp:198-276,278-401,496-515
//   adaptive geometric model                  src/lib/geomcodec.h:40-97, geomcodec.cpp:31-56
//   adaptive binary model                     src/lib/bitcodec.h:28-93, bitcodec.cpp:31-42
//   LL band DPCM (pred)                       src/lib/bandcodec.cpp:62-110
//   block zero-tree (tree / block_enum)       src/lib/bandcodec.cpp:324-604
//   band order inside a plane                 src/lib/wavelet2d.cpp:119-159 (encode), :183-221 (decode)
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "ric_host.h"

namespace ric {

// Encode every plane of ONE image (luma first: Y, Cg, Co -- ric.cpp:163-168) from its quantised band
// arenas (what the encode stage produced) into `out`: the payload exactly as a .ric file holds it after
// the 9-byte header.  The arenas are consumed: zero-tree markers are cleared/propagated in place, as the
// reference's tree<encode> does.  Returns the payload size, or -1 if `cap` bytes were not enough.
long entropy_encode_image(const HostGeom &g, char *image_arena, uint8_t *out, size_t cap);

// Same payload through the parallel pre-pass + hinted walker (ric_entropy_core.h), the form the device stage uses.
// Requires arenas that come from the encode stage; leaves them untouched.
long entropy_encode_image_hinted(const HostGeom &g, const char *image_arena, uint8_t *out, size_t cap);

// Decode one image's payload into signed quantised band arenas (what the decode stage consumes).
// `payload` need not be padded.  Returns 0, or -1 on a truncated / over-long stream.
int entropy_decode_image(const HostGeom &g, const uint8_t *payload, size_t size, char *image_arena);

// Plane-at-a-time form, the granularity of the reference API: one coder object shared by the planes of an
// image (CMuxCodec, muxcodec.cpp:25-64), one call per plane (the entropy half of CodeBand / DecodeBand).
// `stream` is laid out as the reference's buffer: bytes 0-1 carry the coder's start word (a .ric file
// drops them, ric.cpp:176,203-205), the payload starts at stream + 2.
struct MuxEncoder;
struct MuxDecoder;
MuxEncoder *mux_encoder_new(uint8_t *stream, size_t cap, unsigned first_word);
void mux_encoder_plane(MuxEncoder *m, const HostGeom &g, char *plane_arena);
long mux_encoder_finish(MuxEncoder *m);  // offset of the end pointer from `stream` (CMuxCodec::endCoding), -1 on overflow
void mux_encoder_free(MuxEncoder *m);
MuxDecoder *mux_decoder_new(const uint8_t *stream, size_t size);
int mux_decoder_plane(MuxDecoder *m, const HostGeom &g, char *plane_arena);
void mux_decoder_free(MuxDecoder *m);

// The coder's static tables (a POD block) for the device-side copy; used by the GPU entropy stage only.
const void *entropy_tables(size_t *bytes);

}  // namespace ric
