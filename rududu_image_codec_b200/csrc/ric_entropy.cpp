// ric_entropy.cpp -- host entropy stage of the .ric format (one image per host thread); the coder itself is
// ric_entropy_core.h.  See ric_entropy.h for the reference map.
#include "ric_entropy.h"

#include <vector>

#include "ric_entropy_core.h"

namespace ric {

using namespace ent;

static const Tables kTables;  // built when the library is loaded
const void *entropy_tables(size_t *bytes)
{
	*bytes = sizeof(Tables);
	return &kTables;
}

long entropy_encode_image(const HostGeom &g, char *image_arena, uint8_t *out, size_t cap)
{
	MuxWriter w(out, cap);
	WPort io(w, &kTables);
	for (int i = 0; i < g.channels; i++) {
		const int plane = g.channels == 3 ? 2 - i : 0;  // Y, Cg, Co
		walk_plane(io, g, image_arena + (size_t)plane * g.arena_bytes);
	}
	uint8_t *end = w.finish();
	return w.overflow() ? -1 : (long)(end - out);
}

// The hinted encoder on the host: the pre-pass runs as a plain loop, then the hinted walker.  It exists so
// that the device path's walker (same source) can be checked without a GPU; the arenas are left untouched.
long entropy_encode_image_hinted(const HostGeom &g, const char *image_arena, uint8_t *out, size_t cap)
{
	std::vector<BlockHint> hints(g.flag_bytes);
	MuxWriter w(out, cap);
	WPort io(w, &kTables);
	for (int i = 0; i < g.channels; i++) {
		const int plane = g.channels == 3 ? 2 - i : 0;
		const char *pl = image_arena + (size_t)plane * g.arena_bytes;
		for (int id = 0; id < 3 * g.nlev; id++)
			for (int by = 0; by < (g.band[id].dimy + 3) / 4; by++)
				for (int bx = 0; bx < g.flag_bw[id]; bx++) hints[hint_slot(g, id, bx, by)] = make_hint(g, kTables, pl, id, bx, by);
		walk_plane_hinted(io, g, const_cast<char *>(pl), hints.data());
	}
	uint8_t *end = w.finish();
	return w.overflow() ? -1 : (long)(end - out);
}

// ---- plane-at-a-time objects (the reference's CMuxCodec + CodeBand / DecodeBand granularity) ----
struct MuxEncoder {
	uint8_t *stream;
	MuxWriter w;
	MuxEncoder(uint8_t *s, size_t cap, unsigned first_word) : stream(s), w(s + 2, cap - 2, first_word) {}
};
struct MuxDecoder {
	MuxReader r;
	MuxDecoder(const uint8_t *s, size_t size) : r(s + 2, size - 2) {}
};

MuxEncoder *mux_encoder_new(uint8_t *stream, size_t cap, unsigned first_word)
{
	return cap < 8 ? nullptr : new MuxEncoder(stream, cap, first_word);
}
void mux_encoder_plane(MuxEncoder *m, const HostGeom &g, char *plane_arena)
{
	WPort io(m->w, &kTables);
	walk_plane(io, g, plane_arena);
}
long mux_encoder_finish(MuxEncoder *m)
{
	uint8_t *end = m->w.finish();
	m->stream[0] = m->w.lead(0);
	m->stream[1] = m->w.lead(1);
	return m->w.overflow() ? -1 : (long)(end - m->stream);
}
void mux_encoder_free(MuxEncoder *m) { delete m; }

MuxDecoder *mux_decoder_new(const uint8_t *stream, size_t size) { return size < 4 ? nullptr : new MuxDecoder(stream, size); }
int mux_decoder_plane(MuxDecoder *m, const HostGeom &g, char *plane_arena)
{
	RPort io(m->r, &kTables);
	walk_plane(io, g, plane_arena);
	return m->r.overrun() ? -1 : 0;
}
void mux_decoder_free(MuxDecoder *m) { delete m; }

int entropy_decode_image(const HostGeom &g, const uint8_t *payload, size_t size, char *image_arena)
{
	MuxReader r(payload, size);
	RPort io(r, &kTables);
	for (int i = 0; i < g.channels; i++) {
		const int plane = g.channels == 3 ? 2 - i : 0;
		walk_plane(io, g, image_arena + (size_t)plane * g.arena_bytes);
	}
	return r.overrun() ? -1 : 0;
}

}  // namespace ric
