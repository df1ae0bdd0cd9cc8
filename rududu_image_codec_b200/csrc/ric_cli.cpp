// ric_cli.cpp -- command-line front end over the C ABI: PNM <-> .ric on the GPU.
//
// Mirrors the reference tool's interface (src/ric/ric.cpp:300-360): -i <input> [-o <output>] [-q 0..31]
// [-t 0|1|2] [-d]; an input whose name contains ".ric" is decompressed, anything else is compressed;
// default output names as ric.cpp:331-349.  Image I/O is binary PNM only (P5 gray / P6 colour, maxval
// <= 255) -- the reference reads whatever CImg can; that dependency is not reproduced.
// Extra: -g <device> selects the GPU.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/ric_b200.h"

namespace {

const int kLevels = 5, kLevelChg = 1;  // WAV_LEVELS, WAV_LEVELS - 4 (ric.cpp:36,159)
const int kShift = 4;                  // SHIFT (ric.cpp:37)

[[noreturn]] void die(const std::string &msg)
{
	fprintf(stderr, "ric_b200: %s\n", msg.c_str());
	exit(1);
}

void check(int rc, const char *what)
{
	if (rc) die(std::string(what) + ": " + ric_last_error());
}

std::vector<uint8_t> read_file(const std::string &path)
{
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) die("cannot open " + path);
	std::vector<uint8_t> data;
	uint8_t buf[1 << 16];
	for (size_t n; (n = fread(buf, 1, sizeof buf, f)) > 0;) data.insert(data.end(), buf, buf + n);
	fclose(f);
	return data;
}

void write_file(const std::string &path, const uint8_t *a, size_t na, const uint8_t *b = 0, size_t nb = 0)
{
	FILE *f = fopen(path.c_str(), "wb");
	if (!f) die("cannot create " + path);
	if (fwrite(a, 1, na, f) != na || (nb && fwrite(b, 1, nb, f) != nb)) die("short write to " + path);
	fclose(f);
}

// Binary PNM -> planar u8.  Returns channels (1 or 3).
int read_pnm(const std::string &path, int &w, int &h, std::vector<uint8_t> &planar)
{
	const std::vector<uint8_t> d = read_file(path);
	size_t pos = 0;
	auto token = [&]() {
		for (;;) {
			while (pos < d.size() && (d[pos] == ' ' || d[pos] == '\t' || d[pos] == '\r' || d[pos] == '\n')) pos++;
			if (pos < d.size() && d[pos] == '#') { while (pos < d.size() && d[pos] != '\n') pos++; continue; }
			break;
		}
		std::string t;
		while (pos < d.size() && d[pos] > ' ') t.push_back((char)d[pos++]);
		return t;
	};
	const std::string magic = token();
	if (magic != "P5" && magic != "P6") die(path + ": not a binary PGM/PPM (P5/P6)");
	const int ch = magic == "P6" ? 3 : 1;
	w = atoi(token().c_str());
	h = atoi(token().c_str());
	const int maxval = atoi(token().c_str());
	if (w < 1 || h < 1 || maxval < 1 || maxval > 255) die(path + ": unsupported PNM header (8-bit samples only)");
	pos++;  // the single whitespace byte after maxval
	const size_t n = (size_t)w * h;
	if (d.size() < pos + n * ch) die(path + ": truncated pixel data");
	planar.resize(n * ch);
	const uint8_t *px = d.data() + pos;
	if (ch == 1) memcpy(planar.data(), px, n);
	else
		for (size_t i = 0; i < n; i++)
			for (int c = 0; c < 3; c++) planar[c * n + i] = px[3 * i + c];
	return ch;
}

void write_pnm(const std::string &path, int w, int h, int ch, const uint8_t *planar)
{
	char head[64];
	const int hn = snprintf(head, sizeof head, "P%d\n%d %d\n255\n", ch == 3 ? 6 : 5, w, h);
	const size_t n = (size_t)w * h;
	std::vector<uint8_t> px(n * ch);
	if (ch == 1) memcpy(px.data(), planar, n);
	else
		for (size_t i = 0; i < n; i++)
			for (int c = 0; c < 3; c++) px[3 * i + c] = planar[c * n + i];
	write_file(path, (const uint8_t *)head, (size_t)hn, px.data(), px.size());
}

inline int clip255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

// Error-diffusion rounding of the 12.4 fixed-point gray plane to 8 bits (ric.cpp:51-74): the first and
// last column and the last row are rounded plainly; elsewhere the rounding error is pushed right and to
// the three pixels below with weights 7/16, 3/16, 5/16, 1/16 built from shifts.  All stores are 16-bit.
void dither_plane(int16_t *p, int w, int h, uint8_t *out)
{
	const int half = 1 << (kShift - 1);
	auto plain = [&](int16_t v) { return (int16_t)clip255(128 + ((v + half) >> kShift)); };
	int16_t *row = p;
	for (int y = 0; y < h - 1; y++, row += w) {
		row[0] = plain(row[0]);
		for (int x = 1; x < w - 1; x++) {
			int16_t t = (int16_t)(row[x] + half);
			const int16_t q = (int16_t)(t >> kShift);
			t = (int16_t)(t - (q << kShift));
			row[x + 1] = (int16_t)(row[x + 1] + ((t >> 1) - (t >> 4)));
			row[x + w - 1] = (int16_t)(row[x + w - 1] + ((t >> 3) + (t >> 4)));
			row[x + w] = (int16_t)(row[x + w] + ((t >> 2) + (t >> 4)));
			row[x + w + 1] = (int16_t)(row[x + w + 1] + (t >> 4));
			row[x] = (int16_t)clip255(q + 128);
		}
		row[w - 1] = plain(row[w - 1]);
	}
	for (int x = 0; x < w; x++) row[x] = plain(row[x]);
	for (size_t i = 0; i < (size_t)w * h; i++) out[i] = (uint8_t)p[i];
}

void usage()
{
	fprintf(stderr,
	        "usage: ric_b200 -i <input> [-o <output>] [-q 0..31] [-t 0|1|2] [-d] [-g <gpu>]\n"
	        "  input *.ric  -> decompress to PNM (default <input>.pnm)\n"
	        "  other input  -> binary PGM/PPM, compress (default <input minus extension>.ric)\n"
	        "  -q  quantiser, 0 = lossless ... 31 (default 9)     -t  0 cdf97, 1 cdf53, 2 haar (default 1 when -q 0)\n"
	        "  -d  dithered output (decompression of gray images only)\n");
}

}  // namespace

int main(int argc, char **argv)
{
	std::string in, out;
	int q = 9, trans = -1, device = 0;
	bool dither = false;
	for (int i = 1; i < argc; i++) {
		const std::string a = argv[i];
		auto value = [&]() { if (i + 1 >= argc) { usage(); exit(1); } return std::string(argv[++i]); };
		if (a == "-i") in = value();
		else if (a == "-o") out = value();
		else if (a == "-q") q = atoi(value().c_str());
		else if (a == "-t") trans = atoi(value().c_str());
		else if (a == "-g") device = atoi(value().c_str());
		else if (a == "-d") dither = true;
		else { usage(); return a == "-h" || a == "-help" || a == "--help" ? 0 : 1; }
	}
	if (in.empty()) { usage(); return 1; }
	if (q < 0 || q > 31) die("-q must be in 0..31");
	if (trans < 0) trans = q == 0 ? 1 : 0;
	if (trans > 2) trans = 0;  // ric.cpp:316-317

	const bool decode = in.find(".ric") != std::string::npos;  // ric.cpp:331
	if (out.empty()) {
		out = in;
		if (decode) out += ".pnm";
		else {
			const size_t dot = in.find_last_of('.'), slash = in.find_last_of('/');
			if (dot != std::string::npos && (slash == std::string::npos || slash < dot)) out.resize(dot);
			out += ".ric";
		}
	}

	ric_ctx *ctx = 0;
	if (!decode) {
		int w, h;
		std::vector<uint8_t> px;
		const int ch = read_pnm(in, w, h, px);
		check(ric_create(&ctx, device, w, h, ch, kLevels, kLevelChg, 32, trans, 1), "ric_create");
		std::vector<uint8_t> file((size_t)w * h * ch * 2 + 4096);
		size_t size = 0;
		check(ric_compress_u8(ctx, px.data(), 1, q, file.data(), file.size(), &size, 1), "ric_compress_u8");
		write_file(out, file.data(), size);
		fprintf(stderr, "%s: %dx%d %s, q=%d -> %zu bytes (%.3f bpp)\n", out.c_str(), w, h, ch == 3 ? "colour" : "gray", q, size,
		        8.0 * size / ((double)w * h));
	} else {
		const std::vector<uint8_t> file = read_file(in);
		int w, h, fq, color, ft;
		if (file.size() < RIC_HEADER_BYTES) die(in + ": too short");
		check(ric_header_parse(file.data(), &w, &h, &fq, &color, &ft), "ric_header_parse");
		const int ch = color ? 3 : 1;
		check(ric_create(&ctx, device, w, h, ch, kLevels, kLevelChg, 32, ft, 1), "ric_create");
		std::vector<uint8_t> px((size_t)w * h * ch);
		if (dither && !color && fq != 0) {
			// The dithered path needs the 16-bit plane before the final rounding: entropy stage on the host,
			// TSUQi and the inverse transform on the GPU through the plane-level calls, diffusion on the host.
			ric_info inf;
			check(ric_get_info(ctx, &inf), "ric_get_info");
			void *arena = 0;
			check(ric_host_alloc(&arena, inf.arena_bytes), "ric_host_alloc");
			check(ric_entropy_decode(w, h, 1, kLevels, kLevelChg, 32, file.data() + RIC_HEADER_BYTES, file.size() - RIC_HEADER_BYTES, arena),
			      "ric_entropy_decode");
			int Q, lambda;
			check(ric_plane_quant(fq, 1, 0, &Q, &lambda), "ric_plane_quant");
			check(ric_tsuqi(ctx, Q, arena), "ric_tsuqi");
			std::vector<int16_t> plane((size_t)w * h);
			check(ric_transform_inv(ctx, arena, plane.data(), w), "ric_transform_inv");
			dither_plane(plane.data(), w, h, px.data());
			ric_host_free(arena);
		} else {
			const size_t size = file.size();
			check(ric_decompress_u8(ctx, file.data(), size, &size, 1, px.data(), 1), "ric_decompress_u8");
		}
		write_pnm(out, w, h, ch, px.data());
		fprintf(stderr, "%s: %dx%d %s, q=%d\n", out.c_str(), w, h, ch == 3 ? "colour" : "gray", fq);
	}
	ric_destroy(ctx);
	return 0;
}
