// End-to-end pipeline of BASELINE configs[3] in miniature: a batch of RGB images is encoded on the
// GPU (ric_encode_u8_stream) while host worker threads entropy-code the chunks that have already
// landed in pinned memory with the REFERENCE's own coder (oracle/_ref: CBandCodec::pred/tree +
// CMuxCodec, unmodified), and every resulting .ric payload is compared with the reference's
// monolithic CompressImage path.  TEST ONLY: the product never links oracle/.
// usage: pipeline_test W H N q images.u8   (N planar RGB u8 images)
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <queue>
#include <thread>
#include <vector>

#include "ric_b200.h"

extern "C" {  // oracle/ref_harness.cpp
void *ref_wavelet_new(int w, int h, int levels, int level_chg, int t);
void ref_wavelet_free(void *h);
void ref_band_set(void *h, int id, const void *src);
void *ref_codec_new_enc(unsigned char *buf);
long ref_codec_end(void *c, unsigned char *buf);
void ref_codec_free(void *c);
void ref_entropy_encode(void *h, void *codec);
long ref_compress(const uint8_t *src, int w, int h, int ch, int q, int t, int levels, int level_chg, uint8_t *out, long cap);
}

struct Chunk { int first, count; };
static std::mutex mu;
static std::condition_variable cv;
static std::queue<Chunk> todo;
static std::atomic<int> landed{0};

static void on_chunk(void *, int first, int count)  // CUDA callback thread: no CUDA calls here
{
	{ std::lock_guard<std::mutex> l(mu); todo.push({first, count}); }
	landed += count;
	cv.notify_one();
}

int main(int argc, char **argv)
{
	if (argc != 6) return 2;
	const int w = atoi(argv[1]), h = atoi(argv[2]), n = atoi(argv[3]), q = atoi(argv[4]);
	const size_t img = (size_t)3 * w * h;
	std::vector<uint8_t> src(img * n);
	FILE *f = fopen(argv[5], "rb");
	if (!f || fread(src.data(), 1, src.size(), f) != src.size()) return 2;
	fclose(f);

	ric_ctx *ctx;
	if (ric_create(&ctx, 0, w, h, 3, 5, 1, 32, RIC_CDF97, n)) { fprintf(stderr, "%s\n", ric_last_error()); return 1; }
	ric_info inf;
	ric_get_info(ctx, &inf);
	void *arenas, *pinned_src;
	ric_host_alloc(&arenas, inf.image_arena_bytes * n);
	ric_host_alloc(&pinned_src, src.size());
	memcpy(pinned_src, src.data(), src.size());

	std::vector<std::vector<uint8_t>> payload(n);
	std::atomic<int> coded{0};
	auto worker = [&]() {
		void *wav = ref_wavelet_new(w, h, 5, 1, 0);
		std::vector<unsigned char> buf(img * 2 + 4096);
		for (;;) {
			Chunk c;
			{
				std::unique_lock<std::mutex> l(mu);
				cv.wait(l, [&] { return !todo.empty() || coded.load() + 0 >= n; });
				if (todo.empty()) break;
				c = todo.front();
				todo.pop();
			}
			for (int i = c.first; i < c.first + c.count; i++) {
				void *codec = ref_codec_new_enc(buf.data());
				const int order[3] = {2, 1, 0};  // Y, Cg, Co  (ric.cpp:163-168)
				for (int p : order) {
					char *a = (char *)arenas + (size_t)i * inf.image_arena_bytes + (size_t)p * inf.arena_bytes;
					for (int id = 0; id < inf.nbands; id++) {
						ric_band_info b;
						ric_get_band(ctx, id, &b);
						ref_band_set(wav, id, a + b.offset);
					}
					ref_entropy_encode(wav, codec);
				}
				long sz = ref_codec_end(codec, buf.data());
				ref_codec_free(codec);
				payload[i].assign(buf.begin() + 2, buf.begin() + sz);
				coded++;
			}
			cv.notify_all();
		}
		ref_wavelet_free(wav);
	};
	auto t0 = std::chrono::steady_clock::now();
	std::vector<std::thread> pool;
	for (int k = 0; k < 4; k++) pool.emplace_back(worker);
	if (ric_encode_u8_stream(ctx, (const uint8_t *)pinned_src, n, q, arenas, on_chunk, nullptr)) { fprintf(stderr, "%s\n", ric_last_error()); return 1; }
	const int landed_at_return = landed.load();  // < n: the call returned before the GPU finished
	ric_sync(ctx);
	auto t1 = std::chrono::steady_clock::now();
	{ std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return coded.load() >= n; }); }
	cv.notify_all();
	for (auto &t : pool) t.join();
	auto t2 = std::chrono::steady_clock::now();

	int bad = 0;
	std::vector<uint8_t> want(img * 2 + 4096);
	for (int i = 0; i < n; i++) {
		long sz = ref_compress(src.data() + img * i, w, h, 3, q, 0, 5, 1, want.data(), (long)want.size());
		if (sz != (long)payload[i].size() || memcmp(want.data(), payload[i].data(), sz)) { bad++; fprintf(stderr, "image %d: payload differs (%ld vs %zu bytes)\n", i, sz, payload[i].size()); }
	}
	printf("%s n=%d landed_at_return=%d gpu_ms=%.2f total_ms=%.2f\n", bad ? "FAIL" : "ok", n, landed_at_return,
	       std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t0).count());
	ric_host_free(arenas);
	ric_host_free(pinned_src);
	ric_destroy(ctx);
	return bad ? 1 : 0;
}
