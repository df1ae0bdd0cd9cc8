// CPU test of the callback -> worker hand-off of ric_compress_u8 (csrc/ric_landed.h): chunks of one batch alternate
// over three streams and may land out of order; a worker must never be handed an image whose chunk has not landed.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>

#include "../../rududu_image_codec_b200/csrc/ric_landed.h"

int main()
{
	int fails = 0;
	for (int round = 0; round < 50; round++) {
		const int step = 4, n = 30, nchunks = (n + step - 1) / step;
		ric::LandedQueue q;
		q.total = n;
		std::vector<std::atomic<int>> is_landed(n), taken(n);
		for (int i = 0; i < n; i++) { is_landed[i] = 0; taken[i] = 0; }
		std::atomic<int> early{0};
		std::vector<std::thread> pool;
		for (int t = 0; t < 5; t++)
			pool.emplace_back([&] {
				for (int i; (i = q.take()) >= 0;) {
					if (!is_landed[i]) early++;
					taken[i]++;
				}
			});
		// chunk k lands at "time" (k % 3 == 0 ? late : early): the order 1, 2, 0, 4, 5, 3, ...
		std::vector<int> order;
		for (int k0 = 0; k0 < nchunks; k0 += 3) {
			for (int d : {1, 2, 0}) if (k0 + d < nchunks) order.push_back(k0 + d);
		}
		for (int k : order) {
			const int first = k * step, count = std::min(step, n - first);
			for (int i = first; i < first + count; i++) is_landed[i] = 1;
			ric::LandedQueue::landed(&q, first, count);
			if (round % 2) std::this_thread::sleep_for(std::chrono::microseconds(200));
		}
		for (auto &t : pool) t.join();
		for (int i = 0; i < n; i++) if (taken[i] != 1) { printf("image %d handed out %d times\n", i, (int)taken[i]); fails++; }
		if (early) { printf("round %d: %d images handed out before they landed\n", round, (int)early); fails++; }
	}
	{  // cancelled call: waiting workers leave
		ric::LandedQueue q;
		q.total = 8;
		std::atomic<int> got{0};
		std::thread w([&] { while (q.take() >= 0) got++; });
		ric::LandedQueue::landed(&q, 4, 2);
		std::this_thread::sleep_for(std::chrono::milliseconds(20));
		q.release_all();
		w.join();
		if (got != 2) { printf("cancel: %d images handed out, want 2\n", (int)got); fails++; }
	}
	printf(fails ? "landed_test: FAIL\n" : "landed_test: ok\n");
	return fails != 0;
}
