// ric_landed.h -- hand-off between the CUDA callback thread and the host entropy workers of ric_compress_u8.
// (Host side of the streaming pipeline; no counterpart in the reference, which codes one image at a time.)
#pragma once
#include <condition_variable>
#include <mutex>
#include <utility>
#include <vector>

namespace ric {

struct LandedQueue {  // images whose arenas have reached pinned memory
	// The chunks alternate over three streams, so they may land OUT OF ORDER: each landed range is queued as it is,
	// never summarised as "everything below first + count".
	std::mutex mu;
	std::condition_variable cv;
	std::vector<std::pair<int, int>> ranges;  // [begin, end) of landed, not yet handed out images
	int handed = 0, total = 0;
	bool released = false;  // error path: hand out nothing more
	static void landed(void *user, int first, int count)  // CUDA callback thread: no CUDA calls here
	{
		LandedQueue *q = (LandedQueue *)user;
		{ std::lock_guard<std::mutex> l(q->mu); q->ranges.emplace_back(first, first + count); }
		q->cv.notify_all();
	}
	void release_all() { { std::lock_guard<std::mutex> l(mu); released = true; } cv.notify_all(); }
	int take()  // an image whose arenas have landed, or -1 when all have been handed out (or the call was cancelled)
	{
		std::unique_lock<std::mutex> l(mu);
		cv.wait(l, [&] { return !ranges.empty() || handed >= total || released; });
		if (ranges.empty()) return -1;
		const int i = ranges.front().first++;
		if (ranges.front().first == ranges.front().second) ranges.erase(ranges.begin());
		if (++handed >= total) cv.notify_all();  // workers still waiting have nothing left to wait for
		return i;
	}
};

}  // namespace ric
