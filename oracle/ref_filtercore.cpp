// ref_filtercore.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the REFERENCE's own FilterCore.h and ProgressBar.h, in place from
// /root/reference (never copied into this repo), against interface shims for the
// headers of the absent diskerror/c_lib and Boost (oracle/ref_shim/).  What this
// buys: the index / edge / narrowing logic of apply_filter_range
// (FilterCore.h:20-79) that runs here is the reference's, not a restatement.
// What it does not buy: the dot product fms() and the taps are shims, so the
// arithmetic is still restated -- parity stays "unpinned" (DESIGN.md).
//
// Output goes to oracle/_ref/libref_filtercore.so (git-ignored, travels with
// gpurun).  Built only where /root/reference exists (oracle/Makefile).
#include <iomanip>   // ProgressBar.h:30 uses setprecision without including it
#include <thread>
#include <vector>
#include <functional>
#include <FilterCore.h>

using namespace Diskerror;

extern "C" {

// apply_filter_range (FilterCore.h:20-27) on caller-owned planar buffers.
__attribute__((visibility("default")))
void ref_apply_filter_range(const float* x, long long N, const double* taps, long long n_taps,
                            float* y, long long start, long long end)
{
	VectorMath<boost::float32_t> channel(x, (size_t) N);
	WindowedSinc<boost::float64_t> sinc(taps, (size_t) n_taps);
	VectorMath<boost::float32_t> out((size_t) N);
	apply_filter_range(channel, sinc, out, start, end, nullptr);
	for (long long i = start; i < end; ++i) y[i] = out[(size_t) i];
}

// The per-channel thread fan-out of process_file (ProcessFile.cp:60-83) around
// the reference's apply_filter_range: chunk = N / threads, last thread takes
// the remainder.  Whole channel [0, N).
__attribute__((visibility("default")))
void ref_filter_channel_threads(const float* x, long long N, const double* taps, long long n_taps,
                                float* y, unsigned num_threads)
{
	VectorMath<boost::float32_t> channel(x, (size_t) N);
	WindowedSinc<boost::float64_t> sinc(taps, (size_t) n_taps);
	VectorMath<boost::float32_t> out((size_t) N);
	std::vector<std::thread> threads;
	const long long chunk = N / num_threads;
	for (unsigned i = 0; i < num_threads; ++i) {
		long long s = i * chunk;
		long long e = (i == num_threads - 1) ? N : s + chunk;
		threads.emplace_back(apply_filter_range, std::cref(channel), std::cref(sinc), std::ref(out),
		                     s, e, nullptr);
	}
	for (auto& t : threads) t.join();
	for (long long i = 0; i < N; ++i) y[i] = out[(size_t) i];
}

}
