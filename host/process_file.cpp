#include "process_file.hpp"

#include <algorithm>
#include <atomic>
#include <barrier>
#include <exception>
#include <format>
#include <iostream>
#include <mutex>
#include <thread>

#include "../include/fir_gpu.h"
#include "audio_container.hpp"
#include "errors.hpp"

namespace lowcut {
namespace {

std::mutex g_io; // serialises stdout between the per-GPU workers

void say(const std::string& s)
{
	std::lock_guard<std::mutex> l(g_io);
	std::cout << s << std::endl;
}

void check(int rc, const char* what)
{
	if (rc != FIR_GPU_OK) throw GpuError(rc, std::format("{}: {}", what, fir_gpu_last_error()));
}

struct PinnedBytes {
	unsigned char* p = nullptr;
	explicit PinnedBytes(size_t n)
	{
		p = static_cast<unsigned char*>(fir_gpu_host_alloc(n ? n : 1));
		if (!p) throw GpuError(FIR_GPU_ERR_NOMEM, std::format("pinned host buffer of {} bytes: {}", n, fir_gpu_last_error()));
	}
	~PinnedBytes() { fir_gpu_host_free(p); }
	PinnedBytes(const PinnedBytes&) = delete;
	PinnedBytes& operator=(const PinnedBytes&) = delete;
};

struct KernelHandle {
	fir_gpu_kernel* k = nullptr;
	int64_t half_len = 0;
	KernelHandle(fir_gpu_ctx* ctx, double fc, double bw) { check(fir_gpu_build_kernel(ctx, fc, bw, &k, &half_len), "fir_gpu_build_kernel"); }
	~KernelHandle() { fir_gpu_kernel_free(k); }
	KernelHandle(const KernelHandle&) = delete;
	KernelHandle& operator=(const KernelHandle&) = delete;
};

fir_gpu_pcm make_fmt(const PcmLayout& l, int64_t frames, int64_t halo_l, int64_t halo_r)
{
	fir_gpu_pcm f{};
	f.frames = frames;
	f.channels = l.channels;
	f.bits = l.bits;
	f.big_endian = l.big_endian ? 1 : 0;
	f.halo_left = halo_l;
	f.halo_right = halo_r;
	return f;
}

std::string timing_line(fir_gpu_ctx* ctx)
{
	fir_gpu_timing t{};
	if (fir_gpu_last_timing(ctx, &t) != FIR_GPU_OK) return "";
	return std::format("  device time: h2d {:.2f} ms, decode {:.2f} ms, fir {:.2f} ms, encode {:.2f} ms, d2h {:.2f} ms",
	                   t.h2d_ms, t.decode_ms, t.fir_ms, t.encode_ms, t.d2h_ms);
}

// Contiguous blocks of ceil(frames/G) frames rounded up to 16 (so every block starts
// on a 128-byte FP64 boundary and the summation grouping of the DMMA kernel is the
// same as in the unsharded run).
struct Block {
	int64_t start, frames, halo_l, halo_r;
};

std::vector<Block> plan_blocks(int64_t total, size_t world, int64_t half_len)
{
	int64_t per = (total + (int64_t) world - 1) / (int64_t) world;
	per = (per + 15) / 16 * 16;
	std::vector<Block> b;
	for (size_t r = 0; r < world; ++r) {
		const int64_t s = std::min<int64_t>((int64_t) r * per, total), e = std::min<int64_t>(s + per, total);
		b.push_back({s, e - s, std::min(half_len, s), std::min(half_len, total - e)});
	}
	return b;
}

void run_file(const std::filesystem::path& input_path, const std::filesystem::path& output_path,
              const FilterOptions& opts, const std::vector<fir_gpu_ctx*>& ctxs)
{
	auto status = [&](const std::string& s) {
		if (opts.verbose) say(s);
	};

	status("Opening input file.");                                   // ProcessFile.cp:33-35
	AudioContainer in(input_path);
	const PcmLayout& l = in.pcm();
	say("Processing file: " + input_path.filename().string());       // ProcessFile.cp:37 (unconditional)
	status(std::format("  {} {} ch, {} bit {}, {} Hz, {} frames, {} chunks", in.type_name(), l.channels, l.bits,
	                   l.big_endian ? "big-endian" : "little-endian", l.sample_rate, l.frames, in.chunks().size()));
	if (!(l.sample_rate > 0.0)) throw FormatError(input_path.string() + ": sample rate is zero");

	status("Reading samples.");                                      // ProcessFile.cp:39-41
	PinnedBytes pcm(l.payload_bytes);
	in.read_payload(0, l.payload_bytes, pcm.p);

	const uint64_t fb = (uint64_t) l.channels * (l.bits / 8);
	const double fc = opts.freq / l.sample_rate, bw = opts.slope / l.sample_rate; // ProcessFile.cp:48-49

	// sample-block mode only pays when every GPU gets a sizeable block
	size_t world = ctxs.size();
	if (world > 1 && (int64_t) l.frames < (int64_t) world * (1 << 18)) world = 1;

	status("Creating sinc kernel for this file's sample rate.");
	double scale = 1.0, peak = 0.0;
	if (l.frames > 0 && world == 1) {
		fir_gpu_ctx* ctx = ctxs[0];
		KernelHandle k(ctx, fc, bw);
		status(std::format("  {} taps", fir_gpu_kernel_num_taps(k.k)));
		status("Filtering.");
		const fir_gpu_pcm fmt = make_fmt(l, (int64_t) l.frames, 0, 0);
		check(fir_gpu_apply(ctx, k.k, pcm.p, &fmt), "fir_gpu_apply");
		check(fir_gpu_peak(ctx, &peak), "fir_gpu_peak");             // ProcessFile.cp:92-96
		scale = scale_for_peak(peak, opts.normalize);                // ProcessFile.cp:98-101
		if (scale != 1.0) status("Doing audio normalize.");
		check(fir_gpu_encode(ctx, scale, pcm.p), "fir_gpu_encode");  // the H2D copy is complete: reuse the buffer
		status(timing_line(ctx));
	} else if (l.frames > 0) {
		// contiguous sample blocks with (taps-1) halo, one host thread per GPU
		std::vector<std::unique_ptr<KernelHandle>> ks(world);
		for (size_t r = 0; r < world; ++r) ks[r] = std::make_unique<KernelHandle>(ctxs[r], fc, bw);
		const std::vector<Block> blocks = plan_blocks((int64_t) l.frames, world, ks[0]->half_len);
		status(std::format("  {} taps, {} sample blocks of up to {} frames, halo {} frames each side",
		                   fir_gpu_kernel_num_taps(ks[0]->k), world, blocks[0].frames, ks[0]->half_len));
		status("Filtering.");
		std::vector<double> peaks(world, 0.0);
		std::vector<std::exception_ptr> errs(world);
		std::atomic<bool> failed{false};
		std::barrier sync((std::ptrdiff_t) world);
		std::vector<std::thread> th;
		for (size_t r = 0; r < world; ++r)
			th.emplace_back([&, r] {
				const Block& b = blocks[r];
				try {
					if (b.frames) {
						const fir_gpu_pcm fmt = make_fmt(l, b.frames, b.halo_l, b.halo_r);
						check(fir_gpu_apply(ctxs[r], ks[r]->k, pcm.p + (uint64_t) (b.start - b.halo_l) * fb, &fmt),
						      "fir_gpu_apply");
						check(fir_gpu_peak(ctxs[r], &peaks[r]), "fir_gpu_peak");
					}
				} catch (...) {
					errs[r] = std::current_exception();
					failed = true;
				}
				// every upload is complete and every peak known beyond this point:
				// the blocks may now be overwritten in place with the common scale
				sync.arrive_and_wait();
				if (failed || !b.frames) return;
				try {
					const double pk = *std::max_element(peaks.begin(), peaks.end());
					check(fir_gpu_encode(ctxs[r], scale_for_peak(pk, opts.normalize), pcm.p + (uint64_t) b.start * fb),
					      "fir_gpu_encode");
				} catch (...) {
					errs[r] = std::current_exception();
				}
			});
		for (auto& t : th) t.join();
		for (auto& e : errs)
			if (e) std::rethrow_exception(e);
		peak = *std::max_element(peaks.begin(), peaks.end());
		scale = scale_for_peak(peak, opts.normalize);
		if (scale != 1.0) status("Doing audio normalize.");
		for (size_t r = 0; r < world; ++r) status(std::format("  GPU {}:{}", r, timing_line(ctxs[r])));
	}
	status(std::format("  peak {:.9f}, scale {:.9f}", peak, scale));

	status("Writing output file.");                                  // ProcessFile.cp:104-117
	const int fd = in.create_output(output_path);                    // every non-sample byte, verbatim
	try {
		AudioContainer::write_payload(fd, l, 0, l.payload_bytes, pcm.p);
	} catch (...) {
		AudioContainer::close_output(fd);
		throw;
	}
	AudioContainer::close_output(fd);
	status("");
}

} // namespace

double scale_for_peak(double peak, bool normalize)
{
	return ((peak > 1.0 || normalize) && peak > 0.0) ? 1.0 / peak : 1.0;
}

GpuPool::GpuPool(unsigned want)
{
	const int n = fir_gpu_device_count();
	if (n <= 0)
		throw GpuError(FIR_GPU_ERR_NO_DEVICE, "no usable B200 (sm_100) device; lowcut has no CPU path");
	const int use = want == 0 ? n : std::min<int>((int) want, n);
	// device_count() counts usable devices; walk the ordinals until `use` contexts exist
	for (int d = 0; (int) ctx_.size() < use && d < 64; ++d) {
		fir_gpu_ctx* c = nullptr;
		const int rc = fir_gpu_create(d, &c);
		if (rc == FIR_GPU_OK) ctx_.push_back(c);
		else if (rc == FIR_GPU_ERR_INVALID) break; // past the last ordinal
	}
	if (ctx_.empty()) throw GpuError(FIR_GPU_ERR_NO_DEVICE, std::string("cannot create a GPU context: ") + fir_gpu_last_error());
}

GpuPool::~GpuPool()
{
	for (fir_gpu_ctx* c : ctx_) fir_gpu_destroy(c);
}

void process_file(const std::filesystem::path& input_path, const std::filesystem::path& output_path,
                  const FilterOptions& opts, GpuPool& pool)
{
	std::vector<fir_gpu_ctx*> ctxs;
	for (size_t i = 0; i < pool.size(); ++i) ctxs.push_back(pool.ctx(i));
	run_file(input_path, output_path, opts, ctxs);
}

void process_batch(const std::vector<std::pair<std::filesystem::path, std::filesystem::path>>& jobs,
                   const FilterOptions& opts, GpuPool& pool)
{
	const size_t workers = std::min(pool.size(), jobs.size());
	if (workers <= 1) {
		for (const auto& j : jobs) run_file(j.first, j.second, opts, {pool.ctx(0)});
		return;
	}
	std::atomic<size_t> next{0};
	std::atomic<bool> failed{false};
	std::vector<std::exception_ptr> errs(workers);
	std::vector<std::thread> th;
	for (size_t w = 0; w < workers; ++w)
		th.emplace_back([&, w] {
			try {
				for (size_t i = next++; i < jobs.size() && !failed; i = next++)
					run_file(jobs[i].first, jobs[i].second, opts, {pool.ctx(w)});
			} catch (...) {
				errs[w] = std::current_exception();
				failed = true;
			}
		});
	for (auto& t : th) t.join();
	for (auto& e : errs)
		if (e) std::rethrow_exception(e);
}

} // namespace lowcut
