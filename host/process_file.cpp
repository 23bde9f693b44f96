#include "process_file.hpp"

#include <algorithm>
#include <atomic>
#include <barrier>
#include <chrono>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <new>
#include <exception>
#include <format>
#include <iostream>
#include <mutex>
#include <thread>

#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include "../include/fir_gpu.h"
#include "audio_container.hpp"
#include "errors.hpp"

namespace lowcut {
namespace {

std::mutex g_io; // serialises stdout between the per-GPU workers
const auto g_loaded = std::chrono::steady_clock::now();

void say(const std::string& s)
{
	std::lock_guard<std::mutex> l(g_io);
	std::cout << s << std::endl;
}

// LOWCUT_TRACE=1: where the host threads of this process spent their time, summed over all files
// and lanes, printed once at the end (tools/cli_timing.py shows it next to the wall time).
struct Trace {
	enum Phase { Read, Feed, WaitFir, CreateOutput, EncodeDownload, Write, Open, N };
	std::atomic<int64_t> ns[N] = {};
	std::atomic<int64_t> files{0};
	double t_fork = 0.0, t_ready = 0.0; // seconds since program start: this worker forked / its contexts were up
	bool on = std::getenv("LOWCUT_TRACE") != nullptr;
	static const char* name(int p)
	{
		static const char* n[N] = {"read file -> pinned", "feed (upload calls)", "wait for FIR / peak", "create output + allocate",
		                           "encode + download", "write pinned -> file", "open + parse"};
		return n[p];
	}
	void report()
	{
		if (!on || !files) return;
		std::string s = std::format("  trace (pid {}, {} files; forked at {:.3f} s, contexts ready at {:.3f} s, done at {:.3f} s; "
		                            "thread-seconds):", (long) ::getpid(), files.load(), t_fork, t_ready, uptime());
		for (int p = 0; p < N; ++p) s += std::format(" {} {:.3f};", name(p), (double) ns[p] * 1e-9);
		std::lock_guard<std::mutex> l(g_io);
		std::cout << s << std::endl;
	}
} g_trace;

struct Timed {
	Trace::Phase p;
	std::chrono::steady_clock::time_point t0;
	explicit Timed(Trace::Phase ph) : p(ph)
	{
		if (g_trace.on) t0 = std::chrono::steady_clock::now();
	}
	~Timed()
	{
		if (g_trace.on)
			g_trace.ns[p] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
	}
};

void check(int rc, const char* what)
{
	if (rc != FIR_GPU_OK) throw GpuError(rc, std::format("{}: {}", what, fir_gpu_last_error()));
}

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

// Text progress bar on stdout, driven by the library's per-chunk completion callback
// (the reference drives ProgressBar.h:18-55 from apply_filter_range, FilterCore.h:38-54).
// Only drawn on a terminal, so logs and pipes stay clean.
struct Progress {
	std::atomic<int64_t> done{0};
	int64_t total = 0;
	bool tty = false;
	int last_pct = -1;
	void draw()
	{
		if (!tty || total <= 0) return;
		const int pct = (int) (100.0 * (double) done.load() / (double) total);
		std::lock_guard<std::mutex> l(g_io);
		if (pct == last_pct) return;
		last_pct = pct;
		const int w = 60, fill = pct * w / 100;
		std::cout << '\r' << '[' << std::string(fill, '#') << std::string(w - fill, ' ') << "] " << pct << "%" << std::flush;
		if (pct >= 100) std::cout << std::endl;
	}
};

struct BlockProgress {
	Progress* bar;
	int64_t last = 0;
};

extern "C" void on_chunk_done(int64_t done_frames, int64_t, void* user)
{
	auto* b = static_cast<BlockProgress*>(user);
	b->bar->done += done_frames - b->last;
	b->last = done_frames;
	b->bar->draw();
}

constexpr uint64_t PIECE_BYTES = 16ull << 20; // file <-> pinned buffer <-> device granularity (pinning memory is slow: keep it small)

// One sample block of the file on one GPU, streamed: file reads overlap upload and FIR.
void filter_block(fir_gpu_ctx* ctx, const fir_gpu_kernel* k, const AudioContainer& in, const Block& b, double* peak,
                  Progress* bar, GpuPool& pool, size_t slot)
{
	const PcmLayout& l = in.pcm();
	const uint64_t fb = (uint64_t) l.channels * (l.bits / 8);
	const fir_gpu_pcm fmt = make_fmt(l, b.frames, b.halo_l, b.halo_r);
	BlockProgress bp{bar};
	// `bp` is the user pointer of callbacks that the CUDA runtime runs later, from its own thread:
	// whatever way this function is left (a short read, a failed call), the hook is removed and
	// the stream drained before `bp` goes out of scope
	struct HookGuard {
		fir_gpu_ctx* ctx;
		~HookGuard()
		{
			fir_gpu_set_progress(ctx, nullptr, nullptr);
			fir_gpu_synchronize(ctx);
		}
	} unhook{ctx};
	check(fir_gpu_set_progress(ctx, on_chunk_done, &bp), "fir_gpu_set_progress");
	check(fir_gpu_apply_begin(ctx, k, &fmt), "fir_gpu_apply_begin");
	const uint64_t first = (uint64_t) (b.start - b.halo_l) * fb, total = (uint64_t) (b.halo_l + b.frames + b.halo_r) * fb;
	unsigned char* bufs[2] = {pool.staging(slot, 0, std::min(total, PIECE_BYTES)),
	                          pool.staging(slot, 1, std::min(total, PIECE_BYTES))};
	int i = 0;
	for (uint64_t off = 0; off < total; off += PIECE_BYTES, i ^= 1) {
		const uint64_t n = std::min(PIECE_BYTES, total - off);
		{
			Timed tt(Trace::Read);
			in.read_payload(first + off, n, bufs[i]);          // while the previous piece uploads / filters
		}
		Timed tt(Trace::Feed);
		check(fir_gpu_apply_feed(ctx, bufs[i], n), "fir_gpu_apply_feed");
	}
	Timed tt(Trace::WaitFir);
	check(fir_gpu_apply_end(ctx), "fir_gpu_apply_end");
	check(fir_gpu_peak(ctx, peak), "fir_gpu_peak");            // ProcessFile.cp:92-96; waits for the FIR
}

// Encode the block with the common scale and write it into the output's sample chunk,
// piece by piece: the file write of one piece overlaps the encode + download of the next.
void encode_block(fir_gpu_ctx* ctx, const PcmLayout& l, const Block& b, double scale, int out_fd, GpuPool& pool,
                  size_t slot)
{
	const uint64_t fb = (uint64_t) l.channels * (l.bits / 8);
	const int64_t piece = std::max<int64_t>(2, (int64_t) (PIECE_BYTES / fb) & ~(int64_t) 1);
	unsigned char* bufs[2] = {pool.staging(slot, 2, (uint64_t) std::min(piece, b.frames) * fb),
	                          pool.staging(slot, 3, (uint64_t) std::min(piece, b.frames) * fb)};
	std::thread writer;
	std::exception_ptr werr;
	int i = 0;
	for (int64_t f = 0; f < b.frames; f += piece, i ^= 1) {
		const int64_t n = std::min(piece, b.frames - f);
		{
			Timed tt(Trace::EncodeDownload);
			check(fir_gpu_encode_range(ctx, scale, f, n, bufs[i]), "fir_gpu_encode_range");
		}
		if (writer.joinable()) writer.join(); // the other buffer is on its way to the file
		if (werr) std::rethrow_exception(werr);
		writer = std::thread([&, f, n, i] {
			try {
				Timed tt(Trace::Write);
				AudioContainer::write_payload(out_fd, l, (uint64_t) (b.start + f) * fb, (uint64_t) n * fb, bufs[i]);
			} catch (...) {
				werr = std::current_exception();
			}
		});
	}
	if (writer.joinable()) writer.join();
	if (werr) std::rethrow_exception(werr);
}

// ctxs[r] lives in pool slot slots[r] (its pinned staging buffers are taken from there).
void run_file(const std::filesystem::path& input_path, const std::filesystem::path& output_path,
              const FilterOptions& opts, GpuPool& pool, const std::vector<fir_gpu_ctx*>& ctxs,
              const std::vector<size_t>& slots)
{
	const auto t_start = std::chrono::steady_clock::now();
	auto since = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
	auto status = [&](const std::string& s) {
		if (opts.verbose) say(s);
	};
	auto stamp = [&](const char* what) {
		if (opts.verbose) say(std::format("  [{:8.3f} s] {}", since(), what));
	};

	status("Opening input file.");                                   // ProcessFile.cp:33-35
	const auto t_open = std::chrono::steady_clock::now();
	AudioContainer in(input_path);
	if (g_trace.on) {
		g_trace.ns[Trace::Open] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_open).count();
		++g_trace.files;
	}
	const PcmLayout& l = in.pcm();
	say("Processing file: " + input_path.filename().string());       // ProcessFile.cp:37 (unconditional)
	status(std::format("  {} {} ch, {} bit {}, {} Hz, {} frames, {} chunks", in.type_name(), l.channels, l.bits,
	                   l.big_endian ? "big-endian" : "little-endian", l.sample_rate, l.frames, in.chunks().size()));
	if (!(l.sample_rate > 0.0)) throw FormatError(input_path.string() + ": sample rate is zero");
	const double fc = opts.freq / l.sample_rate, bw = opts.slope / l.sample_rate; // ProcessFile.cp:48-49
	// the output appears under its name only once it is complete (the reference writes in place and
	// leaves a partial file behind when it fails)
	std::filesystem::path part_path = output_path;
	part_path += ".part";

	// sample-block mode only pays when every GPU gets a sizeable block
	size_t world = ctxs.size();
	if (world > 1 && (int64_t) l.frames < (int64_t) world * (1 << 18)) world = 1;

	// the output starts as a copy of every non-sample byte (ProcessFile.cp:104-112);
	// done first so that it overlaps nothing it could disturb
	status("Reading samples.");                                      // ProcessFile.cp:39-41 (streamed below)
	status("Creating sinc kernel for this file's sample rate.");
	double scale = 1.0, peak = 0.0;
	int out_fd = -1;
	if (l.frames > 0) {
		// The output is created (every non-sample byte copied, ProcessFile.cp:104-112, and its pages
		// allocated) on a thread of its own while the samples are read and filtered: it is a .part
		// file until the result is complete, so nothing is lost by creating it before the peak is known.
		std::exception_ptr out_err;
		std::thread out_creator([&] {
			try {
				Timed tt(Trace::CreateOutput);
				out_fd = in.create_output(part_path);
			} catch (...) {
				out_err = std::current_exception();
			}
		});
		struct JoinOnExit {
			std::thread& t;
			~JoinOnExit()
			{
				if (t.joinable()) t.join();
			}
		} join_creator{out_creator};
		std::vector<fir_gpu_kernel*> ks(world);
		long long half_len = 0;
		for (size_t r = 0; r < world; ++r) ks[r] = pool.kernel(slots[r], fc, bw, &half_len);
		const std::vector<Block> blocks = plan_blocks((int64_t) l.frames, world, half_len);
		status(std::format("  {} taps{}", fir_gpu_kernel_num_taps(ks[0]),
		                   world > 1 ? std::format(", {} sample blocks of up to {} frames, halo {} frames each side", world,
		                                           blocks[0].frames, half_len)
		                             : std::string()));
		stamp("kernel built");
		status("Filtering.");
		Progress bar;
		bar.total = (int64_t) l.frames;
		bar.tty = ::isatty(1) != 0;
		std::vector<double> peaks(world, 0.0);
		std::vector<std::exception_ptr> errs(world);
		std::atomic<bool> failed{false};
		std::barrier sync((std::ptrdiff_t) world + 1);
		// Sample-block mode: the peak is one ncclAllReduce(max) over the blocks' device scalars.
		// The communicator comes up on its own thread while the blocks are read and filtered.
		std::vector<fir_gpu_ctx*> block_ctxs(ctxs.begin(), ctxs.begin() + (std::ptrdiff_t) world);
		// Bringing an NCCL communicator up costs a second or so: it is worth its 8 bytes when the blocks
		// filter for at least that long (the start-up hides under the FIR); a short file forced onto
		// several GPUs takes the max of the n scalars on the host instead.  LOWCUT_NCCL=1 / 0 overrides.
		const double block_seconds =
			2.0 * (double) fir_gpu_kernel_num_taps(ks[0]) * (double) blocks[0].frames * l.channels / 35.0e12;
		bool use_nccl = world > 1 && block_seconds >= 1.0;
		if (const char* e = std::getenv("LOWCUT_NCCL")) use_nccl = world > 1 && std::atoi(e) != 0;
		std::thread comm_up;
		if (use_nccl) comm_up = std::thread([&] { fir_gpu_comm_prepare(block_ctxs.data(), (int) world); });
		std::vector<std::thread> th;
		for (size_t r = 0; r < world; ++r)
			th.emplace_back([&, r] {
				const Block& b = blocks[r];
				try {
					if (b.frames) filter_block(ctxs[r], ks[r], in, b, &peaks[r], &bar, pool, slots[r]);
				} catch (...) {
					errs[r] = std::current_exception();
					failed = true;
				}
				sync.arrive_and_wait(); // every peak is known: main decides the scale, creates the output
				sync.arrive_and_wait();
				if (failed || !b.frames) return;
				try {
					encode_block(ctxs[r], l, b, scale, out_fd, pool, slots[r]);
				} catch (...) {
					errs[r] = std::current_exception();
					failed = true;
				}
			});
		sync.arrive_and_wait();
		if (comm_up.joinable()) comm_up.join();
		out_creator.join(); // the output exists (or could not be created) before anything is decided about it
		if (out_err) {
			errs.push_back(out_err);
			failed = true;
		}
		stamp("filtered, peak known");
		try {
			if (!failed) {
				// ProcessFile.cp:92-96 is a max over the whole file: all-reduce MAX of the blocks' peaks
				// (NCCL over NVLink, on the device scalars).  Without a loadable NCCL the n doubles that
				// fir_gpu_peak already brought back are compared here instead.
				const char* how = "single block";
				if (use_nccl && fir_gpu_allreduce_peak(block_ctxs.data(), (int) world, &peak) == FIR_GPU_OK) {
					how = "ncclAllReduce(max) over the blocks";
					if (peak != *std::max_element(peaks.begin(), peaks.end()))
						throw GpuError(FIR_GPU_ERR_STATE, "the all-reduced peak is not the max of the block peaks");
				} else {
					if (world > 1)
						how = use_nccl ? "host max of the blocks' peaks (NCCL could not be loaded)"
						               : "host max of the blocks' peaks (blocks too short to hide an NCCL start-up; LOWCUT_NCCL=1 forces it)";
					peak = *std::max_element(peaks.begin(), peaks.end());
				}
				status(std::format("  peak exchange: {}", how));
				scale = scale_for_peak(peak, opts.normalize);             // ProcessFile.cp:98-101
				if (scale != 1.0) status("Doing audio normalize.");
				status("Writing output file.");
			}
		} catch (...) {
			errs.push_back(std::current_exception());
			failed = true;
		}
		stamp("output file created");
		sync.arrive_and_wait();
		for (auto& t : th) t.join();
		if (out_fd >= 0) AudioContainer::close_output(out_fd);
		stamp("encoded and written");
		for (auto& e : errs)
			if (e) {
				std::error_code ec;
				std::filesystem::remove(part_path, ec); // never leave a half-written output behind
				std::rethrow_exception(e);
			}
		try {
			std::filesystem::rename(part_path, output_path);
		} catch (...) {
			std::error_code ec;
			std::filesystem::remove(part_path, ec); // never leave a .part behind, whatever went wrong
			throw;
		}
		for (size_t r = 0; r < world; ++r) status(std::format("  GPU {}:{}", r, timing_line(ctxs[r])));
	} else {
		status("Writing output file.");
		AudioContainer::close_output(in.create_output(part_path));
		std::filesystem::rename(part_path, output_path);
	}
	status(std::format("  peak {:.9f}, scale {:.9f}, {:.3f} s", peak, scale,
	                   std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count()));
	status("");
}

} // namespace

double uptime() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - g_loaded).count(); }

double scale_for_peak(double peak, bool normalize)
{
	return ((peak > 1.0 || normalize) && peak > 0.0) ? 1.0 / peak : 1.0;
}

GpuPool::GpuPool(unsigned want)
{
	const int n = fir_gpu_device_count();
	if (n <= 0)
		throw GpuError(FIR_GPU_ERR_NO_DEVICE, "no usable B200 (sm_100) device; lowcut has no CPU path");
	forced_ = want != 0;
	const int use = want == 0 ? n : std::min<int>((int) want, n);
	for (int d = 0; d < use; ++d) ordinals_.push_back(d);
	lanes_.resize(LANES * ordinals_.size());
}

GpuPool::~GpuPool()
{
	for (Lane& l : lanes_) {
		for (const Lane::CachedKernel& c : l.kernels) fir_gpu_kernel_free(c.k);
		for (unsigned char* b : l.buf) fir_gpu_host_free(b);
		fir_gpu_destroy(l.ctx);
	}
}

std::vector<fir_gpu_ctx*> GpuPool::acquire(size_t n_devices, size_t subs, std::vector<size_t>* slots)
{
	n_devices = std::max<size_t>(1, std::min(n_devices, ordinals_.size()));
	std::vector<size_t> want;
	subs = std::clamp<size_t>(subs, 1, LANES);
	for (size_t sub = 0; sub < subs; ++sub)
		for (size_t d = 0; d < n_devices; ++d) want.push_back(LANES * d + sub);
	std::vector<std::thread> th;
	std::vector<std::string> errs(lanes_.size());
	for (size_t slot : want)
		if (!lanes_[slot].ctx)
			th.emplace_back([this, slot, &errs] {
				if (fir_gpu_create(ordinals_[slot / LANES], &lanes_[slot].ctx) != FIR_GPU_OK) errs[slot] = fir_gpu_last_error();
			});
	for (auto& t : th) t.join();
	std::vector<fir_gpu_ctx*> out;
	for (size_t slot : want) {
		if (!lanes_[slot].ctx) throw GpuError(FIR_GPU_ERR_NO_DEVICE, "cannot create a GPU context: " + errs[slot]);
		out.push_back(lanes_[slot].ctx);
	}
	if (slots) *slots = want;
	return out;
}

fir_gpu_ctx* GpuPool::acquire_slot(size_t slot)
{
	Lane& l = lanes_.at(slot); // a lane belongs to one thread: no lock
	if (!l.ctx && fir_gpu_create(ordinals_[slot / LANES], &l.ctx) != FIR_GPU_OK)
		throw GpuError(FIR_GPU_ERR_NO_DEVICE, std::string("cannot create a GPU context: ") + fir_gpu_last_error());
	return l.ctx;
}

fir_gpu_kernel* GpuPool::kernel(size_t slot, double fc, double bw, long long* half_len)
{
	Lane& l = lanes_.at(slot); // a lane is used by one thread at a time: no lock
	for (const Lane::CachedKernel& c : l.kernels)
		if (c.fc == fc && c.bw == bw) {
			*half_len = c.half_len;
			return c.k;
		}
	fir_gpu_kernel* k = nullptr;
	int64_t h = 0;
	check(fir_gpu_build_kernel(l.ctx, fc, bw, &k, &h), "fir_gpu_build_kernel");
	l.kernels.push_back({fc, bw, k, (long long) h});
	*half_len = (long long) h;
	return k;
}

unsigned char* GpuPool::staging(size_t slot, int which, size_t bytes)
{
	Lane& l = lanes_.at(slot);
	if (l.cap[which] < bytes) {
		fir_gpu_host_free(l.buf[which]);
		l.buf[which] = static_cast<unsigned char*>(fir_gpu_host_alloc(bytes));
		l.cap[which] = l.buf[which] ? bytes : 0;
		if (!l.buf[which])
			throw GpuError(FIR_GPU_ERR_NOMEM, std::format("pinned host buffer of {} bytes: {}", bytes, fir_gpu_last_error()));
	}
	return l.buf[which];
}

// Seconds of FIR one B200 needs for this much work (35 TFLOP/s measured), used to
// decide how many GPUs are worth starting.
static double fir_seconds(double taps, double frames, double channels)
{
	return 2.0 * taps * frames * channels / 35.0e12;
}

static double estimate_file_seconds(const std::filesystem::path& p, const FilterOptions& opts)
{
	AudioContainer in(p);
	in.advise_willneed(); // start the read-ahead now: it runs while the CUDA context comes up
	const PcmLayout& l = in.pcm();
	if (!(l.sample_rate > 0.0) || !(opts.slope > 0.0)) return 0.0;
	return fir_seconds(4.0 * l.sample_rate / opts.slope + 1.0, (double) l.frames, l.channels);
}

// How many GPUs a job of `seconds` of FIR is worth.  Every GPU a run touches costs about half a
// second of wall time that no parallelism buys back: the driver creates (and, at exit, destroys)
// contexts one after the other, machine-wide -- measured on the 8-GPU box: 0.26 s + 0.24 s per
// device, in one process or in eight.  With g GPUs a job takes about seconds/g + GPU_COST*g, least
// at g = sqrt(seconds / GPU_COST): 8 GPUs from ~30 s of FIR upwards (config 3: 29 s), 3 for the
// 256-file batch of config 4 (3.9 s), 1 for a ten-minute file.
constexpr double GPU_COST_SECONDS = 0.5;

static size_t gpus_worth(size_t limit, double seconds)
{
	const double g = std::round(std::sqrt(std::max(seconds, 0.0) / GPU_COST_SECONDS));
	return (size_t) std::clamp<double>(g, 1.0, (double) std::max<size_t>(limit, 1));
}

static size_t gpus_worth_starting(const GpuPool& pool, double seconds)
{
	if (pool.forced()) return pool.limit();
	return gpus_worth(pool.limit(), seconds);
}

void process_file(const std::filesystem::path& input_path, const std::filesystem::path& output_path,
                  const FilterOptions& opts, GpuPool& pool)
{
	const size_t world = gpus_worth_starting(pool, estimate_file_seconds(input_path, opts));
	std::vector<size_t> slots;
	const std::vector<fir_gpu_ctx*> ctxs = pool.acquire(world, 1, &slots);
	if (opts.verbose) say(std::format("  [{:8.3f} s since start] {} GPU context(s) ready", uptime(), ctxs.size()));
	run_file(input_path, output_path, opts, pool, ctxs, slots);
}

void process_batch(const std::vector<std::pair<std::filesystem::path, std::filesystem::path>>& jobs,
                   const FilterOptions& opts, GpuPool& pool)
{
	double seconds = 0.0;
	for (const auto& j : jobs) seconds += estimate_file_seconds(j.first, opts);
	const size_t gpus = std::min(gpus_worth_starting(pool, seconds), jobs.size());
	// several lanes per GPU: while one file filters, others read, upload, download, write.
	// Per file the host side (page-cache read, output creation, write) costs a few times the
	// FIR of a short file, so up to LANES files are in flight per GPU (LOWCUT_LANES overrides).
	size_t lanes = GpuPool::LANES;
	if (const char* e = std::getenv("LOWCUT_LANES")) lanes = (size_t) std::max(1, std::atoi(e));
	lanes = std::clamp<size_t>(std::min(lanes, (jobs.size() + gpus - 1) / gpus), 1, GpuPool::LANES);
	std::atomic<size_t> next{0};
	std::atomic<bool> failed{false};
	std::mutex err_mutex;
	std::vector<std::exception_ptr> errs;
	std::vector<std::thread> th;
	// Every lane is a thread that brings up ITS context and then pulls files: the first lane that
	// exists starts filtering while the driver is still creating the others (it creates contexts
	// one after the other: 4 lanes on one device took 1.4 s to be "all ready", the first 0.3 s).
	std::atomic<bool> first_ready{false};
	auto lane_loop = [&](size_t slot) {
		try {
			fir_gpu_ctx* ctx = pool.acquire_slot(slot);
			if (!first_ready.exchange(true)) g_trace.t_ready = uptime();
			for (size_t i = next++; i < jobs.size() && !failed; i = next++)
				run_file(jobs[i].first, jobs[i].second, opts, pool, {ctx}, {slot});
		} catch (...) {
			std::lock_guard<std::mutex> l(err_mutex);
			errs.push_back(std::current_exception());
			failed = true;
		}
	};
	for (size_t sub = 0; sub < lanes; ++sub) // lane 0 of every device first, then the second lanes, ...
		for (size_t d = 0; d < gpus; ++d) th.emplace_back(lane_loop, GpuPool::LANES * d + sub);
	for (auto& t : th) t.join();
	g_trace.report();
	for (auto& e : errs)
		if (e) std::rethrow_exception(e);
}

size_t visible_device_count_without_cuda(std::vector<std::string>* visible_ids)
{
	std::error_code ec;
	size_t n = 0;
	for (auto it = std::filesystem::directory_iterator("/proc/driver/nvidia/gpus", ec);
	     !ec && it != std::filesystem::directory_iterator(); it.increment(ec))
		++n;
	if (n == 0) {
		// containers often hide /proc/driver/nvidia but must expose the device nodes CUDA opens:
		// /dev/nvidia0, /dev/nvidia1, ... (not nvidiactl, nvidia-uvm, ...)
		ec.clear();
		for (auto it = std::filesystem::directory_iterator("/dev", ec); !ec && it != std::filesystem::directory_iterator();
		     it.increment(ec)) {
			const std::string name = it->path().filename().string();
			if (name.size() > 6 && name.compare(0, 6, "nvidia") == 0 &&
			    name.find_first_not_of("0123456789", 6) == std::string::npos)
				++n;
		}
	}
	std::vector<std::string> ids;
	if (const char* e = std::getenv("CUDA_VISIBLE_DEVICES")) {
		// the listed entries (ordinals or UUIDs), up to the first invalid one, as CUDA reads it
		std::string s(e), item;
		for (size_t i = 0; i <= s.size(); ++i) {
			if (i == s.size() || s[i] == ',') {
				if (item.empty()) break;
				const bool ordinal = item.find_first_not_of("0123456789") == std::string::npos;
				if (ordinal && (size_t) std::atoll(item.c_str()) >= n) break;
				ids.push_back(item);
				item.clear();
			} else {
				item += s[i];
			}
		}
	} else {
		for (size_t i = 0; i < n; ++i) ids.push_back(std::to_string(i));
	}
	if (n == 0) ids.clear();
	if (visible_ids) *visible_ids = ids;
	return ids.size();
}

unsigned restrict_devices_for_file(const std::filesystem::path& input_path, const FilterOptions& opts, unsigned want_gpus)
{
	std::vector<std::string> ids;
	const size_t n_dev = visible_device_count_without_cuda(&ids);
	if (n_dev == 0) return want_gpus; // cannot tell without CUDA: the pool will ask it
	const size_t world = want_gpus ? std::min<size_t>(want_gpus, n_dev)
	                               : gpus_worth(n_dev, estimate_file_seconds(input_path, opts));
	std::string list;
	for (size_t i = 0; i < world; ++i) list += (i ? "," : "") + ids[i];
	::setenv("CUDA_VISIBLE_DEVICES", list.c_str(), 1);
	return (unsigned) world;
}

namespace {

// The page the worker processes of a batch share: the next job to hand out, and whether anybody failed.
struct BatchShared {
	std::atomic<size_t> next;
	std::atomic<int> failed;
};
static_assert(std::atomic<size_t>::is_always_lock_free && std::atomic<int>::is_always_lock_free);

constexpr int EXIT_NO_DEVICE = 3; // a worker that found no usable device (and so did nothing)

// One worker process: the only device it sees is its own; up to LANES files in flight on it.
[[noreturn]] void batch_worker(const std::vector<std::pair<std::filesystem::path, std::filesystem::path>>& jobs,
                               const FilterOptions& opts, BatchShared* sh, size_t lanes)
{
	int code = EXIT_SUCCESS;
	g_trace.t_fork = uptime();
	try {
		std::unique_ptr<GpuPool> pool_ptr;
		try {
			pool_ptr = std::make_unique<GpuPool>(1);
		} catch (const GpuError&) {
			// this worker's device cannot be used (a device node that is not ours to open, a GPU that is
			// not a B200): it has taken no job -- the other workers share out all of them
			std::_Exit(EXIT_NO_DEVICE);
		}
		GpuPool& pool = *pool_ptr;
		lanes = std::clamp<size_t>(lanes, 1, GpuPool::LANES);
		std::vector<std::exception_ptr> errs(lanes);
		std::vector<std::thread> th;
		std::atomic<bool> first_ready{false};
		for (size_t w = 0; w < lanes; ++w)
			th.emplace_back([&, w] {
				try {
					fir_gpu_ctx* ctx = pool.acquire_slot(w); // this lane's context; the lanes that exist already work
					if (!first_ready.exchange(true)) g_trace.t_ready = uptime();
					for (size_t i = sh->next++; i < jobs.size() && !sh->failed; i = sh->next++)
						run_file(jobs[i].first, jobs[i].second, opts, pool, {ctx}, {w});
				} catch (...) {
					errs[w] = std::current_exception();
					sh->failed = 1; // nobody starts another file (the reference stops at the first error, main.cp:157)
				}
			});
		for (auto& t : th) t.join();
		for (auto& e : errs)
			if (e) std::rethrow_exception(e);
	} catch (const std::exception& e) {
		sh->failed = 1;
		std::lock_guard<std::mutex> l(g_io);
		std::cerr << e.what() << std::endl;
		code = EXIT_FAILURE;
	}
	g_trace.report();
	std::cout.flush();
	std::cerr.flush();
	std::_Exit(code); // the files are closed; tearing contexts down one by one buys nothing
}

} // namespace

size_t process_batch(const std::vector<std::pair<std::filesystem::path, std::filesystem::path>>& jobs,
                     const FilterOptions& opts, unsigned want_gpus)
{
	double seconds = 0.0;
	for (const auto& j : jobs) seconds += estimate_file_seconds(j.first, opts);
	std::vector<std::string> ids;
	size_t n_dev = visible_device_count_without_cuda(&ids);
	if (const char* e = std::getenv("LOWCUT_WORKER_DEVICES")) {
		// test hook: the CUDA_VISIBLE_DEVICES value of each worker process, e.g. "0,0" = two workers
		// sharing one GPU (exercises the fork / shared-counter path on a one-GPU box)
		ids.clear();
		std::string s(e), item;
		for (size_t i = 0; i <= s.size(); ++i) {
			if (i == s.size() || s[i] == ',') {
				if (!item.empty()) ids.push_back(item);
				item.clear();
			} else {
				item += s[i];
			}
		}
		n_dev = ids.size();
	}
	size_t gpus = want_gpus ? std::min<size_t>(want_gpus, std::max<size_t>(n_dev, 1)) : gpus_worth(n_dev, seconds);
	gpus = std::min(gpus, jobs.size());
	if (n_dev == 0 || gpus <= 1 || std::getenv("LOWCUT_SINGLE_PROCESS")) {
		// one GPU (or no way to count them without CUDA): everything in this process
		GpuPool pool(n_dev == 0 ? want_gpus : (unsigned) std::max<size_t>(gpus, 1));
		if (!opts.verbose) say(std::format("Using up to {} GPU(s).", pool.limit()));
		process_batch(jobs, opts, pool);
		return pool.limit();
	}
	if (!opts.verbose) say(std::format("Using up to {} GPU(s).", gpus));
	if (opts.verbose || std::getenv("LOWCUT_TRACE")) say(std::format("  one worker process per GPU ({} processes)", gpus));
	size_t lanes = GpuPool::LANES;
	if (const char* e = std::getenv("LOWCUT_LANES")) lanes = (size_t) std::max(1, std::atoi(e));
	lanes = std::min(lanes, (jobs.size() + gpus - 1) / gpus);

	void* page = ::mmap(nullptr, sizeof(BatchShared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
	if (page == MAP_FAILED) throw std::runtime_error("cannot map the batch's shared page");
	BatchShared* sh = new (page) BatchShared{};
	std::cout.flush();
	std::cerr.flush();
	std::vector<pid_t> kids;
	for (size_t g = 0; g < gpus; ++g) {
		const pid_t pid = ::fork();
		if (pid < 0) {
			sh->failed = 1;
			break;
		}
		if (pid == 0) {
			::setenv("CUDA_VISIBLE_DEVICES", ids[g].c_str(), 1); // before this process first touches CUDA
			batch_worker(jobs, opts, sh, lanes);
		}
		kids.push_back(pid);
	}
	bool bad = kids.size() != gpus;
	size_t worked = 0;
	for (pid_t pid : kids) {
		int st = 0;
		while (::waitpid(pid, &st, 0) < 0 && errno == EINTR) {}
		if (WIFEXITED(st) && WEXITSTATUS(st) == EXIT_SUCCESS) ++worked;
		else if (WIFEXITED(st) && WEXITSTATUS(st) == EXIT_NO_DEVICE) continue; // did nothing, took nothing
		else {
			if (WIFSIGNALED(st)) std::cerr << std::format("a GPU worker process died of signal {}", WTERMSIG(st)) << std::endl;
			bad = true;
		}
	}
	bad = bad || sh->failed;
	const bool all_done = sh->next.load() >= jobs.size();
	::munmap(page, sizeof(BatchShared));
	if (bad) throw BatchFailed();
	if (worked == 0 || !all_done)
		throw GpuError(FIR_GPU_ERR_NO_DEVICE, "no usable B200 (sm_100) device; lowcut has no CPU path");
	return worked;
}

} // namespace lowcut
