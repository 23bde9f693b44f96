// process_file.hpp -- per-file orchestration of the lowcut hot path on B200s.
//
// Mirrors the reference's ProcessFile.h (FilterOptions :13-19, process_file :21-25);
// the arithmetic of ProcessFile.cp:41-101,117 runs behind the C-ABI of
// include/fir_gpu.h, the container work of :34-35,105-116 stays here.
#pragma once
#include <filesystem>
#include <memory>
#include <string>
#include <utility>
#include <vector>

struct fir_gpu_ctx;
struct fir_gpu_kernel;

namespace lowcut {

struct FilterOptions {
	double freq = 15.0;       // -f, Hz
	double slope = 10.0;      // -s, Hz
	bool normalize = false;   // -n
	bool verbose = false;     // -v
	unsigned num_threads = 0; // -t: accepted, unused (the device grid replaces the thread fan-out)
};

// The B200s this run may use, one fir_gpu context (and one set of pinned staging
// buffers) per device, created on first use: starting a CUDA context costs far more
// than filtering a short file, so a GPU is only brought up when there is work for it.
// Throws GpuError if there is no usable device: this program has no CPU path.
class GpuPool {
public:
	explicit GpuPool(unsigned want /* 0 = choose by the amount of work, up to every usable device */);
	~GpuPool();
	GpuPool(const GpuPool&) = delete;
	GpuPool& operator=(const GpuPool&) = delete;
	size_t limit() const { return ordinals_.size(); } // devices that may be used
	bool forced() const { return forced_; }             // -g N given: use exactly that many when possible
	// Lanes: up to LANES contexts per device (slot = LANES*device + sub).  Block mode uses
	// sub 0 of the first n devices; batch mode several subs, so that the file I/O, upload
	// and download of some files overlap the FIR of another on the same GPU.  Created in
	// parallel on demand.
	static constexpr size_t LANES = 4;
	std::vector<fir_gpu_ctx*> acquire(size_t n_devices, size_t subs, std::vector<size_t>* slots = nullptr);
	// One lane's context, created by the thread that will use it (batch mode: every lane brings up
	// its own context and starts on the files at once; the others are still being created).
	fir_gpu_ctx* acquire_slot(size_t slot);
	// Two pinned buffers of `bytes` for upload and two for download, per lane, reused across files.
	unsigned char* staging(size_t slot, int which /*0..3*/, size_t bytes);
	// The low-cut kernel for (fc, bw) on that lane, built once and kept: files of one batch
	// mostly share a sample rate (ProcessFile.cp:48-50 rebuilds it per file).
	struct fir_gpu_kernel* kernel(size_t slot, double fc, double bw, long long* half_len);

private:
	struct Lane {
		fir_gpu_ctx* ctx = nullptr;
		unsigned char* buf[4] = {nullptr, nullptr, nullptr, nullptr};
		size_t cap[4] = {0, 0, 0, 0};
		struct CachedKernel {
			double fc, bw;
			struct fir_gpu_kernel* k;
			long long half_len;
		};
		std::vector<CachedKernel> kernels;
	};
	std::vector<int> ordinals_;
	std::vector<Lane> lanes_;
	bool forced_ = false;
};

// One file (ProcessFile.cp:27-120).  A file long enough is split into contiguous
// sample blocks with (taps-1) halo across every GPU of the pool (one peak
// max-reduction in between); a short one runs on the pool's first GPU.
void process_file(const std::filesystem::path& input_path, const std::filesystem::path& output_path,
                  const FilterOptions& opts, GpuPool& pool);

// Batch scenario (main.cp:132-147): whole files dealt to GPUs, no communication.  The first
// failure stops the hand-out of further files and is rethrown once the files in flight are done.
// This overload runs everything in the calling process (one context per lane and device).
void process_batch(const std::vector<std::pair<std::filesystem::path, std::filesystem::path>>& jobs,
                   const FilterOptions& opts, GpuPool& pool);

// The same, deciding itself how many GPUs are worth starting (want_gpus = 0) and -- when that is
// more than one -- running ONE PROCESS PER GPU: the CUDA contexts then come up in parallel
// (inside one process the driver creates them one after the other, ~0.5 s each on an 8-GPU box,
// several times the FIR of a whole batch), and the files are handed out from a counter in shared
// memory.  Must be called before anything in this process has touched CUDA.  Returns the number
// of GPUs used.  Throws BatchFailed when a worker process failed (its message is already on stderr).
size_t process_batch(const std::vector<std::pair<std::filesystem::path, std::filesystem::path>>& jobs,
                     const FilterOptions& opts, unsigned want_gpus);

// How many GPUs one file is worth (sample-block mode needs >= 1 s of FIR per extra GPU), decided
// WITHOUT touching CUDA, and CUDA_VISIBLE_DEVICES narrowed to exactly those devices: cuInit then
// initialises one device instead of eight (0.1 s instead of 0.6 s on an 8-GPU box -- more than
// filtering a ten-minute file takes).  Returns the number to hand to GpuPool (0 = could not tell).
unsigned restrict_devices_for_file(const std::filesystem::path& input_path, const FilterOptions& opts, unsigned want_gpus);

// NVIDIA devices this process may use, WITHOUT initialising CUDA (so that the answer can be had
// before a fork): the entries of /proc/driver/nvidia/gpus (or, where a container hides that, the
// /dev/nvidia<N> device nodes), narrowed by CUDA_VISIBLE_DEVICES.  0 = unknown: callers fall back
// to asking CUDA.
size_t visible_device_count_without_cuda(std::vector<std::string>* visible_ids = nullptr);

// Seconds since the program was loaded (the -v time stamps; start-up cost is part of what a user waits for).
double uptime();

// ProcessFile.cp:98: normalise when the peak exceeds full scale or -n is given.
double scale_for_peak(double peak, bool normalize);

} // namespace lowcut
