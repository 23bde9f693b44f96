// lowcut -- command-line host of the B200 low-cut FIR (scenarios of the reference's
// main.cp:84-148; exit codes and messages of its catch ladder :153-164).
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <filesystem>
#include <format>
#include <iostream>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "errors.hpp"
#include "options.hpp"
#include "process_file.hpp"

namespace fs = std::filesystem;
using namespace lowcut;

// Every output file is closed: leave without tearing the CUDA contexts and the pinned
// buffers down one by one (0.4 s for nothing; the OS reclaims them).
[[noreturn]] static void leave_now()
{
	std::cout.flush();
	std::cerr.flush();
	std::_Exit(EXIT_SUCCESS);
}

int main(int argc, char** argv)
{
	int exit_val = EXIT_SUCCESS;
	try {
		const CliOptions cli = parse_cli(argc, argv);
		if (cli.help) throw StopNoError(help_text());             // main.cp:64-66

		FilterOptions opts{cli.freq, cli.slope, cli.normalize, cli.verbose, cli.num_threads};
		// main.cp:75-76 keeps a thread count even though nothing here uses it
		if (opts.num_threads == 0) opts.num_threads = (unsigned) std::floor(std::thread::hardware_concurrency() * 0.7);
		if (opts.num_threads == 0) opts.num_threads = 4;

		std::vector<fs::path> paths(cli.paths.begin(), cli.paths.end());
		if (paths.size() < 2) throw UsageError("Invalid number of parameters. Need at least 2.");   // main.cp:150
		if (!(opts.slope > 0.0)) throw UsageError("Filter slope width must be greater than 0 Hz.");
		if (!(opts.freq > 0.0)) throw UsageError("Filter cutoff frequency must be greater than 0 Hz.");

		if (paths.size() == 2) {
			// Scenario 1: input file -> output file (main.cp:84-110)
			const fs::path& in = paths[0];
			const fs::path& out = paths[1];
			if (!fs::exists(in) || !fs::is_regular_file(in)) throw FileNotFound(in.string());
			if (fs::exists(out) && fs::is_directory(out))
				throw UsageError("With two parameters the second parameter must be a file path, not a directory.");
			if (in.extension() != out.extension())
				throw UsageError("Input and output file types (WAVE or AIFF) must be the same (extensions must match).");
			// the reference removes an existing output BEFORE it opens the input (main.cp:107): with
			// -O and input == output that destroys the only copy.  Refuse instead.
			if (fs::exists(out) && fs::equivalent(in, out))
				throw UsageError("Input and output are the same file: " + in.string());
			if (fs::exists(out) && !cli.overwrite) throw FileExists(out.string());
			// decide how many GPUs this file is worth and hide the others from CUDA before it starts
			GpuPool pool(restrict_devices_for_file(in, opts, cli.gpus));
			// main.cp:69-72 prints its resource line only when -v is NOT given; kept as is
			if (!opts.verbose) std::cout << std::format("Using up to {} GPU(s).", pool.limit()) << std::endl;
			// the reference removes an existing output up front (main.cp:107); here the result is
			// written to <out>.part and renamed over the old file only once it is complete, so a
			// failure leaves the old output in place
			if (opts.verbose) std::cout << std::format("  [{:8.3f} s since start] devices counted", uptime()) << std::endl;
			process_file(in, out, opts, pool);
			if (opts.verbose) std::cout << std::format("  [{:8.3f} s since start] done", uptime()) << std::endl;
			leave_now();
		} else {
			// Scenario 2: input files -> output directory (main.cp:112-148)
			const fs::path& dest = paths.back();
			if (fs::exists(dest)) {
				if (!fs::is_directory(dest))
					throw UsageError(std::format("Destination exists but is not a directory: {}", dest.string()));
			} else {
				if (dest.has_extension())
					throw UsageError(std::format(
						"Destination directory '{}' does not exist and has a suffix. Undefined scenario.", dest.string()));
				if (!opts.verbose) std::cout << std::format("Creating directory: {}", dest.string()) << std::endl;
				fs::create_directories(dest);
			}
			// the reference validates each file as it reaches it (main.cp:132-147); the
			// per-GPU workers run files concurrently, so validate them all up front
			std::vector<std::pair<fs::path, fs::path>> jobs;
			std::set<fs::path> claimed; // destinations of this run
			for (size_t i = 0; i + 1 < paths.size(); ++i) {
				const fs::path& in = paths[i];
				if (!fs::exists(in) || !fs::is_regular_file(in)) throw FileNotFound(in.string());
				fs::path out = dest / in.filename();
				if (fs::exists(out) && fs::equivalent(in, out))
					throw UsageError("Input and output are the same file: " + in.string());
				// Two inputs with one basename (a/x.wav b/x.wav out/) map to one destination.  The
				// reference runs files in sequence, so its second iteration finds the first one's
				// output and stops with FileExists (main.cp:140-142) unless -O is given; with -O the
				// later file wins.  Files run concurrently here, and two lanes writing one .part file
				// would corrupt it, so the clash is settled before anything runs.
				if (!claimed.insert(fs::weakly_canonical(out)).second) {
					if (!cli.overwrite) throw FileExists(out.string());
					for (auto& j : jobs)
						if (fs::weakly_canonical(j.second) == fs::weakly_canonical(out)) j.first = in; // last one wins
					continue;
				}
				if (fs::exists(out) && !cli.overwrite) throw FileExists(out.string());
				jobs.emplace_back(in, out);
			}
			// existing destinations (-O) are replaced one by one, by the rename that completes each
			// file (the reference removes each only when it reaches it, main.cp:144): if a file
			// fails, the old outputs of the files that were not reached are still there.
			// Nothing has touched CUDA yet: process_batch may fork one worker process per GPU.
			process_batch(jobs, opts, cli.gpus);
			leave_now();
		}
	} catch (const StopNoError& e) {
		const std::string s = e.what();
		if (!s.empty()) std::cout << s << std::endl;
	} catch (const BatchFailed&) {
		exit_val = EXIT_FAILURE; // the worker process that failed has already said why
	} catch (const std::exception& e) {
		std::cerr << e.what() << std::endl;
		exit_val = EXIT_FAILURE;
	} catch (...) {
		std::cerr << "Caught an unknown exception." << std::endl;
		exit_val = EXIT_FAILURE;
	}
	return exit_val;
}
