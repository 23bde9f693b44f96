// tools/tmpfs_write_probe.cpp -- how fast can this box create FILES files of MB megabytes each in
// DIR with THREADS threads doing nothing else (ftruncate [+ fallocate] + 16 MB pwrites from a
// private buffer)?  The ceiling the 256-file batch of tools/cli_timing.py is compared with: lowcut
// has to put the same 22 GB of fresh pages into the same file system.
//   tmpfs_write_probe DIR THREADS FILES MB [fallocate=0|1]
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

int main(int argc, char** argv)
{
	if (argc < 5) return 2;
	const std::string dir = argv[1];
	const int threads = std::atoi(argv[2]), files = std::atoi(argv[3]);
	const size_t bytes = (size_t) std::atol(argv[4]) * 1000000;
	const bool prealloc = argc > 5 && std::atoi(argv[5]) != 0;
	std::atomic<int> next{0};
	const auto t0 = std::chrono::steady_clock::now();
	std::vector<std::thread> th;
	for (int t = 0; t < threads; ++t)
		th.emplace_back([&] {
			std::vector<char> buf(16u << 20, 1);
			for (int f = next++; f < files; f = next++) {
				const std::string p = dir + "/probe_" + std::to_string(f);
				const int fd = ::open(p.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
				if (fd < 0) std::exit(1);
				if (::ftruncate(fd, (off_t) bytes)) std::exit(1);
				if (prealloc) (void) ::fallocate(fd, 0, 0, (off_t) bytes);
				for (size_t off = 0; off < bytes; off += buf.size())
					if (::pwrite(fd, buf.data(), std::min(buf.size(), bytes - off), (off_t) off) < 0) std::exit(1);
				::close(fd);
			}
		});
	for (auto& x : th) x.join();
	const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	std::printf("{\"probe\": \"create %d files x %zu MB in %s, %d threads, fallocate %d\", \"seconds\": %.3f, \"gb_per_s\": %.2f}\n",
	            files, bytes / 1000000, dir.c_str(), threads, (int) prealloc, s, (double) files * bytes / s / 1e9);
	for (int f = 0; f < files; ++f) ::unlink((dir + "/probe_" + std::to_string(f)).c_str());
	return 0;
}
