// Test helper: open a malformed file many times through AudioContainer and count the
// descriptors left open afterwards (tests/test_host_cli.py).
#include <dirent.h>

#include <cstdio>
#include <exception>

#include "audio_container.hpp"

static int open_fds()
{
	int n = 0;
	if (DIR* d = opendir("/proc/self/fd")) {
		while (readdir(d)) ++n;
		closedir(d);
	}
	return n;
}

int main(int argc, char** argv)
{
	if (argc < 2) return 2;
	const int before = open_fds();
	int threw = 0;
	for (int i = 0; i < 64; ++i) {
		try {
			lowcut::AudioContainer c(argv[1]);
		} catch (const std::exception&) {
			++threw;
		}
	}
	std::printf("threw=%d leaked=%d\n", threw, open_fds() - before);
	return threw == 64 ? 0 : 1;
}
