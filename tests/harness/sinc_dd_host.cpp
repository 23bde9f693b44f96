// Test harness (CPU): compiles the PRODUCT's double-double tap recipe
// (audio_fir_filter_b200/csrc/sinc_dd.cuh) with g++ and prints the taps, so that the
// arithmetic can be checked against the mpmath golden vectors without a GPU.  This is a
// test of the source, not a CPU path of the product: nothing ships or links it.
//   sinc_dd_host <fc_norm> <M>   ->  M+1 binary64 values on stdout (raw bytes)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../audio_fir_filter_b200/csrc/sinc_dd.cuh"

int main(int argc, char** argv)
{
	if (argc != 3) return 2;
	const double fc = std::strtod(argv[1], nullptr);
	const long long M = std::atoll(argv[2]);
	std::vector<firgpu::dd> lp(M + 1);
	firgpu::dd S = {0.0, 0.0};
	for (long long i = 0; i <= M; ++i) {
		lp[i] = firgpu::dd_lowpass_tap(i, M, fc);
		S = firgpu::dd_add(S, lp[i]);
	}
	std::vector<double> h(M + 1);
	for (long long i = 0; i <= M; ++i) h[i] = firgpu::dd_lowcut_tap(lp[i], S, i == M / 2);
	std::fwrite(h.data(), sizeof(double), h.size(), stdout);
	return 0;
}
