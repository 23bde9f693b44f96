// Interface shim, TEST INFRASTRUCTURE ONLY (see oracle/ref_filtercore.cpp).
// Stands in for diskerror/c_lib's <WindowedSinc.h> (absent, un-pinned).  The
// taps are injected (they come from oracle_build_lowcut); the members are the
// ones FilterCore.h calls: getMo2() (:29) and the three fms() forms (:59,67,74).
// fms sign convention (the only reading consistent with a zero-padded FIR,
// SURVEY.md 3.3): fms(it) = all M+1 taps; fms(it, +n) = first n taps;
// fms(it, -n) = last n taps; each against consecutive samples starting at it.
#pragma once
#include <cstdint>
#include <vector>
namespace Diskerror {
template <typename T>
class WindowedSinc {
	std::vector<T> _h;
	template <typename It>
	T dot(const T* h, It it, std::int64_t n) const {
		T acc = 0;
#pragma omp simd reduction(+ : acc)
		for (std::int64_t k = 0; k < n; ++k) acc += h[k] * static_cast<T>(it[k]);
		return acc;
	}
public:
	WindowedSinc(const T* taps, std::size_t n) : _h(taps, taps + n) {}
	std::size_t getMo2() const { return (_h.size() - 1) / 2; }
	template <typename It> T fms(It it) const { return dot(_h.data(), it, (std::int64_t) _h.size()); }
	template <typename It> T fms(It it, std::int32_t n) const {
		if (n >= 0) return dot(_h.data(), it, n);
		return dot(_h.data() + (_h.size() - (std::size_t)(-n)), it, -n);
	}
};
}
