// Interface shim, TEST INFRASTRUCTURE ONLY (see oracle/ref_filtercore.cpp).
// Stands in for diskerror/c_lib's <VectorMath.h> (absent, un-pinned).  Only the
// members the reference's hot path touches exist: ctor(n) (ProcessFile.cp:58),
// size() (FilterCore.h:28), begin() (:59), operator[] (:59,67,74), max_mag()
// (ProcessFile.cp:94).
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>
namespace Diskerror {
template <typename T>
class VectorMath {
	std::vector<T> _v;
public:
	VectorMath() = default;
	explicit VectorMath(std::size_t n) : _v(n) {}
	VectorMath(const T* p, std::size_t n) : _v(p, p + n) {}
	std::size_t size() const { return _v.size(); }
	const T* begin() const { return _v.data(); }
	T* begin() { return _v.data(); }
	T& operator[](std::size_t i) { return _v[i]; }
	const T& operator[](std::size_t i) const { return _v[i]; }
	T max_mag() const {
		T m = 0;
		for (const T& s : _v) { T a = std::fabs(s); if (a > m) m = a; }
		return m;
	}
};
}
