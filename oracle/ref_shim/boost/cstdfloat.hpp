// Interface shim, TEST INFRASTRUCTURE ONLY (see oracle/ref_filtercore.cpp).
// Stands in for <boost/cstdfloat.hpp>, which the reference includes
// (FilterCore.h:10, ProcessFile.h:9) and which is absent from this image.
// boost::float32_t / float64_t are the IEEE binary32 / binary64 types.
#pragma once
#include <cstdint>
namespace boost {
using float32_t = float;
using float64_t = double;
static_assert(sizeof(float32_t) == 4 && sizeof(float64_t) == 8);
}
