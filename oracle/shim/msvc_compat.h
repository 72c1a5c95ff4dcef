// Build-side shim (test infrastructure, not product code), force-included with -include when the
// UNMODIFIED reference sources under /root/reference are compiled with g++.
// MSVC pulls these headers in transitively and exposes ::sinf & co. inside namespace std.
#pragma once
#include <math.h>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
namespace std {
using ::acosf;
using ::atanf;
using ::cosf;
using ::powf;
using ::sinf;
using ::sqrtf;
}  // namespace std
