// stencil.cuh -- assignment-function stencils shared by deposit and gather.
// Restates PMMethod::TSCAssignmentFunc and the base-cell arithmetic of spreadMass / interpolateField
// (source/pmMethod.cpp:187-198, 207-209, 219-230, 250-258) with the reference's quirks kept:
//   Q1  the TSC/CIC base cell is the TRUNCATED coordinate, NGP uses round();
//   Q2  cells are addressed with the unwrapped flat index x + y*Nx + z*Nx*Ny.
#pragma once

#include "common.cuh"

namespace p3m {

template <int K>
struct StencilInfo;  // K points per axis
template <>
struct StencilInfo<1> { static constexpr int first = 0; };   // NGP
template <>
struct StencilInfo<2> { static constexpr int first = 0; };   // CIC: base, base+1
template <>
struct StencilInfo<3> { static constexpr int first = -1; };  // TSC: base-1 .. base+1

template <typename T, int K>
struct Stencil {
  int x0, y0, z0;  // first cell touched on each axis
  T wx[K], wy[K], wz[K];
  T pref;  // mass prefactor for deposition: m/8 (TSC, source/pmMethod.cpp:261) or m
};

template <typename T, int K>
__device__ __forceinline__ void axis_weights(T p, int& first, T w[K]) {
  if (K == 1) {
    first = (int)round(p);  // std::round: half away from zero (source/pmMethod.cpp:207)
    w[0] = T(1);
  } else if (K == 2) {
    int b = (int)p;
    T d = p - (T)b;
    first = b;
    w[0] = T(1) - d;  // tx = 1 - dx (source/pmMethod.cpp:228)
    w[K - 1] = d;
  } else {
    int b = (int)p;
    T d = p - (T)b;
    first = b - 1;
    w[0] = (T(0.5) - d) * (T(0.5) - d);      // t = -1
    w[K > 1 ? 1 : 0] = T(1.5) - 2 * d * d;   // t =  0
    w[K - 1] = (T(0.5) + d) * (T(0.5) + d);  // t = +1
  }
}

template <typename T, int K>
__device__ __forceinline__ Stencil<T, K> make_stencil(T x, T y, T z, T m) {
  Stencil<T, K> s;
  axis_weights<T, K>(x, s.x0, s.wx);
  axis_weights<T, K>(y, s.y0, s.wy);
  axis_weights<T, K>(z, s.z0, s.wz);
  s.pref = (K == 3) ? m / 8 : m;
  return s;
}

// integer division of small non-negative ints by a runtime divisor through a float reciprocal
// (exact for e < 2^20, d <= 64: the +0.5 keeps the quotient away from integer boundaries)
__device__ __forceinline__ int fast_div(int e, int d, float inv_d) {
  (void)d;
  return __float2int_rz(((float)e + 0.5f) * inv_d);
}

__device__ __forceinline__ int wrap_idx(int a, int n) {
  // include/grid.h:46 mod(): (a % n + n) % n, for |a| < 2n
  a = a < 0 ? a + n : a;
  a = a >= n ? a - n : a;
  if (a < 0 || a >= n) a = ((a % n) + n) % n;
  return a;
}

}  // namespace p3m
