// shortrange.cu -- A7 + A8 + A9: the chaining-mesh short-range particle-particle correction.
//
// Replaces P3MMethod::calculateShortRangeForces / updateSRForcesThreadJob / updateSRForces,
// shortRangeForce(FromTable), initSRForceTable and correctAccelerations
// (source/p3mMethod.cpp:50-57, 168-322) together with the ChainingMesh linked lists
// (source/chainingMesh.cpp:20-84).
//
// The reference walks a half shell of 13 + 1 neighbour cells per cell, applies Newton's third law into
// 13 per-particle neighbour slots (168 B of accumulators per particle) and divides by the mass at the
// end.  Because the chaining cell is never smaller than the cutoff (M = int(box / re)), the result is
// simply "sum over all j with |r_ij| < re"; here every TARGET gathers that sum itself from the full
// 27-cell neighbourhood (no scatter, no atomics, no slots, bit-reproducible run to run), and the
// acceleration mj * f(r) * r_ij is accumulated directly (the reference's  mi*mj*f / mi).
//
//   dense cells (>= kDenseCell particles; the Plummer core puts ~half of all particles in 8 cells):
//     work item = (cell, 64 consecutive targets) handled by ONE WARP, 2 targets per lane in registers.
//     Particles are sorted by (cell, 16^3 sub-cell, id), so consecutive particles are spatially compact;
//     every globally aligned 32-particle group carries a bounding box (k_tile_aabb) and a source group whose
//     box is farther than the cutoff from the box of the warp's targets is skipped WHOLE (exact: it only
//     drops pairs with r >= re).  Surviving groups are staged in the warp's private shared-memory slice
//     (next one prefetched in registers) and read back with broadcast LDS.128; there is no block-wide
//     barrier in the loop.  Items are generated on the device, sorted by cost (heaviest first) and pulled
//     from an atomic queue by a persistent grid.
//   sparse cells: one thread per target walks its 27 cells straight from L1/L2.
//
// Force law (code units, G = 1/(4 pi)): table mode reproduces shortRangeForceFromTable (:240-245): linear
// interpolation in r^2 over 500 entries, multiplied by r_ij (NOT the unit vector).  Coordinates
// are staged in units of the cutoff, so u = |d|^2 = r^2 / re^2 saturates to 1 exactly at the cutoff, where
// the last table entry is (0, 0): that folds the test r^2 < re^2 of :258 into the lookup.  Inner loop, per
// pair: 3 FADD (d), FMUL + FFMA + FFMA.SAT (u), FFMA.RZ with 2^23 (floor(499 u) lands in the mantissa), LEA,
// LDS.64 of (A_t, B_t) from the lane's conflict-free copy of the table, FFMA (f = A + B u), 3 FFMA
// (accumulate) = 13 issue slots with equal masses (+1 FMUL otherwise).  r = 0 contributes exactly 0
// (F_0 = 0), so i == j needs no test.  Bound: FP32 / issue rate, not memory.
#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <type_traits>
#include <vector>

#include "ctx.cuh"

namespace p3m {

template <typename T>
struct alignas(2 * sizeof(T)) V2 {
  T x, y;
};

template <typename T>
__device__ __forceinline__ T ref_force_dev(const SRParams<T>& sp, T r) {
  const T G = T(0.07957747154594767);  // 1 / (4 pi)
  const T a = sp.a;
  if (sp.cloud == P3M_S1) {  // referenceForceS1 :194-201
    if (r >= a) return G / (r * r);
    const T q = r / a;
    return G / (a * a) * (8 * r / a - 9 * r * r / (a * a) + 2 * q * q * q * q);
  }
  const T u = 2 * r / a;  // referenceForceS2 :203-218
  const T u2 = u * u, u3 = u2 * u, u4 = u2 * u2, u5 = u4 * u, u6 = u3 * u3;
  if (u <= 1) return G / (35 * a * a) * (224 * u - 224 * u3 + 70 * u4 + 48 * u5 - 21 * u6);
  if (u <= 2)
    return G / (35 * a * a) *
           (12 / u2 - 224 + 896 * u - 840 * u2 + 224 * u3 + 70 * u4 - 48 * u5 + 7 * u6);
  return G / (r * r);
}

// Handle of the shared-memory table of (A_t, B_t) pairs, replicated COPIES times so that the lanes of a
// warp read disjoint banks: entry t of copy c sits at element t * COPIES + c and lane l uses copy
// l % COPIES.  A 64-bit (fp32 pair) load is served one half warp at a time, a 128-bit (fp64 pair) load
// one quarter warp at a time, so 128 / sizeof(pair) copies make EVERY lookup conflict-free (ncu on the
// single-copy table: 41 % of all shared-memory wavefronts were bank conflicts of this lookup and the
// kernel sat at 0.86 wavefronts/clk/SM, i.e. it was shared-memory bound).
// The argument is u = r^2 / re^2 saturated to [0, 1]:  index = floor(499 u), value = A + B u.
// fp32: fma.rz(u, 499, 2^23) leaves the index in the low mantissa bits; the exponent bits (0x4B000000)
// are folded into an opaque pre-biased per-lane base, so the lookup is FFMA.RZ + LEA + LDS.64.
template <typename T, int COPIES>
struct TableRef;
template <int COPIES>
struct TableRef<float, COPIES> {
  static constexpr int kShift = 3 + (COPIES == 1 ? 0 : COPIES == 8 ? 3 : 4);
  static_assert(COPIES == 1 || COPIES == 8 || COPIES == 16, "copies");
  unsigned base;
  // `slots` are 32 shared-memory words: the biased base makes a round trip through them so that ptxas
  // cannot re-associate the bias back out of the address arithmetic (it would cost an IADD3 per pair).
  // Must be called by all threads of the block, before a __syncthreads().
  __device__ __forceinline__ TableRef(const V2<float>* tab, volatile unsigned* slots) {
    if (threadIdx.x < 32)
      slots[threadIdx.x] = (unsigned)__cvta_generic_to_shared(tab) + (threadIdx.x & (COPIES - 1)) * 8u -
                           (0x4B000000u << kShift);
    base = 0;
  }
  __device__ __forceinline__ void finish(volatile unsigned* slots) { base = slots[threadIdx.x & 31]; }
  // `bits` = fma.rz(u, 499, 2^23) reinterpreted: index in the low mantissa bits, exponent folded into `base`
  __device__ __forceinline__ V2<float> at(unsigned bits) const {
    const unsigned addr = base + (bits << kShift);
    V2<float> e;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(e.x), "=f"(e.y) : "r"(addr));
    return e;
  }
  __device__ __forceinline__ V2<float> get(float u) const {  // 0 <= u <= 1
    const unsigned addr =
        base + ((unsigned)__float_as_int(__fmaf_rz(u, float(kSRTable - 1), 8388608.0f)) << kShift);
    V2<float> e;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(e.x), "=f"(e.y) : "r"(addr));
    return e;
  }
};
template <int COPIES>
struct TableRef<double, COPIES> {
  const V2<double>* tab;
  __device__ __forceinline__ TableRef(const V2<double>* t, volatile unsigned*)
      : tab(t + (threadIdx.x & (COPIES - 1))) {}
  __device__ __forceinline__ void finish(volatile unsigned*) {}
  __device__ __forceinline__ V2<double> get(double u) const {
    return tab[(int)(u * double(kSRTable - 1)) * COPIES];
  }
};

template <typename T>
__device__ __forceinline__ T fma_sat(T a, T b, T c) {
  return fmin(fma(a, b, c), T(1));
}
template <>
__device__ __forceinline__ float fma_sat<float>(float a, float b, float c) {
  float r;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// One target-source pair.
//  TABLE: d = (target - source) / re, so u = |d|^2 = r^2 / re^2 saturates to 1 exactly at the cutoff,
//         where the last table entry is (0, 0): that folds the test r^2 < re^2 of
//         source/p3mMethod.cpp:258 into the lookup.  The table holds re * (A_t, 499 B_t) [* m when all
//         masses are equal], so  acc += f * d  needs no rescaling.  Per pair: 3 FADD, FMUL, FFMA,
//         FFMA.SAT, FFMA.RZ, LEA, LDS.64, FFMA, (FMUL mj), 3 FFMA = 13 (14) issue slots.
//  analytic: d in code units (shortRangeForce :220-238, divided by mi).
template <typename T, bool TABLE, bool COUNT, bool UNIMASS, int COPIES>
__device__ __forceinline__ void pair_acc(T dx, T dy, T dz, T mj, const SRParams<T>& sp,
                                         const TableRef<T, COPIES>& tab, T& ax, T& ay, T& az,
                                         unsigned& n_in) {
  if (TABLE) {
    const T u = fma_sat<T>(dz, dz, fma(dy, dy, dx * dx));
    if (COUNT) n_in += (u < T(1) && u > T(0)) ? 1u : 0u;
    const V2<T> e = tab.get(u);
    T f = fma(e.y, u, e.x);
    if (!UNIMASS) f *= mj;
    ax = fma(f, dx, ax), ay = fma(f, dy, ay), az = fma(f, dz, az);
  } else {
    const T r2 = dx * dx + dy * dy + dz * dz;
    if (COUNT) n_in += (r2 < sp.re2 && r2 > T(0)) ? 1u : 0u;
    if (r2 < sp.re2 && r2 > T(0)) {
      const T r = sqrt(r2);
      const T G = T(0.07957747154594767);
      const T f = mj * (ref_force_dev(sp, r) - G / (r2 + sp.eps2)) / r;
      ax += f * dx, ay += f * dy, az += f * dz;
    }
  }
}

// g_tab holds the reference's interpolation in slope-intercept form (A_t, B_t) over xi = r^2/delta^2;
// shared memory gets it over u = xi / 499 with the output scale folded in.
template <typename T, int COPIES>
__device__ __forceinline__ void load_table(const T* __restrict__ g_tab, V2<T>* s_tab, T scale) {
  for (int k = threadIdx.x; k < kSRTable * COPIES; k += blockDim.x) {
    const int t = k / COPIES;
    s_tab[k] = V2<T>{g_tab[2 * t] * scale, g_tab[2 * t + 1] * (scale * T(kSRTable - 1))};
  }
}

// ---- work items for the dense cells ------------------------------------------------------------------
template <typename T>
__global__ void k_pp_items(const int* __restrict__ cell_start, Geom<T> g, int dense_cell, int* __restrict__ items,
                           unsigned* __restrict__ cost, int* __restrict__ counters) {
  const long long ncells = 1LL << (3 * g.mbits);
  long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const int s = cell_start[c], nq = cell_start[c + 1] - s;
  if (nq < dense_cell) return;
  const int cx = (int)compact3((uint32_t)c), cy = (int)compact3((uint32_t)c >> 1),
            cz = (int)compact3((uint32_t)c >> 2);
  long long src = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = cx + dx, y = cy + dy, z = cz + dz;
        if (x < 0 || y < 0 || z < 0 || x >= g.mx || y >= g.my || z >= g.mz) continue;
        const uint32_t qn = morton3((uint32_t)x, (uint32_t)y, (uint32_t)z);
        src += cell_start[qn + 1] - cell_start[qn];
      }
  const int k = (nq + kPPTargets - 1) / kPPTargets;
  const int base = atomicAdd(&counters[0], k);
  for (int t = 0; t < k; ++t) {
    items[2 * (base + t)] = (int)c;
    items[2 * (base + t) + 1] = s + t * kPPTargets;
    const int nt = min(kPPTargets, nq - t * kPPTargets);
    long long w = (src * nt) >> 10;
    cost[base + t] = (unsigned)min(w, 0xffffffffLL);
  }
}

template <typename T>
struct PPCfg {
  static constexpr int kCopies = 128 / (2 * (int)sizeof(T));  // conflict-free table replication
#ifndef P3M_PP_WARPS
#define P3M_PP_WARPS 16
#endif
#ifndef P3M_PP_UNROLL
#define P3M_PP_UNROLL 8
#endif
  static constexpr int kWarps = P3M_PP_WARPS;                  // warps per CTA
  static constexpr int kUnroll = P3M_PP_UNROLL;                // sources per unrolled inner-loop body
  static constexpr int kCtasPerSm = sizeof(T) == 8 ? 1 : 2;
  static constexpr size_t smem(int sub) {
    return sizeof(T) * 2 * kSRTable * kCopies + sizeof(T) * 4 * (size_t)sub * kWarps + 128;
  }
};

// Dense cells.  Every WARP is autonomous: it pulls (cell, 64 targets) items from the atomic queue, holds
// 2 targets per lane in registers, tests 32 source boxes (SUB particles each) per ballot against the box
// of its own targets, stages each surviving box in its private shared-memory slice (next box prefetched
// into registers meanwhile) and runs the inner loop with broadcast LDS.128.  No block-wide barrier in
// the loop: the warps of a CTA only share the replicated force table.
// Staged coordinates are (p - origin) / re with origin = the centre of the target cell: |p - origin| <= 1.5
// cells, so the subtraction of two nearby fp32 numbers is exact or within half an ulp of the POSITION (the
// granularity the reference's own x_i - x_j works on) and the one rounding of the scaling is 2^-24 relative to
// a number <= 1.5 hc / re ~ 1.6 -- about 1e-7 of the cutoff on every staged coordinate.
template <typename T, bool TABLE, bool COUNT, bool UNIMASS, int SUB>
__global__ void __launch_bounds__(PPCfg<T>::kWarps * 32, PPCfg<T>::kCtasPerSm)
k_pp_tiled(const V4<T>* __restrict__ posm, const int* __restrict__ cell_start,
           const V4<T>* __restrict__ aabb, const V4<T>* __restrict__ gposm,
           const int* __restrict__ gcell_start, const V4<T>* __restrict__ gaabb,
           const int* __restrict__ items, const unsigned* __restrict__ order,
           int* __restrict__ counters, Geom<T> g, SRParams<T> sp, const T* __restrict__ g_tab,
           T uni_mass, V4<T>* __restrict__ acc, V4<T>* __restrict__ acc_sr,
           unsigned long long* __restrict__ pair_counts) {
  constexpr int COPIES = PPCfg<T>::kCopies;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V2<T>* s_tab = reinterpret_cast<V2<T>*>(smem_raw);
  V4<T>* s_src_all = reinterpret_cast<V4<T>*>(smem_raw + sizeof(V2<T>) * kSRTable * COPIES);
  volatile unsigned* s_tb =
      reinterpret_cast<volatile unsigned*>(smem_raw + sizeof(V2<T>) * kSRTable * COPIES +
                                           sizeof(V4<T>) * SUB * PPCfg<T>::kWarps);
  // acc = sum F(u) * (pi - pj) = re * sum F(u) * d
  const T re = sqrt(sp.re2);
  const T scl = TABLE ? T(1) / re : T(1);
  load_table<T, COPIES>(g_tab, s_tab, (TABLE ? re : T(1)) * (UNIMASS ? uni_mass : T(1)));
  TableRef<T, COPIES> tref(s_tab, s_tb);
  __syncthreads();
  tref.finish(s_tb);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  V4<T>* s_src = s_src_all + wid * SUB;
  const int nitems = counters[0];
  const T cut2 = sp.re2 * (T(1) + T(1e-5));
  unsigned long long checked = 0, inrange = 0;
  for (;;) {
    int it = 0;
    if (lane == 0) it = atomicAdd(&counters[1], 1);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= nitems) break;
    const int item = (int)order[it];
    const uint32_t q = (uint32_t)items[2 * item];
    const int t0 = items[2 * item + 1];
    const int tend = min(cell_start[q + 1], t0 + kPPTargets);
    const int i0 = t0 + lane, i1 = i0 + 32;
    const bool v0 = i0 < tend, v1 = i1 < tend;
    V4<T> p0 = posm[v0 ? i0 : t0], p1 = posm[v1 ? i1 : t0];
    // bounding box of this warp's targets (code units)
    T wlo[3], whi[3];
    {
      T lx = min(p0.x, p1.x), ly = min(p0.y, p1.y), lz = min(p0.z, p1.z);
      T hx = max(p0.x, p1.x), hy = max(p0.y, p1.y), hz = max(p0.z, p1.z);
      for (int o = 16; o > 0; o >>= 1) {
        lx = min(lx, __shfl_xor_sync(0xffffffffu, lx, o)), ly = min(ly, __shfl_xor_sync(0xffffffffu, ly, o));
        lz = min(lz, __shfl_xor_sync(0xffffffffu, lz, o)), hx = max(hx, __shfl_xor_sync(0xffffffffu, hx, o));
        hy = max(hy, __shfl_xor_sync(0xffffffffu, hy, o)), hz = max(hz, __shfl_xor_sync(0xffffffffu, hz, o));
      }
      wlo[0] = lx, wlo[1] = ly, wlo[2] = lz, whi[0] = hx, whi[1] = hy, whi[2] = hz;
    }
    const int cx = (int)compact3(q), cy = (int)compact3(q >> 1), cz = (int)compact3(q >> 2);
    // staging frame of this item
    const T ox = TABLE ? (T(cx) + T(0.5)) * g.hcx : T(0);
    const T oy = TABLE ? (T(cy) + T(0.5)) * g.hcy : T(0);
    const T oz = TABLE ? (T(cz) + T(0.5)) * g.hcz : T(0);
    p0.x = (p0.x - ox) * scl, p0.y = (p0.y - oy) * scl, p0.z = (p0.z - oz) * scl;
    p1.x = (p1.x - ox) * scl, p1.y = (p1.y - oy) * scl, p1.z = (p1.z - oz) * scl;
    auto stage = [&](V4<T> v) { return V4<T>{(v.x - ox) * scl, (v.y - oy) * scl, (v.z - oz) * scl, v.w}; };
    T a0x = 0, a0y = 0, a0z = 0, a1x = 0, a1y = 0, a1z = 0;
    unsigned n_in = 0;
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = cx + dx, y = cy + dy, z = cz + dz;
          if (x < 0 || y < 0 || z < 0 || x >= g.mx || y >= g.my || z >= g.mz) continue;
          const uint32_t qn = morton3((uint32_t)x, (uint32_t)y, (uint32_t)z);
          // multi-GPU: cells of a foreign z-layer are served from the ghost copy of the neighbour slab
          const bool own = g.nranks == 1 || layer_owner(g, z) == g.rank;
          const V4<T>* __restrict__ spos = own ? posm : gposm;
          const int* __restrict__ scs = own ? cell_start : gcell_start;
          const V4<T>* __restrict__ sbb = own ? aabb : gaabb;
          const int s = scs[qn], e = scs[qn + 1];
          if (s >= e) continue;
          const int box_first = s / SUB, box_last = (e - 1) / SUB;
          for (int b0 = box_first; b0 <= box_last; b0 += 32) {
            // exact culling, 32 boxes per ballot: a box farther than the cutoff from the targets' box
            // holds no partner of any of them
            const int mybox = b0 + lane;
            bool near_ = false;
            if (mybox <= box_last) {
              const V4<T> blo = sbb[2 * mybox], bhi = sbb[2 * mybox + 1];
              const T gx = max(T(0), max(blo.x - whi[0], wlo[0] - bhi.x));
              const T gy = max(T(0), max(blo.y - whi[1], wlo[1] - bhi.y));
              const T gz = max(T(0), max(blo.z - whi[2], wlo[2] - bhi.z));
              near_ = gx * gx + gy * gy + gz * gz <= cut2;
            }
            unsigned todo = __ballot_sync(0xffffffffu, near_);
            if (todo == 0u) continue;
            // software pipeline over the surviving boxes: registers hold the NEXT box while the
            // current one is consumed from shared memory
            const V4<T> far_{T(-64), T(-64), T(-64), T(0)};  // padding: beyond the cutoff, mass 0
            int k = __ffs(todo) - 1;
            todo &= todo - 1;
            int jb = max(s, (b0 + k) * SUB), je = min(e, (b0 + k + 1) * SUB);
            V4<T> n0 = (jb + lane < je) ? stage(spos[jb + lane]) : far_;
            V4<T> n1 = far_;
            if (SUB > 32) n1 = (jb + 32 + lane < je) ? stage(spos[jb + 32 + lane]) : far_;
            for (;;) {
              const int cnt = je - jb;
              __syncwarp();
              s_src[lane] = n0;
              if (SUB > 32) s_src[lane + 32] = n1;
              __syncwarp();
              const bool more = todo != 0u;
              if (more) {
                k = __ffs(todo) - 1;
                todo &= todo - 1;
                jb = max(s, (b0 + k) * SUB), je = min(e, (b0 + k + 1) * SUB);
                n0 = (jb + lane < je) ? stage(spos[jb + lane]) : far_;
                if (SUB > 32) n1 = (jb + 32 + lane < je) ? stage(spos[jb + 32 + lane]) : far_;
              }
              if (COUNT) checked += (unsigned long long)cnt * ((v0 ? 1 : 0) + (v1 ? 1 : 0));
#pragma unroll PPCfg<T>::kUnroll
              for (int j = 0; j < cnt; ++j) {
                const V4<T> sj = s_src[j];
                unsigned c0 = 0, c1 = 0;
                pair_acc<T, TABLE, COUNT, UNIMASS, COPIES>(p0.x - sj.x, p0.y - sj.y, p0.z - sj.z, sj.w, sp,
                                                           tref, a0x, a0y, a0z, c0);
                pair_acc<T, TABLE, COUNT, UNIMASS, COPIES>(p1.x - sj.x, p1.y - sj.y, p1.z - sj.z, sj.w, sp,
                                                           tref, a1x, a1y, a1z, c1);
                if (COUNT) n_in += (v0 ? c0 : 0u) + (v1 ? c1 : 0u);
              }
              if (!more) break;
            }
          }
        }
    if (!COUNT && v0) {  // the counting instantiation is read-only: p3m_get_pair_counts has no side effects
      acc_sr[i0] = V4<T>{a0x, a0y, a0z, 0};
      V4<T> a = acc[i0];
      acc[i0] = V4<T>{a.x + a0x, a.y + a0y, a.z + a0z, 0};  // correctAccelerations :55
    }
    if (!COUNT && v1) {
      acc_sr[i1] = V4<T>{a1x, a1y, a1z, 0};
      V4<T> a = acc[i1];
      acc[i1] = V4<T>{a.x + a1x, a.y + a1y, a.z + a1z, 0};
    }
    if (COUNT) inrange += n_in;
  }
  if (COUNT) {
    atomicAdd(&pair_counts[0], checked);
    atomicAdd(&pair_counts[1], inrange);
  }
}

// ---- packed-FP32 dense-cell kernel (fp32, tabulated force, equal masses: the reference's default set-up) ------
// Same work decomposition as k_pp_tiled (autonomous warps, 64-target items, exact box culling, private staging
// slice, register prefetch), but the pair body runs on sm_100's two-wide FP32 instructions
// (add / mul / fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2): one lane processes 2 targets x 2 sources per
// iteration, the two SOURCES of a pair sharing each packed instruction.  The scalar body costs 13 issue slots
// per pair and the kernel was issue-bound (ncu r01: issue-active 86 %, FMA pipe 64 %); packed, a pair costs
//   3 FADD2 + FMUL2 + FFMA2 + 2 FFMA.SAT (f32x2 has no .sat) + FFMA2.RZ + 2 LEA + 2 LDS.64 + 2 FFMA + 3 FFMA2
//   = 17 slots per TWO pairs, plus 2 broadcast LDS per FOUR pairs = 9 slots per pair,
// which moves the bound from the issue port to the FMA pipe itself (11 lane-cycles per pair either way).
// Sources are staged NEGATED and interleaved pairwise, (-xa,-xb,-ya,-yb) | (-za,-zb), so that d = p + (-s) is
// one FADD2 per axis with the target coordinate duplicated in a register pair.  Arithmetic per pair is the
// scalar kernel's, operation for operation (same roundings); only the summation order differs (even and odd
// sources of a box are summed separately and added at the end), which stays fixed run to run.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 fma2_rz(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// measured on B200 (C2, ms per short-range phase): 16 warps x 2 CTAs / unroll 4 (64 registers): 111.1;
// 16 x 2 / 8: 110.0; 12 x 2 / 4 (80 registers): 104.9; 12 x 2 / 8: 102.8; 8 x 2 / 8 (126 registers): 101.8 --
// instruction-level parallelism inside a warp pays more than resident warps.
#ifndef P3M_PK_WARPS
#define P3M_PK_WARPS 8
#endif
#ifndef P3M_PK_UNROLL
#define P3M_PK_UNROLL 8
#endif
#ifndef P3M_PK_CTAS
#define P3M_PK_CTAS 2
#endif
struct PackedCfg {
  static constexpr int kCopies = 16;
  static constexpr int kWarps = P3M_PK_WARPS;
  static constexpr int kUnroll = P3M_PK_UNROLL;      // source PAIRS per unrolled inner-loop body
  static constexpr int kCtasPerSm = P3M_PK_CTAS;
  static constexpr int kSub = 32;                    // sources per box (== kPPSub)
  static constexpr int kSliceBytes = kSub * 12;      // per warp: 16 x (xa,xb,ya,yb) + 16 x (za,zb)
  static constexpr size_t smem() { return sizeof(float) * 2 * kSRTable * kCopies + (size_t)kSliceBytes * kWarps + 128; }
};

// one target against one packed source pair; (ax, ay, az) hold the two partial sums side by side
__device__ __forceinline__ void pair2_acc(f32x2 px, f32x2 py, f32x2 pz, f32x2 nsx, f32x2 nsy, f32x2 nsz,
                                          const TableRef<float, PackedCfg::kCopies>& tab, f32x2 k499, f32x2 kbias,
                                          f32x2& ax, f32x2& ay, f32x2& az) {
  const f32x2 dx = add2(px, nsx), dy = add2(py, nsy), dz = add2(pz, nsz);
  const f32x2 t = fma2(dy, dy, mul2(dx, dx));
  float ta, tb, dza, dzb;
  unpack2(t, ta, tb);
  unpack2(dz, dza, dzb);
  const float ua = fma_sat<float>(dza, dza, ta), ub = fma_sat<float>(dzb, dzb, tb);
  const f32x2 u = pack2(ua, ub);
  const f32x2 idx = fma2_rz(u, k499, kbias);  // floor(499 u) in the low mantissa bits of each half
  float ia, ib;
  unpack2(idx, ia, ib);
  const V2<float> ea = tab.at((unsigned)__float_as_int(ia)), eb = tab.at((unsigned)__float_as_int(ib));
  const float fa = fmaf(ea.y, ua, ea.x), fb = fmaf(eb.y, ub, eb.x);
  const f32x2 f = pack2(fa, fb);
  ax = fma2(f, dx, ax), ay = fma2(f, dy, ay), az = fma2(f, dz, az);
}

__global__ void __launch_bounds__(PackedCfg::kWarps * 32, PackedCfg::kCtasPerSm)
k_pp_packed(const V4<float>* __restrict__ posm, const int* __restrict__ cell_start,
            const V4<float>* __restrict__ aabb, const V4<float>* __restrict__ gposm,
            const int* __restrict__ gcell_start, const V4<float>* __restrict__ gaabb,
            const int* __restrict__ items, const unsigned* __restrict__ order, int* __restrict__ counters,
            Geom<float> g, SRParams<float> sp, const float* __restrict__ g_tab, float uni_mass,
            V4<float>* __restrict__ acc, V4<float>* __restrict__ acc_sr) {
  using T = float;
  constexpr int COPIES = PackedCfg::kCopies, SUB = PackedCfg::kSub;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V2<T>* s_tab = reinterpret_cast<V2<T>*>(smem_raw);
  unsigned char* s_slices = smem_raw + sizeof(V2<T>) * kSRTable * COPIES;
  volatile unsigned* s_tb = reinterpret_cast<volatile unsigned*>(s_slices + PackedCfg::kSliceBytes * PackedCfg::kWarps);
  const T re = sqrt(sp.re2);
  const T scl = T(1) / re;
  load_table<T, COPIES>(g_tab, s_tab, re * uni_mass);
  TableRef<T, COPIES> tref(s_tab, s_tb);
  __syncthreads();
  tref.finish(s_tb);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float4* s_xy = reinterpret_cast<float4*>(s_slices + wid * PackedCfg::kSliceBytes);       // 16 x (-xa,-xb,-ya,-yb)
  float2* s_z = reinterpret_cast<float2*>(s_slices + wid * PackedCfg::kSliceBytes + 256);  // 16 x (-za,-zb)
  float* s_xy_w = reinterpret_cast<float*>(s_xy) + (lane >> 1) * 4 + (lane & 1);
  float* s_z_w = reinterpret_cast<float*>(s_z) + lane;
  const int nitems = counters[0];
  const T cut2 = sp.re2 * (T(1) + T(1e-5));
  const f32x2 k499 = pack2(float(kSRTable - 1), float(kSRTable - 1)), kbias = pack2(8388608.0f, 8388608.0f);
  for (;;) {
    int it = 0;
    if (lane == 0) it = atomicAdd(&counters[1], 1);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= nitems) break;
    const int item = (int)order[it];
    const uint32_t q = (uint32_t)items[2 * item];
    const int t0 = items[2 * item + 1];
    const int tend = min(cell_start[q + 1], t0 + kPPTargets);
    const int i0 = t0 + lane, i1 = i0 + 32;
    const bool v0 = i0 < tend, v1 = i1 < tend;
    V4<T> p0 = posm[v0 ? i0 : t0], p1 = posm[v1 ? i1 : t0];
    T wlo[3], whi[3];
    {
      T lx = min(p0.x, p1.x), ly = min(p0.y, p1.y), lz = min(p0.z, p1.z);
      T hx = max(p0.x, p1.x), hy = max(p0.y, p1.y), hz = max(p0.z, p1.z);
      for (int o = 16; o > 0; o >>= 1) {
        lx = min(lx, __shfl_xor_sync(0xffffffffu, lx, o)), ly = min(ly, __shfl_xor_sync(0xffffffffu, ly, o));
        lz = min(lz, __shfl_xor_sync(0xffffffffu, lz, o)), hx = max(hx, __shfl_xor_sync(0xffffffffu, hx, o));
        hy = max(hy, __shfl_xor_sync(0xffffffffu, hy, o)), hz = max(hz, __shfl_xor_sync(0xffffffffu, hz, o));
      }
      wlo[0] = lx, wlo[1] = ly, wlo[2] = lz, whi[0] = hx, whi[1] = hy, whi[2] = hz;
    }
    const int cx = (int)compact3(q), cy = (int)compact3(q >> 1), cz = (int)compact3(q >> 2);
    // staging frame of this item (see k_pp_tiled): centre of the target cell
    const T ox = (T(cx) + T(0.5)) * g.hcx;
    const T oy = (T(cy) + T(0.5)) * g.hcy;
    const T oz = (T(cz) + T(0.5)) * g.hcz;
    const f32x2 p0x = pack2((p0.x - ox) * scl, (p0.x - ox) * scl), p0y = pack2((p0.y - oy) * scl, (p0.y - oy) * scl),
                p0z = pack2((p0.z - oz) * scl, (p0.z - oz) * scl);
    const f32x2 p1x = pack2((p1.x - ox) * scl, (p1.x - ox) * scl), p1y = pack2((p1.y - oy) * scl, (p1.y - oy) * scl),
                p1z = pack2((p1.z - oz) * scl, (p1.z - oz) * scl);
    // negated staged source; padding lanes sit beyond the cutoff of every target (u saturates, F_499 = 0)
    auto stage_neg = [&](V4<T> v) { return V4<T>{-((v.x - ox) * scl), -((v.y - oy) * scl), -((v.z - oz) * scl), 0}; };
    const V4<T> far_{T(64), T(64), T(64), T(0)};
    f32x2 a0x = 0, a0y = 0, a0z = 0, a1x = 0, a1y = 0, a1z = 0;
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = cx + dx, y = cy + dy, z = cz + dz;
          if (x < 0 || y < 0 || z < 0 || x >= g.mx || y >= g.my || z >= g.mz) continue;
          const uint32_t qn = morton3((uint32_t)x, (uint32_t)y, (uint32_t)z);
          const bool own = g.nranks == 1 || layer_owner(g, z) == g.rank;
          const V4<T>* __restrict__ spos = own ? posm : gposm;
          const int* __restrict__ scs = own ? cell_start : gcell_start;
          const V4<T>* __restrict__ sbb = own ? aabb : gaabb;
          const int s = scs[qn], e = scs[qn + 1];
          if (s >= e) continue;
          const int box_first = s / SUB, box_last = (e - 1) / SUB;
          for (int b0 = box_first; b0 <= box_last; b0 += 32) {
            const int mybox = b0 + lane;
            bool near_ = false;
            if (mybox <= box_last) {
              const V4<T> blo = sbb[2 * mybox], bhi = sbb[2 * mybox + 1];
              const T gx = max(T(0), max(blo.x - whi[0], wlo[0] - bhi.x));
              const T gy = max(T(0), max(blo.y - whi[1], wlo[1] - bhi.y));
              const T gz = max(T(0), max(blo.z - whi[2], wlo[2] - bhi.z));
              near_ = gx * gx + gy * gy + gz * gz <= cut2;
            }
            unsigned todo = __ballot_sync(0xffffffffu, near_);
            if (todo == 0u) continue;
            int k = __ffs(todo) - 1;
            todo &= todo - 1;
            int jb = max(s, (b0 + k) * SUB), je = min(e, (b0 + k + 1) * SUB);
            V4<T> n0 = (jb + lane < je) ? stage_neg(spos[jb + lane]) : far_;
            for (;;) {
              const int npair = (je - jb + 1) >> 1;
              __syncwarp();
              s_xy_w[0] = n0.x, s_xy_w[2] = n0.y, *s_z_w = n0.z;
              __syncwarp();
              const bool more = todo != 0u;
              if (more) {
                k = __ffs(todo) - 1;
                todo &= todo - 1;
                jb = max(s, (b0 + k) * SUB), je = min(e, (b0 + k + 1) * SUB);
                n0 = (jb + lane < je) ? stage_neg(spos[jb + lane]) : far_;
              }
#pragma unroll PackedCfg::kUnroll
              for (int j = 0; j < npair; ++j) {
                const float4 sxy = s_xy[j];
                const float2 sz = s_z[j];
                const f32x2 nsx = pack2(sxy.x, sxy.y), nsy = pack2(sxy.z, sxy.w), nsz = pack2(sz.x, sz.y);
                pair2_acc(p0x, p0y, p0z, nsx, nsy, nsz, tref, k499, kbias, a0x, a0y, a0z);
                pair2_acc(p1x, p1y, p1z, nsx, nsy, nsz, tref, k499, kbias, a1x, a1y, a1z);
              }
              if (!more) break;
            }
          }
        }
    // the two half sums are the sums over the even and the odd sources of every staged box
    float lo, hi;
    if (v0) {
      V4<T> r;
      unpack2(a0x, lo, hi), r.x = lo + hi;
      unpack2(a0y, lo, hi), r.y = lo + hi;
      unpack2(a0z, lo, hi), r.z = lo + hi;
      r.w = 0;
      acc_sr[i0] = r;
      const V4<T> a = acc[i0];
      acc[i0] = V4<T>{a.x + r.x, a.y + r.y, a.z + r.z, 0};  // correctAccelerations :55
    }
    if (v1) {
      V4<T> r;
      unpack2(a1x, lo, hi), r.x = lo + hi;
      unpack2(a1y, lo, hi), r.y = lo + hi;
      unpack2(a1z, lo, hi), r.z = lo + hi;
      r.w = 0;
      acc_sr[i1] = r;
      const V4<T> a = acc[i1];
      acc[i1] = V4<T>{a.x + r.x, a.y + r.y, a.z + r.z, 0};
    }
  }
}

// Sparse cells: one thread per target walks its 27 cells straight from L1/L2 (single-copy table).
template <typename T, bool TABLE, bool COUNT>
__global__ void __launch_bounds__(128)
k_pp_sparse(const V4<T>* __restrict__ posm, long long n, const int* __restrict__ cell_start,
            const V4<T>* __restrict__ gposm, const int* __restrict__ gcell_start, Geom<T> g, int dense_cell, SRParams<T> sp, const T* __restrict__ g_tab, V4<T>* __restrict__ acc,
            V4<T>* __restrict__ acc_sr, unsigned long long* __restrict__ pair_counts) {
  __shared__ V2<T> s_tab[kSRTable];
  __shared__ unsigned s_tb[32];
  const T re = sqrt(sp.re2);
  const T scl = TABLE ? T(1) / re : T(1);
  load_table<T, 1>(g_tab, s_tab, TABLE ? re : T(1));
  TableRef<T, 1> tref(s_tab, s_tb);
  __syncthreads();
  tref.finish(s_tb);
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  V4<T> p = posm[i];
  int cx, cy, cz;
  bool inside;
  bin_cell(g, p.x, p.y, p.z, cx, cy, cz, inside);
  const uint32_t q = morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
  if (cell_start[q + 1] - cell_start[q] >= dense_cell) return;  // tiled kernel owns this cell
  T ax = 0, ay = 0, az = 0;
  unsigned n_in = 0;
  unsigned long long checked = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = cx + dx, y = cy + dy, z = cz + dz;
        if (x < 0 || y < 0 || z < 0 || x >= g.mx || y >= g.my || z >= g.mz) continue;
        const uint32_t qn = morton3((uint32_t)x, (uint32_t)y, (uint32_t)z);
        const bool own = g.nranks == 1 || layer_owner(g, z) == g.rank;
        const V4<T>* __restrict__ spos = own ? posm : gposm;
        const int* __restrict__ scs = own ? cell_start : gcell_start;
        const int s = scs[qn], e = scs[qn + 1];
        if (COUNT) checked += (unsigned long long)(e - s);
        for (int j = s; j < e; ++j) {
          const V4<T> sj = spos[j];
          // differences of code-unit coordinates are exact for close pairs (as in the reference)
          pair_acc<T, TABLE, COUNT, false, 1>((p.x - sj.x) * scl, (p.y - sj.y) * scl, (p.z - sj.z) * scl, sj.w,
                                              sp, tref, ax, ay, az, n_in);
        }
      }
  if (!COUNT) {
    acc_sr[i] = V4<T>{ax, ay, az, 0};
    const V4<T> a = acc[i];
    acc[i] = V4<T>{a.x + ax, a.y + ay, a.z + az, 0};
  }
  if (COUNT) {
    atomicAdd(&pair_counts[0], checked);
    atomicAdd(&pair_counts[1], (unsigned long long)n_in);
  }
}

// ---- host side ------------------------------------------------------------------------------------------
// initSRForceTable (source/p3mMethod.cpp:275-294) in the context's precision, same operation order
template <typename T>
static void build_table_host(const p3m_ctx* c, const SRParams<T>& sp, T delta2, T eps, std::vector<T>& F) {
  F.resize(kSRTable);
  const T G = 1 / (4 * (T)3.14159265358979323846);
  auto refS1 = [&](T r) -> T {
    const T a = sp.a;
    if (r >= a) return G / (r * r);
    return G / (a * a) * (8 * r / a - 9 * r * r / (a * a) + 2 * std::pow(r / a, (T)4));
  };
  auto refS2 = [&](T r) -> T {
    const T a = sp.a;
    const T u = 2 * r / a;
    if (u <= 1)
      return G / (35 * std::pow(a, (T)2)) *
             (224 * u - 224 * std::pow(u, (T)3) + 70 * std::pow(u, (T)4) + 48 * std::pow(u, (T)5) -
              21 * std::pow(u, (T)6));
    if (u <= 2)
      return G / (35 * std::pow(a, (T)2)) *
             (12 / std::pow(u, (T)2) - 224 + 896 * u - 840 * std::pow(u, (T)2) +
              224 * std::pow(u, (T)3) + 70 * std::pow(u, (T)4) - 48 * std::pow(u, (T)5) +
              7 * std::pow(u, (T)6));
    return G / (r * r);
  };
  for (int i = 0; i < kSRTable; ++i) {
    const T r2 = i * delta2;
    const T r = std::sqrt(r2);
    const T R = -(c->prm.cloud_shape == P3M_S1 ? refS1(r) : refS2(r));
    const T total = -G / (r * r + eps * eps);
    F[i] = (r == 0) ? 0 : (total - R) / std::sqrt(r * r + eps * eps);
  }
}

template <typename T>
int sr_table_upload(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  SRParams<T>& sp = Sel<T>::sr(c);
  const p3m_params& p = c->prm;
  // code-unit lengths, source/p3mMethod.cpp:35-37,42-44
  const T re = (T)p.cutoff_radius / (T)p.H;
  sp.a = (T)(p.sr_particle_diameter > 0 ? p.sr_particle_diameter : p.particle_diameter) / (T)p.H;
  const T eps = (T)p.softening / (T)p.H;
  const T delta2 = re * re / (kSRTable - 1);
  sp.re2 = re * re;
  sp.inv_delta2 = 1 / delta2;
  sp.eps2 = eps * eps;
  sp.use_table = p.use_sr_table;
  sp.cloud = p.cloud_shape;
  std::vector<T> F;
  build_table_host<T>(c, sp, delta2, eps, F);
  c->sr_table_host.assign(F.begin(), F.end());
  // device layout: slope-intercept form of the reference's interpolation, evaluated in double:
  //   F_t + (xi - t)(F_{t+1} - F_t)  =  A_t + B_t xi,   A_t = F_t - t B_t,  B_t = F_{t+1} - F_t
  // entry 499 = (0, 0): everything at or beyond the cutoff evaluates to zero.
  std::vector<T> pairs(2 * kSRTable);
  for (int t = 0; t < kSRTable - 1; ++t) {
    const double b = (double)F[t + 1] - (double)F[t];
    pairs[2 * t] = (T)((double)F[t] - t * b);
    pairs[2 * t + 1] = (T)b;
  }
  pairs[2 * (kSRTable - 1)] = 0, pairs[2 * (kSRTable - 1) + 1] = 0;
  P3M_CUDA(cudaMemcpyAsync(s.sr_table, pairs.data(), sizeof(T) * 2 * kSRTable, cudaMemcpyHostToDevice,
                           c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

template <typename T>
__global__ void k_iota_zero(unsigned* idx, unsigned* cost, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (unsigned)i, cost[i] = 0u;
}

template <typename T, bool TABLE, bool COUNT, bool UNIMASS>
static int run_pp(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const SRParams<T>& sp = Sel<T>::sr(c);
  const long long n = c->n;
  const long long ncells = 1LL << (3 * g.mbits);
  // upper bound on dense-cell work items: every dense cell holds >= kDenseCell particles
  const int dense = c->tune.dense_cell;
  long long max_items = n / kPPTargets + (ncells < n / dense ? ncells : n / dense) + 16;
  unsigned* cost = reinterpret_cast<unsigned*>(s.keys);  // sort scratch is free between bin_sort calls
  unsigned* cost_sorted = reinterpret_cast<unsigned*>(s.keys_alt);
  unsigned* idx = s.slots;
  unsigned* order = s.slots_alt;
  if (max_items > c->cap) max_items = c->cap;
  P3M_CUDA(cudaMemsetAsync(s.pp_counters, 0, sizeof(int) * 8, c->stream));
  if (COUNT) P3M_CUDA(cudaMemsetAsync(s.pair_counts, 0, sizeof(unsigned long long) * 2, c->stream));
  k_iota_zero<T><<<(unsigned)((max_items + 255) / 256), 256, 0, c->stream>>>(idx, cost, max_items);
  P3M_LAUNCH_CHECK(c);
  k_pp_items<T><<<(unsigned)((ncells + 255) / 256), 256, 0, c->stream>>>(s.cell_start, g, dense, s.pp_items,
                                                                         cost, s.pp_counters);
  P3M_LAUNCH_CHECK(c);
  size_t tmp = s.cub_tmp_bytes;
  P3M_CUDA(cub::DeviceRadixSort::SortPairsDescending(s.cub_tmp, tmp, cost, cost_sorted, idx, order,
                                                     (int)max_items, 0, 32, c->stream));
  c->launches += 5;
  bool packed = false;
  if constexpr (std::is_same<T, float>::value && TABLE && UNIMASS && !COUNT && kPPSub == PackedCfg::kSub) {
    packed = !c->tune.scalar_pp;
    if (packed) {
      const size_t smem = PackedCfg::smem();
      P3M_CUDA(cudaFuncSetAttribute(k_pp_packed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_pp_packed<<<c->num_sms * PackedCfg::kCtasPerSm, PackedCfg::kWarps * 32, smem, c->stream>>>(
          s.posm, s.cell_start, s.aabb, s.gposm, s.gcell_start, s.gaabb, s.pp_items, order, s.pp_counters, g, sp,
          s.sr_table, (float)c->uniform_mass_code, s.acc, s.acc_sr);
      P3M_LAUNCH_CHECK(c);
    }
  }
  if (!COUNT) c->packed_pp = packed;
  if (!packed) {
    auto kern = k_pp_tiled<T, TABLE, COUNT, UNIMASS, kPPSub>;
    const size_t smem = PPCfg<T>::smem(kPPSub);
    P3M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<c->num_sms * PPCfg<T>::kCtasPerSm, PPCfg<T>::kWarps * 32, smem, c->stream>>>(
        s.posm, s.cell_start, s.aabb, s.gposm, s.gcell_start, s.gaabb, s.pp_items, order, s.pp_counters, g, sp,
        s.sr_table, (T)c->uniform_mass_code, s.acc, s.acc_sr, s.pair_counts);
    P3M_LAUNCH_CHECK(c);
  }
  k_pp_sparse<T, TABLE, COUNT><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(
      s.posm, n, s.cell_start, s.gposm, s.gcell_start, g, dense, sp, s.sr_table, s.acc, s.acc_sr, s.pair_counts);
  P3M_LAUNCH_CHECK(c);
  return 0;
}

template <typename T>
int short_range(p3m_ctx* c) {
  if (!c->prm.p3m) return 0;
  if (!c->have_particles || !c->sorted) return fail(P3M_ESTATE, "p3m_short_range: particles not sorted");
  if (c->n == 0) return 0;
  if (!c->have_acc && !c->count_pairs) {
    // no mesh part for the current particle order (p3m_gather was not run since the last sort): the
    // short-range part is then the whole acceleration
    P3M_CUDA(cudaMemsetAsync(Sel<T>::st(c).acc, 0, sizeof(V4<T>) * (size_t)c->n, c->stream));
    c->have_acc = true;
  }
  phase_begin(c, PH_SHORT_RANGE);
  int r;
  const bool table = c->prm.use_sr_table != 0, count = c->count_pairs != 0;
  // all masses equal (the reference's samplers: masses = M / n): the mass is folded into the table
  const bool uni = table && c->uniform_mass;
  if (uni && !count) r = run_pp<T, true, false, true>(c);
  else if (uni && count) r = run_pp<T, true, true, true>(c);
  else if (table && !count) r = run_pp<T, true, false, false>(c);
  else if (table && count) r = run_pp<T, true, true, false>(c);
  else if (!table && !count) r = run_pp<T, false, false, false>(c);
  else r = run_pp<T, false, true, false>(c);
  phase_end(c, PH_SHORT_RANGE);
  return r;
}

template int short_range<float>(p3m_ctx*);
template int short_range<double>(p3m_ctx*);
template int sr_table_upload<float>(p3m_ctx*);
template int sr_table_upload<double>(p3m_ctx*);

}  // namespace p3m
