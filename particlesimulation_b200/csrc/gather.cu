// gather.cu -- A5 + A6: PMMethod::findFieldInCells / getFieldInCell and
// PMMethod::updateAccelerations / interpolateField (source/pmMethod.cpp:279-338, 352-390), fused.
//
// The reference materialises a Vec3 field mesh (12 B/cell written, then re-read at random by every
// particle).  Here a CTA walks a run of cell-sorted particles; for each aligned block of binning cells
// it stages the POTENTIAL tile (+ finite-difference halo) in shared memory, differences it once into an
// E tile, and the block's particles interpolate from shared memory.  HBM traffic drops from
// 16 B/cell + 12 B/cell + 24 B/particle to ~4 B/cell + 28 B/particle; the field mesh is never written.
// p3m_gradient() still produces the explicit field mesh for callers of Grid::getField.
//
// Quirks kept: truncated base cell, field cells addressed by the UNWRAPPED flat index
// (only the potential wraps periodically, source/grid.cpp:46-48,66-68).  Particles whose stencil leaves
// [0,N) on some axis -- where the unwrapped index aliases into a neighbouring row -- take the exact
// per-cell path below instead of the tile.
#include <algorithm>

#include "ctx.cuh"
#include "stencil.cuh"

namespace p3m {

// -E = grad(phi) by centred differences with periodic wrap (source/pmMethod.cpp:352-371)
template <typename T, int FD>
__device__ __forceinline__ void field_at(const T* __restrict__ phi, const Geom<T>& g, int x, int y,
                                         int z, T& fx, T& fy, T& fz) {
  const long long sx = 1, sy = g.nx, sz = (long long)g.nx * g.ny;
  auto P = [&](int a, int b, int cc) -> T {
    const int zl = pot_plane(g, cc);  // multi-GPU: `phi` holds the planes of this rank's particle slab
    return zl < 0 ? T(0) : phi[wrap_idx(a, g.nx) * sx + wrap_idx(b, g.ny) * sy + zl * sz];
  };
  if (FD == 1) {
    fx = T(-0.5) * (P(x + 1, y, z) - P(x - 1, y, z));
    fy = T(-0.5) * (P(x, y + 1, z) - P(x, y - 1, z));
    fz = T(-0.5) * (P(x, y, z + 1) - P(x, y, z - 1));
  } else {
    const T k = T(-1.0) / 12;
    fx = k * (-P(x + 2, y, z) + 8 * P(x + 1, y, z) - 8 * P(x - 1, y, z) + P(x - 2, y, z));
    fy = k * (-P(x, y + 2, z) + 8 * P(x, y + 1, z) - 8 * P(x, y - 1, z) + P(x, y - 2, z));
    fz = k * (-P(x, y, z + 2) + 8 * P(x, y, z + 1) - 8 * P(x, y, z - 1) + P(x, y, z - 2));
  }
}

// exact restatement for one particle straight from global memory (any position)
template <typename T, int K, int FD>
__device__ __forceinline__ void gather_direct(const Stencil<T, K>& s, const Geom<T>& g,
                                              const T* __restrict__ phi, T& ax, T& ay, T& az) {
  const T scale = (K == 3) ? T(0.125) : T(1);
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = 0; b < K; ++b)
#pragma unroll
      for (int cc = 0; cc < K; ++cc) {
        const T w = scale * ((s.wx[a] * s.wy[b]) * s.wz[cc]);
        long long flat = (long long)(s.x0 + a) + (long long)(s.y0 + b) * g.nx +
                         (long long)(s.z0 + cc) * g.nx * g.ny;
        if (flat < 0 || flat >= g.M) continue;  // out of the arrays: undefined in the reference
        const int x = (int)(flat % g.nx), y = (int)((flat / g.nx) % g.ny),
                  z = (int)(flat / ((long long)g.nx * g.ny));
        T fx, fy, fz;
        field_at<T, FD>(phi, g, x, y, z, fx, fy, fz);
        ax += w * fx, ay += w * fy, az += w * fz;
      }
}

// Interior particles that are not served from a shared-memory tile (sparse blocks): the same arithmetic as the
// tile path, straight from the potential mesh in global memory / L2 with run-time strides -- one base pointer, no
// per-point index arithmetic.  Returns false when the stencil + finite-difference halo touches the periodic wrap,
// leaves the arrays or (several GPUs) the planes held locally: the caller then takes gather_direct.
template <typename T, int K, int FD>
__device__ __forceinline__ bool gather_interior(const Stencil<T, K>& s, const Geom<T>& g,
                                                const T* __restrict__ phi, T& ax, T& ay, T& az) {
  if (s.x0 < FD || s.y0 < FD || s.z0 < FD || s.x0 + K + FD > g.nx || s.y0 + K + FD > g.ny || s.z0 + K + FD > g.nz)
    return false;
  const int zl0 = pot_plane(g, s.z0 - FD), zl1 = pot_plane(g, s.z0 + K - 1 + FD);
  if (zl0 < 0 || zl1 != zl0 + K - 1 + 2 * FD) return false;
  const long long SY = g.nx, SZ = (long long)g.nx * g.ny;
  const T* base = phi + s.x0 + (long long)s.y0 * SY + (long long)(zl0 + FD) * SZ;
  const T scale = (K == 3) ? T(0.125) : T(1);
#pragma unroll
  for (int a = 0; a < K; ++a)
#pragma unroll
    for (int b = 0; b < K; ++b)
#pragma unroll
      for (int cc = 0; cc < K; ++cc) {
        const T* q = base + a + b * SY + cc * SZ;
        if (FD == 1) {
          const T wgt = (s.wx[a] * s.wy[b]) * s.wz[cc];
          ax += wgt * (q[1] - q[-1]), ay += wgt * (q[SY] - q[-SY]), az += wgt * (q[SZ] - q[-SZ]);
        } else {
          const T wgt = scale * ((s.wx[a] * s.wy[b]) * s.wz[cc]);
          const T k = T(-1.0) / 12;
          ax += wgt * (k * (-q[2] + 8 * q[1] - 8 * q[-1] + q[-2]));
          ay += wgt * (k * (-q[2 * SY] + 8 * q[SY] - 8 * q[-SY] + q[-2 * SY]));
          az += wgt * (k * (-q[2 * SZ] + 8 * q[SZ] - 8 * q[-SZ] + q[-2 * SZ]));
        }
      }
  if (FD == 1) {
    const T f = T(-0.5) * scale;
    ax *= f, ay *= f, az *= f;
  }
  return true;
}

// externalField in original units -> code units (source/pmMethod.cpp:386-388,
// source/externalFields.cpp:4-15, include/unitConversions.h:22-24)
template <typename T>
__device__ __forceinline__ void add_external(const Geom<T>& g, T x, T y, T z, T& ax, T& ay, T& az) {
  if (g.ext_kind != P3M_EXT_SPH_RAD_DECR) return;
  const T dx = g.H * x - g.ecx, dy = g.H * y - g.ecy, dz = g.H * z - g.ecz;
  const T r = sqrt(dx * dx + dy * dy + dz * dz);
  T gg;
  if (r > g.eR)
    gg = -g.G * g.eM / (r * r);
  else
    gg = -(g.G * g.eM / (g.eR * g.eR * g.eR)) * r * (4 - 3 * r / g.eR);
  const T ex = gg * (dx / r), ey = gg * (dy / r), ez = gg * (dz / r);
  ax += g.DT * g.DT * ex / g.H, ay += g.DT * g.DT * ey / g.H, az += g.DT * g.DT * ez / g.H;
}

template <typename T, int K, int FD>
__global__ void __launch_bounds__(256, 2)
k_gather(const V4<T>* __restrict__ posm, long long n, int chunk, const int* __restrict__ cell_start,
         Geom<T> g, const T* __restrict__ phi, V4<T>* __restrict__ acc) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int tile_cap = g.tex * g.tey * g.tez;
  // E tile as (Ex, Ey, Ez, -) records: one vector LDS per stencil point; potential tile behind it
  V4<T>* sE = reinterpret_cast<V4<T>*>(smem_raw);
  T* sphi = reinterpret_cast<T*>(sE + tile_cap);
  long long cur = (long long)blockIdx.x * chunk;
  const long long chunk_end = min(n, cur + (long long)chunk);
  const long long ncells = 1LL << (3 * g.mbits);
  const int bs3 = 3 * g.bshift;
  const T scale = (K == 3) ? T(0.125) : T(1);

  while (cur < chunk_end) {  // CTA-uniform loop over tile segments
    const V4<T> p0 = posm[cur];
    int cx, cy, cz;
    bool inside;
    bin_cell(g, p0.x, p0.y, p0.z, cx, cy, cz, inside);
    const uint32_t blk = morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz) >> bs3;
    long long blk_end_cell = ((long long)blk + 1) << bs3;
    if (blk_end_cell > ncells) blk_end_cell = ncells;
    long long seg_end = min(chunk_end, (long long)cell_start[blk_end_cell]);
    if (seg_end <= cur) seg_end = cur + 1;
    const int count = (int)(seg_end - cur);

    if (count < 2 * (long long)g.tile_min) {
      for (long long i = cur + tid; i < seg_end; i += blockDim.x) {
        const V4<T> p = posm[i];
        const Stencil<T, K> s = make_stencil<T, K>(p.x, p.y, p.z, p.w);
        T ax = 0, ay = 0, az = 0;
        gather_direct<T, K, FD>(s, g, phi, ax, ay, az);
        add_external(g, p.x, p.y, p.z, ax, ay, az);
        acc[i] = V4<T>{ax, ay, az, 0};
      }
      cur = seg_end;
      continue;
    }

    int lo[3], ext[3];
    tile_box(g, (int)compact3(blk), (int)compact3(blk >> 1), (int)compact3(blk >> 2), lo, ext);
    const int px = ext[0] + 2 * FD, py = ext[1] + 2 * FD, pz = ext[2] + 2 * FD;
    const int pelems = px * py * pz, elems = ext[0] * ext[1] * ext[2];
    {
      const float ipx = 1.0f / (float)px, ipy = 1.0f / (float)py;
      for (int e = tid; e < pelems; e += blockDim.x) {
        const int q = fast_div(e, px, ipx);
        const int ix = e - q * px;
        const int iz = fast_div(q, py, ipy);
        const int iy = q - iz * py;
        const int gx = wrap_idx(lo[0] - FD + ix, g.nx), gy = wrap_idx(lo[1] - FD + iy, g.ny),
                  gz = pot_plane(g, lo[2] - FD + iz);
        sphi[e] = gz < 0 ? T(0) : phi[(long long)gx + (long long)gy * g.nx + (long long)gz * g.nx * g.ny];
      }
    }
    __syncthreads();
    {
      const float i0 = 1.0f / (float)ext[0], i1 = 1.0f / (float)ext[1];
      const int sy = px, sz = px * py;
      for (int e = tid; e < elems; e += blockDim.x) {
        const int q = fast_div(e, ext[0], i0);
        const int ix = e - q * ext[0];
        const int iz = fast_div(q, ext[1], i1);
        const int iy = q - iz * ext[1];
        const T* c0 = sphi + (iz + FD) * sz + (iy + FD) * sy + (ix + FD);
        T fx, fy, fz;
        if (FD == 1) {
          fx = T(-0.5) * (c0[1] - c0[-1]);
          fy = T(-0.5) * (c0[sy] - c0[-sy]);
          fz = T(-0.5) * (c0[sz] - c0[-sz]);
        } else {
          const T k = T(-1.0) / 12;
          fx = k * (-c0[2] + 8 * c0[1] - 8 * c0[-1] + c0[-2]);
          fy = k * (-c0[2 * sy] + 8 * c0[sy] - 8 * c0[-sy] + c0[-2 * sy]);
          fz = k * (-c0[2 * sz] + 8 * c0[sz] - 8 * c0[-sz] + c0[-2 * sz]);
        }
        sE[e] = V4<T>{fx, fy, fz, 0};
      }
    }
    __syncthreads();

    for (long long i = cur + tid; i < seg_end; i += blockDim.x) {
      const V4<T> p = posm[i];
      const Stencil<T, K> s = make_stencil<T, K>(p.x, p.y, p.z, p.w);
      const int rx = s.x0 - lo[0], ry = s.y0 - lo[1], rz = s.z0 - lo[2];
      const bool fits = rx >= 0 && ry >= 0 && rz >= 0 && rx + K <= ext[0] && ry + K <= ext[1] &&
                        rz + K <= ext[2] && s.x0 >= 0 && s.y0 >= 0 && s.z0 >= 0 &&
                        s.x0 + K <= g.nx && s.y0 + K <= g.ny && s.z0 + K <= g.nz;
      T ax = 0, ay = 0, az = 0;
      if (fits) {
        const int base = (rz * ext[1] + ry) * ext[0] + rx;
#pragma unroll
        for (int a = 0; a < K; ++a)
#pragma unroll
          for (int b = 0; b < K; ++b)
#pragma unroll
            for (int cc = 0; cc < K; ++cc) {
              const T w = scale * ((s.wx[a] * s.wy[b]) * s.wz[cc]);
              const V4<T> ev = sE[base + (cc * ext[1] + b) * ext[0] + a];
              ax += w * ev.x, ay += w * ev.y, az += w * ev.z;
            }
      } else {
        gather_direct<T, K, FD>(s, g, phi, ax, ay, az);
      }
      add_external(g, p.x, p.y, p.z, ax, ay, az);
      acc[i] = V4<T>{ax, ay, az, 0};
    }
    __syncthreads();
    cur = seg_end;
  }
}

// ---- PM-only contexts: compile-time tile, potential tile only -------------------------------------------
// The binning cells are 8^3 mesh cells and particles are sorted by (Morton(binning cell), mesh cell,
// id), so 4 consecutive binning cells = one 16 x 16 x 8 block of mesh cells = one contiguous run of
// particles.  A CTA stages the POTENTIAL of that block (+ assignment and finite-difference halo, periodic
// wrap applied while staging) and every particle differences and interpolates straight from it: no E
// tile, one barrier per block, every shared-memory offset a compile-time immediate.  The arithmetic per
// stencil point is the reference's (E = -1/2 (phi[+1] - phi[-1]), then sum w E); the power-of-two
// factors -1/2 and 1/8 are applied once at the end, which is exact in binary floating point.
template <int FD>
struct PmTile {
  static constexpr int TX = 16, TY = 16, TZ = 8;
  static constexpr int HALO = 1 + FD;  // assignment stencil reach + finite-difference reach
  // row pitch 24 = 8 mod 16: the 4 rows of 8 cells that 32 consecutive sorted particles cover (sub key =
  // mesh cell, x fastest, inside an 8^3 binning cell) fall into 4 disjoint groups of 8 banks
  static constexpr int PX = 24, PY = TY + 2 * HALO, PZ = TZ + 2 * HALO;
  static constexpr int WX = TX + 2 * HALO;  // staged extent in x (<= PX)
  static constexpr int ELEMS = PX * PY * PZ;
};
constexpr int kGatherPmDirect = 24;  // blocks with fewer particles read the potential straight from L2

// non-empty 16 x 16 x 8 blocks (4 consecutive binning cells), compacted
__global__ void k_occupied_blocks(const int* __restrict__ cell_start, long long ncells, int* __restrict__ list,
                                  int* __restrict__ counter) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (4 * b >= ncells) return;
  const long long c1 = 4 * b + 4 < ncells ? 4 * b + 4 : ncells;
  if (cell_start[c1] > cell_start[4 * b]) list[atomicAdd(counter, 1)] = (int)b;
}

// 4-byte asynchronous copy global -> shared (LDGSTS): no register staging, completes in the background
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Persistent CTAs walk the compacted block list with a static stride.  While a block's particles are
// interpolated from one shared-memory buffer, the potential tile of the CTA's NEXT block streams into the
// other buffer with asynchronous copies (periodic wrap applied per element), and the list / cell_start
// entries of the block after that are already in flight: no global-memory latency sits on the critical path.
template <typename T, int K, int FD>
__global__ void __launch_bounds__(256, 3)
k_gather_pm(const V4<T>* __restrict__ posm, const int* __restrict__ cell_start, const int* __restrict__ list,
            const int* __restrict__ counter, Geom<T> g, const T* __restrict__ phi, V4<T>* __restrict__ acc) {
  using TL = PmTile<FD>;
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* const sbuf0 = reinterpret_cast<T*>(smem_raw);
  const int tid = threadIdx.x;
  const long long ncells = 1LL << (3 * g.mbits);
  const int nocc = *counter, G = gridDim.x;
  const T scale = (K == 3) ? T(0.125) : T(1);
  constexpr int SY = TL::PX, SZ = TL::PX * TL::PY;

  auto origin = [&](int blk, int& ox, int& oy, int& oz) {
    const uint32_t c0 = 4u * (uint32_t)blk;
    ox = ((int)compact3(c0) << kPmTileShift) - TL::HALO;
    oy = ((int)compact3(c0 >> 1) << kPmTileShift) - TL::HALO;
    oz = ((int)compact3(c0 >> 2) << kPmTileShift) - TL::HALO;
  };
  auto stage_async = [&](int blk, T* dst) {
    int ox, oy, oz;
    origin(blk, ox, oy, oz);
    for (int el = tid; el < TL::ELEMS; el += 256) {
      const int iz = el / (TL::PX * TL::PY), r = el - iz * (TL::PX * TL::PY);
      const int iy = r / TL::PX, ix = r - iy * TL::PX;
      if (ix >= TL::WX) continue;  // pitch padding
      const int gx = wrap_idx(ox + ix, g.nx), gy = wrap_idx(oy + iy, g.ny), gz = pot_plane(g, oz + iz);
      if (gz < 0) {
        dst[el] = T(0);
      } else {
        const T* src = phi + ((long long)gx + (long long)gy * g.nx + (long long)gz * g.nx * g.ny);
        if (sizeof(T) == 4) {
          cp_async_4(dst + el, src);
        } else {
          cp_async_4(reinterpret_cast<float*>(dst + el), reinterpret_cast<const float*>(src));
          cp_async_4(reinterpret_cast<float*>(dst + el) + 1, reinterpret_cast<const float*>(src) + 1);
        }
      }
    }
  };
  auto range_of = [&](int blk, int& s, int& e) {
    const long long c0 = 4LL * blk;
    s = cell_start[c0], e = cell_start[c0 + 4 < ncells ? c0 + 4 : ncells];
  };

  int w = blockIdx.x;
  if (w >= nocc) return;
  int blk = list[w];
  int blk_n = w + G < nocc ? list[w + G] : -1;
  int s, e;
  range_of(blk, s, e);
  stage_async(blk, sbuf0);
  cp_async_commit();
  for (int it = 0; w < nocc; w += G, ++it) {
    T* sphi = sbuf0 + (it & 1) * TL::ELEMS;
    // next block: its tile starts streaming now; the block after that is looked up for the next iteration
    int s_n = 0, e_n = 0;
    const int blk_nn = w + 2 * G < nocc ? list[w + 2 * G] : -1;
    if (blk_n >= 0) {
      range_of(blk_n, s_n, e_n);
      stage_async(blk_n, sbuf0 + ((it & 1) ^ 1) * TL::ELEMS);
    }
    cp_async_commit();
    int i = s + tid;
    V4<T> pn = i < e ? posm[i] : V4<T>{0, 0, 0, 0};
    cp_async_wait<1>();  // everything but the newest group: this block's tile has landed
    __syncthreads();
    int ox, oy, oz;
    origin(blk, ox, oy, oz);
    const bool direct = e - s < kGatherPmDirect;  // too few particles to be worth a tile: straight from L2
    for (; i < e; i += 256) {
      const V4<T> p = pn;
      if (i + 256 < e) pn = posm[i + 256];  // register prefetch of the next particle
      const Stencil<T, K> st = make_stencil<T, K>(p.x, p.y, p.z, p.w);
      const int rx = st.x0 - ox, ry = st.y0 - oy, rz = st.z0 - oz;
      const bool fits = !direct && rx >= FD && ry >= FD && rz >= FD && rx + K + FD <= TL::WX &&
                        ry + K + FD <= TL::PY && rz + K + FD <= TL::PZ && st.x0 >= 0 && st.y0 >= 0 && st.z0 >= 0 &&
                        st.x0 + K <= g.nx && st.y0 + K <= g.ny && st.z0 + K <= g.nz;
      T ax = 0, ay = 0, az = 0;
      if (fits) {
        const T* base = sphi + (rz * TL::PY + ry) * TL::PX + rx;
#pragma unroll
        for (int a = 0; a < K; ++a)
#pragma unroll
          for (int b = 0; b < K; ++b)
#pragma unroll
            for (int cc = 0; cc < K; ++cc) {
              const T* q = base + (cc * TL::PY + b) * TL::PX + a;
              if (FD == 1) {
                const T wgt = (st.wx[a] * st.wy[b]) * st.wz[cc];
                ax += wgt * (q[1] - q[-1]), ay += wgt * (q[SY] - q[-SY]), az += wgt * (q[SZ] - q[-SZ]);
              } else {
                const T wgt = scale * ((st.wx[a] * st.wy[b]) * st.wz[cc]);
                const T k = T(-1.0) / 12;
                ax += wgt * (k * (-q[2] + 8 * q[1] - 8 * q[-1] + q[-2]));
                ay += wgt * (k * (-q[2 * SY] + 8 * q[SY] - 8 * q[-SY] + q[-2 * SY]));
                az += wgt * (k * (-q[2 * SZ] + 8 * q[SZ] - 8 * q[-SZ] + q[-2 * SZ]));
              }
            }
        if (FD == 1) {
          const T f = T(-0.5) * scale;
          ax *= f, ay *= f, az *= f;
        }
      } else if (!gather_interior<T, K, FD>(st, g, phi, ax, ay, az)) {
        ax = ay = az = 0;
        gather_direct<T, K, FD>(st, g, phi, ax, ay, az);
      }
      add_external(g, p.x, p.y, p.z, ax, ay, az);
      acc[i] = V4<T>{ax, ay, az, 0};
    }
    __syncthreads();  // this buffer is overwritten by the copies issued in the next iteration
    blk = blk_n, blk_n = blk_nn, s = s_n, e = e_n;
  }
  cp_async_wait<0>();
}

// ---- P3M contexts: the same pipeline on blocks of chaining cells ---------------------------------------------
// Particles are sorted by (Morton(chaining cell), sub-cell, id), so an aligned block of (1 << gshift)^3 chaining
// cells is one contiguous run.  Its mesh footprint is not a fixed box (the chaining cell is box / M / H mesh
// cells, ~2.1 with the demo parameters, not an integer), so origin and extent are evaluated per block
// (tile_box) and the tile has a compile-time PITCH of kGatherTile in x and y: every stencil offset is still an
// immediate.  Persistent CTAs, potential tile of the next block streaming in with cp.async while the current
// block's particles are interpolated (fp32: two buffers; fp64: one, staged synchronously), finite differences
// taken on the fly from the potential tile (no E tile), one barrier per block.  Replaces the round-1 k_gather
// (synchronous staging, E tile, two barriers per block, 12 KB tiles sized for the deposit kernel).
constexpr int kGatherP3mChunk = 4096;  // particles per work item: a heavy block (cluster core) is shared by many CTAs

__global__ void k_occupied_p3m_blocks(const int* __restrict__ cell_start, long long ncells, int shift3,
                                      int2* __restrict__ list, int* __restrict__ counter) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long c0 = b << shift3;
  if (c0 >= ncells) return;
  const long long c1 = c0 + (1LL << shift3) < ncells ? c0 + (1LL << shift3) : ncells;
  const int np = cell_start[c1] - cell_start[c0];
  if (np <= 0) return;
  const int chunks = (np + kGatherP3mChunk - 1) / kGatherP3mChunk;
  const int at = atomicAdd(counter, chunks);
  for (int k = 0; k < chunks; ++k) list[at + k] = make_int2((int)b, k);
}

template <typename T, int K, int FD, int NBUF>
__global__ void __launch_bounds__(256, 2)
k_gather_p3m(const V4<T>* __restrict__ posm, const int* __restrict__ cell_start, const int2* __restrict__ list,
             const int* __restrict__ counter, Geom<T> g, const T* __restrict__ phi, V4<T>* __restrict__ acc) {
  constexpr int P = kGatherTile, ELEMS = P * P * P;
  constexpr int SY = P, SZ = P * P;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* const sbuf0 = reinterpret_cast<T*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long ncells = 1LL << (3 * g.mbits);
  const int shift3 = 3 * g.gshift;
  const int nocc = *counter, G = gridDim.x;
  const T scale = (K == 3) ? T(0.125) : T(1);
  Geom<T> gb = g;
  gb.bshift = g.gshift;

  // tile origin (mesh cell of tile element 0) and staged extents of block `blk`
  auto box_of = [&](int blk, int o[3], int w[3]) {
    const uint32_t c0 = (uint32_t)blk << shift3;
    int lo[3], ext[3];
    tile_box(gb, (int)(compact3(c0) >> g.gshift), (int)(compact3(c0 >> 1) >> g.gshift),
             (int)(compact3(c0 >> 2) >> g.gshift), lo, ext);
    for (int d = 0; d < 3; ++d) o[d] = lo[d] - FD, w[d] = min(ext[d] + 2 * FD, P);
  };
  // one warp per tile row: the index arithmetic is per ROW (warp-uniform), the lanes copy consecutive cells
  auto stage = [&](int blk, T* dst, bool async) {
    int o[3], w[3];
    box_of(blk, o, w);
    const int rows = w[1] * w[2];
    for (int r = wid; r < rows; r += 8) {
      const int iz = r / w[1], iy = r - iz * w[1];
      const int gy = wrap_idx(o[1] + iy, g.ny), gz = pot_plane(g, o[2] + iz);
      T* drow = dst + iz * SZ + iy * SY;
      if (lane < w[0]) {
        if (gz < 0) {
          drow[lane] = T(0);
        } else {
          const T* src = phi + ((long long)wrap_idx(o[0] + lane, g.nx) + (long long)gy * g.nx + (long long)gz * g.nx * g.ny);
          if (!async) {
            drow[lane] = *src;
          } else if (sizeof(T) == 4) {
            cp_async_4(drow + lane, src);
          } else {
            cp_async_4(reinterpret_cast<float*>(drow + lane), reinterpret_cast<const float*>(src));
            cp_async_4(reinterpret_cast<float*>(drow + lane) + 1, reinterpret_cast<const float*>(src) + 1);
          }
        }
      }
    }
  };
  auto range_of = [&](int2 item, int& s, int& e) {
    const long long c0 = (long long)item.x << shift3;
    s = cell_start[c0] + item.y * kGatherP3mChunk;
    e = min(s + kGatherP3mChunk, cell_start[c0 + (1LL << shift3) < ncells ? c0 + (1LL << shift3) : ncells]);
  };

  int w = blockIdx.x;
  if (w >= nocc) return;
  const int2 none = make_int2(-1, 0);
  int2 item = list[w];
  int2 item_n = w + G < nocc ? list[w + G] : none;
  int blk = item.x, blk_n = item_n.x;
  int s, e;
  range_of(item, s, e);
  if (NBUF == 2) {
    stage(blk, sbuf0, true);
    cp_async_commit();
  }
  for (int it = 0; w < nocc; w += G, ++it) {
    T* sphi = sbuf0 + (NBUF == 2 ? (it & 1) * ELEMS : 0);
    int s_n = 0, e_n = 0;
    const int2 item_nn = w + 2 * G < nocc ? list[w + 2 * G] : none;
    if (blk_n >= 0) range_of(item_n, s_n, e_n);
    if (NBUF == 2) {
      if (blk_n >= 0) stage(blk_n, sbuf0 + ((it & 1) ^ 1) * ELEMS, true);
      cp_async_commit();
    } else {
      stage(blk, sphi, false);
    }
    int i = s + tid;
    V4<T> pn = i < e ? posm[i] : V4<T>{0, 0, 0, 0};
    if (NBUF == 2) cp_async_wait<1>();  // everything but the newest group: this block's tile has landed
    __syncthreads();
    int o[3], wv[3];
    box_of(blk, o, wv);
    const bool direct = e - s < kGatherPmDirect;
    for (; i < e; i += 256) {
      const V4<T> p = pn;
      if (i + 256 < e) pn = posm[i + 256];
      const Stencil<T, K> st = make_stencil<T, K>(p.x, p.y, p.z, p.w);
      const int rx = st.x0 - o[0], ry = st.y0 - o[1], rz = st.z0 - o[2];
      const bool fits = !direct && rx >= FD && ry >= FD && rz >= FD && rx + K + FD <= wv[0] && ry + K + FD <= wv[1] &&
                        rz + K + FD <= wv[2] && st.x0 >= 0 && st.y0 >= 0 && st.z0 >= 0 && st.x0 + K <= g.nx &&
                        st.y0 + K <= g.ny && st.z0 + K <= g.nz;
      T ax = 0, ay = 0, az = 0;
      if (fits) {
        const T* base = sphi + (rz * P + ry) * P + rx;
#pragma unroll
        for (int a = 0; a < K; ++a)
#pragma unroll
          for (int b = 0; b < K; ++b)
#pragma unroll
            for (int cc = 0; cc < K; ++cc) {
              const T* q = base + (cc * P + b) * P + a;
              if (FD == 1) {
                const T wgt = (st.wx[a] * st.wy[b]) * st.wz[cc];
                ax += wgt * (q[1] - q[-1]), ay += wgt * (q[SY] - q[-SY]), az += wgt * (q[SZ] - q[-SZ]);
              } else {
                const T wgt = scale * ((st.wx[a] * st.wy[b]) * st.wz[cc]);
                const T k = T(-1.0) / 12;
                ax += wgt * (k * (-q[2] + 8 * q[1] - 8 * q[-1] + q[-2]));
                ay += wgt * (k * (-q[2 * SY] + 8 * q[SY] - 8 * q[-SY] + q[-2 * SY]));
                az += wgt * (k * (-q[2 * SZ] + 8 * q[SZ] - 8 * q[-SZ] + q[-2 * SZ]));
              }
            }
        if (FD == 1) {
          const T f = T(-0.5) * scale;
          ax *= f, ay *= f, az *= f;
        }
      } else if (!gather_interior<T, K, FD>(st, g, phi, ax, ay, az)) {
        ax = ay = az = 0;
        gather_direct<T, K, FD>(st, g, phi, ax, ay, az);
      }
      add_external(g, p.x, p.y, p.z, ax, ay, az);
      acc[i] = V4<T>{ax, ay, az, 0};
    }
    __syncthreads();  // this buffer is overwritten next
    item_n = item_nn;
    blk = blk_n, blk_n = item_nn.x, s = s_n, e = e_n;
  }
  if (NBUF == 2) cp_async_wait<0>();
}

template <typename T, int K, int FD>
static int launch_gather_p3m(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const long long ncells = 1LL << (3 * g.mbits);
  const int shift3 = 3 * g.gshift;
  const long long nblocks = (ncells + (1LL << shift3) - 1) >> shift3;
  // work items (block, chunk of its particles) in pp_items, which the short-range kernels rebuild later for
  // themselves: at most nblocks + n / kGatherP3mChunk pairs of ints, within its 2 * (cap / 64 + ncells + 16) ints
  int2* list = reinterpret_cast<int2*>(s.pp_items);
  int* counter = s.pp_counters + 7;
  P3M_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), c->stream));
  k_occupied_p3m_blocks<<<(unsigned)((nblocks + 255) / 256), 256, 0, c->stream>>>(s.cell_start, ncells, shift3, list, counter);
  P3M_LAUNCH_CHECK(c);
  constexpr int NBUF = sizeof(T) == 4 ? 2 : 1;
  auto kern = k_gather_p3m<T, K, FD, NBUF>;
  const size_t smem = (size_t)NBUF * sizeof(T) * kGatherTile * kGatherTile * kGatherTile;
  P3M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<c->num_sms * 2, 256, smem, c->stream>>>(s.posm, s.cell_start, list, counter, g, s.pot_part, s.acc);
  P3M_LAUNCH_CHECK(c);
  return 0;
}

template <typename T, int K, int FD>
static int launch_gather_pm(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const long long ncells = 1LL << (3 * g.mbits);
  const long long nblocks = (ncells + 3) / 4;
  int* list = s.pp_items + ncells;  // free in PM-only contexts (first ncells entries: deposit's cell list)
  int* counter = s.pp_counters + 7;
  P3M_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), c->stream));
  k_occupied_blocks<<<(unsigned)((nblocks + 255) / 256), 256, 0, c->stream>>>(s.cell_start, ncells, list, counter);
  P3M_LAUNCH_CHECK(c);
  auto kern = k_gather_pm<T, K, FD>;
  const size_t smem = 2 * sizeof(T) * PmTile<FD>::ELEMS;
  P3M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<c->num_sms * 3, 256, smem, c->stream>>>(s.posm, s.cell_start, list, counter, g, s.pot_part, s.acc);
  P3M_LAUNCH_CHECK(c);
  return 0;
}

template <typename T, int K, int FD>
static int launch_gather(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const long long n = c->n;
  long long want = (n + (long long)c->num_sms * 8 - 1) / ((long long)c->num_sms * 8);
  int chunk = (int)((want + 255) / 256 * 256);
  if (chunk < 256) chunk = 256;
  if (chunk > kGatherChunk) chunk = kGatherChunk;
  const long long blocks = (n + chunk - 1) / chunk;
  const size_t smem = sizeof(T) * ((size_t)(g.tex + 2 * FD) * (g.tey + 2 * FD) * (g.tez + 2 * FD) +
                                   4 * (size_t)g.tex * g.tey * g.tez);
  auto kern = k_gather<T, K, FD>;
  P3M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)blocks, 256, smem, c->stream>>>(s.posm, n, chunk, s.cell_start, g, s.pot_part,
                                                   s.acc);
  P3M_LAUNCH_CHECK(c);
  return 0;
}

template <typename T>
int gather(p3m_ctx* c) {
  if (!c->have_particles || !c->sorted) return fail(P3M_ESTATE, "p3m_gather: particles not sorted");
  if (!c->have_potential) return fail(P3M_ESTATE, "p3m_gather: no potential (call p3m_poisson)");
  const Geom<T>& g = Sel<T>::g(c);
  phase_begin(c, PH_GATHER);
  int r = 0;
  const bool pm_tiles = !g.p3m && g.tile_shift == kPmTileShift && g.sbits == g.tile_shift;
  if (c->n > 0 && pm_tiles) {
    const int k = g.is == P3M_TSC ? 3 : (g.is == P3M_CIC ? 2 : 1);
    const int fd = g.fds == P3M_TWO_POINT ? 1 : 2;
    if (k == 3 && fd == 1) r = launch_gather_pm<T, 3, 1>(c);
    else if (k == 3 && fd == 2) r = launch_gather_pm<T, 3, 2>(c);
    else if (k == 2 && fd == 1) r = launch_gather_pm<T, 2, 1>(c);
    else if (k == 2 && fd == 2) r = launch_gather_pm<T, 2, 2>(c);
    else if (k == 1 && fd == 1) r = launch_gather_pm<T, 1, 1>(c);
    else r = launch_gather_pm<T, 1, 2>(c);
  } else if (c->n > 0 && g.p3m && g.gshift >= 0 && !c->tune.old_gather) {
    const int k = g.is == P3M_TSC ? 3 : (g.is == P3M_CIC ? 2 : 1);
    const int fd = g.fds == P3M_TWO_POINT ? 1 : 2;
    if (k == 3 && fd == 1) r = launch_gather_p3m<T, 3, 1>(c);
    else if (k == 3 && fd == 2) r = launch_gather_p3m<T, 3, 2>(c);
    else if (k == 2 && fd == 1) r = launch_gather_p3m<T, 2, 1>(c);
    else if (k == 2 && fd == 2) r = launch_gather_p3m<T, 2, 2>(c);
    else if (k == 1 && fd == 1) r = launch_gather_p3m<T, 1, 1>(c);
    else r = launch_gather_p3m<T, 1, 2>(c);
  } else if (c->n > 0) {
    const int k = g.is == P3M_TSC ? 3 : (g.is == P3M_CIC ? 2 : 1);
    const int fd = g.fds == P3M_TWO_POINT ? 1 : 2;
    if (k == 3 && fd == 1) r = launch_gather<T, 3, 1>(c);
    else if (k == 3 && fd == 2) r = launch_gather<T, 3, 2>(c);
    else if (k == 2 && fd == 1) r = launch_gather<T, 2, 1>(c);
    else if (k == 2 && fd == 2) r = launch_gather<T, 2, 2>(c);
    else if (k == 1 && fd == 1) r = launch_gather<T, 1, 1>(c);
    else r = launch_gather<T, 1, 2>(c);
  }
  phase_end(c, PH_GATHER);
  if (r == 0) c->have_acc = true;
  return r;
}

// ---- explicit field mesh (PMMethod::findFieldInCells, source/pmMethod.cpp:373-382) -----------------
template <typename T, int FD>
__global__ void k_gradient(const T* __restrict__ phi, Geom<T> g, T* __restrict__ field) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= g.M) return;
  const int x = (int)(idx % g.nx), y = (int)((idx / g.nx) % g.ny),
            z = (int)(idx / ((long long)g.nx * g.ny));
  T fx, fy, fz;
  field_at<T, FD>(phi, g, x, y, z, fx, fy, fz);
  field[3 * idx] = fx, field[3 * idx + 1] = fy, field[3 * idx + 2] = fz;
}

template <typename T>
int gradient(p3m_ctx* c) {
  if (!c->have_potential) return fail(P3M_ESTATE, "p3m_gradient: no potential (call p3m_poisson)");
  if (c->slab) return fail(P3M_ESTATE, "p3m_gradient: the explicit field mesh is not built on a slab-decomposed mesh");
  P3M_TRY(complete_potential<T>(c));  // the explicit field mesh needs every plane of the potential
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  if (!s.field) P3M_CUDA(cudaMalloc((void**)&s.field, sizeof(T) * 3 * (size_t)g.M));
  phase_begin(c, PH_GRADIENT);
  const unsigned blocks = (unsigned)((g.M + 255) / 256);
  if (g.fds == P3M_TWO_POINT)
    k_gradient<T, 1><<<blocks, 256, 0, c->stream>>>(s.potential, g, s.field);
  else
    k_gradient<T, 2><<<blocks, 256, 0, c->stream>>>(s.potential, g, s.field);
  P3M_LAUNCH_CHECK(c);
  phase_end(c, PH_GRADIENT);
  c->have_field = true;
  return 0;
}

template int gather<float>(p3m_ctx*);
template int gather<double>(p3m_ctx*);
template int gradient<float>(p3m_ctx*);
template int gradient<double>(p3m_ctx*);

}  // namespace p3m
