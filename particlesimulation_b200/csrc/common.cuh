// common.cuh -- shared types of the B200-native P3M force step (sm_100a only).
//
// Data layout in HBM:
//   particles  SoA of 16-byte (fp32) / 32-byte (fp64) records, kept PERSISTENTLY in cell-sorted order:
//                posm[i] = (x, y, z, mass)   code units        one LDG.128 per particle
//                vel[i]  = (vx, vy, vz, -)
//                acc[i]  = (ax, ay, az, -)
//                id[i]   = original particle index (index into the caller's arrays)
//   meshes     real density / potential of Nx*Ny*Nz (x fastest, as include/grid.h:52-54 of the
//              reference), half-spectrum complex (Nx/2+1)*Ny*Nz for the R2C transform, and the
//              real, Hermitian-symmetrised influence function on the half spectrum.
//   cells      cell_start[] indexed by the Morton code of the binning cell (chaining-mesh cell for
//              P3M, 8^3 mesh-cell tile for PM-only).
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

#include <string>

#include "../../include/p3m_b200.h"

namespace p3m {

template <typename T>
struct alignas(4 * sizeof(T)) V4 {
  T x, y, z, w;
};

template <typename T>
struct CufftTypes;
template <>
struct CufftTypes<float> {
  using real = cufftReal;
  using cplx = cufftComplex;
};
template <>
struct CufftTypes<double> {
  using real = cufftDoubleReal;
  using cplx = cufftDoubleComplex;
};

// ---- geometry handed to kernels by value -------------------------------------------------------
template <typename T>
struct Geom {
  // mesh (Grid, include/grid.h:12)
  int nx, ny, nz;
  long long M;
  int is;   // P3M_NGP / CIC / TSC
  int fds;  // P3M_TWO_POINT / FOUR_POINT
  // binning mesh: chaining mesh (p3m) or tiles of 2^tile_shift mesh cells (PM only)
  int p3m;
  int mx, my, mz;   // binning cells per axis
  int mbits;        // Morton bits per axis: 2^mbits >= max(mx,my,mz)
  int idbits;       // bits of the particle id packed under the Morton code in the sort key
  int sbits;        // sub-cell bits per axis sorted between the cell code and the id (P3M: groups
                    // particles spatially INSIDE a chaining cell so short-range tiles are compact)
  T hcx, hcy, hcz;  // chaining cell size in code units (source/chainingMesh.cpp:13-15)
  int tile_shift;   // PM only: binning cell = (1 << tile_shift)^3 mesh cells
  // deposit / gather tile: an aligned block of (1 << bshift)^3 binning cells, contiguous in key order
  int bshift;
  int tex, tey, tez;  // max tile extent in mesh cells (incl. assignment halo)
  int tile_min;       // segments with fewer particles bypass the tile (direct global path)
  // P3M contexts, gather: aligned block of (1 << gshift)^3 chaining cells whose mesh footprint (assignment +
  // finite-difference halo included) fits the 24^3 potential tile of k_gather_p3m; -1: no such block
  int gshift;
  // external field (source/externalFields.cpp:4-15) in ORIGINAL units, plus unit factors
  int ext_kind;
  T ecx, ecy, ecz, eR, eM, G;
  T H, DT;
  T boxx, boxy, boxz;  // effective box, original units (escape check)
  int unit_roundtrip;
  // z-slab decomposition of the particles over nranks GPUs (dist.cu); nranks == 1: everything local.
  // Binning-cell layers [cut[r], cut[r+1]) along z belong to rank r (the last rank also takes every
  // layer above cut[nranks]); lay0/lay1 = this rank's own range.
  int nranks, rank;
  int cut[P3M_MAX_RANKS + 1];
  int lay0, lay1;
  // planes held by the buffers the particle kernels address (single GPU: the whole mesh)
  long long den_off, den_len;  // density buffer = global flat indices [den_off, den_off + den_len)
  int pot_z0, pot_nz;          // potential buffer = unwrapped planes [pot_z0, pot_z0 + pot_nz)
};

// plane of the local potential buffer holding (periodic) mesh plane z, or -1 if it is not held
template <typename T>
__device__ __forceinline__ int pot_plane(const Geom<T>& g, int z) {
  int zl = z - g.pot_z0;
  zl = zl < 0 ? zl + g.nz : zl;
  zl = zl >= g.nz ? zl - g.nz : zl;
  if (zl < 0 || zl >= g.nz) zl = ((zl % g.nz) + g.nz) % g.nz;
  return zl < g.pot_nz ? zl : -1;
}

// owner rank of a binning-cell layer
template <typename T>
__host__ __device__ inline int layer_owner(const Geom<T>& g, int cz) {
  int r = 0;
  for (int k = 1; k < g.nranks; ++k) r += (cz >= g.cut[k]) ? 1 : 0;
  return r;
}

template <typename T>
struct SRParams {
  T re2;         // cutoff^2, code units (source/p3mMethod.cpp:35,258)
  T inv_delta2;  // 1 / deltaSquared (source/p3mMethod.cpp:44)
  T a;           // particle diameter, code units
  T eps2;        // softening^2
  int use_table;
  int cloud;
};

// ---- Morton (z-order) code of a binning cell, 10 bits per axis ------------------------------------
__host__ __device__ inline uint32_t spread3(uint32_t v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__host__ __device__ inline uint32_t compact3(uint32_t v) {
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030C30C3u;
  v = (v | (v >> 4)) & 0x0300F00Fu;
  v = (v | (v >> 8)) & 0x030000FFu;
  v = (v | (v >> 16)) & 0x3ffu;
  return v;
}
__host__ __device__ inline uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
  return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}

// ---- binning cell of a position: bit-exact restatement of the reference's index arithmetic --------
//   mesh cell      (int)pos                      source/pmMethod.cpp:250-252
//   chaining cell  int(pos / HC)  (a DIVISION)   source/chainingMesh.cpp:25-29, SURVEY Q3
template <typename T>
__host__ __device__ inline void bin_cell(const Geom<T>& g, T x, T y, T z, int& cx, int& cy, int& cz,
                                         bool& inside) {
  if (g.p3m) {
    cx = (int)(x / g.hcx);
    cy = (int)(y / g.hcy);
    cz = (int)(z / g.hcz);
  } else {
    cx = ((int)x) >> g.tile_shift;
    cy = ((int)y) >> g.tile_shift;
    cz = ((int)z) >> g.tile_shift;
  }
  inside = cx >= 0 && cy >= 0 && cz >= 0 && cx < g.mx && cy < g.my && cz < g.mz && x >= 0 &&
           y >= 0 && z >= 0;
  cx = min(max(cx, 0), g.mx - 1);
  cy = min(max(cy, 0), g.my - 1);
  cz = min(max(cz, 0), g.mz - 1);
}

// mesh-cell origin and extent of the tile owned by block (bx,by,bz) of binning cells
template <typename T>
__host__ __device__ inline void tile_box(const Geom<T>& g, int bx, int by, int bz, int lo[3],
                                         int ext[3]) {
  const int B = 1 << g.bshift;
  if (g.p3m) {
    const T h[3] = {g.hcx, g.hcy, g.hcz};
    const int b[3] = {bx, by, bz};
    for (int d = 0; d < 3; ++d) {
      int l = (int)floor((double)(b[d] * B) * (double)h[d]) - 1;
      int u = (int)floor((double)((b[d] + 1) * B) * (double)h[d]) + 1;
      lo[d] = l;
      ext[d] = u - l + 1;
    }
  } else {
    const int s = g.tile_shift;
    lo[0] = ((bx * B) << s) - 1, lo[1] = ((by * B) << s) - 1, lo[2] = ((bz * B) << s) - 1;
    ext[0] = ext[1] = ext[2] = (B << s) + 2;
  }
}

// ---- error plumbing --------------------------------------------------------------------------------
int fail(int code, const char* fmt, ...);
#define P3M_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return ::p3m::fail(P3M_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,            \
                         cudaGetErrorString(e__));                                          \
  } while (0)
#define P3M_FFT(expr)                                                                       \
  do {                                                                                      \
    cufftResult r__ = (expr);                                                               \
    if (r__ != CUFFT_SUCCESS)                                                               \
      return ::p3m::fail(P3M_ECUDA, "%s:%d %s -> cufft error %d", __FILE__, __LINE__, #expr, \
                         (int)r__);                                                         \
  } while (0)
#define P3M_TRY(expr)        \
  do {                       \
    int r__ = (expr);        \
    if (r__ != 0) return r__; \
  } while (0)

constexpr int kPmTileShift = 3;      // PM-only binning cell = 8^3 mesh cells
constexpr int kDepositChunk = 512;   // particles per warp work item (deposit)
constexpr int kGatherChunk = 2048;   // particles per CTA work item (gather)
constexpr int kTileMinCount = 24;    // segments shorter than this take the direct (global) path
constexpr int kMaxTileBytes = 12288; // per-warp deposit tile budget
constexpr int kSRTable = 500;        // tabulatedValuesCnt (source/p3mMethod.cpp:42)
constexpr int kDenseCell = 32;       // chaining cells with >= this many particles use the warp-per-64-targets PP kernel
                                     // (measured: 64 -> 32 takes the uniform C5 set from 11.3 to 9.2 ms, C2 unchanged)
constexpr int kPPTargets = 64;       // targets per dense-cell work item (one warp, 2 per lane)
#ifndef P3M_PP_SUB
#define P3M_PP_SUB 32
#endif
constexpr int kPPSub = P3M_PP_SUB;           // particles per bounding box / staged source group (globally aligned)
constexpr int kSubBits = 4;          // 16^3 sub-cells per chaining cell in the sort key
constexpr int kGatherTile = 24;      // pitch and maximum extent of the potential tile of k_gather_p3m

}  // namespace p3m
