// sort_kernels.cuh -- kernels shared by the local cell sort (binsort.cu) and the ghost-particle sort
// of the multi-GPU path (dist.cu).
#pragma once

#include "ctx.cuh"

namespace p3m {

// ---- A0: sort keys ---------------------------------------------------------------------------------
// KeyT = uint64_t: (Morton(cell), sub-cell, particle id) -- the order is a pure function of positions and ids.
// KeyT = uint32_t (PM-only contexts): (Morton(tile), mesh cell in the tile) without the id; the radix sort
// is stable, so ties keep their previous relative order (id order right after an upload).  Half the key
// bytes and 4 instead of 7 radix passes on a 512^3 mesh.
// (Morton(cell), sub-cell) code of a position, without the id bits
template <typename T>
__device__ __forceinline__ uint64_t cell_key(const Geom<T>& g, const V4<T>& p, bool& inside) {
  int cx, cy, cz;
  bin_cell(g, p.x, p.y, p.z, cx, cy, cz, inside);
  uint64_t m = morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
  if (g.sbits && !g.p3m) {
    // PM only: the mesh cell inside the 8^3 tile, x fastest -- neighbouring lanes then touch
    // neighbouring shared-memory words in the deposit / gather tiles
    const int mask = (1 << g.tile_shift) - 1;
    const int lx = ((int)p.x) & mask, ly = ((int)p.y) & mask, lz = ((int)p.z) & mask;
    m = (m << (3 * g.sbits)) | (uint64_t)((lz << (2 * g.tile_shift)) | (ly << g.tile_shift) | lx);
  } else if (g.sbits) {
    // position inside the chaining cell in units of 1/2^sbits of the cell: x/HC - cx is exact in
    // floating point (cx is the truncation of the same quotient), so the sub-cell is reproducible
    const int S = 1 << g.sbits;
    int sx = (int)((p.x / g.hcx - (T)cx) * (T)S), sy = (int)((p.y / g.hcy - (T)cy) * (T)S),
        sz = (int)((p.z / g.hcz - (T)cz) * (T)S);
    sx = min(max(sx, 0), S - 1), sy = min(max(sy, 0), S - 1), sz = min(max(sz, 0), S - 1);
    m = (m << (3 * g.sbits)) | morton3((uint32_t)sx, (uint32_t)sy, (uint32_t)sz);
  }
  return m;
}

// highest mesh plane (int)z of a particle, clamped to [0, 2^30)
template <typename T>
__device__ __forceinline__ int plane_of(const V4<T>& p) {
  const float zf = (float)p.z;
  return zf < 0.f ? 0 : (zf > 1e9f ? 0x3fffffff : (int)zf);
}
// block maximum -> *zmax with ONE guarded atomic per CTA (most CTAs only read); call from every thread of the CTA
__device__ __forceinline__ void block_zmax(int z, int* __restrict__ zmax) {
  __shared__ int s_zmax;
  if (threadIdx.x == 0) s_zmax = -1;
  __syncthreads();
  const int wz = __reduce_max_sync(0xffffffffu, z);
  if ((threadIdx.x & 31) == 0 && wz >= 0) atomicMax(&s_zmax, wz);
  __syncthreads();
  if (threadIdx.x == 0 && s_zmax > *reinterpret_cast<volatile int*>(zmax)) atomicMax(zmax, s_zmax);
}

template <typename T, typename KeyT>
__global__ void k_keys(const V4<T>* __restrict__ posm, const int* __restrict__ id, long long n,
                       Geom<T> g, KeyT* __restrict__ keys, uint32_t* __restrict__ slots,
                       int* __restrict__ flags, int* __restrict__ zmax = nullptr) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int z = -1;
  if (i < n) {
    bool inside;
    const V4<T> pp = posm[i];
    const uint64_t m = cell_key(g, pp, inside);
    z = plane_of(pp);
    if (!inside) flags[1] = 1;
    if (sizeof(KeyT) == 8)
      keys[i] = (KeyT)((m << g.idbits) | (uint64_t)(uint32_t)id[i]);
    else
      keys[i] = (KeyT)m;
    slots[i] = (uint32_t)i;
  }
  if (zmax) block_zmax(z, zmax);  // uniform over the CTA
}

template <typename T>
__global__ void k_permute(const uint32_t* __restrict__ slots, long long n,
                          const V4<T>* __restrict__ posm, const V4<T>* __restrict__ vel,
                          const int* __restrict__ id, V4<T>* __restrict__ posm_o,
                          V4<T>* __restrict__ vel_o, int* __restrict__ id_o) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = slots[i];
  posm_o[i] = posm[s];
  vel_o[i] = vel[s];
  id_o[i] = id[s];
}

// cell_start[c] = first sorted slot whose cell code is >= c (lower bound), c in [0, ncells]
template <typename KeyT>
static __global__ void k_cell_start(const KeyT* __restrict__ keys, long long n, int idbits,
                             long long ncells, int* __restrict__ cell_start) {
  long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c > ncells) return;
  const uint64_t target = (uint64_t)c << idbits;
  long long lo = 0, hi = n;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (keys[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  cell_start[c] = (int)lo;
}

// bounding box of every globally aligned group of kPPSub consecutive (sorted) particles: one warp per
// group, coalesced 128-bit loads.  Consumed by the short-range kernel to skip source tiles
// that cannot reach a target group.
template <typename T>
__global__ void __launch_bounds__(256)
k_tile_aabb(const V4<T>* __restrict__ posm, long long n, V4<T>* __restrict__ aabb) {
  const long long tile = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long b = tile * kPPSub;
  if (b >= n) return;
  T lx = 1e30, ly = 1e30, lz = 1e30, hx = -1e30, hy = -1e30, hz = -1e30;
  for (int k = lane; k < kPPSub && b + k < n; k += 32) {
    const V4<T> p = posm[b + k];
    lx = min(lx, p.x), ly = min(ly, p.y), lz = min(lz, p.z);
    hx = max(hx, p.x), hy = max(hy, p.y), hz = max(hz, p.z);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lx = min(lx, __shfl_xor_sync(0xffffffffu, lx, o)), ly = min(ly, __shfl_xor_sync(0xffffffffu, ly, o));
    lz = min(lz, __shfl_xor_sync(0xffffffffu, lz, o)), hx = max(hx, __shfl_xor_sync(0xffffffffu, hx, o));
    hy = max(hy, __shfl_xor_sync(0xffffffffu, hy, o)), hz = max(hz, __shfl_xor_sync(0xffffffffu, hz, o));
  }
  if (lane == 0) {
    aabb[2 * tile] = V4<T>{lx, ly, lz, 0};
    aabb[2 * tile + 1] = V4<T>{hx, hy, hz, 0};
  }
}


}  // namespace p3m
