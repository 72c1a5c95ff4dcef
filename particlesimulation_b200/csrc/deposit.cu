// deposit.cu -- A1: mass assignment, PMMethod::spreadMass + Grid::assignDensity / clearDensity
// (source/pmMethod.cpp:200-277, source/grid.cpp:28-36).
//
// The reference takes one std::mutex per touched cell (27 lock/unlock pairs per TSC particle); its
// CUDA mirror issues 27 global float atomics per particle (source/PMMethodGPU.cu:312-336).  sm_100a has no
// native shared-memory float add (atomicAdd on __shared__ float compiles to an ATOMS.CAST.SPIN loop), so
// both kernels here get exclusivity from the work assignment and update shared-memory tiles with plain
// LDS / FADD / STS; tiles are flushed to the global mesh with native REDG.ADD (fp32 and fp64), zeros skipped.
//   k_deposit_pm  PM-only contexts, CIC / TSC: one thread per mesh cell, register accumulation, TMA
//                 bulk-copy double buffering (see the comment above the kernel);
//   k_deposit     P3M contexts and NGP: a WARP owns a run of consecutive sorted particles and a private tile
//                 (the mesh footprint of one aligned block of binning cells).  Per batch of 32 particles the
//                 lanes sharing a base cell are found with match.any:
//                   - all lanes in one cell (dense cores): the 27 contributions are summed across the warp
//                     with shuffles and written once (warp-aggregated);
//                   - otherwise the lowest lane of each group pulls the other members through shuffles and
//                     sums their stencils in registers, so ONE round of read-modify-writes serves the batch
//                     (leaders hold distinct cells, hence distinct addresses per stencil point);
//                 runs too short to amortise a tile (sparse outskirts) go straight to global REDG.
// Algorithmic HBM traffic: 16 B/particle read + 4 B/cell written (+ 4 B/cell memset).
#include <cstdlib>

#include "async_copy.cuh"
#include <algorithm>

#include "ctx.cuh"
#include "stencil.cuh"

namespace p3m {

template <typename T, int K>
__device__ __forceinline__ void deposit_direct(const Stencil<T, K>& s, const Geom<T>& g,
                                               T* __restrict__ density) {
#pragma unroll
  for (int a = 0; a < K; ++a) {
    const T t1 = s.pref * s.wx[a];
#pragma unroll
    for (int b = 0; b < K; ++b) {
      const T t2 = t1 * s.wy[b];
#pragma unroll
      for (int cc = 0; cc < K; ++cc) {
        const T t3 = t2 * s.wz[cc];
        // Grid::getIndx, unwrapped (include/grid.h:52-54)
        long long flat = (long long)(s.x0 + a) + (long long)(s.y0 + b) * g.nx +
                         (long long)(s.z0 + cc) * g.nx * g.ny;
        // (multi-GPU: `density` holds the planes of this rank's particle slab only)
        if (flat >= 0 && flat < g.M) {
          flat -= g.den_off;
          if (flat >= 0 && flat < g.den_len) atomicAdd(&density[flat], t3);
        }
      }
    }
  }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, int K>
__global__ void __launch_bounds__(256)
k_deposit(const V4<T>* __restrict__ posm, long long n, int chunk, const int* __restrict__ cell_start,
          Geom<T> g, T* __restrict__ density) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int tile_cap = g.tex * g.tey * g.tez;
  T* tile = reinterpret_cast<T*>(smem_raw) + (size_t)wib * tile_cap;
  const long long warp_id = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
  long long cur = warp_id * chunk;
  const long long chunk_end = min(n, cur + (long long)chunk);
  const long long ncells = 1LL << (3 * g.mbits);
  const int bs3 = 3 * g.bshift;

  while (cur < chunk_end) {  // warp-uniform loop over the tile segments of this chunk
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

    if (count < g.tile_min) {
      const long long i = cur + lane;
      if (i < seg_end) {
        const V4<T> p = posm[i];
        const Stencil<T, K> s = make_stencil<T, K>(p.x, p.y, p.z, p.w);
        deposit_direct<T, K>(s, g, density);
      }
      cur = seg_end;
      continue;
    }

    int lo[3], ext[3];
    tile_box(g, (int)compact3(blk), (int)compact3(blk >> 1), (int)compact3(blk >> 2), lo, ext);
    const int elems = ext[0] * ext[1] * ext[2];
    for (int e = lane; e < elems; e += 32) tile[e] = T(0);
    __syncwarp();

    for (long long b0 = cur; b0 < seg_end; b0 += 32) {
      const long long i = b0 + lane;
      const bool valid = i < seg_end;
      Stencil<T, K> s;
      int li = -1 - lane;
      bool fits = false;
      T px = 0, py = 0, pz = 0, pm = 0;
      if (valid) {
        const V4<T> p = posm[i];
        px = p.x, py = p.y, pz = p.z, pm = p.w;
        s = make_stencil<T, K>(p.x, p.y, p.z, p.w);
        const int rx = s.x0 - lo[0], ry = s.y0 - lo[1], rz = s.z0 - lo[2];
        fits = rx >= 0 && ry >= 0 && rz >= 0 && rx + K <= ext[0] && ry + K <= ext[1] &&
               rz + K <= ext[2];
        if (fits)
          li = (rz * ext[1] + ry) * ext[0] + rx;
        else
          deposit_direct<T, K>(s, g, density);  // does not fit the tile (rounding slop / stray)
      }
      const unsigned fmask = __ballot_sync(0xffffffffu, fits);
      if (fmask == 0) continue;
      const int leader = __ffs(fmask) - 1;
      const int li_lead = __shfl_sync(0xffffffffu, li, leader);
      const bool uniform = __all_sync(0xffffffffu, !fits || li == li_lead) && __popc(fmask) >= 8;
      if (uniform) {
        // warp-aggregated: one cell for the whole batch
        T mine = T(0);
        int q = 0;
#pragma unroll
        for (int a = 0; a < K; ++a) {
          const T t1 = fits ? s.pref * s.wx[a] : T(0);
#pragma unroll
          for (int b = 0; b < K; ++b) {
            const T t2 = fits ? t1 * s.wy[b] : T(0);
#pragma unroll
            for (int cc = 0; cc < K; ++cc, ++q) {
              const T t3 = fits ? t2 * s.wz[cc] : T(0);
              const T tot = warp_sum<T>(t3);
              if (lane == q) mine = tot;
            }
          }
        }
        if (lane < K * K * K) {
          const int a = lane / (K * K), b = (lane / K) % K, cc = lane % K;
          tile[li_lead + (cc * ext[1] + b) * ext[0] + a] += mine;
        }
        __syncwarp();
      } else {
        // Lanes sharing a base cell: the lowest lane of each group (the leader) pulls the other members'
        // particles through shuffles and sums all their stencils in registers, so that ONE round of
        // read-modify-writes serves the whole batch -- leaders hold distinct base cells, hence distinct
        // addresses for any given stencil point.
        const unsigned grp = __match_any_sync(0xffffffffu, li) & fmask;
        const bool leader = fits && (__ffs(grp) - 1) == lane;
        const int members = fits ? __popc(grp) : 0;
        const int maxmembers = __reduce_max_sync(0xffffffffu, members);
        T v[K * K * K];
        {
          int q = 0;
#pragma unroll
          for (int a = 0; a < K; ++a) {
            const T t1 = s.pref * s.wx[a];
#pragma unroll
            for (int b = 0; b < K; ++b) {
              const T t2 = t1 * s.wy[b];
#pragma unroll
              for (int cc = 0; cc < K; ++cc, ++q) v[q] = t2 * s.wz[cc];
            }
          }
        }
        unsigned rest = leader ? (grp & ~(1u << lane)) : 0u;
        for (int r = 1; r < maxmembers; ++r) {
          const bool pull = rest != 0u;
          const int src = pull ? __ffs(rest) - 1 : lane;
          rest &= rest - 1u;
          const T qx = __shfl_sync(0xffffffffu, px, src), qy = __shfl_sync(0xffffffffu, py, src),
                  qz = __shfl_sync(0xffffffffu, pz, src), qm = __shfl_sync(0xffffffffu, pm, src);
          if (pull) {
            const Stencil<T, K> o = make_stencil<T, K>(qx, qy, qz, qm);
            int q = 0;
#pragma unroll
            for (int a = 0; a < K; ++a) {
              const T t1 = o.pref * o.wx[a];
#pragma unroll
              for (int b = 0; b < K; ++b) {
                const T t2 = t1 * o.wy[b];
#pragma unroll
                for (int cc = 0; cc < K; ++cc, ++q) v[q] += t2 * o.wz[cc];
              }
            }
          }
        }
        {
          int q = 0;
#pragma unroll
          for (int a = 0; a < K; ++a)
#pragma unroll
            for (int b = 0; b < K; ++b)
#pragma unroll
              for (int cc = 0; cc < K; ++cc, ++q) {
                if (leader) tile[li + (cc * ext[1] + b) * ext[0] + a] += v[q];
                __syncwarp();  // orders the RMW of different lanes on overlapping stencils
              }
        }
      }
    }
    __syncwarp();

    // flush: native REDG.ADD to the global mesh, zeros skipped
    const float inv0 = 1.0f / (float)ext[0], inv1 = 1.0f / (float)ext[1];
    for (int e = lane; e < elems; e += 32) {
      const T v = tile[e];
      if (v != T(0)) {
        const int q = fast_div(e, ext[0], inv0);
        const int ix = e - q * ext[0];
        const int iz = fast_div(q, ext[1], inv1);
        const int iy = q - iz * ext[1];
        long long flat = (long long)(lo[0] + ix) + (long long)(lo[1] + iy) * g.nx +
                         (long long)(lo[2] + iz) * g.nx * g.ny;
        if (flat >= 0 && flat < g.M) {
          flat -= g.den_off;
          if (flat >= 0 && flat < g.den_len) atomicAdd(&density[flat], v);
        }
      }
    }
    __syncwarp();
    cur = seg_end;
  }
}

// ---- PM-only contexts: one thread per mesh cell ----------------------------------------------------------
// Binning cells are 8^3 mesh cells and the sort key orders the particles of a binning cell by mesh cell
// (x fastest), so each mesh cell is a contiguous run.  A CTA of 512 threads takes one binning cell:
//   1. stage its particles in shared memory and find every mesh cell's run from the head / tail positions
//      (cell id differs from the neighbour's) -- no atomics, no scan;
//   2. thread t sums the K^3 stencil contributions of all particles of mesh cell t IN REGISTERS;
//   3. K^3 read-modify-write rounds on a 10^3 shared tile: in one round every thread adds the same stencil
//      offset, so all addresses are distinct (no atomics needed -- sm_100a has no native shared float add);
//      rounds along x only interact inside a warp (__syncwarp), a change of the (y, z) offset is a
//      __syncthreads.  Row pitch 24 keeps the 4 rows of a warp in disjoint banks;
//   4. the tile goes to the global mesh with native REDG.ADD, zeros skipped.
// Duplicates, rank rounds and match.any of the generic kernel disappear; the work per particle is the
// register FMAs only.  Non-empty binning cells come from a compacted list (k_occupied_cells), walked by a
// persistent grid.
__global__ void k_occupied_cells(const int* __restrict__ cell_start, long long ncells, int* __restrict__ list,
                                 int* __restrict__ counter) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  if (cell_start[c + 1] > cell_start[c]) list[atomicAdd(counter, 1)] = (int)c;
}

template <typename T, int K>
__global__ void __launch_bounds__(512, 2)
k_deposit_pm(const V4<T>* __restrict__ posm, const int* __restrict__ cell_start, const int* __restrict__ list,
             const int* __restrict__ counter, Geom<T> g, T* __restrict__ density) {
  constexpr int TE = 10, TP = 24, CAP = sizeof(T) == 8 ? 256 : 1024, B = 1 << kPmTileShift;
  constexpr int O = (K == 3) ? 0 : 1;  // tile coordinate of stencil point 0 of local cell 0 (tile origin = cell - 1)
  __shared__ __align__(16) V4<T> sp_all[2][CAP];  // double buffer: item k+1 lands while item k is processed
  __shared__ T tile[TE * TE * TP];
  __shared__ short cid[CAP + 2];
  __shared__ short first[B * B * B], last[B * B * B];
  __shared__ uint64_t bar[2];
  __shared__ int meta[2][3];  // (cell, s, e) of the item whose first CAP particles are in buffer b
  const int tid = threadIdx.x;
  const int lx = tid & (B - 1), ly = (tid >> kPmTileShift) & (B - 1), lz = tid >> (2 * kPmTileShift);
  const int nocc = *counter;
  const int G = gridDim.x;
  if (tid == 0) {
    mbar_init(&bar[0], 1), mbar_init(&bar[1], 1);
    mbar_init_fence();
  }
  __syncthreads();
  // thread 0 runs the copy pipeline: `cell_next` is the binning cell of the NEXT item
  int cell_next = -1;
  if (tid == 0 && (int)blockIdx.x < nocc) {
    const int c0 = list[blockIdx.x];
    const int s0 = cell_start[c0], e0 = cell_start[c0 + 1];
    meta[0][0] = c0, meta[0][1] = s0, meta[0][2] = e0;
    const uint32_t bytes = (uint32_t)min(CAP, e0 - s0) * (uint32_t)sizeof(V4<T>);
    mbar_arrive_expect_tx(&bar[0], bytes);
    bulk_copy_g2s(sp_all[0], posm + s0, bytes, &bar[0]);
    if ((int)blockIdx.x + G < nocc) cell_next = list[blockIdx.x + G];
  }
  __syncthreads();
  int it = 0;
  for (int w = blockIdx.x; w < nocc; w += G, ++it) {
    const int bsel = it & 1;
    V4<T>* sp = sp_all[bsel];
    // metadata of the next item: loads issued now, consumed after the first barrier below
    int s_n = 0, e_n = 0, cell_nn = -1;
    if (tid == 0 && cell_next >= 0) {
      s_n = cell_start[cell_next], e_n = cell_start[cell_next + 1];
      if (w + 2 * G < nocc) cell_nn = list[w + 2 * G];
    }
    const uint32_t cell = (uint32_t)meta[bsel][0];
    const int s = meta[bsel][1], e = meta[bsel][2];
    const int ox = (int)compact3(cell) << kPmTileShift, oy = (int)compact3(cell >> 1) << kPmTileShift,
              oz = (int)compact3(cell >> 2) << kPmTileShift;
    for (int el = tid; el < TE * TE * TP; el += 512) tile[el] = T(0);
    T* mytile = tile + ((lz + O) * TE + (ly + O)) * TP + lx + O;
    mbar_wait(&bar[bsel], (uint32_t)((it >> 1) & 1));  // this item's particles have landed
    for (int pass = s; pass < e; pass += CAP) {
      const int n = min(CAP, e - pass);
      first[tid] = 0, last[tid] = 0;
      if (tid == 0) cid[0] = -2, cid[n + 1] = -3;
      bool stray = false;
      for (int k = tid; k < n; k += 512) {
        V4<T> p;
        if (pass == s) {
          p = sp[k];
        } else {  // runs longer than CAP: later passes are loaded synchronously
          p = posm[pass + k];
          sp[k] = p;
        }
        const int bx = (int)p.x - ox, by = (int)p.y - oy, bz = (int)p.z - oz;  // base cell
        const bool in = bx >= 0 && by >= 0 && bz >= 0 && bx < B && by < B && bz < B;
        cid[k + 1] = in ? (short)((bz * B + by) * B + bx) : (short)-1;
        stray |= !in;
      }
      const bool any_stray = __syncthreads_or(stray);
      if (pass == s && tid == 0 && cell_next >= 0) {
        // the other buffer was released by the barrier at the end of the previous item: start the next copy
        meta[bsel ^ 1][0] = cell_next, meta[bsel ^ 1][1] = s_n, meta[bsel ^ 1][2] = e_n;
        const uint32_t bytes = (uint32_t)min(CAP, e_n - s_n) * (uint32_t)sizeof(V4<T>);
        proxy_fence_async();
        mbar_arrive_expect_tx(&bar[bsel ^ 1], bytes);
        bulk_copy_g2s(sp_all[bsel ^ 1], posm + s_n, bytes, &bar[bsel ^ 1]);
      }
      if (any_stray) {
        // a particle binned here by clamping (outside the mesh): exact global path for this pass
        for (int k = tid; k < n; k += 512) {
          const V4<T> p = sp[k];
          const Stencil<T, K> st = make_stencil<T, K>(p.x, p.y, p.z, p.w);
          deposit_direct<T, K>(st, g, density);
        }
        __syncthreads();
        continue;
      }
      for (int k = tid; k < n; k += 512) {
        const int c0 = cid[k + 1];
        if (c0 != cid[k]) first[c0] = (short)k;
        if (c0 != cid[k + 2]) last[c0] = (short)(k + 1);
      }
      __syncthreads();
      // registers: the K^3 stencil sums of this thread's mesh cell (runs longer than CAP simply take several
      // passes, each followed by its own read-modify-write rounds)
      T acc[K * K * K];
#pragma unroll
      for (int q = 0; q < K * K * K; ++q) acc[q] = T(0);
      const int j0 = first[tid], j1 = last[tid];
      for (int j = j0; j < j1; ++j) {
        const V4<T> p = sp[j];
        const Stencil<T, K> st = make_stencil<T, K>(p.x, p.y, p.z, p.w);
#pragma unroll
        for (int a = 0; a < K; ++a) {
          const T t1 = st.pref * st.wx[a];
#pragma unroll
          for (int b = 0; b < K; ++b) {
            const T t2 = t1 * st.wy[b];
#pragma unroll
            for (int cc = 0; cc < K; ++cc) acc[(a * K + b) * K + cc] += t2 * st.wz[cc];
          }
        }
      }
      const bool mine = j1 > j0;
#pragma unroll
      for (int cc = 0; cc < K; ++cc)
#pragma unroll
        for (int b = 0; b < K; ++b) {
          __syncthreads();  // rows of other warps (the first one also fences sp / cid / first / last)
#pragma unroll
          for (int a = 0; a < K; ++a) {
            if (mine) mytile[(cc * TE + b) * TP + a] += acc[(a * K + b) * K + cc];
            __syncwarp();  // x neighbours live in the same warp
          }
        }
    }
    cell_next = cell_nn;
    __syncthreads();
    for (int el = tid; el < TE * TE * TE; el += 512) {
      const int iz = el / (TE * TE), r = el - iz * (TE * TE), iy = r / TE, ix = r - iy * TE;
      const T v = tile[(iz * TE + iy) * TP + ix];
      if (v != T(0)) {
        // Grid::getIndx, unwrapped (include/grid.h:52-54)
        long long flat = (long long)(ox - 1 + ix) + (long long)(oy - 1 + iy) * g.nx +
                         (long long)(oz - 1 + iz) * g.nx * g.ny;
        if (flat >= 0 && flat < g.M) {
          flat -= g.den_off;
          if (flat >= 0 && flat < g.den_len) atomicAdd(&density[flat], v);
        }
      }
    }
    __syncthreads();
  }
}

template <typename T, int K>
static int launch_deposit_pm(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const long long ncells = 1LL << (3 * g.mbits);
  int* list = s.pp_items;  // free in PM-only contexts; sized >= ncells entries
  int* counter = s.pp_counters + 6;
  P3M_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), c->stream));
  k_occupied_cells<<<(unsigned)((ncells + 255) / 256), 256, 0, c->stream>>>(s.cell_start, ncells, list, counter);
  P3M_LAUNCH_CHECK(c);
  k_deposit_pm<T, K><<<c->num_sms * 2, 512, 0, c->stream>>>(s.posm, s.cell_start, list, counter, g, s.dens_part);
  P3M_LAUNCH_CHECK(c);
  return 0;
}

template <typename T, int K>
static int launch_deposit(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const long long n = c->n;
  // chunk: enough warps to fill the machine for small N, kDepositChunk for large N
  long long want = (n + (long long)c->num_sms * 64 - 1) / ((long long)c->num_sms * 64);
  int chunk = (int)((want + 31) / 32 * 32);
  if (chunk < 64) chunk = 64;
  if (chunk > kDepositChunk) chunk = kDepositChunk;
  const long long warps = (n + chunk - 1) / chunk;
  const int wpb = 8;
  const size_t smem = (size_t)wpb * g.tex * g.tey * g.tez * sizeof(T);
  auto kern = k_deposit<T, K>;
  P3M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, smem, c->stream>>>(s.posm, n, chunk,
                                                                         s.cell_start, g, s.dens_part);
  P3M_LAUNCH_CHECK(c);
  return 0;
}

template <typename T>
int deposit(p3m_ctx* c) {
  if (!c->have_particles) return fail(P3M_ESTATE, "p3m_deposit: no particles set");
  if (!c->sorted) return fail(P3M_ESTATE, "p3m_deposit: call p3m_bin_sort first");
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  phase_begin(c, PH_DEPOSIT);
  // Grid::clearDensity (source/grid.cpp:34-36)
  // single GPU with a known occupied z range (State::zocc, from the sort): only the planes that were or will be
  // written are cleared -- the padding half of the mesh stays zero from the first clear
  {
    size_t clear = (size_t)g.den_len;
    const size_t plane = (size_t)g.nx * g.ny;
    if (c->nranks == 1 && s.zocc > 0 && s.dens_dirty >= 0) clear = plane * (size_t)std::max(s.zocc, s.dens_dirty);
    if (clear > (size_t)g.den_len) clear = (size_t)g.den_len;
    P3M_CUDA(cudaMemsetAsync(s.dens_part, 0, sizeof(T) * clear, c->stream));
    s.dens_dirty = c->nranks == 1 && s.zocc > 0 ? s.zocc : -1;
    s.dens_occ = c->nranks == 1 ? s.zocc : 0;
  }
  c->launches++;
  int r = 0;
  const bool pm_cells = !g.p3m && g.tile_shift == kPmTileShift && g.sbits == g.tile_shift && g.is != P3M_NGP &&
                        !c->tune.old_deposit;
  if (c->n > 0 && pm_cells) {
    r = g.is == P3M_TSC ? launch_deposit_pm<T, 3>(c) : launch_deposit_pm<T, 2>(c);
  } else if (c->n > 0) {
    if (g.is == P3M_TSC)
      r = launch_deposit<T, 3>(c);
    else if (g.is == P3M_CIC)
      r = launch_deposit<T, 2>(c);
    else
      r = launch_deposit<T, 1>(c);
  }
  phase_end(c, PH_DEPOSIT);
  // multi-GPU: planes go to the ranks that own them in the FFT slab decomposition (or, replicated-mesh
  // fallback, one all-reduce of the full mesh)
  if (r == 0 && c->nranks > 1) r = c->slab ? slab_reduce_density<T>(c) : dist_allreduce_density<T>(c);
  c->have_density = (r == 0);
  return r;
}

template int deposit<float>(p3m_ctx*);
template int deposit<double>(p3m_ctx*);

}  // namespace p3m
