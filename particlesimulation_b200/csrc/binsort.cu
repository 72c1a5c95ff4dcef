// binsort.cu -- particle storage, unit conversion on upload/download, cell binning and the z-order
// cell sort.
//
// Replaces: the vector<Particle> copy in PMMethod::PMMethod (source/pmMethod.cpp:57-59),
// stateToCodeUnits / massToCodeUnits (source/unitConversions.cpp:23-71), and
// ChainingMesh::fill / fillWithYSorting (source/chainingMesh.cpp:20-58) -- the linked-list insertion
// becomes one stable radix sort on (Morton(cell), particle id) followed by a gather of the particle
// records, so every cell is a contiguous, id-ordered slice of the particle arrays.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ctx.cuh"
#include "sort_kernels.cuh"

namespace p3m {

template <typename T>
static int dev_alloc(T** p, size_t count) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  P3M_CUDA(cudaMalloc((void**)p, sizeof(T) * (count ? count : 1)));
  return 0;
}

template <typename T>
int alloc_particles(p3m_ctx* c, long long n) {
  State<T>& s = Sel<T>::st(c);
  if (n <= c->cap) return 0;
  long long cap = n + n / 8 + 1024;
  P3M_TRY(dev_alloc(&s.posm, cap));
  P3M_TRY(dev_alloc(&s.posm_alt, cap));
  P3M_TRY(dev_alloc(&s.vel, cap));
  P3M_TRY(dev_alloc(&s.vel_alt, cap));
  P3M_TRY(dev_alloc(&s.acc, cap));
  P3M_TRY(dev_alloc(&s.acc_sr, cap));
  P3M_TRY(dev_alloc(&s.id, cap));
  P3M_TRY(dev_alloc(&s.id_alt, cap));
  // sort scratch: also used for the ghost sort, which can hold up to 2 * cap particles (dist.cu)
  const long long scap = c->nranks > 1 ? 2 * cap : cap;
  P3M_TRY(dev_alloc(&s.keys, scap));
  P3M_TRY(dev_alloc(&s.keys_alt, scap));
  P3M_TRY(dev_alloc(&s.slots, scap));
  P3M_TRY(dev_alloc(&s.slots_alt, scap));
  P3M_TRY(dev_alloc(&s.skeys, cap));
  P3M_TRY(dev_alloc(&s.skeys_alt, cap));
  P3M_TRY(dev_alloc(&s.inc_hist, 1024 * (size_t)(cap / 4096 + 2) + 1024));
  if (!s.inc_counts) P3M_TRY(dev_alloc(&s.inc_counts, 8));
  if (!s.inc_counts_host) P3M_CUDA(cudaMallocHost((void**)&s.inc_counts_host, sizeof(int) * 8));
  c->order_valid = false;
  P3M_TRY(dev_alloc(&s.aabb, 2 * (cap / kPPSub + 8)));
  P3M_TRY(dev_alloc(&s.pp_items, 2 * (cap / kPPTargets + ((size_t)1 << (3 * Sel<T>::g(c).mbits)) + 16)));
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, s.keys, s.keys_alt, s.slots, s.slots_alt, (int)scap, 0,
                                  64, c->stream);
  if (s.cub_tmp) cudaFree(s.cub_tmp);
  s.cub_tmp = nullptr;
  P3M_CUDA(cudaMalloc(&s.cub_tmp, tmp + 16));
  s.cub_tmp_bytes = tmp;
  c->cap = cap;
  return 0;
}

// ---- upload: original units -> code units, exactly the reference's fp32 operations -----------------
//   pos / H                       include/unitConversions.h:8-10
//   DT * v / H                    include/unitConversions.h:15-17
//   factor * m, factor = DT*DT*4*pi*G/(H*H*H) evaluated left to right   include/unitConversions.h:42-44
template <typename T>
__global__ void k_upload(const float* __restrict__ pos, const float* __restrict__ vel,
                         const float* __restrict__ mass, long long n, int units, T H, T DT,
                         T mass_factor, V4<T>* __restrict__ posm, V4<T>* __restrict__ velo,
                         V4<T>* __restrict__ acc, V4<T>* __restrict__ acc_sr, int* __restrict__ id) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  T x = (T)pos[3 * i], y = (T)pos[3 * i + 1], z = (T)pos[3 * i + 2];
  T vx = 0, vy = 0, vz = 0;
  if (vel) vx = (T)vel[3 * i], vy = (T)vel[3 * i + 1], vz = (T)vel[3 * i + 2];
  T m = (T)mass[i];
  if (units == P3M_UNITS_ORIGINAL) {
    x = x / H, y = y / H, z = z / H;
    vx = DT * vx / H, vy = DT * vy / H, vz = DT * vz / H;
    m = mass_factor * m;
  }
  posm[i] = V4<T>{x, y, z, m};
  velo[i] = V4<T>{vx, vy, vz, 0};
  acc[i] = V4<T>{0, 0, 0, 0};
  acc_sr[i] = V4<T>{0, 0, 0, 0};
  id[i] = (int)i;
}

template <typename T>
int upload_particles(p3m_ctx* c, const float* pos, const float* vel, const float* mass, long long n,
                     int units) {
  P3M_TRY(alloc_particles<T>(c, n));
  State<T>& s = Sel<T>::st(c);
  // staging: reuse the sort scratch (keys: 8 B * cap >= 3 floats/particle? no -> dedicated alloc)
  float* stage = nullptr;
  P3M_CUDA(cudaMallocAsync((void**)&stage, sizeof(float) * 7 * (size_t)(n ? n : 1), c->stream));
  P3M_CUDA(cudaMemcpyAsync(stage, pos, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
  if (vel)
    P3M_CUDA(cudaMemcpyAsync(stage + 3 * n, vel, sizeof(float) * 3 * n, cudaMemcpyHostToDevice,
                             c->stream));
  P3M_CUDA(cudaMemcpyAsync(stage + 6 * n, mass, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
  const Geom<T>& g = Sel<T>::g(c);
  T mf = c->f64 ? (T)c->mass_factor64 : (T)c->mass_factor32;
  if (n > 0) {
    k_upload<T><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        stage, vel ? stage + 3 * n : nullptr, stage + 6 * n, n, units, g.H, g.DT, mf, s.posm, s.vel,
        s.acc, s.acc_sr, s.id);
    P3M_LAUNCH_CHECK(c);
  }
  P3M_CUDA(cudaFreeAsync(stage, c->stream));
  // the caller may reuse (pinned) pos / vel / mass as soon as this returns: wait for the copies
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  c->n = n;
  c->n_global = n;
  c->have_particles = true;
  c->sorted = false;
  c->order_valid = false;
  c->have_acc = true;  // all zero, in upload order
  {
    // equal masses?  (positive floats order like their bit patterns)
    float lo = n > 0 ? mass[0] : 0.f, hi = lo;
    for (long long i = 1; i < n; ++i) lo = mass[i] < lo ? mass[i] : lo, hi = mass[i] > hi ? mass[i] : hi;
    c->mass_lo = lo, c->mass_hi = hi;
    c->uniform_mass = n > 0 && lo == hi && lo > 0.f;
    T m = (T)lo;
    if (units == P3M_UNITS_ORIGINAL) m = mf * m;
    c->uniform_mass_code = (double)m;
  }
  // multi-GPU: every rank was handed the whole set; balance the layer cuts on it (identically on every rank),
  // then keep the particles of this rank's z-slab
  if (c->nranks > 1) {
    P3M_TRY(dist_balance_cuts<T>(c, pos, n, units));
    P3M_TRY(dist_migrate<T>(c, false));
  }
  return 0;
}

// A subset with explicit global ids (multi-GPU end-to-end path: each rank uploads what it holds; the
// next p3m_bin_sort migrates anything that is not in this rank's slab).
template <typename T>
__global__ void k_set_ids(const int32_t* __restrict__ ids, long long n, int* __restrict__ id) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) id[i] = ids[i];
}

template <typename T>
int upload_particles_ids(p3m_ctx* c, const float* pos, const float* vel, const float* mass, const int32_t* ids,
                         long long n, int units) {
  const long long ng = c->n_global;
  if (ng <= 0) return fail(P3M_ESTATE, "p3m_set_particles_ids: call p3m_set_particles once first (global count)");
  if (n > c->cap) return fail(P3M_ERANGE, "subset of %lld particles exceeds the capacity %lld", n, c->cap);
  const int nr = c->nranks;
  c->nranks = 1;  // plain upload, no filtering
  int r = upload_particles<T>(c, pos, vel, mass, n, units);
  c->nranks = nr;
  if (r != 0) return r;
  c->n_global = ng;
  if (nr > 1) {
    // the subset may be uniform while the global set is not: agree on min / max over all ranks
    int v[2] = {-0x7fffffff, -0x7fffffff};
    if (n > 0 && c->mass_lo > 0.f) {
      memcpy(&v[0], &c->mass_hi, 4);
      int lo_bits;
      memcpy(&lo_bits, &c->mass_lo, 4);
      v[1] = -lo_bits;
    } else if (n > 0) {
      v[0] = 0x7fffffff, v[1] = 0;  // non-positive mass somewhere: never uniform
    }
    P3M_TRY(dist_allreduce_host_imax(c, v, 2));
    c->uniform_mass = v[0] == -v[1] && v[0] > 0;
    if (c->uniform_mass) {
      float mg;
      memcpy(&mg, &v[0], 4);
      T m = (T)mg;
      if (units == P3M_UNITS_ORIGINAL) m = (c->f64 ? (T)c->mass_factor64 : (T)c->mass_factor32) * m;
      c->uniform_mass_code = (double)m;
    }
  }
  State<T>& s = Sel<T>::st(c);
  if (n > 0) {
    int32_t* stage = nullptr;
    P3M_CUDA(cudaMallocAsync((void**)&stage, sizeof(int32_t) * (size_t)n, c->stream));
    P3M_CUDA(cudaMemcpyAsync(stage, ids, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    k_set_ids<T><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(stage, n, s.id);
    P3M_LAUNCH_CHECK(c);
    P3M_CUDA(cudaFreeAsync(stage, c->stream));
    P3M_CUDA(cudaStreamSynchronize(c->stream));  // `ids` may be pinned memory the caller rewrites next
  }
  return 0;
}

// local particles in local (sorted) order with their global ids
template <typename T>
__global__ void k_download_local(const V4<T>* __restrict__ posm, const V4<T>* __restrict__ vel,
                                 const V4<T>* __restrict__ acc, long long n, int units, T H, T DT,
                                 float* __restrict__ pos_o, float* __restrict__ vel_o, float* __restrict__ acc_o) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool orig = units == P3M_UNITS_ORIGINAL;
  if (pos_o) {
    V4<T> p = posm[i];
    if (orig) p.x = H * p.x, p.y = H * p.y, p.z = H * p.z;
    pos_o[3 * i] = (float)p.x, pos_o[3 * i + 1] = (float)p.y, pos_o[3 * i + 2] = (float)p.z;
  }
  if (vel_o) {
    V4<T> v = vel[i];
    if (orig) v.x = H * v.x / DT, v.y = H * v.y / DT, v.z = H * v.z / DT;
    vel_o[3 * i] = (float)v.x, vel_o[3 * i + 1] = (float)v.y, vel_o[3 * i + 2] = (float)v.z;
  }
  if (acc_o) {
    V4<T> a = acc[i];
    if (orig) a.x = H * a.x / (DT * DT), a.y = H * a.y / (DT * DT), a.z = H * a.z / (DT * DT);
    acc_o[3 * i] = (float)a.x, acc_o[3 * i + 1] = (float)a.y, acc_o[3 * i + 2] = (float)a.z;
  }
}

template <typename T>
int download_local(p3m_ctx* c, int32_t* ids, float* pos, float* vel, float* acc, int units) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const long long n = c->n;
  if (n == 0) return 0;
  float* stage = nullptr;
  P3M_CUDA(cudaMallocAsync((void**)&stage, sizeof(float) * 9 * (size_t)n, c->stream));
  k_download_local<T><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
      s.posm, s.vel, s.acc, n, units, g.H, g.DT, pos ? stage : nullptr, vel ? stage + 3 * n : nullptr,
      acc ? stage + 6 * n : nullptr);
  P3M_LAUNCH_CHECK(c);
  if (pos) P3M_CUDA(cudaMemcpyAsync(pos, stage, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  if (vel) P3M_CUDA(cudaMemcpyAsync(vel, stage + 3 * n, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  if (acc) P3M_CUDA(cudaMemcpyAsync(acc, stage + 6 * n, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  if (ids) P3M_CUDA(cudaMemcpyAsync(ids, s.id, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaFreeAsync(stage, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- download: scatter back to original particle order ---------------------------------------------
//   H * pos, H * v / DT, H * a / (DT*DT)      include/unitConversions.h:11-13,18-20,26-28
template <typename T, typename O>
__global__ void k_download(const V4<T>* __restrict__ posm, const V4<T>* __restrict__ vel,
                           const V4<T>* __restrict__ acc, const int* __restrict__ id, long long n,
                           int units, T H, T DT, O* __restrict__ pos_o, O* __restrict__ vel_o,
                           O* __restrict__ acc_o) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long j = id[i];
  if (pos_o) {
    V4<T> p = posm[i];
    if (units == P3M_UNITS_ORIGINAL) p.x = H * p.x, p.y = H * p.y, p.z = H * p.z;
    pos_o[3 * j] = (O)p.x, pos_o[3 * j + 1] = (O)p.y, pos_o[3 * j + 2] = (O)p.z;
  }
  if (vel_o) {
    V4<T> v = vel[i];
    if (units == P3M_UNITS_ORIGINAL) v.x = H * v.x / DT, v.y = H * v.y / DT, v.z = H * v.z / DT;
    vel_o[3 * j] = (O)v.x, vel_o[3 * j + 1] = (O)v.y, vel_o[3 * j + 2] = (O)v.z;
  }
  if (acc_o) {
    V4<T> a = acc[i];
    if (units == P3M_UNITS_ORIGINAL)
      a.x = H * a.x / (DT * DT), a.y = H * a.y / (DT * DT), a.z = H * a.z / (DT * DT);
    acc_o[3 * j] = (O)a.x, acc_o[3 * j + 1] = (O)a.y, acc_o[3 * j + 2] = (O)a.z;
  }
}

template <typename T, typename O>
int download_particles(p3m_ctx* c, O* pos, O* vel, O* acc, int units) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  // output arrays are indexed by the ORIGINAL (global) particle id; with several ranks each one fills
  // the entries of the particles it holds and leaves the others zero
  const long long nl = c->n;
  const long long n = c->nranks > 1 ? c->n_global : nl;
  if (n == 0) return 0;
  O* stage = nullptr;
  P3M_CUDA(cudaMallocAsync((void**)&stage, sizeof(O) * 9 * (size_t)n, c->stream));
  if (c->nranks > 1) P3M_CUDA(cudaMemsetAsync(stage, 0, sizeof(O) * 9 * (size_t)n, c->stream));
  if (nl > 0) {
    k_download<T, O><<<(unsigned)((nl + 255) / 256), 256, 0, c->stream>>>(
        s.posm, s.vel, s.acc, s.id, nl, units, g.H, g.DT, pos ? stage : nullptr,
        vel ? stage + 3 * n : nullptr, acc ? stage + 6 * n : nullptr);
    P3M_LAUNCH_CHECK(c);
  }
  if (pos) P3M_CUDA(cudaMemcpyAsync(pos, stage, sizeof(O) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  if (vel)
    P3M_CUDA(cudaMemcpyAsync(vel, stage + 3 * n, sizeof(O) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  if (acc)
    P3M_CUDA(cudaMemcpyAsync(acc, stage + 6 * n, sizeof(O) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaFreeAsync(stage, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

template <typename T>
int bin_sort(p3m_ctx* c) {
  if (!c->have_particles) return fail(P3M_ESTATE, "p3m_bin_sort: no particles set");
  State<T>& s = Sel<T>::st(c);
  Geom<T>& g = Sel<T>::g(c);
  if (c->nranks > 1) P3M_TRY(dist_migrate<T>(c, true));  // particles that left this rank's z-slab
  const long long n = c->n;
  const long long nid = c->nranks > 1 ? c->n_global : n;  // ids are global
  int idbits = 1;
  while ((1LL << idbits) < nid) ++idbits;
  g.idbits = idbits;
  g.sbits = 0;
  if (g.p3m) {
    g.sbits = kSubBits;
    if (c->tune.subbits >= 0) g.sbits = c->tune.subbits;  // tuning hook (measurements only)
    // prefer a key that fits 32 bits (cell code + sub-cell, no id: see below): 16^3 sub-cells up to 64 chaining
    // cells per axis, 8^3 up to 128, 4^3 up to 256
    if (!c->tune.long_key)
      while (g.sbits > 2 && 3 * g.mbits + 3 * g.sbits > 32) --g.sbits;
    while (g.sbits > 0 && 3 * g.mbits + 3 * g.sbits + idbits > 62) --g.sbits;
  } else if (3 * g.mbits + 3 * g.tile_shift + idbits <= 62) {
    g.sbits = g.tile_shift;  // PM: sub key = mesh cell inside the tile (sort_kernels.cuh)
  }
  // 32-bit key (cell code, sub-cell) WITHOUT the id whenever it fits: the radix sort is stable, so particles of
  // one sub-cell keep their previous relative order -- id order right after an upload / generation (ids ascend
  // there), arrival order afterwards.  Half the key bytes and 4 instead of 7 radix passes.  The order stays a
  // deterministic function of the history of the run; it is a pure function of (positions, ids) after the first
  // sort only (tests/test_gpu_parity.py::test_cells_and_sort_order_bit_exact).
  const bool short_key = 3 * g.mbits + 3 * g.sbits <= 32 && !c->tune.long_key;
  const int lowbits = (short_key ? 0 : idbits) + 3 * g.sbits;
  const int keybits = lowbits + 3 * g.mbits;
  const long long ncells = 1LL << (3 * g.mbits);
  uint32_t* keys32 = reinterpret_cast<uint32_t*>(s.keys);
  uint32_t* keys32_alt = reinterpret_cast<uint32_t*>(s.keys_alt);
  phase_begin(c, PH_BINSORT);
  if (n > 0) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    size_t tmp = s.cub_tmp_bytes;
    bool merged = false;
    // highest mesh plane a particle sits in (single GPU: bounds the planes that are cleared / transformed)
    const bool want_zmax = c->nranks == 1 && c->fused_z && !c->tune.no_prune;
    s.inc_counts_host[3] = 0;
    if (want_zmax) {
      const int init = -1;
      P3M_CUDA(cudaMemcpyAsync(s.inc_counts + 2, &init, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    if (s.inc_backoff > 0) --s.inc_backoff;
    else if (short_key && c->order_valid && !c->tune.full_sort && s.skeys_n == n && s.skeys_bits == keybits)
      P3M_TRY(sort_incremental<T>(c, keybits, &merged));  // movers only (incsort.cu)
    if (!merged) {
      // the occupied-plane bound is complete once k_keys has run: it is fetched right behind that kernel and waited
      // for (this copy only) at the end of the call, while the radix sort and the permutation are already queued
      auto fetch_zmax = [&]() -> int {
        if (!want_zmax || s.inc_counts_host[3]) return 0;
        if (!s.zmax_event) P3M_CUDA(cudaEventCreateWithFlags(&s.zmax_event, cudaEventDisableTiming));
        P3M_CUDA(cudaMemcpyAsync(s.inc_counts_host, s.inc_counts, sizeof(int) * 3, cudaMemcpyDeviceToHost, c->stream));
        P3M_CUDA(cudaEventRecord(s.zmax_event, c->stream));
        return 0;
      };
      if (short_key) {
        k_keys<T, uint32_t><<<blocks, 256, 0, c->stream>>>(s.posm, s.id, n, g, keys32, s.slots, s.flags,
                                                           want_zmax ? s.inc_counts + 2 : nullptr);
        P3M_LAUNCH_CHECK(c);
        P3M_TRY(fetch_zmax());
        // sorted keys go straight into the persistent array the next (incremental) sort compares against
        P3M_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_tmp, tmp, keys32, s.skeys, s.slots, s.slots_alt, (int)n, 0,
                                                 keybits, c->stream));
      } else {
        k_keys<T, uint64_t><<<blocks, 256, 0, c->stream>>>(s.posm, s.id, n, g, s.keys, s.slots, s.flags,
                                                           want_zmax ? s.inc_counts + 2 : nullptr);
        P3M_LAUNCH_CHECK(c);
        P3M_TRY(fetch_zmax());
        P3M_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_tmp, tmp, s.keys, s.keys_alt, s.slots, s.slots_alt,
                                                 (int)n, 0, keybits, c->stream));
      }
      c->launches += (keybits + 7) / 8 + 1;
      c->full_sorts++;
      k_permute<T><<<blocks, 256, 0, c->stream>>>(s.slots_alt, n, s.posm, s.vel, s.id, s.posm_alt,
                                                  s.vel_alt, s.id_alt);
      P3M_LAUNCH_CHECK(c);
      std::swap(s.posm, s.posm_alt);
      std::swap(s.vel, s.vel_alt);
      std::swap(s.id, s.id_alt);
    }
    c->incr_sort = merged;
    c->have_acc = false;  // acc / acc_sr were not permuted: stale until the next gather
    if (g.p3m) {
      const long long tiles = (n + kPPSub - 1) / kPPSub;
      k_tile_aabb<T><<<(unsigned)((tiles * 32 + 255) / 256), 256, 0, c->stream>>>(s.posm, n, s.aabb);
      P3M_LAUNCH_CHECK(c);
    }
  }
  if (short_key)
    k_cell_start<uint32_t><<<(unsigned)((ncells + 1 + 255) / 256), 256, 0, c->stream>>>(s.skeys, n, lowbits, ncells,
                                                                                       s.cell_start);
  else
    k_cell_start<uint64_t><<<(unsigned)((ncells + 1 + 255) / 256), 256, 0, c->stream>>>(s.keys_alt, n, lowbits, ncells,
                                                                                       s.cell_start);
  P3M_LAUNCH_CHECK(c);
  s.zocc = 0;
  if (n > 0 && c->nranks == 1 && c->fused_z && !c->tune.no_prune) {
    if (!s.inc_counts_host[3] && s.zmax_event) P3M_CUDA(cudaEventSynchronize(s.zmax_event));  // full-sort path: see above
    const int zmax = s.inc_counts_host[2];
    // TSC / CIC write planes (int)z - 1 .. (int)z + 1 (one more when the unwrapped flat index of a particle at the
    // top of y aliases into the next plane): planes [0, zmax + 3) may hold density; rounded up to 16 so
    // that the cached batched FFT plans change rarely
    if (zmax >= 0 && zmax + 3 < c->prm.nz) s.zocc = std::min(c->prm.nz, (zmax + 3 + 15) / 16 * 16);
  }
  phase_end(c, PH_BINSORT);
  c->sorted = true;
  c->order_valid = short_key && n > 0;
  s.skeys_n = n, s.skeys_bits = keybits;
  if (c->nranks > 1) P3M_TRY(dist_ghosts<T>(c));  // boundary-layer particles of the neighbour slabs
  return 0;
}

// ---- test / diagnostics readback: the reference's own flat indices -----------------------------------
template <typename T>
__global__ void k_cells_out(const V4<T>* __restrict__ posm, const int* __restrict__ id, long long n,
                            Geom<T> g, int* __restrict__ mesh_cell, int* __restrict__ chain_cell,
                            int* __restrict__ order) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  V4<T> p = posm[i];
  int j = id[i];
  if (mesh_cell) {
    // (int)pos, flat = x + y*Nx + z*Nx*Ny   source/pmMethod.cpp:250-252, include/grid.h:52-54
    int x = (int)p.x, y = (int)p.y, z = (int)p.z;
    mesh_cell[j] = x + y * g.nx + z * g.nx * g.ny;
  }
  if (chain_cell) {
    int cc = -1;
    if (g.p3m) {
      // source/chainingMesh.cpp:25-29,79-84
      int cx = (int)(p.x / g.hcx), cy = (int)(p.y / g.hcy), cz = (int)(p.z / g.hcz);
      if (!(cx < 0 || cy < 0 || cz < 0 || cx >= g.mx || cy >= g.my || cz >= g.mz))
        cc = cx + cy * g.mx + cz * g.mx * g.my;
    }
    chain_cell[j] = cc;
  }
  if (order) order[i] = j;
}

template <typename T>
int get_cells(p3m_ctx* c, int32_t* mesh_cell, int32_t* chain_cell, int32_t* order) {
  State<T>& s = Sel<T>::st(c);
  const long long n = c->n;                                      // local particles
  const long long ng = c->nranks > 1 ? c->n_global : n;          // id-indexed outputs
  if (ng == 0) return 0;
  int* stage = nullptr;
  P3M_CUDA(cudaMallocAsync((void**)&stage, sizeof(int) * (2 * (size_t)ng + (size_t)n + 1), c->stream));
  if (c->nranks > 1) P3M_CUDA(cudaMemsetAsync(stage, 0xff, sizeof(int) * 2 * (size_t)ng, c->stream));
  if (n > 0) {
    k_cells_out<T><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        s.posm, s.id, n, Sel<T>::g(c), mesh_cell ? stage : nullptr, chain_cell ? stage + ng : nullptr,
        order ? stage + 2 * ng : nullptr);
    P3M_LAUNCH_CHECK(c);
  }
  if (mesh_cell)
    P3M_CUDA(cudaMemcpyAsync(mesh_cell, stage, sizeof(int) * ng, cudaMemcpyDeviceToHost, c->stream));
  if (chain_cell)
    P3M_CUDA(cudaMemcpyAsync(chain_cell, stage + ng, sizeof(int) * ng, cudaMemcpyDeviceToHost, c->stream));
  if (order && n > 0)
    P3M_CUDA(cudaMemcpyAsync(order, stage + 2 * ng, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaFreeAsync(stage, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// acc[i] += a[id[i]]  (host-callback external field, include/unitConversions.h:22-24 for the units)
template <typename T>
__global__ void k_add_acc(V4<T>* __restrict__ acc, const int* __restrict__ id, long long n,
                          const float* __restrict__ a, int units, T H, T DT) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long j = id[i];
  T ax = (T)a[3 * j], ay = (T)a[3 * j + 1], az = (T)a[3 * j + 2];
  if (units == P3M_UNITS_ORIGINAL) ax = DT * DT * ax / H, ay = DT * DT * ay / H, az = DT * DT * az / H;
  V4<T> v = acc[i];
  v.x += ax, v.y += ay, v.z += az;
  acc[i] = v;
}

template <typename T>
int add_acceleration(p3m_ctx* c, const float* a, int units) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const long long n = c->n;
  if (n == 0) return 0;
  float* stage = nullptr;
  P3M_CUDA(cudaMallocAsync((void**)&stage, sizeof(float) * 3 * (size_t)n, c->stream));
  P3M_CUDA(cudaMemcpyAsync(stage, a, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
  k_add_acc<T><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(s.acc, s.id, n, stage, units, g.H, g.DT);
  P3M_LAUNCH_CHECK(c);
  P3M_CUDA(cudaFreeAsync(stage, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

template <typename T>
void free_state(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  slab_free<T>(c);
  void* ptrs[] = {s.posm,   s.posm_alt,  s.vel,      s.vel_alt,  s.acc,        s.acc_sr,  s.id,
                  s.id_alt, s.keys,      s.keys_alt, s.slots,    s.slots_alt,  s.cub_tmp, s.cell_start,
                  s.density, s.potential, s.spectrum, s.green,    s.field,      s.sr_table, s.pp_items, s.aabb, s.gposm, s.gposm_alt, s.gid, s.gid_alt, s.gcell_start, s.gaabb,
                  s.pp_counters, s.pair_counts, s.flags, s.diag, s.skeys, s.skeys_alt, s.inc_hist, s.inc_counts};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (s.inc_counts_host) cudaFreeHost(s.inc_counts_host);
  for (auto& bp : s.batch_plans) cufftDestroy(bp.h);
  if (s.zmax_event) cudaEventDestroy(s.zmax_event);
  if (s.twiddle_z) cudaFree(s.twiddle_z);
  if (s.plans) {
    cufftDestroy(s.plan_fwd);
    cufftDestroy(s.plan_inv);
  }
  s = State<T>();
}

#define INST(T)                                                                                  \
  template int alloc_particles<T>(p3m_ctx*, long long);                                          \
  template int upload_particles<T>(p3m_ctx*, const float*, const float*, const float*, long long, int); \
  template int upload_particles_ids<T>(p3m_ctx*, const float*, const float*, const float*, const int32_t*, long long, int); \
  template int download_local<T>(p3m_ctx*, int32_t*, float*, float*, float*, int); \
  template int download_particles<T, float>(p3m_ctx*, float*, float*, float*, int);              \
  template int download_particles<T, double>(p3m_ctx*, double*, double*, double*, int);          \
  template int bin_sort<T>(p3m_ctx*);                                                            \
  template int get_cells<T>(p3m_ctx*, int32_t*, int32_t*, int32_t*);                             \
  template int add_acceleration<T>(p3m_ctx*, const float*, int);                                 \
  template void free_state<T>(p3m_ctx*);
INST(float)
INST(double)

}  // namespace p3m
