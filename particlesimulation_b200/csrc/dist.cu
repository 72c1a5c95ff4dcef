// dist.cu -- multi-GPU: z-slab decomposition of the PARTICLES over 1/2/4/8 GPUs with NCCL
// (nothing of this exists in the reference, which is single-process).
//
// One process per GPU.  Binning-cell layers along z (chaining-mesh layers for P3M, 8-cell tile layers
// for PM) are cut into contiguous ranges, one per rank.  Per force evaluation:
//   1. migration   particles whose layer now belongs to another rank are sent there: a 1-pass radix
//                  sort on the destination rank groups them, counts travel with one small all-gather,
//                  payloads (x,y,z,m | vx,vy,vz | id) with grouped ncclSend/ncclRecv straight out of /
//                  into the particle arrays.  Arrival order does not matter: the (cell, sub-cell, id)
//                  sort that follows restores a deterministic order.
//   2. ghosts      (P3M) the particles of the boundary layer on each side of a slab are copied to the
//                  neighbouring rank (16 B + id each) and sorted there by the same key; the short-range
//                  kernels read neighbour cells of foreign layers from these ghost arrays.  Forces on
//                  ghosts are never computed (gather form), so nothing is sent back.
//   3. mesh        round 1: every rank deposits its own particles into a full-size mesh and the density
//                  is summed with ONE ncclAllReduce (4 B/cell); the Poisson solve is then replicated.
//                  This is exact and costs no halo logic, but the mesh does not shrink with the rank
//                  count: fine up to ~512^3; the slab-decomposed FFT (all-to-all transpose) replaces
//                  this step next.
// Diagnostics and the escape flag are all-reduced so that every rank takes the same decisions.
#include <nccl.h>

#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.cuh"
#include "sort_kernels.cuh"

namespace p3m {

#define P3M_NCCL(expr)                                                                        \
  do {                                                                                        \
    ncclResult_t r__ = (expr);                                                                \
    if (r__ != ncclSuccess)                                                                   \
      return ::p3m::fail(P3M_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,              \
                         ncclGetErrorString(r__));                                            \
  } while (0)

constexpr int kCntSeg = 0;     // [0..P]   segment starts of the destination-sorted particles
constexpr int kCntMine = 16;   // [16..16+P) my per-destination counts, [16+P] my particle capacity
constexpr int kCntAll = 32;    // [32..32+P*(P+1)) everybody's counts + capacity (row = source rank, stride P+1)
constexpr int kCntGhost = 112;  // [112, 113] my ghost counts (to rank-1, to rank+1), [114] my ghost receive capacity
constexpr int kCntGhostAll = 116;  // [116..116+3P) everybody's ghost counts and capacities
constexpr int kCntTotal = 160;

template <typename T>
static void set_cuts(p3m_ctx* c) {
  Geom<T>& g = Sel<T>::g(c);
  const int P = c->nranks;
  int layers = g.mz;
  if (!g.p3m) {
    // only the lower part of the mesh is occupied (particles live in [0, box]); cut that part
    const double occupied = (double)c->prm.box[2] / (double)c->prm.H;
    int l = (int)std::ceil(occupied / (double)(1 << g.tile_shift));
    if (l < layers) layers = l;
    if (layers < P) layers = g.mz;
  }
  g.nranks = P, g.rank = c->rank;
  for (int r = 0; r <= P3M_MAX_RANKS; ++r) g.cut[r] = (int)(((long long)(r < P ? r : P) * layers) / P);
  g.lay0 = g.cut[c->rank];
  g.lay1 = c->rank == P - 1 ? 0x7fffffff : g.cut[c->rank + 1];
  if (c->rank == 0) g.lay0 = -0x7fffffff;
}

// Balanced cuts.  p3m_set_particles hands every rank the SAME full particle set, so every rank computes the same
// per-layer weights on the host and the same cuts -- nothing is communicated.
//   PM-only:  weight of a binning layer = its particle count (mesh kernels are linear in the particles);
//   P3M:      + the short-range work of the layer: sum over its chaining cells of n_cell * (particles in the
//             27-cell neighbourhood), i.e. the pair evaluations the gather-form PP kernels will make; a particle
//             is charged as kPairsPerParticle pair evaluations of mesh-side work.
// cut[r] = first layer at which the cumulative weight reaches r/P of the total, kept strictly increasing so that
// every rank owns at least one layer.  A single layer is never split, so a structure thinner than one layer
// along z still lands on one rank.
template <typename T>
int dist_cuts_from_weights(p3m_ctx* c, const double* weight, int layers);

template <typename T>
int dist_balance_cuts(p3m_ctx* c, const float* pos, long long n, int units) {
  if (c->nranks <= 1 || n <= 0 || c->tune.static_cuts) return 0;
  Geom<T>& g = Sel<T>::g(c);
  const int P = c->nranks;
  set_cuts<T>(c);  // geometric cuts: defines the number of layers being cut
  const int layers = g.cut[P];
  // mesh-side work of one particle in units of one pair evaluation (sort + deposit + gather + integrate ~0.15-0.5
  // ns vs ~0.5-1 ps per pair on B200).  A caller that re-uploads all particles every step (bench.py's e2e loop)
  // is transfer-bound instead and wants a much larger value: P3M_TUNE_PARTICLE_WEIGHT overrides it.
  const double kPairsPerParticle = c->tune.particle_weight;
  const double h[3] = {g.p3m ? (double)g.hcx : (double)(1 << g.tile_shift),
                       g.p3m ? (double)g.hcy : (double)(1 << g.tile_shift),
                       g.p3m ? (double)g.hcz : (double)(1 << g.tile_shift)};
  const double u = units == P3M_UNITS_ORIGINAL ? 1.0 / (double)c->prm.H : 1.0;
  const double inv[3] = {u / h[0], u / h[1], u / h[2]};
  std::vector<double> weight((size_t)layers, 0.0);
  auto clampi = [](int v, int hi) { return v < 0 ? 0 : (v >= hi ? hi - 1 : v); };
  if (!g.p3m || c->tune.count_cuts) {
    for (long long i = 0; i < n; ++i)
      weight[(size_t)clampi((int)std::floor((double)pos[3 * i + 2] * inv[2]), layers)] += 1.0;
  } else {
    const int mx = g.mx, my = g.my;
    const size_t plane = (size_t)mx * my, cells = plane * (size_t)layers;
    std::vector<unsigned> cnt(cells, 0u), sx(cells), sy(cells);
    for (long long i = 0; i < n; ++i) {
      const int x = clampi((int)std::floor((double)pos[3 * i] * inv[0]), mx);
      const int y = clampi((int)std::floor((double)pos[3 * i + 1] * inv[1]), my);
      const int z = clampi((int)std::floor((double)pos[3 * i + 2] * inv[2]), layers);
      cnt[(size_t)z * plane + (size_t)y * mx + x]++;
    }
    // separable 3 x 3 x 3 box sum (non-periodic, like the chaining mesh)
    for (size_t r = 0; r < cells; r += mx)
      for (int x = 0; x < mx; ++x)
        sx[r + x] = cnt[r + x] + (x > 0 ? cnt[r + x - 1] : 0u) + (x + 1 < mx ? cnt[r + x + 1] : 0u);
    for (int z = 0; z < layers; ++z)
      for (int y = 0; y < my; ++y) {
        const size_t r = (size_t)z * plane + (size_t)y * mx;
        for (int x = 0; x < mx; ++x)
          sy[r + x] = sx[r + x] + (y > 0 ? sx[r + x - mx] : 0u) + (y + 1 < my ? sx[r + x + mx] : 0u);
      }
    for (int z = 0; z < layers; ++z) {
      double w = 0.0;
      for (size_t k = 0; k < plane; ++k) {
        const size_t i = (size_t)z * plane + k;
        if (!cnt[i]) continue;
        const double s27 = (double)sy[i] + (z > 0 ? (double)sy[i - plane] : 0.0) + (z + 1 < layers ? (double)sy[i + plane] : 0.0);
        w += (double)cnt[i] * (kPairsPerParticle + s27);
      }
      weight[(size_t)z] = w;
    }
  }
  return dist_cuts_from_weights<T>(c, weight.data(), layers);
}

// cut[r] = first layer at which the cumulative weight reaches r/P of the total (see dist_balance_cuts)
template <typename T>
int dist_cuts_from_weights(p3m_ctx* c, const double* weight, int layers) {
  Geom<T>& g = Sel<T>::g(c);
  const int P = c->nranks;
  double total = 0.0;
  for (int z = 0; z < layers; ++z) total += weight[z];
  int cut[P3M_MAX_RANKS + 1];
  cut[0] = 0;
  double cum = 0.0;
  int l = 0;
  for (int r = 1; r < P; ++r) {
    const double want = total * r / P;
    while (l < layers && cum + weight[l] <= want) cum += weight[l++];
    // the layer holding the quantile goes to whichever side leaves the smaller excess
    int k = l;
    if (l < layers && want - cum > cum + weight[l] - want) k = l + 1;
    if (k <= cut[r - 1]) k = cut[r - 1] + 1;
    if (k > layers - (P - r)) k = layers - (P - r);
    cut[r] = k;
  }
  for (int r = P; r <= P3M_MAX_RANKS; ++r) cut[r] = layers;
  for (int r = 0; r <= P3M_MAX_RANKS; ++r) g.cut[r] = cut[r];
  g.lay0 = c->rank == 0 ? -0x7fffffff : g.cut[c->rank];
  g.lay1 = c->rank == P - 1 ? 0x7fffffff : g.cut[c->rank + 1];
  return slab_replan<T>(c);
}

void dist_set_cuts(p3m_ctx* c) {
  if (c->f64) set_cuts<double>(c); else set_cuts<float>(c);
}

int comm_unique_id(void* out) {
  ncclUniqueId id;
  P3M_NCCL(ncclGetUniqueId(&id));
  memcpy(out, &id, sizeof(id));
  return 0;
}

int dist_init(p3m_ctx* c, const void* unique_id, int rank, int nranks) {
  if (nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks)
    return fail(P3M_EINVAL, "p3m_create_dist: rank %d of %d (1..8 ranks supported)", rank, nranks);
  c->rank = rank, c->nranks = nranks;
  if (c->f64) set_cuts<double>(c); else set_cuts<float>(c);
  const int layers = c->f64 ? c->g64.cut[nranks] : c->g32.cut[nranks];
  if (layers < nranks) return fail(P3M_EINVAL, "only %d binning layers along z for %d ranks", layers, nranks);
  if (nranks == 1) return 0;
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm;
  P3M_NCCL(ncclCommInitRank(&comm, nranks, id, rank));
  c->nccl_comm = comm;
  P3M_CUDA(cudaMalloc((void**)&c->dist_counts, sizeof(int) * kCntTotal));
  P3M_CUDA(cudaMemset(c->dist_counts, 0, sizeof(int) * kCntTotal));
  P3M_CUDA(cudaMallocHost((void**)&c->dist_counts_host, sizeof(int) * kCntTotal));
  return 0;
}

void dist_destroy(p3m_ctx* c) {
  if (c->nccl_comm) ncclCommDestroy((ncclComm_t)c->nccl_comm);
  c->nccl_comm = nullptr;
  if (c->dist_counts) cudaFree(c->dist_counts);
  if (c->dist_counts_host) cudaFreeHost(c->dist_counts_host);
  c->dist_counts = nullptr, c->dist_counts_host = nullptr;
}

int dist_allreduce(p3m_ctx* c, void* buf, size_t count, int kind) {
  if (c->nranks <= 1) return 0;
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  if (kind == 0)
    P3M_NCCL(ncclAllReduce(buf, buf, count, ncclInt, ncclMax, comm, c->stream));
  else
    P3M_NCCL(ncclAllReduce(buf, buf, count, ncclDouble, ncclSum, comm, c->stream));
  c->launches++;
  return 0;
}

// small host-side integer max over the ranks (blocking)
int dist_allreduce_host_imax(p3m_ctx* c, int* v, int count) {
  if (c->nranks <= 1) return 0;
  int* d = c->dist_counts + 120;
  P3M_CUDA(cudaMemcpyAsync(d, v, sizeof(int) * count, cudaMemcpyHostToDevice, c->stream));
  P3M_TRY(dist_allreduce(c, d, (size_t)count, 0));
  P3M_CUDA(cudaMemcpyAsync(v, d, sizeof(int) * count, cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

template <typename T>
int dist_allreduce_density(p3m_ctx* c) {
  if (c->nranks <= 1) return 0;
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  phase_begin(c, PH_COMM);
  P3M_NCCL(ncclAllReduce(s.density, s.density, (size_t)g.M, sizeof(T) == 8 ? ncclDouble : ncclFloat, ncclSum,
                         (ncclComm_t)c->nccl_comm, c->stream));
  c->launches++;
  phase_end(c, PH_COMM);
  return 0;
}

// ---- migration ----------------------------------------------------------------------------------------
template <typename T>
__global__ void k_dest(const V4<T>* __restrict__ posm, long long n, Geom<T> g, uint32_t* __restrict__ dest,
                       uint32_t* __restrict__ slots) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V4<T> p = posm[i];
  int cx, cy, cz;
  bool inside;
  bin_cell(g, p.x, p.y, p.z, cx, cy, cz, inside);
  dest[i] = (uint32_t)layer_owner(g, cz);
  slots[i] = (uint32_t)i;
}

__global__ void k_seg_start(const uint32_t* __restrict__ dest_sorted, long long n, int P, int cap,
                            int* __restrict__ cnt) {
  const int d = threadIdx.x;
  if (d > P) return;
  long long lo = 0, hi = n;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (dest_sorted[mid] < (uint32_t)d) lo = mid + 1; else hi = mid;
  }
  cnt[kCntSeg + d] = (int)lo;
  __syncthreads();
  if (d < P) cnt[kCntMine + d] = cnt[kCntSeg + d + 1] - cnt[kCntSeg + d];
  if (d == P) cnt[kCntMine + P] = cap;  // travels with the counts: capacities differ between ranks
}

template <typename T>
int dist_migrate(p3m_ctx* c, bool exchange) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const int P = c->nranks, me = c->rank;
  const long long n = c->n;
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  phase_begin(c, PH_COMM);
  uint32_t* dest = reinterpret_cast<uint32_t*>(s.keys);
  uint32_t* dest_sorted = reinterpret_cast<uint32_t*>(s.keys_alt);
  int bits = 1;
  while ((1 << bits) < P) ++bits;
  if (n > 0) {
    k_dest<T><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(s.posm, n, g, dest, s.slots);
    P3M_LAUNCH_CHECK(c);
    size_t tmp = s.cub_tmp_bytes;
    P3M_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_tmp, tmp, dest, dest_sorted, s.slots, s.slots_alt, (int)n, 0,
                                             bits, c->stream));
    c->launches += 3;
  }
  k_seg_start<<<1, 32, 0, c->stream>>>(dest_sorted, n, P, (int)(c->cap < 0x7fffffffLL ? c->cap : 0x7fffffffLL),
                                       c->dist_counts);
  P3M_LAUNCH_CHECK(c);
  if (exchange) {
    P3M_NCCL(ncclAllGather(c->dist_counts + kCntMine, c->dist_counts + kCntAll, P + 1, ncclInt, comm, c->stream));
    c->launches++;
  }
  P3M_CUDA(cudaMemcpyAsync(c->dist_counts_host, c->dist_counts, sizeof(int) * kCntTotal, cudaMemcpyDeviceToHost,
                           c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  const int* h = c->dist_counts_host;
  const int* seg = h + kCntSeg;
  long long n_new = h[kCntMine + me];
  if (exchange) {
    // every rank evaluates EVERY rank's capacity from the same all-gathered matrix, so that all of
    // them fail together instead of one leaving the others waiting in a collective
    for (int dst = 0; dst < P; ++dst) {
      long long tot = 0;
      for (int src = 0; src < P; ++src) tot += h[kCntAll + src * (P + 1) + dst];
      const long long cap_dst = h[kCntAll + dst * (P + 1) + P];
      if (tot > cap_dst)
        return fail(P3M_ERANGE, "rank %d would hold %lld particles, capacity %lld", dst, tot, cap_dst);
      if (dst == me) n_new = tot;
    }
  }
  if (n_new > c->cap)
    return fail(P3M_ERANGE, "rank %d would hold %lld particles, capacity %lld", me, n_new, c->cap);
  if (n > 0) {
    k_permute<T><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(s.slots_alt, n, s.posm, s.vel, s.id,
                                                                    s.posm_alt, s.vel_alt, s.id_alt);
    P3M_LAUNCH_CHECK(c);
  }
  // stayers to the front of the primary arrays, arrivals appended behind them in source-rank order
  const long long keep = h[kCntMine + me];
  // the keys the stayers are sorted by travel with them (incremental re-sort, incsort.cu)
  if (exchange) P3M_TRY(migrate_sorted_keys<T>(c, s.slots_alt + seg[me], keep, n_new));
  else c->order_valid = false;
  if (keep > 0) {
    P3M_CUDA(cudaMemcpyAsync(s.posm, s.posm_alt + seg[me], sizeof(V4<T>) * keep, cudaMemcpyDeviceToDevice, c->stream));
    P3M_CUDA(cudaMemcpyAsync(s.vel, s.vel_alt + seg[me], sizeof(V4<T>) * keep, cudaMemcpyDeviceToDevice, c->stream));
    P3M_CUDA(cudaMemcpyAsync(s.id, s.id_alt + seg[me], sizeof(int) * keep, cudaMemcpyDeviceToDevice, c->stream));
  }
  if (exchange) {
    long long off = keep;
    long long sent = 0;
    for (int p = 0; p < P; ++p)
      if (p != me) sent += h[kCntMine + p];
    c->stat_migrated = (double)sent;
    c->stat_mig_bytes = (double)sent * (double)(2 * sizeof(V4<T>) + sizeof(int));
    P3M_NCCL(ncclGroupStart());
    for (int p = 0; p < P; ++p) {
      if (p == me) continue;
      const long long ns = h[kCntMine + p], nr = h[kCntAll + p * (P + 1) + me];
      if (ns > 0) {
        P3M_NCCL(ncclSend(s.posm_alt + seg[p], sizeof(V4<T>) * ns, ncclChar, p, comm, c->stream));
        P3M_NCCL(ncclSend(s.vel_alt + seg[p], sizeof(V4<T>) * ns, ncclChar, p, comm, c->stream));
        P3M_NCCL(ncclSend(s.id_alt + seg[p], sizeof(int) * ns, ncclChar, p, comm, c->stream));
      }
      if (nr > 0) {
        P3M_NCCL(ncclRecv(s.posm + off, sizeof(V4<T>) * nr, ncclChar, p, comm, c->stream));
        P3M_NCCL(ncclRecv(s.vel + off, sizeof(V4<T>) * nr, ncclChar, p, comm, c->stream));
        P3M_NCCL(ncclRecv(s.id + off, sizeof(int) * nr, ncclChar, p, comm, c->stream));
        off += nr;
      }
    }
    P3M_NCCL(ncclGroupEnd());
    c->launches++;
  }
  c->n = n_new;
  c->sorted = false;
  c->have_acc = false;  // accelerations do not travel
  phase_end(c, PH_COMM);
  return 0;
}

// ---- ghosts -------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_ghost_pack(const V4<T>* __restrict__ posm, const int* __restrict__ id, long long n, Geom<T> g,
                             V4<T>* __restrict__ pack_pos, int* __restrict__ pack_id, long long half,
                             int* __restrict__ counters) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V4<T> p = posm[i];
  int cx, cy, cz;
  bool inside;
  bin_cell(g, p.x, p.y, p.z, cx, cy, cz, inside);
  if (g.rank > 0 && cz == g.cut[g.rank]) {  // my lowest layer: ghost of rank-1
    const int k = atomicAdd(&counters[0], 1);
    pack_pos[k] = p, pack_id[k] = id[i];
  }
  if (g.rank < g.nranks - 1 && cz == g.cut[g.rank + 1] - 1) {  // my highest layer: ghost of rank+1
    const int k = atomicAdd(&counters[1], 1);
    pack_pos[half + k] = p, pack_id[half + k] = id[i];
  }
}

template <typename T>
__global__ void k_permute_ghost(const uint32_t* __restrict__ slots, long long n, const V4<T>* __restrict__ pos,
                                const int* __restrict__ id, V4<T>* __restrict__ pos_o, int* __restrict__ id_o) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = slots[i];
  pos_o[i] = pos[s];
  id_o[i] = id[s];
}

template <typename T>
static int dev_realloc(T** p, size_t count) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  P3M_CUDA(cudaMalloc((void**)p, sizeof(T) * (count ? count : 1)));
  return 0;
}

template <typename T>
int dist_ghosts(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  s.n_ghost = 0;
  if (!g.p3m || c->nranks <= 1) return 0;
  const int P = c->nranks, me = c->rank;
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  const long long ncells = 1LL << (3 * g.mbits);
  if (s.ghost_cap < 2 * c->cap) {
    // worst case: one layer holds (almost) every particle of a rank and goes to both neighbours, and
    // both neighbours' boundary layers arrive here: 2 * cap on each side
    P3M_TRY(dev_realloc(&s.gposm, 2 * c->cap));
    P3M_TRY(dev_realloc(&s.gposm_alt, 2 * c->cap));
    P3M_TRY(dev_realloc(&s.gid, 2 * c->cap));
    P3M_TRY(dev_realloc(&s.gid_alt, 2 * c->cap));
    P3M_TRY(dev_realloc(&s.gcell_start, ncells + 2));
    P3M_TRY(dev_realloc(&s.gaabb, 2 * (2 * c->cap / kPPSub + 8)));
    s.ghost_cap = 2 * c->cap;
  }
  phase_begin(c, PH_COMM);
  const long long n = c->n, half = s.ghost_cap / 2;
  int* counters = c->dist_counts + kCntGhost;
  {
    // [2] = how many ghosts this rank can receive; capacities differ between ranks after p3m_set_particles_ids
    const long long rc = s.ghost_cap < 0x7fffffffLL ? s.ghost_cap : 0x7fffffffLL;
    const int init[3] = {0, 0, (int)rc};
    P3M_CUDA(cudaMemcpyAsync(counters, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  }
  if (n > 0) {
    k_ghost_pack<T><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(s.posm, s.id, n, g, s.gposm_alt, s.gid_alt,
                                                                       half, counters);
    P3M_LAUNCH_CHECK(c);
  }
  P3M_NCCL(ncclAllGather(counters, c->dist_counts + kCntGhostAll, 3, ncclInt, comm, c->stream));
  P3M_CUDA(cudaMemcpyAsync(c->dist_counts_host + kCntGhostAll, c->dist_counts + kCntGhostAll, sizeof(int) * 3 * P,
                           cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  const int* h = c->dist_counts_host + kCntGhostAll;
  const long long send_lo = h[3 * me], send_hi = h[3 * me + 1];
  const long long recv_lo = me > 0 ? h[3 * (me - 1) + 1] : 0;      // rank-1's highest layer
  const long long recv_hi = me < P - 1 ? h[3 * (me + 1)] : 0;      // rank+1's lowest layer
  for (int r = 0; r < P; ++r) {  // every rank reaches the same verdict about every rank: nobody is left in a collective
    const long long in = (r > 0 ? h[3 * (r - 1) + 1] : 0) + (r < P - 1 ? h[3 * (r + 1)] : 0);
    if (in > h[3 * r + 2])
      return fail(P3M_ERANGE, "rank %d would receive %lld ghost particles, capacity %d", r, in, h[3 * r + 2]);
  }
  P3M_NCCL(ncclGroupStart());
  if (me > 0) {
    if (send_lo > 0) {
      P3M_NCCL(ncclSend(s.gposm_alt, sizeof(V4<T>) * send_lo, ncclChar, me - 1, comm, c->stream));
      P3M_NCCL(ncclSend(s.gid_alt, sizeof(int) * send_lo, ncclChar, me - 1, comm, c->stream));
    }
    if (recv_lo > 0) {
      P3M_NCCL(ncclRecv(s.gposm, sizeof(V4<T>) * recv_lo, ncclChar, me - 1, comm, c->stream));
      P3M_NCCL(ncclRecv(s.gid, sizeof(int) * recv_lo, ncclChar, me - 1, comm, c->stream));
    }
  }
  if (me < P - 1) {
    if (send_hi > 0) {
      P3M_NCCL(ncclSend(s.gposm_alt + half, sizeof(V4<T>) * send_hi, ncclChar, me + 1, comm, c->stream));
      P3M_NCCL(ncclSend(s.gid_alt + half, sizeof(int) * send_hi, ncclChar, me + 1, comm, c->stream));
    }
    if (recv_hi > 0) {
      P3M_NCCL(ncclRecv(s.gposm + recv_lo, sizeof(V4<T>) * recv_hi, ncclChar, me + 1, comm, c->stream));
      P3M_NCCL(ncclRecv(s.gid + recv_lo, sizeof(int) * recv_hi, ncclChar, me + 1, comm, c->stream));
    }
  }
  P3M_NCCL(ncclGroupEnd());
  c->launches += 2;
  const long long ng = recv_lo + recv_hi;
  s.n_ghost = ng;
  c->stat_ghost_bytes = (double)(send_lo + send_hi) * (double)(sizeof(V4<T>) + sizeof(int));
  // sort the ghosts by the same (cell, sub-cell, id) key and index them by cell
  if (ng > 0) {
    const unsigned blocks = (unsigned)((ng + 255) / 256);
    k_keys<T, uint64_t><<<blocks, 256, 0, c->stream>>>(s.gposm, s.gid, ng, g, s.keys, s.slots, s.flags);
    P3M_LAUNCH_CHECK(c);
    size_t tmp = s.cub_tmp_bytes;
    const int keybits = g.idbits + 3 * g.sbits + 3 * g.mbits;
    P3M_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_tmp, tmp, s.keys, s.keys_alt, s.slots, s.slots_alt, (int)ng, 0,
                                             keybits, c->stream));
    c->launches += (keybits + 7) / 8 + 1;
    k_permute_ghost<T><<<blocks, 256, 0, c->stream>>>(s.slots_alt, ng, s.gposm, s.gid, s.gposm_alt, s.gid_alt);
    P3M_LAUNCH_CHECK(c);
    std::swap(s.gposm, s.gposm_alt);
    std::swap(s.gid, s.gid_alt);
    const long long tiles = (ng + kPPSub - 1) / kPPSub;
    k_tile_aabb<T><<<(unsigned)((tiles * 32 + 255) / 256), 256, 0, c->stream>>>(s.gposm, ng, s.gaabb);
    P3M_LAUNCH_CHECK(c);
  }
  k_cell_start<uint64_t><<<(unsigned)((ncells + 1 + 255) / 256), 256, 0, c->stream>>>(s.keys_alt, ng, g.idbits + 3 * g.sbits,
                                                                           ncells, s.gcell_start);
  P3M_LAUNCH_CHECK(c);
  phase_end(c, PH_COMM);
  return 0;
}

template int dist_cuts_from_weights<float>(p3m_ctx*, const double*, int);
template int dist_cuts_from_weights<double>(p3m_ctx*, const double*, int);
template int dist_balance_cuts<float>(p3m_ctx*, const float*, long long, int);
template int dist_balance_cuts<double>(p3m_ctx*, const float*, long long, int);
template int dist_migrate<float>(p3m_ctx*, bool);
template int dist_migrate<double>(p3m_ctx*, bool);
template int dist_ghosts<float>(p3m_ctx*);
template int dist_ghosts<double>(p3m_ctx*);
template int dist_allreduce_density<float>(p3m_ctx*);
template int dist_allreduce_density<double>(p3m_ctx*);

}  // namespace p3m
