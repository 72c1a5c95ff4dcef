// incsort.cu -- A0, the per-step re-sort: merge the MOVERS instead of sorting everything again.
//
// The particle arrays stay in key order from one step to the next (binsort.cu) and a particle moves less than
// a cell per step, so after a drift most particles still carry the key they were sorted by.  With the 32-bit
// key (Morton(cell), sub-cell; no id) the re-sort is therefore:
//
//   k_inc_keys           new key of every particle; "mover" = new key != the key it was sorted by (skeys, which
//                        is kept from the previous sort; particles that arrived by migration carry 0xffffffff).
//                        The non-movers ("stayers") are a subsequence of a sorted sequence with unchanged keys:
//                        still sorted.  Writes a mover bit mask and per-CTA mover counts.
//   k_excl_scan          exclusive scan of the per-CTA counts, total m (one CTA)
//   k_inc_split          mover prefix per 32-particle word (moff), compacted movers (key, old index) in index order
//   k_rx_hist/scan/scatter   hand-written stable LSD radix sort (digits of <= 10 bits) of the m movers only
//   k_inc_place_movers   sorted mover j lands at  j + #stayers before it   (binary search in the OLD key array,
//                        which is sorted over all old indices; stayers before an old index = index - movers
//                        before it, from moff + the mask word)
//   k_inc_place_stayers  stayer i lands at  (i - movers before i) + #movers before it  (binary search in the
//                        sorted movers, narrowed per CTA to the bracket of its first and last old key)
//
// Ties are ordered by the old index on both sides, which is exactly what a STABLE sort of the whole array by
// the new key would do: the result is identical, element for element, to the full radix sort
// (tests/test_gpu_parity.py::test_incremental_sort_equals_full_sort).  m = 0: nothing is moved at all.
// The reference has no counterpart (linked-list insertion per step, source/chainingMesh.cpp:20-58).
#include "ctx.cuh"
#include "sort_kernels.cuh"

namespace p3m {
namespace {

constexpr int kIncBlock = 1024;  // particles per CTA of the key / split / place kernels = 32 mask words
constexpr int kRxTile = 4096;    // movers per CTA of a radix pass (8 warps x 16 rounds x 32 lanes)

template <typename T>
__global__ void __launch_bounds__(256)
k_inc_keys(const V4<T>* __restrict__ posm, long long n, Geom<T> g, const uint32_t* __restrict__ skeys,
           uint32_t* __restrict__ keys, uint32_t* __restrict__ mask, int* __restrict__ blockcnt,
           int* __restrict__ flags, int* __restrict__ zmax) {
  const long long base = blockIdx.x * (long long)kIncBlock;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int movers = 0, zloc = -1;
#pragma unroll
  for (int r = 0; r < kIncBlock / 256; ++r) {
    const long long i = base + r * 256 + threadIdx.x;
    bool mv = false;
    if (i < n) {
      bool inside;
      const V4<T> pp = posm[i];
      const uint32_t k = (uint32_t)cell_key(g, pp, inside);
      zloc = max(zloc, plane_of(pp));
      if (!inside) flags[1] = 1;
      keys[i] = k;
      mv = k != skeys[i];
    }
    const unsigned b = __ballot_sync(0xffffffffu, mv);
    if (lane == 0 && base + r * 256 + wid * 32 < n) mask[(base >> 5) + r * 8 + wid] = b;
    movers += __popc(b);  // identical on every lane of the warp
  }
  __shared__ int red[8];
  if (lane == 0) red[wid] = movers;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    blockcnt[blockIdx.x] = t;
  }
  if (zmax) block_zmax(zloc, zmax);
}

// in-place exclusive scan of `count` ints by ONE CTA of 1024 threads, 8 consecutive entries per thread and
// iteration; total[0] = sum
__global__ void __launch_bounds__(1024) k_excl_scan(int* __restrict__ data, int count, int* __restrict__ total) {
  constexpr int ITEMS = 8;
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < count; base += 1024 * ITEMS) {
    const int first = base + threadIdx.x * ITEMS;
    int v[ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      v[k] = first + k < count ? data[first + k] : 0;
      sum += v[k];
    }
    int x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    if (wid == 0) {
      const int w = wsum[lane];
      int s = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      wsum[lane] = s - w;
    }
    __syncthreads();
    int run = carry_s + wsum[wid] + x - sum;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      if (first + k < count) data[first + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = carry_s;
}

__global__ void __launch_bounds__(256)
k_inc_split(const uint32_t* __restrict__ keys, long long n, const uint32_t* __restrict__ mask,
            const int* __restrict__ blockoff, int* __restrict__ moff, uint32_t* __restrict__ mv_key,
            uint32_t* __restrict__ mv_idx) {
  __shared__ int woff[32];
  __shared__ unsigned wmask[32];
  const long long base = blockIdx.x * (long long)kIncBlock;
  const long long w0 = base >> 5, nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x < 32) {
    const long long w = w0 + threadIdx.x;
    const unsigned mk = w < nwords ? mask[w] : 0u;
    const int c = __popc(mk);
    int x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    const int excl = blockoff[blockIdx.x] + x - c;
    woff[threadIdx.x] = excl;
    wmask[threadIdx.x] = mk;
    if (w < nwords) moff[w] = excl;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kIncBlock / 256; ++r) {
    const long long i = base + r * 256 + threadIdx.x;
    const int wl = r * 8 + wid;
    const unsigned mk = wmask[wl];
    if (i < n && ((mk >> lane) & 1u)) {
      const int j = woff[wl] + __popc(mk & ((1u << lane) - 1u));
      mv_key[j] = keys[i];
      mv_idx[j] = (uint32_t)i;
    }
  }
}

// ---- stable LSD radix sort of the movers: (key, old index) pairs, `rb` <= 10 bits per pass -------------
constexpr int kRxMaxDigits = 1024;

__global__ void __launch_bounds__(256)
k_rx_hist(const uint32_t* __restrict__ key, int m, int shift, int rb, int nblk, int* __restrict__ hist) {
  __shared__ int h[kRxMaxDigits];
  const int nd = 1 << rb;
  for (int d = threadIdx.x; d < nd; d += 256) h[d] = 0;
  __syncthreads();
  const int base = blockIdx.x * kRxTile;
  for (int k = threadIdx.x; k < kRxTile; k += 256) {
    const int i = base + k;
    if (i < m) atomicAdd(&h[(key[i] >> shift) & (uint32_t)(nd - 1)], 1);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < nd; d += 256) hist[d * nblk + blockIdx.x] = h[d];  // digit-major rows
}

// one CTA per digit: exclusive scan of its row of per-CTA counts in place, row total to rowtot[digit]
__global__ void __launch_bounds__(256) k_rx_scan_rows(int* __restrict__ hist, int nblk, int* __restrict__ rowtot) {
  __shared__ int wsum[8];
  __shared__ int carry_s;
  int* row = hist + (size_t)blockIdx.x * nblk;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 256) {
    const int i = base + threadIdx.x;
    const int v = i < nblk ? row[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    int before = carry_s;
    for (int w = 0; w < wid; ++w) before += wsum[w];
    if (i < nblk) row[i] = before + x - v;
    __syncthreads();
    if (threadIdx.x == 255) carry_s = before + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) rowtot[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(256)
k_rx_scatter(const uint32_t* __restrict__ key, const uint32_t* __restrict__ idx, int m, int shift, int rb, int nblk,
             const int* __restrict__ hist, const int* __restrict__ rowtot, uint32_t* __restrict__ key_o,
             uint32_t* __restrict__ idx_o) {
  __shared__ int wc[8][kRxMaxDigits];
  __shared__ int dbase[kRxMaxDigits];
  __shared__ int chunk[8];
  const int nd = 1 << rb;
  const uint32_t dmask = (uint32_t)(nd - 1);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int w = 0; w < 8; ++w)
    for (int d = threadIdx.x; d < nd; d += 256) wc[w][d] = 0;
  // start of every digit in the output = exclusive scan of the row totals (nd <= 1024: 4 per thread)
  {
    const int per = nd / 256 > 0 ? nd / 256 : 1;
    int v[4] = {0, 0, 0, 0}, sum = 0;
    for (int k = 0; k < per; ++k) {
      const int d = threadIdx.x * per + k;
      v[k] = d < nd ? rowtot[d] : 0;
      sum += v[k];
    }
    int x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) chunk[wid] = x;
    __syncthreads();
    int run = x - sum;
    for (int w = 0; w < wid; ++w) run += chunk[w];
    for (int k = 0; k < per; ++k) {
      const int d = threadIdx.x * per + k;
      if (d < nd) dbase[d] = run;
      run += v[k];
    }
  }
  __syncthreads();
  // warp w owns elements [w * 512, (w + 1) * 512) of the tile, in 16 rounds of 32 consecutive ones
  const int wbase = blockIdx.x * kRxTile + wid * (kRxTile / 8);
  uint32_t kk[16], ii[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int i = wbase + r * 32 + lane;
    const bool ok = i < m;
    kk[r] = ok ? key[i] : 0u;
    ii[r] = ok ? idx[i] : 0u;
    const int d = ok ? (int)((kk[r] >> shift) & dmask) : nd;
    const unsigned mm = __match_any_sync(0xffffffffu, d);
    if (ok && lane == __ffs(mm) - 1) wc[wid][d] += __popc(mm);
    __syncwarp();
  }
  __syncthreads();
  // start of (digit, warp) in the output: digit base + CTAs before this one + warps before this one
  for (int d = threadIdx.x; d < nd; d += 256) {
    int run = dbase[d] + hist[d * nblk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int c = wc[w][d];
      wc[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int i = wbase + r * 32 + lane;
    const bool ok = i < m;
    const int d = ok ? (int)((kk[r] >> shift) & dmask) : nd;
    const unsigned mm = __match_any_sync(0xffffffffu, d);
    if (ok) {
      const int pos = wc[wid][d] + __popc(mm & ((1u << lane) - 1u));
      key_o[pos] = kk[r];
      idx_o[pos] = ii[r];
    }
    __syncwarp();
    if (ok && lane == __ffs(mm) - 1) wc[wid][d] += __popc(mm);
    __syncwarp();
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_inc_place_movers(const uint32_t* __restrict__ mv_key, const uint32_t* __restrict__ mv_idx, int m,
                   const uint32_t* __restrict__ skeys, long long n, const uint32_t* __restrict__ mask,
                   const int* __restrict__ moff, const V4<T>* __restrict__ posm, const V4<T>* __restrict__ vel,
                   const int* __restrict__ id, V4<T>* __restrict__ posm_o, V4<T>* __restrict__ vel_o,
                   int* __restrict__ id_o, uint32_t* __restrict__ skeys_o) {
  __shared__ long long bracket[2];
  const int j0 = blockIdx.x * blockDim.x;
  const int j = j0 + threadIdx.x;
  // first old index i with (skeys[i], i) >= (k, idx): everything before it sorts before this mover
  auto first_not_before = [&](uint32_t k, uint32_t idx, long long lo, long long hi) {
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      const uint32_t sk = skeys[mid];
      if (sk < k || (sk == k && mid < (long long)idx)) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  if (threadIdx.x < 2) {
    // the sorted movers of this CTA are monotone in (key, index): its first and last one bracket the others
    const int jj = threadIdx.x == 0 ? j0 : min(j0 + (int)blockDim.x, m) - 1;
    bracket[threadIdx.x] = first_not_before(mv_key[jj], mv_idx[jj], 0, n);
  }
  __syncthreads();
  if (j >= m) return;
  const uint32_t k = mv_key[j], idx = mv_idx[j];
  const long long lo = first_not_before(k, idx, bracket[0], bracket[1]);
  const long long movers_before = lo < n ? (long long)moff[lo >> 5] + __popc(mask[lo >> 5] & ((1u << (lo & 31)) - 1u)) : m;
  const long long dest = (long long)j + (lo - movers_before);
  posm_o[dest] = posm[idx];
  vel_o[dest] = vel[idx];
  id_o[dest] = id[idx];
  skeys_o[dest] = k;
}

constexpr int kStayStage = 1024;  // sorted movers of a CTA's bracket staged in shared memory (8 KB)

template <typename T>
__global__ void __launch_bounds__(256)
k_inc_place_stayers(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ skeys, long long n,
                    const uint32_t* __restrict__ mask, const int* __restrict__ moff,
                    const uint32_t* __restrict__ mv_key, const uint32_t* __restrict__ mv_idx, int m,
                    const V4<T>* __restrict__ posm, const V4<T>* __restrict__ vel, const int* __restrict__ id,
                    V4<T>* __restrict__ posm_o, V4<T>* __restrict__ vel_o, int* __restrict__ id_o,
                    uint32_t* __restrict__ skeys_o) {
  __shared__ int bracket[2];
  __shared__ uint32_t sk_key[kStayStage], sk_idx[kStayStage];
  const long long base = blockIdx.x * (long long)kIncBlock;
  const long long last = (base + kIncBlock < n ? base + kIncBlock : n) - 1;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 2) {
    // the old keys are sorted over ALL old indices, so the first and the last one of this CTA bracket every
    // stayer in between (a stayer's key is its old key)
    const long long i = threadIdx.x == 0 ? base : last;
    const uint32_t k = skeys[i];
    int lo = 0, hi = m;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const uint32_t mk = mv_key[mid];
      if (mk < k || (mk == k && (long long)mv_idx[mid] < i)) lo = mid + 1; else hi = mid;
    }
    bracket[threadIdx.x] = lo;
  }
  __syncthreads();
  const int b0 = bracket[0], b1 = bracket[1];
  const bool staged = b1 - b0 <= kStayStage;
  if (staged)
    for (int k = threadIdx.x; k < b1 - b0; k += 256) sk_key[k] = mv_key[b0 + k], sk_idx[k] = mv_idx[b0 + k];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kIncBlock / 256; ++r) {
    const long long i = base + r * 256 + threadIdx.x;
    if (i >= n) continue;
    const unsigned mk = mask[i >> 5];
    if ((mk >> lane) & 1u) continue;  // movers are placed by k_inc_place_movers
    const uint32_t k = keys[i];
    // number of sorted movers that order before (k, i), searched inside the bracket
    int lo = 0, hi = b1 - b0;
    if (staged) {
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const uint32_t q = sk_key[mid];
        if (q < k || (q == k && (long long)sk_idx[mid] < i)) lo = mid + 1; else hi = mid;
      }
    } else {
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const uint32_t q = mv_key[b0 + mid];
        if (q < k || (q == k && (long long)mv_idx[b0 + mid] < i)) lo = mid + 1; else hi = mid;
      }
    }
    const long long dest = i - ((long long)moff[i >> 5] + __popc(mk & ((1u << lane) - 1u))) + b0 + lo;
    posm_o[dest] = posm[i];
    vel_o[dest] = vel[i];
    id_o[dest] = id[i];
    skeys_o[dest] = k;
  }
}

}  // namespace

// stayers of a migration keep their sorted-by key (gathered through the same slots as the particle arrays);
// arrivals get 0xffffffff, which makes every one of them a mover and keeps the old-key array sorted
__global__ void k_migrate_skeys(const uint32_t* __restrict__ slots, long long keep, long long n_new,
                                const uint32_t* __restrict__ skeys, uint32_t* __restrict__ skeys_o) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_new) return;
  skeys_o[i] = i < keep ? skeys[slots[i]] : 0xffffffffu;
}

template <typename T>
int migrate_sorted_keys(p3m_ctx* c, const uint32_t* stayer_slots, long long keep, long long n_new) {
  State<T>& s = Sel<T>::st(c);
  if (!c->order_valid) return 0;
  if (n_new > 0) {
    k_migrate_skeys<<<(unsigned)((n_new + 255) / 256), 256, 0, c->stream>>>(stayer_slots, keep, n_new, s.skeys, s.skeys_alt);
    P3M_LAUNCH_CHECK(c);
    std::swap(s.skeys, s.skeys_alt);
  }
  s.skeys_n = n_new;
  return 0;
}

// Returns with *done = true when the arrays are in key order again and s.skeys holds the sorted keys; false when
// the caller has to run the full sort (too many movers).
template <typename T>
int sort_incremental(p3m_ctx* c, int keybits, bool* done) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const long long n = c->n;
  *done = false;
  if (n <= 0) return 0;
  const long long scap = c->nranks > 1 ? 2 * c->cap : c->cap;
  uint32_t* keys32 = reinterpret_cast<uint32_t*>(s.keys);
  uint32_t* upper = keys32 + scap;                       // second half of the 64-bit key scratch
  const long long W = (c->cap + 31) / 32 + 1;
  uint32_t* mask = upper;
  int* moff = reinterpret_cast<int*>(upper + W);
  int* blockcnt = reinterpret_cast<int*>(upper + 2 * W);
  const int nb = (int)((n + kIncBlock - 1) / kIncBlock);
  k_inc_keys<T><<<nb, 256, 0, c->stream>>>(s.posm, n, g, s.skeys, keys32, mask, blockcnt, s.flags, s.inc_counts + 2);
  P3M_LAUNCH_CHECK(c);
  k_excl_scan<<<1, 1024, 0, c->stream>>>(blockcnt, nb, s.inc_counts);
  P3M_LAUNCH_CHECK(c);
  P3M_CUDA(cudaMemcpyAsync(s.inc_counts_host, s.inc_counts, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  const int m = s.inc_counts_host[0];
  s.inc_counts_host[3] = 1;  // [2] = highest occupied plane is fresh (bin_sort)
  c->stat_movers = (double)m;
  // many movers (e.g. the 16^3 sub-cell key of a warm P3M set): the full sort is cheaper -- measured break-even on
  // B200 at 2^24 particles: ~1/10 of the particles (merge 0.97 ms vs radix sort 0.89 ms at 11.5 % movers); do not
  // even try for the next few steps (the attempt costs the key pass)
  if ((long long)m * c->tune.inc_sort_den > n) {
    s.inc_backoff = 8;
    return 0;
  }
  *done = true;
  c->incr_sorts++;
  if (m == 0) return 0;                // nothing moved: arrays and keys stay as they are
  uint32_t* mvk[2] = {reinterpret_cast<uint32_t*>(s.keys_alt), reinterpret_cast<uint32_t*>(s.keys_alt) + scap};
  uint32_t* mvi[2] = {s.slots, s.slots_alt};
  k_inc_split<<<nb, 256, 0, c->stream>>>(keys32, n, mask, blockcnt, moff, mvk[0], mvi[0]);
  P3M_LAUNCH_CHECK(c);
  const int nblk = (m + kRxTile - 1) / kRxTile;
  const int passes = (keybits + 9) / 10;                 // digits of at most 10 bits
  const int rb = (keybits + passes - 1) / passes;
  int* rowtot = s.inc_hist + (size_t)kRxMaxDigits * (size_t)(c->cap / kRxTile + 2);
  int cur = 0;
  for (int shift = 0; shift < keybits; shift += rb) {
    k_rx_hist<<<nblk, 256, 0, c->stream>>>(mvk[cur], m, shift, rb, nblk, s.inc_hist);
    P3M_LAUNCH_CHECK(c);
    k_rx_scan_rows<<<1 << rb, 256, 0, c->stream>>>(s.inc_hist, nblk, rowtot);
    P3M_LAUNCH_CHECK(c);
    k_rx_scatter<<<nblk, 256, 0, c->stream>>>(mvk[cur], mvi[cur], m, shift, rb, nblk, s.inc_hist, rowtot, mvk[cur ^ 1],
                                              mvi[cur ^ 1]);
    P3M_LAUNCH_CHECK(c);
    cur ^= 1;
  }
  k_inc_place_movers<T><<<(m + 255) / 256, 256, 0, c->stream>>>(mvk[cur], mvi[cur], m, s.skeys, n, mask, moff, s.posm, s.vel,
                                                              s.id, s.posm_alt, s.vel_alt, s.id_alt, s.skeys_alt);
  P3M_LAUNCH_CHECK(c);
  k_inc_place_stayers<T><<<nb, 256, 0, c->stream>>>(keys32, s.skeys, n, mask, moff, mvk[cur], mvi[cur], m, s.posm, s.vel,
                                                   s.id, s.posm_alt, s.vel_alt, s.id_alt, s.skeys_alt);
  P3M_LAUNCH_CHECK(c);
  std::swap(s.posm, s.posm_alt);
  std::swap(s.vel, s.vel_alt);
  std::swap(s.id, s.id_alt);
  std::swap(s.skeys, s.skeys_alt);
  return 0;
}

template int sort_incremental<float>(p3m_ctx*, int, bool*);
template int sort_incremental<double>(p3m_ctx*, int, bool*);
template int migrate_sorted_keys<float>(p3m_ctx*, const uint32_t*, long long, long long);
template int migrate_sorted_keys<double>(p3m_ctx*, const uint32_t*, long long, long long);

}  // namespace p3m
