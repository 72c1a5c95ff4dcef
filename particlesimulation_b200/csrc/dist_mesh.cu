// dist_mesh.cu -- multi-GPU: the slab-decomposed mesh and the distributed FFT Poisson solve
// (nothing of this exists in the reference, which is single-process).
//
// FFT decomposition: rank r owns the mesh planes z in [r*nz/P, (r+1)*nz/P) of density and potential.
// The particle decomposition (dist.cu) cuts binning-cell layers, and particles only occupy the lower
// part of the mesh (box = half the mesh per axis in the reference's set-ups), so the planes a rank's
// particles touch are NOT the planes it owns.  Everything below is derived from the static layer cuts,
// identically on every rank, so no sizes are ever negotiated at run time:
//
//   deposit     -> dens_part : planes [den_z0, den_z0 + den_nz) touched by this rank's particles
//   slab_reduce_density     : every overlap (source planes x owner's FFT planes) travels with one grouped
//                             ncclSend/ncclRecv and is summed into the owner's slab (fixed order -> the
//                             result does not depend on arrival order)
//   slab_poisson            : batched 2-D R2C over the local planes -> pack -> all-to-all (grouped
//                             ncclSend/ncclRecv over NVLink) -> strided 1-D C2C along z on [kx, ky_local, kz]
//                             -> multiply by the locally stored share of the influence function ->
//                             inverse 1-D -> all-to-all back -> unpack -> batched 2-D C2R
//   slab_spread_potential   : the unwrapped planes [pot_z0, pot_z0 + pot_nz) every rank's gather needs
//                             (assignment stencil + finite-difference halo, periodic in z) are sent from
//                             their owners straight into pot_part
//
// Per GPU and solve the all-to-all moves 2 x 8 B x (nx/2+1) x ny x nz / P x (P-1)/P bytes; mesh memory is
// M/P per array, so a 1024^3 mesh on 8 GPUs holds 512 MiB per real slab.
#include <nccl.h>

#include <algorithm>
#include <cmath>

#include "ctx.cuh"

namespace p3m {

#define P3M_NCCL(expr)                                                                        \
  do {                                                                                        \
    ncclResult_t r__ = (expr);                                                                \
    if (r__ != ncclSuccess)                                                                   \
      return ::p3m::fail(P3M_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,              \
                         ncclGetErrorString(r__));                                            \
  } while (0)

namespace {

template <typename T>
ncclDataType_t nccl_real() { return sizeof(T) == 8 ? ncclDouble : ncclFloat; }

// [kx, ky, zl] -> P chunks [kx, ky_local, zl]   (forward) or back (inverse)
template <typename C, bool PACK>
__global__ void k_transpose_pack(C* __restrict__ planes, C* __restrict__ chunks, int nxh, int ny, int nyl,
                                 int nzl) {
  const long long total = (long long)nxh * ny * nzl;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kx = (int)(i % nxh);
    const long long q = i / nxh;
    const int ky = (int)(q % ny), zl = (int)(q / ny);
    const int d = ky / nyl, kyl = ky - d * nyl;
    const long long j = (long long)d * nxh * nyl * nzl + kx + (long long)nxh * (kyl + (long long)nyl * zl);
    if (PACK) chunks[j] = planes[i]; else planes[i] = chunks[j];
  }
}

template <typename C, typename T>
__global__ void k_multiply_t(C* __restrict__ spec, const T* __restrict__ table, long long count) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    C v = spec[i];
    const T g = table[i];
    v.x *= g, v.y *= g;
    spec[i] = v;
  }
}

template <typename T>
__global__ void k_add_planes(T* __restrict__ dst, const T* __restrict__ src, long long count) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] += src[i];
}

inline cufftResult exec_r2c(cufftHandle p, float* in, cufftComplex* out) { return cufftExecR2C(p, in, out); }
inline cufftResult exec_r2c(cufftHandle p, double* in, cufftDoubleComplex* out) { return cufftExecD2Z(p, in, out); }
inline cufftResult exec_c2r(cufftHandle p, cufftComplex* in, float* out) { return cufftExecC2R(p, in, out); }
inline cufftResult exec_c2r(cufftHandle p, cufftDoubleComplex* in, double* out) { return cufftExecZ2D(p, in, out); }
inline cufftResult exec_c2c(cufftHandle p, cufftComplex* d, int dir) { return cufftExecC2C(p, d, d, dir); }
inline cufftResult exec_c2c(cufftHandle p, cufftDoubleComplex* d, int dir) { return cufftExecZ2Z(p, d, d, dir); }

inline int grid_for(long long count, int sms) {
  long long b = (count + 255) / 256;
  const long long cap = (long long)sms * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

// ---- which planes a rank's FFT slab holds ------------------------------------------------------------------------
// The particles live in the lower half of the mesh along every axis (isolated boundary conditions by zero padding:
// mesh = 2 x box), so with contiguous FFT slabs the upper half of the ranks would own nothing but empty planes while
// the density of everybody's particles travels to the lower half and the potential back.  Each rank therefore owns
// TWO runs of nzl / 2 planes: run 0 in the occupied half, [r h, (r+1) h), next to its own particles for count-
// balanced cuts, and run 1 in the padding half, [nz/2 + r h, nz/2 + (r+1) h), h = nzl / 2 (`split`; odd nzl: one
// contiguous run as before).  Local plane zl < h belongs to run 0, the others to run 1.
struct SlabRuns {
  int nruns, len;      // runs per rank, planes per run
  int nz, nzl;
  __host__ __device__ int first(int rank, int run) const { return nruns == 1 ? rank * nzl : (run == 0 ? rank * len : nz / 2 + rank * len); }
};
static SlabRuns slab_runs(const p3m_ctx* c, int nz) {
  SlabRuns r;
  r.nz = nz, r.nzl = nz / c->nranks;
  const bool split = c->slab_split;
  r.nruns = split ? 2 : 1, r.len = split ? r.nzl / 2 : r.nzl;
  return r;
}

// ---- static plane ranges -------------------------------------------------------------------------------------
template <typename T>
static void plane_ranges(p3m_ctx* c) {
  const Geom<T>& g = Sel<T>::g(c);
  const int P = c->nranks;
  const int FD = g.fds == P3M_TWO_POINT ? 1 : 2;
  // z extent of the binning layers in mesh cells
  const double hz = g.p3m ? (double)g.hcz : (double)(1 << g.tile_shift);
  const int layers = g.cut[P];
  int top = (int)std::ceil(layers * hz) + (g.p3m ? 2 : (1 << g.tile_shift));
  if (top > g.nz) top = g.nz;
  for (int r = 0; r < P; ++r) {
    int zlo = r == 0 ? 0 : (int)std::floor(g.cut[r] * hz);
    int zhi = r == P - 1 ? top : (int)std::ceil(g.cut[r + 1] * hz) + 1;  // +1: rounding slop of pos / HC
    if (r > 0) zlo -= 1;
    zlo = std::max(zlo, 0), zhi = std::min(zhi, g.nz);
    // density: base cell in [zlo, zhi), stencil reaches one plane further on each side
    const int d0 = std::max(zlo - 1, 0), d1 = std::min(zhi + 1, g.nz);
    c->den_z0[r] = d0, c->den_nz[r] = d1 - d0;
    // potential: field planes zlo-1 .. zhi, each differencing +-FD planes, periodic -> unwrapped range
    int p0 = zlo - 1 - FD, p1 = zhi + 1 + FD;
    if (p1 - p0 >= g.nz) p0 = 0, p1 = g.nz;
    c->pot_z0[r] = p0, c->pot_nz[r] = p1 - p0;
  }
}

// planes this rank receives in slab_reduce_density: the overlaps of every OTHER rank's deposit range with
// my FFT slab.  With work-balanced cuts several thin particle slabs (each with its own halo planes) can sit
// inside one FFT slab, so this is computed from the ranges instead of bounded by a formula.
static size_t stage_planes_needed(const p3m_ctx* c, int nzl) {
  const SlabRuns sr = slab_runs(c, nzl * c->nranks);
  size_t need = 0;
  for (int p = 0; p < c->nranks; ++p) {
    if (p == c->rank) continue;
    for (int run = 0; run < sr.nruns; ++run) {
      const int lo = std::max(c->den_z0[p], sr.first(c->rank, run));
      const int hi = std::min(c->den_z0[p] + c->den_nz[p], sr.first(c->rank, run) + sr.len);
      if (hi > lo) need += (size_t)(hi - lo);
    }
  }
  return need + 1;
}

template <typename T>
void slab_free(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  if (!c->slab) return;
  void* ptrs[] = {s.dens_part, s.pot_part, s.den_stage, s.spectrum_t, s.pack};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  s.dens_part = s.pot_part = s.den_stage = nullptr;
  s.spectrum_t = s.pack = nullptr;
  if (s.plan_z_made) cufftDestroy(s.plan_z);
  s.plan_z_made = false;
}

// The layer cuts changed (dist_balance_cuts): re-derive the plane ranges of the particle-side buffers and
// re-allocate them.  The FFT slabs themselves never move.
template <typename T>
int slab_replan(p3m_ctx* c) {
  if (!c->slab) return 0;
  State<T>& s = Sel<T>::st(c);
  Geom<T>& g = Sel<T>::g(c);
  const int me = c->rank;
  const size_t plane = (size_t)g.nx * g.ny;
  const int old_den = c->den_nz[me], old_pot = c->pot_nz[me];
  const size_t old_stage = stage_planes_needed(c, g.nz / c->nranks);
  plane_ranges<T>(c);
  g.den_off = (long long)c->den_z0[me] * (long long)plane;
  g.den_len = (long long)c->den_nz[me] * (long long)plane;
  g.pot_z0 = c->pot_z0[me], g.pot_nz = c->pot_nz[me];
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  if (c->den_nz[me] > old_den) {
    cudaFree(s.dens_part);
    s.dens_part = nullptr;
    P3M_CUDA(cudaMalloc((void**)&s.dens_part, sizeof(T) * plane * (size_t)c->den_nz[me]));
  }
  if (c->pot_nz[me] > old_pot) {
    cudaFree(s.pot_part);
    s.pot_part = nullptr;
    P3M_CUDA(cudaMalloc((void**)&s.pot_part, sizeof(T) * plane * (size_t)c->pot_nz[me]));
  }
  const size_t new_stage = stage_planes_needed(c, g.nz / c->nranks);
  if (new_stage > old_stage) {
    cudaFree(s.den_stage);
    s.den_stage = nullptr;
    P3M_CUDA(cudaMalloc((void**)&s.den_stage, sizeof(T) * plane * new_stage));
  }
  P3M_CUDA(cudaMemsetAsync(s.pot_part, 0, sizeof(T) * plane * (size_t)c->pot_nz[me], c->stream));
  c->have_potential = false;
  return 0;
}

// Allocates the slab-sized meshes and the three cuFFT plans.  Called instead of the full-mesh allocation.
template <typename T>
int slab_setup(p3m_ctx* c) {
  using cplx = typename State<T>::cplx;
  State<T>& s = Sel<T>::st(c);
  Geom<T>& g = Sel<T>::g(c);
  const int P = c->nranks, me = c->rank;
  const int nxh = g.nx / 2 + 1, nyl = g.ny / P, nzl = g.nz / P;
  const size_t plane = (size_t)g.nx * g.ny;
  plane_ranges<T>(c);
  g.den_off = (long long)c->den_z0[me] * (long long)plane;
  g.den_len = (long long)c->den_nz[me] * (long long)plane;
  g.pot_z0 = c->pot_z0[me], g.pot_nz = c->pot_nz[me];
  const size_t spec = (size_t)nxh * g.ny * nzl;  // == nxh * nyl * nz
  P3M_CUDA(cudaMalloc((void**)&s.density, sizeof(T) * plane * nzl));
  P3M_CUDA(cudaMalloc((void**)&s.potential, sizeof(T) * plane * nzl));
  P3M_CUDA(cudaMalloc((void**)&s.dens_part, sizeof(T) * plane * (size_t)c->den_nz[me]));
  P3M_CUDA(cudaMalloc((void**)&s.pot_part, sizeof(T) * plane * (size_t)c->pot_nz[me]));
  // received density planes (re-sized by slab_replan when the cuts move)
  P3M_CUDA(cudaMalloc((void**)&s.den_stage, sizeof(T) * plane * stage_planes_needed(c, nzl)));
  P3M_CUDA(cudaMalloc((void**)&s.spectrum, sizeof(cplx) * spec));
  P3M_CUDA(cudaMalloc((void**)&s.spectrum_t, sizeof(cplx) * spec));
  P3M_CUDA(cudaMalloc((void**)&s.pack, sizeof(cplx) * spec));
  P3M_CUDA(cudaMalloc((void**)&s.green, sizeof(T) * spec));
  P3M_CUDA(cudaMemsetAsync(s.density, 0, sizeof(T) * plane * nzl, c->stream));
  P3M_CUDA(cudaMemsetAsync(s.potential, 0, sizeof(T) * plane * nzl, c->stream));
  P3M_CUDA(cudaMemsetAsync(s.pot_part, 0, sizeof(T) * plane * (size_t)c->pot_nz[me], c->stream));
  const bool dbl = sizeof(T) == 8;
  int n2[2] = {g.ny, g.nx};
  s.fft_chunk = fft_chunk_planes((long long)nxh * g.ny * sizeof(cplx), nzl, c->tune.fft_chunk_bytes);
  P3M_FFT(cufftPlanMany(&s.plan_fwd, 2, n2, nullptr, 1, 0, nullptr, 1, 0, dbl ? CUFFT_D2Z : CUFFT_R2C, s.fft_chunk));
  P3M_FFT(cufftPlanMany(&s.plan_inv, 2, n2, nullptr, 1, 0, nullptr, 1, 0, dbl ? CUFFT_Z2D : CUFFT_C2R, s.fft_chunk));
  s.plans = true;
  P3M_FFT(cufftSetStream(s.plan_fwd, c->stream));
  P3M_FFT(cufftSetStream(s.plan_inv, c->stream));
  if (!c->fused_z) {  // z leg through cuFFT (nz not a power of two): strided batched 1-D transforms
    int n1[1] = {g.nz};
    int embed[1] = {g.nz};
    const int stride = nxh * nyl;
    P3M_FFT(cufftPlanMany(&s.plan_z, 1, n1, embed, stride, 1, embed, stride, 1, dbl ? CUFFT_Z2Z : CUFFT_C2C, stride));
    s.plan_z_made = true;
    P3M_FFT(cufftSetStream(s.plan_z, c->stream));
  }
  return 0;
}

// ---- density: particle slabs -> FFT slabs -----------------------------------------------------------------------
template <typename T>
int slab_reduce_density(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const int P = c->nranks, me = c->rank, nzl = g.nz / P;
  const size_t plane = (size_t)g.nx * g.ny;
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  phase_begin(c, PH_COMM);
  const SlabRuns sr = slab_runs(c, g.nz);
  // planes of the deposit range of particle rank `src` that fall into run `run` of the FFT slab of rank `dst`
  auto overlap = [&](int src, int dst, int run, int& lo, int& hi) {
    lo = std::max(c->den_z0[src], sr.first(dst, run));
    hi = std::min(c->den_z0[src] + c->den_nz[src], sr.first(dst, run) + sr.len);
    return hi > lo;
  };
  P3M_CUDA(cudaMemsetAsync(s.density, 0, sizeof(T) * plane * nzl, c->stream));
  size_t stage_off[P3M_MAX_RANKS][2] = {{0}};
  size_t off = 0;
  double sent = 0;
  P3M_NCCL(ncclGroupStart());
  for (int p = 0; p < P; ++p) {
    if (p == me) continue;
    for (int run = 0; run < sr.nruns; ++run) {
      int lo, hi;
      if (overlap(me, p, run, lo, hi)) {
        sent += (double)(hi - lo) * (double)plane * sizeof(T);
        P3M_NCCL(ncclSend(s.dens_part + (size_t)(lo - c->den_z0[me]) * plane, (size_t)(hi - lo) * plane, nccl_real<T>(), p,
                          comm, c->stream));
      }
      if (overlap(p, me, run, lo, hi)) {
        stage_off[p][run] = off;
        P3M_NCCL(ncclRecv(s.den_stage + off, (size_t)(hi - lo) * plane, nccl_real<T>(), p, comm, c->stream));
        off += (size_t)(hi - lo) * plane;
      }
    }
  }
  P3M_NCCL(ncclGroupEnd());
  c->launches++;
  c->stat_den_bytes = sent;
  // sum in rank order: bit-reproducible
  for (int p = 0; p < P; ++p)
    for (int run = 0; run < sr.nruns; ++run) {
      int lo, hi;
      if (!overlap(p, me, run, lo, hi)) continue;
      const T* src = p == me ? s.dens_part + (size_t)(lo - c->den_z0[me]) * plane : s.den_stage + stage_off[p][run];
      const long long cnt = (long long)(hi - lo) * (long long)plane;
      const size_t local = (size_t)(run * sr.len + lo - sr.first(me, run));  // local plane of global plane lo
      k_add_planes<T><<<grid_for(cnt, c->num_sms), 256, 0, c->stream>>>(s.density + local * plane, src, cnt);
      P3M_LAUNCH_CHECK(c);
    }
  phase_end(c, PH_COMM);
  return 0;
}

// ---- potential: FFT slabs -> particle slabs (unwrapped, with halo) ------------------------------------------------
template <typename T>
int slab_spread_potential(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const int P = c->nranks, me = c->rank, nzl = g.nz / P;
  const size_t plane = (size_t)g.nx * g.ny;
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  phase_begin(c, PH_COMM);
  const SlabRuns sr = slab_runs(c, g.nz);
  // planes of run `run` of owner `own`, periodic image k, wanted by particle rank `dst`: unwrapped [lo, hi)
  auto overlap = [&](int own, int run, int k, int dst, int& lo, int& hi) {
    lo = std::max(sr.first(own, run) + k * g.nz, c->pot_z0[dst]);
    hi = std::min(sr.first(own, run) + sr.len + k * g.nz, c->pot_z0[dst] + c->pot_nz[dst]);
    return hi > lo;
  };
  // local plane (in the slab of `own`) of unwrapped plane z of image k
  auto local_of = [&](int own, int run, int k, int z) { return (size_t)(run * sr.len + z - k * g.nz - sr.first(own, run)); };
  double sent = 0;
  P3M_NCCL(ncclGroupStart());
  for (int p = 0; p < P; ++p)
    for (int run = 0; run < sr.nruns; ++run)
      for (int k = -1; k <= 1; ++k) {
        int lo, hi;
        if (p != me && overlap(me, run, k, p, lo, hi)) {
          sent += (double)(hi - lo) * (double)plane * sizeof(T);
          P3M_NCCL(ncclSend(s.potential + local_of(me, run, k, lo) * plane, (size_t)(hi - lo) * plane, nccl_real<T>(), p,
                            comm, c->stream));
        }
        if (p != me && overlap(p, run, k, me, lo, hi))
          P3M_NCCL(ncclRecv(s.pot_part + (size_t)(lo - c->pot_z0[me]) * plane, (size_t)(hi - lo) * plane, nccl_real<T>(),
                            p, comm, c->stream));
      }
  P3M_NCCL(ncclGroupEnd());
  c->launches++;
  c->stat_pot_bytes = sent;
  for (int run = 0; run < sr.nruns; ++run)
    for (int k = -1; k <= 1; ++k) {
      int lo, hi;
      if (overlap(me, run, k, me, lo, hi))
        P3M_CUDA(cudaMemcpyAsync(s.pot_part + (size_t)(lo - c->pot_z0[me]) * plane, s.potential + local_of(me, run, k, lo) * plane,
                                 sizeof(T) * (size_t)(hi - lo) * plane, cudaMemcpyDeviceToDevice, c->stream));
    }
  phase_end(c, PH_COMM);
  return 0;
}

// ---- the distributed Poisson solve ----------------------------------------------------------------------------------
// Chunk p of the packed array = [run][plane in run][ky_local][kx] for / from rank p.  In the transposed array
// [kz][ky_local][kx] the planes of run `run` of rank p start at kz = first(p, run): with split slabs a peer's chunk
// lands in two places.  to_t: packed planes -> transposed array; else the way back.
template <typename T>
static int all_to_all(p3m_ctx* c, typename State<T>::cplx* packed, typename State<T>::cplx* transposed, size_t chunk,
                      bool to_t) {
  using cplx = typename State<T>::cplx;
  const int P = c->nranks, me = c->rank;
  const SlabRuns sr = slab_runs(c, c->prm.nz);
  const size_t part = chunk / (size_t)sr.nruns;          // elements of one run of one chunk
  const size_t per_plane = chunk / (size_t)sr.nzl;       // nxh * nyl
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  P3M_NCCL(ncclGroupStart());
  for (int p = 0; p < P; ++p) {
    if (p == me) continue;
    for (int run = 0; run < sr.nruns; ++run) {
      cplx* pk = packed + (size_t)p * chunk + (size_t)run * part;             // my planes of this run, rows of rank p
      cplx* tr = transposed + (size_t)sr.first(p, run) * per_plane;           // planes of rank p's run, my rows
      if (to_t) {
        P3M_NCCL(ncclSend(pk, part * sizeof(cplx), ncclChar, p, comm, c->stream));
        P3M_NCCL(ncclRecv(tr, part * sizeof(cplx), ncclChar, p, comm, c->stream));
      } else {
        P3M_NCCL(ncclSend(tr, part * sizeof(cplx), ncclChar, p, comm, c->stream));
        P3M_NCCL(ncclRecv(pk, part * sizeof(cplx), ncclChar, p, comm, c->stream));
      }
    }
  }
  P3M_NCCL(ncclGroupEnd());
  for (int run = 0; run < sr.nruns; ++run) {
    cplx* pk = packed + (size_t)me * chunk + (size_t)run * part;
    cplx* tr = transposed + (size_t)sr.first(me, run) * per_plane;
    P3M_CUDA(cudaMemcpyAsync(to_t ? tr : pk, to_t ? pk : tr, part * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream));
  }
  c->launches += 2;
  c->stat_a2a_bytes += (double)(P - 1) * (double)chunk * sizeof(cplx);
  return 0;
}

template <typename T>
int slab_poisson(p3m_ctx* c) {
  using cplx = typename State<T>::cplx;
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  const int P = c->nranks;
  const int nxh = g.nx / 2 + 1, nyl = g.ny / P, nzl = g.nz / P;
  const long long spec = (long long)nxh * g.ny * nzl;
  const size_t chunk = (size_t)nxh * nyl * nzl;
  const int grid = grid_for(spec, c->num_sms);
  const size_t plane = (size_t)g.nx * g.ny, splane = (size_t)nxh * g.ny;
  c->stat_a2a_bytes = 0;
  phase_begin(c, PH_FFT_FWD);
  for (int z = 0; z < nzl; z += s.fft_chunk) {
    P3M_FFT(exec_r2c(s.plan_fwd, s.density + plane * z, s.spectrum + splane * z));
    c->launches += 2;
  }
  k_transpose_pack<cplx, true><<<grid, 256, 0, c->stream>>>(s.spectrum, s.pack, nxh, g.ny, nyl, nzl);
  P3M_LAUNCH_CHECK(c);
  c->launches += 1;
  phase_end(c, PH_FFT_FWD);
  phase_begin(c, PH_COMM);
  P3M_TRY(all_to_all<T>(c, s.pack, s.spectrum_t, chunk, true));  // chunk p = planes of rank p: [kx, ky_local, z]
  phase_end(c, PH_COMM);
  if (c->fused_z) {
    phase_begin(c, PH_MULTIPLY);  // forward z FFT + multiply + inverse z FFT in one pass (poisson_z.cu)
    P3M_TRY(fused_z_pass<T>(c, s.spectrum_t, s.green, (long long)nxh * nyl));
    phase_end(c, PH_MULTIPLY);
  } else {
    phase_begin(c, PH_FFT_FWD);
    P3M_FFT(exec_c2c(s.plan_z, s.spectrum_t, CUFFT_FORWARD));
    c->launches++;
    phase_end(c, PH_FFT_FWD);
    phase_begin(c, PH_MULTIPLY);
    k_multiply_t<<<grid, 256, 0, c->stream>>>(s.spectrum_t, s.green, spec);
    P3M_LAUNCH_CHECK(c);
    phase_end(c, PH_MULTIPLY);
    phase_begin(c, PH_FFT_INV);
    P3M_FFT(exec_c2c(s.plan_z, s.spectrum_t, CUFFT_INVERSE));
    c->launches++;
    phase_end(c, PH_FFT_INV);
  }
  phase_begin(c, PH_COMM);
  P3M_TRY(all_to_all<T>(c, s.pack, s.spectrum_t, chunk, false));  // planes of rank p in the transposed array -> chunk p
  phase_end(c, PH_COMM);
  phase_begin(c, PH_FFT_INV);
  k_transpose_pack<cplx, false><<<grid, 256, 0, c->stream>>>(s.spectrum, s.pack, nxh, g.ny, nyl, nzl);
  P3M_LAUNCH_CHECK(c);
  for (int z = 0; z < nzl; z += s.fft_chunk) {
    P3M_FFT(exec_c2r(s.plan_inv, s.spectrum + splane * z, s.potential + plane * z));
    c->launches += 2;
  }
  phase_end(c, PH_FFT_INV);
  return 0;
}

template int slab_replan<float>(p3m_ctx*);
template int slab_replan<double>(p3m_ctx*);
template void slab_free<float>(p3m_ctx*);
template void slab_free<double>(p3m_ctx*);
template int slab_setup<float>(p3m_ctx*);
template int slab_setup<double>(p3m_ctx*);
template int slab_reduce_density<float>(p3m_ctx*);
template int slab_reduce_density<double>(p3m_ctx*);
template int slab_spread_potential<float>(p3m_ctx*);
template int slab_spread_potential<double>(p3m_ctx*);
template int slab_poisson<float>(p3m_ctx*);
template int slab_poisson<double>(p3m_ctx*);

}  // namespace p3m
