// capi.cu -- the extern "C" boundary declared in include/p3m_b200.h.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "ctx.cuh"

namespace p3m {

static thread_local char g_err[1024] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void p3m_tune_load_impl(p3m_tune& t) {
  auto flag = [](const char* name) {
    const char* e = getenv(name);
    return e && *e && !(e[0] == '0' && e[1] == 0);
  };
  if (const char* e = getenv("P3M_TUNE_SUBBITS")) t.subbits = atoi(e);
  t.long_key = flag("P3M_TUNE_LONGKEY");
  t.old_deposit = flag("P3M_TUNE_OLD_DEPOSIT");
  t.old_gather = flag("P3M_TUNE_OLD_GATHER");
  t.static_cuts = flag("P3M_STATIC_CUTS");
  t.count_cuts = flag("P3M_COUNT_CUTS");
  if (const char* e = getenv("P3M_TUNE_PARTICLE_WEIGHT")) t.particle_weight = atof(e);
  if (const char* e = getenv("P3M_TUNE_FFT_CHUNK_MB")) t.fft_chunk_bytes = atoll(e) << 20;
  t.replicated_mesh = flag("P3M_REPLICATED_MESH");
  t.cufft_z = flag("P3M_TUNE_CUFFT_Z");
  t.full_sort = flag("P3M_TUNE_FULL_SORT");
  t.scalar_pp = flag("P3M_TUNE_SCALAR_PP");
  t.z_wide = flag("P3M_TUNE_Z_WIDE");
  t.contig_slabs = flag("P3M_TUNE_CONTIG_SLABS");
  t.no_prune = flag("P3M_TUNE_NO_PRUNE");
  if (const char* e = getenv("P3M_TUNE_INC_SORT_DEN")) t.inc_sort_den = atoi(e) > 1 ? atoi(e) : 2;
  if (const char* e = getenv("P3M_TUNE_A2A_CHUNKS")) t.a2a_chunks = atoi(e);
  if (const char* e = getenv("P3M_TUNE_DENSE_CELL")) t.dense_cell = atoi(e) > 0 ? atoi(e) : 1;
}

static cudaEvent_t timer_event(PhaseTimer& t) {
  if (!t.pool.empty()) {
    cudaEvent_t e = t.pool.back();
    t.pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

void phase_begin(p3m_ctx* c, int ph) {
  if (!c->timing) return;
  PhaseTimer& t = c->timer;
  if (t.cur[ph]) t.pool.push_back(t.cur[ph]);  // begin without end: drop
  t.cur[ph] = timer_event(t);
  cudaEventRecord(t.cur[ph], c->stream);
}

void phase_end(p3m_ctx* c, int ph) {
  if (!c->timing) return;
  PhaseTimer& t = c->timer;
  if (!t.cur[ph]) return;
  cudaEvent_t b = timer_event(t);
  cudaEventRecord(b, c->stream);
  t.open_.push_back(PhaseTimer::Interval{ph, t.cur[ph], b});
  t.cur[ph] = nullptr;
  if (t.open_.size() > 4096) phase_resolve(c);  // bound the pool on very long runs
}

// fold every recorded interval into acc_ms (synchronises the stream once)
void phase_resolve(p3m_ctx* c) {
  PhaseTimer& t = c->timer;
  if (t.open_.empty()) return;
  cudaStreamSynchronize(c->stream);
  for (const PhaseTimer::Interval& iv : t.open_) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, iv.a, iv.b) == cudaSuccess) t.acc_ms[iv.ph] += ms;
    t.pool.push_back(iv.a);
    t.pool.push_back(iv.b);
  }
  t.open_.clear();
}

template <typename T> int escaped_now(p3m_ctx* c, int* escaped);

template <typename T>
static int setup_geometry(p3m_ctx* c) {
  const p3m_params& p = c->prm;
  Geom<T>& g = Sel<T>::g(c);
  g.nx = p.nx, g.ny = p.ny, g.nz = p.nz;
  g.M = (long long)p.nx * p.ny * p.nz;
  g.is = p.assignment, g.fds = p.fd_scheme;
  g.p3m = p.p3m;
  g.H = (T)p.H, g.DT = (T)p.DT, g.G = (T)p.G;
  g.boxx = (T)p.box[0], g.boxy = (T)p.box[1], g.boxz = (T)p.box[2];
  g.ext_kind = p.ext_kind;
  g.ecx = (T)p.ext_center[0], g.ecy = (T)p.ext_center[1], g.ecz = (T)p.ext_center[2];
  g.eR = (T)p.ext_R, g.eM = (T)p.ext_M;
  g.unit_roundtrip = p.unit_roundtrip;
  g.tile_shift = kPmTileShift;
  g.idbits = 1;
  g.sbits = 0;
  if (p.p3m) {
    // ChainingMesh::ChainingMesh, source/chainingMesh.cpp:10-15
    const T box[3] = {(T)p.box[0], (T)p.box[1], (T)p.box[2]};
    int m[3];
    T hc[3];
    for (int d = 0; d < 3; ++d) {
      m[d] = (int)(box[d] / (T)p.cutoff_radius);
      if (m[d] < 1) return fail(P3M_EINVAL, "cutoff radius %g exceeds the box", (double)p.cutoff_radius);
      hc[d] = (box[d] / m[d]) / (T)p.H;
    }
    g.mx = m[0], g.my = m[1], g.mz = m[2];
    g.hcx = hc[0], g.hcy = hc[1], g.hcz = hc[2];
  } else {
    const int t = 1 << kPmTileShift;
    g.mx = (p.nx + t - 1) / t, g.my = (p.ny + t - 1) / t, g.mz = (p.nz + t - 1) / t;
    g.hcx = g.hcy = g.hcz = (T)t;
  }
  int mmax = g.mx > g.my ? g.mx : g.my;
  mmax = mmax > g.mz ? mmax : g.mz;
  g.mbits = 0;
  while ((1 << g.mbits) < mmax) ++g.mbits;
  if (g.mbits > 10) return fail(P3M_EINVAL, "binning mesh %dx%dx%d exceeds 1024 cells per axis", g.mx, g.my, g.mz);
  // tile = aligned block of (1 << bshift)^3 binning cells; largest block whose mesh footprint fits the
  // per-warp shared-memory budget of the deposit kernel
  int best = -1;
  for (int b = g.mbits; b >= 0; --b) {
    g.bshift = b;
    const int B = 1 << b;
    int ex[3] = {0, 0, 0};
    const int nb[3] = {(g.mx + B - 1) / B, (g.my + B - 1) / B, (g.mz + B - 1) / B};
    for (int d = 0; d < 3; ++d)
      for (int i = 0; i < nb[d]; ++i) {
        int lo[3], ext[3];
        tile_box(g, d == 0 ? i : 0, d == 1 ? i : 0, d == 2 ? i : 0, lo, ext);
        if (ext[d] > ex[d]) ex[d] = ext[d];
      }
    const long long elems = (long long)ex[0] * ex[1] * ex[2];
    g.tex = ex[0], g.tey = ex[1], g.tez = ex[2];
    if (elems * (long long)sizeof(T) <= kMaxTileBytes) {
      best = b;
      break;
    }
  }
  // gather tile of P3M contexts: largest block whose footprint + 2-cell finite-difference halo fits kGatherTile^3
  g.gshift = -1;
  if (p.p3m) {
    const int keep = g.bshift;
    for (int b = g.mbits; b >= 0 && g.gshift < 0; --b) {
      g.bshift = b;
      const int B = 1 << b;
      const int nb[3] = {(g.mx + B - 1) / B, (g.my + B - 1) / B, (g.mz + B - 1) / B};
      bool ok = true;
      for (int d = 0; d < 3 && ok; ++d)
        for (int i = 0; i < nb[d] && ok; ++i) {
          int lo[3], ext[3];
          tile_box(g, d == 0 ? i : 0, d == 1 ? i : 0, d == 2 ? i : 0, lo, ext);
          ok = ext[d] + 4 <= kGatherTile;
        }
      if (ok) g.gshift = b;
    }
    g.bshift = keep;
  }
  if (best < 0) {
    // a single binning cell is larger than the tile budget: direct paths only
    g.bshift = 0;
    g.tex = g.tey = g.tez = 1;
    g.tile_min = 0x7fffffff;
  } else {
    g.tile_min = kTileMinCount;
  }
  return 0;
}

template <typename T>
static int create_typed(p3m_ctx* c, const void* uid, int rank, int nranks) {
  P3M_TRY(setup_geometry<T>(c));
  P3M_TRY(dist_init(c, uid, rank, nranks));  // cuts + communicator first: the mesh layout depends on them
  P3M_TRY(alloc_meshes<T>(c));
  if (c->prm.p3m) P3M_TRY(sr_table_upload<T>(c));
  return 0;
}

}  // namespace p3m

using namespace p3m;

void p3m_tune::load() { p3m::p3m_tune_load_impl(*this); }

extern "C" {

const char* p3m_last_error(void) { return g_err; }
int p3m_version(void) { return 100; }

void p3m_default_params(p3m_params* p) {
  memset(p, 0, sizeof(*p));
  p->DT = 1.0f;
  p->G = 1.0f;
  p->H = 1.0f;
  p->assignment = P3M_TSC;
  p->fd_scheme = P3M_TWO_POINT;
  p->greens_function = P3M_S1_OPTIMAL;
  p->cloud_shape = P3M_S1;
  p->use_sr_table = 1;
  p->precision = P3M_F32;
  p->unit_roundtrip = 1;
  p->green_zero_degenerate = 1;
  p->device = -1;
}

static int create_common(const p3m_params* prm, const void* uid, int rank, int nranks, p3m_ctx** out) {
  if (!prm || !out) return fail(P3M_EINVAL, "p3m_create: null argument");
  *out = nullptr;
  if (prm->nx < 2 || prm->ny < 2 || prm->nz < 2) return fail(P3M_EINVAL, "mesh must be at least 2^3");
  if (prm->assignment < P3M_NGP || prm->assignment > P3M_TSC)
    return fail(P3M_EINVAL, "Unkown interpolation scheme");  // source/pmMethod.cpp:275
  if (prm->fd_scheme < P3M_TWO_POINT || prm->fd_scheme > P3M_FOUR_POINT)
    return fail(P3M_EINVAL, "Unknown finite difference type.");  // source/pmMethod.cpp:367
  if (prm->greens_function < P3M_DISCRETE_LAPLACIAN || prm->greens_function > P3M_POOR_MAN)
    return fail(P3M_EINVAL, "not implemented");  // source/pmMethod.cpp:180
  if (prm->p3m && (prm->cloud_shape < P3M_S1 || prm->cloud_shape > P3M_S2))
    return fail(P3M_EINVAL, "not implemented");  // source/p3mMethod.cpp:285
  if (prm->ext_kind < P3M_EXT_NONE || prm->ext_kind > P3M_EXT_SPH_RAD_DECR)
    return fail(P3M_EINVAL, "unknown external field kind %d", prm->ext_kind);
  if (!(prm->H > 0) || !(prm->DT > 0)) return fail(P3M_EINVAL, "H and DT must be positive");
  if (prm->p3m && !(prm->cutoff_radius > 0)) return fail(P3M_EINVAL, "cutoff radius must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(P3M_ENODEV, "no CUDA device available (this library has no CPU fallback)");
  }
  int dev = prm->device;
  if (dev < 0) P3M_CUDA(cudaGetDevice(&dev));
  if (dev >= ndev) return fail(P3M_ENODEV, "device %d out of range (%d devices)", dev, ndev);
  P3M_CUDA(cudaSetDevice(dev));
  p3m_ctx* c = new (std::nothrow) p3m_ctx();
  if (!c) return fail(P3M_EINVAL, "out of host memory");
  c->prm = *prm;
  c->tune.load();
  c->device = dev;
  c->f64 = prm->precision == P3M_F64;
  c->timing = prm->timing != 0;
  cudaDeviceProp prop;
  P3M_CUDA(cudaGetDeviceProperties(&prop, dev));
  c->num_sms = prop.multiProcessorCount;
  P3M_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    // staging buffers come from the stream-ordered pool: keep freed blocks cached across synchronisations
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  {
    // massToCodeUnits factor, left to right as include/unitConversions.h:42-44
    const float DT = prm->DT, H = prm->H, G = prm->G, pi = 3.14159265358979323846f;
    c->mass_factor32 = DT * DT * 4 * pi * G / (H * H * H);
    const double dDT = DT, dH = H, dG = G, dpi = 3.14159265358979323846;
    c->mass_factor64 = dDT * dDT * 4 * dpi * dG / (dH * dH * dH);
  }
  int r = c->f64 ? create_typed<double>(c, uid, rank, nranks) : create_typed<float>(c, uid, rank, nranks);
  if (r != 0) {
    p3m_destroy(c);
    return r;
  }
  *out = c;
  return 0;
}

int p3m_create(const p3m_params* prm, p3m_ctx** out) {
  return create_common(prm, nullptr, 0, 1, out);  // single rank: owns every layer
}

int p3m_slab_cuts(const p3m_params* prm, int nranks, int32_t cuts[9], int32_t* layers_out) {
  if (!prm || !cuts || nranks < 1 || nranks > 8) return fail(P3M_EINVAL, "p3m_slab_cuts: bad argument");
  p3m_ctx c;
  c.prm = *prm;
  c.f64 = false;
  c.rank = 0, c.nranks = nranks;
  int r = setup_geometry<float>(&c);
  if (r != 0) return r;
  dist_set_cuts(&c);
  for (int i = 0; i < 9; ++i) cuts[i] = c.g32.cut[i];
  if (layers_out) *layers_out = c.g32.cut[nranks];
  if (c.g32.cut[nranks] < nranks) return fail(P3M_EINVAL, "only %d binning layers along z for %d ranks", c.g32.cut[nranks], nranks);
  return 0;
}

int p3m_balanced_cuts(const p3m_params* prm, int nranks, const float* pos, int64_t n, int units, int32_t cuts[9],
                      int32_t* layers_out) {
  if (!prm || !cuts || nranks < 1 || nranks > 8 || n < 0 || (n > 0 && !pos))
    return fail(P3M_EINVAL, "p3m_balanced_cuts: bad argument");
  p3m_ctx c;
  c.prm = *prm;
  c.tune.load();
  c.f64 = false;
  c.rank = 0, c.nranks = nranks;
  int r = setup_geometry<float>(&c);
  if (r != 0) return r;
  dist_set_cuts(&c);
  if (c.g32.cut[nranks] < nranks)
    return fail(P3M_EINVAL, "only %d binning layers along z for %d ranks", c.g32.cut[nranks], nranks);
  r = dist_balance_cuts<float>(&c, pos, (long long)n, units);  // host arithmetic only (no slab buffers here)
  if (r != 0) return r;
  for (int i = 0; i < 9; ++i) cuts[i] = c.g32.cut[i];
  if (layers_out) *layers_out = c.g32.cut[nranks];
  return 0;
}

int p3m_comm_unique_id(void* out) {
  if (!out) return fail(P3M_EINVAL, "null output");
  return comm_unique_id(out);
}

int p3m_create_dist(const p3m_params* prm, const void* uid, int rank, int nranks, p3m_ctx** out) {
  if (nranks < 1 || nranks > P3M_MAX_RANKS || rank < 0 || rank >= nranks)
    return fail(P3M_EINVAL, "p3m_create_dist: rank %d of %d (1..%d ranks supported)", rank, nranks, P3M_MAX_RANKS);
  if (!uid && nranks > 1) return fail(P3M_EINVAL, "p3m_create_dist: null unique id");
  return create_common(prm, uid, rank, nranks, out);
}

int p3m_get_local(p3m_ctx* c, int32_t* ids, float* pos, float* vel, float* acc, int units) {
  if (!c) return fail(P3M_EINVAL, "null context");
  P3M_CUDA(cudaSetDevice(c->device));
  if (!c->have_particles) return fail(P3M_ESTATE, "no particles set");
  if (acc && c->n > 0 && !c->have_acc) return fail(P3M_ESTATE, "p3m_get_local: accelerations are stale (re-sorted since the last p3m_gather)");
  return P3M_DISPATCH(c, download_local, ids, pos, vel, acc, units);
}

int p3m_set_particles_ids(p3m_ctx* c, const float* pos, const float* vel, const float* mass, const int32_t* ids,
                          int64_t n, int units) {
  if (!c) return fail(P3M_EINVAL, "null context");
  P3M_CUDA(cudaSetDevice(c->device));
  if (n < 0 || (n > 0 && (!pos || !mass || !ids))) return fail(P3M_EINVAL, "p3m_set_particles_ids: bad argument");
  int r = P3M_DISPATCH(c, upload_particles_ids, pos, vel, mass, ids, (long long)n, units);
  if (r == 0) {
    int* flags = c->f64 ? c->s64.flags : c->s32.flags;
    P3M_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * 4, c->stream));
  }
  return r;
}

int64_t p3m_num_global(const p3m_ctx* c) { return c ? (c->nranks > 1 ? c->n_global : c->n) : 0; }

int p3m_rank_info(p3m_ctx* c, int64_t out[8]) {
  if (!c || !out) return fail(P3M_EINVAL, "null argument");
  out[0] = c->rank, out[1] = c->nranks;
  out[2] = c->f64 ? c->g64.cut[c->rank] : c->g32.cut[c->rank];
  out[3] = c->f64 ? c->g64.cut[c->rank + 1 > 8 ? 8 : c->rank + 1] : c->g32.cut[c->rank + 1 > 8 ? 8 : c->rank + 1];
  out[4] = c->f64 ? c->s64.n_ghost : c->s32.n_ghost;
  const int nzl = c->prm.nz / (c->slab ? c->nranks : 1);
  out[5] = c->slab ? 1 : 0, out[6] = c->slab ? (int64_t)c->rank * nzl : 0, out[7] = nzl;
  return 0;
}

int p3m_destroy(p3m_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  dist_destroy(c);
  free_state<float>(c);
  free_state<double>(c);
  phase_resolve(c);
  for (cudaEvent_t e : c->timer.pool) cudaEventDestroy(e);
  for (int i = 0; i < P3M_NPHASE; ++i)
    if (c->timer.cur[i]) cudaEventDestroy(c->timer.cur[i]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

#define CHECK_CTX(c)                                            \
  do {                                                          \
    if (!(c)) return fail(P3M_EINVAL, "null context");          \
    P3M_CUDA(cudaSetDevice((c)->device));                       \
  } while (0)

int p3m_set_particles(p3m_ctx* c, const float* pos, const float* vel, const float* mass, int64_t n,
                      int units) {
  CHECK_CTX(c);
  if (n < 0 || n > 0x7fffffffLL) return fail(P3M_EINVAL, "particle count %lld out of range", (long long)n);
  if (n > 0 && (!pos || !mass)) return fail(P3M_EINVAL, "p3m_set_particles: null pos/mass");
  if (units != P3M_UNITS_ORIGINAL && units != P3M_UNITS_CODE) return fail(P3M_EINVAL, "bad units");
  int r = P3M_DISPATCH(c, upload_particles, pos, vel, mass, (long long)n, units);
  if (r == 0) {
    State<float>& s = c->s32;
    (void)s;
    int* flags = c->f64 ? c->s64.flags : c->s32.flags;
    P3M_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * 4, c->stream));
  }
  return r;
}

static_assert(sizeof(p3m_ic) == 104 && sizeof(p3m_params) == 120, "ABI structs are mirrored by ctypes (capi.py)");

int p3m_generate_particles(p3m_ctx* c, const p3m_ic* ic) {
  CHECK_CTX(c);
  return P3M_DISPATCH(c, generate_particles, ic);
}

int p3m_sample_particles(const p3m_ic* ic, int64_t first, int64_t count, float* pos, float* vel, float* mass) {
  return sample_particles(ic, (long long)first, (long long)count, pos, vel, mass);
}

int p3m_get_particles(p3m_ctx* c, float* pos, float* vel, float* acc, int units) {
  CHECK_CTX(c);
  if (!c->have_particles) return fail(P3M_ESTATE, "no particles set");
  if (acc && c->n > 0 && !c->have_acc) return fail(P3M_ESTATE, "p3m_get_particles: accelerations are stale (re-sorted since the last p3m_gather)");
  return c->f64 ? download_particles<double, float>(c, pos, vel, acc, units)
                : download_particles<float, float>(c, pos, vel, acc, units);
}

int p3m_get_particles_f64(p3m_ctx* c, double* pos, double* vel, double* acc, int units) {
  CHECK_CTX(c);
  if (!c->have_particles) return fail(P3M_ESTATE, "no particles set");
  if (acc && c->n > 0 && !c->have_acc) return fail(P3M_ESTATE, "p3m_get_particles_f64: accelerations are stale (re-sorted since the last p3m_gather)");
  return c->f64 ? download_particles<double, double>(c, pos, vel, acc, units)
                : download_particles<float, double>(c, pos, vel, acc, units);
}

int64_t p3m_num_particles(const p3m_ctx* c) { return c ? c->n : 0; }

int p3m_green_init(p3m_ctx* c) {
  CHECK_CTX(c);
  return P3M_DISPATCH(c, green_init);
}

int p3m_set_green_table(p3m_ctx* c, const float* t) {
  CHECK_CTX(c);
  if (!t) return fail(P3M_EINVAL, "null table");
  return c->f64 ? green_set<double, float>(c, t) : green_set<float, float>(c, t);
}

int p3m_set_green_table_f64(p3m_ctx* c, const double* t) {
  CHECK_CTX(c);
  if (!t) return fail(P3M_EINVAL, "null table");
  return c->f64 ? green_set<double, double>(c, t) : green_set<float, double>(c, t);
}

int p3m_get_green_table(p3m_ctx* c, double* t) {
  CHECK_CTX(c);
  if (!t) return fail(P3M_EINVAL, "null table");
  return P3M_DISPATCH(c, green_get, t);
}

int p3m_bin_sort(p3m_ctx* c) {
  CHECK_CTX(c);
  return P3M_DISPATCH(c, bin_sort);
}
int p3m_deposit(p3m_ctx* c) {
  CHECK_CTX(c);
  return P3M_DISPATCH(c, deposit);
}
int p3m_poisson(p3m_ctx* c) {
  CHECK_CTX(c);
  return P3M_DISPATCH(c, poisson);
}
int p3m_gradient(p3m_ctx* c) {
  CHECK_CTX(c);
  return P3M_DISPATCH(c, gradient);
}
int p3m_gather(p3m_ctx* c) {
  CHECK_CTX(c);
  return P3M_DISPATCH(c, gather);
}
int p3m_short_range(p3m_ctx* c) {
  CHECK_CTX(c);
  return P3M_DISPATCH(c, short_range);
}

int p3m_force(p3m_ctx* c) {
  CHECK_CTX(c);
  if (!c->have_green) P3M_TRY(p3m_green_init(c));
  P3M_TRY(P3M_DISPATCH(c, bin_sort));
  P3M_TRY(P3M_DISPATCH(c, deposit));
  P3M_TRY(P3M_DISPATCH(c, poisson));
  P3M_TRY(P3M_DISPATCH(c, gather));
  P3M_TRY(P3M_DISPATCH(c, short_range));
  return 0;
}

int p3m_kick(p3m_ctx* c, float f) {
  CHECK_CTX(c);
  if (c->have_particles && c->n > 0 && !c->have_acc)
    return fail(P3M_ESTATE, "p3m_kick: accelerations are stale (particles were re-sorted since the last p3m_gather)");
  return P3M_DISPATCH(c, kick, (double)f);
}
int p3m_drift(p3m_ctx* c) {
  CHECK_CTX(c);
  return P3M_DISPATCH(c, drift);
}

int p3m_step(p3m_ctx* c, int steps, int* steps_done) {
  CHECK_CTX(c);
  if (steps < 0) return fail(P3M_EINVAL, "negative step count");
  int* flags = c->f64 ? c->s64.flags : c->s32.flags;
  P3M_CUDA(cudaMemsetAsync(flags + 2, 0, sizeof(int), c->stream));
  for (int t = 0; t < steps; ++t) {
    P3M_TRY(p3m_drift(c));   // updatePositions (+ unit round trip + escape test)
    P3M_TRY(p3m_force(c));   // pmMethodStep [+ calculateShortRangeForces + correctAccelerations]
    P3M_TRY(p3m_kick(c, 1.0f));  // updateVelocities
  }
  int h[4] = {0, 0, 0, 0};
  P3M_CUDA(cudaMemcpyAsync(h, flags, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  if (steps_done) *steps_done = h[2];
  return 0;
}

int p3m_escaped(p3m_ctx* c, int* escaped) {
  CHECK_CTX(c);
  if (!escaped) return fail(P3M_EINVAL, "null output");
  if (!c->have_particles) return fail(P3M_ESTATE, "no particles set");
  return c->f64 ? escaped_now<double>(c, escaped) : escaped_now<float>(c, escaped);
}

int p3m_add_acceleration(p3m_ctx* c, const float* a, int units) {
  CHECK_CTX(c);
  if (!a) return fail(P3M_EINVAL, "null acceleration array");
  if (!c->have_particles) return fail(P3M_ESTATE, "no particles set");
  if (c->n > 0 && !c->have_acc) return fail(P3M_ESTATE, "p3m_add_acceleration: accelerations are stale (re-sorted since the last p3m_gather)");
  return P3M_DISPATCH(c, add_acceleration, a, units);
}

int p3m_fft3d_c2c(int nz, int ny, int nx, const float* in, float* out, int inverse) {
  return fft3d_c2c(nz, ny, nx, in, out, inverse);
}

int p3m_diagnostics(p3m_ctx* c, double out[11]) {
  CHECK_CTX(c);
  return P3M_DISPATCH(c, diagnostics, out);
}

int p3m_get_density(p3m_ctx* c, float* m) {
  CHECK_CTX(c);
  return c->f64 ? get_mesh<double, float>(c, c->s64.density, m, c->g64.M)
                : get_mesh<float, float>(c, c->s32.density, m, c->g32.M);
}
int p3m_get_potential(p3m_ctx* c, float* m) {
  CHECK_CTX(c);
  P3M_TRY(P3M_DISPATCH(c, complete_potential));  // planes the pruned solve skipped
  return c->f64 ? get_mesh<double, float>(c, c->s64.potential, m, c->g64.M)
                : get_mesh<float, float>(c, c->s32.potential, m, c->g32.M);
}
int p3m_get_field(p3m_ctx* c, float* m) {
  CHECK_CTX(c);
  if (!c->have_field) return fail(P3M_ESTATE, "call p3m_gradient first");
  return c->f64 ? get_mesh<double, float>(c, c->s64.field, m, 3 * c->g64.M)
                : get_mesh<float, float>(c, c->s32.field, m, 3 * c->g32.M);
}
int p3m_get_density_f64(p3m_ctx* c, double* m) {
  CHECK_CTX(c);
  return c->f64 ? get_mesh<double, double>(c, c->s64.density, m, c->g64.M)
                : get_mesh<float, double>(c, c->s32.density, m, c->g32.M);
}
int p3m_get_potential_f64(p3m_ctx* c, double* m) {
  CHECK_CTX(c);
  P3M_TRY(P3M_DISPATCH(c, complete_potential));
  return c->f64 ? get_mesh<double, double>(c, c->s64.potential, m, c->g64.M)
                : get_mesh<float, double>(c, c->s32.potential, m, c->g32.M);
}
int p3m_set_density(p3m_ctx* c, const float* m) {
  CHECK_CTX(c);
  int r = c->f64 ? set_mesh<double, float>(c, c->s64.density, m, c->g64.M)
                 : set_mesh<float, float>(c, c->s32.density, m, c->g32.M);
  if (r == 0) c->have_density = true;
  // a density that did not come from the particles: nothing is known about its empty planes
  c->s32.dens_occ = c->s64.dens_occ = 0, c->s32.dens_dirty = c->s64.dens_dirty = -1;
  return r;
}
int p3m_set_potential(p3m_ctx* c, const float* m) {
  CHECK_CTX(c);
  int r = c->f64 ? set_mesh<double, float>(c, c->s64.potential, m, c->g64.M)
                 : set_mesh<float, float>(c, c->s32.potential, m, c->g32.M);
  if (r == 0 && c->slab) r = c->f64 ? slab_spread_potential<double>(c) : slab_spread_potential<float>(c);
  if (r == 0) c->have_potential = true;
  c->s32.pot_partial = c->s64.pot_partial = false;
  return r;
}

int p3m_get_cells(p3m_ctx* c, int32_t* mesh_cell, int32_t* chain_cell, int32_t* order) {
  CHECK_CTX(c);
  if (!c->have_particles) return fail(P3M_ESTATE, "no particles set");
  return P3M_DISPATCH(c, get_cells, mesh_cell, chain_cell, order);
}

int p3m_get_chaining_dims(p3m_ctx* c, int32_t dims[3]) {
  if (!c) return fail(P3M_EINVAL, "null context");
  if (!c->prm.p3m) return fail(P3M_ESTATE, "PM-only context has no chaining mesh");
  dims[0] = c->f64 ? c->g64.mx : c->g32.mx;
  dims[1] = c->f64 ? c->g64.my : c->g32.my;
  dims[2] = c->f64 ? c->g64.mz : c->g32.mz;
  return 0;
}

int p3m_get_binning(p3m_ctx* c, int32_t out[8]) {
  if (!c || !out) return fail(P3M_EINVAL, "null argument");
  if (c->f64) {
    const Geom<double>& g = c->g64;
    const int32_t v[8] = {g.mx, g.my, g.mz, g.mbits, g.sbits, g.idbits, g.bshift, g.p3m};
    memcpy(out, v, sizeof(v));
  } else {
    const Geom<float>& g = c->g32;
    const int32_t v[8] = {g.mx, g.my, g.mz, g.mbits, g.sbits, g.idbits, g.bshift, g.p3m};
    memcpy(out, v, sizeof(v));
  }
  return 0;
}

int p3m_chaining_neighbors(const int32_t M[3], int32_t cell, int32_t nb[14]) {
  // ChainingMesh::getNeighborsAndSelf / tripleToFlatIndex, source/chainingMesh.cpp:60-84
  if (!M || !nb || cell < 0 || cell >= M[0] * M[1] * M[2]) return fail(P3M_EINVAL, "bad cell");
  const int cx = cell % M[0], cy = (cell / M[0]) % M[1], cz = cell / (M[0] * M[1]);
  auto tri = [&](int x, int y, int z) {
    if (x < 0 || y < 0 || z < 0 || x >= M[0] || y >= M[1] || z >= M[2]) return -1;
    return x + y * M[0] + z * M[0] * M[1];
  };
  int i = 0;
  for (int t = -1; t <= 1; ++t)
    for (int s = -1; s <= 1; ++s) nb[i++] = tri(cx + t, cy - 1, cz + s);
  for (int s = -1; s <= 1; ++s) nb[i++] = tri(cx + s, cy, cz - 1);
  nb[12] = tri(cx - 1, cy, cz);
  nb[13] = cell;
  return 0;
}

int p3m_get_acc_parts(p3m_ctx* c, double* acc_pm, double* acc_sr) {
  CHECK_CTX(c);
  if (!c->have_particles) return fail(P3M_ESTATE, "no particles set");
  if (c->n > 0 && !c->have_acc) return fail(P3M_ESTATE, "p3m_get_acc_parts: accelerations are stale (re-sorted since the last p3m_gather)");
  return P3M_DISPATCH(c, get_acc_parts, acc_pm, acc_sr);
}

int p3m_get_sample(p3m_ctx* c, const int32_t* ids, int64_t m, double* pos, double* acc, double* acc_sr) {
  CHECK_CTX(c);
  if (!c->have_particles) return fail(P3M_ESTATE, "no particles set");
  if (m < 0 || (m > 0 && !ids)) return fail(P3M_EINVAL, "p3m_get_sample: bad argument");
  for (int64_t k = 1; k < m; ++k)
    if (ids[k] <= ids[k - 1]) return fail(P3M_EINVAL, "p3m_get_sample: ids must be strictly ascending");
  if ((acc || acc_sr) && c->n > 0 && !c->have_acc)
    return fail(P3M_ESTATE, "p3m_get_sample: accelerations are stale (re-sorted since the last p3m_gather)");
  return P3M_DISPATCH(c, get_sample, ids, (long long)m, pos, acc, acc_sr);
}

int p3m_get_sr_table(p3m_ctx* c, double* t) {
  if (!c || !t) return fail(P3M_EINVAL, "null argument");
  if (c->sr_table_host.size() != kSRTable) return fail(P3M_ESTATE, "no short-range table (PM-only context)");
  memcpy(t, c->sr_table_host.data(), sizeof(double) * kSRTable);
  return 0;
}

static const char* kPhaseNames[P3M_NPHASE] = {"binSort",      "spreadMass",       "forwardFFT",
                                              "fourierPotential", "inverseFFT",   "updateAccelerations",
                                              "shortRangeForcesCalc", "integrate", "fieldInCells",
                                              "comm"};
const char* p3m_phase_name(int i) { return (i >= 0 && i < P3M_NPHASE) ? kPhaseNames[i] : ""; }

int p3m_get_phase_ms(p3m_ctx* c, float ms[P3M_NPHASE], int reset) {
  if (!c) return fail(P3M_EINVAL, "null context");
  phase_resolve(c);
  for (int i = 0; i < P3M_NPHASE; ++i) {
    ms[i] = c->timer.acc_ms[i];
    if (reset) c->timer.acc_ms[i] = 0;
  }
  return 0;
}

int p3m_get_pair_counts(p3m_ctx* c, uint64_t* checked, uint64_t* in_range) {
  CHECK_CTX(c);
  if (!c->prm.p3m) return fail(P3M_ESTATE, "PM-only context");
  if (!c->have_particles || !c->sorted)
    return fail(P3M_ESTATE, "p3m_get_pair_counts: call after p3m_short_range / p3m_force (particles must be sorted)");
  c->count_pairs = 1;  // counting instantiation of the same kernels: reads only, no collective
  int r = P3M_DISPATCH(c, short_range);
  c->count_pairs = 0;
  if (r != 0) return r;
  unsigned long long h[2];
  void* src = c->f64 ? (void*)c->s64.pair_counts : (void*)c->s32.pair_counts;
  P3M_CUDA(cudaMemcpyAsync(h, src, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  if (checked) *checked = h[0];
  if (in_range) *in_range = h[1];
  return 0;
}

int p3m_direct_sum(p3m_ctx* c, int mode, const double* tpos, int64_t m, double eps, double* out) {
  CHECK_CTX(c);
  if (!c->have_particles) return fail(P3M_ESTATE, "no particles set");
  if (m < 0 || (m > 0 && (!tpos || !out))) return fail(P3M_EINVAL, "p3m_direct_sum: bad argument");
  if (mode != P3M_SUM_SHORT_RANGE && mode != P3M_SUM_NEWTON && mode != P3M_SUM_CUTOFF_SHELL)
    return fail(P3M_EINVAL, "p3m_direct_sum: unknown mode %d", mode);
  if (mode != P3M_SUM_NEWTON && !c->prm.p3m) return fail(P3M_ESTATE, "p3m_direct_sum: PM-only context has no short-range law");
  return P3M_DISPATCH(c, direct_sum, mode, tpos, (long long)m, eps, out);
}

int p3m_get_stats(p3m_ctx* c, double out[P3M_NSTAT]) {
  if (!c || !out) return fail(P3M_EINVAL, "null argument");
  for (int i = 0; i < P3M_NSTAT; ++i) out[i] = 0;
  out[0] = c->fused_z, out[1] = c->slab, out[2] = c->uniform_mass && c->prm.use_sr_table, out[3] = c->packed_pp;
  out[4] = c->incr_sort, out[5] = c->stat_migrated, out[6] = (double)(c->f64 ? c->s64.n_ghost : c->s32.n_ghost);
  out[7] = c->stat_a2a_bytes, out[8] = c->stat_den_bytes, out[9] = c->stat_pot_bytes, out[10] = c->stat_mig_bytes;
  out[11] = c->stat_ghost_bytes, out[12] = c->stat_movers, out[13] = (double)c->full_sorts, out[14] = (double)c->incr_sorts;
  return 0;
}

int64_t p3m_launch_count(const p3m_ctx* c) { return c ? c->launches : 0; }
void* p3m_stream(p3m_ctx* c) { return c ? (void*)c->stream : nullptr; }
int p3m_synchronize(p3m_ctx* c) {
  CHECK_CTX(c);
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

}  // extern "C"
