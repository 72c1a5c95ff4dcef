// integrate.cu -- A10 + A11 + N1: leapfrog, the per-step unit round trip, the escape test and the
// SimInfo diagnostics, all on device-resident particles.
//
// Replaces setHalfStepVelocities / updateVelocities / updatePositions (source/leapfrog.cpp:5-24),
// the stateToOriginalUnits / stateToCodeUnits pair the run loops execute around recording
// (source/pmMethod.cpp:94,113; source/p3mMethod.cpp:108,137), PMMethod::escapedComputationalBox
// (source/pmMethod.cpp:146-156) and SimInfo (source/simInfo.cpp:50-127).
#include "ctx.cuh"

namespace p3m {

// v += f * a.  flags[0] != 0 (a particle escaped) freezes the state, as the reference's `break` does.
template <typename T>
__global__ void k_kick(V4<T>* __restrict__ vel, const V4<T>* __restrict__ acc, long long n, T f,
                       int* __restrict__ flags, int count_step) {
  if (flags[0]) return;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i == 0 && count_step) flags[2] += 1;  // completed steps
  if (i >= n) return;
  V4<T> v = vel[i];
  const V4<T> a = acc[i];
  v.x += f * a.x, v.y += f * a.y, v.z += f * a.z;  // leapfrog.cpp:7,18 (dt = 1)
  vel[i] = v;
}

// x += v, then optionally x -> H*x -> /H and v -> H*v/DT -> DT*v/H (fp rounding of the run loop),
// then the box test 0 <= x_orig <= box (source/pmMethod.cpp:146-149).
template <typename T>
__global__ void k_drift(V4<T>* __restrict__ posm, V4<T>* __restrict__ vel, long long n, Geom<T> g,
                        int* __restrict__ flags, int* __restrict__ flag_out) {
  if (flags[0]) return;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  V4<T> p = posm[i];
  V4<T> v = vel[i];
  p.x += v.x, p.y += v.y, p.z += v.z;  // leapfrog.cpp:23 (dt = 1)
  T ox = g.H * p.x, oy = g.H * p.y, oz = g.H * p.z;
  if (g.unit_roundtrip) {
    p.x = ox / g.H, p.y = oy / g.H, p.z = oz / g.H;
    v.x = g.DT * (g.H * v.x / g.DT) / g.H;
    v.y = g.DT * (g.H * v.y / g.DT) / g.H;
    v.z = g.DT * (g.H * v.z / g.DT) / g.H;
    vel[i] = v;
  }
  posm[i] = p;
  const bool in = ox >= 0 && ox <= g.boxx && oy >= 0 && oy <= g.boxy && oz >= 0 && oz <= g.boxz;
  if (!in) *flag_out = 1;  // published to flags[0] by the next kernel boundary
}

__global__ void k_publish_escape(int* flags) {
  if (flags[3]) flags[0] = 1;
}

template <typename T>
__global__ void k_escaped(const V4<T>* __restrict__ posm, long long n, Geom<T> g, int* __restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V4<T> p = posm[i];
  const T ox = g.H * p.x, oy = g.H * p.y, oz = g.H * p.z;
  const bool in = ox >= 0 && ox <= g.boxx && oy >= 0 && oy <= g.boxy && oz >= 0 && oz <= g.boxz;
  if (!in) *out = 1;
}

template <typename T>
int kick(p3m_ctx* c, double f) {
  if (!c->have_particles) return fail(P3M_ESTATE, "p3m_kick: no particles set");
  State<T>& s = Sel<T>::st(c);
  if (c->n == 0) return 0;
  phase_begin(c, PH_INTEGRATE);
  k_kick<T><<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(s.vel, s.acc, c->n, (T)f, s.flags,
                                                                  f == 1.0 ? 1 : 0);
  P3M_LAUNCH_CHECK(c);
  phase_end(c, PH_INTEGRATE);
  return 0;
}

template <typename T>
int drift(p3m_ctx* c) {
  if (!c->have_particles) return fail(P3M_ESTATE, "p3m_drift: no particles set");
  State<T>& s = Sel<T>::st(c);
  if (c->n == 0 && c->nranks == 1) return 0;
  phase_begin(c, PH_INTEGRATE);
  if (c->n > 0) {
    k_drift<T><<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(s.posm, s.vel, c->n, Sel<T>::g(c),
                                                                     s.flags, s.flags + 3);
    P3M_LAUNCH_CHECK(c);
  }
  phase_end(c, PH_INTEGRATE);
  if (c->nranks > 1) {  // every rank must take the same decision; the wait for the slowest rank is `comm` time
    phase_begin(c, PH_COMM);
    P3M_TRY(dist_allreduce(c, s.flags + 3, 1, 0));
    phase_end(c, PH_COMM);
  }
  k_publish_escape<<<1, 1, 0, c->stream>>>(s.flags);
  P3M_LAUNCH_CHECK(c);
  c->sorted = false;
  return 0;
}

// ---- SimInfo on the device ----------------------------------------------------------------------------
template <typename T>
__device__ inline double ext_potential(const Geom<T>& g, double x, double y, double z) {
  // sphRadDecrFieldPotential, source/externalFields.cpp:17-24
  if (g.ext_kind != P3M_EXT_SPH_RAD_DECR) return 0.0;
  const double dx = x - (double)g.ecx, dy = y - (double)g.ecy, dz = z - (double)g.ecz;
  const double R = (double)g.eR, M = (double)g.eM, G = (double)g.G;
  const double r = sqrt(dx * dx + dy * dy + dz * dz);
  if (r > R) return -G * M / r;
  const double u = r / R;
  return G * M / R * (-2 + u * u * (2 - u));
}

template <int NV>
__device__ inline void block_accumulate(double (&v)[NV], double* __restrict__ out) {
  __shared__ double red[NV][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[k][w] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[threadIdx.x][i];
    atomicAdd(&out[threadIdx.x], t);
  }
}

// out: [0] external PE, [1] KE, [2..4] p, [5..7] L, [8..10] external force  (original units)
template <typename T>
__global__ void __launch_bounds__(256)
k_diag_particles(const V4<T>* __restrict__ posm, const V4<T>* __restrict__ vel,
                 const V4<T>* __restrict__ acc, long long n, Geom<T> g, double mass_factor,
                 double* __restrict__ out) {
  double v[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) v[k] = 0;
  const double H = (double)g.H, DT = (double)g.DT;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const V4<T> p = posm[i], hv = vel[i], a = acc[i];
    const double m = (double)p.w / mass_factor;  // massToOriginalUnits
    const double x = H * (double)p.x, y = H * (double)p.y, z = H * (double)p.z;
    const double s = H / DT;
    const double vx = s * (double)hv.x, vy = s * (double)hv.y, vz = s * (double)hv.z;
    // integer-step velocity v + a/2 (source/leapfrog.cpp:10-14), original units
    const double ix = s * ((double)hv.x + 0.5 * (double)a.x), iy = s * ((double)hv.y + 0.5 * (double)a.y),
                 iz = s * ((double)hv.z + 0.5 * (double)a.z);
    v[0] += m * ext_potential(g, x, y, z);
    v[1] += 0.5 * m * (vx * vx + vy * vy + vz * vz);  // half-step velocity, simInfo.cpp:94-101
    v[2] += m * ix, v[3] += m * iy, v[4] += m * iz;    // simInfo.cpp:103-109
    v[5] += m * (y * iz - z * iy), v[6] += m * (z * ix - x * iz), v[7] += m * (x * iy - y * ix);
    if (g.ext_kind == P3M_EXT_SPH_RAD_DECR) {  // totalExternalForceOrigUnits, pmMethod.cpp:158-162
      const double dx = x - (double)g.ecx, dy = y - (double)g.ecy, dz = z - (double)g.ecz;
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      const double R = (double)g.eR, M = (double)g.eM, G = (double)g.G;
      const double gg = r > R ? -G * M / (r * r) : -(G * M / (R * R * R)) * r * (4 - 3 * r / R);
      v[8] += m * gg * dx / r, v[9] += m * gg * dy / r, v[10] += m * gg * dz / r;
    }
  }
  block_accumulate<11>(v, out);
}

// out[11] += sum rho * phi (code units)
template <typename T>
__global__ void __launch_bounds__(256)
k_diag_mesh(const T* __restrict__ rho, const T* __restrict__ phi, long long first, long long M,
            double* __restrict__ out) {
  double v[1] = {0};
  for (long long i = first + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M;
       i += (long long)gridDim.x * blockDim.x)
    v[0] += (double)rho[i] * (double)phi[i];
  block_accumulate<1>(v, out + 11);
}

template <typename T>
int diagnostics(p3m_ctx* c, double* out) {
  if (!c->have_particles) return fail(P3M_ESTATE, "p3m_diagnostics: no particles set");
  if (c->n > 0 && !c->have_acc)
    return fail(P3M_ESTATE, "p3m_diagnostics: accelerations are stale (re-sorted since the last p3m_gather)");
  State<T>& s = Sel<T>::st(c);
  const Geom<T>& g = Sel<T>::g(c);
  P3M_CUDA(cudaMemsetAsync(s.diag, 0, sizeof(double) * 16, c->stream));
  const double mf = c->f64 ? c->mass_factor64 : (double)c->mass_factor32;
  const int blocks = c->num_sms * 4;
  if (c->n > 0) {
    k_diag_particles<T><<<blocks, 256, 0, c->stream>>>(s.posm, s.vel, s.acc, c->n, g, mf, s.diag);
    P3M_LAUNCH_CHECK(c);
  }
  if (c->have_density && c->have_potential) {
    // the mesh is replicated on every rank: each sums its share of the cells
    // (slab-decomposed mesh: density / potential ARE this rank's planes)
    const long long share = (g.M + c->nranks - 1) / c->nranks;
    const long long first = c->slab ? 0 : share * c->rank;
    const long long last = c->slab ? g.M / c->nranks : (first + share < g.M ? first + share : g.M);
    k_diag_mesh<T><<<blocks, 256, 0, c->stream>>>(s.density, s.potential, first, last, s.diag);
    P3M_LAUNCH_CHECK(c);
  }
  P3M_TRY(dist_allreduce(c, s.diag, 16, 1));
  double h[16];
  P3M_CUDA(cudaMemcpyAsync(h, s.diag, sizeof(double) * 16, cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  // PE = 0.5 * H^3 * sum rho_orig * phi_orig + external   (source/simInfo.cpp:56-69)
  const double H = (double)g.H, DT = (double)g.DT, G = (double)g.G;
  const double pi = 3.14159265358979323846;
  const double internal = h[11] / (DT * DT * 4 * pi * G) * (H * H / (DT * DT));
  out[0] = 0.5 * H * H * H * internal + h[0];
  for (int k = 1; k < 11; ++k) out[k] = h[k];
  return 0;
}

template <typename T>
__global__ void k_acc_parts(const V4<T>* __restrict__ acc, const V4<T>* __restrict__ acc_sr,
                            const int* __restrict__ id, long long n, double* __restrict__ pm,
                            double* __restrict__ sr) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long j = id[i];
  const V4<T> a = acc[i], b = acc_sr[i];
  if (pm) pm[3 * j] = (double)a.x - (double)b.x, pm[3 * j + 1] = (double)a.y - (double)b.y, pm[3 * j + 2] = (double)a.z - (double)b.z;
  if (sr) sr[3 * j] = (double)b.x, sr[3 * j + 1] = (double)b.y, sr[3 * j + 2] = (double)b.z;
}

template <typename T>
int get_acc_parts(p3m_ctx* c, double* acc_pm, double* acc_sr) {
  State<T>& s = Sel<T>::st(c);
  const long long nl = c->n;
  const long long n = c->nranks > 1 ? c->n_global : nl;  // id-indexed outputs
  if (n == 0) return 0;
  double* stage = nullptr;
  P3M_CUDA(cudaMallocAsync((void**)&stage, sizeof(double) * 6 * (size_t)n, c->stream));
  if (c->nranks > 1) P3M_CUDA(cudaMemsetAsync(stage, 0, sizeof(double) * 6 * (size_t)n, c->stream));
  if (nl > 0) {
    k_acc_parts<T><<<(unsigned)((nl + 255) / 256), 256, 0, c->stream>>>(
        s.acc, s.acc_sr, s.id, nl, acc_pm ? stage : nullptr, acc_sr ? stage + 3 * n : nullptr);
    P3M_LAUNCH_CHECK(c);
  }
  if (acc_pm) P3M_CUDA(cudaMemcpyAsync(acc_pm, stage, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  if (acc_sr) P3M_CUDA(cudaMemcpyAsync(acc_sr, stage + 3 * n, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaFreeAsync(stage, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// rows of the particles whose id is in the ascending list `ids` (binary search per local particle)
template <typename T>
__global__ void k_sample_rows(const V4<T>* __restrict__ posm, const V4<T>* __restrict__ acc,
                              const V4<T>* __restrict__ acc_sr, const int* __restrict__ id, long long n,
                              const int* __restrict__ ids, int m, double* __restrict__ pos_o, double* __restrict__ acc_o,
                              double* __restrict__ sr_o) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int me = id[i];
  int lo = 0, hi = m;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ids[mid] < me) lo = mid + 1; else hi = mid;
  }
  if (lo >= m || ids[lo] != me) return;
  if (pos_o) {
    const V4<T> p = posm[i];
    pos_o[3 * lo] = (double)p.x, pos_o[3 * lo + 1] = (double)p.y, pos_o[3 * lo + 2] = (double)p.z;
  }
  if (acc_o) {
    const V4<T> a = acc[i];
    acc_o[3 * lo] = (double)a.x, acc_o[3 * lo + 1] = (double)a.y, acc_o[3 * lo + 2] = (double)a.z;
  }
  if (sr_o) {
    const V4<T> a = acc_sr[i];
    sr_o[3 * lo] = (double)a.x, sr_o[3 * lo + 1] = (double)a.y, sr_o[3 * lo + 2] = (double)a.z;
  }
}

template <typename T>
int get_sample(p3m_ctx* c, const int32_t* ids, long long m, double* pos, double* acc, double* acc_sr) {
  State<T>& s = Sel<T>::st(c);
  if (m == 0) return 0;
  int* d_ids = nullptr;
  double* d_out = nullptr;
  P3M_CUDA(cudaMallocAsync((void**)&d_ids, sizeof(int) * (size_t)m, c->stream));
  P3M_CUDA(cudaMallocAsync((void**)&d_out, sizeof(double) * 9 * (size_t)m, c->stream));
  P3M_CUDA(cudaMemcpyAsync(d_ids, ids, sizeof(int) * (size_t)m, cudaMemcpyHostToDevice, c->stream));
  P3M_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * 9 * (size_t)m, c->stream));
  if (c->n > 0) {
    k_sample_rows<T><<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(
        s.posm, s.acc, s.acc_sr, s.id, c->n, d_ids, (int)m, pos ? d_out : nullptr, acc ? d_out + 3 * m : nullptr,
        acc_sr ? d_out + 6 * m : nullptr);
    P3M_LAUNCH_CHECK(c);
  }
  if (pos) P3M_CUDA(cudaMemcpyAsync(pos, d_out, sizeof(double) * 3 * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
  if (acc) P3M_CUDA(cudaMemcpyAsync(acc, d_out + 3 * m, sizeof(double) * 3 * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
  if (acc_sr) P3M_CUDA(cudaMemcpyAsync(acc_sr, d_out + 6 * m, sizeof(double) * 3 * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaFreeAsync(d_ids, c->stream));
  P3M_CUDA(cudaFreeAsync(d_out, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

template <typename T>
int escaped_now(p3m_ctx* c, int* escaped) {
  State<T>& s = Sel<T>::st(c);
  P3M_CUDA(cudaMemsetAsync(s.pp_counters + 4, 0, sizeof(int), c->stream));
  if (c->n > 0) {
    k_escaped<T><<<(unsigned)((c->n + 255) / 256), 256, 0, c->stream>>>(s.posm, c->n, Sel<T>::g(c),
                                                                       s.pp_counters + 4);
    P3M_LAUNCH_CHECK(c);
  }
  P3M_TRY(dist_allreduce(c, s.pp_counters + 4, 1, 0));
  P3M_CUDA(cudaMemcpyAsync(escaped, s.pp_counters + 4, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

template int kick<float>(p3m_ctx*, double);
template int kick<double>(p3m_ctx*, double);
template int drift<float>(p3m_ctx*);
template int drift<double>(p3m_ctx*);
template int diagnostics<float>(p3m_ctx*, double*);
template int diagnostics<double>(p3m_ctx*, double*);
template int get_acc_parts<float>(p3m_ctx*, double*, double*);
template int get_acc_parts<double>(p3m_ctx*, double*, double*);
template int get_sample<float>(p3m_ctx*, const int32_t*, long long, double*, double*, double*);
template int get_sample<double>(p3m_ctx*, const int32_t*, long long, double*, double*, double*);
template int escaped_now<float>(p3m_ctx*, int*);
template int escaped_now<double>(p3m_ctx*, int*);

}  // namespace p3m
