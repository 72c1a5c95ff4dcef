// poisson.cu -- A2 + A3 + A4: the Poisson solve.
//
//   A4  PMMethod::initGreensFunction + GreenOptimal / GreenDiscreteLaplacian / GreenPoorMan
//       (source/pmMethod.cpp:164-185, source/greensFunctions.cpp:8-220) evaluated on the device.
//       The reference needs ~17 us per cell on a CPU core for the optimal influence function (125
//       alias terms); it is infeasible there for >= 512^3.
//   A2  Grid::fftDensity / invFftPotential through FFTAdapter<float> (source/grid.cpp:50-56): the
//       reference runs a full complex-to-complex transform on a real density; here cuFFT R2C / C2R on
//       the half spectrum (half the bytes, half the flops).
//   A3  PMMethod::findFourierPotential (source/pmMethod.cpp:340-350): one in-place multiply of the half
//       spectrum by the real table; the adapters' 1/length of the inverse transform
//       (include/kissFFTAdapter.h:22-28) is folded into the table.
//
// R2C vs C2C: the reference takes .real() of a C2C inverse (source/grid.cpp:66-68) and its
// table is not exactly Hermitian (alias sum truncated around un-centred k), which is equivalent to the
// real operator G_sym(k) = (G(k) + G(-k mod N)) / 2.  The table stored here is G_sym / M.
//
// The table is always EVALUATED in fp64 (once per run) and rounded to the mesh precision.
#include <cstdlib>

#include "ctx.cuh"

namespace p3m {

namespace {

struct GreenCfg {
  int nx, ny, nz;
  int is, fds, gfunc;
  double a;  // particle diameter, code units (source/pmMethod.cpp:56)
  int zero_degenerate;
};

__device__ inline double d_sinc(double x) { return x == 0.0 ? 1.0 : sin(x) / x; }

__device__ double green_optimal(const GreenCfg& c, int kx, int ky, int kz) {
  // source/greensFunctions.cpp:122-189, restated on the non-zero lanes of its complex arithmetic:
  // D = i d, R_n = -i k_n S^2(|k_n| a / 2) / |k_n|^2  =>  G = sum_i d_i r_i / (|d|^2 (sum U^2)^2)
  if (kx == 0 && ky == 0 && kz == 0) return 0.0;
  if (c.zero_degenerate && (2 * kx) % c.nx == 0 && (2 * ky) % c.ny == 0 && (2 * kz) % c.nz == 0)
    return 0.0;  // D vanishes identically: 0/0 in the reference
  const double pi = 3.14159265358979323846;
  const double k[3] = {2 * pi * kx / c.nx, 2 * pi * ky / c.ny, 2 * pi * kz / c.nz};
  double denomSum = 1.0;
  if (c.is == P3M_TSC) {  // :102-108
    for (int i = 0; i < 3; ++i) {
      const double s = sin(k[i] / 2), s2 = s * s;
      denomSum *= (1 - s2 + 2.0 / 15 * s2 * s2);
    }
  } else if (c.is == P3M_CIC) {  // :110-116
    for (int i = 0; i < 3; ++i) {
      const double co = cos(k[i] / 2);
      denomSum *= (1 + 2 * co * co);
    }
    denomSum *= 1.0 / 27;
  }
  double d[3];
  for (int i = 0; i < 3; ++i) {
    if (c.fds == P3M_TWO_POINT)
      d[i] = sin(k[i]);  // :70-78
    else
      d[i] = 4.0 / 3 * sin(k[i]) + (1 - 4.0 / 3) * sin(2 * k[i]) / 2.0;  // :80-89
  }
  const double dnorm = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const int pw = c.is == P3M_TSC ? 3 : (c.is == P3M_CIC ? 2 : 1);
  const bool s2cloud = c.gfunc == P3M_S2_OPTIMAL;
  double num[3] = {0, 0, 0};
  // per-axis factors sinc^(2p)(k_n/2) are separable: precompute the 5 aliases of each axis
  double kn[3][5], u2[3][5];
  for (int i = 0; i < 3; ++i)
    for (int n = -2; n <= 2; ++n) {
      const double kk = k[i] + 2 * pi * n;
      const double s = d_sinc(kk / 2);
      double w = s;
      for (int q = 1; q < pw; ++q) w *= s;
      kn[i][n + 2] = kk;
      u2[i][n + 2] = w * w;
    }
  for (int n1 = 0; n1 < 5; ++n1)
    for (int n2 = 0; n2 < 5; ++n2)
      for (int n3 = 0; n3 < 5; ++n3) {
        const double usq = u2[0][n1] * u2[1][n2] * u2[2][n3];  // (prod sinc^p)^2, :166-172
        const double k2 = kn[0][n1] * kn[0][n1] + kn[1][n2] * kn[1][n2] + kn[2][n3] * kn[2][n3];
        const double kl = sqrt(k2);
        const double u = kl * c.a / 2;
        double s;
        if (!s2cloud)
          s = -3 / (u * u * u) * (u * cos(u) - sin(u));  // S1Fourier :42-45
        else
          s = 12 / (u * u * u * u) * (2 - 2 * cos(u) - u * sin(u));  // S2Fourier :47-50
        const double f = usq * s * s / k2;  // :62-64 and :178-180
        num[0] += -kn[0][n1] * f, num[1] += -kn[1][n2] * f, num[2] += -kn[2][n3] * f;
      }
  const double numerator = d[0] * num[0] + d[1] * num[1] + d[2] * num[2];
  return numerator / (dnorm * denomSum * denomSum);
}

__device__ double green_value(const GreenCfg& c, int kx, int ky, int kz) {
  const double pi = 3.14159265358979323846;
  if (c.gfunc == P3M_DISCRETE_LAPLACIAN) {  // :191-200
    if (kx == 0 && ky == 0 && kz == 0) return 0.0;
    const double sx = sin(pi * kx / c.nx), sy = sin(pi * ky / c.ny), sz = sin(pi * kz / c.nz);
    return -0.25 / (sx * sx + sy * sy + sz * sz);
  }
  if (c.gfunc == P3M_POOR_MAN) {  // :202-220
    if (kx == 0 && ky == 0 && kz == 0) return 0.0;
    const int ki = (kx <= c.nx / 2) ? kx : kx - c.nx;
    const int kj = (ky <= c.ny / 2) ? ky : ky - c.ny;
    const int kk = (kz <= c.nz / 2) ? kz : kz - c.nz;
    const double a = 2 * pi * ki / c.nx, b = 2 * pi * kj / c.ny, cc = 2 * pi * kk / c.nz;
    return -1.0 / (a * a + b * b + cc * cc);
  }
  return green_optimal(c, kx, ky, kz);
}

template <typename T>
__global__ void k_green(GreenCfg c, long long half_count, T* __restrict__ table) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= half_count) return;
  const int nxh = c.nx / 2 + 1;
  const int kx = (int)(idx % nxh), ky = (int)((idx / nxh) % c.ny), kz = (int)(idx / ((long long)nxh * c.ny));
  const double g1 = green_value(c, kx, ky, kz);
  const double g2 = green_value(c, (c.nx - kx) % c.nx, (c.ny - ky) % c.ny, (c.nz - kz) % c.nz);
  const double M = (double)c.nx * c.ny * c.nz;
  table[idx] = (T)(0.5 * (g1 + g2) / M);
}

template <typename T, typename I>
__global__ void k_green_from_full(const I* __restrict__ full, int nx, int ny, int nz,
                                  long long half_count, T* __restrict__ table) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= half_count) return;
  const int nxh = nx / 2 + 1;
  const int kx = (int)(idx % nxh), ky = (int)((idx / nxh) % ny), kz = (int)(idx / ((long long)nxh * ny));
  const long long a = kx + (long long)ky * nx + (long long)kz * nx * ny;
  const long long b = (nx - kx) % nx + (long long)((ny - ky) % ny) * nx + (long long)((nz - kz) % nz) * nx * ny;
  const double M = (double)nx * ny * nz;
  table[idx] = (T)(0.5 * ((double)full[a] + (double)full[b]) / M);
}

template <typename T>
__global__ void k_green_to_full(const T* __restrict__ table, int nx, int ny, int nz,
                                double* __restrict__ full) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long M = (long long)nx * ny * nz;
  if (idx >= M) return;
  int kx = (int)(idx % nx), ky = (int)((idx / nx) % ny), kz = (int)(idx / ((long long)nx * ny));
  const int nxh = nx / 2 + 1;
  if (kx >= nxh) kx = nx - kx, ky = (ny - ky) % ny, kz = (nz - kz) % nz;
  full[idx] = (double)table[kx + (long long)ky * nxh + (long long)kz * nxh * ny] * (double)M;
}

// slab-decomposed mesh: this rank's share of the table in the transposed layout [kx, ky_local, kz]
template <typename T>
__global__ void k_green_slab(GreenCfg c, int ky0, int nyl, long long count, T* __restrict__ table) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= count) return;
  const int nxh = c.nx / 2 + 1;
  const int kx = (int)(idx % nxh), ky = ky0 + (int)((idx / nxh) % nyl), kz = (int)(idx / ((long long)nxh * nyl));
  const double g1 = green_value(c, kx, ky, kz);
  const double g2 = green_value(c, (c.nx - kx) % c.nx, (c.ny - ky) % c.ny, (c.nz - kz) % c.nz);
  const double M = (double)c.nx * c.ny * c.nz;
  table[idx] = (T)(0.5 * (g1 + g2) / M);
}

template <typename T, typename I>
__global__ void k_green_slab_from_full(const I* __restrict__ full, int nx, int ny, int nz, int ky0, int nyl,
                                       long long count, T* __restrict__ table) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= count) return;
  const int nxh = nx / 2 + 1;
  const int kx = (int)(idx % nxh), ky = ky0 + (int)((idx / nxh) % nyl), kz = (int)(idx / ((long long)nxh * nyl));
  const long long a = kx + (long long)ky * nx + (long long)kz * nx * ny;
  const long long b = (nx - kx) % nx + (long long)((ny - ky) % ny) * nx + (long long)((nz - kz) % nz) * nx * ny;
  const double M = (double)nx * ny * nz;
  table[idx] = (T)(0.5 * ((double)full[a] + (double)full[b]) / M);
}

template <typename C, typename T>
__global__ void k_multiply(C* __restrict__ spec, const T* __restrict__ table, long long count) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= count) return;
  C v = spec[idx];
  const T g = table[idx];
  v.x *= g, v.y *= g;
  spec[idx] = v;
}

inline long long half_count(const p3m_params& p) {
  return (long long)(p.nx / 2 + 1) * p.ny * p.nz;
}

}  // namespace

template <typename T>
int green_init(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  const p3m_params& p = c->prm;
  GreenCfg cfg{p.nx, p.ny, p.nz, p.assignment, p.fd_scheme, p.greens_function,
               (double)p.particle_diameter / (double)p.H, p.green_zero_degenerate};
  if (!c->f64) cfg.a = (double)(p.particle_diameter / p.H);  // lengthToCodeUnits in fp32 (:56)
  const long long hc = half_count(p);
  if (c->slab) {
    const int nyl = p.ny / c->nranks;
    const long long cnt = hc / c->nranks;
    k_green_slab<T><<<(unsigned)((cnt + 127) / 128), 128, 0, c->stream>>>(cfg, c->rank * nyl, nyl, cnt, s.green);
  } else {
    k_green<T><<<(unsigned)((hc + 127) / 128), 128, 0, c->stream>>>(cfg, hc, s.green);
  }
  P3M_LAUNCH_CHECK(c);
  c->have_green = true;
  return 0;
}

template <typename T, typename I>
int green_set(p3m_ctx* c, const I* full) {
  State<T>& s = Sel<T>::st(c);
  const p3m_params& p = c->prm;
  const long long M = (long long)p.nx * p.ny * p.nz, hc = half_count(p);
  I* stage = nullptr;
  P3M_CUDA(cudaMallocAsync((void**)&stage, sizeof(I) * (size_t)M, c->stream));
  P3M_CUDA(cudaMemcpyAsync(stage, full, sizeof(I) * (size_t)M, cudaMemcpyHostToDevice, c->stream));
  if (c->slab) {
    const int nyl = p.ny / c->nranks;
    const long long cnt = hc / c->nranks;
    k_green_slab_from_full<T, I><<<(unsigned)((cnt + 255) / 256), 256, 0, c->stream>>>(
        stage, p.nx, p.ny, p.nz, c->rank * nyl, nyl, cnt, s.green);
  } else {
    k_green_from_full<T, I><<<(unsigned)((hc + 255) / 256), 256, 0, c->stream>>>(stage, p.nx, p.ny, p.nz,
                                                                               hc, s.green);
  }
  P3M_LAUNCH_CHECK(c);
  P3M_CUDA(cudaFreeAsync(stage, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  c->have_green = true;
  return 0;
}

template <typename T>
int green_get(p3m_ctx* c, double* full) {
  if (!c->have_green) return fail(P3M_ESTATE, "p3m_get_green_table: table not initialised");
  if (c->slab) return fail(P3M_ESTATE, "p3m_get_green_table: the table is distributed (slab-decomposed mesh)");
  State<T>& s = Sel<T>::st(c);
  const p3m_params& p = c->prm;
  const long long M = (long long)p.nx * p.ny * p.nz;
  double* stage = nullptr;
  P3M_CUDA(cudaMallocAsync((void**)&stage, sizeof(double) * (size_t)M, c->stream));
  k_green_to_full<T><<<(unsigned)((M + 255) / 256), 256, 0, c->stream>>>(s.green, p.nx, p.ny, p.nz, stage);
  P3M_LAUNCH_CHECK(c);
  P3M_CUDA(cudaMemcpyAsync(full, stage, sizeof(double) * (size_t)M, cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaFreeAsync(stage, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

static cufftResult exec_fwd(cufftHandle p, float* in, cufftComplex* out) { return cufftExecR2C(p, in, out); }
static cufftResult exec_fwd(cufftHandle p, double* in, cufftDoubleComplex* out) { return cufftExecD2Z(p, in, out); }
static cufftResult exec_inv(cufftHandle p, cufftComplex* in, float* out) { return cufftExecC2R(p, in, out); }
static cufftResult exec_inv(cufftHandle p, cufftDoubleComplex* in, double* out) { return cufftExecZ2D(p, in, out); }

// planes per batched 2-D transform.  Measured on B200 (512^3): splitting the batch so that the intermediate of
// cuFFT's two passes would stay in L2 is SLOWER than one batch over all planes (2.39 vs 1.73 ms per solve),
// so the default is one batch; P3M_TUNE_FFT_CHUNK_MB re-enables the split for measurements.
int fft_chunk_planes(long long plane_bytes, int planes, long long budget) {
  int cp = 1;
  while (cp * 2 <= planes && (long long)(cp * 2) * plane_bytes <= budget && planes % (cp * 2) == 0) cp *= 2;
  return cp;
}

template <typename T>
int alloc_meshes(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  Geom<T>& g = Sel<T>::g(c);
  const p3m_params& p = c->prm;
  const long long hc = half_count(p);
  // several ranks: slab-decomposed mesh whenever the planes and the ky rows divide evenly
  c->slab = c->nranks > 1 && p.nz % c->nranks == 0 && p.ny % c->nranks == 0 && !c->tune.replicated_mesh;
  // every rank owns a run of planes in the occupied half of the mesh and one in the zero-padding half (dist_mesh.cu)
  c->slab_split = c->slab && (p.nz / c->nranks) % 2 == 0 && !c->tune.contig_slabs;
  c->fused_z = fused_z_supported(p.nz) && !c->tune.cufft_z;
  if (c->fused_z) P3M_TRY(fused_z_init<T>(c));
  if (c->slab) {
    P3M_TRY(slab_setup<T>(c));
  } else {
    P3M_CUDA(cudaMalloc((void**)&s.density, sizeof(T) * (size_t)g.M));
    P3M_CUDA(cudaMalloc((void**)&s.potential, sizeof(T) * (size_t)g.M));
    P3M_CUDA(cudaMalloc((void**)&s.spectrum, sizeof(typename State<T>::cplx) * (size_t)hc));
    P3M_CUDA(cudaMalloc((void**)&s.green, sizeof(T) * (size_t)hc));
    s.dens_part = s.density, s.pot_part = s.potential;
    g.den_off = 0, g.den_len = g.M, g.pot_z0 = 0, g.pot_nz = g.nz;
    P3M_CUDA(cudaMemsetAsync(s.density, 0, sizeof(T) * (size_t)g.M, c->stream));
    P3M_CUDA(cudaMemsetAsync(s.potential, 0, sizeof(T) * (size_t)g.M, c->stream));
    const bool dbl = sizeof(T) == 8;
    if (c->fused_z) {
      // batched 2-D transforms of the planes; the z leg is k_poisson_z
      s.fft_chunk = fft_chunk_planes((long long)(p.nx / 2 + 1) * p.ny * sizeof(typename State<T>::cplx), p.nz, c->tune.fft_chunk_bytes);
      int n2[2] = {p.ny, p.nx};
      P3M_FFT(cufftPlanMany(&s.plan_fwd, 2, n2, nullptr, 1, 0, nullptr, 1, 0, dbl ? CUFFT_D2Z : CUFFT_R2C, s.fft_chunk));
      P3M_FFT(cufftPlanMany(&s.plan_inv, 2, n2, nullptr, 1, 0, nullptr, 1, 0, dbl ? CUFFT_Z2D : CUFFT_C2R, s.fft_chunk));
    } else {
      // adapters get dims {Nz, Ny, Nx} (source/demos.cpp:758-759): x is the fastest axis
      P3M_FFT(cufftPlan3d(&s.plan_fwd, p.nz, p.ny, p.nx, dbl ? CUFFT_D2Z : CUFFT_R2C));
      P3M_FFT(cufftPlan3d(&s.plan_inv, p.nz, p.ny, p.nx, dbl ? CUFFT_Z2D : CUFFT_C2R));
    }
    s.plans = true;
    P3M_FFT(cufftSetStream(s.plan_fwd, c->stream));
    P3M_FFT(cufftSetStream(s.plan_inv, c->stream));
  }
  P3M_CUDA(cudaMalloc((void**)&s.cell_start, sizeof(int) * (((size_t)1 << (3 * g.mbits)) + 2)));
  P3M_CUDA(cudaMalloc((void**)&s.sr_table, sizeof(T) * 2 * kSRTable));
  P3M_CUDA(cudaMalloc((void**)&s.pp_counters, sizeof(int) * 8));
  P3M_CUDA(cudaMalloc((void**)&s.pair_counts, sizeof(unsigned long long) * 4));  // [2]: estimated dense-cell pairs
  P3M_CUDA(cudaMalloc((void**)&s.flags, sizeof(int) * 4));
  P3M_CUDA(cudaMalloc((void**)&s.diag, sizeof(double) * 16));
  P3M_CUDA(cudaMemsetAsync(s.flags, 0, sizeof(int) * 4, c->stream));
  P3M_CUDA(cudaMemsetAsync(s.pair_counts, 0, sizeof(unsigned long long) * 2, c->stream));
  return 0;
}

// batched 2-D plan over `batch` planes (kind 0: R2C, 1: C2R), cached per (kind, batch)
template <typename T>
static int batch_plan(p3m_ctx* c, int kind, int batch, cufftHandle* out) {
  State<T>& s = Sel<T>::st(c);
  for (const auto& bp : s.batch_plans)
    if (bp.kind == kind && bp.batch == batch) {
      *out = bp.h;
      return 0;
    }
  const bool dbl = sizeof(T) == 8;
  int n2[2] = {c->prm.ny, c->prm.nx};
  cufftHandle h;
  P3M_FFT(cufftPlanMany(&h, 2, n2, nullptr, 1, 0, nullptr, 1, 0,
                        kind == 0 ? (dbl ? CUFFT_D2Z : CUFFT_R2C) : (dbl ? CUFFT_Z2D : CUFFT_C2R), batch));
  P3M_FFT(cufftSetStream(h, c->stream));
  s.batch_plans.push_back({kind, batch, h});
  *out = h;
  return 0;
}

// the pruned solve transformed back only the planes the gather reads; a caller that wants the whole potential mesh
// (readback, explicit gradient) gets the remaining planes now -- the spectrum still holds them
template <typename T>
int complete_potential(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  if (!s.pot_partial) return 0;
  const p3m_params& p = c->prm;
  const size_t plane = (size_t)p.nx * p.ny, splane = (size_t)(p.nx / 2 + 1) * p.ny;
  const int mid = p.nz - s.pot_tail - s.pot_lo;
  if (mid > 0) {
    cufftHandle pm;
    P3M_TRY(batch_plan<T>(c, 1, mid, &pm));
    P3M_FFT(exec_inv(pm, s.spectrum + splane * (size_t)s.pot_lo, s.potential + plane * (size_t)s.pot_lo));
    c->launches += 2;
  }
  s.pot_partial = false;
  return 0;
}

template <typename T>
int poisson(p3m_ctx* c) {
  if (!c->have_green) return fail(P3M_ESTATE, "p3m_poisson: call p3m_green_init first");
  if (!c->have_density) return fail(P3M_ESTATE, "p3m_poisson: no density (call p3m_deposit)");
  State<T>& s = Sel<T>::st(c);
  const long long hc = half_count(c->prm);
  if (c->slab) {
    P3M_TRY(slab_poisson<T>(c));
    P3M_TRY(slab_spread_potential<T>(c));
    c->have_potential = true;
    return 0;
  }
  if (c->fused_z) {
    const p3m_params& p = c->prm;
    const size_t plane = (size_t)p.nx * p.ny, splane = (size_t)(p.nx / 2 + 1) * p.ny;
    // pruned z range (State::zocc): planes >= zocc hold no density -- their 2-D transform is zero and is neither
    // computed nor read by the z pass; of the result only the planes the gather touches are transformed back
    const int FD = c->prm.fd_scheme == P3M_TWO_POINT ? 1 : 2;
    const int lo = s.dens_occ + FD + 2, tail = FD + 2;
    bool prune = s.dens_occ > 0 && lo + tail + 16 <= p.nz && !c->tune.no_prune;
    s.pot_partial = false;
    cufftHandle pf = 0, pl = 0, pt = 0;
    // a plan that cannot be made (work-area memory) is no reason to fail the solve: transform every plane instead
    if (prune && (batch_plan<T>(c, 0, s.dens_occ, &pf) != 0 || batch_plan<T>(c, 1, lo, &pl) != 0 ||
                  batch_plan<T>(c, 1, tail, &pt) != 0)) {
      prune = false;
      cudaGetLastError();
    }
    if (prune) {
      phase_begin(c, PH_FFT_FWD);
      P3M_FFT(exec_fwd(pf, s.density, s.spectrum));
      c->launches += 2;
      phase_end(c, PH_FFT_FWD);
      phase_begin(c, PH_MULTIPLY);  // forward z FFT + multiply + inverse z FFT; input planes >= zocc taken as zero
      P3M_TRY(fused_z_pass<T>(c, s.spectrum, s.green, (long long)splane, s.dens_occ));
      phase_end(c, PH_MULTIPLY);
      phase_begin(c, PH_FFT_INV);
      P3M_FFT(exec_inv(pl, s.spectrum, s.potential));
      P3M_FFT(exec_inv(pt, s.spectrum + splane * (size_t)(p.nz - tail), s.potential + plane * (size_t)(p.nz - tail)));
      c->launches += 4;
      phase_end(c, PH_FFT_INV);
      s.pot_partial = true, s.pot_lo = lo, s.pot_tail = tail;
      c->have_potential = true;
      return 0;
    }
    phase_begin(c, PH_FFT_FWD);
    for (int z = 0; z < p.nz; z += s.fft_chunk) {
      P3M_FFT(exec_fwd(s.plan_fwd, s.density + plane * z, s.spectrum + splane * z));
      c->launches += 2;
    }
    phase_end(c, PH_FFT_FWD);
    phase_begin(c, PH_MULTIPLY);  // forward z FFT + multiply + inverse z FFT
    P3M_TRY(fused_z_pass<T>(c, s.spectrum, s.green, (long long)splane));
    phase_end(c, PH_MULTIPLY);
    phase_begin(c, PH_FFT_INV);
    for (int z = 0; z < p.nz; z += s.fft_chunk) {
      P3M_FFT(exec_inv(s.plan_inv, s.spectrum + splane * z, s.potential + plane * z));
      c->launches += 2;
    }
    phase_end(c, PH_FFT_INV);
    c->have_potential = true;
    return 0;
  }
  phase_begin(c, PH_FFT_FWD);
  P3M_FFT(exec_fwd(s.plan_fwd, s.density, s.spectrum));
  c->launches += 2;
  phase_end(c, PH_FFT_FWD);
  phase_begin(c, PH_MULTIPLY);
  k_multiply<<<(unsigned)((hc + 255) / 256), 256, 0, c->stream>>>(s.spectrum, s.green, hc);
  P3M_LAUNCH_CHECK(c);
  phase_end(c, PH_MULTIPLY);
  phase_begin(c, PH_FFT_INV);
  P3M_FFT(exec_inv(s.plan_inv, s.spectrum, s.potential));
  c->launches += 2;
  phase_end(c, PH_FFT_INV);
  c->have_potential = true;
  return 0;
}

template <typename T, typename O>
__global__ void k_convert(const T* __restrict__ in, O* __restrict__ out, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (O)in[i];
}

// Full-mesh readback.  Slab-decomposed mesh: `dev` is this rank's FFT slab; its planes are placed at their
// position in the full array and every other plane reads 0 (sum over ranks = the full mesh).
template <typename T, typename O>
int get_mesh(p3m_ctx* c, const T* dev, O* out, long long count) {
  if (!dev) return fail(P3M_ESTATE, "mesh not available");
  O* stage = nullptr;
  P3M_CUDA(cudaMallocAsync((void**)&stage, sizeof(O) * (size_t)count, c->stream));
  if (c->slab) {
    const long long share = count / c->nranks;
    P3M_CUDA(cudaMemsetAsync(stage, 0, sizeof(O) * (size_t)count, c->stream));
    if (c->slab_split) {  // run 0 in the lower half, run 1 in the upper half of the full array
      const long long half = share / 2;
      k_convert<T, O><<<(unsigned)((half + 255) / 256), 256, 0, c->stream>>>(dev, stage + half * c->rank, half);
      k_convert<T, O><<<(unsigned)((half + 255) / 256), 256, 0, c->stream>>>(dev + half, stage + count / 2 + half * c->rank, half);
    } else {
      k_convert<T, O><<<(unsigned)((share + 255) / 256), 256, 0, c->stream>>>(dev, stage + share * c->rank, share);
    }
  } else {
    k_convert<T, O><<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(dev, stage, count);
  }
  P3M_LAUNCH_CHECK(c);
  P3M_CUDA(cudaMemcpyAsync(out, stage, sizeof(O) * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaFreeAsync(stage, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

template <typename T, typename I>
int set_mesh(p3m_ctx* c, T* dev, const I* in, long long count) {
  I* stage = nullptr;
  P3M_CUDA(cudaMallocAsync((void**)&stage, sizeof(I) * (size_t)count, c->stream));
  P3M_CUDA(cudaMemcpyAsync(stage, in, sizeof(I) * (size_t)count, cudaMemcpyHostToDevice, c->stream));
  if (c->slab) {  // keep this rank's planes of the full array
    const long long share = count / c->nranks;
    if (c->slab_split) {
      const long long half = share / 2;
      k_convert<I, T><<<(unsigned)((half + 255) / 256), 256, 0, c->stream>>>(stage + half * c->rank, dev, half);
      k_convert<I, T><<<(unsigned)((half + 255) / 256), 256, 0, c->stream>>>(stage + count / 2 + half * c->rank, dev + half, half);
    } else {
      k_convert<I, T><<<(unsigned)((share + 255) / 256), 256, 0, c->stream>>>(stage + share * c->rank, dev, share);
    }
  } else {
    k_convert<I, T><<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(stage, dev, count);
  }
  P3M_LAUNCH_CHECK(c);
  P3M_CUDA(cudaFreeAsync(stage, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

__global__ void k_scale_c(cufftComplex* d, long long n, float s) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) d[i].x *= s, d[i].y *= s;
}

int fft3d_c2c(int nz, int ny, int nx, const float* in, float* out, int inverse) {
  if (nz < 1 || ny < 1 || nx < 1 || !in || !out) return fail(P3M_EINVAL, "p3m_fft3d_c2c: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(P3M_ENODEV, "no CUDA device available (this library has no CPU fallback)");
  }
  const long long n = (long long)nz * ny * nx;
  cufftComplex* d = nullptr;
  P3M_CUDA(cudaMalloc((void**)&d, sizeof(cufftComplex) * (size_t)n));
  cufftHandle plan;
  cufftResult r = cufftPlan3d(&plan, nz, ny, nx, CUFFT_C2C);
  if (r != CUFFT_SUCCESS) {
    cudaFree(d);
    return fail(P3M_ECUDA, "cufftPlan3d failed (%d)", (int)r);
  }
  int rc = 0;
  do {
    if (cudaMemcpy(d, in, sizeof(cufftComplex) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) { rc = fail(P3M_ECUDA, "H2D copy failed"); break; }
    if (cufftExecC2C(plan, d, d, inverse ? CUFFT_INVERSE : CUFFT_FORWARD) != CUFFT_SUCCESS) { rc = fail(P3M_ECUDA, "cufftExecC2C failed"); break; }
    if (inverse) k_scale_c<<<(unsigned)((n + 255) / 256), 256>>>(d, n, 1.0f / (float)n);
    if (cudaMemcpy(out, d, sizeof(cufftComplex) * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = fail(P3M_ECUDA, "D2H copy failed"); break; }
  } while (0);
  cufftDestroy(plan);
  cudaFree(d);
  return rc;
}

#define INST(T)                                                         \
  template int green_init<T>(p3m_ctx*);                                 \
  template int green_set<T, float>(p3m_ctx*, const float*);             \
  template int green_set<T, double>(p3m_ctx*, const double*);           \
  template int green_get<T>(p3m_ctx*, double*);                         \
  template int alloc_meshes<T>(p3m_ctx*);                               \
  template int poisson<T>(p3m_ctx*);                                    \
  template int complete_potential<T>(p3m_ctx*);                              \
  template int get_mesh<T, float>(p3m_ctx*, const T*, float*, long long);   \
  template int get_mesh<T, double>(p3m_ctx*, const T*, double*, long long); \
  template int set_mesh<T, float>(p3m_ctx*, T*, const float*, long long);   \
  template int set_mesh<T, double>(p3m_ctx*, T*, const double*, long long);
INST(float)
INST(double)

}  // namespace p3m
