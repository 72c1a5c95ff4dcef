// ics.cu -- N3: initial conditions on the device.
//
// The reference samples its initial conditions on the host, one particle at a time, from
// std::default_random_engine streams with a Newton solve per particle (source/plummerSampler.cpp:11-83,
// source/diskSamplerLinear.cpp:10-74, source/utils.cpp:51-68): minutes for 2^26 particles, and every rank of a
// multi-GPU run would have to hold the whole set.  Here particle i is a pure function of (seed, i) -- a
// counter-based generator (SplitMix64 finaliser over (seed, i, draw)) -- so a rank generates exactly the
// particles of its own z-slab, and the per-layer work weights that define the slabs are evaluated on the
// device as well.  The DISTRIBUTIONS are the reference's (inverse-CDF radii, q^2 (1 - q^2)^3.5 speeds, linear
// surface density, circular speeds from bulge + disk field); its random STREAMS are implementation-defined
//  and are not reproduced.  tests/test_ics.py checks the distributions against
// particlesimulation_b200/ics.py (numpy) and, where it is built, against the compiled reference's samplers.
#include <cmath>
#include <vector>

#include "ctx.cuh"

namespace p3m {

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// draw number `d` of particle i: uniform in (0, 1), 53 bits
__device__ inline double u01(uint64_t seed, uint64_t i, uint32_t d) {
  const uint64_t h = mix64(mix64(seed ^ (i * 0xD1342543DE82EF95ull)) + (uint64_t)d * 0x9E3779B97F4A7C15ull);
  return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ inline void isotropic(double u_phi, double u_cos, double& x, double& y, double& z) {
  const double phi = 2.0 * 3.14159265358979323846 * u_phi;
  const double ct = 1.0 - 2.0 * u_cos;  // == cos(acos(1 - 2u)), source/plummerSampler.cpp:53
  const double st = sqrt(fmax(0.0, 1.0 - ct * ct));
  x = st * cos(phi), y = st * sin(phi), z = ct;
}

__device__ inline double plummer_radius(double u, double a, double r_max, int truncate) {
  // M(r)/M = r^3 / (r^2 + a^2)^(3/2) = u  =>  r = a (u^(-2/3) - 1)^(-1/2)   source/plummerSampler.cpp:49
  if (truncate) {
    const double f = r_max * r_max * r_max / pow(r_max * r_max + a * a, 1.5);
    u *= f;
  }
  double r = a / sqrt(fmax(pow(u, -2.0 / 3.0) - 1.0, 1e-300));
  return r > r_max ? r_max : r;  // :50-52
}

// speed fraction q = v / v_esc with density q^2 (1 - q^2)^(7/2) (source/plummerSampler.cpp:39-41; the reference
// inverts the CDF with Newton, here by rejection: same law)
__device__ inline double plummer_q(uint64_t seed, uint64_t i, uint32_t first_draw) {
  for (uint32_t k = 0; k < 64; ++k) {
    const double x = u01(seed, i, first_draw + 2 * k), y = 0.1 * u01(seed, i, first_draw + 2 * k + 1);
    const double w = 1.0 - x * x;
    if (y < x * x * w * w * w * sqrt(w)) return x;
  }
  return 0.5;
}

__device__ inline void bulge_field(double px, double py, double pz, double R, double M, double G, double& gx,
                                   double& gy, double& gz) {
  // sphRadDecrField, source/externalFields.cpp:4-15 (centre at the origin here)
  const double r = sqrt(px * px + py * py + pz * pz);
  const double g = r > R ? -G * M / (r * r) : -(G * M / (R * R * R)) * r * (4 - 3 * r / R);
  const double s = r > 0 ? g / r : 0.0;
  gx = s * px, gy = s * py, gz = s * pz;
}

// particle i of the set, original units, relative quantities in double
__device__ inline void ic_particle(const p3m_ic& ic, uint64_t i, double pos[3], double vel[3]) {
  const double pi = 3.14159265358979323846;
  const double cx = ic.center[0], cy = ic.center[1], cz = ic.center[2];
  vel[0] = vel[1] = vel[2] = 0.0;
  if (ic.kind == P3M_IC_PLUMMER) {
    const double r = plummer_radius(u01(ic.seed, i, 1), ic.a, ic.r_max, ic.truncate);
    double dx, dy, dz;
    isotropic(u01(ic.seed, i, 0), u01(ic.seed, i, 2), dx, dy, dz);
    pos[0] = cx + r * dx, pos[1] = cy + r * dy, pos[2] = cz + r * dz;
    const double vesc = sqrt(2.0 * ic.G * ic.total_mass / sqrt(r * r + (double)ic.a * ic.a));  // :68-69
    const double v = plummer_q(ic.seed, i, 16) * vesc;
    isotropic(u01(ic.seed, i, 3), u01(ic.seed, i, 4), dx, dy, dz);
    vel[0] = v * dx, vel[1] = v * dy, vel[2] = v * dz;
  } else if (ic.kind == P3M_IC_DISK_LINEAR) {
    // surface density ~ (rd - r) on [r0, rd]: root of include/diskSamplerLinear.h:25-30 by bisection
    const double rd = ic.rd, r0 = ic.r0, cdf = u01(ic.seed, i, 1);
    const double k0 = cdf * (rd - r0) * (rd - r0) * (2 * r0 + rd) - 2 * r0 * r0 * r0 + 3 * rd * r0 * r0;
    double lo = r0, hi = rd;
    for (int it = 0; it < 60; ++it) {
      const double m = 0.5 * (lo + hi);
      const double f = 2 * m * m * m - 3 * rd * m * m + k0;  // decreasing on [0, rd]
      if (f > 0) lo = m; else hi = m;
    }
    const double r = 0.5 * (lo + hi), phi = 2 * pi * u01(ic.seed, i, 0);
    const double x = r * cos(phi), y = r * sin(phi), z = (ic.thickness / 2) * (2 * u01(ic.seed, i, 2) - 1);
    pos[0] = cx + x, pos[1] = cy + y, pos[2] = cz + z;
    // circular speed from the bulge + disk field, source/diskSamplerLinear.cpp:37-65
    const double rr = sqrt(x * x + y * y + z * z), rho = sqrt(x * x + y * y);
    const double ra = rr / rd, sigma0 = 3 * ic.md / (pi * rd * rd), kk = 2.5, h = 0.66, aa = -kk / (h * h);
    const double gd = -ic.G * sigma0 * (aa * (ra - h) * (ra - h) + kk);
    double gx, gy, gz;
    bulge_field(x, y, z, ic.rb, ic.mb, ic.G, gx, gy, gz);
    if (rho > 0) gx += gd * x / rho, gy += gd * y / rho;
    const double gval = sqrt(gx * gx + gy * gy + gz * gz);
    const double v = rr > 0 ? sqrt(gval * rho * rho / rr) : 0.0;
    if (rho > 0) vel[0] = -v * y / rho, vel[1] = v * x / rho;
  } else if (ic.kind == P3M_IC_UNIFORM) {
    for (int d = 0; d < 3; ++d) pos[d] = ic.lo[d] + ((double)ic.hi[d] - ic.lo[d]) * u01(ic.seed, i, d);
    if (ic.vel_sigma > 0) {  // Box-Muller
      const double r1 = sqrt(-2.0 * log(u01(ic.seed, i, 3))), t1 = 2 * pi * u01(ic.seed, i, 4);
      const double r2 = sqrt(-2.0 * log(u01(ic.seed, i, 5))), t2 = 2 * pi * u01(ic.seed, i, 6);
      vel[0] = ic.vel_sigma * r1 * cos(t1), vel[1] = ic.vel_sigma * r1 * sin(t1), vel[2] = ic.vel_sigma * r2 * cos(t2);
    }
  } else {  // P3M_IC_DISK_HALO: BASELINE configs[3], see particlesimulation_b200/ics.py clustered_disk_halo
    const uint64_t nd = (uint64_t)ic.n / 2;
    const double M = ic.total_mass, a = ic.a;
    if (i < nd) {
      // disk in the x-z plane (normal along y) so that z-slabs cut through it; triangular radial law
      const double r = ic.rd * (1.0 - sqrt(1.0 - u01(ic.seed, i, 1)));
      const double phi = 2 * pi * u01(ic.seed, i, 0), cs = cos(phi), sn = sin(phi);
      pos[0] = cx + r * cs, pos[2] = cz + r * sn, pos[1] = cy + ic.thickness * (u01(ic.seed, i, 2) - 0.5);
      const double q = 1.0 - r / ic.rd;
      const double menc = M * (0.5 * (1.0 - q * q) + 0.5 * r * r * r / pow(r * r + a * a, 1.5));
      const double v = sqrt(ic.G * menc / fmax(r, 1e-3));
      vel[0] = -v * sn, vel[2] = v * cs;
    } else {
      const double r = plummer_radius(u01(ic.seed, i, 1), a, ic.r_max, 1);
      double dx, dy, dz;
      isotropic(u01(ic.seed, i, 0), u01(ic.seed, i, 2), dx, dy, dz);
      pos[0] = cx + r * dx, pos[1] = cy + r * dy, pos[2] = cz + r * dz;
      const double vesc = sqrt(2.0 * ic.G * M / sqrt(r * r + a * a));
      isotropic(u01(ic.seed, i, 3), u01(ic.seed, i, 4), dx, dy, dz);
      vel[0] = 0.5 * vesc * dx, vel[1] = 0.5 * vesc * dy, vel[2] = 0.5 * vesc * dz;
    }
  }
}

// code-unit record of particle i, with exactly the fp32 operations of k_upload (binsort.cu) on the fp32 values
// a host sampler would have handed to p3m_set_particles
template <typename T>
__device__ inline void ic_record(const p3m_ic& ic, uint64_t i, T H, T DT, T mass_factor, V4<T>& posm, V4<T>& vel) {
  double p[3], v[3];
  ic_particle(ic, i, p, v);
  const float m = ic.total_mass / (float)ic.n;
  T x = (T)(float)p[0], y = (T)(float)p[1], z = (T)(float)p[2];
  T vx = (T)(float)v[0], vy = (T)(float)v[1], vz = (T)(float)v[2];
  posm = V4<T>{x / H, y / H, z / H, mass_factor * (T)m};
  vel = V4<T>{DT * vx / H, DT * vy / H, DT * vz / H, 0};
}

// host readback (p3m_sample_particles)
__global__ void k_ic_sample(p3m_ic ic, long long first, long long count, float* __restrict__ pos,
                            float* __restrict__ vel, float* __restrict__ mass) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= count) return;
  double p[3], v[3];
  ic_particle(ic, (uint64_t)(first + k), p, v);
  if (pos) pos[3 * k] = (float)p[0], pos[3 * k + 1] = (float)p[1], pos[3 * k + 2] = (float)p[2];
  if (vel) vel[3 * k] = (float)v[0], vel[3 * k + 1] = (float)v[1], vel[3 * k + 2] = (float)v[2];
  if (mass) mass[k] = ic.total_mass / (float)ic.n;
}

// single rank: the whole set straight into the particle arrays
template <typename T>
__global__ void k_ic_fill_all(p3m_ic ic, long long n, T H, T DT, T mf, V4<T>* __restrict__ posm,
                              V4<T>* __restrict__ vel, V4<T>* __restrict__ acc, V4<T>* __restrict__ acc_sr,
                              int* __restrict__ id) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  V4<T> p, v;
  ic_record<T>(ic, (uint64_t)i, H, DT, mf, p, v);
  posm[i] = p, vel[i] = v, acc[i] = V4<T>{0, 0, 0, 0}, acc_sr[i] = V4<T>{0, 0, 0, 0}, id[i] = (int)i;
}

// several ranks, pass 1: occupancy histogram of the binning cells (P3M) or of the layers (PM) of the WHOLE set
template <typename T>
__global__ void k_ic_hist(p3m_ic ic, long long n, Geom<T> g, T H, T DT, T mf, int layers, int by_cell,
                          unsigned* __restrict__ cnt) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  V4<T> p, v;
  ic_record<T>(ic, (uint64_t)i, H, DT, mf, p, v);
  int cx, cy, cz;
  bool inside;
  bin_cell(g, p.x, p.y, p.z, cx, cy, cz, inside);
  cz = min(cz, layers - 1);
  if (by_cell)
    atomicAdd(&cnt[((size_t)cz * g.my + cy) * g.mx + cx], 1u);
  else
    atomicAdd(&cnt[cz], 1u);
}

// pass 2 (P3M): weight of layer z = sum over its cells of n_cell * (w_particle + particles in the 27 cells), the
// same figure dist_balance_cuts computes on the host; one CTA per layer, fixed summation order
template <typename T>
__global__ void __launch_bounds__(256)
k_ic_layer_weight(const unsigned* __restrict__ cnt, int mx, int my, int layers, double w_particle,
                  double* __restrict__ weight) {
  const int z = blockIdx.x;
  __shared__ double red[256];
  double w = 0.0;
  for (int k = threadIdx.x; k < mx * my; k += 256) {
    const int x = k % mx, y = k / mx;
    const unsigned c0 = cnt[((size_t)z * my + y) * mx + x];
    if (!c0) continue;
    double s27 = 0.0;
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int X = x + dx, Y = y + dy, Z = z + dz;
          if (X < 0 || Y < 0 || Z < 0 || X >= mx || Y >= my || Z >= layers) continue;
          s27 += (double)cnt[((size_t)Z * my + Y) * mx + X];
        }
    w += (double)c0 * (w_particle + s27);
  }
  red[threadIdx.x] = w;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) weight[z] = red[0];
}

// pass 3: count (out == nullptr) or append the particles of this rank's layers
template <typename T>
__global__ void k_ic_fill_local(p3m_ic ic, long long n, Geom<T> g, T H, T DT, T mf, int* __restrict__ counter,
                                long long cap, V4<T>* __restrict__ posm, V4<T>* __restrict__ vel,
                                V4<T>* __restrict__ acc, V4<T>* __restrict__ acc_sr, int* __restrict__ id) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  bool mine = false;
  V4<T> p, v;
  if (i < n) {
    ic_record<T>(ic, (uint64_t)i, H, DT, mf, p, v);
    int cx, cy, cz;
    bool inside;
    bin_cell(g, p.x, p.y, p.z, cx, cy, cz, inside);
    mine = layer_owner(g, cz) == g.rank;
  }
  // warp-aggregated append
  const unsigned m = __ballot_sync(0xffffffffu, mine);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (!mine || !posm) return;
  const long long k = base + __popc(m & ((1u << lane) - 1u));
  if (k >= cap) return;
  posm[k] = p, vel[k] = v, acc[k] = V4<T>{0, 0, 0, 0}, acc_sr[k] = V4<T>{0, 0, 0, 0}, id[k] = (int)i;
}

int check_ic(const p3m_ic* ic) {
  if (!ic) return fail(P3M_EINVAL, "null initial-condition descriptor");
  if (ic->kind < P3M_IC_PLUMMER || ic->kind > P3M_IC_DISK_HALO) return fail(P3M_EINVAL, "unknown initial-condition kind %d", ic->kind);
  if (ic->n < 0 || ic->n > 0x7fffffffLL) return fail(P3M_EINVAL, "particle count %lld out of range", (long long)ic->n);
  if ((ic->kind == P3M_IC_PLUMMER || ic->kind == P3M_IC_DISK_HALO) && !(ic->a > 0 && ic->r_max > 0))
    return fail(P3M_EINVAL, "Plummer sampler needs a > 0 and r_max > 0");
  if ((ic->kind == P3M_IC_DISK_LINEAR || ic->kind == P3M_IC_DISK_HALO) && !(ic->rd > 0))
    return fail(P3M_EINVAL, "disk sampler needs rd > 0");
  return 0;
}

}  // namespace

template <typename T>
int generate_particles(p3m_ctx* c, const p3m_ic* icp) {
  P3M_TRY(check_ic(icp));
  const p3m_ic ic = *icp;
  const long long n = ic.n;
  Geom<T>& g = Sel<T>::g(c);
  const T mf = c->f64 ? (T)c->mass_factor64 : (T)c->mass_factor32;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  long long n_local = n, n_max = 0;
  if (c->nranks > 1 && n > 0) {
    // (1) work weights of the binning layers from the whole set, on this rank's own device
    dist_set_cuts(c);  // geometric cuts: defines how many layers are being cut
    const int layers = g.cut[c->nranks];
    if (!c->tune.static_cuts) {
      const bool cells = g.p3m && !c->tune.count_cuts;
      const size_t ncnt = cells ? (size_t)g.mx * g.my * layers : (size_t)layers;
      unsigned* cnt = nullptr;
      double* wdev = nullptr;
      P3M_CUDA(cudaMallocAsync((void**)&cnt, sizeof(unsigned) * ncnt, c->stream));
      P3M_CUDA(cudaMallocAsync((void**)&wdev, sizeof(double) * layers, c->stream));
      P3M_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned) * ncnt, c->stream));
      k_ic_hist<T><<<blocks, 256, 0, c->stream>>>(ic, n, g, g.H, g.DT, mf, layers, cells ? 1 : 0, cnt);
      P3M_LAUNCH_CHECK(c);
      std::vector<double> weight((size_t)layers, 0.0);
      if (cells) {
        k_ic_layer_weight<T><<<layers, 256, 0, c->stream>>>(cnt, g.mx, g.my, layers, c->tune.particle_weight, wdev);
        P3M_LAUNCH_CHECK(c);
        P3M_CUDA(cudaMemcpyAsync(weight.data(), wdev, sizeof(double) * layers, cudaMemcpyDeviceToHost, c->stream));
        P3M_CUDA(cudaStreamSynchronize(c->stream));
      } else {
        std::vector<unsigned> h((size_t)layers);
        P3M_CUDA(cudaMemcpyAsync(h.data(), cnt, sizeof(unsigned) * layers, cudaMemcpyDeviceToHost, c->stream));
        P3M_CUDA(cudaStreamSynchronize(c->stream));
        for (int z = 0; z < layers; ++z) weight[(size_t)z] = (double)h[(size_t)z];
      }
      P3M_CUDA(cudaFreeAsync(cnt, c->stream));
      P3M_CUDA(cudaFreeAsync(wdev, c->stream));
      P3M_TRY(dist_cuts_from_weights<T>(c, weight.data(), layers));
    }
    // (2) how many particles fall into every rank's layers (final cuts): this rank's own count, and the largest
    // one, which sizes the arrays of EVERY rank -- equal capacities keep the migration / ghost-layer capacity
    // checks symmetric, and a thin slab next to a cluster core receives that core's boundary layer as ghosts
    unsigned* lcnt = nullptr;
    P3M_CUDA(cudaMallocAsync((void**)&lcnt, sizeof(unsigned) * (size_t)layers, c->stream));
    P3M_CUDA(cudaMemsetAsync(lcnt, 0, sizeof(unsigned) * (size_t)layers, c->stream));
    k_ic_hist<T><<<blocks, 256, 0, c->stream>>>(ic, n, g, g.H, g.DT, mf, layers, 0, lcnt);
    P3M_LAUNCH_CHECK(c);
    std::vector<unsigned> hl((size_t)layers);
    P3M_CUDA(cudaMemcpyAsync(hl.data(), lcnt, sizeof(unsigned) * (size_t)layers, cudaMemcpyDeviceToHost, c->stream));
    P3M_CUDA(cudaStreamSynchronize(c->stream));
    P3M_CUDA(cudaFreeAsync(lcnt, c->stream));
    long long per_rank[P3M_MAX_RANKS] = {0};
    for (int z = 0; z < layers; ++z) per_rank[layer_owner(g, z)] += hl[(size_t)z];
    n_local = per_rank[c->rank];
    for (int r = 0; r < c->nranks; ++r) n_max = per_rank[r] > n_max ? per_rank[r] : n_max;
  }
  // capacity: head-room for migration (a rank may hold up to ~1.5 x the largest initial share)
  const long long want = c->nranks > 1 ? n_max + n_max / 2 + 4096 : n_local;
  P3M_TRY(alloc_particles<T>(c, want));
  State<T>& s = Sel<T>::st(c);
  if (n > 0) {
    if (c->nranks > 1) {
      int* counter = s.pp_counters + 5;
      P3M_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), c->stream));
      k_ic_fill_local<T><<<blocks, 256, 0, c->stream>>>(ic, n, g, g.H, g.DT, mf, counter, c->cap, s.posm, s.vel, s.acc,
                                                       s.acc_sr, s.id);
    } else {
      k_ic_fill_all<T><<<blocks, 256, 0, c->stream>>>(ic, n, g.H, g.DT, mf, s.posm, s.vel, s.acc, s.acc_sr, s.id);
    }
    P3M_LAUNCH_CHECK(c);
  }
  c->n = n_local;
  c->n_global = n;
  c->have_particles = true;
  c->sorted = false;
  c->order_valid = false;
  c->have_acc = true;
  // every particle carries total_mass / n: the equal-mass force table applies
  const float m = n > 0 ? ic.total_mass / (float)n : 0.f;
  c->mass_lo = c->mass_hi = m;
  c->uniform_mass = n > 0 && m > 0.f;
  c->uniform_mass_code = (double)(mf * (T)m);
  P3M_CUDA(cudaMemsetAsync(s.flags, 0, sizeof(int) * 4, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int sample_particles(const p3m_ic* icp, long long first, long long count, float* pos, float* vel, float* mass) {
  P3M_TRY(check_ic(icp));
  if (first < 0 || count < 0 || first + count > icp->n) return fail(P3M_EINVAL, "p3m_sample_particles: range outside the set");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(P3M_ENODEV, "no CUDA device available (this library has no CPU fallback)");
  }
  if (count == 0) return 0;
  float* d = nullptr;
  P3M_CUDA(cudaMalloc((void**)&d, sizeof(float) * 7 * (size_t)count));
  k_ic_sample<<<(unsigned)((count + 255) / 256), 256>>>(*icp, first, count, pos ? d : nullptr, vel ? d + 3 * count : nullptr,
                                                       mass ? d + 6 * count : nullptr);
  int rc = 0;
  if (cudaGetLastError() != cudaSuccess) rc = fail(P3M_ECUDA, "sampler launch failed");
  if (!rc && pos && cudaMemcpy(pos, d, sizeof(float) * 3 * count, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(P3M_ECUDA, "D2H copy failed");
  if (!rc && vel && cudaMemcpy(vel, d + 3 * count, sizeof(float) * 3 * count, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(P3M_ECUDA, "D2H copy failed");
  if (!rc && mass && cudaMemcpy(mass, d + 6 * count, sizeof(float) * count, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(P3M_ECUDA, "D2H copy failed");
  cudaFree(d);
  return rc;
}

template int generate_particles<float>(p3m_ctx*, const p3m_ic*);
template int generate_particles<double>(p3m_ctx*, const p3m_ic*);

}  // namespace p3m
