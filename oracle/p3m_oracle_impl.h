/* oracle/p3m_oracle_impl.h -- TEST INFRASTRUCTURE ONLY.  Body of the C restatement, included twice
 * by p3m_oracle.c with (R, S, SIN, COS, POW, SQRT, ROUND) bound to float / double.  See
 * p3m_oracle.h for the contract.  Reference citations are relative to /root/reference/.
 *
 * The float instantiation keeps the reference's operation ORDER (left-to-right evaluation, the same
 * libm calls, no FMA contraction: built with -ffp-contract=off) so that it tracks the compiled
 * reference to within a few ulp; where the reference uses std::complex arithmetic whose real or
 * imaginary lane is identically zero, only the non-zero lane is restated (derivation in comments).
 */

#define FN2(a, b) a##_##b
#define FN1(a, b) FN2(a, b)
#define FN(name) FN1(name, S)

typedef struct {
  R x, y, z;
} FN(V3);

static inline size_t FN(flat)(const OrcParams* p, int x, int y, int z) {
  /* include/grid.h:52-54  getIndx: x fastest */
  return (size_t)((long)x + (long)y * p->nx + (long)z * p->nx * p->ny);
}

static inline int FN(wrap)(int a, int b) { /* include/grid.h:46 mod() */ return (a % b + b) % b; }

/* ------------------------------------------------------------------------------------------- */
/* unit conversions: include/unitConversions.h:8-50                                            */

void FN(orc_to_code_units)(const OrcParams* p, const float* pos, const float* vel,
                           const float* mass, R* pos_c, R* vel_c, R* mass_c) {
  R H = p->H, DT = p->DT, G = p->G;
  for (int i = 0; i < 3 * p->n; ++i) {
    if (pos_c) pos_c[i] = (R)pos[i] / H;               /* :8-10 pos / H */
    if (vel_c) vel_c[i] = vel ? DT * (R)vel[i] / H : 0; /* :15-17 DT * v / H */
  }
  if (mass_c)
    for (int i = 0; i < p->n; ++i) /* :42-44 */
      mass_c[i] = DT * DT * 4 * (R)PI_R * G / (H * H * H) * (R)mass[i];
}

/* ------------------------------------------------------------------------------------------- */
/* Green's functions: source/greensFunctions.cpp                                               */

static R FN(sinc)(R x) { /* :8-13 */ return x == 0 ? 1 : SIN(x) / x; }

static R FN(assign_fourier)(int is, const R k[3]) {
  /* :15-40 TSCFourier / CICFourier / NGPFourier */
  R prod = 1;
  for (int i = 0; i < 3; ++i) {
    R s = FN(sinc)(k[i] / 2);
    prod *= (is == 2) ? POW(s, 3) : (is == 1) ? POW(s, 2) : s;
  }
  return prod;
}

static R FN(cloud_fourier)(int cs, R k, R a) {
  R u = k * a / 2;
  if (cs == 0) /* :42-45 S1Fourier */
    return -3 / POW(u, 3) * (u * COS(u) - SIN(u));
  /* :47-50 S2Fourier */
  return 12 / POW(u, 4) * (2 - 2 * COS(u) - u * SIN(u));
}

static R FN(alias_sum)(int is, const R k[3]) {
  if (is == 2) { /* :102-108 TSCAliasSum */
    R sum = 1;
    for (int i = 0; i < 3; ++i)
      sum *= (1 - POW(SIN(k[i] / 2), 2) + (R)2.0 / 15 * POW(SIN(k[i] / 2), 4));
    return sum;
  }
  if (is == 1) { /* :110-116 CICAliasSum */
    R sum = 1;
    for (int i = 0; i < 3; ++i) sum *= (1 + 2 * POW(COS(k[i] / 2), 2));
    return ((R)1.0 / 27) * sum;
  }
  return 1; /* :118-120 */
}

/* :122-189 GreenOptimal.  D = i*d (pure imaginary, :70-89), R_n = -i*k_n*s^2/|k_n|^2 (pure
 * imaginary, :52-68), so with conj() in dotProduct (:91-100):  D . conj(sum u^2 R) = sum_i d_i *
 * r_i (real), |D|^2 = sum d_i^2, and the returned complex has zero imaginary part. */
static R FN(green_optimal)(const OrcParams* p, int kx, int ky, int kz, R a) {
  if (kx == 0 && ky == 0 && kz == 0) return 0;
  if (p->greenZeroDegenerate && (2 * kx) % p->nx == 0 && (2 * ky) % p->ny == 0 &&
      (2 * kz) % p->nz == 0)
    return 0; /* not in the reference: see OrcParams.greenZeroDegenerate */
  const R pi = (R)PI_R;
  const int is = p->is, cs = (p->gfunc == 2) ? 1 : 0;
  R k[3] = {2 * pi * (R)kx / p->nx, 2 * pi * (R)ky / p->ny, 2 * pi * (R)kz / p->nz};
  R denomSum = FN(alias_sum)(is, k);
  R d[3];
  if (p->fds == 0) { /* D2Fourier :70-78 */
    for (int i = 0; i < 3; ++i) d[i] = SIN(k[i]);
  } else { /* D4Fourier :80-89 */
    R alpha = (R)4.0 / 3;
    for (int i = 0; i < 3; ++i) d[i] = alpha * SIN(k[i]) + (1 - alpha) * SIN(2 * k[i]) / (R)2.0;
  }
  R dnorm = 0;
  for (int i = 0; i < 3; ++i) dnorm += d[i] * d[i];
  R num[3] = {0, 0, 0};
  for (int n1 = -2; n1 <= 2; ++n1)
    for (int n2 = -2; n2 <= 2; ++n2)
      for (int n3 = -2; n3 <= 2; ++n3) {
        R kn[3] = {k[0] + 2 * pi * n1, k[1] + 2 * pi * n2, k[2] + 2 * pi * n3};
        R u = FN(assign_fourier)(is, kn);
        R u2 = POW(u, 2);
        R kl = SQRT(kn[0] * kn[0] + kn[1] * kn[1] + kn[2] * kn[2]);
        R s = FN(cloud_fourier)(cs, kl, a);
        R s2 = POW(s, 2);
        for (int i = 0; i < 3; ++i) num[i] += u2 * (-kn[i] * s2 / (kl * kl));
      }
  R numerator = 0;
  for (int i = 0; i < 3; ++i) numerator += d[i] * num[i];
  R denominator = dnorm * denomSum * denomSum;
  return numerator / denominator;
}

static R FN(green_laplacian)(const OrcParams* p, int kx, int ky, int kz) {
  /* :191-200 */
  if (kx == 0 && ky == 0 && kz == 0) return 0;
  const R pi = (R)PI_R;
  R sx = SIN(pi * kx / p->nx), sy = SIN(pi * ky / p->ny), sz = SIN(pi * kz / p->nz);
  return (R)-0.25 / (sx * sx + sy * sy + sz * sz);
}

static R FN(green_poor_man)(const OrcParams* p, int i, int j, int k) {
  /* :202-220 */
  if (i == 0 && j == 0 && k == 0) return 0;
  const R pi = (R)PI_R;
  int ki = (i <= p->nx / 2) ? i : i - p->nx;
  int kj = (j <= p->ny / 2) ? j : j - p->ny;
  int kk = (k <= p->nz / 2) ? k : k - p->nz;
  R kx = 2 * pi * ki / p->nx, ky = 2 * pi * kj / p->ny, kz = 2 * pi * kk / p->nz;
  R k2 = kx * kx + ky * ky + kz * kz;
  return (R)-1.0 / k2;
}

void FN(orc_green)(const OrcParams* p, R* green) {
  /* source/pmMethod.cpp:164-185; particleDiameter in code units :56 */
  const R a = (R)p->particleDiameter / (R)p->H;
  const long M = (long)p->nx * p->ny * p->nz;
#pragma omp parallel for schedule(dynamic, 256)
  for (long idx = 0; idx < M; ++idx) {
    int kx = (int)(idx % p->nx), ky = (int)((idx / p->nx) % p->ny),
        kz = (int)(idx / ((long)p->nx * p->ny));
    R g;
    if (p->gfunc == 0)
      g = FN(green_laplacian)(p, kx, ky, kz);
    else if (p->gfunc == 3)
      g = FN(green_poor_man)(p, kx, ky, kz);
    else
      g = FN(green_optimal)(p, kx, ky, kz, a);
    green[idx] = g;
  }
}

/* ------------------------------------------------------------------------------------------- */
/* mass assignment: source/pmMethod.cpp:187-277                                                */

static inline R FN(tsc_w)(R x, int t) {
  /* :187-198 TSCAssignmentFunc (each factor is 2x the true TSC weight; hence the /8) */
  if (t == 1) return ((R)0.5 + x) * ((R)0.5 + x);
  if (t == 0) return (R)1.5 - 2 * x * x;
  return ((R)0.5 - x) * ((R)0.5 - x);
}

/* Out-of-array indices are undefined behaviour in the reference (vector::operator[] past the end,
 * only reachable for particles outside [1, N-2] cells); the oracle skips them instead of crashing. */
#define DEP(idx, v)                                        \
  do {                                                     \
    size_t i__ = (idx);                                    \
    if (i__ < M) density[i__] += (v);                      \
  } while (0)

void FN(orc_deposit)(const OrcParams* p, const R* pos, const R* mass, R* density) {
  const size_t M = (size_t)p->nx * p->ny * p->nz;
  memset(density, 0, M * sizeof(R)); /* grid.cpp:34-36 clearDensity */
  for (int i = 0; i < p->n; ++i) {
    R px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2], d = mass[i];
    if (p->is == 0) { /* NGP :204-213 */
      int x = (int)ROUND(px), y = (int)ROUND(py), z = (int)ROUND(pz);
      DEP(FN(flat)(p, x, y, z), d);
    } else if (p->is == 1) { /* CIC :216-244, no periodic wrap (grid.cpp:28-32) */
      int x = (int)px, y = (int)py, z = (int)pz;
      R dx = px - x, dy = py - y, dz = pz - z;
      R tx = 1 - dx, ty = 1 - dy, tz = 1 - dz;
      DEP(FN(flat)(p, x, y, z), d * tx * ty * tz);
      DEP(FN(flat)(p, x + 1, y, z), d * dx * ty * tz);
      DEP(FN(flat)(p, x, y + 1, z), d * tx * dy * tz);
      DEP(FN(flat)(p, x, y, z + 1), d * tx * ty * dz);
      DEP(FN(flat)(p, x + 1, y + 1, z), d * dx * dy * tz);
      DEP(FN(flat)(p, x + 1, y, z + 1), d * dx * ty * dz);
      DEP(FN(flat)(p, x, y + 1, z + 1), d * tx * dy * dz);
      DEP(FN(flat)(p, x + 1, y + 1, z + 1), d * dx * dy * dz);
    } else { /* TSC :247-272, truncation base */
      int x = (int)px, y = (int)py, z = (int)pz;
      R dx = px - x, dy = py - y, dz = pz - z;
      for (int t1 = -1; t1 <= 1; ++t1) {
        R T1 = (d / 8) * FN(tsc_w)(dx, t1);
        for (int t2 = -1; t2 <= 1; ++t2) {
          R T2 = T1 * FN(tsc_w)(dy, t2);
          for (int t3 = -1; t3 <= 1; ++t3) {
            R T3 = T2 * FN(tsc_w)(dz, t3);
            DEP(FN(flat)(p, x + t1, y + t2, z + t3), T3);
          }
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------- */
/* FFT.  The reference delegates to an FFTAdapter (include/FFTAdapter.h:6-15); the contract pinned
 * by test/fftAdaptersTest.cpp:6-22 is: forward unnormalised with exp(-i k x), inverse divided by
 * the length.  Which FFT library computes it is not part of the algorithm (FFTW / kissfft /
 * pocketfft in the reference), so this is a plain mixed-radix Cooley-Tukey in R precision with
 * twiddles rounded from double.                                                                 */

typedef struct {
  R re, im;
} FN(cpx);

static void FN(fft_rec)(int n, int stride, const FN(cpx) * in, FN(cpx) * out, const FN(cpx) * tw,
                        int twstride, FN(cpx) * scratch) {
  /* decimation in time on the smallest prime factor; tw[j*twstride] = exp(-+2 pi i j / n) */
  if (n == 1) {
    out[0] = in[0];
    return;
  }
  int pfac = 2;
  while (n % pfac) ++pfac;
  int m = n / pfac;
  for (int q = 0; q < pfac; ++q)
    FN(fft_rec)(m, stride * pfac, in + (size_t)q * stride, out + (size_t)q * m, tw,
                twstride * pfac, scratch);
  if (pfac == 2) {
    for (int k = 0; k < m; ++k) {
      FN(cpx) w = tw[(size_t)k * twstride], a = out[k], b = out[m + k];
      R tr = b.re * w.re - b.im * w.im, ti = b.re * w.im + b.im * w.re;
      out[k].re = a.re + tr, out[k].im = a.im + ti;
      out[m + k].re = a.re - tr, out[m + k].im = a.im - ti;
    }
    return;
  }
  for (int k = 0; k < m; ++k) {
    for (int q = 0; q < pfac; ++q) scratch[q] = out[(size_t)q * m + k];
    for (int r = 0; r < pfac; ++r) {
      R sr = 0, si = 0;
      long kk = k + (long)r * m;
      for (int q = 0; q < pfac; ++q) {
        FN(cpx) w = tw[(size_t)((kk * q) % n) * twstride];
        sr += scratch[q].re * w.re - scratch[q].im * w.im;
        si += scratch[q].re * w.im + scratch[q].im * w.re;
      }
      out[(size_t)r * m + k].re = sr, out[(size_t)r * m + k].im = si;
    }
  }
}

static void FN(fft_axis)(FN(cpx) * data, int n, long count_outer, long stride_line, long stride_el,
                         long n_inner, int inverse) {
  /* transforms every line of length n with element stride stride_el */
  FN(cpx)* tw = (FN(cpx)*)malloc(sizeof(FN(cpx)) * (size_t)n);
  for (int j = 0; j < n; ++j) {
    double ang = (inverse ? 2.0 : -2.0) * 3.14159265358979323846 * j / n;
    tw[j].re = (R)cos(ang), tw[j].im = (R)sin(ang);
  }
#pragma omp parallel
  {
    FN(cpx)* in = (FN(cpx)*)malloc(sizeof(FN(cpx)) * (size_t)n * 2 + sizeof(FN(cpx)) * 64);
    FN(cpx)* out = in + n;
    FN(cpx)* scratch = out + n;
#pragma omp for collapse(2)
    for (long o = 0; o < count_outer; ++o)
      for (long i = 0; i < n_inner; ++i) {
        FN(cpx)* base = data + o * stride_line + i;
        for (int j = 0; j < n; ++j) in[j] = base[(size_t)j * stride_el];
        FN(fft_rec)(n, 1, in, out, tw, 1, scratch);
        for (int j = 0; j < n; ++j) base[(size_t)j * stride_el] = out[j];
      }
    free(in);
  }
  free(tw);
}

static void FN(fft3d)(const OrcParams* p, FN(cpx) * data, int inverse) {
  long nx = p->nx, ny = p->ny, nz = p->nz;
  FN(fft_axis)(data, (int)nx, ny * nz, nx, 1, 1, inverse);       /* x lines */
  FN(fft_axis)(data, (int)ny, nz, nx * ny, nx, nx, inverse);     /* y lines */
  FN(fft_axis)(data, (int)nz, 1, 0, nx * ny, nx * ny, inverse);  /* z lines */
}

void FN(orc_poisson)(const OrcParams* p, const R* density, const R* green, R* potential) {
  const size_t M = (size_t)p->nx * p->ny * p->nz;
  FN(cpx)* buf = (FN(cpx)*)malloc(sizeof(FN(cpx)) * M);
  for (size_t i = 0; i < M; ++i) buf[i].re = density[i], buf[i].im = 0;
  FN(fft3d)(p, buf, 0); /* grid.cpp:50-52 fftDensity */
  for (size_t i = 0; i < M; ++i) {
    /* pmMethod.cpp:340-350: potentialFourier = densityFourier * G, G = (g, 0) */
    buf[i].re = buf[i].re * green[i], buf[i].im = buf[i].im * green[i];
  }
  buf[0].re = 0, buf[0].im = 0;
  FN(fft3d)(p, buf, 1); /* grid.cpp:54-56 + adapter's /length (kissFFTAdapter.h:22-28) */
  const R len = (R)M;
  for (size_t i = 0; i < M; ++i) potential[i] = buf[i].re / len; /* grid.cpp:66-68 .real() */
  free(buf);
}

/* ------------------------------------------------------------------------------------------- */
/* field on the mesh: source/pmMethod.cpp:352-382; only getPotential wraps (grid.cpp:66-68)      */

static inline R FN(pot)(const OrcParams* p, const R* phi, int x, int y, int z) {
  return phi[(size_t)FN(wrap)(x, p->nx) + (size_t)FN(wrap)(y, p->ny) * p->nx +
             (size_t)FN(wrap)(z, p->nz) * p->nx * p->ny];
}

void FN(orc_field)(const OrcParams* p, const R* phi, R* field) {
#pragma omp parallel for
  for (int z = 0; z < p->nz; ++z)
    for (int y = 0; y < p->ny; ++y)
      for (int x = 0; x < p->nx; ++x) {
        R fx, fy, fz;
        if (p->fds == 0) { /* :355-358 */
          fx = (R)-0.5 * (FN(pot)(p, phi, x + 1, y, z) - FN(pot)(p, phi, x - 1, y, z));
          fy = (R)-0.5 * (FN(pot)(p, phi, x, y + 1, z) - FN(pot)(p, phi, x, y - 1, z));
          fz = (R)-0.5 * (FN(pot)(p, phi, x, y, z + 1) - FN(pot)(p, phi, x, y, z - 1));
        } else { /* :359-365 */
          const R c = (R)-1.0 / 12;
          fx = c * (-FN(pot)(p, phi, x + 2, y, z) + 8 * FN(pot)(p, phi, x + 1, y, z) -
                    8 * FN(pot)(p, phi, x - 1, y, z) + FN(pot)(p, phi, x - 2, y, z));
          fy = c * (-FN(pot)(p, phi, x, y + 2, z) + 8 * FN(pot)(p, phi, x, y + 1, z) -
                    8 * FN(pot)(p, phi, x, y - 1, z) + FN(pot)(p, phi, x, y - 2, z));
          fz = c * (-FN(pot)(p, phi, x, y, z + 2) + 8 * FN(pot)(p, phi, x, y, z + 1) -
                    8 * FN(pot)(p, phi, x, y, z - 1) + FN(pot)(p, phi, x, y, z - 2));
        }
        size_t i = FN(flat)(p, x, y, z);
        field[3 * i] = fx, field[3 * i + 1] = fy, field[3 * i + 2] = fz;
      }
}

/* ------------------------------------------------------------------------------------------- */
/* gather: source/pmMethod.cpp:279-338, 384-390; external field source/externalFields.cpp:4-15  */

static FN(V3) FN(ext_field)(const OrcParams* p, FN(V3) pos) {
  FN(V3) out = {0, 0, 0};
  if (p->extKind != 1) return out;
  R cx = p->extCenter[0], cy = p->extCenter[1], cz = p->extCenter[2];
  R Rb = p->extR, Mb = p->extM, G = p->G;
  R dx = pos.x - cx, dy = pos.y - cy, dz = pos.z - cz;
  R r = SQRT(dx * dx + dy * dy + dz * dz);
  R g;
  if (r > Rb)
    g = -G * Mb / (r * r);
  else
    g = -(G * Mb / POW(Rb, 3)) * r * (4 - 3 * r / Rb);
  out.x = g * (dx / r), out.y = g * (dy / r), out.z = g * (dz / r);
  return out;
}

static R FN(ext_potential)(const OrcParams* p, FN(V3) pos) {
  /* externalFields.cpp:17-24 sphRadDecrFieldPotential */
  if (p->extKind != 1) return 0;
  R dx = pos.x - p->extCenter[0], dy = pos.y - p->extCenter[1], dz = pos.z - p->extCenter[2];
  R Rb = p->extR, Mb = p->extM, G = p->G;
  R r = SQRT(dx * dx + dy * dy + dz * dz);
  if (r > Rb) return -G * Mb / r;
  R u = r / Rb;
  return G * Mb / Rb * (-2 + u * u * (2 - u));
}

static inline FN(V3) FN(fld)(const OrcParams* p, const R* f, int x, int y, int z) {
  size_t i = FN(flat)(p, x, y, z); /* grid.cpp:46-48 getField: NOT wrapped */
  FN(V3) v = {0, 0, 0};
  if (i < (size_t)p->nx * p->ny * p->nz) v.x = f[3 * i], v.y = f[3 * i + 1], v.z = f[3 * i + 2];
  return v;
}

#define ACCW(acc, w, v) ((acc).x += (w) * (v).x, (acc).y += (w) * (v).y, (acc).z += (w) * (v).z)

static FN(V3) FN(interpolate)(const OrcParams* p, const R* field, R x, R y, R z) {
  FN(V3) a = {0, 0, 0};
  if (p->is == 0) { /* :284-289 */
    return FN(fld)(p, field, (int)ROUND(x), (int)ROUND(y), (int)ROUND(z));
  } else if (p->is == 1) { /* :291-310 */
    int xi = (int)x, yi = (int)y, zi = (int)z;
    R dx = x - xi, dy = y - yi, dz = z - zi, tx = 1 - dx, ty = 1 - dy, tz = 1 - dz;
    FN(V3) v;
    /* a = first term, then + in source order (operator+ chains left to right) */
    v = FN(fld)(p, field, xi, yi, zi);
    a.x = tx * ty * tz * v.x, a.y = tx * ty * tz * v.y, a.z = tx * ty * tz * v.z;
    v = FN(fld)(p, field, xi + 1, yi, zi);
    ACCW(a, dx * ty * tz, v);
    v = FN(fld)(p, field, xi, yi + 1, zi);
    ACCW(a, tx * dy * tz, v);
    v = FN(fld)(p, field, xi, yi, zi + 1);
    ACCW(a, tx * ty * dz, v);
    v = FN(fld)(p, field, xi + 1, yi + 1, zi);
    ACCW(a, dx * dy * tz, v);
    v = FN(fld)(p, field, xi + 1, yi, zi + 1);
    ACCW(a, dx * ty * dz, v);
    v = FN(fld)(p, field, xi, yi + 1, zi + 1);
    ACCW(a, tx * dy * dz, v);
    v = FN(fld)(p, field, xi + 1, yi + 1, zi + 1);
    ACCW(a, dx * dy * dz, v);
    return a;
  }
  /* TSC :312-333 */
  int xi = (int)x, yi = (int)y, zi = (int)z;
  R dx = x - xi, dy = y - yi, dz = z - zi;
  for (int t1 = -1; t1 <= 1; ++t1) {
    R T1 = FN(tsc_w)(dx, t1);
    for (int t2 = -1; t2 <= 1; ++t2) {
      R T2 = T1 * FN(tsc_w)(dy, t2);
      for (int t3 = -1; t3 <= 1; ++t3) {
        R T3 = T2 * FN(tsc_w)(dz, t3);
        FN(V3) v = FN(fld)(p, field, xi + t1, yi + t2, zi + t3);
        ACCW(a, (R)0.125 * T3, v);
      }
    }
  }
  return a;
}

void FN(orc_gather)(const OrcParams* p, const R* pos, const R* field, R* acc) {
  const R H = p->H, DT = p->DT;
#pragma omp parallel for
  for (int i = 0; i < p->n; ++i) {
    R x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
    FN(V3) a = FN(interpolate)(p, field, x, y, z);
    /* :386-388  + accelerationToCodeUnits(externalField(positionToOriginalUnits(pos,H)),H,DT) */
    FN(V3) po = {H * x, H * y, H * z};
    FN(V3) e = FN(ext_field)(p, po);
    if (p->extKind == 1) {
      a.x += DT * DT * e.x / H, a.y += DT * DT * e.y / H, a.z += DT * DT * e.z / H;
    } else {
      a.x += 0, a.y += 0, a.z += 0;
    }
    acc[3 * i] = a.x, acc[3 * i + 1] = a.y, acc[3 * i + 2] = a.z;
  }
}

/* ------------------------------------------------------------------------------------------- */
/* short range: source/p3mMethod.cpp:194-294, source/chainingMesh.cpp                           */

typedef struct {
  R re, a, eps, delta2; /* code units: p3mMethod.cpp:35-44 */
  int M[3];
  R HC[3];
} FN(SR);

static FN(SR) FN(sr_setup)(const OrcParams* p) {
  FN(SR) s;
  s.re = (R)p->cutoffRadius / (R)p->H;   /* :35 */
  s.a = (R)p->particleDiameter / (R)p->H; /* :36 */
  s.eps = (R)p->softening / (R)p->H;      /* :37 */
  s.delta2 = s.re * s.re / (500 - 1);     /* :42-44 */
  for (int d = 0; d < 3; ++d) {
    /* chainingMesh.cpp:10-15 (box and cutoff are fp32 inputs; the division is done in R) */
    s.M[d] = (int)((R)p->box[d] / (R)p->cutoffRadius);
    s.HC[d] = ((R)p->box[d] / s.M[d]) / (R)p->H;
  }
  return s;
}

static R FN(ref_force)(const FN(SR) * s, int cs, R r) {
  const R a = s->a;
  const R G = 1 / (4 * (R)PI_R);
  if (cs == 0) { /* referenceForceS1 :194-201 */
    if (r >= a) return G / (r * r);
    return G / (a * a) * (8 * r / a - 9 * r * r / (a * a) + 2 * POW(r / a, 4));
  }
  /* referenceForceS2 :203-218 */
  const R u = 2 * r / a;
  if (u <= 1)
    return G / (35 * POW(a, 2)) *
           (224 * u - 224 * POW(u, 3) + 70 * POW(u, 4) + 48 * POW(u, 5) - 21 * POW(u, 6));
  if (u <= 2)
    return G / (35 * POW(a, 2)) *
           (12 / POW(u, 2) - 224 + 896 * u - 840 * POW(u, 2) + 224 * POW(u, 3) + 70 * POW(u, 4) -
            48 * POW(u, 5) + 7 * POW(u, 6));
  return G / (r * r);
}

static void FN(sr_table)(const OrcParams* p, const FN(SR) * s, R* table) {
  /* initSRForceTable :275-294 */
  for (int i = 0; i < 500; ++i) {
    R r2 = i * s->delta2;
    R r = SQRT(r2);
    R Rf = -FN(ref_force)(s, p->cloudShape, r);
    R G = 1 / (4 * (R)PI_R);
    R total = -G / (r * r + s->eps * s->eps);
    table[i] = (r == 0) ? 0 : (total - Rf) / SQRT(r * r + s->eps * s->eps);
  }
}

void FN(orc_sr_table)(const OrcParams* p, R* table500) {
  FN(SR) s = FN(sr_setup)(p);
  FN(sr_table)(p, &s, table500);
}

static inline int FN(cell_of)(const FN(SR) * s, const R* pos, int i) {
  /* chainingMesh.cpp:25-29 (division, not reciprocal multiply: SURVEY Q3) + :79-84 */
  int cx = (int)(pos[3 * i] / s->HC[0]);
  int cy = (int)(pos[3 * i + 1] / s->HC[1]);
  int cz = (int)(pos[3 * i + 2] / s->HC[2]);
  if (cx < 0 || cy < 0 || cz < 0 || cx >= s->M[0] || cy >= s->M[1] || cz >= s->M[2]) return -1;
  return cx + cy * s->M[0] + cz * s->M[0] * s->M[1];
}

void FN(orc_chaining_cells)(const OrcParams* p, const R* pos, int* dims, int* cell) {
  FN(SR) s = FN(sr_setup)(p);
  dims[0] = s.M[0], dims[1] = s.M[1], dims[2] = s.M[2];
  if (cell)
    for (int i = 0; i < p->n; ++i) cell[i] = FN(cell_of)(&s, pos, i);
}

typedef struct {
  R y;
  int id;
} FN(YI);

static int FN(cmp_yi)(const void* a, const void* b) {
  const FN(YI)*u = (const FN(YI)*)a, *v = (const FN(YI)*)b;
  if (u->y < v->y) return -1;
  if (u->y > v->y) return 1;
  return (u->id > v->id) - (u->id < v->id);
}

/* Cell lists in the order the reference's linked lists are walked.  fillWithYSorting
 * (chainingMesh.cpp:20-43) inserts particle i after every node with y <= y_i, i.e. each list ends
 * up ascending in (y, particle id); fill (:45-58) pushes at the head, i.e. descending particle id.
 * The O(n^2) insertion is replaced by a counting sort + per-cell sort with the same result. */
static void FN(build_cells)(const OrcParams* p, const FN(SR) * s, const R* pos, int** start_out,
                            int** ids_out) {
  const int nc = s->M[0] * s->M[1] * s->M[2];
  int* start = (int*)calloc((size_t)nc + 1, sizeof(int));
  int* ids = (int*)malloc(sizeof(int) * (size_t)(p->n > 0 ? p->n : 1));
  int* cell = (int*)malloc(sizeof(int) * (size_t)(p->n > 0 ? p->n : 1));
  for (int i = 0; i < p->n; ++i) {
    cell[i] = FN(cell_of)(s, pos, i);
    if (cell[i] >= 0) start[cell[i] + 1]++;
  }
  for (int c = 0; c < nc; ++c) start[c + 1] += start[c];
  int* fillp = (int*)malloc(sizeof(int) * (size_t)(nc + 1));
  memcpy(fillp, start, sizeof(int) * (size_t)(nc + 1));
  for (int i = 0; i < p->n; ++i)
    if (cell[i] >= 0) ids[fillp[cell[i]]++] = i;
  for (int c = 0; c < nc; ++c) {
    int b = start[c], e = start[c + 1];
    if (e - b < 2) continue;
    if (p->ySort) {
      FN(YI)* tmp = (FN(YI)*)malloc(sizeof(FN(YI)) * (size_t)(e - b));
      for (int k = b; k < e; ++k) tmp[k - b].y = pos[3 * ids[k] + 1], tmp[k - b].id = ids[k];
      qsort(tmp, (size_t)(e - b), sizeof(FN(YI)), FN(cmp_yi));
      for (int k = b; k < e; ++k) ids[k] = tmp[k - b].id;
      free(tmp);
    } else {
      for (int k = 0; k < (e - b) / 2; ++k) {
        int t = ids[b + k];
        ids[b + k] = ids[e - 1 - k];
        ids[e - 1 - k] = t;
      }
    }
  }
  free(fillp);
  free(cell);
  *start_out = start;
  *ids_out = ids;
}

void FN(orc_chaining_order)(const OrcParams* p, const R* pos, int* order) {
  FN(SR) s = FN(sr_setup)(p);
  int *start, *ids;
  FN(build_cells)(p, &s, pos, &start, &ids);
  const int nc = s.M[0] * s.M[1] * s.M[2];
  memcpy(order, ids, sizeof(int) * (size_t)start[nc]);
  free(start);
  free(ids);
}

static inline FN(V3) FN(pair_force)(const OrcParams* p, const FN(SR) * s, const R* table, FN(V3) rij,
                                    R r2, R mi, R mj) {
  FN(V3) f;
  if (p->useTable) { /* shortRangeForceFromTable :240-245 */
    R ksi = r2 / s->delta2;
    int t = (int)ksi;
    R F = mi * mj * (table[t] + (ksi - t) * (table[t + 1] - table[t]));
    f.x = F * rij.x, f.y = F * rij.y, f.z = F * rij.z;
    return f;
  }
  /* shortRangeForce :220-238 */
  R len = SQRT(r2);
  FN(V3) dir = {rij.x / len, rij.y / len, rij.z / len};
  R G = 1 / (4 * (R)PI_R);
  R Rf = FN(ref_force)(s, p->cloudShape, len);
  R cR = -mi * mj * Rf;
  R cT = -G * mi * mj / (len * len + s->eps * s->eps);
  f.x = cT * dir.x - cR * dir.x, f.y = cT * dir.y - cR * dir.y, f.z = cT * dir.z - cR * dir.z;
  return f;
}

void FN(orc_sr_forces)(const OrcParams* p, const R* pos, const R* mass, R* sr) {
  FN(SR) s = FN(sr_setup)(p);
  R table[500];
  if (p->useTable) FN(sr_table)(p, &s, table);
  int *start, *ids;
  FN(build_cells)(p, &s, pos, &start, &ids);
  const int nc = s.M[0] * s.M[1] * s.M[2];
  const size_t n = (size_t)p->n;
  /* Particle::shortRangeForce and ::shortRangeFromNeighbor[13] (include/particle.h:11-12),
   * zeroed each call (:170-175) */
  FN(V3)* own = (FN(V3)*)calloc(n ? n : 1, sizeof(FN(V3)));
  FN(V3)* slot = (FN(V3)*)calloc((n ? n : 1) * 13, sizeof(FN(V3)));
  const R re2 = s.re * s.re;
  /* updateSRForcesThreadJob :296-322: every (particle, slot) accumulator has exactly one writer
   * cell, so cells may run concurrently without changing any summation order. */
#pragma omp parallel for schedule(dynamic, 1)
  for (int q = 0; q < nc; ++q) {
    if (start[q] == start[q + 1]) continue;
    int nb[14];
    orc_chaining_neighbors(s.M, q, nb);
    for (int li = 0; li < 14; ++li) {
      int qn = nb[li];
      if (qn == -1) continue;
      for (int a = start[q]; a < start[q + 1]; ++a) {
        const int i = ids[a];
        for (int b = start[qn]; b < start[qn + 1]; ++b) {
          const int j = ids[b];
          if (p->ySort && pos[3 * j + 1] - pos[3 * i + 1] > s.re) break; /* :309-313 */
          if (i == j) continue;                                           /* :253-255 */
          FN(V3) rij = {pos[3 * i] - pos[3 * j], pos[3 * i + 1] - pos[3 * j + 1],
                        pos[3 * i + 2] - pos[3 * j + 2]};
          R r2 = rij.x * rij.x + rij.y * rij.y + rij.z * rij.z;
          if (r2 >= re2) continue; /* :258-260 */
          FN(V3) f = FN(pair_force)(p, &s, table, rij, r2, mass[i], mass[j]);
          own[i].x += f.x, own[i].y += f.y, own[i].z += f.z; /* :269 */
          if (qn != q) {                                     /* :270-272 */
            FN(V3)* t = &slot[(size_t)j * 13 + li];
            t->x += -1 * f.x, t->y += -1 * f.y, t->z += -1 * f.z;
          }
        }
      }
    }
  }
  for (size_t i = 0; i < n; ++i) {
    /* correctAccelerations :52-54: accumulate(slots, zero) + shortRangeForce */
    FN(V3) t = {0, 0, 0};
    for (int k = 0; k < 13; ++k) {
      t.x += slot[i * 13 + k].x, t.y += slot[i * 13 + k].y, t.z += slot[i * 13 + k].z;
    }
    sr[3 * i] = t.x + own[i].x, sr[3 * i + 1] = t.y + own[i].y, sr[3 * i + 2] = t.z + own[i].z;
  }
  free(own);
  free(slot);
  free(start);
  free(ids);
}

/* ------------------------------------------------------------------------------------------- */
/* whole force evaluation and run loop                                                          */

void FN(orc_force)(const OrcParams* p, int p3m, const R* green, const R* pos, const R* mass,
                   R* density, R* potential, R* acc) {
  const size_t M = (size_t)p->nx * p->ny * p->nz;
  R* field = (R*)malloc(sizeof(R) * 3 * M);
  /* pmMethodStep pmMethod.cpp:137-144 */
  FN(orc_deposit)(p, pos, mass, density);
  FN(orc_poisson)(p, density, green, potential);
  FN(orc_field)(p, potential, field);
  FN(orc_gather)(p, pos, field, acc);
  free(field);
  if (p3m) {
    R* sr = (R*)malloc(sizeof(R) * 3 * (size_t)(p->n ? p->n : 1));
    FN(orc_sr_forces)(p, pos, mass, sr);
    for (int i = 0; i < p->n; ++i) /* correctAccelerations p3mMethod.cpp:55 */
      for (int d = 0; d < 3; ++d) acc[3 * i + d] += sr[3 * i + d] / mass[i];
    free(sr);
  }
}

int FN(orc_run)(const OrcParams* p, int p3m, const float* pos0, const float* vel0,
                const float* mass0, int simLength, R* diag, R* pos_out, R* vel_out, R* acc_out) {
  const int n = p->n;
  const size_t M = (size_t)p->nx * p->ny * p->nz;
  const R H = p->H, DT = p->DT, G = p->G, pi = (R)PI_R;
  R* pos = (R*)malloc(sizeof(R) * 3 * (size_t)n);
  R* vel = (R*)malloc(sizeof(R) * 3 * (size_t)n);
  R* acc = (R*)calloc(3 * (size_t)n, sizeof(R));
  R* intv = (R*)malloc(sizeof(R) * 3 * (size_t)n);
  R* mass = (R*)malloc(sizeof(R) * (size_t)n);
  R* green = (R*)malloc(sizeof(R) * M);
  R* density = (R*)malloc(sizeof(R) * M);
  R* potential = (R*)malloc(sizeof(R) * M);
  R expected[3] = {0, 0, 0};
  for (int i = 0; i < 3 * n; ++i) pos[i] = pos0[i], vel[i] = vel0 ? vel0[i] : 0, intv[i] = vel[i];
  for (int i = 0; i < n; ++i) mass[i] = mass0[i];
  if (diag) /* SimInfo::setInitialMomentum simInfo.cpp:120-122 (integerStepVelocity = velocity) */
    for (int i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) expected[d] += mass[i] * intv[3 * i + d];

  /* pmMethod.cpp:72-77 / p3mMethod.cpp:73-91 */
  for (int i = 0; i < 3 * n; ++i) pos[i] = pos[i] / H, vel[i] = DT * vel[i] / H;
  for (int i = 0; i < n; ++i) mass[i] = DT * DT * 4 * pi * G / (H * H * H) * mass[i];
  FN(orc_green)(p, green);
  FN(orc_force)(p, p3m, green, pos, mass, density, potential, acc);
  for (int i = 0; i < 3 * n; ++i) vel[i] += (R)0.5 * 1 * acc[i]; /* leapfrog.cpp:5-8, dt = 1 */

  int rows = 0;
  for (int t = 0; t <= simLength; ++t) {
    for (int i = 0; i < 3 * n; ++i) pos[i] += 1 * vel[i]; /* updatePositions leapfrog.cpp:21-24 */
    if (diag)
      for (int i = 0; i < 3 * n; ++i) { /* leapfrog.cpp:10-14 + unitConversions.cpp:49-54 */
        intv[i] = vel[i] + (R)0.5 * 1 * acc[i];
        intv[i] = H * intv[i] / DT;
      }
    for (int i = 0; i < 3 * n; ++i) pos[i] = H * pos[i], vel[i] = H * vel[i] / DT; /* :108 */
    int escaped = 0;
    for (int i = 0; i < n; ++i) /* pmMethod.cpp:146-156 */
      for (int d = 0; d < 3; ++d)
        if (!(pos[3 * i + d] >= 0 && pos[3 * i + d] <= p->box[d])) escaped = 1;
    if (diag) {
      R* row = diag + 12 * (size_t)rows;
      for (int i = 0; i < n; ++i) /* unitConversions.h:45-47 */
        mass[i] = (H * H * H) / (DT * DT * 4 * pi * G) * mass[i];
      R ext[3] = {0, 0, 0};
      for (int i = 0; i < n; ++i) { /* totalExternalForceOrigUnits pmMethod.cpp:158-162 */
        FN(V3) po = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
        FN(V3) e = FN(ext_field)(p, po);
        ext[0] += mass[i] * e.x, ext[1] += mass[i] * e.y, ext[2] += mass[i] * e.z;
      }
      for (int d = 0; d < 3; ++d) expected[d] += DT * ext[d]; /* simInfo.cpp:124-127 */
      R mom[3] = {0, 0, 0}, L[3] = {0, 0, 0}, ke = 0, external = 0, internal = 0;
      for (int i = 0; i < n; ++i) { /* simInfo.cpp:94-118 */
        const R* x = pos + 3 * i;
        const R* v = intv + 3 * i;
        for (int d = 0; d < 3; ++d) mom[d] += mass[i] * v[d];
        L[0] += mass[i] * (x[1] * v[2] - x[2] * v[1]);
        L[1] += mass[i] * (x[2] * v[0] - x[0] * v[2]);
        L[2] += mass[i] * (x[0] * v[1] - x[1] * v[0]);
        const R* hv = vel + 3 * i;
        ke += (R)0.5 * mass[i] * (hv[0] * hv[0] + hv[1] * hv[1] + hv[2] * hv[2]);
        FN(V3) po = {x[0], x[1], x[2]};
        external += mass[i] * FN(ext_potential)(p, po);
      }
      for (size_t c = 0; c < M; ++c) /* simInfo.cpp:56-61 */
        internal += (density[c] / (DT * DT * 4 * pi * G)) * (potential[c] * H * H / (DT * DT));
      R vol = H * H * H;
      row[0] = (R)0.5 * vol * internal + external; /* simInfo.cpp:68-69 */
      row[1] = ke;
      for (int d = 0; d < 3; ++d) row[2 + d] = mom[d], row[5 + d] = L[d], row[8 + d] = expected[d];
      row[11] = (R)escaped;
    }
    ++rows;
    if (escaped) break; /* pmMethod.cpp:108-111 */
    for (int i = 0; i < 3 * n; ++i) pos[i] = pos[i] / H, vel[i] = DT * vel[i] / H; /* :113 */
    if (diag) {
      for (int i = 0; i < n; ++i) mass[i] = DT * DT * 4 * pi * G / (H * H * H) * mass[i];
      for (int i = 0; i < 3 * n; ++i) intv[i] = DT * intv[i] / H;
    }
    FN(orc_force)(p, p3m, green, pos, mass, density, potential, acc);
    for (int i = 0; i < 3 * n; ++i) vel[i] += 1 * acc[i]; /* updateVelocities leapfrog.cpp:16-19 */
  }
  for (int i = 0; i < 3 * n; ++i) {
    if (pos_out) pos_out[i] = pos[i];
    if (vel_out) vel_out[i] = vel[i];
    if (acc_out) acc_out[i] = acc[i];
  }
  free(pos), free(vel), free(acc), free(intv), free(mass), free(green), free(density),
      free(potential);
  return rows;
}

#undef ACCW
#undef DEP
#undef FN
#undef FN1
#undef FN2
