// directsum.cu -- N4: the O(N m) brute-force accuracy oracle on the device.
//
// The reference's only physical accuracy pin is "P3M (or PM) versus the direct sum" (source/ppMethod.cpp:88-125
// ppMethodLeapfrog, source/demos.cpp:593-727, script/p3m_accuracy.py:32-36 -- mean relative force error).  This
// file evaluates, in DOUBLE precision and without any of the machinery of the fast path (no chaining mesh, no
// sort order, no bounding boxes, no replicated tables, no packed arithmetic), for m target points:
//   * the short-range sum the PP kernels compute (tabulated or analytic law of the context), and
//   * the softened Newtonian direct sum in code units.
// One CTA per target walks all particles of the rank; partial sums are reduced in a fixed order.
#include <vector>

#include "ctx.cuh"

namespace p3m {

namespace {

struct SumCfg {
  int mode, use_table, cloud;
  double re2, inv_delta2, a, eps2;
};

__device__ inline double ref_force_d(const SumCfg& s, double r) {
  const double G = 0.07957747154594767;  // 1 / (4 pi)
  const double a = s.a;
  if (s.cloud == P3M_S1) {  // referenceForceS1, source/p3mMethod.cpp:194-201
    if (r >= a) return G / (r * r);
    const double q = r / a;
    return G / (a * a) * (8 * r / a - 9 * r * r / (a * a) + 2 * q * q * q * q);
  }
  const double u = 2 * r / a;  // referenceForceS2, :203-218
  const double u2 = u * u, u3 = u2 * u, u4 = u2 * u2, u5 = u4 * u, u6 = u3 * u3;
  if (u <= 1) return G / (35 * a * a) * (224 * u - 224 * u3 + 70 * u4 + 48 * u5 - 21 * u6);
  if (u <= 2) return G / (35 * a * a) * (12 / u2 - 224 + 896 * u - 840 * u2 + 224 * u3 + 70 * u4 - 48 * u5 + 7 * u6);
  return G / (r * r);
}

template <typename T>
__global__ void __launch_bounds__(256)
k_direct_sum(const V4<T>* __restrict__ posm, long long n, const double* __restrict__ tpos, SumCfg s,
             const double* __restrict__ table, double* __restrict__ out) {
  const long long t = blockIdx.x;
  const double tx = tpos[3 * t], ty = tpos[3 * t + 1], tz = tpos[3 * t + 2];
  const double G = 0.07957747154594767;
  double ax = 0, ay = 0, az = 0;
  for (long long j = threadIdx.x; j < n; j += 256) {
    const V4<T> p = posm[j];
    const double dx = tx - (double)p.x, dy = ty - (double)p.y, dz = tz - (double)p.z;  // r_ij = x_i - x_j
    const double r2 = dx * dx + dy * dy + dz * dz;
    if (!(r2 > 0.0)) continue;
    double f;
    if (s.mode == P3M_SUM_CUTOFF_SHELL) {
      // the short-range law is discontinuous at the cutoff: a source this close to it is in or out depending on
      // the rounding of r^2 (in the reference's own fp32 arithmetic as well); count them
      if (fabs(r2 / s.re2 - 1.0) <= s.eps2) ax += 1.0;
      continue;
    }
    if (s.mode == P3M_SUM_NEWTON) {
      const double q = r2 + s.eps2;
      f = -G * (double)p.w / (q * sqrt(q));
    } else {
      if (!(r2 < s.re2)) continue;
      if (s.use_table) {  // shortRangeForceFromTable :240-245
        const double xi = r2 * s.inv_delta2;
        int k = (int)xi;
        k = k > kSRTable - 2 ? kSRTable - 2 : k;
        f = (double)p.w * (table[k] + (xi - k) * (table[k + 1] - table[k]));
      } else {  // shortRangeForce :220-238, divided by m_i
        const double r = sqrt(r2);
        f = (double)p.w * (ref_force_d(s, r) - G / (r2 + s.eps2)) / r;
      }
    }
    ax += f * dx, ay += f * dy, az += f * dz;
  }
  __shared__ double red[3][256];
  red[0][threadIdx.x] = ax, red[1][threadIdx.x] = ay, red[2][threadIdx.x] = az;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w)
      for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x < 3) out[3 * t + threadIdx.x] = red[threadIdx.x][0];
}

}  // namespace

template <typename T>
int direct_sum(p3m_ctx* c, int mode, const double* tpos, long long m, double eps, double* out) {
  if (m == 0) return 0;
  State<T>& s = Sel<T>::st(c);
  const SRParams<T>& sp = Sel<T>::sr(c);
  SumCfg cfg{mode, sp.use_table, sp.cloud, (double)sp.re2, (double)sp.inv_delta2, (double)sp.a,
             mode == P3M_SUM_NEWTON ? eps * eps : (double)sp.eps2};
  if (mode == P3M_SUM_CUTOFF_SHELL) cfg.eps2 = eps > 0 ? eps : 2e-6;  // relative half-width of the shell in r^2
  double *d_t = nullptr, *d_out = nullptr, *d_tab = nullptr;
  P3M_CUDA(cudaMallocAsync((void**)&d_t, sizeof(double) * 3 * (size_t)m, c->stream));
  P3M_CUDA(cudaMallocAsync((void**)&d_out, sizeof(double) * 3 * (size_t)m, c->stream));
  P3M_CUDA(cudaMallocAsync((void**)&d_tab, sizeof(double) * kSRTable, c->stream));
  P3M_CUDA(cudaMemcpyAsync(d_t, tpos, sizeof(double) * 3 * (size_t)m, cudaMemcpyHostToDevice, c->stream));
  std::vector<double> tab(kSRTable, 0.0);
  if (c->sr_table_host.size() == (size_t)kSRTable) tab = c->sr_table_host;
  P3M_CUDA(cudaMemcpyAsync(d_tab, tab.data(), sizeof(double) * kSRTable, cudaMemcpyHostToDevice, c->stream));
  k_direct_sum<T><<<(unsigned)m, 256, 0, c->stream>>>(s.posm, c->n, d_t, cfg, d_tab, d_out);
  P3M_LAUNCH_CHECK(c);
  P3M_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * 3 * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
  P3M_CUDA(cudaFreeAsync(d_t, c->stream));
  P3M_CUDA(cudaFreeAsync(d_out, c->stream));
  P3M_CUDA(cudaFreeAsync(d_tab, c->stream));
  P3M_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

template int direct_sum<float>(p3m_ctx*, int, const double*, long long, double, double*);
template int direct_sum<double>(p3m_ctx*, int, const double*, long long, double, double*);

}  // namespace p3m
