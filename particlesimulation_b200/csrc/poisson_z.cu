// poisson_z.cu -- the z leg of the Poisson solve as ONE hand-written pass (A2 + A3 fused):
//     forward 1-D FFT along z  ->  multiply by the influence function  ->  inverse 1-D FFT along z
// on the half spectrum produced by the batched 2-D R2C transforms (cuFFT) of the mesh planes.
//
// The reference runs forward FFT, a separate multiply loop over 3 complex meshes and the inverse FFT
// (source/grid.cpp:50-56, source/pmMethod.cpp:340-350); a library 3-D transform + multiply kernel streams the
// 8 B/cell half spectrum through HBM five times for these three steps (z pass of the forward transform,
// multiply, z pass of the inverse).  Here each column kz = 0..Nz-1 of `C` adjacent (kx, ky) points is
// loaded once, transformed in registers + shared memory (Stockham radix-8/4 stages, twiddles from a table
// evaluated in double), scaled by G_sym/M while it sits in registers, transformed back and stored once:
// 8 B read + 8 B written + 4 B of table per spectrum element.
//
// Layout: element (col, kz) at spec[col + ncols * kz], col = kx + nxh * ky -- the natural cuFFT layout on one
// GPU (ncols = nxh * ny) and the transposed slab layout [kx, ky_local, kz] of dist_mesh.cu
// (ncols = nxh * ny / P).  The `C` columns of a CTA are contiguous in memory for every kz (64-128 B runs).
//
// A thread owns 8 elements of one column.  In a radix-R stage it performs 8/R butterflies
//     j = t + u N/8,  inputs j + r N/R,  twiddle W_N^{(j mod Ns) r N/(Ns R)},  outputs (j - j mod Ns) R + j mod Ns + r Ns
// (Stockham autosort: natural order in, natural order out).  The last forward stage leaves exactly the
// elements in registers that the first inverse stage (radices reversed) consumes, so the multiply costs no
// exchange.  The inverse is conj(FFT(conj(.))).
#include <cstdlib>
#include <type_traits>

#include "ctx.cuh"

namespace p3m {

template <typename T>
struct Cx {
  T x, y;
};

namespace {

template <typename T>
__device__ __forceinline__ Cx<T> cmul(Cx<T> a, Cx<T> b) {
  return Cx<T>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}

template <typename T>
__device__ __forceinline__ void dft2(Cx<T>& a, Cx<T>& b) {
  const Cx<T> t = a;
  a = Cx<T>{t.x + b.x, t.y + b.y};
  b = Cx<T>{t.x - b.x, t.y - b.y};
}

// forward 4-point DFT in place, natural order
template <typename T>
__device__ __forceinline__ void dft4(Cx<T>& v0, Cx<T>& v1, Cx<T>& v2, Cx<T>& v3) {
  const Cx<T> t0{v0.x + v2.x, v0.y + v2.y}, t1{v0.x - v2.x, v0.y - v2.y};
  const Cx<T> t2{v1.x + v3.x, v1.y + v3.y}, d{v1.x - v3.x, v1.y - v3.y};
  const Cx<T> t3{d.y, -d.x};  // -i d
  v0 = Cx<T>{t0.x + t2.x, t0.y + t2.y};
  v2 = Cx<T>{t0.x - t2.x, t0.y - t2.y};
  v1 = Cx<T>{t1.x + t3.x, t1.y + t3.y};
  v3 = Cx<T>{t1.x - t3.x, t1.y - t3.y};
}

template <typename T, int R>
struct Dft;
template <typename T>
struct Dft<T, 2> {
  static __device__ __forceinline__ void run(Cx<T>* v) { dft2(v[0], v[1]); }
};
template <typename T>
struct Dft<T, 4> {
  static __device__ __forceinline__ void run(Cx<T>* v) { dft4(v[0], v[1], v[2], v[3]); }
};
template <typename T>
struct Dft<T, 8> {
  static __device__ __forceinline__ void run(Cx<T>* v) {
    dft4(v[0], v[2], v[4], v[6]);  // E_k in v[2k]
    dft4(v[1], v[3], v[5], v[7]);  // O_k in v[2k+1]
    const T h = T(0.70710678118654752440);
    const Cx<T> o0 = v[1];
    const Cx<T> o1{(v[3].x + v[3].y) * h, (v[3].y - v[3].x) * h};    // W8^1 O_1
    const Cx<T> o2{v[5].y, -v[5].x};                                   // W8^2 O_2 = -i O_2
    const Cx<T> o3{(v[7].y - v[7].x) * h, -(v[7].x + v[7].y) * h};   // W8^3 O_3
    const Cx<T> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    v[0] = Cx<T>{e0.x + o0.x, e0.y + o0.y}, v[4] = Cx<T>{e0.x - o0.x, e0.y - o0.y};
    v[1] = Cx<T>{e1.x + o1.x, e1.y + o1.y}, v[5] = Cx<T>{e1.x - o1.x, e1.y - o1.y};
    v[2] = Cx<T>{e2.x + o2.x, e2.y + o2.y}, v[6] = Cx<T>{e2.x - o2.x, e2.y - o2.y};
    v[3] = Cx<T>{e3.x + o3.x, e3.y + o3.y}, v[7] = Cx<T>{e3.x - o3.x, e3.y - o3.y};
  }
};

// stage radices for N = 2^LOGN (everything below is resolved at compile time: no index arithmetic at run time)
template <int LOGN> struct Plan;
template <> struct Plan<4>  { static constexpr int n = 2; static constexpr int r(int s) { constexpr int t[4] = {4, 4, 1, 1}; return t[s]; } };
template <> struct Plan<5>  { static constexpr int n = 2; static constexpr int r(int s) { constexpr int t[4] = {8, 4, 1, 1}; return t[s]; } };
template <> struct Plan<6>  { static constexpr int n = 2; static constexpr int r(int s) { constexpr int t[4] = {8, 8, 1, 1}; return t[s]; } };
template <> struct Plan<7>  { static constexpr int n = 3; static constexpr int r(int s) { constexpr int t[4] = {8, 4, 4, 1}; return t[s]; } };
template <> struct Plan<8>  { static constexpr int n = 3; static constexpr int r(int s) { constexpr int t[4] = {8, 8, 4, 1}; return t[s]; } };
template <> struct Plan<9>  { static constexpr int n = 3; static constexpr int r(int s) { constexpr int t[4] = {8, 8, 8, 1}; return t[s]; } };
template <> struct Plan<10> { static constexpr int n = 4; static constexpr int r(int s) { constexpr int t[4] = {8, 8, 4, 4}; return t[s]; } };

// radix of stage S and the product Ns of the radices before it, forward order or reversed (inverse)
template <int LOGN, bool INV>
constexpr int stage_radix(int S) { return Plan<LOGN>::r(INV ? Plan<LOGN>::n - 1 - S : S); }
template <int LOGN, bool INV>
constexpr int stage_ns(int S) {
  int ns = 1;
  for (int i = 0; i < S; ++i) ns *= stage_radix<LOGN, INV>(i);
  return ns;
}

// natural index of register slot (u, r) of a radix-R stage: j + r N/R with j = t + u N/8
template <int N, int R>
__device__ __forceinline__ int slot_index(int t, int u, int r) {
  return t + u * (N / 8) + r * (N / R);
}

// Shared-memory slot of element `pos` of column c.  The lanes of a warp hold C columns x 32/C consecutive
// threads t, whose elements are either consecutive (stride 1) or one radix apart (store of a first stage: 8,
// or 4 when the plan ends in a radix-4 stage, which opens the inverse).  Row pos lives at row pos + pos/8, so
// strides 1 and 8 walk through consecutive rows of C words and a warp covers 32 distinct banks; stride 4 is
// left with a 2-way conflict on that one store (tests/test_zpass_model.py enumerates every access).  A plain
// odd pitch left 43 % of the wavefronts as conflicts (ncu r01n).
template <int C>
__device__ __forceinline__ int sidx(int pos, int c) {
  return (pos + (pos >> 3)) * C + c;
}

template <typename T, int N, int C, int R>
__device__ __forceinline__ void stage_load(Cx<T> (&e)[8], const T* re, const T* im, int t, int c) {
#pragma unroll
  for (int u = 0; u < 8 / R; ++u)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int pos = sidx<C>(slot_index<N, R>(t, u, r), c);
      e[u * R + r] = Cx<T>{re[pos], im[pos]};
    }
}

template <typename T, int N, int R, int Ns>
__device__ __forceinline__ void stage_compute(Cx<T> (&e)[8], const Cx<T>* tw, int t) {
#pragma unroll
  for (int u = 0; u < 8 / R; ++u) {
    if (Ns > 1) {
      const int k = (t + u * (N / 8)) & (Ns - 1);
      const int step = k * (N / (R * Ns));  // k N / (Ns R)
#pragma unroll
      for (int r = 1; r < R; ++r) e[u * R + r] = cmul(e[u * R + r], tw[r * step]);
    }
    Dft<T, R>::run(&e[u * R]);
  }
}

template <typename T, int N, int C, int R, int Ns>
__device__ __forceinline__ void stage_store(const Cx<T> (&e)[8], T* re, T* im, int t, int c) {
#pragma unroll
  for (int u = 0; u < 8 / R; ++u) {
    const int j = t + u * (N / 8);
    const int k = j & (Ns - 1);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int pos = sidx<C>(j0 + r * Ns, c);
      re[pos] = e[u * R + r].x, im[pos] = e[u * R + r].y;
    }
  }
}

// stages S .. n-1 of one transform; the first stage starts from registers, the last one ends in registers
template <typename T, int LOGN, int C, bool INV, int S>
__device__ __forceinline__ void run_stages(Cx<T> (&e)[8], const Cx<T>* tw, T* re, T* im, int t, int c) {
  constexpr int N = 1 << LOGN;
  constexpr int R = stage_radix<LOGN, INV>(S), Ns = stage_ns<LOGN, INV>(S);
  if (S > 0) {
    stage_load<T, N, C, R>(e, re, im, t, c);
    __syncthreads();  // everybody has read the previous stage before anyone overwrites it
  }
  stage_compute<T, N, R, Ns>(e, tw, t);
  if constexpr (S < Plan<LOGN>::n - 1) {
    stage_store<T, N, C, R, Ns>(e, re, im, t, c);
    __syncthreads();
    run_stages<T, LOGN, C, INV, S + 1>(e, tw, re, im, t, c);
  }
}

template <typename T, int LOGN, int C>
__global__ void __launch_bounds__(C * (1 << LOGN) / 8)
k_poisson_z(Cx<T>* __restrict__ spec, const T* __restrict__ green, const Cx<T>* __restrict__ twg, long long ncols,
            int nz_in) {
  constexpr int N = 1 << LOGN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cx<T>* tw = reinterpret_cast<Cx<T>*>(smem_raw);
  T* re = reinterpret_cast<T*>(tw + N);
  T* im = re + (size_t)(N + N / 8) * C;
  const int c = threadIdx.x % C, t = threadIdx.x / C;
  const long long col = (long long)blockIdx.x * C + c;
  const bool valid = col < ncols;
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = twg[i];
  Cx<T> e[8];
  constexpr int R0 = Plan<LOGN>::r(0), RL = Plan<LOGN>::r(Plan<LOGN>::n - 1);
  Cx<T>* __restrict__ base = spec + (valid ? col : 0);
  const T* __restrict__ gbase = green + (valid ? col : 0);

  // ---- forward: first stage reads global memory -------------------------------------------------------------
#pragma unroll
  for (int u = 0; u < 8 / R0; ++u)
#pragma unroll
    for (int r = 0; r < R0; ++r)
      // planes >= nz_in are known to be zero (zero padding of the isolated boundary conditions): not even read
      e[u * R0 + r] = valid && slot_index<N, R0>(t, u, r) < nz_in ? base[ncols * slot_index<N, R0>(t, u, r)] : Cx<T>{0, 0};
  run_stages<T, LOGN, C, false, 0>(e, tw, re, im, t, c);

  // ---- influence function (A3) on the registers, conjugated for the inverse ------------------------------
#pragma unroll
  for (int u = 0; u < 8 / RL; ++u)
#pragma unroll
    for (int r = 0; r < RL; ++r) {
      const T gk = valid ? gbase[ncols * slot_index<N, RL>(t, u, r)] : T(0);
      e[u * RL + r] = Cx<T>{e[u * RL + r].x * gk, -(e[u * RL + r].y * gk)};
    }

  // ---- inverse = conj(FFT(conj(.))), radices reversed; its last stage (radix R0, Ns = N / R0) leaves the
  //      natural indices j + r N/R0 in the registers
  run_stages<T, LOGN, C, true, 0>(e, tw, re, im, t, c);
  if (valid) {
#pragma unroll
    for (int u = 0; u < 8 / R0; ++u)
#pragma unroll
      for (int r = 0; r < R0; ++r)
        base[ncols * slot_index<N, R0>(t, u, r)] = Cx<T>{e[u * R0 + r].x, -e[u * R0 + r].y};
  }
}

template <typename T>
__global__ void k_twiddles(Cx<T>* tw, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double s, c;
  sincospi(2.0 * (double)i / (double)N, &s, &c);
  tw[i] = Cx<T>{(T)c, (T)(-s)};  // W_N^i = exp(-2 pi i / N)
}

int log2_exact(int N) {
  int k = 0;
  while ((1 << k) < N) ++k;
  return (1 << k) == N ? k : -1;
}

template <typename T, int LOGN, int C>
int launch_z(p3m_ctx* c, void* spec, const T* green, long long ncols, int nz_in) {
  State<T>& s = Sel<T>::st(c);
  constexpr int N = 1 << LOGN;
  const size_t smem = sizeof(Cx<T>) * (size_t)N + 2 * sizeof(T) * (size_t)(N + N / 8) * C;
  auto kern = k_poisson_z<T, LOGN, C>;
  P3M_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (ncols + C - 1) / C;
  kern<<<(unsigned)blocks, C * N / 8, smem, c->stream>>>(reinterpret_cast<Cx<T>*>(spec), green,
                                                        reinterpret_cast<const Cx<T>*>(s.twiddle_z), ncols,
                                                        nz_in > 0 && nz_in < N ? nz_in : N);
  P3M_LAUNCH_CHECK(c);
  return 0;
}

}  // namespace

// can the fused z pass handle this mesh?
bool fused_z_supported(int nz) {
  const int k = log2_exact(nz);
  return k >= 4 && k <= 10;
}

template <typename T>
int fused_z_init(p3m_ctx* c) {
  State<T>& s = Sel<T>::st(c);
  const int N = c->prm.nz;
  if (s.twiddle_z) cudaFree(s.twiddle_z);
  s.twiddle_z = nullptr;
  P3M_CUDA(cudaMalloc(&s.twiddle_z, sizeof(Cx<T>) * (size_t)N));
  k_twiddles<T><<<(N + 127) / 128, 128, 0, c->stream>>>(reinterpret_cast<Cx<T>*>(s.twiddle_z), N);
  P3M_LAUNCH_CHECK(c);
  return 0;
}

// forward z FFT, multiply by `green`, inverse z FFT, in place on `spec` ([col + ncols * kz]).
// N/8 threads per column, C columns per CTA (one contiguous C*8-byte run for every kz), 512 threads.
template <typename T>
int fused_z_pass(p3m_ctx* c, void* spec, const T* green, long long ncols, int nz_in) {
  constexpr bool D = sizeof(T) == 8;
  switch (log2_exact(c->prm.nz)) {
    case 4: return launch_z<T, 4, 32>(c, spec, green, ncols, nz_in);
    case 5: return launch_z<T, 5, 32>(c, spec, green, ncols, nz_in);
    case 6: return launch_z<T, 6, 32>(c, spec, green, ncols, nz_in);
    case 7: return launch_z<T, 7, 32>(c, spec, green, ncols, nz_in);
    case 8: return launch_z<T, 8, D ? 8 : 16>(c, spec, green, ncols, nz_in);
    case 9:
      if (c->tune.z_wide && !D) return launch_z<T, 9, 16>(c, spec, green, ncols, nz_in);  // 128-byte runs, 1024 threads
      return launch_z<T, 9, D ? 4 : 8>(c, spec, green, ncols, nz_in);
    case 10:
      if (c->tune.z_wide && !D) return launch_z<T, 10, 8>(c, spec, green, ncols, nz_in);
      return launch_z<T, 10, 4>(c, spec, green, ncols, nz_in);
    default: return fail(P3M_EINVAL, "fused z pass: nz = %d is not a power of two in [16, 1024]", c->prm.nz);
  }
}

template int fused_z_init<float>(p3m_ctx*);
template int fused_z_init<double>(p3m_ctx*);
template int fused_z_pass<float>(p3m_ctx*, void*, const float*, long long, int);
template int fused_z_pass<double>(p3m_ctx*, void*, const double*, long long, int);

}  // namespace p3m
