// ctx.cuh -- the opaque context behind include/p3m_b200.h.
#pragma once

#include <vector>

#include "common.cuh"

namespace p3m {

// Per-phase CUDA-event timing that never blocks the host: phase_begin / phase_end only RECORD events (taken
// from a pool) on the context's stream; p3m_get_phase_ms synchronises once and folds every recorded interval
// into acc_ms.  Enabling it therefore does not serialise host and device between phases.
struct PhaseTimer {
  struct Interval { int ph; cudaEvent_t a, b; };
  std::vector<Interval> open_;      // recorded, not yet resolved
  std::vector<cudaEvent_t> pool;    // free events
  cudaEvent_t cur[P3M_NPHASE] = {};  // begin event of a phase that has not ended yet
  float acc_ms[P3M_NPHASE] = {};
};

template <typename T>
struct State {
  using cplx = typename CufftTypes<T>::cplx;
  // particles, sorted order (double-buffered for the permutation)
  V4<T>*posm = nullptr, *posm_alt = nullptr;
  V4<T>*vel = nullptr, *vel_alt = nullptr;
  V4<T>* acc = nullptr;     // total acceleration (mesh + short range), code units
  V4<T>* acc_sr = nullptr;  // short-range part alone (diagnostics / parity tests)
  int *id = nullptr, *id_alt = nullptr;
  // sort scratch
  uint64_t *keys = nullptr, *keys_alt = nullptr;
  uint32_t *slots = nullptr, *slots_alt = nullptr;
  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  // incremental re-sort (incsort.cu): the 32-bit keys the arrays are currently sorted by (kept from sort to sort),
  // radix histograms of the mover sort, mover count (device + pinned host copy)
  uint32_t *skeys = nullptr, *skeys_alt = nullptr;
  long long skeys_n = -1;
  int skeys_bits = 0;
  int inc_backoff = 0;  // full sorts still to do before the incremental path is tried again
  int* inc_hist = nullptr;
  int* inc_counts = nullptr;
  int* inc_counts_host = nullptr;
  cudaEvent_t zmax_event = nullptr;  // the D2H copy of the occupied-plane bound has landed (full-sort path)
  int* cell_start = nullptr;  // (1 << 3*mbits) + 1 entries
  V4<T>* aabb = nullptr;      // 2 per kPPTile-particle tile: (lo.xyz, -), (hi.xyz, -)
  // ghost particles of the two neighbouring z-slabs (dist.cu), sorted by the same key
  V4<T>*gposm = nullptr, *gposm_alt = nullptr;
  int *gid = nullptr, *gid_alt = nullptr;
  int* gcell_start = nullptr;
  V4<T>* gaabb = nullptr;
  long long n_ghost = 0, ghost_cap = 0;
  // meshes
  T* density = nullptr;    // M
  T* potential = nullptr;  // M
  cplx* spectrum = nullptr;  // (nx/2+1)*ny*nz
  T* green = nullptr;        // (nx/2+1)*ny*nz, symmetrised, includes 1/M
  T* field = nullptr;        // 3M (only after p3m_gradient)
  // slab-decomposed mesh (nranks > 1, dist_mesh.cu).  `density` / `potential` are then this rank's FFT
  // slab (planes [rank*nz/P, (rank+1)*nz/P)); the particle kernels work on the planes their own z-slab
  // of particles touches: dens_part (planes g.den_z0 ..) and pot_part (unwrapped planes g.pot_z0 ..).
  // Single GPU: dens_part == density, pot_part == potential.
  T* dens_part = nullptr;
  T* pot_part = nullptr;
  T* den_stage = nullptr;       // received density planes before they are summed into the slab
  cplx* spectrum_t = nullptr;   // transposed half spectrum [kx, ky_local, kz]
  cplx* pack = nullptr;         // all-to-all staging, P chunks of [kx, ky_local, z_local]
  cufftHandle plan_z = 0;
  bool plan_z_made = false;
  // short range
  T* sr_table = nullptr;  // 2*kSRTable: (F[t], F[t+1]-F[t]) pairs; entry 499 = (0,0)
  int* pp_items = nullptr;     // (cell, first target) pairs of the dense-cell kernel; PM-only contexts: the
                               // occupied-cell / occupied-block lists of deposit and gather
  int* pp_counters = nullptr;  // [0] items, [1] next item, [4] escape probe, [6] / [7] list lengths (PM)
  unsigned long long* pair_counts = nullptr;  // [0] checked, [1] in range
  int* flags = nullptr;    // [0] escaped, [1] out-of-mesh particles seen
  double* diag = nullptr;  // 16 doubles
  cufftHandle plan_fwd = 0, plan_inv = 0;
  bool plans = false;
  // Pruned z range (single GPU): the padding half of the mesh (isolated boundary conditions) holds no density, so
  // only the planes [0, zocc) are cleared, deposited into and 2-D transformed forward, and only the planes the gather
  // reads -- [0, pot_lo) and the periodic tail [nz - pot_tail, nz) -- are transformed back; the others are completed
  // on demand (potential readback, explicit gradient).  zocc comes from the particles (max plane seen by the sort,
  // rounded up to 16), 0 = unknown = no pruning.  Batched plans are cached per plane count.
  int zocc = 0;
  int dens_occ = 0;           // the same bound for what `density` currently holds (0: unknown, e.g. p3m_set_density)
  int dens_dirty = -1;        // leading planes of `density` that may be non-zero (-1: all)
  bool pot_partial = false;
  int pot_lo = 0, pot_tail = 0;
  struct BatchPlan { int kind, batch; cufftHandle h; };
  std::vector<BatchPlan> batch_plans;
  // fused z pass (poisson_z.cu): plan_fwd / plan_inv are then batched 2-D transforms over `fft_chunk` planes
  void* twiddle_z = nullptr;  // W_nz^k, k < nz
  int fft_chunk = 0;
};

}  // namespace p3m

// Measurement switches, read from the environment ONCE when the context is created --
// never on the per-step path.  Each selects an alternative, equally tested code path so that A/B numbers
// come from one build; none is needed for normal use.
struct p3m_tune {
  int subbits = -1;             // P3M_TUNE_SUBBITS: sub-cell bits of the P3M sort key (-1 = default)
  bool long_key = false;        // P3M_TUNE_LONGKEY: 64-bit sort key in PM-only contexts
  bool old_deposit = false;     // P3M_TUNE_OLD_DEPOSIT: warp-private-tile deposit in PM-only contexts
  bool old_gather = false;      // P3M_TUNE_OLD_GATHER: round-1 gather kernel (E tile, synchronous staging) in P3M contexts
  bool static_cuts = false;     // P3M_STATIC_CUTS: geometric layer cuts
  bool count_cuts = false;      // P3M_COUNT_CUTS: cuts by particle count also for P3M
  double particle_weight = 300.0;  // P3M_TUNE_PARTICLE_WEIGHT: mesh-side work of a particle, in pair evaluations
  long long fft_chunk_bytes = 1ll << 60;  // P3M_TUNE_FFT_CHUNK_MB: split the 2-D plane batches
  bool replicated_mesh = false; // P3M_REPLICATED_MESH: full mesh + all-reduce instead of slabs
  bool no_prune = false;        // P3M_TUNE_NO_PRUNE: clear / transform every plane of the mesh (single GPU)
  bool contig_slabs = false;    // P3M_TUNE_CONTIG_SLABS: one contiguous run of planes per rank (round-1 FFT slab layout)
  bool cufft_z = false;         // P3M_TUNE_CUFFT_Z: z leg through cuFFT + multiply kernel
  bool full_sort = false;       // P3M_TUNE_FULL_SORT: radix-sort from scratch every step
  bool scalar_pp = false;       // P3M_TUNE_SCALAR_PP: scalar-FFMA dense-cell kernel instead of the packed one
  int inc_sort_den = 12;        // P3M_TUNE_INC_SORT_DEN: the movers are merged when they are fewer than n / this, else full sort
  bool z_wide = false;          // P3M_TUNE_Z_WIDE: twice the columns per CTA in k_poisson_z at nz >= 512 (128-byte runs, 1024 threads)
  int a2a_chunks = 0;           // P3M_TUNE_A2A_CHUNKS: plane chunks of the overlapped slab FFT (0 = default)
  int dense_cell = p3m::kDenseCell;  // P3M_TUNE_DENSE_CELL: chaining cells with at least this many particles take the
                                // warp-per-64-targets kernel, the others the thread-per-target kernel
  void load();
};

struct p3m_ctx {
  p3m_params prm;
  p3m_tune tune;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool f64 = false;
  long long n = 0;        // particles
  long long cap = 0;      // allocated particle capacity
  bool have_particles = false, have_green = false, sorted = false, have_field = false;
  bool have_density = false, have_potential = false;
  // acc[] / acc_sr[] belong to the CURRENT particle order.  A re-sort or a migration permutes positions,
  // velocities and ids but not the accelerations, so it clears this flag; p3m_gather (or p3m_short_range on
  // its own) sets it again.  p3m_kick, acceleration readbacks and diagnostics refuse to run without it.
  bool have_acc = false;
  // the particle arrays are in the order of State::skeys (set by p3m_bin_sort with the 32-bit key, kept through
  // drifts and migrations, cleared by uploads / generation): the next sort only has to merge the movers
  bool order_valid = false;
  long long launches = 0;
  int steps_since_sort = 0;
  p3m::Geom<float> g32;
  p3m::Geom<double> g64;
  p3m::SRParams<float> sr32;
  p3m::SRParams<double> sr64;
  p3m::State<float> s32;
  p3m::State<double> s64;
  std::vector<double> sr_table_host;  // 500 entries, reference layout
  float mass_factor32 = 0;
  double mass_factor64 = 0;
  int num_sms = 148;
  p3m::PhaseTimer timer;
  bool timing = false;
  int count_pairs = 0;
  // all particle masses equal (true for every sampler of the reference: masses = M / n); lets the
  // short-range kernel fold the mass into its force table.  Decided on the host at upload time over the
  // GLOBAL particle set (multi-GPU: all-reduced), value in code units.
  bool uniform_mass = false;
  double uniform_mass_code = 0;
  float mass_lo = 0, mass_hi = 0;
  // p3m_get_stats: what the last force evaluation exchanged / how the last re-sort went
  double stat_migrated = 0, stat_a2a_bytes = 0, stat_den_bytes = 0, stat_pot_bytes = 0, stat_mig_bytes = 0,
         stat_ghost_bytes = 0, stat_movers = -1;
  long long full_sorts = 0, incr_sorts = 0;
  bool packed_pp = false, incr_sort = false;
  // multi-GPU (z-slabs of particles, NCCL): dist.cu
  void* nccl_comm = nullptr;  // ncclComm_t
  int rank = 0, nranks = 1;
  long long n_global = 0;
  bool fused_z = false;       // z leg of the Poisson solve = k_poisson_z (power-of-two nz), else cuFFT
  bool slab = false;          // slab-decomposed mesh + distributed FFT (else: replicated mesh, all-reduce)
  bool slab_split = false;    // every rank's FFT slab = one run of planes in the occupied half + one in the padding half
  // static plane ranges of every rank (identical on all ranks): density planes deposited by the
  // particle slab, unwrapped potential planes its gather needs
  int den_z0[P3M_MAX_RANKS] = {0}, den_nz[P3M_MAX_RANKS] = {0}, pot_z0[P3M_MAX_RANKS] = {0}, pot_nz[P3M_MAX_RANKS] = {0};
  int* dist_counts = nullptr;       // device: nranks + 1 segment starts, then nranks*nranks counts, then 4 ghost counts * nranks
  int* dist_counts_host = nullptr;  // pinned mirror
};

namespace p3m {

template <typename T>
struct Sel;
template <>
struct Sel<float> {
  static State<float>& st(p3m_ctx* c) { return c->s32; }
  static Geom<float>& g(p3m_ctx* c) { return c->g32; }
  static SRParams<float>& sr(p3m_ctx* c) { return c->sr32; }
};
template <>
struct Sel<double> {
  static State<double>& st(p3m_ctx* c) { return c->s64; }
  static Geom<double>& g(p3m_ctx* c) { return c->g64; }
  static SRParams<double>& sr(p3m_ctx* c) { return c->sr64; }
};

// phase ids (p3m_get_phase_ms)
enum Phase {
  PH_BINSORT = 0,
  PH_DEPOSIT,
  PH_FFT_FWD,
  PH_MULTIPLY,
  PH_FFT_INV,
  PH_GATHER,
  PH_SHORT_RANGE,
  PH_INTEGRATE,
  PH_GRADIENT,
  PH_COMM
};

void phase_begin(p3m_ctx* c, int ph);
void phase_end(p3m_ctx* c, int ph);
void phase_resolve(p3m_ctx* c);

// implemented one per .cu file, templated on the arithmetic type
template <typename T> int alloc_particles(p3m_ctx* c, long long n);
template <typename T> int alloc_meshes(p3m_ctx* c);
template <typename T> void free_state(p3m_ctx* c);
template <typename T> int upload_particles(p3m_ctx* c, const float* pos, const float* vel, const float* mass, long long n, int units);
template <typename T, typename O> int download_particles(p3m_ctx* c, O* pos, O* vel, O* acc, int units);
template <typename T> int bin_sort(p3m_ctx* c);
template <typename T> int deposit(p3m_ctx* c);
template <typename T> int green_init(p3m_ctx* c);
template <typename T, typename I> int green_set(p3m_ctx* c, const I* full);
template <typename T> int green_get(p3m_ctx* c, double* full);
template <typename T> int poisson(p3m_ctx* c);
template <typename T> int gradient(p3m_ctx* c);
template <typename T> int gather(p3m_ctx* c);
template <typename T> int short_range(p3m_ctx* c);
template <typename T> int sr_table_upload(p3m_ctx* c);
template <typename T> int kick(p3m_ctx* c, double f);
template <typename T> int drift(p3m_ctx* c);
template <typename T> int diagnostics(p3m_ctx* c, double* out);
template <typename T> int escaped_now(p3m_ctx* c, int* escaped);
template <typename T> int get_cells(p3m_ctx* c, int32_t* mesh_cell, int32_t* chain_cell, int32_t* order);
template <typename T> int get_acc_parts(p3m_ctx* c, double* acc_pm, double* acc_sr);
template <typename T> int get_sample(p3m_ctx* c, const int32_t* ids, long long m, double* pos, double* acc, double* acc_sr);
template <typename T> int add_acceleration(p3m_ctx* c, const float* a, int units);
int fft3d_c2c(int nz, int ny, int nx, const float* in, float* out, int inverse);
// dist.cu
int comm_unique_id(void* out128);
void dist_set_cuts(p3m_ctx* c);
int dist_init(p3m_ctx* c, const void* unique_id, int rank, int nranks);
void dist_destroy(p3m_ctx* c);
template <typename T> int dist_migrate(p3m_ctx* c, bool exchange);
template <typename T> int dist_ghosts(p3m_ctx* c);
template <typename T> int dist_allreduce_density(p3m_ctx* c);
// poisson_z.cu: forward z FFT + influence-function multiply + inverse z FFT in one pass
bool fused_z_supported(int nz);
template <typename T> int fused_z_init(p3m_ctx* c);
template <typename T> int fused_z_pass(p3m_ctx* c, void* spec, const T* green, long long ncols, int nz_in = 0);
template <typename T> int complete_potential(p3m_ctx* c);  // poisson.cu: materialise the planes the pruned solve skipped
int fft_chunk_planes(long long plane_bytes, int planes, long long budget);
// dist_mesh.cu: slab-decomposed mesh
template <typename T> int slab_setup(p3m_ctx* c);             // after dist_init: plane ranges, buffers, plans
template <typename T> void slab_free(p3m_ctx* c);
template <typename T> int slab_replan(p3m_ctx* c);           // after the layer cuts moved
// dist.cu: equal-COUNT layer cuts from the full particle set every rank was handed (clustered sets)
template <typename T> int dist_balance_cuts(p3m_ctx* c, const float* pos, long long n, int units);
template <typename T> int sort_incremental(p3m_ctx* c, int keybits, bool* done);  // incsort.cu
template <typename T> int migrate_sorted_keys(p3m_ctx* c, const uint32_t* stayer_slots, long long keep, long long n_new);
template <typename T> int dist_cuts_from_weights(p3m_ctx* c, const double* weight, int layers);  // + slab_replan
// ics.cu: device-side initial conditions
template <typename T> int generate_particles(p3m_ctx* c, const p3m_ic* ic);
int sample_particles(const p3m_ic* ic, long long first, long long count, float* pos, float* vel, float* mass);
// directsum.cu: N4 brute-force accuracy oracle
template <typename T> int direct_sum(p3m_ctx* c, int mode, const double* tpos, long long m, double eps, double* out);
template <typename T> int slab_reduce_density(p3m_ctx* c);    // dens_part of all ranks -> density slabs
template <typename T> int slab_poisson(p3m_ctx* c);           // distributed FFT Poisson solve
template <typename T> int slab_spread_potential(p3m_ctx* c);  // potential slabs -> pot_part of all ranks

int dist_allreduce(p3m_ctx* c, void* buf, size_t count, int kind /*0 int max, 1 double sum*/);
int dist_allreduce_host_imax(p3m_ctx* c, int* v, int count /* <= 8 */);
template <typename T> int upload_particles_ids(p3m_ctx* c, const float* pos, const float* vel, const float* mass, const int32_t* ids, long long n, int units);
template <typename T> int download_local(p3m_ctx* c, int32_t* ids, float* pos, float* vel, float* acc, int units);
template <typename T, typename O> int get_mesh(p3m_ctx* c, const T* dev, O* out, long long count);
template <typename T, typename I> int set_mesh(p3m_ctx* c, T* dev, const I* in, long long count);

#define P3M_DISPATCH(c, fn, ...) ((c)->f64 ? fn<double>((c), ##__VA_ARGS__) : fn<float>((c), ##__VA_ARGS__))

#define P3M_LAUNCH_CHECK(c)                   \
  do {                                        \
    (c)->launches++;                          \
    P3M_CUDA(cudaGetLastError());             \
  } while (0)

}  // namespace p3m
