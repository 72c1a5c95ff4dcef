/* include/p3m_b200.h -- the drop-in boundary of the B200-native P3M / PM force step.
 *
 * A plain C ABI (extern "C", pointers + sizes, no C++/torch types) implemented by
 * particlesimulation_b200/libp3m_b200.so (hand-written sm_100a CUDA kernels + cuFFT).  The reference
 * (AleksyBalazinski/ParticleSimulation) has no FFI layer -- its boundary is the C++ class API -- so
 * each entry point below names the reference member function(s) whose body it replaces; the
 * API-compatible C++ classes in particlesimulation_b200/host/ (PMMethod, P3MMethod, ...) are thin
 * callers of this ABI.  Reference citations are file:line under the reference checkout.
 *
 * Conventions
 *  - every function returns 0 on success, a negative P3M_E* code otherwise; p3m_last_error() gives
 *    the message of the last failure on the calling thread.  No exceptions cross the boundary.
 *  - a context is bound to one GPU and must be driven by one host thread at a time (the reference's
 *    run loops are not re-entrant either: file-scope timers, shared Grid).
 *  - "host" pointers are ordinary host memory (pinned or pageable); particle arrays are packed xyz
 *    triples (Vec3 layout, include/vec3.h:5-8) in ORIGINAL particle order, i.e. index i here is
 *    particles[i] of PMMethod::getParticles() (include/pmMethod.h:33).
 *  - meshes are x-fastest, flat = x + y*Nx + z*Nx*Ny (include/grid.h:52-54).
 *  - all work is queued on the context's CUDA stream; readbacks synchronise that stream.
 *  - there is NO CPU fallback: creating a context without a CUDA device fails with P3M_ENODEV.
 */
#ifndef P3M_B200_H
#define P3M_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P3M_OK 0
#define P3M_EINVAL (-1)   /* bad argument / unknown enum (reference: std::invalid_argument) */
#define P3M_ENODEV (-2)   /* no usable CUDA device */
#define P3M_ECUDA (-3)    /* CUDA / cuFFT / NCCL runtime error */
#define P3M_ESTATE (-4)   /* call order violated (e.g. force before particles were set) */
#define P3M_ERANGE (-5)   /* a particle left the region the mesh kernels can address */

/* include/pmConfig.h:3-5, include/greensFunctions.h:7 -- same enumerator order as the reference */
enum { P3M_NGP = 0, P3M_CIC = 1, P3M_TSC = 2 };
enum { P3M_TWO_POINT = 0, P3M_FOUR_POINT = 1 };
enum { P3M_DISCRETE_LAPLACIAN = 0, P3M_S1_OPTIMAL = 1, P3M_S2_OPTIMAL = 2, P3M_POOR_MAN = 3 };
enum { P3M_S1 = 0, P3M_S2 = 1 };
enum { P3M_EXT_NONE = 0, P3M_EXT_SPH_RAD_DECR = 1 }; /* source/externalFields.cpp:4-15 */
enum { P3M_F32 = 0, P3M_F64 = 1 };
enum { P3M_UNITS_ORIGINAL = 0, P3M_UNITS_CODE = 1 };

/* Everything the constructors PMMethod(...) (include/pmMethod.h:14-26) and P3MMethod(...)
 * (include/p3mMethod.h:13-27) take, as a POD.  Physical quantities are in ORIGINAL units exactly as
 * the reference's callers pass them (source/demos.cpp:736-776, 897-951). */
typedef struct p3m_params {
  int32_t nx, ny, nz;        /* Grid(gridPoints) include/grid.h:12 */
  float box[3];              /* effectiveBoxSize == compBoxSize */
  float H, DT, G;            /* cell size, time step, gravitational constant */
  int32_t assignment;        /* P3M_NGP | P3M_CIC | P3M_TSC */
  int32_t fd_scheme;         /* P3M_TWO_POINT | P3M_FOUR_POINT */
  int32_t greens_function;   /* P3M_DISCRETE_LAPLACIAN | ... */
  float particle_diameter;   /* PMMethod particleDiameter / P3MMethod particleDiameter */
  int32_t p3m;               /* 0: PM only (PMMethod), 1: P3M (P3MMethod wraps PMMethod) */
  float cutoff_radius;       /* P3MMethod cutoffRadius (re) */
  float softening;           /* P3MMethod softeningLength */
  int32_t cloud_shape;       /* P3M_S1 | P3M_S2 */
  int32_t use_sr_table;      /* P3MMethod useSRForceTable (default true, include/p3mMethod.h:25) */
  int32_t ext_kind;          /* externalField as a POD: P3M_EXT_NONE | P3M_EXT_SPH_RAD_DECR */
  float ext_center[3];
  float ext_R, ext_M;        /* sphRadDecrField(pos, center, R, M, G) */
  int32_t precision;         /* P3M_F32 (the reference's precision) | P3M_F64 */
  int32_t unit_roundtrip;    /* 1: reproduce the per-step code->original->code rounding of the run
                                loops (source/pmMethod.cpp:94,113); 0: stay in code units */
  int32_t green_zero_degenerate; /* 1 (recommended): G := 0 where every k_i in {0, N_i/2}: GreenOptimal
                                is 0/0 rounding noise there and the mode carries no force */
  int32_t device;            /* CUDA device ordinal; -1 = current device */
  int32_t timing;            /* 1: bracket every phase with CUDA events (p3m_get_phase_ms) */
  float sr_particle_diameter; /* P3MMethod's own particleDiameter (source/p3mMethod.cpp:36: the cloud size of
                                the short-range reference force) when it differs from PMMethod's (the cloud
                                size of the influence function, source/pmMethod.cpp:56); 0 = the same */
} p3m_params;

typedef struct p3m_ctx p3m_ctx;

const char* p3m_last_error(void);
int p3m_version(void);

/* defaults as the reference demos use them: TSC, TWO_POINT, S1_OPTIMAL, DT = 1, table on */
void p3m_default_params(p3m_params* p);

/* PMMethod::PMMethod / P3MMethod::P3MMethod + Grid::Grid + FFTAdapter construction
 * (source/pmMethod.cpp:30-60, source/p3mMethod.cpp:19-48, source/grid.cpp:5-19,
 * source/chainingMesh.cpp:6-18): allocates device meshes, cuFFT plans, the short-range table. */
int p3m_create(const p3m_params* params, p3m_ctx** out);
int p3m_destroy(p3m_ctx* ctx);

/* ---- multi-GPU: z-slab decomposition of the particles, one process per GPU, NCCL.
 * Nothing of this exists in the reference.  Rank 0 obtains an id with p3m_comm_unique_id and hands it
 * to every rank by any means (bench.py: torch.distributed broadcast); all ranks then call
 * p3m_create_dist collectively.  Afterwards the SAME calls as on one GPU drive the run (they become
 * collective): p3m_set_particles takes the full particle set on every rank and keeps this rank's
 * slab; p3m_force / p3m_step migrate particles, exchange ghost layers, move density planes to the ranks
 * that own them in the slab-decomposed FFT (2-D FFTs on local planes, all-to-all transpose, 1-D FFTs
 * along z, and back) and return the potential planes each rank's particles need, all over NVLink;
 * id-indexed readbacks fill the entries of the particles the rank holds and leave the rest 0, mesh
 * readbacks fill the planes of the rank's FFT slab and leave the rest 0 (sum over ranks = full mesh).
 * The mesh is slab-decomposed when nz and ny are multiples of nranks; otherwise every rank keeps a full
 * mesh and the density is all-reduced (small meshes only). */
#define P3M_UNIQUE_ID_BYTES 128
#define P3M_MAX_RANKS 8   /* one NVSwitch node; larger rank counts are rejected with P3M_EINVAL */
int p3m_comm_unique_id(void* out_bytes128);
int p3m_create_dist(const p3m_params* params, const void* unique_id_bytes128, int rank, int nranks,
                    p3m_ctx** out);
/* this rank's particles (p3m_num_particles of them) in local order with their global ids */
int p3m_get_local(p3m_ctx* ctx, int32_t* ids, float* pos, float* vel, float* acc, int units);
/* upload a subset with explicit global ids (e.g. what p3m_get_local returned); the next force
 * evaluation migrates whatever does not belong to this rank's slab */
int p3m_set_particles_ids(p3m_ctx* ctx, const float* pos, const float* vel, const float* mass,
                          const int32_t* ids, int64_t n, int units);
int64_t p3m_num_global(const p3m_ctx* ctx);
/* host-only: the binning layers along z and their cuts for `nranks` ranks (cuts[r]..cuts[r+1] belongs to
 * rank r); needs no device, identical on every rank.  layers_out = number of layers that are cut. */
int p3m_slab_cuts(const p3m_params* params, int nranks, int32_t cuts[9], int32_t* layers_out);
/* host-only: the WORK-BALANCED cuts p3m_set_particles derives from a full particle set (packed xyz, `units` as
 * for p3m_set_particles): particle count per layer for PM-only parameters, estimated pair evaluations of the
 * short-range kernels + particles for P3M.  Deterministic, so every rank computes the same cuts from the same
 * set without communicating.  Needs no device. */
int p3m_balanced_cuts(const p3m_params* params, int nranks, const float* pos, int64_t n, int units,
                      int32_t cuts[9], int32_t* layers_out);
/* out = {rank, nranks, first owned binning layer, one past the last, ghost particles held,
 *        1 if the mesh is slab-decomposed (distributed FFT) / 0 if replicated, first mesh plane of this
 *        rank's FFT slab, number of planes in it} */
int p3m_rank_info(p3m_ctx* ctx, int64_t out[8]);

/* Particle upload.  The host arrays (pageable or pinned) may be reused as soon as the call returns.
 * units = P3M_UNITS_ORIGINAL applies stateToCodeUnits + massToCodeUnits on the
 * device (source/unitConversions.cpp:23-71, include/unitConversions.h:8-50), as the head of run()
 * does (source/pmMethod.cpp:72-73).  vel may be NULL (zeros).  Replaces the constructor's copy of
 * `state`/`masses` into vector<Particle> and PMMethodGPU::copyParticlesHostToDevice
 * (source/PMMethodGPU.cu:241-247). */
int p3m_set_particles(p3m_ctx* ctx, const float* pos, const float* vel, const float* mass,
                      int64_t n, int units);
/* ---- N3: initial conditions generated ON THE DEVICE ---------------------------------
 * The DISTRIBUTIONS of the reference's samplers -- PlummerSampler::sample (source/plummerSampler.cpp:11-83),
 * DiskSamplerLinear::sample (source/diskSamplerLinear.cpp:10-74), a uniform cube (BASELINE configs[2]) and the
 * clustered disk + halo mix of configs[3] -- from a COUNTER-BASED generator: particle i of a set is a pure
 * function of (seed, i), so every rank can produce exactly its own share and any index range can be re-created
 * anywhere.  (The reference's own random streams are implementation-defined, SURVEY Q11: the distributions are
 * reproduced, not the streams.)  Quantities are in ORIGINAL units, as the samplers' callers pass them. */
enum { P3M_IC_PLUMMER = 0, P3M_IC_DISK_LINEAR = 1, P3M_IC_UNIFORM = 2, P3M_IC_DISK_HALO = 3 };
typedef struct p3m_ic {
  int32_t kind;
  int32_t truncate;      /* Plummer radii beyond r_max: 0 = moved onto the r_max shell as the reference does
                            (source/plummerSampler.cpp:50-52), 1 = the CDF is truncated at r_max (no shell) */
  uint64_t seed;
  int64_t n;             /* particles of the whole set (all ranks together) */
  float center[3];
  float total_mass;      /* every particle gets total_mass / n (the reference's callers: masses(n, M / n)) */
  float G;
  float a, r_max;        /* PLUMMER: scale radius, cut radius; DISK_HALO: the halo's */
  float rb, mb, rd, md, thickness, r0; /* DISK_LINEAR: bulge radius / mass, disk radius / mass, thickness, inner
                                          radius; DISK_HALO: rd, thickness of its disk (in the x-z plane) */
  float lo[3], hi[3];    /* UNIFORM: positions in [lo, hi) */
  float vel_sigma;       /* UNIFORM: isotropic Gaussian velocity dispersion (0 = at rest) */
} p3m_ic;
/* Fills the context with the set: replaces p3m_set_particles.  With several ranks the call is collective-free:
 * every rank evaluates the per-layer work weights of the whole set on its own device (identical results), derives
 * the same work-balanced cuts, and keeps the particles of its own z-slab -- no host array, no full upload. */
int p3m_generate_particles(p3m_ctx* ctx, const p3m_ic* ic);
/* Particles [first, first + count) of the set as packed host arrays (any pointer may be NULL); needs a CUDA
 * device but no context.  Used by the tests and to feed the same set to the CPU reference. */
int p3m_sample_particles(const p3m_ic* ic, int64_t first, int64_t count, float* pos, float* vel, float* mass);

/* Download in original particle order; any pointer may be NULL.  acc is always in code units when
 * units == P3M_UNITS_CODE, else converted with accelerationToOriginalUnits.  Replaces
 * getParticles() / copyParticlesDeviceToHost (source/PMMethodGPU.cu:244-247). */
int p3m_get_particles(p3m_ctx* ctx, float* pos, float* vel, float* acc, int units);
int p3m_get_particles_f64(p3m_ctx* ctx, double* pos, double* vel, double* acc, int units);
int64_t p3m_num_particles(const p3m_ctx* ctx);

/* PMMethod::initGreensFunction (source/pmMethod.cpp:164-185) + GreenOptimal /
 * GreenDiscreteLaplacian / GreenPoorMan (source/greensFunctions.cpp:122-220), evaluated on the
 * device; stored Hermitian-symmetrised with the inverse FFT's 1/M folded in. */
int p3m_green_init(p3m_ctx* ctx);
/* Optional: install a caller-computed table instead (M values, the real part the reference stores
 * in Grid::greensFunction).  Used by the parity tests to isolate the FFT from the table. */
int p3m_set_green_table(p3m_ctx* ctx, const float* green_full_mesh);
int p3m_set_green_table_f64(p3m_ctx* ctx, const double* green_full_mesh);
/* Symmetrised table expanded back to the full mesh, without the 1/M factor. */
int p3m_get_green_table(p3m_ctx* ctx, double* green_full_mesh);

/* ---- the phases of one force evaluation, individually callable (each is parity-tested) ------- */
/* A0: mesh cell + chaining-mesh cell of every particle (source/pmMethod.cpp:250-252,
 * source/chainingMesh.cpp:25-29) and the cell sort that replaces ChainingMesh::fill /
 * fillWithYSorting (source/chainingMesh.cpp:20-58). */
int p3m_bin_sort(p3m_ctx* ctx);
/* A1: PMMethod::spreadMass (source/pmMethod.cpp:200-277) */
int p3m_deposit(p3m_ctx* ctx);
/* A2+A3: Grid::fftDensity, PMMethod::findFourierPotential, Grid::invFftPotential
 * (source/grid.cpp:50-56, source/pmMethod.cpp:340-350) */
int p3m_poisson(p3m_ctx* ctx);
/* A5: PMMethod::findFieldInCells (source/pmMethod.cpp:373-382) into an explicit field mesh.  Only
 * needed by callers that read Grid::getField; p3m_gather computes the differences on the fly. */
int p3m_gradient(p3m_ctx* ctx);
/* A5+A6: PMMethod::updateAccelerations (source/pmMethod.cpp:384-390): finite differences of the
 * potential fused with interpolation to the particles, plus the external field. */
int p3m_gather(p3m_ctx* ctx);
/* A7-A9: P3MMethod::calculateShortRangeForces + correctAccelerations
 * (source/p3mMethod.cpp:50-57,168-192,247-322).  No-op (P3M_OK) for a PM-only context. */
int p3m_short_range(p3m_ctx* ctx);
/* PMMethod::pmMethodStep (source/pmMethod.cpp:137-144) [+ the two P3M calls above]:
 * bin_sort, deposit, poisson, gather, short_range. */
int p3m_force(p3m_ctx* ctx);

/* A10: leapfrog free functions (source/leapfrog.cpp:5-24), code units.
 * Accelerations belong to the particle order they were computed in: p3m_bin_sort (and a multi-GPU migration)
 * permutes positions, velocities and ids only, so p3m_kick, p3m_diagnostics and every acceleration readback
 * return P3M_ESTATE between a re-sort and the next p3m_gather.  p3m_short_range on its own (no p3m_gather
 * since the sort) starts from zero accelerations. */
int p3m_kick(p3m_ctx* ctx, float dt_factor);   /* v += dt_factor * a  (0.5 = setHalfStepVelocities) */
int p3m_drift(p3m_ctx* ctx);                    /* x += v (+ unit round trip, + escape check) */
/* A11: `steps` iterations of the run-loop body without recording (source/pmMethod.cpp:87-121,
 * source/p3mMethod.cpp:101-153): drift, [escape check], force, kick.  Stops early when a particle
 * escaped the computational box (PMMethod::escapedComputationalBox, source/pmMethod.cpp:146-156);
 * *steps_done (may be NULL) receives the number of completed steps. */
int p3m_step(p3m_ctx* ctx, int steps, int* steps_done);
/* PMMethod::escapedComputationalBox on the current positions. */
int p3m_escaped(p3m_ctx* ctx, int* escaped);

/* Slow path for an arbitrary host `externalField` std::function (source/pmMethod.cpp:384-390): the
 * caller evaluates it on downloaded positions and adds the result (3N values, original particle
 * order; units as given) to the accelerations of the last force evaluation. */
int p3m_add_acceleration(p3m_ctx* ctx, const float* acc_xyz, int units);

/* FFTAdapter<float>::fft / ifft (include/FFTAdapter.h:11-14) on host buffers of nz*ny*nx interleaved
 * complex floats, x fastest: forward unnormalised, inverse divided by the length -- the contract of
 * test/fftAdaptersTest.cpp:6-22.  Needs no context; backs the CuFFTAdapter host class. */
int p3m_fft3d_c2c(int nz, int ny, int nx, const float* in, float* out, int inverse);

/* N1: SimInfo diagnostics on the device (source/simInfo.cpp:50-127, source/pmMethod.cpp:97-105):
 * out[0]=PE out[1]=KE out[2..4]=momentum out[5..7]=angular momentum out[8..10]=total external force
 * (original units; momentum uses the integer-step velocity v + a/2). */
int p3m_diagnostics(p3m_ctx* ctx, double out[11]);

/* ---- readbacks for tests / Grid back-fill (PMMethodGPU::copyGridDensityToHost etc.,
 * source/PMMethodGPU.cu:117-134) ------------------------------------------------------------- */
int p3m_get_density(p3m_ctx* ctx, float* mesh);      /* M values, code units */
int p3m_get_potential(p3m_ctx* ctx, float* mesh);    /* M values */
int p3m_get_field(p3m_ctx* ctx, float* mesh_xyz);    /* 3M values (after p3m_gradient) */
int p3m_get_density_f64(p3m_ctx* ctx, double* mesh);
int p3m_get_potential_f64(p3m_ctx* ctx, double* mesh);
int p3m_set_density(p3m_ctx* ctx, const float* mesh); /* test hook: feed the Poisson solve */
int p3m_set_potential(p3m_ctx* ctx, const float* mesh);
/* per particle (original order): PM mesh cell x + y*Nx + z*Nx*Ny of the truncated position, the
 * chaining-mesh cell (or -1 for PM-only), and `order` = particle ids in sorted order. */
int p3m_get_cells(p3m_ctx* ctx, int32_t* mesh_cell, int32_t* chain_cell, int32_t* order);
int p3m_get_chaining_dims(p3m_ctx* ctx, int32_t dims[3]);
/* sort geometry after p3m_bin_sort: out = {binning cells x, y, z, Morton bits per axis, sub-cell bits
 * per axis, particle-id bits, tile block shift, p3m}.  The sort order is ascending in
 * (Morton(binning cell), Morton(sub-cell), particle id). */
int p3m_get_binning(p3m_ctx* ctx, int32_t out[8]);
/* ChainingMesh::getNeighborsAndSelf (source/chainingMesh.cpp:60-84): host-side geometry helper */
int p3m_chaining_neighbors(const int32_t dims[3], int32_t cell, int32_t out14[14]);
/* mesh part of the acceleration and short-range acceleration (= total SR force / mass), code units */
int p3m_get_acc_parts(p3m_ctx* ctx, double* acc_pm, double* acc_sr);
int p3m_get_sr_table(p3m_ctx* ctx, double* table500);
/* rows of a SAMPLE of particles, by global id (ids ascending): code-unit position, total acceleration and its
 * short-range part, 3 doubles each (any output may be NULL).  Rows of particles this rank does not hold are
 * zero, so the sum over ranks is the full answer.  For parity figures at sizes where id-indexed readbacks of
 * the whole set are not wanted (bench.py). */
int p3m_get_sample(p3m_ctx* ctx, const int32_t* ids_ascending, int64_t m, double* pos, double* acc, double* acc_sr);

/* ---- N4: brute-force accuracy oracle on the device ------------------------------
 * Double-precision sums over ALL particles this rank holds, for `m` target points given in code units -- no
 * chaining mesh, no sort order, no culling, no table replication: an independent evaluation of what the
 * short-range kernels compute, and the O(N^2) reference the thesis' accuracy experiments compare P3M with
 * (source/ppMethod.cpp:88-125 `ppMethodLeapfrog`, source/demos.cpp:593-727, script/p3m_accuracy.py:32-36).
 *   P3M_SUM_SHORT_RANGE  a_i = sum_j m_j f(r_ij) r_ij over r_ij < cutoff, f = the context's short-range law
 *                        (tabulated, source/p3mMethod.cpp:240-245, or analytic, :220-238)
 *   P3M_SUM_NEWTON       a_i = -G_c sum_j m_j r_ij / (r_ij^2 + eps^2)^(3/2), G_c = 1/(4 pi): softened Newtonian
 *                        gravity in code units (source/ppMethod.cpp:100-113 with G -> G_c)
 *   P3M_SUM_CUTOFF_SHELL acc[3i] = number of sources whose r_ij^2 lies within a relative `softening_code` (default
 *                        2e-6) of cutoff^2.  The short-range law is discontinuous at the cutoff (it is simply
 *                        truncated, source/p3mMethod.cpp:258), so whether such a source counts is decided by the
 *                        rounding of r^2 -- in the reference's fp32 arithmetic as much as here; parity figures
 *                        are quoted over the targets without one (bench.py)
 * Pairs at zero distance are skipped (a target that is one of the particles does not attract itself).
 * Multi-GPU: every rank returns the contribution of its own particles; the caller adds the ranks up. */
enum { P3M_SUM_SHORT_RANGE = 0, P3M_SUM_NEWTON = 1, P3M_SUM_CUTOFF_SHELL = 2 };
int p3m_direct_sum(p3m_ctx* ctx, int mode, const double* target_pos_code, int64_t m, double softening_code,
                   double* acc_code);

/* ---- measurement -------------------------------------------------------------------------- */
#define P3M_NPHASE 10
/* accumulated CUDA-event milliseconds per phase since the last reset (timing must be enabled):
 * 0 binSort 1 spreadMass 2 forwardFFT 3 fourierPotential 4 inverseFFT 5 fieldInCells+
 * updateAccelerations (fused gather) 6 shortRangeForcesCalc 7 integrate 8 fieldInCells (explicit
 * gradient) 9 comm -- names follow the reference's measureTime accumulators
 * (source/pmMethod.cpp:21-28, source/p3mMethod.cpp:14-17). */
int p3m_get_phase_ms(p3m_ctx* ctx, float ms[P3M_NPHASE], int reset);
const char* p3m_phase_name(int i);
/* exact pair statistics of the current (sorted) particle set: pairs examined by the short-range kernels and
 * pairs within the cutoff.  Runs a counting instantiation of the same kernels that only READS the particle
 * state (accelerations are not touched) and involves no collective; multi-GPU: this rank's targets only.
 * Needs sorted particles (P3M_ESTATE otherwise): call it after p3m_short_range / p3m_force / p3m_step. */
int p3m_get_pair_counts(p3m_ctx* ctx, uint64_t* checked, uint64_t* in_range);
/* which code paths are active and what the last force evaluation exchanged (this rank's figures):
 *  [0] fused z pass (k_poisson_z) in use   [1] slab-decomposed mesh   [2] equal-mass table in use
 *  [3] packed-FP32 dense-cell kernel in use [4] incremental re-sort enabled
 *  [5] particles this rank sent away in the last migration   [6] ghost particles held
 *  [7] bytes sent in the two all-to-all transposes of the last solve   [8] density-plane bytes sent
 *  [9] potential-plane bytes sent   [10] migration bytes sent   [11] ghost bytes sent
 *  [12] particles that changed cell in the last re-sort (-1: full sort)   [13] full sorts so far
 *  [14] incremental re-sorts so far   [15] reserved */
#define P3M_NSTAT 16
int p3m_get_stats(p3m_ctx* ctx, double out[P3M_NSTAT]);
/* kernels launched by this context so far (bench.py's gpu_launches) */
int64_t p3m_launch_count(const p3m_ctx* ctx);
/* the context's stream as a cudaStream_t, for event timing by the caller */
void* p3m_stream(p3m_ctx* ctx);
int p3m_synchronize(p3m_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* P3M_B200_H */
