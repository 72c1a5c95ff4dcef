// pm_p3m_methods.cpp -- PMMethod / P3MMethod of the reference's API over the C ABI (include/p3m_b200.h).
//
// Replaces the bodies of source/pmMethod.cpp:30-162 and source/p3mMethod.cpp:19-166.  The run loops keep
// the reference's sequencing (unit change, Green table, first force, half kick, then per step: drift,
// diagnostics, record, escape test, force, kick) but every phase is one C-ABI call on device-resident
// particles; host vectors are only filled for recording and for getParticles().
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <stdexcept>

#include "../../../include/p3m_b200.h"
#include "../include/particle_simulation_b200.hpp"

namespace {

void check(int rc, const char* what) {
  if (rc == P3M_OK) return;
  std::string msg = std::string(what) + ": " + p3m_last_error();
  // unknown enums are std::invalid_argument in the reference (source/pmMethod.cpp:180,275,336,367)
  if (rc == P3M_EINVAL) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

}  // namespace

struct PMMethod::Impl {
  p3m_params prm{};
  p3m_ctx* ctx = nullptr;
  // Host mirror of the reference's `particles` member.  Like the reference's vector its units are the CALLER's
  // business: the constructor fills it in original units, run() converts it, pmMethodStep() takes it as code
  // units (source/pmMethod.cpp:137-144 runs on whatever the vector holds; P3MMethod::run converts it through
  // getParticles() first, source/p3mMethod.cpp:73-74).  Whenever the device state is copied back the mirror is
  // in code units (positions, velocities, masses), which is what the reference's vector holds after the same
  // calls; `hostUnitsCode` only records that so that download() and run() know.
  std::vector<Particle> particles;
  std::vector<float> massOriginal;  // masses in original units (constructor / head of run())
  bool hostUnitsCode = false;
  bool hostIsNewer = true;          // host vector must be uploaded before the next device call
  bool deviceIsNewer = false;       // device state must be downloaded before host code reads the vector
  Grid* grid = nullptr;
  std::unique_ptr<CuFFTAdapter> ownedFft;
  std::unique_ptr<Grid> ownedGrid;
  bool extIsZero = true;     // no field at all: empty std::function, or ExternalFieldDesc::none() installed
  bool extOnDevice = false;  // an ExternalFieldDesc was installed: the device evaluates it, the callable is unused
  std::vector<std::complex<float>> gridDensity, gridPotential;  // PMMethodGPU::getGridDensity / getGridPotential
  std::vector<float> scratch, scratch2;

  ~Impl() {
    if (ctx) p3m_destroy(ctx);
  }

  void ensureCtx() {
    if (ctx) return;
    check(p3m_create(&prm, &ctx), "p3m_create");
  }
  void dropCtx() {
    if (ctx) p3m_destroy(ctx);
    ctx = nullptr;
    hostIsNewer = true;
  }
  // `units`: how the operation that needs the particles on the device interprets the mirror
  void upload(int units) {
    ensureCtx();
    if (!hostIsNewer) return;
    const size_t n = particles.size();
    scratch.resize(7 * n);
    float *pos = scratch.data(), *vel = pos + 3 * n, *mass = vel + 3 * n;
    for (size_t i = 0; i < n; ++i) {
      const Particle& p = particles[i];
      pos[3 * i] = p.position.x, pos[3 * i + 1] = p.position.y, pos[3 * i + 2] = p.position.z;
      vel[3 * i] = p.velocity.x, vel[3 * i + 1] = p.velocity.y, vel[3 * i + 2] = p.velocity.z;
      mass[i] = p.mass;  // same units as the positions: the mirror is converted as a whole
    }
    check(p3m_set_particles(ctx, pos, vel, mass, (int64_t)n, units), "p3m_set_particles");
    hostIsNewer = false;
    deviceIsNewer = false;
    hostUnitsCode = true;  // from here on the device state (code units) is what download() mirrors
    if (units == P3M_UNITS_CODE)
      for (size_t i = 0; i < n; ++i) massOriginal[i] = massToOriginalUnits(particles[i].mass, prm.H, prm.DT, prm.G);
    else
      for (size_t i = 0; i < n; ++i) massOriginal[i] = particles[i].mass;
  }
  // host vector <- device, in the units the reference's vector would be in at this point
  void download() {
    if (!ctx || !deviceIsNewer) return;
    const size_t n = particles.size();
    scratch.resize(9 * n);
    float *pos = scratch.data(), *vel = pos + 3 * n, *acc = vel + 3 * n;
    check(p3m_get_particles(ctx, pos, vel, acc, P3M_UNITS_CODE), "p3m_get_particles");
    const float H = prm.H, DT = prm.DT;
    for (size_t i = 0; i < n; ++i) {
      Particle& p = particles[i];
      p.position = Vec3{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
      p.velocity = Vec3{vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]};
      p.acceleration = Vec3{acc[3 * i], acc[3 * i + 1], acc[3 * i + 2]};  // always code units in the reference
      p.integerStepVelocity = p.velocity + 0.5f * p.acceleration;         // source/leapfrog.cpp:10-14, code units
      p.mass = massToCodeUnits(massOriginal[i], H, DT, prm.G);
    }
    (void)H, (void)DT;
    deviceIsNewer = false;
  }
  void hostCallbackField(const std::function<Vec3(Vec3)>& field) {
    // slow path for arbitrary std::function fields (source/pmMethod.cpp:384-390)
    const size_t n = particles.size();
    scratch.resize(3 * n);
    scratch2.resize(3 * n);
    check(p3m_get_particles(ctx, scratch.data(), nullptr, nullptr, P3M_UNITS_ORIGINAL), "p3m_get_particles");
    for (size_t i = 0; i < n; ++i) {
      Vec3 g = field(Vec3{scratch[3 * i], scratch[3 * i + 1], scratch[3 * i + 2]});
      scratch2[3 * i] = g.x, scratch2[3 * i + 1] = g.y, scratch2[3 * i + 2] = g.z;
    }
    check(p3m_add_acceleration(ctx, scratch2.data(), P3M_UNITS_ORIGINAL), "p3m_add_acceleration");
  }
};

static void fillParams(p3m_params& prm, std::tuple<int, int, int> gridPoints,
                       std::tuple<float, float, float> box, float H, float DT, float G, InterpolationScheme is,
                       FiniteDiffScheme fds, GreensFunction gFunc, float particleDiameter) {
  p3m_default_params(&prm);
  prm.nx = std::get<0>(gridPoints), prm.ny = std::get<1>(gridPoints), prm.nz = std::get<2>(gridPoints);
  prm.box[0] = std::get<0>(box), prm.box[1] = std::get<1>(box), prm.box[2] = std::get<2>(box);
  prm.H = H, prm.DT = DT, prm.G = G;
  prm.assignment = (int)is, prm.fd_scheme = (int)fds, prm.greens_function = (int)gFunc;
  prm.particle_diameter = particleDiameter;
  prm.p3m = 0;
}


PMMethod::PMMethod(const std::vector<Vec3>& state, const std::vector<float>& masses,
                   const std::tuple<float, float, float> effectiveBoxSize,
                   const std::function<Vec3(Vec3)> externalField,
                   const std::function<float(Vec3)> externalPotential, const float H, const float DT,
                   const float G, const InterpolationScheme is, const FiniteDiffScheme fds,
                   const GreensFunction gFunc, const float particleDiameter, Grid& grid)
    : impl(new Impl), H(H), DT(DT), G(G), externalField(externalField), externalPotential(externalPotential) {
  fillParams(impl->prm, grid.getGridPoints(), effectiveBoxSize, H, DT, G, is, fds, gFunc, particleDiameter);
  impl->grid = &grid;
  const size_t n = masses.size();
  impl->particles.reserve(n);
  for (size_t i = 0; i < n; ++i) impl->particles.emplace_back(state[i], state[n + i], masses[i]);
  impl->massOriginal = masses;
  // a callable cannot be proven zero by sampling it: only an EMPTY std::function means "no field"
  impl->extIsZero = !externalField;
}

PMMethod::PMMethod(const std::vector<Vec3>& state, const std::vector<float>& masses,
                   const std::tuple<float, float, float> effectiveBoxSize,
                   const std::function<Vec3(Vec3)> externalField,
                   const std::function<float(Vec3)> externalPotential, const float H, const float DT,
                   const float G, const InterpolationScheme is, const FiniteDiffScheme fds,
                   const GreensFunction gFunc, const float particleDiameter,
                   std::tuple<int, int, int> gridPoints)
    : impl(new Impl), H(H), DT(DT), G(G), externalField(externalField), externalPotential(externalPotential) {
  fillParams(impl->prm, gridPoints, effectiveBoxSize, H, DT, G, is, fds, gFunc, particleDiameter);
  impl->ownedFft.reset(new CuFFTAdapter(std::array<int, 3>{std::get<2>(gridPoints), std::get<1>(gridPoints),
                                                           std::get<0>(gridPoints)}));
  impl->ownedGrid.reset(new Grid(gridPoints, *impl->ownedFft));
  impl->grid = impl->ownedGrid.get();
  const size_t n = masses.size();
  impl->particles.reserve(n);
  for (size_t i = 0; i < n; ++i) impl->particles.emplace_back(state[i], state[n + i], masses[i]);
  impl->massOriginal = masses;
  impl->extIsZero = !externalField;
}

PMMethod::~PMMethod() = default;

void PMMethod::setExternalFieldDescriptor(const ExternalFieldDesc& d) {
  impl->prm.ext_kind = (int)d.kind;
  impl->prm.ext_center[0] = d.center.x, impl->prm.ext_center[1] = d.center.y, impl->prm.ext_center[2] = d.center.z;
  impl->prm.ext_R = d.R, impl->prm.ext_M = d.M;
  impl->extOnDevice = true;  // the descriptor replaces the callable (none(): explicitly no field)
  impl->extIsZero = d.kind == ExternalFieldDesc::NONE;
  if (impl->ctx) {
    impl->download();
    impl->dropCtx();
  }
}

void PMMethod::setPrecision(bool fp64) {
  impl->prm.precision = fp64 ? P3M_F64 : P3M_F32;
  if (impl->ctx) {
    impl->download();
    impl->dropCtx();
  }
}

p3m_ctx* PMMethod::context() {
  impl->ensureCtx();
  return impl->ctx;
}

std::vector<Particle>& PMMethod::getParticles() {
  impl->download();
  impl->hostIsNewer = true;  // the caller may edit the vector (P3MMethod does in the reference)
  return impl->particles;
}

void PMMethod::copyParticlesDeviceToHost() { impl->download(); }
// include_gpu/PMMethodGPU.h:46-47; the reference calls it with the vector already in code units
// (source/p3mMethod.cpp:77-79,138-140)
void PMMethod::copyParticlesHostToDevice() {
  impl->hostIsNewer = true;
  impl->upload(P3M_UNITS_CODE);
}

void PMMethod::initGreensFunction() {
  impl->ensureCtx();
  check(p3m_green_init(impl->ctx), "p3m_green_init");
}

// source/pmMethod.cpp:137-144.  Like the reference it works on the particles as they are (code units
// once run() has converted them); after the call the accelerations live on the device.
void PMMethod::pmMethodStep() {
  impl->upload(P3M_UNITS_CODE);
  check(p3m_bin_sort(impl->ctx), "p3m_bin_sort");
  check(p3m_deposit(impl->ctx), "p3m_deposit");
  check(p3m_poisson(impl->ctx), "p3m_poisson");
  check(p3m_gather(impl->ctx), "p3m_gather");
  if (!impl->extIsZero && !impl->extOnDevice) impl->hostCallbackField(externalField);
  impl->deviceIsNewer = true;
}

// source/pmMethod.cpp:146-156: the positions AS THEY ARE in the vector against the box (the run loops call it
// right after stateToOriginalUnits)
bool PMMethod::escapedComputationalBox() {
  impl->download();
  const float bx = impl->prm.box[0], by = impl->prm.box[1], bz = impl->prm.box[2];
  for (const Particle& p : impl->particles) {
    const Vec3& x = p.position;
    if (!(x.x >= 0 && x.x <= bx && x.y >= 0 && x.y <= by && x.z >= 0 && x.z <= bz)) return true;
  }
  return false;
}

// source/pmMethod.cpp:158-162: sum of p.mass * externalField(p.position) over the vector AS IT IS (the run loops
// call it with positions and masses back in original units)
Vec3 PMMethod::totalExternalForceOrigUnits() {
  if (impl->extIsZero) return Vec3::zero();
  impl->download();
  Vec3 total = Vec3::zero();
  const p3m_params& q = impl->prm;
  const Vec3 c = Vec3::create(q.ext_center[0], q.ext_center[1], q.ext_center[2]);
  for (const Particle& p : impl->particles)
    total += p.mass * (externalField ? externalField(p.position) : sphRadDecrField(p.position, c, q.ext_R, q.ext_M, G));
  return total;
}

void PMMethod::copyGridDensityToHost() {
  impl->ensureCtx();
  const int M = impl->grid->getLength();
  impl->scratch2.resize(M);
  check(p3m_get_density(impl->ctx, impl->scratch2.data()), "p3m_get_density");
  auto& d = impl->grid->densityStorage();
  impl->gridDensity.resize(M);
  for (int i = 0; i < M; ++i) impl->gridDensity[i] = d[i] = std::complex<float>(impl->scratch2[i], 0.0f);
}

void PMMethod::copyGridPotentialToHost() {
  impl->ensureCtx();
  const int M = impl->grid->getLength();
  impl->scratch2.resize(M);
  check(p3m_get_potential(impl->ctx, impl->scratch2.data()), "p3m_get_potential");
  auto& d = impl->grid->potentialStorage();
  impl->gridPotential.resize(M);
  for (int i = 0; i < M; ++i) impl->gridPotential[i] = d[i] = std::complex<float>(impl->scratch2[i], 0.0f);
}

// include/pmMethod.h:37 is const; the meshes live on the device, so the host Grid is refreshed first
// (impl is a pointer: the refresh does not touch this object's own members)
const Grid& PMMethod::getGrid() const {
  if (impl->ctx) {
    PMMethod* self = const_cast<PMMethod*>(this);
    self->copyGridDensityToHost();
    self->copyGridPotentialToHost();
  }
  return *impl->grid;
}
// include_gpu/PMMethodGPU.h:38-39: the vectors as of the last copyGrid*ToHost()
const std::vector<std::complex<float>>& PMMethod::getGridDensity() const { return impl->gridDensity; }
const std::vector<std::complex<float>>& PMMethod::getGridPotential() const { return impl->gridPotential; }

// The shared body of PMMethod::run (source/pmMethod.cpp:62-135) and P3MMethod::run
// (source/p3mMethod.cpp:59-166).
void PMMethod::runLoop(StateRecorder& rec, int simLength, bool diagnostics, bool recordField, bool p3m) {
  Impl& s = *impl;
  const size_t n = s.particles.size();
  SimInfo simInfo;
  if (s.hostUnitsCode)
    throw std::runtime_error("run(): the particles are in code units already (run() converts them itself, "
                             "source/pmMethod.cpp:72-73)");
  if (diagnostics) simInfo.setInitialMomentum(s.particles);  // sum m * v, original units
  if (s.prm.p3m != (p3m ? 1 : 0)) throw std::logic_error("run(): context mode mismatch");
  s.hostIsNewer = true;
  s.upload(P3M_UNITS_ORIGINAL);  // stateToCodeUnits + massToCodeUnits happen on the device
  const bool callback = !s.extIsZero && !s.extOnDevice;
  auto force = [&]() {
    check(p3m_force(s.ctx), "p3m_force");  // pmMethodStep [+ short range + correctAccelerations]
    if (callback) s.hostCallbackField(externalField);
  };
  check(p3m_green_init(s.ctx), "p3m_green_init");
  force();
  check(p3m_kick(s.ctx, 0.5f), "p3m_kick");  // setHalfStepVelocities
  // only StateRecorder members the reference's class has too (include/stateRecorder.h:26-36): in
  // reference-tree mode `rec` IS the reference's recorder
  static_assert(sizeof(Vec3) == 3 * sizeof(float), "Vec3 is three packed floats");
  std::vector<Vec3> posbuf(n);
  float* buf = reinterpret_cast<float*>(posbuf.data());
  for (int t = 0; t <= simLength; ++t) {
    std::cout << "progress: " << float(t) / simLength << '\r';
    std::cout.flush();
    if (recordField) {
      s.deviceIsNewer = true;
      s.download();  // accelerations in code units; recordField converts (source/stateRecorder.cpp:102-108)
      rec.recordField(s.particles, H, DT);
    }
    check(p3m_drift(s.ctx), "p3m_drift");  // updatePositions (+ the unit round trip, + escape flag)
    check(p3m_get_particles(s.ctx, buf, nullptr, nullptr, P3M_UNITS_ORIGINAL), "p3m_get_particles");
    rec.recordPositions(posbuf.begin(), posbuf.end());
    if (diagnostics) {
      double d[11];
      check(p3m_diagnostics(s.ctx, d), "p3m_diagnostics");
      Vec3 ext{(float)d[8], (float)d[9], (float)d[10]};
      double pe = d[0];
      if (callback) {
        ext = Vec3::zero();
        for (size_t i = 0; i < n; ++i) {
          Vec3 pos{buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]};
          ext += s.massOriginal[i] * externalField(pos);
          if (externalPotential) pe += s.massOriginal[i] * externalPotential(pos);
        }
      }
      rec.recordExpectedMomentum(simInfo.updateExpectedMomentum(ext, DT));
      rec.recordTotalMomentum(Vec3{(float)d[2], (float)d[3], (float)d[4]});
      rec.recordTotalAngularMomentum(Vec3{(float)d[5], (float)d[6], (float)d[7]});
      rec.recordEnergy((float)pe, (float)d[1]);
    }
    int escaped = 0;
    check(p3m_escaped(s.ctx, &escaped), "p3m_escaped");
    if (escaped) {
      std::cout << "Particle moved outside the computational box.\n";
      break;
    }
    force();
    check(p3m_kick(s.ctx, 1.0f), "p3m_kick");  // updateVelocities
  }
  float ms[P3M_NPHASE];
  if (s.prm.timing && p3m_get_phase_ms(s.ctx, ms, 0) == P3M_OK)
    for (int i = 0; i < P3M_NPHASE; ++i) std::cout << p3m_phase_name(i) << ": " << ms[i] << " ms\n";
  s.deviceIsNewer = true;
}

std::string PMMethod::run(StateRecorder& stateRecorder, const int simLength, bool collectDiagnostics,
                          bool recordField) {
  if (impl->prm.p3m) {  // a P3MMethod was attached earlier; PMMethod::run is mesh-only
    impl->download();
    impl->prm.p3m = 0;
    impl->dropCtx();
  }
  runLoop(stateRecorder, simLength, collectDiagnostics, recordField, false);
  return stateRecorder.flush();
}

// ---- P3MMethod ------------------------------------------------------------------------------------------
P3MMethod::P3MMethod(PMMethod& pm, std::tuple<float, float, float> compBoxSize, float cutoffRadius,
                     float particleDiameter, float H, float softeningLength, CloudShape cloudShape,
                     bool useSRForceTable, bool /*enableYSorting*/)
    : pmMethod(pm) {
  // enableYSorting only orders the reference's linked lists (an early-exit optimisation of its CPU
  // loop, source/p3mMethod.cpp:309-313); the result is the same sum, so it has no effect here.
  p3m_params& prm = pm.impl->prm;
  if (pm.impl->ctx) {
    pm.impl->download();
    pm.impl->dropCtx();
  }
  prm.p3m = 1;
  prm.box[0] = std::get<0>(compBoxSize), prm.box[1] = std::get<1>(compBoxSize), prm.box[2] = std::get<2>(compBoxSize);
  prm.cutoff_radius = cutoffRadius;
  prm.softening = softeningLength;
  prm.cloud_shape = (int)cloudShape;
  prm.use_sr_table = useSRForceTable ? 1 : 0;
  // the reference keeps its own particleDiameter for the short-range reference force (source/p3mMethod.cpp:36)
  // next to PMMethod's for the influence function (source/pmMethod.cpp:56): honoured when they differ
  prm.sr_particle_diameter = particleDiameter == prm.particle_diameter ? 0.0f : particleDiameter;
  // cutoff, diameter, softening and the chaining mesh are converted with the H passed here
  // (source/p3mMethod.cpp:35-37, source/chainingMesh.cpp:13-15) while the particles are in PMMethod's code
  // units: a different H is an inconsistent set-up in the reference too -- refused instead of ignored
  if (H != pm.getH())
    throw std::invalid_argument("P3MMethod: H differs from the H of the PMMethod it wraps");
}

void P3MMethod::forceStep() {
  PMMethod::Impl& s = *pmMethod.impl;
  s.upload(P3M_UNITS_CODE);  // like pmMethodStep(): the vector is taken as code units
  check(p3m_force(s.ctx), "p3m_force");
  if (!s.extIsZero && !s.extOnDevice) s.hostCallbackField(pmMethod.externalField);
  s.deviceIsNewer = true;
}

void P3MMethod::run(StateRecorder& stateRecorder, const int simLength, bool collectDiagnostics, bool recordField) {
  pmMethod.impl->prm.p3m = 1;
  pmMethod.runLoop(stateRecorder, simLength, collectDiagnostics, recordField, true);
  stateRecorder.flush();
}
