// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// A C-ABI driver around the UNMODIFIED reference CPU sources, which are compiled where they lie
// under /root/reference by oracle/Makefile (nothing from the reference is copied into this repo).
// It plays the role of source/demos.cpp (which does not compile with g++):
// it builds KissFFTAdapter -> Grid -> PMMethod -> P3MMethod exactly as the demos do
// (/root/reference/source/demos.cpp:736-776, 897-951, 1416-1447) and dumps every intermediate of
// the P3M force step as raw arrays so the C restatement (oracle/p3m_oracle.c) and the CUDA path can
// be pinned against the real thing.
//
// Output: oracle/_ref/libp3m_ref.so (git-ignored, travels to the GPU box with the snapshot).
#include <array>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <numbers>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

// The reference keeps spreadMass/calculateShortRangeForces/chainingMesh private; this driver (and
// only this translation unit) opens them up so each phase can be observed on its own.  Access
// specifiers do not change the object layout, and the reference's own .cpp files are compiled
// untouched.
#define private public
#include "chainingMesh.h"
#include "p3mMethod.h"
#include "pmMethod.h"
#undef private

#include "diskSampler.h"
#include "diskSamplerLinear.h"
#include "externalFields.h"
#include "kissFFTAdapter.h"
#include "leapfrog.h"
#include "plummerSampler.h"
#include "simInfo.h"
#include "stateRecorder.h"
#include "unitConversions.h"

// measureTime accumulators are file-scope globals in the reference
// (/root/reference/source/pmMethod.cpp:21-28, /root/reference/source/p3mMethod.cpp:14-17).
extern float spreadMassTimeMs_, forwardFFTTimeMs_, fourierPotentialTimeMs_, inverseFFTTimeMs_,
    fieldInCellsTimeMs_, updateAccelerationsTimeMs_, pmTimeMs_;
extern float chainingMeshSetupTimeMs_, shortRangeForcesCalcTimeMs_, pmStepTimeMs_,
    correctAccelerationsTimeMs_;

void correctAccelerations(std::vector<Particle>& particles);  // source/p3mMethod.cpp:50-57

extern "C" {

// Mirrored field-for-field by tests/refapi.py (ctypes).
struct RefParams {
  int n;
  int nx, ny, nz;
  float box[3];  // effectiveBoxSize / compBoxSize, original units
  float H, DT, G;
  int is;     // InterpolationScheme: 0 NGP 1 CIC 2 TSC      (include/pmConfig.h:3)
  int fds;    // FiniteDiffScheme: 0 TWO_POINT 1 FOUR_POINT  (include/pmConfig.h:4)
  int gfunc;  // GreensFunction: 0 DISCRETE_LAPLACIAN 1 S1_OPTIMAL 2 S2_OPTIMAL 3 POOR_MAN
  float particleDiameter;  // original units
  float cutoffRadius;      // original units
  float softening;         // original units
  int cloudShape;          // 0 S1, 1 S2 (include/greensFunctions.h:7)
  int useTable;
  int ySort;
  int extKind;  // 0 none, 1 sphRadDecrField(center, R, M, G)
  float extCenter[3];
  float extR, extM;
  int greenZeroDegenerate;  // OrcParams only; the reference has no such switch (ignored here)
};

}  // extern "C"

namespace {

struct Sim {
  std::array<int, 3> dims;
  KissFFTAdapter<float> fft;
  Grid grid;
  PMMethod pm;

  static std::function<Vec3(Vec3)> field(const RefParams& p) {
    if (p.extKind == 1) {
      Vec3 c = Vec3(p.extCenter[0], p.extCenter[1], p.extCenter[2]);
      float R = p.extR, M = p.extM, G = p.G;
      return [c, R, M, G](Vec3 pos) -> Vec3 { return sphRadDecrField(pos, c, R, M, G); };
    }
    return [](Vec3) -> Vec3 { return Vec3::zero(); };
  }
  static std::function<float(Vec3)> potential(const RefParams& p) {
    if (p.extKind == 1) {
      Vec3 c = Vec3(p.extCenter[0], p.extCenter[1], p.extCenter[2]);
      float R = p.extR, M = p.extM, G = p.G;
      return [c, R, M, G](Vec3 pos) -> float { return sphRadDecrFieldPotential(pos, c, R, M, G); };
    }
    return [](Vec3) -> float { return 0; };
  }
  static std::vector<Vec3> state(const RefParams& p, const float* pos, const float* vel) {
    std::vector<Vec3> s(2 * (size_t)p.n);
    for (int i = 0; i < p.n; ++i) {
      s[i] = Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
      s[p.n + i] = vel ? Vec3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]) : Vec3::zero();
    }
    return s;
  }

  Sim(const RefParams& p, const float* pos, const float* vel, const float* mass)
      : dims{p.nz, p.ny, p.nx},  // adapters get {Nz,Ny,Nx}: source/demos.cpp:758-759
        fft(dims.data(), 3),
        grid(std::make_tuple(p.nx, p.ny, p.nz), fft),
        pm(state(p, pos, vel),
           std::vector<float>(mass, mass + p.n),
           std::make_tuple(p.box[0], p.box[1], p.box[2]),
           field(p),
           potential(p),
           p.H,
           p.DT,
           p.G,
           (InterpolationScheme)p.is,
           (FiniteDiffScheme)p.fds,
           (GreensFunction)p.gfunc,
           p.particleDiameter,
           grid) {}
};

void dumpMesh(const RefParams& p, Sim& s, float* density, float* potential, float* field,
              float* green) {
  const Grid& g = s.pm.getGrid();
  size_t i = 0;
  for (int z = 0; z < p.nz; ++z)
    for (int y = 0; y < p.ny; ++y)
      for (int x = 0; x < p.nx; ++x, ++i) {
        if (density) density[i] = g.getDensity(x, y, z);
        if (potential) potential[i] = g.getPotential(x, y, z);
        if (field) {
          Vec3 f = g.getField(x, y, z);
          field[3 * i] = f.x, field[3 * i + 1] = f.y, field[3 * i + 2] = f.z;
        }
        if (green) {
          auto c = g.getGreensFunction(x, y, z);
          green[2 * i] = c.real(), green[2 * i + 1] = c.imag();
        }
      }
}

void dumpVec(const std::vector<Particle>& ps, Vec3 Particle::*m, float* out) {
  if (!out) return;
  for (size_t i = 0; i < ps.size(); ++i) {
    const Vec3& v = ps[i].*m;
    out[3 * i] = v.x, out[3 * i + 1] = v.y, out[3 * i + 2] = v.z;
  }
}

P3MMethod makeP3M(const RefParams& p, PMMethod& pm) {
  return P3MMethod(pm, std::make_tuple(p.box[0], p.box[1], p.box[2]), p.cutoffRadius,
                   p.particleDiameter, p.H, p.softening, (CloudShape)p.cloudShape, p.useTable != 0,
                   p.ySort != 0);
}

}  // namespace

extern "C" {

// Influence-function table exactly as PMMethod::initGreensFunction fills it
// (source/pmMethod.cpp:164-185); green = M interleaved (re, im).
int ref_green(const RefParams* p, float* green) {
  std::vector<float> pos(3, 1.0f), mass(1, 1.0f);
  RefParams q = *p;
  q.n = 1;
  Sim s(q, pos.data(), nullptr, mass.data());
  s.pm.initGreensFunction();
  dumpMesh(q, s, nullptr, nullptr, nullptr, green);
  return 0;
}

// One PM force evaluation, sequenced as the head of PMMethod::run (source/pmMethod.cpp:72-75).
// All outputs optional.  pos_code (3N) / mass_code (N) = particle state after the unit change,
// density/potential (M), field (3M), acc (3N, code units), green (2M).
int ref_pm_force(const RefParams* p, const float* pos, const float* vel, const float* mass,
                 float* pos_code, float* mass_code, float* density, float* potential, float* field,
                 float* acc, float* green) {
  Sim s(*p, pos, vel, mass);
  auto& ps = s.pm.getParticles();
  stateToCodeUnits(ps, p->H, p->DT);
  massToCodeUnits(ps, p->H, p->DT, p->G);
  s.pm.initGreensFunction();
  s.pm.pmMethodStep();
  dumpMesh(*p, s, density, potential, field, green);
  dumpVec(ps, &Particle::position, pos_code);
  dumpVec(ps, &Particle::acceleration, acc);
  if (mass_code)
    for (int i = 0; i < p->n; ++i) mass_code[i] = ps[i].mass;
  return 0;
}

// One full P3M force evaluation, sequenced as the head of P3MMethod::run
// (source/p3mMethod.cpp:73-86).  acc_pm = mesh part, sr_force = own + sum of the 13 neighbour slots
// (what correctAccelerations adds before dividing by mass), acc = corrected total (code units).
// cell (N) = chaining-mesh cell of each particle, order (N) = particle ids listed cell by cell in
// the linked-list order the reference walks them (source/chainingMesh.cpp:20-58).
// ftable (500) = short-range force table (source/p3mMethod.cpp:275-294), when useTable.
int ref_p3m_force(const RefParams* p, const float* pos, const float* vel, const float* mass,
                  float* acc_pm, float* sr_force, float* acc, int* cell, int* order, int* mesh_dims,
                  float* ftable) {
  Sim s(*p, pos, vel, mass);
  auto& ps = s.pm.getParticles();
  P3MMethod p3m = makeP3M(*p, s.pm);
  stateToCodeUnits(ps, p->H, p->DT);
  massToCodeUnits(ps, p->H, p->DT, p->G);
  s.pm.initGreensFunction();
  s.pm.pmMethodStep();
  dumpVec(ps, &Particle::acceleration, acc_pm);
  p3m.calculateShortRangeForces(ps);
  if (sr_force)
    for (int i = 0; i < p->n; ++i) {
      Vec3 t = ps[i].shortRangeForce;
      for (const auto& v : ps[i].shortRangeFromNeighbor) t += v;
      sr_force[3 * i] = t.x, sr_force[3 * i + 1] = t.y, sr_force[3 * i + 2] = t.z;
    }
  correctAccelerations(ps);
  dumpVec(ps, &Particle::acceleration, acc);
  ChainingMesh& cm = p3m.chainingMesh;
  if (mesh_dims) mesh_dims[0] = cm.Mx, mesh_dims[1] = cm.My, mesh_dims[2] = cm.Mz;
  if (cell || order) {
    int k = 0;
    for (int q = 0; q < cm.getSize(); ++q)
      for (auto* node = cm.getParticlesInCell(q); node != nullptr; node = node->next) {
        if (cell) cell[node->particleId] = q;
        if (order) order[k] = node->particleId;
        ++k;
      }
  }
  if (ftable && p->useTable) std::memcpy(ftable, p3m.FTable.data(), 500 * sizeof(float));
  return 0;
}

// ChainingMesh geometry alone: cell of every particle given CODE-unit positions
// (source/chainingMesh.cpp:6-29), plus getNeighborsAndSelf of one cell (:60-84).
int ref_chaining_cells(const RefParams* p, const float* pos_code, int* cell, int* mesh_dims) {
  ChainingMesh cm(std::make_tuple(p->box[0], p->box[1], p->box[2]), p->cutoffRadius, p->H, p->n);
  mesh_dims[0] = cm.Mx, mesh_dims[1] = cm.My, mesh_dims[2] = cm.Mz;
  for (int i = 0; i < p->n; ++i) {
    int cx = int(pos_code[3 * i] / cm.HCx);
    int cy = int(pos_code[3 * i + 1] / cm.HCy);
    int cz = int(pos_code[3 * i + 2] / cm.HCz);
    cell[i] = cm.tripleToFlatIndex(cx, cy, cz);
  }
  return 0;
}

int ref_chaining_neighbors(const RefParams* p, int cellIdx, int* out14) {
  ChainingMesh cm(std::make_tuple(p->box[0], p->box[1], p->box[2]), p->cutoffRadius, p->H, 1);
  auto nb = cm.getNeighborsAndSelf(cellIdx);
  for (int i = 0; i < 14; ++i) out14[i] = nb[i];
  return 0;
}

// Whole run loop (PMMethod::run source/pmMethod.cpp:62-135 or P3MMethod::run
// source/p3mMethod.cpp:59-166) with diagnostics on; the reference writes energy.txt, momentum.txt,
// expected_momentum.txt, angular_momentum.txt and positions.dat into out_dir.  Final particle state
// is returned in the units the loop leaves it in (code units; see the loop tail).
int ref_run(const RefParams* p, const float* pos, const float* vel, const float* mass,
            int simLength, int p3m, int diagnostics, const char* out_dir, float* pos_out,
            float* vel_out, float* acc_out) {
  std::filesystem::create_directories(out_dir);
  Sim s(*p, pos, vel, mass);
  {
    StateRecorder rec(p->n, simLength + 1, out_dir);
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());  // the loop prints a progress line per step
    if (p3m) {
      P3MMethod m = makeP3M(*p, s.pm);
      m.run(rec, simLength, diagnostics != 0, false);
    } else {
      s.pm.run(rec, simLength, diagnostics != 0, false);
    }
    std::cout.rdbuf(old);
  }
  auto& ps = s.pm.getParticles();
  dumpVec(ps, &Particle::position, pos_out);
  dumpVec(ps, &Particle::velocity, vel_out);
  dumpVec(ps, &Particle::acceleration, acc_out);
  return 0;
}

// CPU baseline timing: `steps` force+integrate steps after one untimed initial force evaluation,
// sequenced like the body of the reference run loops without recording/diagnostics
// (drift -> force -> kick).  ms[0..10] = spreadMass, forwardFFT, fourierPotential, inverseFFT,
// fieldInCells, updateAccelerations, chainingMeshSetup, shortRangeForcesCalc (incl. setup),
// correctAccelerations, integrate, total wall  -- all summed over the timed steps.
// green_init_ms = the once-per-run influence-function setup (not part of a step).
int ref_time_steps(const RefParams* p, const float* pos, const float* vel, const float* mass,
                   int steps, int p3m, float* ms, float* green_init_ms) {
  using clk = std::chrono::steady_clock;
  Sim s(*p, pos, vel, mass);
  auto& ps = s.pm.getParticles();
  stateToCodeUnits(ps, p->H, p->DT);
  massToCodeUnits(ps, p->H, p->DT, p->G);
  auto g0 = clk::now();
  s.pm.initGreensFunction();
  *green_init_ms = std::chrono::duration<float, std::milli>(clk::now() - g0).count();
  std::unique_ptr<P3MMethod> m;
  if (p3m) m = std::make_unique<P3MMethod>(makeP3M(*p, s.pm));
  s.pm.pmMethodStep();
  if (p3m) {
    m->calculateShortRangeForces(ps);
    correctAccelerations(ps);
  }
  setHalfStepVelocities(ps);
  spreadMassTimeMs_ = forwardFFTTimeMs_ = fourierPotentialTimeMs_ = inverseFFTTimeMs_ = 0;
  fieldInCellsTimeMs_ = updateAccelerationsTimeMs_ = chainingMeshSetupTimeMs_ = 0;
  float sr = 0, corr = 0, integ = 0;
  auto t0 = clk::now();
  for (int t = 0; t < steps; ++t) {
    auto a = clk::now();
    updatePositions(ps);
    auto b = clk::now();
    s.pm.pmMethodStep();
    auto c = clk::now();
    if (p3m) {
      m->calculateShortRangeForces(ps);
      auto d = clk::now();
      correctAccelerations(ps);
      auto e = clk::now();
      sr += std::chrono::duration<float, std::milli>(d - c).count();
      corr += std::chrono::duration<float, std::milli>(e - d).count();
    }
    auto f = clk::now();
    updateVelocities(ps);
    auto g = clk::now();
    integ += std::chrono::duration<float, std::milli>((b - a) + (g - f)).count();
  }
  float total = std::chrono::duration<float, std::milli>(clk::now() - t0).count();
  ms[0] = spreadMassTimeMs_, ms[1] = forwardFFTTimeMs_, ms[2] = fourierPotentialTimeMs_;
  ms[3] = inverseFFTTimeMs_, ms[4] = fieldInCellsTimeMs_, ms[5] = updateAccelerationsTimeMs_;
  ms[6] = chainingMeshSetupTimeMs_, ms[7] = sr, ms[8] = corr, ms[9] = integ, ms[10] = total;
  return 0;
}

// Initial-condition samplers of the reference (implementation-defined RNG, SURVEY Q11: arrays are
// generated once here and fed identically to every implementation).
int ref_sample_plummer(unsigned seed, const float* center, float a, float rMax, float M, float G,
                       int n, float* pos, float* vel) {
  PlummerSampler sampler(seed);
  // call order used by the demos: (center, a, rMax, M, G, n) -- source/demos.cpp:1350, SURVEY Q12
  auto st = sampler.sample(Vec3(center[0], center[1], center[2]), a, rMax, M, G, n);
  for (int i = 0; i < n; ++i) {
    pos[3 * i] = st[i].x, pos[3 * i + 1] = st[i].y, pos[3 * i + 2] = st[i].z;
    vel[3 * i] = st[n + i].x, vel[3 * i + 1] = st[n + i].y, vel[3 * i + 2] = st[n + i].z;
  }
  return 0;
}

int ref_sample_disk_linear(unsigned seed, const float* center, float rb, float mb, float rd,
                           float md, float thickness, float G, int n, float* pos, float* vel) {
  DiskSamplerLinear sampler(seed);
  auto st = sampler.sample(Vec3(center[0], center[1], center[2]), rb, mb, rd, md, thickness, G, n);
  for (int i = 0; i < n; ++i) {
    pos[3 * i] = st[i].x, pos[3 * i + 1] = st[i].y, pos[3 * i + 2] = st[i].z;
    vel[3 * i] = st[n + i].x, vel[3 * i + 1] = st[n + i].y, vel[3 * i + 2] = st[n + i].z;
  }
  return 0;
}

int ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
