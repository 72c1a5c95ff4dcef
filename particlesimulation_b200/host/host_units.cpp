// host_units.cpp -- drives the HOST-side classes of the reference API that surround the hot path
// (ChainingMesh, GreenOptimal / GreenDiscreteLaplacian / GreenPoorMan, the leapfrog free functions,
// LeapfrogStepper, unit conversions, SimInfo, Grid) and, with a GPU, the PMMethodGPU extras (getGrid()
// back-fill, copyGrid*ToHost, getGridDensity / getGridPotential, copyParticles*, SimInfo::potentialEnergy
// overloads).  It only dumps what these classes return; tests/test_host_units.py compares the dump with the
// unmodified reference (oracle/_ref) and with numpy restatements.  Test infrastructure, built by the Makefile.
//
//   host_units cpu <in.bin> <out.bin>      no GPU needed
//   host_units gpu <in.bin> <out.bin>      needs a CUDA device
//   in.bin : int32 n, float32 box[3], cutoff, H, DT, G, then pos[3n], vel[3n], mass[n] (original units)
//   out.bin: records  name[32] | dtype ('f' float32, 'i' int32) | int64 count | data
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <numeric>

#include "chainingMesh.h"
#include "greensFunctions.h"
#include "grid.h"
#include "leapfrog.h"
#include "p3mMethod.h"
#include "pmMethod.h"
#include "simInfo.h"
#include "unitConversions.h"
#include "cuFFTAdapter.h"

static FILE* g_out;

static void put(const char* name, char dtype, const void* data, long long count) {
  char nm[32] = {0};
  std::strncpy(nm, name, 31);
  std::fwrite(nm, 1, 32, g_out);
  std::fwrite(&dtype, 1, 1, g_out);
  std::fwrite(&count, sizeof(long long), 1, g_out);
  std::fwrite(data, 4, (size_t)count, g_out);
}
static void putf(const char* name, const std::vector<float>& v) { put(name, 'f', v.data(), (long long)v.size()); }
static void puti(const char* name, const std::vector<int>& v) { put(name, 'i', v.data(), (long long)v.size()); }
static void putv(const char* name, const std::vector<Particle>& ps, Vec3 Particle::*m) {
  std::vector<float> v;
  for (const auto& p : ps) v.push_back((p.*m).x), v.push_back((p.*m).y), v.push_back((p.*m).z);
  putf(name, v);
}

struct Input {
  int n;
  float box[3], cutoff, H, DT, G;
  std::vector<Vec3> state;
  std::vector<float> masses;
};

static bool readInput(const char* path, Input& in) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  bool ok = std::fread(&in.n, 4, 1, f) == 1 && std::fread(in.box, 4, 3, f) == 3 && std::fread(&in.cutoff, 4, 1, f) == 1 &&
            std::fread(&in.H, 4, 1, f) == 1 && std::fread(&in.DT, 4, 1, f) == 1 && std::fread(&in.G, 4, 1, f) == 1;
  if (ok) {
    in.state.resize(2 * (size_t)in.n);
    in.masses.resize(in.n);
    ok = std::fread(in.state.data(), sizeof(Vec3), 2 * (size_t)in.n, f) == 2 * (size_t)in.n &&
         std::fread(in.masses.data(), 4, in.n, f) == (size_t)in.n;
  }
  std::fclose(f);
  return ok;
}

static std::vector<Particle> particlesOf(const Input& in) {
  std::vector<Particle> ps;
  for (int i = 0; i < in.n; ++i) ps.emplace_back(in.state[i], in.state[in.n + i], in.masses[i]);
  return ps;
}

static int runCpu(const Input& in) {
  auto box = std::make_tuple(in.box[0], in.box[1], in.box[2]);
  // ---- ChainingMesh on code-unit positions -------------------------------------------------------------
  std::vector<Particle> ps = particlesOf(in);
  stateToCodeUnits(ps, in.H, in.DT);
  putv("pos_code", ps, &Particle::position);
  putv("vel_code", ps, &Particle::velocity);
  ChainingMesh cm(box, in.cutoff, in.H, in.n);
  auto [Mx, My, Mz] = cm.getLength();
  puti("cm_dims", {Mx, My, Mz, cm.getSize()});
  for (int sorted = 0; sorted < 2; ++sorted) {
    if (sorted) cm.fillWithYSorting(ps); else cm.fill(ps);
    std::vector<int> cellOf(in.n, -1), listOrder;
    for (int c = 0; c < cm.getSize(); ++c)
      for (auto* node = cm.getParticlesInCell(c); node; node = node->next) {
        cellOf[node->particleId] = c;
        listOrder.push_back(node->particleId);
      }
    puti(sorted ? "cm_cell_ysort" : "cm_cell", cellOf);
    puti(sorted ? "cm_order_ysort" : "cm_order", listOrder);
  }
  std::vector<int> nb;
  for (int c = 0; c < cm.getSize(); ++c) {
    auto a = cm.getNeighborsAndSelf(c);
    nb.insert(nb.end(), a.begin(), a.end());
  }
  puti("cm_neighbors", nb);
  // ---- single-mode influence functions on an 8 x 6 x 4 mesh --------------------------------------------------
  auto dims = std::make_tuple(8, 6, 4);
  std::vector<float> gl, gp, g1, g2;
  for (int kz = 0; kz < 4; ++kz)
    for (int ky = 0; ky < 6; ++ky)
      for (int kx = 0; kx < 8; ++kx) {
        gl.push_back(GreenDiscreteLaplacian(kx, ky, kz, dims).real());
        gp.push_back(GreenPoorMan(kx, ky, kz, dims).real());
        g1.push_back(GreenOptimal(InterpolationScheme::TSC, kx, ky, kz, dims, 3.0f, CloudShape::S1, FiniteDiffScheme::TWO_POINT).real());
        g2.push_back(GreenOptimal(InterpolationScheme::CIC, kx, ky, kz, dims, 2.5f, CloudShape::S2, FiniteDiffScheme::FOUR_POINT).real());
      }
  putf("green_laplacian", gl), putf("green_poorman", gp), putf("green_s1_tsc_2pt", g1), putf("green_s2_cic_4pt", g2);
  // ---- leapfrog free functions + LeapfrogStepper --------------------------------------------------------------
  std::vector<Particle> lp = particlesOf(in);
  for (int i = 0; i < in.n; ++i) lp[i].acceleration = Vec3::create(0.01f * (i % 7), -0.02f * (i % 5), 0.005f * (i % 3));
  setHalfStepVelocities(lp);
  putv("lf_half_vel", lp, &Particle::velocity);
  updatePositions(lp);
  putv("lf_pos", lp, &Particle::position);
  updateVelocities(lp, 0.5f);
  putv("lf_vel", lp, &Particle::velocity);
  setIntegerStepVelocities(lp);
  putv("lf_int_vel", lp, &Particle::integerStepVelocity);
  std::vector<Particle> st = particlesOf(in);
  LeapfrogStepper stepper([](std::vector<Particle>& x) {
    for (auto& p : x) p.acceleration = Vec3::create(-0.001f * p.position.x, -0.001f * p.position.y, -0.001f * p.position.z);
  });
  AbstractStepper<Particle>& abstractStepper = stepper;
  abstractStepper.doStep(st, 1.0f);
  abstractStepper.doStep(st, 0.5f);
  putv("stepper_pos", st, &Particle::position);
  putv("stepper_vel", st, &Particle::velocity);
  // ---- unit conversions + SimInfo ---------------------------------------------------------------------------------
  std::vector<Particle> up = particlesOf(in);
  massToCodeUnits(up, in.H, in.DT, in.G);
  std::vector<float> mc;
  for (auto& p : up) mc.push_back(p.mass);
  putf("mass_code", mc);
  massToOriginalUnits(up, in.H, in.DT, in.G);
  std::vector<Vec3> sv = in.state;
  stateToCodeUnits(sv, in.H, in.DT);
  stateToOriginalUnits(sv, in.H, in.DT);
  put("state_roundtrip", 'f', sv.data(), 3 * (long long)sv.size());
  std::vector<float> scal = {densityToCodeUnits(2.0f, in.DT, in.G), densityToOriginalUnits(2.0f, in.DT, in.G),
                             potentialToOriginalUnits(2.0f, in.H, in.DT), lengthToCodeUnits(2.0f, in.H),
                             massToCodeUnits(2.0f, in.H, in.DT, in.G), massToOriginalUnits(2.0f, in.H, in.DT, in.G)};
  putf("unit_scalars", scal);
  std::vector<Particle> sp = particlesOf(in);
  for (auto& p : sp) p.acceleration = Vec3::create(0.01f, 0.02f, -0.01f);
  setIntegerStepVelocities(sp);
  Vec3 mom = SimInfo::totalMomentum(sp), L = SimInfo::totalAngularMomentum(sp);
  std::vector<Vec3> s2 = in.state;
  const int m = std::min(in.n, 200);  // the direct-sum overload is O(n^2)
  std::vector<Vec3> sub(s2.begin(), s2.begin() + m);
  std::vector<float> subm(in.masses.begin(), in.masses.begin() + m);
  std::vector<Vec3> subv(s2.begin() + in.n, s2.begin() + in.n + m);
  Vec3 mom2 = SimInfo::totalMomentum(subv.begin(), subv.end(), subm);
  SimInfo info;
  info.setInitialMomentum(sp);
  Vec3 em = info.updateExpectedMomentum(Vec3::create(1, 2, 3), 0.5f);
  std::vector<float> si = {SimInfo::kineticEnergy(sp), mom.x, mom.y, mom.z, L.x, L.y, L.z,
                           SimInfo::potentialEnergy(sub.begin(), sub.end(), subm, in.G),
                           SimInfo::kineticEnergy(subv.begin(), subv.end(), subm, in.G), mom2.x, mom2.y, mom2.z,
                           em.x, em.y, em.z};
  putf("siminfo", si);
  return 0;
}

static int runGpu(const Input& in) {
  auto box = std::make_tuple(in.box[0], in.box[1], in.box[2]);
  const auto gridPoints = std::make_tuple(32, 32, 16);
  Vec3 center = Vec3::create(in.box[0] / 2, in.box[1] / 2, in.box[2] / 2);
  const float rb = 3.0f, mb = 60.0f, G = in.G;
  auto field = [=](Vec3 p) -> Vec3 { return sphRadDecrField(p, center, rb, mb, G); };
  auto pot = [=](Vec3 p) -> float { return sphRadDecrFieldPotential(p, center, rb, mb, G); };
  // (1) reference-style constructor with a caller-owned Grid; getGrid() const back-fills it
  std::array<int, 3> dims = {16, 32, 32};
  CuFFTAdapter fft(dims);
  Grid grid(gridPoints, fft);
  PMMethod pm(in.state, in.masses, box, field, pot, in.H, in.DT, in.G, InterpolationScheme::TSC,
              FiniteDiffScheme::TWO_POINT, GreensFunction::DISCRETE_LAPLACIAN, 0, grid);
  auto& ps = pm.getParticles();
  stateToCodeUnits(ps, in.H, in.DT);  // what the head of run() does to the vector (source/pmMethod.cpp:72-73)
  massToCodeUnits(ps, in.H, in.DT, in.G);
  pm.initGreensFunction();
  pm.pmMethodStep();
  const PMMethod& cpm = pm;
  const Grid& g = cpm.getGrid();
  std::vector<float> rho, phi;
  auto [nx, ny, nz] = g.getGridPoints();
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int x = 0; x < nx; ++x) rho.push_back(g.getDensity(x, y, z)), phi.push_back(g.getPotential(x, y, z));
  putf("grid_density", rho), putf("grid_potential", phi);
  std::vector<float> wrap = {g.getPotential(-1, ny, nz + 1), g.getPotential(nx - 1, 0, 1)};
  putf("grid_potential_wrap", wrap);
  auto& after = pm.getParticles();
  putv("pm_acc", after, &Particle::acceleration);
  putv("pm_pos", after, &Particle::position);
  // back to original units IN PLACE, as the run loops do before the diagnostics (source/pmMethod.cpp:94-105)
  stateToOriginalUnits(after, in.H, in.DT);
  massToOriginalUnits(after, in.H, in.DT, in.G);
  Vec3 ext = pm.totalExternalForceOrigUnits();
  std::vector<float> misc = {SimInfo::potentialEnergy(g, after, pm.getExternalPotential(), in.H, in.DT, in.G), ext.x, ext.y, ext.z,
                             pm.escapedComputationalBox() ? 1.0f : 0.0f, pm.getH(), pm.getDT(), pm.getG()};
  // (2) CUDA-build spelling: mesh size instead of a Grid, explicit copies, vector getters
  PMMethodGPU gpu(in.state, in.masses, box, field, pot, in.H, in.DT, in.G, InterpolationScheme::TSC,
                  FiniteDiffScheme::TWO_POINT, GreensFunction::DISCRETE_LAPLACIAN, 0, gridPoints);
  auto& gp = gpu.getParticles();
  stateToCodeUnits(gp, in.H, in.DT);
  massToCodeUnits(gp, in.H, in.DT, in.G);
  gpu.initGreensFunction();
  gpu.copyParticlesHostToDevice();
  gpu.pmMethodStep();
  gpu.copyParticlesDeviceToHost();
  gpu.copyGridDensityToHost();
  gpu.copyGridPotentialToHost();
  std::vector<float> rho2, phi2;
  for (auto& v : gpu.getGridDensity()) rho2.push_back(v.real());
  for (auto& v : gpu.getGridPotential()) phi2.push_back(v.real());
  putf("gpu_density", rho2), putf("gpu_potential", phi2);
  std::vector<Particle> gorig = gpu.getParticles();
  putv("gpu_acc", gorig, &Particle::acceleration);
  stateToOriginalUnits(gorig, in.H, in.DT);
  massToOriginalUnits(gorig, in.H, in.DT, in.G);
  misc.push_back(SimInfo::potentialEnergy(gpu.getGridDensity(), gpu.getGridPotential(), gorig, gpu.getExternalPotential(),
                                          in.H, in.DT, in.G));
  // (3) P3MMethod argument checks: its own particleDiameter is honoured, a different H is refused
  int refused = 0;
  try {
    P3MMethod bad(gpu, box, in.cutoff, 3 * in.H, 2 * in.H, 0.5f, CloudShape::S1);
  } catch (const std::invalid_argument&) {
    refused = 1;
  }
  misc.push_back((float)refused);
  putf("misc", misc);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s <cpu|gpu> <in.bin> <out.bin>\n", argv[0]);
    return 2;
  }
  Input in;
  if (!readInput(argv[2], in)) {
    std::fprintf(stderr, "cannot read %s\n", argv[2]);
    return 2;
  }
  g_out = std::fopen(argv[3], "wb");
  if (!g_out) return 2;
  int rc = 1;
  try {
    rc = std::strcmp(argv[1], "gpu") == 0 ? runGpu(in) : runCpu(in);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
  }
  std::fclose(g_out);
  return rc;
}
