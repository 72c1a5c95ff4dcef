// support.cpp -- the small host-side pieces of the reference API that surround the hot path: Vec3,
// Grid (host view), CuFFTAdapter, ChainingMesh (host geometry), leapfrog free functions, unit
// conversions, single-mode Green functions, SimInfo, StateRecorder (file formats), external fields.
// All new code; the cited reference lines give the behaviour each piece reproduces.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <stdexcept>

#include "../../../include/p3m_b200.h"
#include "../include/particle_simulation_b200.hpp"


#if !P3M_B200_REFERENCE_TREE  // reference-tree mode: the reference's own vec3.cpp provides these
// ---- Vec3 (include/vec3.h, source/vec3.cpp) ------------------------------------------------------------
float Vec3::getMagnitude() const { return std::sqrt(x * x + y * y + z * z); }

char* Vec3::toString(char* singleBuf, std::size_t singleBufSize, char* vecBuf, std::size_t vecBufSize) const {
  // "%.6e %.6e %.6e" (source/vec3.cpp:58-75)
  (void)singleBuf;
  (void)singleBufSize;
  std::snprintf(vecBuf, vecBufSize, "%.6e %.6e %.6e", x, y, z);
  return vecBuf;
}

#endif  // !P3M_B200_REFERENCE_TREE

// ---- unit conversions (source/unitConversions.cpp:7-71) ---------------------------------------------------
void stateToCodeUnits(std::vector<Vec3>& state, float H, float DT) {
  const size_t n = state.size() / 2;  // [0, n) positions, [n, 2n) velocities
  for (size_t i = 0; i < n; ++i) state[i] = positionToCodeUntits(state[i], H);
  for (size_t i = n; i < state.size(); ++i) state[i] = velocityToCodeUntits(state[i], H, DT);
}
void stateToOriginalUnits(std::vector<Vec3>& state, float H, float DT) {
  const size_t n = state.size() / 2;
  for (size_t i = 0; i < n; ++i) state[i] = positionToOriginalUnits(state[i], H);
  for (size_t i = n; i < state.size(); ++i) state[i] = velocityToOriginalUnits(state[i], H, DT);
}
void velocitiesToCodeUnits(std::vector<Vec3>& v, float H, float DT) {
  for (auto& x : v) x = velocityToCodeUntits(x, H, DT);
}
void velocitiesToOriginalUnits(std::vector<Vec3>& v, float H, float DT) {
  for (auto& x : v) x = velocityToOriginalUnits(x, H, DT);
}
void integerStepVelocitiesToOriginalUnits(std::vector<Particle>& ps, float H, float DT) {
  for (auto& p : ps) p.integerStepVelocity = velocityToOriginalUnits(p.integerStepVelocity, H, DT);
}
void integerStepVelocitiesToCodeUnits(std::vector<Particle>& ps, float H, float DT) {
  for (auto& p : ps) p.integerStepVelocity = velocityToCodeUntits(p.integerStepVelocity, H, DT);
}

void stateToCodeUnits(std::vector<Particle>& ps, float H, float DT) {
  for (auto& p : ps) p.position = positionToCodeUntits(p.position, H), p.velocity = velocityToCodeUntits(p.velocity, H, DT);
}
void stateToOriginalUnits(std::vector<Particle>& ps, float H, float DT) {
  for (auto& p : ps) p.position = positionToOriginalUnits(p.position, H), p.velocity = velocityToOriginalUnits(p.velocity, H, DT);
}
void massToCodeUnits(std::vector<Particle>& ps, float H, float DT, float G) {
  for (auto& p : ps) p.mass = massToCodeUnits(p.mass, H, DT, G);
}
void massToOriginalUnits(std::vector<Particle>& ps, float H, float DT, float G) {
  for (auto& p : ps) p.mass = massToOriginalUnits(p.mass, H, DT, G);
}

// ---- leapfrog (source/leapfrog.cpp:5-24) on host vectors ---------------------------------------------------
void setHalfStepVelocities(std::vector<Particle>& ps, float dt) {
  for (auto& p : ps) p.velocity += 0.5f * dt * p.acceleration;
}
void setIntegerStepVelocities(std::vector<Particle>& ps, float dt) {
  for (auto& p : ps) p.integerStepVelocity = p.velocity + 0.5f * dt * p.acceleration;
}
void updateVelocities(std::vector<Particle>& ps, float dt) {
  for (auto& p : ps) p.velocity += dt * p.acceleration;
}
void updatePositions(std::vector<Particle>& ps, float dt) {
  for (auto& p : ps) p.position += dt * p.velocity;
}
void LeapfrogStepper::doStep(std::vector<Particle>& x, float dt) {
  // drift, force, kick -- the loop body of source/p3mMethod.cpp:101-153 with half-step velocities
  updatePositions(x, dt);
  force(x);
  updateVelocities(x, dt);
}

#if !P3M_B200_REFERENCE_TREE  // reference-tree mode: the reference's externalFields.cpp stays in the build
// ---- external fields (source/externalFields.cpp:4-24) ----------------------------------------------------
Vec3 sphRadDecrField(Vec3 pos, Vec3 center, float R, float M, float G) {
  const Vec3 d = pos - center;
  const float r = d.getMagnitude();
  const float g = r > R ? -G * M / (r * r) : -(G * M / std::pow(R, 3.0f)) * r * (4 - 3 * r / R);
  return g * (d / r);
}
float sphRadDecrFieldPotential(Vec3 pos, Vec3 center, float R, float M, float G) {
  const float r = (pos - center).getMagnitude();
  if (r > R) return -G * M / r;
  const float u = r / R;
  return G * M / R * (-2 + u * u * (2 - u));
}
#endif  // !P3M_B200_REFERENCE_TREE

// ---- single-mode Green functions (source/greensFunctions.cpp:122-220), evaluated in double ------------------
static double sincd(double x) { return x == 0 ? 1.0 : std::sin(x) / x; }

std::complex<float> GreenDiscreteLaplacian(int kx, int ky, int kz, std::tuple<int, int, int> dims) {
  if (kx == 0 && ky == 0 && kz == 0) return 0;
  const double pi = 3.14159265358979323846;
  const double sx = std::sin(pi * kx / std::get<0>(dims)), sy = std::sin(pi * ky / std::get<1>(dims)),
               sz = std::sin(pi * kz / std::get<2>(dims));
  return (float)(-0.25 / (sx * sx + sy * sy + sz * sz));
}

std::complex<float> GreenPoorMan(int i, int j, int k, std::tuple<int, int, int> dims) {
  if (i == 0 && j == 0 && k == 0) return 0.0f;
  const double pi = 3.14159265358979323846;
  auto [Nx, Ny, Nz] = dims;
  const int ki = (i <= Nx / 2) ? i : i - Nx, kj = (j <= Ny / 2) ? j : j - Ny, kk = (k <= Nz / 2) ? k : k - Nz;
  const double a = 2 * pi * ki / Nx, b = 2 * pi * kj / Ny, c = 2 * pi * kk / Nz;
  return (float)(-1.0 / (a * a + b * b + c * c));
}

std::complex<float> GreenOptimal(InterpolationScheme is, int kx, int ky, int kz, std::tuple<int, int, int> dims,
                                 float a, CloudShape cs, FiniteDiffScheme fds) {
  if (kx == 0 && ky == 0 && kz == 0) return 0;
  if ((int)is < 0 || (int)is > 2) throw std::invalid_argument("Not implemented");
  if ((int)fds < 0 || (int)fds > 1) throw std::invalid_argument("not implemented");
  const double pi = 3.14159265358979323846;
  const int N[3] = {std::get<0>(dims), std::get<1>(dims), std::get<2>(dims)};
  const int kk[3] = {kx, ky, kz};
  double k[3], d[3], denom = 1;
  for (int i = 0; i < 3; ++i) k[i] = 2 * pi * kk[i] / N[i];
  for (int i = 0; i < 3; ++i) {
    const double s = std::sin(k[i] / 2), c = std::cos(k[i] / 2);
    if (is == InterpolationScheme::TSC) denom *= 1 - s * s + 2.0 / 15 * s * s * s * s;
    if (is == InterpolationScheme::CIC) denom *= (1 + 2 * c * c) / 3.0;
    d[i] = fds == FiniteDiffScheme::TWO_POINT ? std::sin(k[i]) : 4.0 / 3 * std::sin(k[i]) - std::sin(2 * k[i]) / 6.0;
  }
  const double dnorm = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const int pw = is == InterpolationScheme::TSC ? 3 : (is == InterpolationScheme::CIC ? 2 : 1);
  double num[3] = {0, 0, 0};
  for (int n1 = -2; n1 <= 2; ++n1)
    for (int n2 = -2; n2 <= 2; ++n2)
      for (int n3 = -2; n3 <= 2; ++n3) {
        const double kn[3] = {k[0] + 2 * pi * n1, k[1] + 2 * pi * n2, k[2] + 2 * pi * n3};
        double u = 1;
        for (int i = 0; i < 3; ++i) u *= std::pow(sincd(kn[i] / 2), pw);
        const double k2 = kn[0] * kn[0] + kn[1] * kn[1] + kn[2] * kn[2], kl = std::sqrt(k2), x = kl * a / 2;
        const double s = cs == CloudShape::S1 ? -3 / (x * x * x) * (x * std::cos(x) - std::sin(x))
                                              : 12 / (x * x * x * x) * (2 - 2 * std::cos(x) - x * std::sin(x));
        for (int i = 0; i < 3; ++i) num[i] += -u * u * kn[i] * s * s / k2;
      }
  return (float)((d[0] * num[0] + d[1] * num[1] + d[2] * num[2]) / (dnorm * denom * denom));
}

// ---- CuFFTAdapter ---------------------------------------------------------------------------------------------
CuFFTAdapter::CuFFTAdapter(std::array<int, 3> dims) : dims(dims) {}
CuFFTAdapter::CuFFTAdapter(int* d, int ndims) : dims{1, 1, 1} {
  if (ndims < 1 || ndims > 3) throw std::invalid_argument("CuFFTAdapter: 1 to 3 dimensions");
  for (int i = 0; i < ndims; ++i) dims[3 - ndims + i] = d[i];
}
static std::vector<std::complex<float>>& runFft(const std::array<int, 3>& dims, std::vector<std::complex<float>>& in,
                                                std::vector<std::complex<float>>& out, int inverse) {
  const size_t n = (size_t)dims[0] * dims[1] * dims[2];
  if (in.size() != n) throw std::invalid_argument("CuFFTAdapter: input length does not match the dims");
  out.resize(n);
  if (p3m_fft3d_c2c(dims[0], dims[1], dims[2], reinterpret_cast<const float*>(in.data()),
                    reinterpret_cast<float*>(out.data()), inverse) != P3M_OK)
    throw std::runtime_error(std::string("CuFFTAdapter: ") + p3m_last_error());
  return out;
}
std::vector<std::complex<float>>& CuFFTAdapter::fft(std::vector<std::complex<float>>& in,
                                                    std::vector<std::complex<float>>& out) {
  return runFft(dims, in, out, 0);
}
std::vector<std::complex<float>>& CuFFTAdapter::ifft(std::vector<std::complex<float>>& in,
                                                     std::vector<std::complex<float>>& out) {
  return runFft(dims, in, out, 1);
}

// ---- Grid (include/grid.h, source/grid.cpp) -------------------------------------------------------------------
Grid::Grid(std::tuple<int, int, int> g, FFTAdapter<float>& fftAdapter)
    : gridPointsX(std::get<0>(g)), gridPointsY(std::get<1>(g)), gridPointsZ(std::get<2>(g)),
      length(gridPointsX * gridPointsY * gridPointsZ), field(length), density(length), densityFourier(length),
      potential(length), potentialFourier(length), greensFunction(length), fftAdapter(fftAdapter) {}

std::tuple<int, int, int> Grid::indexTripleFromFlat(int f) const {
  return std::make_tuple(f % gridPointsX, (f / gridPointsX) % gridPointsY, f / (gridPointsX * gridPointsY));
}
int Grid::wrapped(int i, int j, int k) const {
  auto m = [](int a, int b) { return (a % b + b) % b; };
  return m(i, gridPointsX) + m(j, gridPointsY) * gridPointsX + m(k, gridPointsZ) * gridPointsX * gridPointsY;
}
void Grid::assignDensity(int x, int y, int z, float d) { density[flat(x, y, z)] += d; }
void Grid::clearDensity() { std::fill(density.begin(), density.end(), std::complex<float>(0, 0)); }
float Grid::getDensity(int x, int y, int z) const { return density[flat(x, y, z)].real(); }
void Grid::assignField(int x, int y, int z, Vec3 v) { field[flat(x, y, z)] = v; }
Vec3 Grid::getField(int x, int y, int z) const { return field[flat(x, y, z)]; }
const std::vector<std::complex<float>>& Grid::fftDensity() { return fftAdapter.fft(density, densityFourier); }
const std::vector<std::complex<float>>& Grid::invFftPotential() { return fftAdapter.ifft(potentialFourier, potential); }
void Grid::setPotentialFourier(int i, int j, int k, std::complex<float> v) { potentialFourier[flat(i, j, k)] = v; }
std::complex<float> Grid::getDensityFourier(int i, int j, int k) const { return densityFourier[flat(i, j, k)]; }
float Grid::getPotential(int i, int j, int k) const { return potential[wrapped(i, j, k)].real(); }
std::complex<float> Grid::getGreensFunction(int i, int j, int k) const { return greensFunction[flat(i, j, k)]; }
void Grid::setGreensFunction(int i, int j, int k, std::complex<float> v) { greensFunction[flat(i, j, k)] = v; }

// ---- ChainingMesh (include/chainingMesh.h, source/chainingMesh.cpp) on the host --------------------------------
ChainingMesh::ChainingMesh(std::tuple<float, float, float> box, float cutoffRadius, float H, int N)
    : Mx(int(std::get<0>(box) / cutoffRadius)), My(int(std::get<1>(box) / cutoffRadius)),
      Mz(int(std::get<2>(box) / cutoffRadius)), HCx(lengthToCodeUnits(std::get<0>(box) / Mx, H)),
      HCy(lengthToCodeUnits(std::get<1>(box) / My, H)), HCz(lengthToCodeUnits(std::get<2>(box) / Mz, H)),
      size(Mx * My * Mz), hoc(size, nullptr), nodePool(new LLNode[N > 0 ? N : 1]) {}

int ChainingMesh::cellOf(const Particle& p) const {
  const int cx = int(p.position.x / HCx), cy = int(p.position.y / HCy), cz = int(p.position.z / HCz);
  if (cx < 0 || cy < 0 || cz < 0 || cx >= Mx || cy >= My || cz >= Mz) return -1;
  return cx + cy * Mx + cz * Mx * My;
}

void ChainingMesh::fill(const std::vector<Particle>& ps) {
  std::fill(hoc.begin(), hoc.end(), nullptr);
  for (int i = 0; i < (int)ps.size(); ++i) {  // head insertion: lists end up in descending id
    const int c = cellOf(ps[i]);
    if (c < 0) continue;
    nodePool[i] = LLNode(i, hoc[c]);
    hoc[c] = &nodePool[i];
  }
}

void ChainingMesh::fillWithYSorting(const std::vector<Particle>& ps) {
  // same lists as the reference's sorted insertion (ascending y, ties by id), built by sorting
  std::fill(hoc.begin(), hoc.end(), nullptr);
  std::vector<int> order(ps.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ps[a].position.y < ps[b].position.y; });
  for (auto it = order.rbegin(); it != order.rend(); ++it) {
    const int i = *it, c = cellOf(ps[i]);
    if (c < 0) continue;
    nodePool[i] = LLNode(i, hoc[c]);
    hoc[c] = &nodePool[i];
  }
}

std::array<int, 14> ChainingMesh::getNeighborsAndSelf(int cellIdx) const {
  const int32_t dims[3] = {Mx, My, Mz};
  int32_t nb[14];
  if (p3m_chaining_neighbors(dims, cellIdx, nb) != P3M_OK) throw std::out_of_range(p3m_last_error());
  std::array<int, 14> out;
  std::copy(nb, nb + 14, out.begin());
  return out;
}

// ---- SimInfo (source/simInfo.cpp:50-127) on host vectors ----------------------------------------------------------
float SimInfo::kineticEnergy(const std::vector<Particle>& ps) {
  float ke = 0;
  for (const auto& p : ps) ke += 0.5f * p.mass * p.velocity.getMagnitudeSquared();
  return ke;
}
Vec3 SimInfo::totalMomentum(const std::vector<Particle>& ps) {
  Vec3 m;
  for (const auto& p : ps) m += p.mass * p.integerStepVelocity;
  return m;
}
Vec3 SimInfo::totalAngularMomentum(const std::vector<Particle>& ps) {
  Vec3 L;
  for (const auto& p : ps) L += p.mass * p.position.cross(p.integerStepVelocity);
  return L;
}
float SimInfo::potentialEnergy(const Grid& grid, const std::vector<Particle>& ps,
                               std::function<float(Vec3)> externalPotential, float H, float DT, float G) {
  double internal = 0;
  auto [nx, ny, nz] = grid.getGridPoints();
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int x = 0; x < nx; ++x)
        internal += (double)densityToOriginalUnits(grid.getDensity(x, y, z), DT, G) *
                    (double)potentialToOriginalUnits(grid.getPotential(x, y, z), H, DT);
  double external = 0;
  if (externalPotential)
    for (const auto& p : ps) external += p.mass * externalPotential(p.position);
  return (float)(0.5 * H * H * H * internal + external);
}
// the CUDA-build overload (include/simInfo.h:30-36, source/simInfo.cpp:72-92): density / potential vectors as
// PMMethodGPU::getGridDensity / getGridPotential return them (real parts, code units)
float SimInfo::potentialEnergy(const std::vector<std::complex<float>>& gridDensity,
                               const std::vector<std::complex<float>>& gridPotential, const std::vector<Particle>& ps,
                               std::function<float(Vec3)> externalPotential, float H, float DT, float G) {
  double internal = 0;
  const size_t M = std::min(gridDensity.size(), gridPotential.size());
  for (size_t i = 0; i < M; ++i)
    internal += (double)densityToOriginalUnits(gridDensity[i].real(), DT, G) *
                (double)potentialToOriginalUnits(gridPotential[i].real(), H, DT);
  double external = 0;
  if (externalPotential)
    for (const auto& p : ps) external += p.mass * externalPotential(p.position);
  return (float)(0.5 * H * H * H * internal + external);
}
// state-vector overloads (source/simInfo.cpp:4-48): direct pair sum with the reference's fixed softening 0.01
float SimInfo::potentialEnergy(std::vector<Vec3>::iterator posBegin, std::vector<Vec3>::iterator posEnd,
                               const std::vector<float>& masses, float G) {
  const float eps = 0.01f;
  float pe = 0;
  int i = 0;
  for (auto a = posBegin; a != posEnd; ++a, ++i) {
    int j = i + 1;
    for (auto b = a + 1; b != posEnd; ++b, ++j)
      pe += (-1) * G * masses[i] * masses[j] / std::sqrt((*a - *b).getMagnitudeSquared() + eps * eps);
  }
  return pe;
}
float SimInfo::kineticEnergy(std::vector<Vec3>::iterator vBegin, std::vector<Vec3>::iterator vEnd,
                             const std::vector<float>& masses, float /*G*/) {
  float ke = 0;
  int i = 0;
  for (auto v = vBegin; v != vEnd; ++v, ++i) ke += 0.5f * masses[i] * v->getMagnitudeSquared();
  return ke;
}
Vec3 SimInfo::totalMomentum(std::vector<Vec3>::iterator vBegin, std::vector<Vec3>::iterator vEnd,
                            const std::vector<float>& masses) {
  Vec3 m = Vec3::zero();
  int i = 0;
  for (auto v = vBegin; v != vEnd; ++v, ++i) m += masses[i] * (*v);
  return m;
}
void SimInfo::setInitialMomentum(const std::vector<Particle>& particles) { expectedMomentum = totalMomentum(particles); }
Vec3 SimInfo::updateExpectedMomentum(Vec3 externalForce, float DT) {
  expectedMomentum += DT * externalForce;
  return expectedMomentum;
}

#if !P3M_B200_REFERENCE_TREE  // reference-tree mode: the reference's stateRecorder.cpp stays in the build
// ---- StateRecorder (source/stateRecorder.cpp; formats read by script/load_data.py:15-37) ---------------------------
StateRecorder::StateRecorder(int particlesCnt, int framesCnt, const std::filesystem::path& outputDirPath,
                             const char* positionsFile, const char* energyFile, const char* momentumFile,
                             const char* expectedMomentumFile, const char* angularMomentumFile,
                             const char* fieldFile, int /*maxRecords*/)
    : dir(outputDirPath), particlesCnt(particlesCnt), framesCnt(framesCnt) {
  std::filesystem::create_directories(dir);
  positions.open(dir / positionsFile, std::ios::binary | std::ios::trunc);
  field.open(dir / fieldFile, std::ios::binary | std::ios::trunc);
  energy.open(dir / energyFile, std::ios::trunc);
  momentum.open(dir / momentumFile, std::ios::trunc);
  expectedMomentum.open(dir / expectedMomentumFile, std::ios::trunc);
  angularMomentum.open(dir / angularMomentumFile, std::ios::trunc);
  for (std::ofstream* f : {&positions, &field}) {  // header: int32 n, int32 frames
    f->write(reinterpret_cast<const char*>(&this->particlesCnt), sizeof(int));
    f->write(reinterpret_cast<const char*>(&this->framesCnt), sizeof(int));
  }
}
StateRecorder::~StateRecorder() { flush(); }

void StateRecorder::recordPositions(const float* xyz, std::size_t n) {
  positions.write(reinterpret_cast<const char*>(xyz), sizeof(float) * 3 * n);
}
void StateRecorder::recordPositions(std::vector<Vec3>::iterator b, std::vector<Vec3>::iterator e) {
  for (auto it = b; it != e; ++it) positions.write(reinterpret_cast<const char*>(&*it), sizeof(Vec3));
}
void StateRecorder::recordPositions(const std::vector<Particle>& ps) {
  for (const auto& p : ps) positions.write(reinterpret_cast<const char*>(&p.position), sizeof(Vec3));
}
void StateRecorder::recordField(const float* a, std::size_t n) {
  field.write(reinterpret_cast<const char*>(a), sizeof(float) * 3 * n);
}
void StateRecorder::recordField(const std::vector<Particle>& ps, float H, float DT) {
  for (const auto& p : ps) {
    const Vec3 a = accelerationToOriginalUnits(p.acceleration, H, DT);
    field.write(reinterpret_cast<const char*>(&a), sizeof(Vec3));
  }
}
void StateRecorder::recordEnergy(float pe, float ke) { energy << std::to_string(pe) << ' ' << std::to_string(ke) << '\n'; }
void StateRecorder::writeVec(std::ofstream& f, Vec3 v) {
  char buf[160];
  std::snprintf(buf, sizeof(buf), "%.6e %.6e %.6e\n", v.x, v.y, v.z);
  f << buf;
}
void StateRecorder::recordTotalMomentum(Vec3 v) { writeVec(momentum, v); }
void StateRecorder::recordExpectedMomentum(Vec3 v) { writeVec(expectedMomentum, v); }
void StateRecorder::recordTotalAngularMomentum(Vec3 v) { writeVec(angularMomentum, v); }
std::string StateRecorder::flush() {
  for (std::ofstream* f : {&positions, &field, &energy, &momentum, &expectedMomentum, &angularMomentum}) f->flush();
  return dir.string();
}
#endif  // !P3M_B200_REFERENCE_TREE
