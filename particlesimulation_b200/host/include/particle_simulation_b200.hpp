// particle_simulation_b200.hpp -- the reference's C++ API surface for the P3M / PM path, re-hosted on
// the B200 library.  A caller of AleksyBalazinski/ParticleSimulation (its source/demos.cpp, or any
// code written against include/pmMethod.h, include/p3mMethod.h, include/grid.h, include/FFTAdapter.h,
// include/greensFunctions.h, include/chainingMesh.h, include/leapfrog.h, include/simInfo.h,
// include/unitConversions.h) compiles against these declarations unchanged: same class names,
// constructor argument order, member names and error behaviour.  The bodies (host/src/*.cpp) are thin
// callers of the C ABI in include/p3m_b200.h; particles live on the GPU and the host
// std::vector<Particle> is only materialised when the caller asks for it.
//
// Two ways to build against it:
//  * INSIDE the reference checkout ("reference-tree mode"):  -I <this dir> -I <reference>/include.
//    The headers this directory provides (pmMethod.h, p3mMethod.h, PMMethodGPU.h, grid.h, greensFunctions.h,
//    chainingMesh.h, leapfrog.h, simInfo.h, unitConversions.h, cuFFTAdapter.h) shadow the reference's; the
//    plain data types the rest of the reference tree shares -- Vec3, Particle, the pmConfig enums,
//    StateRecorder, FFTAdapter<T>, AbstractStepper<T>, the external-field functions -- are NOT redefined
//    here: the reference's own vec3.h / particle.h / pmConfig.h / stateRecorder.h / FFTAdapter.h /
//    abstractStepper.h / externalFields.h are included, so samplers, Barnes-Hut, the direct-sum method and
//    the demos keep compiling in the same translation unit, and host/src/*.cpp are compiled in place of
//    pmMethod.cpp, p3mMethod.cpp, grid.cpp, greensFunctions.cpp, chainingMesh.cpp, leapfrog.cpp,
//    simInfo.cpp and unitConversions.cpp.
//  * WITHOUT the reference ("standalone mode"):  -I <this dir> -I <this dir>/standalone.  The same plain
//    types are defined below and implemented in host/src/support.cpp (libparticlesim_host.so).
#pragma once

#include <array>
#include <complex>
#include <cstddef>
#include <filesystem>
#include <fstream>
#include <functional>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

struct p3m_ctx;  // include/p3m_b200.h

#if !defined(P3M_B200_STANDALONE) && __has_include("p3m_b200_standalone.h")
#include "p3m_b200_standalone.h"
#endif
#if !defined(P3M_B200_STANDALONE) && __has_include("vec3.h") && __has_include("particle.h") && \
    __has_include("pmConfig.h") && __has_include("stateRecorder.h") && __has_include("FFTAdapter.h") && \
    __has_include("abstractStepper.h") && __has_include("externalFields.h")
#define P3M_B200_REFERENCE_TREE 1
#include "FFTAdapter.h"
#include "abstractStepper.h"
#include "externalFields.h"
#include "particle.h"
#include "pmConfig.h"
#include "stateRecorder.h"
#include "vec3.h"
#else
#define P3M_B200_REFERENCE_TREE 0
#endif

#if !P3M_B200_REFERENCE_TREE
// ---- include/vec3.h ---------------------------------------------------------------------------------
struct Vec3 {
  float x{};
  float y{};
  float z{};

  static Vec3 create(float x, float y, float z) { return Vec3{x, y, z}; }
  static Vec3 zero() { return Vec3{0, 0, 0}; }
  Vec3& operator+=(const Vec3 o) {
    x += o.x, y += o.y, z += o.z;
    return *this;
  }
  Vec3& operator/=(float s) {
    x /= s, y /= s, z /= s;
    return *this;
  }
  Vec3 cross(const Vec3 o) const { return Vec3{y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x}; }
  float getMagnitudeSquared() const { return x * x + y * y + z * z; }
  float getMagnitude() const;
  char* toString(char* singleBuf, std::size_t singleBufSize, char* vecBuf, std::size_t vecBufSize) const;
};
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(float s, const Vec3& a) { return Vec3{s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator/(const Vec3& a, float s) { return Vec3{a.x / s, a.y / s, a.z / s}; }

// ---- include/particle.h ------------------------------------------------------------------------------
struct Particle {
  Vec3 position;
  Vec3 velocity;
  Vec3 acceleration;
  float mass;
  Vec3 integerStepVelocity;
  Vec3 shortRangeForce;                           // here: the TOTAL short-range force on the particle
  std::array<Vec3, 13> shortRangeFromNeighbor{};  // kept for source compatibility; always zero (the
                                                  // GPU path gathers, it has no Newton-3 slots)
  Particle(Vec3 position, Vec3 velocity, float mass)
      : position(position), velocity(velocity), mass(mass), integerStepVelocity(velocity) {}
};

// ---- include/pmConfig.h, include/greensFunctions.h -------------------------------------------------------
enum class InterpolationScheme { NGP, CIC, TSC };
enum class FiniteDiffScheme { TWO_POINT, FOUR_POINT };
enum class GreensFunction { DISCRETE_LAPLACIAN, S1_OPTIMAL, S2_OPTIMAL, POOR_MAN };
#endif  // !P3M_B200_REFERENCE_TREE

// ---- include/greensFunctions.h ---------------------------------------------------------------------------
enum class CloudShape { S1, S2 };

// Single-mode evaluations with the reference's signatures (host, double inside, float out); the run
// loops use the device table (p3m_green_init) instead of calling these M times.
std::complex<float> GreenOptimal(InterpolationScheme is, int kx, int ky, int kz, std::tuple<int, int, int> dims,
                                 float a, CloudShape cs, FiniteDiffScheme fds);
std::complex<float> GreenDiscreteLaplacian(int kx, int ky, int kz, std::tuple<int, int, int> dims);
std::complex<float> GreenPoorMan(int kx, int ky, int kz, std::tuple<int, int, int> dims);

#if !P3M_B200_REFERENCE_TREE
// ---- include/FFTAdapter.h ------------------------------------------------------------------------------
template <typename T>
class FFTAdapter {
 public:
  virtual ~FFTAdapter() {}
  virtual std::vector<std::complex<T>>& fft(std::vector<std::complex<T>>& in, std::vector<std::complex<T>>& out) = 0;
  virtual std::vector<std::complex<T>>& ifft(std::vector<std::complex<T>>& in, std::vector<std::complex<T>>& out) = 0;
};

#endif  // !P3M_B200_REFERENCE_TREE

// The GPU backend as a proper FFTAdapter<float> (the reference's CuFFTAdapter does not derive from
// the interface).  dims = {Nz, Ny, Nx} as every reference adapter takes them
// (source/demos.cpp:758-759); forward unnormalised, inverse divided by the length
// (test/fftAdaptersTest.cpp:6-22).
class CuFFTAdapter : public FFTAdapter<float> {
 public:
  explicit CuFFTAdapter(std::array<int, 3> dims);
  CuFFTAdapter(int* dims, int ndims);
  std::vector<std::complex<float>>& fft(std::vector<std::complex<float>>& in,
                                        std::vector<std::complex<float>>& out) override;
  std::vector<std::complex<float>>& ifft(std::vector<std::complex<float>>& in,
                                         std::vector<std::complex<float>>& out) override;

 private:
  std::array<int, 3> dims;
};

// ---- include/grid.h ------------------------------------------------------------------------------------
// Host view of the mesh.  The solvers keep the meshes on the device; a Grid handed to PMMethod is
// back-filled (density, potential, field on request) whenever host code reads it.
class Grid {
 public:
  Grid(std::tuple<int, int, int> gridPoints, FFTAdapter<float>& fftAdapter);

  std::tuple<int, int, int> indexTripleFromFlat(int flatIndex) const;
  void assignDensity(int x, int y, int z, float density);
  void clearDensity();
  float getDensity(int x, int y, int z) const;
  void assignField(int x, int y, int z, Vec3 fieldVal);
  Vec3 getField(int x, int y, int z) const;
  int getLength() const { return length; }
  std::tuple<int, int, int> getGridPoints() const { return std::make_tuple(gridPointsX, gridPointsY, gridPointsZ); }
  const std::vector<std::complex<float>>& fftDensity();
  const std::vector<std::complex<float>>& invFftPotential();
  void setPotentialFourier(int i, int j, int k, std::complex<float> value);
  std::complex<float> getDensityFourier(int i, int j, int k) const;
  float getPotential(int i, int j, int k) const;
  std::complex<float> getGreensFunction(int i, int j, int k) const;
  void setGreensFunction(int i, int j, int k, std::complex<float> value);

  // back-fill hooks used by PMMethod
  std::vector<std::complex<float>>& densityStorage() { return density; }
  std::vector<std::complex<float>>& potentialStorage() { return potential; }
  std::vector<std::complex<float>>& greensStorage() { return greensFunction; }
  std::vector<Vec3>& fieldStorage() { return field; }

 private:
  int wrapped(int i, int j, int k) const;
  int flat(int i, int j, int k) const { return i + j * gridPointsX + k * gridPointsX * gridPointsY; }
  int gridPointsX, gridPointsY, gridPointsZ, length;
  std::vector<Vec3> field;
  std::vector<std::complex<float>> density, densityFourier, potential, potentialFourier, greensFunction;
  FFTAdapter<float>& fftAdapter;
};

#if !P3M_B200_REFERENCE_TREE
// ---- include/stateRecorder.h (file formats of SURVEY section 8f row N2) ------------------------------
class StateRecorder {
 public:
  StateRecorder(int particlesCnt, int framesCnt, const std::filesystem::path& outputDirPath,
                const char* positionsFile = "positions.dat", const char* energyFile = "energy.txt",
                const char* momentumFile = "momentum.txt",
                const char* expectedMomentumFile = "expected_momentum.txt",
                const char* angularMomentumFile = "angular_momentum.txt", const char* fieldFile = "field.dat",
                int maxRecords = 500);
  ~StateRecorder();
  void recordPositions(std::vector<Vec3>::iterator begin, std::vector<Vec3>::iterator end);
  void recordPositions(const std::vector<Particle>& particles);
  void recordPositions(const float* xyz, std::size_t n);  // packed triples straight from the device copy
  void recordEnergy(float pe, float ke);
  void recordTotalMomentum(Vec3 momentum);
  void recordExpectedMomentum(Vec3 expectedMomentum);
  void recordTotalAngularMomentum(Vec3 angularMomentum);
  void recordField(const std::vector<Particle>& particles, float H, float DT);
  void recordField(const float* acc_xyz_original_units, std::size_t n);
  std::string flush();

 private:
  void writeVec(std::ofstream& f, Vec3 v);
  std::filesystem::path dir;
  std::ofstream positions, energy, momentum, expectedMomentum, angularMomentum, field;
  int particlesCnt, framesCnt;
};

#endif  // !P3M_B200_REFERENCE_TREE

// ---- include/simInfo.h -------------------------------------------------------------------------------
class SimInfo {
 public:
  // state-vector overloads (include/simInfo.h:11-24; used by the reference's direct-sum method)
  static float potentialEnergy(std::vector<Vec3>::iterator posBegin, std::vector<Vec3>::iterator posEnd,
                               const std::vector<float>& masses, float G);
  static float kineticEnergy(std::vector<Vec3>::iterator vBegin, std::vector<Vec3>::iterator vEnd,
                             const std::vector<float>& masses, float G);
  static Vec3 totalMomentum(std::vector<Vec3>::iterator vBegin, std::vector<Vec3>::iterator vEnd,
                            const std::vector<float>& masses);
  // mesh overloads (include/simInfo.h:26-36): from a host Grid, or from the density / potential vectors of
  // the CUDA-build PMMethodGPU (getGridDensity / getGridPotential), as source/p3mMethod.cpp:121-122 calls it
  static float potentialEnergy(const Grid& grid, const std::vector<Particle>& particles,
                               std::function<float(Vec3)> externalPotential, float H, float DT, float G);
  static float potentialEnergy(const std::vector<std::complex<float>>& gridDensity,
                               const std::vector<std::complex<float>>& gridPotential,
                               const std::vector<Particle>& particles, std::function<float(Vec3)> externalPotential,
                               float H, float DT, float G);
  static float kineticEnergy(const std::vector<Particle>& particles);
  static Vec3 totalMomentum(const std::vector<Particle>& particles);
  static Vec3 totalAngularMomentum(const std::vector<Particle>& particles);
  void setInitialMomentum(const std::vector<Particle>& particles);
  Vec3 updateExpectedMomentum(Vec3 externalForce, float DT);

 private:
  Vec3 expectedMomentum;
};

#if !P3M_B200_REFERENCE_TREE
// ---- source/externalFields.cpp ---------------------------------------------------------------------------
Vec3 sphRadDecrField(Vec3 pos, Vec3 center, float R, float M, float G);
float sphRadDecrFieldPotential(Vec3 pos, Vec3 center, float R, float M, float G);

#endif  // !P3M_B200_REFERENCE_TREE

// An external field the device can evaluate itself.  std::function fields cannot cross to the GPU: an EMPTY
// std::function means "no field"; any other callable is evaluated on the host once per particle per step
// (download positions, call, upload accelerations -- correct for every field, but slow).  Installing a
// descriptor (PMMethod::setExternalFieldDescriptor) moves the evaluation into the gather kernel; the
// descriptor then REPLACES the callable, so pass the matching one (none() for an identically zero lambda).
struct ExternalFieldDesc {
  enum Kind { NONE = 0, SPH_RAD_DECR = 1 } kind = NONE;
  Vec3 center{};
  float R = 0, M = 0;
  static ExternalFieldDesc none() { return ExternalFieldDesc{}; }
  static ExternalFieldDesc sphRadDecr(Vec3 center, float R, float M) { return ExternalFieldDesc{SPH_RAD_DECR, center, R, M}; }
};

// ---- include/pmMethod.h ------------------------------------------------------------------------------------
class PMMethod {
 public:
  // the reference's constructor (include/pmMethod.h:14-26)
  PMMethod(const std::vector<Vec3>& state, const std::vector<float>& masses,
           const std::tuple<float, float, float> effectiveBoxSize, const std::function<Vec3(Vec3)> externalField,
           const std::function<float(Vec3)> externalPotential, const float H, const float DT, const float G,
           const InterpolationScheme is, const FiniteDiffScheme fds, const GreensFunction gFunc,
           const float particleDiameter, Grid& grid);
  // the reference's CUDA-build constructor (include_gpu/PMMethodGPU.h:14-26): mesh size instead of a Grid
  PMMethod(const std::vector<Vec3>& state, const std::vector<float>& masses,
           const std::tuple<float, float, float> effectiveBoxSize, const std::function<Vec3(Vec3)> externalField,
           const std::function<float(Vec3)> externalPotential, const float H, const float DT, const float G,
           const InterpolationScheme is, const FiniteDiffScheme fds, const GreensFunction gFunc,
           const float particleDiameter, std::tuple<int, int, int> gridPoints);
  ~PMMethod();
  PMMethod(const PMMethod&) = delete;
  PMMethod& operator=(const PMMethod&) = delete;

  std::string run(StateRecorder& stateRecorder, const int simLength, bool collectDiagnostics = false,
                  bool recordField = false);

  std::vector<Particle>& getParticles();  // downloads; the next device call re-uploads the vector
  float getH() const { return H; }
  float getDT() const { return DT; }
  float getG() const { return G; }
  const Grid& getGrid() const;  // back-fills density and potential from the device (include/pmMethod.h:37)
  std::function<float(Vec3)> getExternalPotential() const { return externalPotential; }
  void pmMethodStep();
  bool escapedComputationalBox();
  Vec3 totalExternalForceOrigUnits();
  void initGreensFunction();

  // PMMethodGPU extras (include_gpu/PMMethodGPU.h:38-54)
  // the getters return the vectors as of the last copyGrid*ToHost(), like the reference's
  const std::vector<std::complex<float>>& getGridDensity() const;
  const std::vector<std::complex<float>>& getGridPotential() const;
  void copyParticlesDeviceToHost();
  void copyParticlesHostToDevice();
  void copyGridPotentialToHost();
  void copyGridDensityToHost();

  // new: device-evaluable external field, precision, access to the context
  void setExternalFieldDescriptor(const ExternalFieldDesc& d);
  void setPrecision(bool fp64);
  p3m_ctx* context();

 private:
  friend class P3MMethod;
  struct Impl;
  std::unique_ptr<Impl> impl;
  float H, DT, G;
  std::function<Vec3(Vec3)> externalField;
  std::function<float(Vec3)> externalPotential;
  void runLoop(StateRecorder& rec, int simLength, bool diagnostics, bool recordField, bool p3m);
};
using PMMethodGPU = PMMethod;  // the reference's `#ifdef CUDA` spelling

// ---- include/chainingMesh.h -----------------------------------------------------------------------------------
class ChainingMesh {
 public:
  struct LLNode {
    int particleId;
    LLNode* next;
    LLNode(int particleId, LLNode* next) : particleId(particleId), next(next) {}
    LLNode() = default;
  };
  ChainingMesh(std::tuple<float, float, float> compBoxSize, float cutoffRadius, float H, int N);
  void fillWithYSorting(const std::vector<Particle>& particles);
  void fill(const std::vector<Particle>& particles);
  std::array<int, 14> getNeighborsAndSelf(int cellIdx) const;
  LLNode* getParticlesInCell(int cellIdx) { return hoc[cellIdx]; }
  int getSize() const { return size; }
  std::tuple<int, int, int> getLength() const { return std::make_tuple(Mx, My, Mz); }

 private:
  int cellOf(const Particle& p) const;
  int Mx, My, Mz;
  float HCx, HCy, HCz;
  int size;
  std::vector<LLNode*> hoc;
  std::unique_ptr<LLNode[]> nodePool;
};

// ---- include/p3mMethod.h --------------------------------------------------------------------------------------
class P3MMethod {
 public:
  P3MMethod(PMMethod& pmMethod, std::tuple<float, float, float> compBoxSize, float cutoffRadius,
            float particleDiameter, float H, float softeningLength, CloudShape cloudShape,
            bool useSRForceTable = true, bool enableYSorting = true);
  void run(StateRecorder& stateRecorder, const int simLength, bool collectDiagnostics = false,
           bool recordField = false);
  // one P3M force evaluation on the current particles (pmMethodStep + short range + correction)
  void forceStep();

 private:
  PMMethod& pmMethod;
};

// ---- include/leapfrog.h, include/abstractStepper.h, include/unitConversions.h ----------------------------------
void setHalfStepVelocities(std::vector<Particle>& particles, float dt = 1.0f);
void setIntegerStepVelocities(std::vector<Particle>& particles, float dt = 1.0f);
void updateVelocities(std::vector<Particle>& particles, float dt = 1.0f);
void updatePositions(std::vector<Particle>& particles, float dt = 1.0f);

#if !P3M_B200_REFERENCE_TREE
template <typename T>
class AbstractStepper {
 public:
  virtual ~AbstractStepper() {}
  virtual void doStep(std::vector<T>& x, float dt) = 0;
};

#endif  // !P3M_B200_REFERENCE_TREE

// Kick-drift-kick leapfrog as an AbstractStepper over Particle (the reference only implements the
// interface for RK4; its leapfrog is the four free functions above).
class LeapfrogStepper : public AbstractStepper<Particle> {
 public:
  explicit LeapfrogStepper(std::function<void(std::vector<Particle>&)> computeAccelerations)
      : force(std::move(computeAccelerations)) {}
  void doStep(std::vector<Particle>& x, float dt) override;

 private:
  std::function<void(std::vector<Particle>&)> force;
};

// include/unitConversions.h:8-71 (all of it: the reference's stateRecorder.cpp, ppMethod.cpp and barnesHut.cpp
// include this header too)
inline Vec3 positionToCodeUntits(const Vec3& pos, float H) { return pos / H; }
inline Vec3 positionToOriginalUnits(const Vec3& pos, float H) { return H * pos; }
inline Vec3 velocityToCodeUntits(const Vec3& v, float H, float DT) { return DT * v / H; }
inline Vec3 velocityToOriginalUnits(const Vec3& v, float H, float DT) { return H * v / DT; }
inline Vec3 accelerationToCodeUnits(const Vec3& a, float H, float DT) { return DT * DT * a / H; }
inline Vec3 accelerationToOriginalUnits(const Vec3& a, float H, float DT) { return H * a / (DT * DT); }
inline constexpr float kP3mPiF = 3.14159265358979323846f;  // == std::numbers::pi_v<float>
inline float densityToCodeUnits(float density, float DT, float G) { return DT * DT * 4 * kP3mPiF * G * density; }
inline float densityToOriginalUnits(float density, float DT, float G) { return density / (DT * DT * 4 * kP3mPiF * G); }
inline float potentialToOriginalUnits(float potential, float H, float DT) { return potential * H * H / (DT * DT); }
inline float massToCodeUnits(float m, float H, float DT, float G) { return DT * DT * 4 * kP3mPiF * G / (H * H * H) * m; }
inline float massToOriginalUnits(float m, float H, float DT, float G) { return (H * H * H) / (DT * DT * 4 * kP3mPiF * G) * m; }
inline float lengthToCodeUnits(float x, float H) { return x / H; }
void stateToCodeUnits(std::vector<Vec3>& state, float H, float DT);
void stateToOriginalUnits(std::vector<Vec3>& state, float H, float DT);
void stateToCodeUnits(std::vector<Particle>& particles, float H, float DT);
void stateToOriginalUnits(std::vector<Particle>& particles, float H, float DT);
void velocitiesToCodeUnits(std::vector<Vec3>& velocities, float H, float DT);
void velocitiesToOriginalUnits(std::vector<Vec3>& velocities, float H, float DT);
void integerStepVelocitiesToOriginalUnits(std::vector<Particle>& particles, float H, float DT);
void integerStepVelocitiesToCodeUnits(std::vector<Particle>& particles, float H, float DT);
void massToCodeUnits(std::vector<Particle>& particles, float H, float DT, float G);
void massToOriginalUnits(std::vector<Particle>& particles, float H, float DT, float G);
