// demo_main.cpp -- a caller written against the REFERENCE API (PMMethod / P3MMethod / Grid / FFTAdapter /
// StateRecorder), in the style of the reference's demos (source/demos.cpp:729-777 galaxy-sim-pm,
// :897-951 galaxy-sim-p3m, :1416-1450 cluster-sim-p3m), compiled against the drop-in headers.  Initial
// conditions come from a file so that the same arrays can be fed to the reference.
//
//   demo_host <pm-disk|p3m-disk|p3m-plummer|fft-roundtrip> <ic.bin> <outdir> <simLength> [nx ny nz]
//   ic.bin: int32 n, then pos[3n], vel[3n], mass[n] as float32
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "p3mMethod.h"
#include "pmMethod.h"
#include "stateRecorder.h"

static bool readIC(const char* path, std::vector<Vec3>& state, std::vector<float>& masses) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  int n = 0;
  if (std::fread(&n, sizeof(int), 1, f) != 1) return false;
  state.resize(2 * (size_t)n);
  masses.resize(n);
  bool ok = std::fread(state.data(), sizeof(Vec3), n, f) == (size_t)n &&
            std::fread(state.data() + n, sizeof(Vec3), n, f) == (size_t)n &&
            std::fread(masses.data(), sizeof(float), n, f) == (size_t)n;
  std::fclose(f);
  return ok;
}

static int fftRoundTrip() {
  // test/fftAdaptersTest.cpp:6-22 with the GPU adapter
  int dims[3] = {2, 2, 2};
  CuFFTAdapter adapter(dims, 3);
  std::vector<std::complex<float>> in(8), mid(8), out(8);
  for (int i = 0; i < 8; ++i) in[i] = std::complex<float>(float(i + 1), float(8 - i));
  adapter.fft(in, mid);
  adapter.ifft(mid, out);
  double err = 0;
  for (int i = 0; i < 8; ++i) err = std::max(err, (double)std::abs(out[i] - in[i]));
  std::printf("fft round trip max error %.3e, sum mode %.1f (expect 36)\n", err, mid[0].real());
  return err < 1e-6 && std::abs(mid[0].real() - 36.0f) < 1e-4 ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc >= 2 && std::strcmp(argv[1], "fft-roundtrip") == 0) return fftRoundTrip();
  if (argc < 5) {
    std::fprintf(stderr, "usage: %s <pm-disk|p3m-disk|p3m-plummer> <ic.bin> <outdir> <simLength> [nx ny nz]\n", argv[0]);
    return 2;
  }
  const std::string mode = argv[1];
  std::vector<Vec3> state;
  std::vector<float> masses;
  if (!readIC(argv[2], state, masses)) {
    std::fprintf(stderr, "cannot read %s\n", argv[2]);
    return 2;
  }
  const int n = (int)masses.size();
  const int simLength = std::atoi(argv[4]);
  try {
    if (mode == "pm-disk" || mode == "p3m-disk") {
      // galaxy units: G(solar mass), kpc, Myr
      Vec3 galaxyCenter = Vec3::create(30, 30, 15);
      float rb = 3.0f, mb = 60.0f, G = 4.5e-3f;
      auto externalField = [=](Vec3 pos) -> Vec3 { return sphRadDecrField(pos, galaxyCenter, rb, mb, G); };
      auto externalPotential = [=](Vec3 pos) -> float { return sphRadDecrFieldPotential(pos, galaxyCenter, rb, mb, G); };
      auto gridPoints = argc >= 8 ? std::make_tuple(std::atoi(argv[5]), std::atoi(argv[6]), std::atoi(argv[7]))
                                  : std::make_tuple(128, 128, 64);
      std::array<int, 3> dims = {std::get<2>(gridPoints), std::get<1>(gridPoints), std::get<0>(gridPoints)};
      auto effectiveBoxSize = std::make_tuple(60.0f, 60.0f, 30.0f);
      float H = std::get<0>(effectiveBoxSize) / (std::get<0>(gridPoints) / 2);
      float DT = 1;
      CuFFTAdapter fftAdapter(dims);
      Grid grid(gridPoints, fftAdapter);
      StateRecorder stateRecorder(n, simLength + 1, argv[3]);
      if (mode == "pm-disk") {
        PMMethod pm(state, masses, effectiveBoxSize, externalField, externalPotential, H, DT, G,
                    InterpolationScheme::TSC, FiniteDiffScheme::TWO_POINT, GreensFunction::DISCRETE_LAPLACIAN, 0,
                    grid);
        if (!std::getenv("DEMO_HOST_CALLBACK"))  // default: the bulge field runs on the device
          pm.setExternalFieldDescriptor(ExternalFieldDesc::sphRadDecr(galaxyCenter, rb, mb));
        pm.run(stateRecorder, simLength, true /*diagnostics*/);
      } else {
        float a = 3 * H;
        PMMethod pm(state, masses, effectiveBoxSize, externalField, externalPotential, H, DT, G,
                    InterpolationScheme::TSC, FiniteDiffScheme::TWO_POINT, GreensFunction::S1_OPTIMAL, a, grid);
        pm.setExternalFieldDescriptor(ExternalFieldDesc::sphRadDecr(galaxyCenter, rb, mb));
        float re = 0.7f * a;
        float softeningLength = 1.5f;
        P3MMethod p3m(pm, effectiveBoxSize, re, a, H, softeningLength, CloudShape::S1);
        p3m.run(stateRecorder, simLength, true /*diagnostics*/);
      }
    } else if (mode == "p3m-plummer") {
      // cluster units: M(solar mass), pc, kyr
      float G = 4.5e-3f;
      auto externalField = [](Vec3) -> Vec3 { return Vec3::zero(); };
      auto externalPotential = [](Vec3) -> float { return 0; };
      auto gridPoints = argc >= 8 ? std::make_tuple(std::atoi(argv[5]), std::atoi(argv[6]), std::atoi(argv[7]))
                                  : std::make_tuple(128, 128, 128);
      auto effectiveBoxSize = std::make_tuple(60.0f, 60.0f, 60.0f);
      float H = std::get<0>(effectiveBoxSize) / (std::get<0>(gridPoints) / 2);
      float DT = 1;
      float particleDiam = 3 * H;
      PMMethod pm(state, masses, effectiveBoxSize, externalField, externalPotential, H, DT, G,
                  InterpolationScheme::TSC, FiniteDiffScheme::TWO_POINT, GreensFunction::S1_OPTIMAL, particleDiam,
                  gridPoints);
      // a non-empty callable is evaluated on the host every step; declare the zero field instead
      if (!std::getenv("DEMO_HOST_CALLBACK")) pm.setExternalFieldDescriptor(ExternalFieldDesc::none());
      float re = 0.7f * particleDiam;
      float softeningLength = 0.5f;
      P3MMethod p3m(pm, effectiveBoxSize, re, particleDiam, H, softeningLength, CloudShape::S1);
      StateRecorder stateRecorder(n, simLength + 1, argv[3]);
      p3m.run(stateRecorder, simLength, true /*diagnostics*/);
      // final state through the reference's accessor
      auto& ps = pm.getParticles();
      FILE* f = std::fopen((std::string(argv[3]) + "/final.bin").c_str(), "wb");
      for (const auto& p : ps) std::fwrite(&p.position, sizeof(Vec3), 1, f);
      for (const auto& p : ps) std::fwrite(&p.velocity, sizeof(Vec3), 1, f);
      std::fclose(f);
    } else {
      std::fprintf(stderr, "unknown mode %s\n", mode.c_str());
      return 2;
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  std::printf("\ndone\n");
  return 0;
}
